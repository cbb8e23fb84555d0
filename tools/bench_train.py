"""Training step of BASELINE.json configs[4] through the product's training path: EdgeLine-YOLO-s, batch 64 per GPU, synthetic 640^2,
8 boxes per image, forward (CUDA DWT / merge / gated residual / attention kernels under autograd) + v8DetectionLoss (DFL kernel, TAL)
+ backward (CUDA backward kernels) + SGD step; DDP (NCCL all-reduce of the gradients) when launched under torchrun.

    python tools/bench_train.py [--scale s --batch 64 --imgsz 640 --steps 10 --warmup 3 --amp]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29520 tools/bench_train.py

Prints one JSON line (rank 0): images/s over all ranks (device-timed, max over ranks) and the split forward / loss / backward+step."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", default="s")
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--imgsz", type=int, default=640)
    ap.add_argument("--nc", type=int, default=80)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--amp", action="store_true", help="bf16 autocast for the convolutions (the custom ops follow the tensor dtype)")
    a = ap.parse_args()

    from edge_yolo_b200 import dist as eld
    from edge_yolo_b200 import modules as M
    from edge_yolo_b200.detection_loss import v8DetectionLoss
    from edge_yolo_b200.model import EdgeLineYOLO

    rank, world, local = eld.env_rank()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    eld.init(dev)
    torch.manual_seed(0)
    model = EdgeLineYOLO(a.scale, a.nc)
    with torch.no_grad():
        for m in model.modules():
            if isinstance(m, M._WaveletEnhancer):
                m.gamma.fill_(0.5)
    model.model[-1].bias_init()
    model = model.to(dev).train()
    crit = v8DetectionLoss(model)
    net = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local], find_unused_parameters=True) if world > 1 else model
    opt = torch.optim.SGD(model.parameters(), lr=0.01, momentum=0.937, nesterov=True, weight_decay=5e-4)

    g = torch.Generator(device=dev).manual_seed(1 + rank)
    x = torch.rand(a.batch, 3, a.imgsz, a.imgsz, device=dev, generator=g)
    nb = 8  # boxes per image
    cls = torch.randint(0, a.nc, (a.batch * nb, 1), device=dev, generator=g).float()
    cxy = 0.1 + 0.8 * torch.rand(a.batch * nb, 2, device=dev, generator=g)
    wh = 0.05 + 0.35 * torch.rand(a.batch * nb, 2, device=dev, generator=g)
    batch = {"batch_idx": torch.arange(a.batch, device=dev).repeat_interleave(nb).float(), "cls": cls, "bboxes": torch.cat([cxy, wh], 1)}

    ev = lambda: torch.cuda.Event(enable_timing=True)
    t_f = t_l = t_b = 0.0

    def step(timed):
        nonlocal t_f, t_l, t_b
        e0, e1, e2, e3 = ev(), ev(), ev(), ev()
        e0.record()
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=a.amp):
            feats = net(x)
        e1.record()
        loss, items = crit([f.float() for f in feats], batch)
        e2.record()
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
        e3.record()
        if timed:
            e3.synchronize()
            t_f += e0.elapsed_time(e1); t_l += e1.elapsed_time(e2); t_b += e2.elapsed_time(e3)
        return items

    for _ in range(a.warmup):
        step(False)
    eld.barrier(dev)
    s0, s1 = ev(), ev()
    s0.record()
    for _ in range(a.steps):
        items = step(True)
    s1.record()
    s1.synchronize()
    eld.barrier(dev)
    ms = eld.max_over_ranks(s0.elapsed_time(s1), dev)
    if rank == 0:
        print(json.dumps({"metric": f"training images/sec (EdgeLine-YOLO-{a.scale}, {a.imgsz}x{a.imgsz}, batch {a.batch}/GPU, {'bf16 autocast' if a.amp else 'fp32'})",
                          "value": world * a.batch * a.steps / (ms * 1e-3), "unit": "images/s", "n_gpus": world, "steps": a.steps, "ms_per_step": ms / a.steps,
                          "forward_ms": t_f / a.steps, "loss_ms": t_l / a.steps, "backward_step_ms": t_b / a.steps,
                          "loss_items": [float(v) for v in items.tolist()], "peak_mem_GB": torch.cuda.max_memory_allocated() / 1e9}), flush=True)
    eld.shutdown()


if __name__ == "__main__":
    main()
