"""Training step of BASELINE.json configs[4] through the product's training path (edge_yolo_b200.train.TrainStep): EdgeLine-YOLO-s, batch 64
per GPU, synthetic 640^2, 8 boxes per image; DDP (NCCL all-reduce of the gradients) when launched under torchrun.  bench.py runs the same leg
as `extra_configs["configs[4]"]`; this script exposes the variants (fp32 / bf16 autocast, NCHW / NHWC) for A/B measurement.

    python tools/bench_train.py [--scale s --batch 64 --imgsz 640 --steps 10 --warmup 3 --no-amp --nchw]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29520 tools/bench_train.py
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", default="s")
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--imgsz", type=int, default=640)
    ap.add_argument("--nc", type=int, default=80)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--no-amp", action="store_true", help="fp32 instead of bf16 autocast")
    ap.add_argument("--nchw", action="store_true", help="contiguous NCHW activations instead of channels_last")
    a = ap.parse_args()

    import bench
    from edge_yolo_b200 import dist as eld

    rank, world, local = eld.env_rank()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    eld.init(dev)
    torch.backends.cudnn.benchmark = True
    res = bench.train_leg({"dev": dev, "rank": rank, "world": world, "local": local}, a.scale, a.batch, a.imgsz, a.nc, a.steps, a.warmup,
                          amp=not a.no_amp, channels_last=not a.nchw)
    if rank == 0:
        print(json.dumps(dict(res, n_gpus=world)), flush=True)
    eld.shutdown()


if __name__ == "__main__":
    main()
