"""Trains EdgeLine-YOLO-n on the synthetic shapes task (tools/synth_data.py) THROUGH THE PRODUCT'S TRAINING PATH on one B200:
CUDA forwards + backwards of the DWT / merge / gated residual / linear attention kernels, `v8DetectionLoss` on the DFL kernel
(+ TaskAlignedAssigner), AdamW.  Writes the checkpoint the mAP parity test loads:

    gpurun -- python tools/train_synth.py --seconds 150 --out gpurun_out/edgeline_n_synth.pt
    cp gpurun_out/edgeline_n_synth.pt tests/golden/

The state dict is stored in fp16 (5 MB); every arm of the parity test loads the same rounded weights."""
import argparse
import math
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from tools import synth_data  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=150)
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--imgsz", type=int, default=256)
    ap.add_argument("--lr", type=float, default=2e-3)
    ap.add_argument("--out", default="gpurun_out/edgeline_n_synth.pt")
    a = ap.parse_args()

    from edge_yolo_b200 import modules as M
    from edge_yolo_b200.detection_loss import v8DetectionLoss
    from edge_yolo_b200.model import EdgeLineYOLO

    dev = torch.device("cuda")
    torch.manual_seed(0)
    model = EdgeLineYOLO("n", synth_data.NC)
    with torch.no_grad():
        for m in model.modules():
            if isinstance(m, M._WaveletEnhancer):
                m.gamma.fill_(0.5)
    model.model[-1].bias_init()
    model = model.to(dev).train()
    crit = v8DetectionLoss(model)
    decay = [p for n, p in model.named_parameters() if p.ndim > 1]
    no_decay = [p for n, p in model.named_parameters() if p.ndim <= 1]
    opt = torch.optim.AdamW([{"params": decay, "weight_decay": 5e-4}, {"params": no_decay, "weight_decay": 0.0}], lr=a.lr, betas=(0.9, 0.999))
    gen = torch.Generator(device=dev).manual_seed(1)
    t0, it, warm = time.time(), 0, 100
    ema = None
    while True:
        el = time.time() - t0
        if el > a.seconds:
            break
        frac = el / a.seconds
        lr = a.lr * min(1.0, (it + 1) / warm) * (0.02 + 0.98 * 0.5 * (1 + math.cos(math.pi * frac)))
        for g in opt.param_groups:
            g["lr"] = lr
        x, t = synth_data.synth_batch(a.batch, a.imgsz, gen, dev)
        feats = model(x)
        loss, items = crit(feats, {"batch_idx": t[:, 0], "cls": t[:, 1], "bboxes": t[:, 2:]})
        opt.zero_grad(set_to_none=True)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(model.parameters(), 10.0)
        opt.step()
        it += 1
        if it % 50 == 0:
            li = items.tolist()
            ema = li if ema is None else [0.9 * e + 0.1 * v for e, v in zip(ema, li)]
            print(f"it {it:5d}  {el:6.1f}s  lr {lr:.2e}  box {li[0]:.3f} cls {li[1]:.3f} dfl {li[2]:.3f}", flush=True)
    torch.cuda.synchronize()
    print(f"{it} iterations, {it * a.batch / (time.time() - t0):.0f} img/s", flush=True)

    # quick self-check on a held-out batch: predict-mode detections vs labels
    model.eval()
    from edge_yolo_b200.nms import non_max_suppression
    from oracle import metrics_ref  # tools/ is measurement + test infrastructure, not the product

    g2 = torch.Generator().manual_seed(4242)
    xv, tv = synth_data.synth_batch(32, a.imgsz, g2, "cpu")
    with torch.no_grad():
        y, _ = model(xv.to(dev))
        dets = non_max_suppression(y, conf_thres=0.001, iou_thres=0.7, max_det=300, multi_label=True)
    m, m50 = metrics_ref.evaluate([d.cpu().numpy() for d in dets], synth_data.labels_xyxy(tv, 32, a.imgsz))
    print(f"held-out mAP50-95 {100 * m:.2f}  mAP50 {100 * m50:.2f}", flush=True)
    os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
    torch.save({k: (v.half() if v.is_floating_point() else v).cpu() for k, v in model.state_dict().items()}, a.out)
    print("saved", a.out, os.path.getsize(a.out))


if __name__ == "__main__":
    main()
