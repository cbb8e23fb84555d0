"""Eager (no CUDA graph) forward passes of the benchmark workload, for ncu:
   ncu --set full -k regex:'dwt|merge|gated|decode|nms|sort|linattn' -s <n> -c <n> python tools/prof_forward.py --passes 2
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from edge_yolo_b200.engine import Predictor, build_model  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--passes", type=int, default=2)
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--imgsz", type=int, default=640)
ap.add_argument("--scale", default="n")
a = ap.parse_args()
model = build_model(a.scale, 80, seed=0)
pred = Predictor(model, a.batch, a.imgsz, use_graph=False)
pred.u8.copy_(torch.randint(0, 256, tuple(pred.u8.shape), dtype=torch.uint8, generator=torch.Generator().manual_seed(1)))
for _ in range(a.passes):
    out, cnt = pred._forward(True)
torch.cuda.synchronize()
print("kept", int(cnt.sum()))
