"""Fused detect chain (decode + candidate emit | select + sort | sweep) of the engine in isolation, in bench.py's regimes.
    python tools/prof_detect.py                          # configs[1]: n, batch 64, 640, predict thresholds, random-init weights
    python tools/prof_detect.py --stress                 # configs[2]: s, batch 32, 1280, nc 10, conf .001, multi_label
    python tools/prof_detect.py --trained                # configs[1] with the committed synthetic-task checkpoint (many classes, few candidates)
Times by CUDA events around back-to-back launches behind a spin kernel (stage 1 on rotating copies of the head maps > L2)."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from edge_yolo_b200 import ops  # noqa: E402
from edge_yolo_b200.engine import Predictor, build_model  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--stress", action="store_true")
ap.add_argument("--trained", action="store_true")
ap.add_argument("--iters", type=int, default=8)
a = ap.parse_args()
dev = torch.device("cuda:0")
if a.stress:
    scale, nc, B, S, nms = "s", 10, 32, 1280, dict(conf=0.001, iou=0.7, max_det=300, multi_label=True)
else:
    scale, nc, B, S, nms = "n", 80, 64, 640, dict(conf=0.25, iou=0.7, max_det=300, multi_label=False)
if a.trained:
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import synth_data  # noqa: E402
    from edge_yolo_b200.model import EdgeLineYOLO  # noqa: E402

    nc = synth_data.NC
    state = {k: (v.float() if v.is_floating_point() else v)
             for k, v in torch.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "edgeline_n_synth.pt"), map_location="cpu").items()}
    model = EdgeLineYOLO("n", nc).eval()
    model.load_state_dict(state, strict=True)
    model = model.fuse(engine=True).to(device=dev, dtype=torch.bfloat16).to(memory_format=torch.channels_last)
else:
    model = build_model(scale, nc, seed=0, device=dev)
pred = Predictor(model, B, S, use_graph=False, **nms)
gen = torch.Generator().manual_seed(1234)
if a.trained:
    x, _ = synth_data.synth_batch(B, S, torch.Generator().manual_seed(1000), "cpu")
    u8 = (x * 255).round().to(torch.uint8).permute(0, 2, 3, 1).contiguous()
else:
    u8 = torch.randint(0, 256, (B, S, S, 3), dtype=torch.uint8, generator=gen)
pred.predict_u8(u8.pin_memory())

calls = []
saved = ops.gfl_detect
ops.gfl_detect = lambda *args, **kw: (calls.append((args, kw)), saved(*args, **kw))[1]
with torch.no_grad():
    pred._forward(False)
ops.gfl_detect = saved
args, kw = calls[0]
kw = dict(kw, workspace=torch.empty(1 << 30, dtype=torch.uint8, device=dev))
out, cnt = saved(*args, **dict(kw, stages=7))
torch.cuda.synchronize()
Bn = args[0][0].shape[0]
n_cand = kw["workspace"][: 4 * Bn].view(torch.int32).clamp(max=kw.get("max_nms", 30000))
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(stages, R=4):
    ts = []
    for _ in range(a.iters):
        saved(*args, **dict(kw, stages=7))
        flush.zero_()
        torch.cuda._sleep(4_000_000)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(R):
            saved(*args, **dict(kw, stages=stages))
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3 / R)
    ts.sort()
    return ts[len(ts) // 2]


res = {"regime": "stress" if a.stress else ("trained" if a.trained else "configs[1]"), "B": Bn, "candidates_per_image": float(n_cand.float().mean()),
       "candidates_max": int(n_cand.max()), "kept_per_image": float(cnt.float().mean()),
       "emit_us": timed(1), "sort_us": timed(2), "sweep_us": timed(4), "chain_us": timed(7)}
print(json.dumps(res))
