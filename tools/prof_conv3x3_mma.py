"""Narrow 3x3 convs (the shared f_h of the early enhancers) on el_conv3x3_mma_fwd: CUDA-event times at the EdgeLine-n / s sites, rotating inputs > L2.
    python tools/prof_conv3x3_mma.py"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from edge_yolo_b200 import ops  # noqa: E402

DEV, PEAK = "cuda", 6544.0


def time_op(fn, sets, iters=12):
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
    ts = []
    for _ in range(iters):
        flush.zero_()
        torch.cuda._sleep(3_000_000)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for s in sets:
            fn(s)
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e-3 / len(sets))
    ts.sort()
    return ts[len(ts) // 2]


rows = []
for B, C, N, hw in [(192, 16, 8, 80), (192, 32, 16, 40), (192, 32, 16, 80), (192, 16, 8, 160)]:
    gen = torch.Generator().manual_seed(1)
    x = torch.randn(B, C, hw, hw, generator=gen).to(DEV, torch.bfloat16).contiguous(memory_format=torch.channels_last)
    w = (torch.randn(N, C, 3, 3, generator=gen) * (9 * C) ** -0.5).to(DEV)
    bias = torch.randn(N, generator=gen).to(DEV)
    nbytes = (x.numel() + B * N * hw * hw) * 2
    R = max(2, min(32, (600 << 20) // nbytes))
    sets = [x] + [x.clone() for _ in range(R - 1)]
    t = time_op(lambda s: ops.conv3x3_mma(s, w, bias=bias, act=1), sets)
    rows.append({"B": B, "C": C, "N": N, "hw": hw, "MB": round(nbytes / 1e6, 2), "us": round(t * 1e6, 2), "GBs": round(nbytes / t / 1e9, 1), "frac": round(nbytes / t / 1e9 / PEAK, 3)})
print(json.dumps(rows, indent=1))
