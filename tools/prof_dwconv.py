"""Depthwise conv sites of the EdgeLine-n engine graph in isolation (CUDA events, rotating inputs > L2).
    python tools/prof_dwconv.py            # current dispatch (k = 3: TMA-pipelined kernel)
    EL_DW_TMA=0 python tools/prof_dwconv.py  # round-1 cp.async tile kernel"""
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


rows, tot = [], 0.0
for (B, C, hw, k, count) in [(64, 32, 80, 3, 4), (64, 16, 160, 3, 1), (64, 64, 40, 3, 4), (64, 64, 20, 3, 8), (64, 64, 80, 3, 1), (64, 32, 40, 3, 4),
                              (64, 128, 40, 3, 1), (64, 256, 20, 3, 1), (64, 16, 160, 7, 1), (64, 32, 80, 7, 1), (64, 64, 40, 7, 1)]:
    gen = torch.Generator().manual_seed(C)
    x = torch.randn(B, C, hw, hw, generator=gen).to(DEV, torch.bfloat16).contiguous(memory_format=torch.channels_last)
    w = ops.pack_dw_weight((torch.randn(C, 1, k, k, generator=gen) * 0.3).to(DEV))
    nbytes = 2 * x.numel() * 2
    R = max(2, min(32, (600 << 20) // nbytes))
    sets = [x] + [x.clone() for _ in range(R - 1)]
    t = time_op(lambda s: ops.dwconv(s, w, k), sets)
    tot += t * count
    rows.append({"B": B, "C": C, "hw": hw, "k": k, "count_in_graph": count, "MB": nbytes / 1e6, "us": t * 1e6, "GBs": nbytes / t / 1e9, "frac": nbytes / t / 1e9 / PEAK})
print(json.dumps({"kernel": "cp.async tile kernel" if os.environ.get("EL_DW_TMA") == "0" else "TMA pipeline (k=3)", "sum_us_weighted": tot * 1e6, "sites": rows}, indent=1))
