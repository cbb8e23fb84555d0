"""Sub-band merge (bands-only engine form) at the eight sites of EdgeLine-n / -s in isolation (CUDA events, rotating inputs > L2).
    python tools/prof_merge.py [n|s]              # shared-memory staged kernel (default dispatch)
    EL_MERGE_SMEM=0 python tools/prof_merge.py    # round-1 register kernel (per-thread L1-cached loads)
    EL_MERGE_SR=4 python tools/prof_merge.py      # fixed source-row chunk"""
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


scale = sys.argv[1] if len(sys.argv) > 1 else "n"
mul = {"n": 1, "s": 2}[scale]
B = 64
alpha = torch.tensor([0.5, 0.2, 0.2, 0.1], device=DEV)
rows, tot, tot_b = [], 0.0, 0
for (c, hw, count) in [(16 * mul, 160, 1), (32 * mul, 80, 2), (64 * mul, 40, 3), (128 * mul, 20, 2)]:
    gen = torch.Generator().manual_seed(c)
    mk = lambda: [torch.randn(B, c // 2, hw // 2, hw // 2, generator=gen).to(DEV, torch.bfloat16).contiguous(memory_format=torch.channels_last) for _ in range(4)]
    nbytes = int(2.5 * B * c * hw * hw * 2)
    R = max(2, min(32, (600 << 20) // nbytes))
    sets = [mk() for _ in range(R)]
    t = time_op(lambda s: ops.wave_merge_bands(*s, alpha, hw, hw), sets)
    tot += t * count
    tot_b += nbytes * count
    rows.append({"B": B, "c": c, "hw": hw, "count_in_graph": count, "MB": nbytes / 1e6, "us": t * 1e6, "GBs": nbytes / t / 1e9, "frac": nbytes / t / 1e9 / PEAK})
print(json.dumps({"kernel": "merge_fwd_x2 (registers)" if os.environ.get("EL_MERGE_SMEM") == "0" else "merge_fwd_x2s (smem staged)",
                  "SR": os.environ.get("EL_MERGE_SR", "auto"), "scale": scale, "sum_us_weighted": tot * 1e6,
                  "all_sites_GBs": tot_b / tot / 1e9, "all_sites_frac": tot_b / tot / 1e9 / PEAK, "sites": rows}, indent=1))
