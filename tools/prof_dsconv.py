"""Depthwise 3x3 -> pointwise 1x1 sites of the EdgeLine-n engine graph: the fused kernel (el_dsconv3_fwd) against the two-kernel path
(el_dwconv_fwd -> el_pwconv_fwd) in isolation (CUDA events, rotating inputs > L2, launches back to back as in the graph).
    python tools/prof_dsconv.py"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from edge_yolo_b200 import ops  # noqa: E402

DEV, PEAK = "cuda", 6544.0


def time_op(fn, sets, iters=10):
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


# (B, C, N, map, depthwise epilogue, count in the n graph)
SITES = [(64, 16, 16, 160, False, 1), (64, 32, 32, 80, False, 2), (64, 32, 32, 40, False, 2), (64, 64, 64, 40, False, 2), (64, 64, 64, 20, False, 4),
         (64, 64, 80, 80, True, 1), (64, 80, 80, 80, True, 1), (64, 128, 80, 40, True, 1), (64, 80, 80, 40, True, 1), (64, 256, 80, 20, True, 1),
         (64, 80, 80, 20, True, 1)]
if len(sys.argv) > 1:
    SITES = [s for s in SITES if str(s[1]) in sys.argv[1].split(",") and (len(sys.argv) < 3 or str(s[3]) in sys.argv[2].split(","))]
rows, tot_f, tot_2 = [], 0.0, 0.0
for (B, C, N, hw, epi, count) in SITES:
    gen = torch.Generator().manual_seed(C + N)
    x = torch.randn(B, C, hw, hw, generator=gen).to(DEV, torch.bfloat16).contiguous(memory_format=torch.channels_last)
    wd = ops.pack_dw_weight((torch.randn(C, 1, 3, 3, generator=gen) * 0.3).to(DEV))
    bd = torch.randn(C, generator=gen).to(DEV) if epi else None
    wp = (torch.randn(N, C, generator=gen) * C ** -0.5).to(DEV)
    bp = torch.randn(N, generator=gen).to(DEV)
    wf, w2 = ops.pack_dsconv3_weight(wp), ops.pack_pw_weight(wp, [C], torch.bfloat16, M=B * hw * hw)
    dact = ops.ACT_SILU if epi else ops.ACT_NONE
    nbytes = x.numel() * 2 + B * N * hw * hw * 2
    R = max(2, min(32, (600 << 20) // nbytes))
    sets = [x] + [x.clone() for _ in range(R - 1)]
    outs = [torch.empty((B, N, hw, hw), device=DEV, dtype=torch.bfloat16, memory_format=torch.channels_last) for _ in range(R)]
    mids = [torch.empty_like(x) for _ in range(R)]
    idx = {id(s): i for i, s in enumerate(sets)}
    t_f = time_op(lambda s: ops.dsconv3(s, wd, wf, N, bias=bp, act=1, dw_bias=bd, dw_act=dact, out=outs[idx[id(s)]]), sets)
    t_2 = time_op(lambda s: ops.pwconv([ops.dwconv(s, wd, 3, bias=bd, act=dact, out=mids[idx[id(s)]])], w2, N, bias=bp, act=1, out=outs[idx[id(s)]]), sets)
    tot_f += t_f * count
    tot_2 += t_2 * count
    rows.append({"B": B, "C": C, "N": N, "hw": hw, "dw_epilogue": epi, "count_in_graph": count, "MB": round(nbytes / 1e6, 2), "fused_us": round(t_f * 1e6, 2),
                 "two_kernel_us": round(t_2 * 1e6, 2), "fused_GBs": round(nbytes / t_f / 1e9, 1), "fused_frac": round(nbytes / t_f / 1e9 / PEAK, 3)})
    print(json.dumps(rows[-1]), flush=True)
print(json.dumps({"sum_fused_us_weighted": round(tot_f * 1e6, 1), "sum_two_kernel_us_weighted": round(tot_2 * 1e6, 1)}))
