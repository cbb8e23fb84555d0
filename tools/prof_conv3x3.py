"""Wide dense 3x3 convs of the engine graph: el_conv3x3_halo_fwd against cuDNN (+ el_bias_act) and the tap-shifted el_conv3x3_fwd, CUDA-event
times with rotating inputs > L2.   python tools/prof_conv3x3.py"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from edge_yolo_b200 import ops  # noqa: E402

DEV, PEAK = "cuda", 6544.0
torch.backends.cudnn.benchmark = True


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
for name, B, C, N, hw in [("head cv2 P3 64->64 @80", 64, 64, 64, 80), ("head cv2 P4 128->64 @40", 64, 128, 64, 40), ("head cv2 P4 64->64 @40", 64, 64, 64, 40),
                          ("head cv2 P5 64->64 @20", 64, 64, 64, 20), ("f_h c=64 64->32 @20 (3B)", 192, 64, 32, 20), ("f_h c=128 128->64 @10 (3B)", 192, 128, 64, 10),
                          ("s-scale 128->128?", 64, 64, 128, 40)]:
    gen = torch.Generator().manual_seed(1)
    x = torch.randn(B, C, hw, hw, generator=gen).to(DEV, torch.bfloat16).contiguous(memory_format=torch.channels_last)
    w = (torch.randn(N, C, 3, 3, generator=gen) * (9 * C) ** -0.5).to(DEV)
    bias = torch.randn(N, generator=gen).to(DEV)
    wb = w.to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    nbytes = (x.numel() + B * N * hw * hw) * 2
    R = max(2, min(32, (600 << 20) // nbytes))
    sets = [x] + [x.clone() for _ in range(R - 1)]
    wpk = ops.pack_conv3x3_halo_weight(w)
    t_halo = time_op(lambda s: ops.conv3x3_halo(s, wpk, N, bias=bias, act=1), sets)
    t_cudnn = time_op(lambda s: ops.bias_act(F.conv2d(s, wb, padding=1), bias, 1), sets)
    err = float((ops.conv3x3_halo(x, wpk, N, bias=bias, act=1).float() - F.silu(F.conv2d(x.float(), wb.float(), bias, padding=1))).abs().max())
    rows.append({"site": name, "B": B, "C": C, "N": N, "hw": hw, "MB": nbytes / 1e6, "halo_us": t_halo * 1e6, "cudnn_bias_act_us": t_cudnn * 1e6,
                 "halo_GBs": nbytes / t_halo / 1e9, "halo_frac": nbytes / t_halo / 1e9 / PEAK, "max_abs_err": err})
print(json.dumps(rows, indent=1))
