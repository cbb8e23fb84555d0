"""Determinism and view-invariance check of the k = 7 depthwise kernels: 50 repeated launches must be bit-identical, and channel-slice
input / output views must give the same bits as dense tensors (used while hunting the ld.global.nc / programmatic-dependent-launch hazard,
DESIGN.md section 5).  python tools/dbg_dw7.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from edge_yolo_b200 import ops
dev, dt, cl = "cuda", torch.bfloat16, torch.channels_last
g = torch.Generator(device=dev).manual_seed(0)
for (B, C, hw) in ((2, 16, 80), (2, 32, 40), (64, 16, 160)):
    x = torch.randn(B, C, hw, hw, device=dev, generator=g).to(dt).contiguous(memory_format=cl)
    wp = ops.pack_dw_weight(torch.randn(C, 1, 7, 7, device=dev, generator=g) * 0.2)
    ref = ops.dwconv(x, wp, 7).clone()
    bad = 0
    for i in range(50):
        o = ops.dwconv(x, wp, 7)
        if not torch.equal(o, ref):
            bad += 1
            d = (o.float() - ref.float()).abs()
            idx = (d > 0).nonzero()
            print("  mismatch run", i, "count", idx.shape[0], "first", idx[0].tolist(), "max", float(d.max()))
    print((B, C, hw), "nondeterministic runs:", bad)
# context variations: channel-slice input / output views vs dense tensors must give the same bits
for (B, C, hw) in ((2, 16, 80), (2, 32, 40)):
    x = torch.randn(B, C, hw, hw, device=dev, generator=g).to(dt).contiguous(memory_format=cl)
    wp = ops.pack_dw_weight(torch.randn(C, 1, 7, 7, device=dev, generator=g) * 0.2)
    bias = torch.randn(C, device=dev, generator=g)
    ref = ops.dwconv(x, wp, 7, bias=bias, act=1)
    big = torch.zeros(B, 3 * C, hw, hw, device=dev, dtype=dt).contiguous(memory_format=cl)
    big[:, C:2 * C] = x
    o1 = ops.dwconv(big[:, C:2 * C], wp, 7, bias=bias, act=1)
    obig = torch.zeros(B, 2 * C, hw, hw, device=dev, dtype=dt).contiguous(memory_format=cl)
    ops.dwconv(x, wp, 7, bias=bias, act=1, out=obig[:, C:])
    print((B, C, hw), "slice-in equal:", torch.equal(o1, ref), " slice-out equal:", torch.equal(obig[:, C:], ref))
