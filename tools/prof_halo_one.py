"""A few launches of el_conv3x3_halo_fwd at 64 -> 64 @ 80 x 80, batch 64, for `ncu --set full -k regex:conv3x3_halo_kernel -s 1 -c 1`."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from edge_yolo_b200 import ops  # noqa: E402

dev, dt, cl = "cuda", torch.bfloat16, torch.channels_last
g = torch.Generator(device=dev).manual_seed(0)
B, C, N, hw = 64, int(os.environ.get("C", 64)), int(os.environ.get("N", 64)), int(os.environ.get("HW", 80))
x = torch.randn(B, C, hw, hw, device=dev, generator=g).to(dt).contiguous(memory_format=cl)
w = torch.randn(N, C, 3, 3, device=dev, generator=g) * (9 * C) ** -0.5
bias = torch.randn(N, device=dev, generator=g)
wpk = ops.pack_conv3x3_halo_weight(w)
for _ in range(3):
    out = ops.conv3x3_halo(x, wpk, N, bias=bias, act=ops.ACT_SILU)
torch.cuda.synchronize()
print("ok")
