"""A few launches of el_conv3x3_halo_fwd at the largest site, for `ncu --set full -k regex:conv3x3_halo`."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from edge_yolo_b200 import ops  # noqa: E402

B, C, N, hw = 64, 64, 64, 80
gen = torch.Generator().manual_seed(1)
x = torch.randn(B, C, hw, hw, generator=gen).to("cuda", torch.bfloat16).contiguous(memory_format=torch.channels_last)
w = (torch.randn(N, C, 3, 3, generator=gen) * (9 * C) ** -0.5).to("cuda")
bias = torch.randn(N, generator=gen).to("cuda")
wpk = ops.pack_conv3x3_halo_weight(w)
for _ in range(3):
    ops.conv3x3_halo(x, wpk, N, bias=bias, act=1)
torch.cuda.synchronize()
