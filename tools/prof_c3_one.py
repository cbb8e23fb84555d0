"""One launch of el_conv3x3_fwd at a wide-channel site for `ncu --set full -k regex:pwconv`."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from edge_yolo_b200 import ops  # noqa: E402

dev, dt, cl = "cuda", torch.bfloat16, torch.channels_last
g = torch.Generator(device=dev).manual_seed(0)
B = 64
for C, N, hw, s in ((64, 64, 80, 1), (128, 64, 40, 1)):
    x = torch.randn(B, C, hw, hw, device=dev, generator=g).to(dt).contiguous(memory_format=cl)
    w = torch.randn(N, C, 3, 3, device=dev, generator=g) * (9 * C) ** -0.5
    bias = torch.randn(N, device=dev, generator=g)
    wpk = ops.pack_conv3x3_weight(w, dt, B * hw * hw)
    for _ in range(2):
        out = ops.conv3x3(x, wpk, N, bias=bias, act=ops.ACT_SILU, stride=s)
torch.cuda.synchronize()
print("ok")
