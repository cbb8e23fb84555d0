"""One launch of el_pwconv_fwd per representative EdgeLine-n shape, for `ncu --set full -k regex:pwconv`."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from edge_yolo_b200 import ops  # noqa: E402

dev, dt, cl = "cuda", torch.bfloat16, torch.channels_last
g = torch.Generator(device=dev).manual_seed(0)
B = 64
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for K, N, hw in ((64, 64, 160), (128, 128, 80), (16, 8, 160), (32, 32, 160)):
    x = torch.randn(B, K, hw, hw, device=dev, generator=g).to(dt).contiguous(memory_format=cl)
    w = torch.randn(N, K, device=dev, generator=g) * (K ** -0.5)
    bias = torch.randn(N, device=dev, generator=g)
    wpk = ops.pack_pw_weight(w, [K], dt, B * hw * hw)
    out = torch.empty(B, N, hw, hw, device=dev, dtype=dt).contiguous(memory_format=cl)
    flush.zero_()
    ops.pwconv([x], wpk, N, bias=bias, act=ops.ACT_SILU, out=out)
torch.cuda.synchronize()
print("ok")
