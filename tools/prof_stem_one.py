"""One launch of the fused uint8 stem (and a few depthwise sites) for `ncu --set full -k regex:stem|dwconv`."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from edge_yolo_b200 import ops  # noqa: E402

dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
img = torch.randint(0, 256, (64, 640, 640, 3), dtype=torch.uint8, device=dev, generator=g)
w = torch.randn(16, 3, 3, 3, device=dev, generator=g) * 0.1 / 255
b = torch.randn(16, device=dev, generator=g)
for _ in range(2):
    out = ops.stem_conv_u8(img, w, b)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    ops.stem_conv_u8(img, w, b, out=out)
e1.record()
e1.synchronize()
print("stem us", e0.elapsed_time(e1) * 1e3 / 5)
