"""k = 7 depthwise conv at the engine's sites: CUDA-event time per launch, inputs rotated through > L2.  EL_DW_TC=0 -> CUDA-core kernel.
   ONE=1: a single launch of the largest site (for ncu --set full)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from edge_yolo_b200 import ops  # noqa: E402

dev, dt, cl = "cuda", torch.bfloat16, torch.channels_last
g = torch.Generator(device=dev).manual_seed(0)
B = 64
K = int(os.environ.get("K", "7"))
sites = [(16, 160), (32, 80), (64, 40), (64, 20)] if K == 7 else [(16, 160), (64, 80), (128, 40), (256, 20)]
if os.environ.get("ONE"):
    sites = sites[:1]
for C, hw in sites:
    R = 1 if os.environ.get("ONE") else max(2, min(16, (600 << 20) // (2 * B * C * hw * hw * 2)))
    xs = [torch.randn(B, C, hw, hw, device=dev, generator=g).to(dt).contiguous(memory_format=cl) for _ in range(R)]
    outs = [torch.empty_like(x) for x in xs]
    wp = ops.pack_dw_weight(torch.randn(C, 1, K, K, device=dev, generator=g) * 0.2)
    for x, o in zip(xs, outs):
        ops.dwconv(x, wp, K, out=o)
    if os.environ.get("ONE"):
        continue
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda._sleep(2_000_000)
        e0.record()
        for x, o in zip(xs, outs):
            ops.dwconv(x, wp, K, out=o)
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3 / R)
    ts.sort()
    nbytes = 2 * B * C * hw * hw * 2
    print(f"k={K} C={C:3d} {hw}x{hw}: {ts[2]:7.2f} us  {nbytes / ts[2] / 1e3:6.0f} GB/s   (EL_DW_TC={os.environ.get('EL_DW_TC', '1')})", flush=True)
torch.cuda.synchronize()
