"""Times the engine's convolution kernels (el_dwconv_fwd, el_pwconv_fwd) at the EdgeLine-n shapes against the PyTorch /
cuDNN path they replace (bf16 NHWC, B=64).  CUDA events, rotating buffers larger than L2.
   python tools/prof_conv.py [dw] [pw]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from edge_yolo_b200 import ops  # noqa: E402

dev, dt, cl = "cuda", torch.bfloat16, torch.channels_last
B = int(os.environ.get("B", "64"))
g = torch.Generator(device=dev).manual_seed(0)
rn = lambda *s: torch.randn(*s, device=dev, generator=g).to(dt).contiguous(memory_format=cl)


def timeit(fn, sets, iters=5):
    for a in sets:
        fn(*a)
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        torch.cuda._sleep(2_000_000)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for a in sets:
            fn(*a)
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3 / len(sets))
    ts.sort()
    return ts[len(ts) // 2]


which = set(sys.argv[1:]) or {"dw", "pw"}
if "dw" in which:
    print("depthwise: C HxW k | ours us (GB/s) | torch us")
    for C, hw, k in ((16, 160, 3), (8, 160, 7), (32, 80, 3), (16, 80, 7), (64, 40, 3), (32, 40, 7), (128, 20, 3), (64, 20, 7), (64, 80, 3), (128, 40, 3), (256, 20, 3)):
        nbytes = 2 * B * C * hw * hw * 2
        R = max(2, min(16, (400 << 20) // nbytes))
        xs = [rn(B, C, hw, hw) for _ in range(R)]
        w = (torch.randn(C, 1, k, k, device=dev, generator=g) * 0.2)
        wp, wb = ops.pack_dw_weight(w), w.to(dt).contiguous(memory_format=cl)
        out = torch.empty_like(xs[0])
        t_ours = timeit(lambda x: ops.dwconv(x, wp, k, out=out), [(x,) for x in xs])
        t_ref = timeit(lambda x: F.conv2d(x, wb, None, 1, k // 2, 1, C), [(x,) for x in xs])
        print(f"  C={C:4d} {hw:3d}x{hw:<3d} k={k} | {t_ours:8.1f} us ({nbytes / t_ours / 1e3:7.1f}) | {t_ref:8.1f} us")
if "pw" in which:
    print("pointwise: K -> N HxW | ours us (GB/s) | torch conv + bias_act us")
    for K, N, hw in ((32, 32, 160), (16, 8, 160), (8, 16, 160), (48, 16, 160), (64, 64, 160), (64, 64, 80), (32, 16, 80), (96, 32, 80), (128, 128, 80),
                     (128, 128, 40), (192, 128, 40), (384, 128, 40), (256, 256, 20), (512, 256, 20), (128, 384, 20), (384, 256, 20)):
        nbytes = B * hw * hw * (K + N) * 2
        R = max(2, min(16, (400 << 20) // nbytes))
        xs = [rn(B, K, hw, hw) for _ in range(R)]
        w = torch.randn(N, K, device=dev, generator=g) * (K ** -0.5)
        bias = torch.randn(N, device=dev, generator=g)
        wpk = ops.pack_pw_weight(w, [K], dt, B * hw * hw)
        wb = w.view(N, K, 1, 1).to(dt).contiguous(memory_format=cl)
        out = torch.empty(B, N, hw, hw, device=dev, dtype=dt).contiguous(memory_format=cl)
        t_ours = timeit(lambda x: ops.pwconv([x], wpk, N, bias=bias, act=ops.ACT_SILU, out=out), [(x,) for x in xs])
        t_ref = timeit(lambda x: ops.bias_act(F.conv2d(x, wb), bias, ops.ACT_SILU), [(x,) for x in xs])
        print(f"  {K:4d}->{N:<4d} {hw:3d}x{hw:<3d} | {t_ours:8.1f} us ({nbytes / t_ours / 1e3:7.1f}) | {t_ref:8.1f} us")
if "c3" in which:
    print("3x3 conv: B C -> N HxW s | ours us | torch conv + bias_act us")
    for Bc, C, N, hw, s in ((64, 16, 32, 320, 2), (64, 32, 64, 160, 2), (64, 64, 128, 80, 2), (64, 128, 256, 40, 2), (192, 16, 8, 80, 1), (192, 32, 16, 40, 1),
                            (192, 64, 32, 20, 1), (192, 128, 64, 10, 1), (64, 64, 64, 80, 1), (64, 128, 64, 40, 1), (64, 256, 64, 20, 1), (64, 64, 64, 40, 1),
                            (64, 64, 64, 20, 1)):
        nb_in = Bc * C * hw * hw * 2
        R = max(2, min(8, (300 << 20) // nb_in))
        xs = [rn(Bc, C, hw, hw) for _ in range(R)]
        w = torch.randn(N, C, 3, 3, device=dev, generator=g) * (9 * C) ** -0.5
        bias = torch.randn(N, device=dev, generator=g)
        ho = (hw - 1) // s + 1
        wb = w.to(dt).contiguous(memory_format=cl)
        tiles = ops.conv3x3_tiles(N, C, Bc, hw, hw, s)
        try:
            wpk = ops.pack_conv3x3_weight(w, dt, Bc * ho * ho)
            t_ours = timeit(lambda x: ops.conv3x3(x, wpk, N, bias=bias, act=ops.ACT_SILU, stride=s), [(x,) for x in xs])
        except Exception as e:  # noqa: BLE001
            t_ours = float("nan")
        t_ref = timeit(lambda x: ops.bias_act(F.conv2d(x, wb, None, s, 1), bias, ops.ACT_SILU), [(x,) for x in xs])
        print(f"  B={Bc:3d} {C:4d}->{N:<4d} {hw:3d}x{hw:<3d} s{s} tiles{tiles} | {t_ours:8.1f} us | {t_ref:8.1f} us")
