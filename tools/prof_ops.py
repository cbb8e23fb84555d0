"""Runs each hot-path kernel of the engine once at the benchmark shapes (B=64, EdgeLine-n, 640^2, bf16 NHWC) for ncu:
   ncu --set full --import-source on -k regex:'merge|dwt|gated|decode|sweep|linattn|sort' -o out python tools/prof_ops.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from edge_yolo_b200 import ops  # noqa: E402

dev, dt, cl = "cuda", torch.bfloat16, torch.channels_last
g = torch.Generator(device=dev).manual_seed(0)
rn = lambda *s: torch.randn(*s, device=dev, generator=g).to(dt).contiguous(memory_format=cl)
B = int(os.environ.get("B", "64"))
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
which = set(os.environ.get("OPS", "dwt,merge,gated,attn,detect,pw,dw,stem").split(","))
for c, hw in ((16, 160), (32, 80), (64, 40), (128, 20)):
    b = rn(B, c, hw, hw)
    bands = [rn(B, c // 2, hw // 2, hw // 2) for _ in range(4)]
    alpha = torch.tensor([0.5, 0.2, 0.2, 0.1], device=dev)
    y = rn(B, c, hw, hw)
    gamma = torch.tensor(0.5, device=dev)
    cat = torch.empty(B, 3 * c, hw, hw, device=dev, dtype=dt).contiguous(memory_format=cl)
    flush.zero_()
    if "dwt" in which:
        ops.dwt_haar(b)
    flush.zero_()
    if "merge" in which:
        ops.wave_merge(b, *bands, alpha)
    flush.zero_()
    if "gated" in which:
        ops.gated_residual(b, y, gamma, inplace=True, out2=cat[:, c : 2 * c])
    if "merge" in which:
        flush.zero_()
        ops.wave_merge_bands(*bands, alpha, hw, hw)
if "pw" in which:  # the engine's 1x1 convs: cv1 with split outputs, concat-free cv2, DSConv.pw with shortcut, fuse + gated residual
    for srcs_c, N, hw, kw in (([64], 64, 160, {}), ([32, 32, 32], 64, 80, {}), ([128], 128, 80, {}), ([16], 32, 160, dict(res=True)), ([256], 256, 20, {})):
        xs = [rn(B, c, hw, hw) for c in srcs_c]
        K = sum(srcs_c)
        w = torch.randn(N, K, device=dev, generator=g) * K ** -0.5
        bias = torch.randn(N, device=dev, generator=g)
        wpk = ops.pack_pw_weight(w, srcs_c, dt, B * hw * hw)
        res = rn(B, N, hw, hw) if kw.get("res") else None
        out = torch.empty(B, N, hw, hw, device=dev, dtype=dt).contiguous(memory_format=cl)
        flush.zero_()
        ops.pwconv(xs, wpk, N, bias=bias, act=ops.ACT_SILU, residual=res, out=out)
if "dw" in which:
    for C, hw, k, epi in ((16, 160, 7, False), (32, 80, 7, False), (64, 40, 7, False), (64, 80, 3, True), (128, 40, 3, True)):  # the engine's sites
        x = rn(B, C, hw, hw)
        wp = ops.pack_dw_weight(torch.randn(C, 1, k, k, device=dev, generator=g) * 0.2)
        bias = torch.randn(C, device=dev, generator=g) if epi else None
        flush.zero_()
        ops.dwconv(x, wp, k, bias=bias, act=ops.ACT_SILU if epi else ops.ACT_NONE)
if "stem" in which:
    img = torch.randint(0, 256, (B, 640, 640, 3), dtype=torch.uint8, device=dev, generator=g)
    w0 = torch.randn(16, 3, 3, 3, device=dev, generator=g) * 0.1 / 255
    b0 = torch.randn(16, device=dev, generator=g)
    flush.zero_()
    ops.stem_conv_u8(img, w0, b0)
if "attn" in which:
    qkv = rn(B, 384, 20, 20)
    flush.zero_()
    ops.linear_attention(qkv, 2)
if "detect" in which:
    sizes = ((80, 80), (40, 40), (20, 20))
    boxes = [rn(B, 64, h, w) * 0.5 for h, w in sizes]
    clss = [rn(B, 80, h, w) * 0.1 for h, w in sizes]  # sigmoid ~0.5 -> with q ~0.5 about half of the anchors pass conf .25
    ws = [(torch.randn(1280, device=dev, generator=g) * 0.1, torch.zeros(64, device=dev), torch.randn(64, device=dev, generator=g) * 0.1,
           torch.zeros(1, device=dev)) for _ in sizes]
    flush.zero_()
    out, cnt = ops.gfl_detect(boxes, clss, ws, [8.0, 16.0, 32.0], conf_thres=0.25, iou_thres=0.7)
    print("kept", int(cnt.sum()))
torch.cuda.synchronize()
