"""The honest "before" of every custom op ON THE SAME GPU: the reference's PyTorch formulation (restated from SURVEY.md section 8a:
grouped-conv DWT, F.interpolate + cat merge, softmax / bmm attention, softmax / topk / conv DGQP + DFL decode, per-image
torchvision NMS loop) timed eagerly on the B200 beside the library's kernel for the same tensors (B = 64, EdgeLine-n @ 640^2 shapes,
bf16 NHWC for the maps, fp32 for decode output / NMS).  Measurement aid, not product code.   python tools/bench_ref_eager.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from edge_yolo_b200 import ops  # noqa: E402
from edge_yolo_b200.nms import non_max_suppression  # noqa: E402

dev, dt, cl = "cuda", torch.bfloat16, torch.channels_last
g = torch.Generator(device=dev).manual_seed(0)
rn = lambda *s: torch.randn(*s, device=dev, generator=g).to(dt).contiguous(memory_format=cl)
B = 64


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    e1.synchronize()
    return e0.elapsed_time(e1) * 1e3 / n


def row(name, ref_us, our_us):
    print(f"{name:44s} reference eager {ref_us:9.1f} us   library {our_us:8.1f} us   x{ref_us / our_us:6.1f}", flush=True)


# ---- a1 DWT: conv2d with a (4C,1,2,2) filter rebuilt by repeat on every call (block.py:3619-3642)
s = 2 ** -0.5
filt = torch.tensor([[[s * s, s * s], [s * s, s * s]], [[s * s, -s * s], [s * s, -s * s]], [[s * s, s * s], [-s * s, -s * s]], [[s * s, -s * s], [-s * s, s * s]]], device=dev).unsqueeze(1)


def dwt_ref(x):
    Bc, C, H, W = x.shape
    y = F.conv2d(x, filt.to(x.dtype).repeat(C, 1, 1, 1), stride=2, groups=C).view(Bc, C, 4, H // 2, W // 2)
    return y[:, :, 0], y[:, :, 1], y[:, :, 2], y[:, :, 3]


for c, hw in ((16, 160), (32, 80), (64, 40), (128, 20)):
    x = rn(B, c, hw, hw)
    row(f"a1 DWT split        ({B},{c},{hw},{hw})", timeit(lambda: dwt_ref(x)), timeit(lambda: ops.dwt_haar(x)))

# ---- a2 merge + gated residual: 4 interpolates, scale, cat, (fuse conv excluded: same conv either way), b + tanh(gamma) * y
alpha = torch.tensor([0.5, 0.2, 0.2, 0.1], device=dev)
gamma = torch.tensor(0.5, device=dev)
for c, hw in ((16, 160), (32, 80), (64, 40), (128, 20)):
    b = rn(B, c, hw, hw)
    bands = [rn(B, c // 2, hw // 2, hw // 2) for _ in range(4)]
    y = rn(B, c, hw, hw)

    def merge_ref():
        w = F.softplus(alpha)
        w = w / (w.sum() + 1e-6)
        ups = [F.interpolate(t, size=(hw, hw), mode="bilinear", align_corners=False) * w[i].to(t.dtype) for i, t in enumerate(bands)]
        cat = torch.cat([b] + ups, 1)
        return cat, b + torch.tanh(gamma).to(b.dtype) * y

    def merge_ours():
        return ops.wave_merge(b, *bands, alpha), ops.gated_residual(b, y, gamma)

    row(f"a2 merge + gate     ({B},{c},{hw},{hw})", timeit(merge_ref), timeit(merge_ours))

# ---- a4 linear attention core (block.py:3360-3373)
qkv = rn(B, 384, 20, 20)


def attn_ref():
    Bc, C3, H, W = qkv.shape
    h, d, N = 2, 64, H * W
    q, k, v = qkv.reshape(Bc, 3, h, d, N).permute(1, 0, 2, 4, 3)
    k = k.softmax(-1)
    q = q.softmax(-2)
    return (q @ (k.transpose(-2, -1) @ v)).transpose(2, 3).reshape(Bc, h * d, H, W)


row(f"a4 attention core   ({B},384,20,20)", timeit(attn_ref), timeit(lambda: ops.linear_attention(qkv, 2)))

# ---- a6 + a7 DGQP + decode (head.py:227-243, 301-345)
sizes, strides, nc = ((80, 80), (40, 40), (20, 20)), [8.0, 16.0, 32.0], 80
boxes = [rn(B, 64, h, w) * 0.5 for h, w in sizes]
clss = [rn(B, nc, h, w) * 0.1 for h, w in sizes]
ws = [(torch.randn(1280, device=dev, generator=g) * 0.1, torch.zeros(64, device=dev), torch.randn(64, device=dev, generator=g) * 0.1, torch.zeros(1, device=dev)) for _ in sizes]
anc = torch.cat([torch.stack(torch.meshgrid(torch.arange(w, device=dev) + 0.5, torch.arange(h, device=dev) + 0.5, indexing="xy"), -1).reshape(-1, 2) for h, w in sizes])
st = torch.cat([torch.full((h * w,), s_, device=dev) for (h, w), s_ in zip(sizes, strides)])
proj = torch.arange(16, device=dev, dtype=torch.float32)


def decode_ref():
    qs = []
    for bx, (w1, b1, w2, b2) in zip(boxes, ws):
        Bc, _, H, W = bx.shape
        prob = bx.float().view(Bc, 4, 16, H, W).softmax(2)
        top = prob.topk(4, dim=2).values
        stat = torch.cat([top, prob.mean(2, keepdim=True)], 2).view(Bc, 20, H, W)
        hid = F.relu(F.conv2d(stat, w1.view(64, 20, 1, 1), b1))
        qs.append(torch.sigmoid(F.conv2d(hid, w2.view(1, 64, 1, 1), b2)))
    x_cat = torch.cat([torch.cat((bx, c_), 1).float().view(B, 64 + nc, -1) for bx, c_ in zip(boxes, clss)], 2)
    box, cls = x_cat.split((64, nc), 1)
    q = torch.cat([t.view(B, 1, -1) for t in qs], 2)
    dist = box.view(B, 4, 16, -1).transpose(2, 1).softmax(1).transpose(1, 2).reshape(B, 4, 16, -1).softmax(2)  # DFL: softmax over the bins ...
    ltrb = (box.view(B, 4, 16, -1).softmax(2) * proj.view(1, 1, 16, 1)).sum(2)                                    # ... and the integral
    lt, rb = ltrb.chunk(2, 1)
    a = anc.t().unsqueeze(0)
    x1y1, x2y2 = a - lt, a + rb
    dbox = torch.cat(((x1y1 + x2y2) / 2, x2y2 - x1y1), 1) * st
    return torch.cat((dbox, cls.sigmoid() * q.clamp(1e-6, 1 - 1e-6)), 1), dist


y_ref = decode_ref()[0]
y_our = ops.gfl_decode(boxes, clss, ws, strides)
print("decode max |ref - library| (scores):", float((y_ref[:, 4:] - y_our[:, 4:]).abs().max()))
row(f"a6+a7 DGQP + decode ({B},144,8400) dense y", timeit(decode_ref), timeit(lambda: ops.gfl_decode(boxes, clss, ws, strides)))

# ---- a9 NMS (utils/ops.py:167-316, predict settings) on the decoded tensor: per-image Python loop + torchvision.ops.nms
import torchvision  # noqa: E402


def nms_ref(pred, conf=0.25, iou=0.7, max_det=300, max_nms=30000, max_wh=7680):
    xc = pred[:, 4:].amax(1) > conf
    pred = pred.transpose(-1, -2).clone()
    xy, wh = pred[..., :2].clone(), pred[..., 2:4].clone()
    pred[..., :2], pred[..., 2:4] = xy - wh / 2, xy + wh / 2
    out = []
    for xi, x in enumerate(pred):
        x = x[xc[xi]]
        if not x.shape[0]:
            out.append(x.new_zeros((0, 6)))
            continue
        box, cls = x[:, :4], x[:, 4:]
        cf, j = cls.max(1, keepdim=True)
        x = torch.cat((box, cf, j.float()), 1)[cf.view(-1) > conf]
        if x.shape[0] > max_nms:
            x = x[x[:, 4].argsort(descending=True)[:max_nms]]
        c = x[:, 5:6] * max_wh
        i = torchvision.ops.nms(x[:, :4] + c, x[:, 4], iou)[:max_det]
        out.append(x[i])
    return out


row(f"a9 NMS predict      ({B},84,8400)", timeit(lambda: nms_ref(y_our), 3), timeit(lambda: non_max_suppression(y_our, conf_thres=0.25, iou_thres=0.7), 3))
row(f"a6..a9 fused detect ({B} images)", timeit(lambda: nms_ref(decode_ref()[0]), 3), timeit(lambda: ops.gfl_detect(boxes, clss, ws, strides, conf_thres=0.25, iou_thres=0.7)))
