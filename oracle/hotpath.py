"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the EdgeLine-YOLO hot path.

Each function restates one reference function in closed form (no reference code is
imported or called here) and cites the reference `file:line` it follows; paths are
relative to `/root/reference/ultralytics/`.  Floating-point ops are written in torch
fp32 on the CPU, the index/ordering work of NMS in numpy + plain C (`nms_ref.c`).
Pinned against reference-generated fixtures by `tests/test_oracle_golden.py`.
"""
from __future__ import annotations

import ctypes
import math
import os
import subprocess

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))

# float32(2^-1/2) squared, rounded to float32: the magnitude of every 2x2 Haar tap that
# `_PywtDWT2D.__init__` builds with einsum on float32 taps (nn/modules/block.py:3597-3609).
HAAR_K = float(np.float32(np.float32(1.0 / math.sqrt(2.0)) * np.float32(1.0 / math.sqrt(2.0))))


# --------------------------------------------------------------------------- a1: DWT
def dwt_haar(x: torch.Tensor):
    """Single-level 2-D Haar analysis, nn/modules/block.py:3619-3642 (`_PywtDWT2D.forward`).

    Even kernel => no padding (block.py:3617), stride 2, floor(H/2) x floor(W/2) outputs.
    Band sign patterns: LL [[+,+],[+,+]], LH [[+,-],[+,-]], HL [[+,+],[-,-]], HH [[+,-],[-,+]].
    """
    H2, W2 = x.shape[-2] // 2, x.shape[-1] // 2
    a = x[..., 0 : 2 * H2 : 2, 0 : 2 * W2 : 2]
    b = x[..., 0 : 2 * H2 : 2, 1 : 2 * W2 : 2]
    c = x[..., 1 : 2 * H2 : 2, 0 : 2 * W2 : 2]
    d = x[..., 1 : 2 * H2 : 2, 1 : 2 * W2 : 2]
    k = HAAR_K
    return k * (a + b + c + d), k * (a - b + c - d), k * (a + b - c - d), k * (a - b - c + d)


def dwt_haar_adjoint(gLL, gLH, gHL, gHH, H: int, W: int):
    """Transpose of `dwt_haar` (what autograd of the grouped stride-2 conv computes)."""
    k = HAAR_K
    g = gLL.new_zeros(*gLL.shape[:-2], H, W)
    H2, W2 = gLL.shape[-2:]
    g[..., 0 : 2 * H2 : 2, 0 : 2 * W2 : 2] = k * (gLL + gLH + gHL + gHH)
    g[..., 0 : 2 * H2 : 2, 1 : 2 * W2 : 2] = k * (gLL - gLH + gHL - gHH)
    g[..., 1 : 2 * H2 : 2, 0 : 2 * W2 : 2] = k * (gLL + gLH - gHL - gHH)
    g[..., 1 : 2 * H2 : 2, 1 : 2 * W2 : 2] = k * (gLL - gLH - gHL + gHH)
    return g


def haar_dwt2d(x: torch.Tensor):
    """`HaarDWT2D.forward`, nn/modules/block.py:225-259: grouped stride-2 conv with exact 0.5 taps; sub-band order
    LL, LH = [[+,+],[-,-]] (vertical difference), HL = [[+,-],[+,-]] (horizontal difference), HH -- LH / HL swapped with respect to
    `_PywtDWT2D`.  H and W even (`y.view(B, C, 4, H // 2, W // 2)` fails otherwise)."""
    a, b = x[..., 0::2, 0::2], x[..., 0::2, 1::2]
    c, d = x[..., 1::2, 0::2], x[..., 1::2, 1::2]
    return 0.5 * (a + b + c + d), 0.5 * (a + b - c - d), 0.5 * (a - b + c - d), 0.5 * (a - b - c + d)


def ihaar_dwt2d(LL, LH, HL, HH):
    """The inverse of `haar_dwt2d` (the Haar bank with 0.5 taps is orthogonal, so the synthesis is the transpose of the analysis).
    PARITY UNPINNED: the reference's `IHaarDWT2D` (block.py:2714-2750) cannot be constructed (Q5: a commented-out class header leaves a
    foreign `__init__(self, dim, num_heads)` / `forward(self, x)` inside it), and the forward it was meant to have feeds a band-major
    `cat` into a transposed conv whose groups expect channel-interleaved bands, which mixes channels.  The definition here is the one the
    module's name and its use in `WaveletMixerMultiLevel` (:2656-2659, reconstruct one level up) require: idwt(dwt(x)) == x.  The centre
    crop of the four bands to their common size is the reference's (:2729-2735)."""
    Hm, Wm = min(t.shape[-2] for t in (LL, LH, HL, HH)), min(t.shape[-1] for t in (LL, LH, HL, HH))

    def crop(t):
        dh, dw = (t.shape[-2] - Hm) // 2, (t.shape[-1] - Wm) // 2
        return t[..., dh : dh + Hm, dw : dw + Wm]

    LL, LH, HL, HH = crop(LL), crop(LH), crop(HL), crop(HH)
    x = LL.new_zeros(*LL.shape[:-2], 2 * Hm, 2 * Wm)
    x[..., 0::2, 0::2] = 0.5 * (LL + LH + HL + HH)
    x[..., 0::2, 1::2] = 0.5 * (LL + LH - HL - HH)
    x[..., 1::2, 0::2] = 0.5 * (LL - LH + HL - HH)
    x[..., 1::2, 1::2] = 0.5 * (LL - LH - HL + HH)
    return x


def wtconv2d(x, base_w, base_b, base_scale, wave_w, wave_scale, stride: int = 1):
    """WTConv2d.forward, nn/modules/conv.py:540-598 (db1 = Haar), SURVEY 8f-3.  `wave_w[i]` (4C,1,k,k) / `wave_scale[i]` (1,4C,1,1) per level.

    The module's filter bank (create_2d_wavelet_filter, conv.py:408-428) orders the sub-bands (LL, [[+,+],[-,-]], [[+,-],[+,-]], HH),
    i.e. bands 1 and 2 swapped with respect to `_PywtDWT2D`; channel index of the depthwise conv = c * 4 + band.  Odd sizes are zero
    padded at the right / bottom before each analysis and cropped after the synthesis; the synthesis is the adjoint of the analysis.
    """
    import torch.nn.functional as F

    k = base_w.shape[-1]
    lls, highs, shapes = [], [], []
    cur = x
    for w, sc in zip(wave_w, wave_scale):
        shapes.append(cur.shape)
        if cur.shape[2] % 2 or cur.shape[3] % 2:
            cur = F.pad(cur, (0, cur.shape[3] % 2, 0, cur.shape[2] % 2))
        LL, LH, HL, HH = dwt_haar(cur)
        B, C, h, w_ = LL.shape
        t = torch.stack((LL, HL, LH, HH), 2).reshape(B, 4 * C, h, w_)
        t = (F.conv2d(t, w, None, padding=k // 2, groups=4 * C) * sc).reshape(B, C, 4, h, w_)
        lls.append(t[:, :, 0])
        highs.append(t[:, :, 1:])
        cur = LL
    nxt = 0
    for _ in range(len(wave_w)):
        ll, hi, shp = lls.pop() + nxt, highs.pop(), shapes.pop()
        nxt = dwt_haar_adjoint(ll, hi[:, :, 1], hi[:, :, 0], hi[:, :, 2], 2 * ll.shape[-2], 2 * ll.shape[-1])[:, :, : shp[2], : shp[3]]
    y = F.conv2d(x, base_w, base_b, padding=k // 2, groups=x.shape[1]) * base_scale + nxt
    return y[:, :, ::stride, ::stride] if stride > 1 else y


# ------------------------------------------------------------------------- a2: merge
def _bilinear_axis(n_in: int, n_out: int):
    """PyTorch `align_corners=False` source index rule used by F.interpolate (block.py:3681-3683)."""
    scale = n_in / n_out
    src = (torch.arange(n_out, dtype=torch.float32) + 0.5) * scale - 0.5
    src = src.clamp_min(0.0)
    i0 = src.floor().long().clamp_max(n_in - 1)
    i1 = (i0 + 1).clamp_max(n_in - 1)
    lam = src - i0.float()
    return i0, i1, lam


def bilinear_resize(x: torch.Tensor, H: int, W: int):
    """Bilinear resize (B,C,h,w)->(B,C,H,W), align_corners=False; block.py:3681-3683."""
    y0, y1, ly = _bilinear_axis(x.shape[-2], H)
    x0, x1, lx = _bilinear_axis(x.shape[-1], W)
    top = x[..., y0, :]
    bot = x[..., y1, :]
    ly = ly.view(-1, 1)
    rows = top * (1.0 - ly) + bot * ly
    return rows[..., x0] * (1.0 - lx) + rows[..., x1] * lx


def band_weights(alpha: torch.Tensor):
    """softplus(alpha) normalised with +1e-6 in the denominator; block.py:3696-3697."""
    w = torch.log1p(torch.exp(alpha.float()))
    return w / (w.sum() + 1e-6)


def wave_merge(b, LLp, LHp, HLp, HHp, alpha):
    """Upsample each processed band to b's size, scale by w[i], concat behind b; block.py:3699-3708."""
    H, W = b.shape[-2:]
    w = band_weights(alpha)
    ups = [bilinear_resize(t.float(), H, W) * w[i] for i, t in enumerate((LLp, LHp, HLp, HHp))]
    return torch.cat([b.float()] + ups, dim=1)


def gated_residual(b, y, gamma):
    """`b + tanh(gamma) * y`; block.py:3710."""
    return b.float() + torch.tanh(gamma.float()) * y.float()


# --------------------------------------------------------------------- a4: attention
def linear_attention_core(qkv: torch.Tensor, heads: int):
    """Softmax-feature-map linear attention between the qkv and proj convs.

    nn/modules/block.py:3364-3372: qkv channel index = t*C + head*d + j; K softmax over d,
    Q softmax over N, ctx = K^T V (d x d), y = Q ctx, output channel = head*d + j.
    """
    B, C3, H, W = qkv.shape
    C, N = C3 // 3, H * W
    d = C // heads
    t = qkv.float().reshape(B, 3, heads, d, N)
    q, k, v = t[:, 0], t[:, 1], t[:, 2]  # (B, h, d, N)
    k = torch.softmax(k, dim=2)  # over d for every token
    p = torch.exp(q - q.amax(dim=3, keepdim=True))  # softmax over N, split in P / s
    s = p.sum(dim=3, keepdim=True)
    ctx = torch.einsum("bhin,bhjn->bhij", k, v)  # (B,h,d_i,d_j)
    y = torch.einsum("bhin,bhij->bhjn", p / s, ctx)  # (B,h,d_j,N)
    return y.reshape(B, C, H, W)


# ------------------------------------------------------------------------ a6: DGQP
def dgqp_stats(box: torch.Tensor, reg_max: int = 16, topk: int = 4):
    """(B,4*reg_max,H,W) -> (B,4*(topk+1),H,W): top-k softmax probs (descending) + mean.

    nn/modules/head.py:232-240; channel = side*(topk+1) + j, j<topk the sorted top-k, j=topk the
    mean over bins (identically 1/reg_max).
    """
    B, _, H, W = box.shape
    prob = torch.softmax(box.float().reshape(B, 4, reg_max, H, W), dim=2)
    top = torch.sort(prob, dim=2, descending=True).values[:, :, :topk]
    mean = prob.sum(dim=2, keepdim=True) / reg_max
    return torch.cat([top, mean], dim=2).reshape(B, 4 * (topk + 1), H, W)


def dgqp_quality(box, w1, b1, w2, b2, reg_max: int = 16, topk: int = 4):
    """Quality map q=(B,1,H,W): 1x1 conv(20->64)+ReLU+1x1 conv(64->1)+sigmoid; head.py:242-243, 847-854."""
    stat = dgqp_stats(box, reg_max, topk)
    hid = torch.einsum("oc,bchw->bohw", w1.float().reshape(w1.shape[0], -1), stat) + b1.float().view(1, -1, 1, 1)
    hid = hid.clamp_min(0.0)
    z = torch.einsum("oc,bchw->bohw", w2.float().reshape(w2.shape[0], -1), hid) + b2.float().view(1, -1, 1, 1)
    return torch.sigmoid(z)


# ---------------------------------------------------------------------- a7: decode
def anchor_grid(shapes, strides):
    """Cell centres (x+.5, y+.5), level-major then row-major; utils/tal.py:333-345."""
    pts, st = [], []
    for (h, w), s in zip(shapes, strides):
        sy, sx = torch.meshgrid(torch.arange(h, dtype=torch.float32) + 0.5, torch.arange(w, dtype=torch.float32) + 0.5, indexing="ij")
        pts.append(torch.stack((sx, sy), -1).reshape(-1, 2))
        st.append(torch.full((h * w,), float(s)))
    return torch.cat(pts), torch.cat(st)


def gfl_decode(boxes, clss, quals, strides, reg_max: int = 16):
    """DFL integral + dist2bbox(xywh) * stride + sigmoid(cls) * clamp(q) -> (B, 4+nc, A).

    head.py:301-345 (`_inference_with_quality`), block.py:87-90 (`DFL.forward`),
    tal.py:348-357 (`dist2bbox`).  `boxes[i]` (B,4*reg_max,Hi,Wi), `clss[i]` (B,nc,Hi,Wi),
    `quals[i]` (B,1,Hi,Wi).
    """
    B = boxes[0].shape[0]
    box = torch.cat([t.float().reshape(B, 4 * reg_max, -1) for t in boxes], 2)
    cls = torch.cat([t.float().reshape(B, t.shape[1], -1) for t in clss], 2)
    q = torch.cat([t.float().reshape(B, 1, -1) for t in quals], 2)
    anc, st = anchor_grid([t.shape[-2:] for t in boxes], strides)
    A = box.shape[2]
    prob = torch.softmax(box.reshape(B, 4, reg_max, A), dim=2)
    ltrb = (prob * torch.arange(reg_max, dtype=torch.float32).view(1, 1, -1, 1)).sum(2)  # (B,4,A)
    ax, ay = anc[:, 0].view(1, A), anc[:, 1].view(1, A)
    x1, y1 = ax - ltrb[:, 0], ay - ltrb[:, 1]
    x2, y2 = ax + ltrb[:, 2], ay + ltrb[:, 3]
    dbox = torch.stack(((x1 + x2) / 2, (y1 + y2) / 2, x2 - x1, y2 - y1), 1) * st.view(1, 1, A)
    score = torch.sigmoid(cls) * q.clamp(1e-6, 1 - 1e-6)
    return torch.cat((dbox, score), 1)


# -------------------------------------------------------------------------- a9: NMS
_LIB = None


def _nms_lib():
    """Compile (once) and load the plain-C greedy NMS in `nms_ref.c`."""
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "libel_oracle.so")
        src = os.path.join(_HERE, "nms_ref.c")
        if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
            subprocess.check_call(["make", "-s", "-C", _HERE, "libel_oracle.so"])
        _LIB = ctypes.CDLL(so)
        _LIB.el_oracle_nms.restype = ctypes.c_int
        _LIB.el_oracle_nms.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_double, ctypes.c_void_p]
    return _LIB


def nms_greedy(boxes: np.ndarray, scores: np.ndarray, iou_thres: float) -> np.ndarray:
    """`torchvision.ops.nms` semantics (third-party; call site utils/ops.py:296).

    Stable descending sort by score, then greedy sweep; a box is dropped iff its IoU with an
    earlier kept box is strictly greater than the threshold; no epsilon in the union; returns
    int64 indices in descending-score order.
    """
    boxes = np.ascontiguousarray(boxes, dtype=np.float32).reshape(-1, 4)
    scores = np.ascontiguousarray(scores, dtype=np.float32).reshape(-1)
    n = boxes.shape[0]
    keep = np.empty(n, dtype=np.int64)
    k = _nms_lib().el_oracle_nms(boxes.ctypes.data, scores.ctypes.data, n, ctypes.c_double(float(iou_thres)), keep.ctypes.data)
    return keep[:k]


def xywh2xyxy(b: np.ndarray) -> np.ndarray:
    """utils/ops.py:416-433, fp32: half = wh / 2; xy - half, xy + half."""
    b = b.astype(np.float32)
    half = b[..., 2:4] / np.float32(2)
    return np.concatenate((b[..., 0:2] - half, b[..., 0:2] + half), -1)


def nms_candidates(pred_img: np.ndarray, conf_thres, multi_label, classes=None, max_nms=30000):
    """Candidate list of one image in the reference's order; utils/ops.py:253-286.

    `pred_img` is (4+nc, A) xywh + class scores.  Returns (n,6) float32 rows
    [x1,y1,x2,y2,score,cls].  Over `max_nms` the reference keeps the top scores through an
    argsort whose tie order is implementation-defined (ops.py:286); this oracle (and the CUDA
    path) use the stable order: descending score, ties by ascending candidate index.
    """
    p = np.ascontiguousarray(pred_img.T, dtype=np.float32)  # (A, 4+nc)
    nc = p.shape[1] - 4
    conf = np.float32(conf_thres)
    p = p[p[:, 4:].max(1) > conf]
    if p.shape[0] == 0:
        return np.zeros((0, 6), np.float32)
    box = xywh2xyxy(p[:, :4])
    cls = p[:, 4:]
    if multi_label and nc > 1:
        i, j = np.nonzero(cls > conf)  # row-major: anchor-major, class-minor
        x = np.concatenate((box[i], cls[i, j][:, None], j[:, None].astype(np.float32)), 1)
    else:
        j = cls.argmax(1)  # first maximum
        c = cls[np.arange(cls.shape[0]), j]
        x = np.concatenate((box, c[:, None], j[:, None].astype(np.float32)), 1)[c > conf]
    if classes is not None:
        x = x[np.isin(x[:, 5], np.asarray(classes, dtype=np.float32))]
    if x.shape[0] > max_nms:
        x = x[np.argsort(-x[:, 4], kind="stable")[:max_nms]]
    return x.astype(np.float32)


def non_max_suppression(pred, conf_thres=0.25, iou_thres=0.45, classes=None, agnostic=False, multi_label=False,
                        max_det=300, max_nms=30000, max_wh=7680):
    """utils/ops.py:167-316 for detection (nm=0, not rotated, no apriori labels, no time limit).

    `pred` (B, 4+nc, A) float32.  Returns (list of (k,6) float32 arrays, list of int64 keep-index
    arrays into each image's candidate list).
    """
    pred = np.asarray(pred, dtype=np.float32)
    outs, keeps = [], []
    for img in pred:
        x = nms_candidates(img, conf_thres, multi_label, classes, max_nms)
        if x.shape[0] == 0:
            outs.append(np.zeros((0, 6), np.float32))
            keeps.append(np.zeros((0,), np.int64))
            continue
        c = x[:, 5:6] * np.float32(0 if agnostic else max_wh)
        keep = nms_greedy(x[:, :4] + c, x[:, 4], iou_thres)[:max_det]
        outs.append(x[keep])
        keeps.append(keep)
    return outs, keeps


# ----------------------------------------------------------------------- a10: QFL
def quality_focal_loss(pred, target, beta: float = 2.0):
    """Elementwise QFL and d(loss)/d(pred); utils/loss.py:49-63.

    loss = BCEwithLogits(x,t) * ([t>0]|t-p|^beta + [t<=0] p^beta), p = sigmoid(x); the modulating
    factor is *not* detached, so the gradient has both terms.
    """
    x, t = pred.double(), target.double().expand_as(pred)
    p = torch.sigmoid(x)
    bce = torch.clamp_min(x, 0) - x * t + torch.log1p(torch.exp(-x.abs()))
    pos = t > 0
    diff = torch.where(pos, (t - p).abs(), p)
    scale = diff.pow(beta)
    loss = bce * scale
    dbce = p - t
    sgn = torch.where(pos, -torch.sign(t - p), torch.ones_like(p))  # d|t-p|/dx = -sign(t-p) p(1-p); d p/dx = p(1-p)
    dscale = beta * diff.pow(beta - 1) * sgn * p * (1 - p)
    grad = dbce * scale + bce * dscale
    return loss.float(), grad.float()


# ----------------------------------------------------------------------- a11: DFL
def dfl_loss(pred_dist, target, reg_max: int = 16):
    """DFLoss forward (n,1) and d(sum of loss)/d(pred_dist) (n*4, reg_max); utils/loss.py:209-224.

    target (n,4) is clamped to [0, reg_max-1-0.01]; tl=trunc, tr=tl+1, wl=tr-t, wr=1-wl;
    loss = mean over the 4 sides of CE(pred,tl)*wl + CE(pred,tr)*wr.
    """
    t = target.float().clamp(0, reg_max - 1 - 0.01)
    tl = t.long()
    tr = tl + 1
    wl = tr.float() - t
    wr = 1.0 - wl
    logp = torch.log_softmax(pred_dist.float(), dim=1)
    ce_l = -logp.gather(1, tl.view(-1, 1)).view_as(t)
    ce_r = -logp.gather(1, tr.view(-1, 1)).view_as(t)
    loss = (ce_l * wl + ce_r * wr).mean(-1, keepdim=True)
    grad = torch.softmax(pred_dist.float(), dim=1)
    grad.scatter_add_(1, tl.view(-1, 1), -wl.view(-1, 1))
    grad.scatter_add_(1, tr.view(-1, 1), -wr.view(-1, 1))
    return loss, grad / 4.0
