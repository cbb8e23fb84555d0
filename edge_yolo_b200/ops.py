"""Tensor-level entry points of the EdgeLine hot path: torch tensors in, CUDA kernels (through the
C ABI of libedgeline_b200.so) out.  PyTorch only provides device memory, streams and autograd
plumbing here; there is no CPU or eager-PyTorch fallback -- non-CUDA tensors are a hard error.

Reference functions replaced (paths under /root/reference/ultralytics/):
  dwt_haar            _PywtDWT2D.forward                        nn/modules/block.py:3619-3642
  wave_merge          _WaveletEnhancer.forward (tail)           nn/modules/block.py:3696-3708
  gated_residual      _WaveletEnhancer.forward (last line)      nn/modules/block.py:3710
  linear_attention    LinearAttention.forward (core)            nn/modules/block.py:3364-3372
  gfl_decode          GF2Detect._compute_quality_from_logits +  nn/modules/head.py:227-243, 301-345
                      _inference_with_quality
  nms_batched / nms   non_max_suppression / torchvision nms     utils/ops.py:167-316 / :296
  qfl / dfl           quality_focal_loss / DFLoss               utils/loss.py:22-70 / :209-224
"""
from __future__ import annotations

import ctypes
from ctypes import c_float, c_int32, c_int64, c_size_t, c_void_p

import torch

from . import _lib
from ._lib import EdgelineError, check

_DTYPES = {torch.float32: _lib.EL_F32, torch.float16: _lib.EL_F16, torch.bfloat16: _lib.EL_BF16}


def _dt(t: torch.Tensor) -> int:
    try:
        return _DTYPES[t.dtype]
    except KeyError:
        raise EdgelineError(f"unsupported dtype {t.dtype}") from None


def _need_cuda(*ts: torch.Tensor) -> None:
    for t in ts:
        if not t.is_cuda:
            raise EdgelineError("edge_yolo_b200 kernels need CUDA tensors (there is no CPU fallback)")


def _stream() -> c_void_p:
    return c_void_p(torch.cuda.current_stream().cuda_stream)


def _i64(vals):
    return (c_int64 * len(vals))(*vals)


def _ptrs(ts):
    return (c_void_p * len(ts))(*[t.data_ptr() for t in ts])


def _channels_last(t: torch.Tensor) -> bool:
    return t.dim() == 4 and t.stride(1) == 1 and t.shape[1] > 1


def _empty_like_layout(ref: torch.Tensor, shape, dtype=None) -> torch.Tensor:
    fmt = torch.channels_last if _channels_last(ref) else torch.contiguous_format
    return torch.empty(shape, device=ref.device, dtype=dtype or ref.dtype, memory_format=fmt)


# ------------------------------------------------------------------------------------ DWT
def _dwt_fwd(x: torch.Tensor) -> torch.Tensor:
    _need_cuda(x)
    B, C, H, W = x.shape
    H2, W2 = H // 2, W // 2
    buf = _empty_like_layout(x, (4 * B, C, H2, W2))  # band-major: LL | LH | HL | HH, each (B,C,H2,W2)
    if buf.numel():
        sn, sc, sh, sw = buf.stride()
        check(_lib.lib().el_dwt_haar_fwd(x.data_ptr(), _i64(x.stride()), buf.data_ptr(), _i64((B * sn, sn, sc, sh, sw)),
                                         B, C, H, W, _dt(x), _stream()), "el_dwt_haar_fwd")
    return buf


def _dwt_bwd(gbuf: torch.Tensor, like: torch.Tensor) -> torch.Tensor:
    B, C, H, W = like.shape
    gbuf = gbuf if gbuf.dtype == like.dtype else gbuf.to(like.dtype)
    gx = _empty_like_layout(like, (B, C, H, W))
    sn, sc, sh, sw = gbuf.stride()
    check(_lib.lib().el_dwt_haar_bwd(gbuf.data_ptr(), _i64((B * sn, sn, sc, sh, sw)), gx.data_ptr(), _i64(gx.stride()),
                                     B, C, H, W, _dt(like), _stream()), "el_dwt_haar_bwd")
    return gx


class _DWT(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        ctx.shape, ctx.cl, ctx.dtype = x.shape, _channels_last(x), x.dtype
        return _dwt_fwd(x)

    @staticmethod
    def backward(ctx, g):
        fmt = torch.channels_last if ctx.cl else torch.contiguous_format
        like = torch.empty(ctx.shape, device=g.device, dtype=ctx.dtype, memory_format=fmt)
        g = g.contiguous(memory_format=fmt)
        return _dwt_bwd(g, like)


def dwt_haar(x: torch.Tensor) -> torch.Tensor:
    """Single-level Haar analysis.  Returns one buffer (4*B, C, H//2, W//2): rows [0,B) = LL, [B,2B) = LH,
    [2B,3B) = HL, [3B,4B) = HH (so `buf[B:]` feeds the shared high-band conv in one call)."""
    if x.requires_grad and torch.is_grad_enabled():
        return _DWT.apply(x)
    return _dwt_fwd(x)


class _IDWT(torch.autograd.Function):
    """Haar synthesis with the analysis as its backward (each is the other's adjoint)."""

    @staticmethod
    def forward(ctx, bands, H, W):
        ctx.cl = _channels_last(bands)
        B = bands.shape[0] // 4
        return _dwt_bwd(bands, _empty_like_layout(bands, (B, bands.shape[1], H, W)))

    @staticmethod
    def backward(ctx, g):
        fmt = torch.channels_last if ctx.cl else torch.contiguous_format
        return _dwt_fwd(g.contiguous(memory_format=fmt)), None, None


def idwt_haar(bands: torch.Tensor, H: int, W: int) -> torch.Tensor:
    """Haar synthesis (adjoint of `dwt_haar`), the single-level case of inverse_2d_wavelet_transform
    (nn/modules/conv.py:438-443).  `bands` is the (4*B, C, H//2, W//2) buffer `dwt_haar` returns."""
    _need_cuda(bands)
    if bands.requires_grad and torch.is_grad_enabled():
        return _IDWT.apply(bands, H, W)
    B = bands.shape[0] // 4
    like = _empty_like_layout(bands, (B, bands.shape[1], H, W))
    return _dwt_bwd(bands, like)


# ----------------------------------------------------------- DWT / IDWT in WTConv2d's coefficient layout
# wavelet_2d_transform / inverse_2d_wavelet_transform (nn/modules/conv.py:430-443) keep the coefficients as (B, C, 4, H/2, W/2) with the
# sub-bands ordered (LL, [[+,+],[-,-]], [[+,-],[+,-]], HH): bands 1 and 2 swapped with respect to `_PywtDWT2D`.  The kernels take
# arbitrary element strides, so the same two entry points serve this layout: the analysis of the TRANSPOSED image yields exactly the
# swapped order, and the (band, B, C, h, w) stride table points straight into the 5-D tensor -- no permute / copy kernels.
def _coeff_view(coeffs: torch.Tensor):
    """(B, C, 4, h, w) tensor -> stride table {band, n, c, h', w'} of its transposed-image view (h' runs over w)."""
    sb, sc, s4, sh, sw = coeffs.stride()
    return (s4, sb, sc, sw, sh)


def _dwt_coeffs_fwd(x: torch.Tensor) -> torch.Tensor:
    _need_cuda(x)
    B, C, H, W = x.shape
    coeffs = torch.empty((B, C, 4, H // 2, W // 2), device=x.device, dtype=x.dtype)
    if coeffs.numel():
        sn, sc, sh, sw = x.stride()
        check(_lib.lib().el_dwt_haar_fwd(x.data_ptr(), _i64((sn, sc, sw, sh)), coeffs.data_ptr(), _i64(_coeff_view(coeffs)),
                                         B, C, W, H, _dt(x), _stream()), "el_dwt_haar_fwd")
    return coeffs


def _dwt_coeffs_bwd(coeffs: torch.Tensor, H: int, W: int) -> torch.Tensor:
    _need_cuda(coeffs)
    B, C = coeffs.shape[:2]
    out = torch.empty((B, C, H, W), device=coeffs.device, dtype=coeffs.dtype)
    sn, sc, sh, sw = out.stride()
    check(_lib.lib().el_dwt_haar_bwd(coeffs.data_ptr(), _i64(_coeff_view(coeffs)), out.data_ptr(), _i64((sn, sc, sw, sh)),
                                     B, C, W, H, _dt(coeffs), _stream()), "el_dwt_haar_bwd")
    return out


class _DWTCoeffs(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        ctx.hw = x.shape[-2:]
        return _dwt_coeffs_fwd(x)

    @staticmethod
    def backward(ctx, g):
        return _dwt_coeffs_bwd(g.contiguous(), *ctx.hw)  # the synthesis is the adjoint of the analysis


class _IDWTCoeffs(torch.autograd.Function):
    @staticmethod
    def forward(ctx, coeffs):
        return _dwt_coeffs_bwd(coeffs, 2 * coeffs.shape[-2], 2 * coeffs.shape[-1])

    @staticmethod
    def backward(ctx, g):
        return _dwt_coeffs_fwd(g.contiguous())


def wavelet_2d_transform(x: torch.Tensor) -> torch.Tensor:
    """Haar case of `wavelet_2d_transform` (conv.py:430-435): (B,C,H,W), H and W even -> (B, C, 4, H/2, W/2)."""
    if x.shape[-1] % 2 or x.shape[-2] % 2:
        raise EdgelineError("wavelet_2d_transform: even H and W expected (WTConv2d pads odd sizes first)")
    return _DWTCoeffs.apply(x) if x.requires_grad and torch.is_grad_enabled() else _dwt_coeffs_fwd(x)


def inverse_2d_wavelet_transform(coeffs: torch.Tensor) -> torch.Tensor:
    """Haar case of `inverse_2d_wavelet_transform` (conv.py:438-443): (B, C, 4, h, w) -> (B, C, 2h, 2w)."""
    coeffs = coeffs.contiguous()
    return _IDWTCoeffs.apply(coeffs) if coeffs.requires_grad and torch.is_grad_enabled() else _dwt_coeffs_bwd(coeffs, 2 * coeffs.shape[-2], 2 * coeffs.shape[-1])


# ---------------------------------------------------------------------------------- merge
def _merge_fwd(b, bands, alpha):
    _need_cuda(b, alpha, *bands)
    B, c, H, W = b.shape
    h, w = bands[0].shape[-2:]
    for t in bands:
        if t.shape != (B, c // 2, h, w) or t.dtype != b.dtype:
            raise EdgelineError(f"wave_merge: band shape/dtype {tuple(t.shape)}/{t.dtype} does not match b {tuple(b.shape)}/{b.dtype}")
    out = _empty_like_layout(b, (B, 3 * c, H, W))
    a = alpha if alpha.dtype == torch.float32 else alpha.float()
    bs = [s for t in bands for s in t.stride()]
    check(_lib.lib().el_wave_merge_fwd(b.data_ptr(), _i64(b.stride()), _ptrs(bands), _i64(bs), a.data_ptr(), out.data_ptr(),
                                       _i64(out.stride()), B, c, H, W, h, w, _dt(b), _stream()), "el_wave_merge_fwd")
    return out


class _Merge(torch.autograd.Function):
    @staticmethod
    def forward(ctx, b, LLp, LHp, HLp, HHp, alpha):
        ctx.save_for_backward(LLp, LHp, HLp, HHp, alpha)
        ctx.b_meta = (b.shape, _channels_last(b))
        return _merge_fwd(b, (LLp, LHp, HLp, HHp), alpha)

    @staticmethod
    def backward(ctx, g):
        LLp, LHp, HLp, HHp, alpha = ctx.saved_tensors
        bands = (LLp, LHp, HLp, HHp)
        (B, c, H, W), cl = ctx.b_meta
        h, w = LLp.shape[-2:]
        fmt = torch.channels_last if cl else torch.contiguous_format
        g = g.contiguous(memory_format=fmt)
        gb = torch.empty((B, c, H, W), device=g.device, dtype=g.dtype, memory_format=fmt)
        gbands = [torch.empty_like(t, memory_format=torch.contiguous_format if not _channels_last(t) else torch.channels_last) for t in bands]
        gw = torch.zeros(4, device=g.device, dtype=torch.float32)
        a = alpha.detach().float()
        bs = [s for t in bands for s in t.stride()]
        gs = [s for t in gbands for s in t.stride()]
        check(_lib.lib().el_wave_merge_bwd(g.data_ptr(), _i64(g.stride()), _ptrs(bands), _i64(bs), a.data_ptr(), gb.data_ptr(),
                                           _i64(gb.stride()), _ptrs(gbands), _i64(gs), gw.data_ptr(), B, c, H, W, h, w, _dt(g),
                                           _stream()), "el_wave_merge_bwd")
        # chain d loss / d w (4 numbers, from the kernel) through w = softplus(alpha) / (sum + 1e-6)
        with torch.enable_grad():
            al = alpha.detach().float().requires_grad_()
            sp = torch.nn.functional.softplus(al)
            wv = sp / (sp.sum() + 1e-6)
            (galpha,) = torch.autograd.grad(wv, al, gw)
        return gb, gbands[0], gbands[1], gbands[2], gbands[3], galpha.to(alpha.dtype)


def wave_merge_bands(LLp, LHp, HLp, HHp, alpha, H: int, W: int) -> torch.Tensor:
    """Inference-engine variant of `wave_merge` without the pass-through copy of b: (B, 2c, H, W) = cat of the four upsampled,
    weighted sub-bands (block.py:3696-3707); the consumer (the `fuse` conv) reads b in place.  No autograd."""
    bands = (LLp, LHp, HLp, HHp)
    _need_cuda(alpha, *bands)
    B, half, h, w = LLp.shape
    for t in bands:
        if t.shape != LLp.shape or t.dtype != LLp.dtype:
            raise EdgelineError("wave_merge_bands: the four bands must share shape and dtype")
    out = torch.empty((B, 4 * half, H, W), device=LLp.device, dtype=LLp.dtype, memory_format=torch.channels_last)
    a = alpha if alpha.dtype == torch.float32 else alpha.float()
    bs = [s for t in bands for s in t.stride()]
    check(_lib.lib().el_wave_merge_fwd(None, None, _ptrs(bands), _i64(bs), a.data_ptr(), out.data_ptr(), _i64(out.stride()), B, 2 * half, H, W,
                                       h, w, _dt(LLp), _stream()), "el_wave_merge_fwd")
    return out


def wave_merge(b, LLp, LHp, HLp, HHp, alpha) -> torch.Tensor:
    """cat[b, up(LLp)*w0, up(LHp)*w1, up(HLp)*w2, up(HHp)*w3] -> (B, 3c, H, W); w from `alpha` on the device."""
    if torch.is_grad_enabled() and any(t.requires_grad for t in (b, LLp, LHp, HLp, HHp, alpha)):
        return _Merge.apply(b, LLp, LHp, HLp, HHp, alpha)
    return _merge_fwd(b, (LLp, LHp, HLp, HHp), alpha)


# ------------------------------------------------------------------------- gated residual
def _gated_fwd(b, y, gamma, out=None, out2=None):
    _need_cuda(b, y, gamma)
    if y.shape != b.shape or y.dtype != b.dtype:
        raise EdgelineError("gated_residual: b and y must have the same shape and dtype")
    B, C, H, W = b.shape
    out = _empty_like_layout(b, b.shape) if out is None else out
    g = gamma if gamma.dtype == torch.float32 else gamma.float()
    check(_lib.lib().el_gated_residual_fwd(b.data_ptr(), _i64(b.stride()), y.data_ptr(), _i64(y.stride()), g.data_ptr(),
                                           out.data_ptr(), _i64(out.stride()), out2.data_ptr() if out2 is not None else None,
                                           _i64(out2.stride()) if out2 is not None else None, B, C, H, W, _dt(b), _stream()),
          "el_gated_residual_fwd")
    return out


class _Gated(torch.autograd.Function):
    @staticmethod
    def forward(ctx, b, y, gamma):
        ctx.save_for_backward(y, gamma)
        return _gated_fwd(b, y, gamma)

    @staticmethod
    def backward(ctx, g):
        y, gamma = ctx.saved_tensors
        g = g if g.dtype == y.dtype else g.to(y.dtype)
        B, C, H, W = y.shape
        gy = _empty_like_layout(y, y.shape)
        ggamma = torch.zeros((), device=y.device, dtype=torch.float32)
        gm = gamma if gamma.dtype == torch.float32 else gamma.float()
        # d/db = g ; d/dy = tanh(gamma) g ; d/dgamma = (1 - tanh^2) <g, y>
        check(_lib.lib().el_gated_residual_bwd(g.data_ptr(), _i64(g.stride()), y.data_ptr(), _i64(y.stride()), gm.data_ptr(), gy.data_ptr(),
                                               _i64(gy.stride()), ggamma.data_ptr(), B, C, H, W, _dt(y), _stream()), "el_gated_residual_bwd")
        return g, gy, ggamma.to(gamma.dtype)


def gated_residual(b, y, gamma, inplace: bool = False, out2: torch.Tensor | None = None) -> torch.Tensor:
    """b + tanh(gamma) * y.  `inplace=True` writes the result over `b`, `out2` receives a second copy (inference only)."""
    if torch.is_grad_enabled() and any(t.requires_grad for t in (b, y, gamma)):
        return _Gated.apply(b, y, gamma)
    return _gated_fwd(b, y, gamma, out=b if inplace else None, out2=out2)


# ------------------------------------------------------------------------------ attention
def _attn_strides(t: torch.Tensor):
    """(batch, channel, token) element strides of a (B, C, H, W) map whose token axis (h*W + w) is uniformly strided."""
    W = t.shape[3]
    if t.stride(2) != W * t.stride(3):
        raise EdgelineError("linear_attention: token axis must be uniformly strided")
    return (t.stride(0), t.stride(1), t.stride(3))


def _linattn_fwd(qkv: torch.Tensor, heads: int) -> torch.Tensor:
    B, C3, H, W = qkv.shape
    C, N = C3 // 3, H * W
    if not _channels_last(qkv) and (qkv.stride(3) != 1 or qkv.stride(2) != W):
        qkv = qkv.contiguous()
    y = _empty_like_layout(qkv, (B, C, H, W))
    check(_lib.lib().el_linattn_fwd(qkv.data_ptr(), _i64(_attn_strides(qkv)), y.data_ptr(), _i64(_attn_strides(y)), B, heads, N, _dt(qkv),
                                    _stream()), "el_linattn_fwd")
    return y


class _LinAttn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, qkv, heads):
        ctx.save_for_backward(qkv)
        ctx.heads = heads
        return _linattn_fwd(qkv, heads)

    @staticmethod
    def backward(ctx, g):
        (qkv,) = ctx.saved_tensors
        B, C3, H, W = qkv.shape
        fmt = torch.channels_last if _channels_last(qkv) else torch.contiguous_format
        q = qkv.contiguous(memory_format=fmt)
        g = g.to(qkv.dtype).contiguous(memory_format=fmt)
        out = torch.empty_like(q, memory_format=fmt)
        check(_lib.lib().el_linattn_bwd(q.data_ptr(), _i64(_attn_strides(q)), g.data_ptr(), _i64(_attn_strides(g)), out.data_ptr(),
                                        _i64(_attn_strides(out)), B, ctx.heads, H * W, _dt(q), _stream()), "el_linattn_bwd")
        return out, None


def linear_attention(qkv: torch.Tensor, heads: int) -> torch.Tensor:
    """qkv (B, 3C, H, W) -> y (B, C, H, W); head_dim must be 64 (the only value the reference model produces)."""
    _need_cuda(qkv)
    C3 = qkv.shape[1]
    if C3 != 3 * heads * 64:
        raise EdgelineError(f"linear_attention: need 3*heads*64 channels, got {C3} with heads={heads}")
    if torch.is_grad_enabled() and qkv.requires_grad:
        return _LinAttn.apply(qkv, heads)
    return _linattn_fwd(qkv, heads)


# --------------------------------------------------------------------------------- decode
def _bias_tables(bias, nl):
    if bias is None:
        return None, None
    box_b, cls_b = bias
    for t in list(box_b) + list(cls_b):
        if t is not None and (t.dtype != torch.float32 or not t.is_contiguous() or not t.is_cuda):
            raise EdgelineError("decode bias vectors must be contiguous fp32 CUDA tensors")
    tab = lambda ts: (c_void_p * nl)(*[t.data_ptr() if t is not None else None for t in ts])
    return tab(box_b), tab(cls_b)


def gfl_decode(boxes, clss, dgqp, strides, want_quality: bool = False, bias=None):
    """boxes[l] (B,64,Hl,Wl), clss[l] (B,nc,Hl,Wl), dgqp[l] = (w1 (64,20), b1 (64), w2 (64), b2 (1)) fp32,
    strides[l] float -> y (B, 4+nc, A) fp32 [, q (B, A) fp32]."""
    nl = len(boxes)
    _need_cuda(*boxes, *clss)
    B, nc = boxes[0].shape[0], clss[0].shape[1]
    dt = _dt(boxes[0])
    hw, bs, cs = [], [], []
    for bx, cl in zip(boxes, clss):
        if bx.shape[1] != 64 or cl.shape[1] != nc or bx.shape[-2:] != cl.shape[-2:] or bx.dtype != boxes[0].dtype or cl.dtype != boxes[0].dtype:
            raise EdgelineError("gfl_decode: inconsistent level shapes / dtypes")
        hw += [bx.shape[2], bx.shape[3]]
        bs += list(bx.stride())
        cs += list(cl.stride())
    A = sum(hw[2 * i] * hw[2 * i + 1] for i in range(nl))
    for w in dgqp:
        for t in w:
            if t.dtype != torch.float32 or not t.is_contiguous() or not t.is_cuda:
                raise EdgelineError("gfl_decode: DGQP weights must be contiguous fp32 CUDA tensors")
    y = torch.empty((B, 4 + nc, A), device=boxes[0].device, dtype=torch.float32)
    q = torch.empty((B, A), device=boxes[0].device, dtype=torch.float32) if want_quality else None
    cols = [_ptrs([w[k] for w in dgqp]) for k in range(4)]
    bb, cb = _bias_tables(bias, nl)
    check(_lib.lib().el_gfl_decode_fwd(nl, _ptrs(boxes), _i64(bs), _ptrs(clss), _i64(cs), (c_int32 * len(hw))(*hw),
                                       (c_float * nl)(*[float(s) for s in strides]), cols[0], cols[1], cols[2], cols[3], bb, cb,
                                       y.data_ptr(), q.data_ptr() if q is not None else None, B, nc, dt, _stream()), "el_gfl_decode_fwd")
    return (y, q) if want_quality else y


def gfl_detect(boxes, clss, dgqp, strides, conf_thres=0.25, iou_thres=0.45, multi_label=False, agnostic=False, classes=None,
               max_det=300, max_nms=30000, max_wh=7680.0, bias=None, workspace=None, stages=7, out=None, cnt=None):
    """Fused decode + NMS of the engine path: same inputs as `gfl_decode`, same outputs as `nms_batched`
    (out (B, max_det, 6), count (B) int32); the dense (B, 4+nc, A) tensor is never written.  Falls back to
    `gfl_decode` + `nms_batched` (identical results) when the head maps are not dense NHWC.

    `stages` (bit 0 decode + candidate emit, bit 1 select + sort, bit 2 sweep) runs part of the chain on the caller's
    `workspace` / `out` / `cnt`: the engine captures stage 1 and stages 6 as two CUDA graphs so that the latency-bound NMS tail
    of batch i overlaps the backbone of batch i+1 (engine.Predictor)."""
    nl = len(boxes)
    _need_cuda(*boxes, *clss)
    B, nc = boxes[0].shape[0], clss[0].shape[1]
    hw, bs, cs = [], [], []
    for bx, cl in zip(boxes, clss):
        hw += [bx.shape[2], bx.shape[3]]
        bs += list(bx.stride())
        cs += list(cl.stride())
    A = sum(hw[2 * i] * hw[2 * i + 1] for i in range(nl))
    L = _lib.lib()
    need = c_size_t()
    check(L.el_gfl_detect_workspace_bytes(B, nc, A, int(bool(multi_label)), int(max_nms), ctypes.byref(need)), "el_gfl_detect_workspace_bytes")
    dev = boxes[0].device
    ws = workspace if workspace is not None and workspace.numel() >= need.value else torch.empty(need.value, device=dev, dtype=torch.uint8)
    if stages != 7 and (workspace is None or ws is not workspace):
        raise EdgelineError("gfl_detect: staged execution needs a caller-owned workspace of el_gfl_detect_workspace_bytes")
    caller_out, caller_cnt = out, cnt
    out = torch.empty((B, max_det, 6), device=dev, dtype=torch.float32) if out is None else out
    cnt = torch.empty((B,), device=dev, dtype=torch.int32) if cnt is None else cnt
    keep = None
    if classes is not None:
        keep = torch.zeros(nc, dtype=torch.int32)
        keep[torch.as_tensor(classes, dtype=torch.long)] = 1
        keep = keep.to(dev)
    cols = [_ptrs([w[k] for w in dgqp]) for k in range(4)]
    bb, cb = _bias_tables(bias, nl)
    st = L.el_gfl_detect_fwd(nl, _ptrs(boxes), _i64(bs), _ptrs(clss), _i64(cs), (c_int32 * len(hw))(*hw),
                             (c_float * nl)(*[float(s) for s in strides]), cols[0], cols[1], cols[2], cols[3], bb, cb, B, nc, _dt(boxes[0]),
                             float(conf_thres), float(iou_thres), int(bool(multi_label)), int(bool(agnostic)),
                             keep.data_ptr() if keep is not None else None, int(max_det), int(max_nms), float(max_wh), int(stages), ws.data_ptr(),
                             need.value, out.data_ptr(), cnt.data_ptr(), None, _stream())
    if st == 2 and stages != 7:
        raise EdgelineError("gfl_detect: staged execution needs dense NHWC head maps")
    if st == 2:  # EL_ERR_UNSUPPORTED: strided / NCHW / unaligned maps -> two-call path with the same kernels downstream
        y = gfl_decode(boxes, clss, dgqp, strides, bias=bias)
        res = nms_batched(y, conf_thres, iou_thres, multi_label=multi_label, agnostic=agnostic, classes=classes, max_det=max_det,
                          max_nms=max_nms, max_wh=max_wh)
        if caller_out is not None:  # the engine's graphs read the results from its own buffers
            caller_out.copy_(res[0])
            caller_cnt.copy_(res[1])
            return caller_out, caller_cnt
        return res
    check(st, "el_gfl_detect_fwd")
    return out, cnt


# ------------------------------------------------------------------------------------ NMS
def nms_batched(pred: torch.Tensor, conf_thres=0.25, iou_thres=0.45, multi_label=False, agnostic=False, classes=None,
                max_det=300, max_nms=30000, max_wh=7680.0, want_index=False):
    """pred (B, 4+nc, A) fp32 -> (out (B, max_det, 6), count (B) int32 [, index (B, max_det) int64]); rows past
    count[b] are undefined.  No host synchronisation."""
    _need_cuda(pred)
    if pred.dtype != torch.float32:
        pred = pred.float()
    pred = pred.contiguous()
    B, no, A = pred.shape
    nc = no - 4
    L = _lib.lib()
    need = c_size_t()
    check(L.el_nms_workspace_bytes(B, nc, A, int(bool(multi_label)), int(max_nms), ctypes.byref(need)), "el_nms_workspace_bytes")
    ws = torch.empty(need.value, device=pred.device, dtype=torch.uint8)
    out = torch.empty((B, max_det, 6), device=pred.device, dtype=torch.float32)
    cnt = torch.empty((B,), device=pred.device, dtype=torch.int32)
    idx = torch.empty((B, max_det), device=pred.device, dtype=torch.int64) if want_index else None
    keep = None
    if classes is not None:
        keep = torch.zeros(nc, dtype=torch.int32)
        keep[torch.as_tensor(classes, dtype=torch.long)] = 1
        keep = keep.to(pred.device)
    check(L.el_nms_batched(pred.data_ptr(), B, nc, A, float(conf_thres), float(iou_thres), int(bool(multi_label)), int(bool(agnostic)),
                           keep.data_ptr() if keep is not None else None, int(max_det), int(max_nms), float(max_wh), ws.data_ptr(),
                           need.value, out.data_ptr(), cnt.data_ptr(), idx.data_ptr() if idx is not None else None, _stream()),
          "el_nms_batched")
    return (out, cnt, idx) if want_index else (out, cnt)


def nms(boxes: torch.Tensor, scores: torch.Tensor, iou_threshold: float) -> torch.Tensor:
    """Same contract as torchvision.ops.nms: int64 keep indices, descending score (one host sync for the count)."""
    _need_cuda(boxes, scores)
    n = boxes.shape[0]
    if n == 0:
        return torch.empty(0, dtype=torch.int64, device=boxes.device)
    boxes = boxes.float().contiguous()
    scores = scores.float().contiguous()
    L = _lib.lib()
    need = c_size_t()
    check(L.el_nms_boxes_workspace_bytes(n, ctypes.byref(need)), "el_nms_boxes_workspace_bytes")
    ws = torch.empty(need.value, device=boxes.device, dtype=torch.uint8)
    keep = torch.empty(n, device=boxes.device, dtype=torch.int64)
    cnt = torch.empty(1, device=boxes.device, dtype=torch.int32)
    check(L.el_nms_boxes(boxes.data_ptr(), scores.data_ptr(), n, float(iou_threshold), ws.data_ptr(), need.value, keep.data_ptr(),
                         cnt.data_ptr(), _stream()), "el_nms_boxes")
    return keep[: int(cnt.item())]


# --------------------------------------------------------------------------------- losses
class _QFL(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, target, beta, reduce_sum):
        _need_cuda(pred, target)
        p = pred.contiguous()
        t = target.expand_as(pred).float().contiguous()
        ctx.save_for_backward(p, t)
        ctx.beta, ctx.reduce_sum = float(beta), reduce_sum
        L, n = _lib.lib(), p.numel()
        if reduce_sum:
            total = torch.empty((), device=p.device, dtype=torch.float32)
            part = torch.empty(L.el_qfl_partials(n), device=p.device, dtype=torch.float32)
            check(L.el_qfl_fwd(p.data_ptr(), t.data_ptr(), n, ctx.beta, _dt(p), None, total.data_ptr(), part.data_ptr(), _stream()), "el_qfl_fwd")
            return total
        loss = torch.empty(p.shape, device=p.device, dtype=torch.float32)
        check(L.el_qfl_fwd(p.data_ptr(), t.data_ptr(), n, ctx.beta, _dt(p), loss.data_ptr(), None, None, _stream()), "el_qfl_fwd")
        return loss

    @staticmethod
    def backward(ctx, g):
        p, t = ctx.saved_tensors
        gp = torch.empty_like(p)
        g = g.float().contiguous()
        args = (None, g.data_ptr()) if ctx.reduce_sum else (g.data_ptr(), None)
        check(_lib.lib().el_qfl_bwd(p.data_ptr(), t.data_ptr(), p.numel(), ctx.beta, _dt(p), args[0], args[1], gp.data_ptr(), _stream()),
              "el_qfl_bwd")
        return gp, None, None, None


def quality_focal_loss(pred, target, beta: float = 2.0, reduction: str = "none"):
    """Same signature as utils/loss.py:22 `quality_focal_loss`; the result is fp32."""
    if reduction == "none":
        return _QFL.apply(pred, target, beta, False)
    total = _QFL.apply(pred, target, beta, True)
    return total / pred.numel() if reduction == "mean" else total


class _DFL(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred_dist, target):
        _need_cuda(pred_dist, target)
        p = pred_dist.contiguous()
        t = target.float().contiguous()
        rows = t.shape[0]
        if p.shape != (rows * 4, 16) or t.shape != (rows, 4):
            raise EdgelineError(f"dfl: expected pred (4n,16) and target (n,4), got {tuple(p.shape)} / {tuple(t.shape)}")
        ctx.save_for_backward(p, t)
        loss = torch.empty((rows, 1), device=p.device, dtype=torch.float32)
        if rows:
            check(_lib.lib().el_dfl_fwd(p.data_ptr(), t.data_ptr(), rows, _dt(p), loss.data_ptr(), _stream()), "el_dfl_fwd")
        return loss

    @staticmethod
    def backward(ctx, g):
        p, t = ctx.saved_tensors
        gp = torch.empty_like(p)
        if t.shape[0]:
            g = g.float().contiguous()
            check(_lib.lib().el_dfl_bwd(p.data_ptr(), t.data_ptr(), t.shape[0], _dt(p), g.data_ptr(), gp.data_ptr(), _stream()), "el_dfl_bwd")
        return gp, None


def dfl_loss(pred_dist: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """DFLoss.__call__ (utils/loss.py:209-224): pred (4n,16) logits, target (n,4) -> (n,1) fp32."""
    return _DFL.apply(pred_dist, target)


class _DFLSide(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, target):
        _need_cuda(pred, target)
        p = pred.reshape(-1, 16).contiguous()
        t = target.reshape(-1).float().contiguous()
        if p.shape[0] != t.shape[0]:
            raise EdgelineError(f"dfl: expected pred (..., 16) and target (...), got {tuple(pred.shape)} / {tuple(target.shape)}")
        ctx.save_for_backward(p, t)
        ctx.pshape = pred.shape
        loss = torch.empty(t.shape[0], device=p.device, dtype=torch.float32)
        if t.shape[0]:
            check(_lib.lib().el_dfl_side_fwd(p.data_ptr(), t.data_ptr(), t.shape[0], _dt(p), loss.data_ptr(), _stream()), "el_dfl_side_fwd")
        return loss.view(target.shape)

    @staticmethod
    def backward(ctx, g):
        p, t = ctx.saved_tensors
        gp = torch.empty_like(p)
        if t.shape[0]:
            g = g.reshape(-1).float().contiguous()
            check(_lib.lib().el_dfl_side_bwd(p.data_ptr(), t.data_ptr(), t.shape[0], _dt(p), g.data_ptr(), gp.data_ptr(), _stream()), "el_dfl_side_bwd")
        return gp.view(ctx.pshape), None


def dfl_side_loss(pred: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """distribution_focal_loss(reduction="none") (utils/loss.py:88-137): pred (..., 16) logits, target (...) in bins -> (...) fp32."""
    return _DFLSide.apply(pred, target)


# --------------------------------------------------------------------------------- ingest
def ingest_u8(src: torch.Tensor, dtype=torch.bfloat16, channels_last: bool = True, out: torch.Tensor | None = None) -> torch.Tensor:
    """uint8 (B,H,W,3) images -> (B,3,H,W) activations / 255 (the predictor's preprocess, engine/predictor.py:117-135)."""
    _need_cuda(src)
    if src.dtype != torch.uint8 or src.dim() != 4 or src.shape[-1] != 3 or not src.is_contiguous():
        raise EdgelineError("ingest_u8: need a contiguous uint8 (B,H,W,3) tensor")
    B, H, W, _ = src.shape
    if out is None:
        out = torch.empty((B, 3, H, W), device=src.device, dtype=dtype,
                          memory_format=torch.channels_last if channels_last else torch.contiguous_format)
    check(_lib.lib().el_ingest_u8(src.data_ptr(), out.data_ptr(), _i64(out.stride()), B, H, W, _dt(out), _stream()), "el_ingest_u8")
    return out


def stem_conv_u8(src: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor, dtype=torch.bfloat16, out: torch.Tensor | None = None) -> torch.Tensor:
    """uint8 (B,H,W,3) images -> SiLU(conv3x3/s2(img/255) + bias) as (B,C0,H/2,W/2) NHWC activations: the predictor's
    preprocess (engine/predictor.py:117-135) fused with layer 0 of the yaml.  weight (C0,3,3,3) fp32 must already carry
    the folded BatchNorm scale and the 1/255."""
    _need_cuda(src, weight, bias)
    if src.dtype != torch.uint8 or src.dim() != 4 or src.shape[-1] != 3 or not src.is_contiguous():
        raise EdgelineError("stem_conv_u8: need a contiguous uint8 (B,H,W,3) tensor")
    B, H, W, _ = src.shape
    C0 = weight.shape[0]
    if weight.dtype != torch.float32 or bias.dtype != torch.float32 or tuple(weight.shape) != (C0, 3, 3, 3) or bias.numel() != C0 \
            or not weight.is_contiguous() or not bias.is_contiguous():
        raise EdgelineError("stem_conv_u8: weight must be contiguous fp32 (C0,3,3,3) and bias fp32 (C0)")
    if out is None:
        out = torch.empty((B, C0, H // 2, W // 2), device=src.device, dtype=dtype, memory_format=torch.channels_last)
    check(_lib.lib().el_stem_conv_u8(src.data_ptr(), weight.data_ptr(), bias.data_ptr(), out.data_ptr(), _i64(out.stride()), B, C0, H, W,
                                     _dt(out), _stream()), "el_stem_conv_u8")
    return out


# ------------------------------------------------------------------------------ epilogues
ACT_NONE, ACT_SILU, ACT_RELU = 0, 1, 2


def bias_act(x: torch.Tensor, bias: torch.Tensor | None, act: int = ACT_SILU, residual: torch.Tensor | None = None,
             out: torch.Tensor | None = None, out2: torch.Tensor | None = None) -> torch.Tensor:
    """out = act(x + bias[c]) (+ residual).  Inference-engine epilogue of a cuDNN conv; `out` defaults to `x` (in place)
    and may be a channel slice of a wider concat buffer.  With `out2`, the first out.shape[1] channels go to `out` and
    the remaining ones to `out2`."""
    _need_cuda(x)
    B, C, H, W = x.shape
    out = x if out is None else out
    if bias is not None and (bias.dtype != torch.float32 or bias.numel() != C or not bias.is_contiguous()):
        raise EdgelineError("bias_act: bias must be a contiguous fp32 vector of C elements")
    if residual is not None and (residual.shape != x.shape or residual.dtype != x.dtype):
        raise EdgelineError("bias_act: residual must match x")
    check(_lib.lib().el_bias_act_fwd(x.data_ptr(), _i64(x.stride()), bias.data_ptr() if bias is not None else None,
                                     residual.data_ptr() if residual is not None else None,
                                     _i64(residual.stride()) if residual is not None else None, out.data_ptr(), _i64(out.stride()),
                                     out2.data_ptr() if out2 is not None else None, _i64(out2.stride()) if out2 is not None else None,
                                     out.shape[1] if out2 is not None else 0, B, C, H, W, int(act), _dt(x), _stream()), "el_bias_act_fwd")
    return out


def pack_dw_weight(weight: torch.Tensor) -> torch.Tensor:
    """Depthwise Conv2d weight (C,1,k,k) -> fp32 (k*k, C) tap-major, the layout el_dwconv_fwd reads."""
    C, one, k, k2 = weight.shape
    if one != 1 or k != k2:
        raise EdgelineError("pack_dw_weight: need a depthwise (C,1,k,k) weight")
    return weight.detach().float().reshape(C, k * k).t().contiguous()


def dwconv(x: torch.Tensor, w_packed: torch.Tensor, k: int, bias: torch.Tensor | None = None, act: int = 0,
           out: torch.Tensor | None = None) -> torch.Tensor:
    """Depthwise k x k conv, stride 1, padding k//2, NHWC activations; optional fused bias + activation (DSConv.dw /
    DWConv of the reference, nn/modules/conv.py:87-112).  `w_packed` from pack_dw_weight."""
    _need_cuda(x, w_packed)
    B, C, H, W = x.shape
    if w_packed.dtype != torch.float32 or tuple(w_packed.shape) != (k * k, C) or not w_packed.is_contiguous():
        raise EdgelineError("dwconv: w_packed must be contiguous fp32 (k*k, C)")
    if bias is not None and (bias.dtype != torch.float32 or bias.numel() != C or not bias.is_contiguous()):
        raise EdgelineError("dwconv: bias must be a contiguous fp32 vector of C elements")
    if out is None:
        out = torch.empty((B, C, H, W), device=x.device, dtype=x.dtype, memory_format=torch.channels_last)
    check(_lib.lib().el_dwconv_fwd(x.data_ptr(), _i64(x.stride()), w_packed.data_ptr(), bias.data_ptr() if bias is not None else None,
                                   out.data_ptr(), _i64(out.stride()), B, C, H, W, int(k), int(act), _dt(x), _stream()), "el_dwconv_fwd")
    return out


def _pw_chunks(src_c):
    """K chunks of el_pwconv_fwd: per source, TMA boxes of 64 channels (16 / 32 when the whole source is that narrow).
    Returns [(first channel in the concatenated K, real channels, box channels)]."""
    chunks, base = [], 0
    for c in src_c:
        if c % 8:
            raise EdgelineError("pwconv: every source needs a multiple of 8 channels")
        bw = 16 if c <= 16 else (32 if c <= 32 else 64)
        for c0 in range(0, c, bw):
            chunks.append((base + c0, min(bw, c - c0), bw))
        base += c
    return chunks


def pack_pw_weight(weight: torch.Tensor, src_c, dtype=torch.bfloat16, M: int = 1 << 30, n_tile: int | None = None) -> torch.Tensor:
    """1x1 conv weight (N, K[,1,1]) with K = sum(src_c) -> the resident UMMA B operand of el_pwconv_fwd: per output-channel
    tile, one K-major tile [n_tile][box channels] per K chunk, stored with the 32/64/128-byte swizzle of its row width
    (16-byte chunk j of row r lands at chunk j ^ ((r * row_bytes >> 7) & (row_bytes / 16 - 1))), each tile padded to 1024 B.
    `M` = pixels of the call site: small maps use narrower output-channel tiles (el_pwconv_tile), so the packing depends on it."""
    w = weight.detach().reshape(weight.shape[0], -1).float()
    N, K = w.shape
    if K != sum(src_c):
        raise EdgelineError("pack_pw_weight: weight K does not match the sources")
    chunks = _pw_chunks(src_c)
    if n_tile is None:
        n_tile = _lib.lib().el_pwconv_tile(N, sum(2 * bw for _, _, bw in chunks), M)
    if n_tile <= 0:
        raise EdgelineError("pack_pw_weight: K too large for a resident weight tile")
    n_tiles = -(-N // n_tile)
    wp = torch.zeros(n_tiles * n_tile, K, device=w.device)
    wp[:N] = w
    rows = torch.arange(n_tile, device=w.device)
    out = []
    for t in range(n_tiles):
        for k0, creal, bw in chunks:
            tile = torch.zeros(n_tile, bw, device=w.device)
            tile[:, :creal] = wp[t * n_tile : (t + 1) * n_tile, k0 : k0 + creal]
            rb = 2 * bw
            f = ((rows * rb) >> 7) & (rb // 16 - 1)
            idx = torch.arange(bw // 8, device=w.device).view(1, -1) ^ f.view(-1, 1)       # dest[r, p] = tile[r, p ^ f(r)]
            sw = tile.view(n_tile, bw // 8, 8).gather(1, idx.view(n_tile, -1, 1).expand(-1, -1, 8)).reshape(-1).to(dtype)
            pad = (-sw.numel() * 2) % 1024 // 2
            out.append(torch.cat([sw, torch.zeros(pad, device=w.device, dtype=dtype)]) if pad else sw)
    return torch.cat(out).contiguous()


def pwconv(srcs, wpk: torch.Tensor, N: int, bias: torch.Tensor | None = None, act: int = ACT_NONE, residual: torch.Tensor | None = None,
           out: torch.Tensor | None = None, out2: torch.Tensor | None = None, res_scale: float = 1.0, up_addend: torch.Tensor | None = None) -> torch.Tensor:
    """1x1 conv over the channel-concatenation of `srcs` (NHWC 16-bit tensors or channel-slice views of the same B,H,W)
    + bias + activation (+ residual) in one tcgen05 GEMM kernel.  With `out2`, channels [0, out.shape[1]) go to `out`
    and the rest to `out2`.  With `residual`, out = residual + res_scale * act(conv + bias) (`out` may be `residual` itself).
    With `up_addend` (B, N, H/2, W/2), out = act(conv + bias + nearest-2x-upsample(up_addend)): the 1x1 conv over
    cat[Upsample(x_low), skip] with the x_low part of the weights applied at low resolution."""
    x0 = srcs[0]
    _need_cuda(*srcs, wpk)
    B, _, H, W = x0.shape
    M = B * H * W

    def pitch(t, what):
        if t.shape[0] != B or t.shape[2] != H or t.shape[3] != W or t.dtype != x0.dtype:
            raise EdgelineError(f"pwconv: {what} must share batch, size and dtype with the first source")
        sn, sc, sh, sw = t.stride()
        if sc != 1 or (W > 1 and sh != W * sw) or (H * W > 1 and B > 1 and sn != H * W * sw):
            raise EdgelineError(f"pwconv: {what} must be an NHWC tensor or a channel slice of one")
        return sw

    if out is None:
        out = torch.empty((B, N, H, W), device=x0.device, dtype=x0.dtype, memory_format=torch.channels_last)
    split = out.shape[1] if out2 is not None else 0
    if bias is not None and (bias.dtype != torch.float32 or bias.numel() != N or not bias.is_contiguous()):
        raise EdgelineError("pwconv: bias must be a contiguous fp32 vector of N elements")
    up_H = up_W = 0
    if up_addend is not None:
        if residual is not None or H % 2 or W % 2 or tuple(up_addend.shape) != (B, N, H // 2, W // 2) or up_addend.dtype != x0.dtype \
                or not up_addend.is_contiguous(memory_format=torch.channels_last):
            raise EdgelineError("pwconv: up_addend must be a dense NHWC (B, N, H/2, W/2) tensor and excludes `residual`")
        residual, up_H, up_W = up_addend, H, W
    n = len(srcs)
    check(_lib.lib().el_pwconv_fwd(n, _ptrs(srcs), _i64([pitch(t, "source") for t in srcs]), (c_int32 * n)(*[t.shape[1] for t in srcs]),
                                   wpk.data_ptr(), bias.data_ptr() if bias is not None else None,
                                   residual.data_ptr() if residual is not None else None,
                                   (residual.stride(3) if up_H else pitch(residual, "residual")) if residual is not None else 0,
                                   float(res_scale), up_H, up_W, out.data_ptr(), pitch(out, "out"), out2.data_ptr() if out2 is not None else None,
                                   pitch(out2, "out2") if out2 is not None else 0, split, M, N, int(act), _dt(x0), _stream()), "el_pwconv_fwd")
    return out


def dsconv3_ok(C: int, N: int) -> bool:
    """el_dsconv3_fwd covers this (C_in, N) site (C 16 / 32 or >= 64, N a multiple of 8 up to 256, pointwise weights resident in shared memory)."""
    return bool(_lib.lib().el_dsconv3_ok(int(C), int(N)))


def pack_dsconv3_weight(pw_weight: torch.Tensor, dtype=torch.bfloat16) -> torch.Tensor:
    """Pointwise weight (N, C[,1,1]) of a depthwise-separable pair -> the resident B operand of el_dsconv3_fwd: the el_pwconv_fwd packing for one
    source of C channels with a single output-channel tile of ceil16(N) rows."""
    N, C = pw_weight.shape[0], pw_weight.shape[1]
    return pack_pw_weight(pw_weight, [C], dtype=dtype, n_tile=-(-N // 16) * 16)


def dsconv3(x: torch.Tensor, dw_packed: torch.Tensor, wpk: torch.Tensor, N: int, bias: torch.Tensor | None = None, act: int = ACT_SILU,
            dw_bias: torch.Tensor | None = None, dw_act: int = ACT_NONE, out: torch.Tensor | None = None) -> torch.Tensor:
    """Depthwise 3x3 (stride 1, padding 1; + dw_bias + dw_act) -> pointwise 1x1 + bias + act in one kernel (DSConv.forward with k = 3,
    nn/modules/conv.py:100-104; DWConv -> Conv(1x1) of the class towers, nn/modules/head.py:66-71).  `dw_packed` from pack_dw_weight,
    `wpk` from pack_dsconv3_weight; NHWC 16-bit activations (channel-slice views allowed)."""
    _need_cuda(x, dw_packed, wpk)
    B, C, H, W = x.shape
    if dw_packed.dtype != torch.float32 or tuple(dw_packed.shape) != (9, C) or not dw_packed.is_contiguous():
        raise EdgelineError("dsconv3: dw_packed must be contiguous fp32 (9, C)")
    for name, b, n in (("bias", bias, N), ("dw_bias", dw_bias, C)):
        if b is not None and (b.dtype != torch.float32 or b.numel() != n or not b.is_contiguous()):
            raise EdgelineError(f"dsconv3: {name} must be a contiguous fp32 vector of {n} elements")
    if out is None:
        out = torch.empty((B, N, H, W), device=x.device, dtype=x.dtype, memory_format=torch.channels_last)
    check(_lib.lib().el_dsconv3_fwd(x.data_ptr(), _i64(x.stride()), C, dw_packed.data_ptr(), dw_bias.data_ptr() if dw_bias is not None else None, int(dw_act),
                                    wpk.data_ptr(), bias.data_ptr() if bias is not None else None, int(act), out.data_ptr(), _i64(out.stride()),
                                    B, H, W, N, _dt(x), _stream()), "el_dsconv3_fwd")
    return out


def conv3x3_tiles(N: int, C: int, B: int, H: int, W: int, stride: int = 1):
    """(n_tile, n_tiles) el_conv3x3_fwd uses for this site."""
    Ho, Wo = (H - 1) // stride + 1, (W - 1) // stride + 1
    n_tile = _lib.lib().el_conv3x3_tile(N, C, B * Ho * Wo)
    return n_tile, (-(-N // n_tile) if n_tile > 0 else 0)


def pack_conv3x3_weight(weight: torch.Tensor, dtype=torch.bfloat16, M: int = 1 << 30) -> torch.Tensor:
    """Conv2d weight (N, C, 3, 3) -> resident UMMA tiles of el_conv3x3_fwd: the (N, 9*C) matrix in (ky, kx, c) order, packed like a
    1x1 conv over nine C-channel sources (one per filter tap)."""
    N, C, kh, kw = weight.shape
    if (kh, kw) != (3, 3):
        raise EdgelineError("pack_conv3x3_weight: need a (N, C, 3, 3) weight")
    return pack_pw_weight(weight.detach().permute(0, 2, 3, 1).reshape(N, 9 * C), [C] * 9, dtype, M, n_tile=_lib.lib().el_conv3x3_tile(N, C, M))


def conv3x3(x: torch.Tensor, wpk: torch.Tensor, N: int, bias: torch.Tensor | None = None, act: int = ACT_NONE, stride: int = 1,
            out: torch.Tensor | None = None) -> torch.Tensor:
    """Dense 3x3 conv (padding 1, stride 1 / 2) + bias + activation as a TMA + tcgen05 implicit GEMM (NHWC 16-bit activations)."""
    _need_cuda(x, wpk)
    B, C, H, W = x.shape
    Ho, Wo = (H - 1) // stride + 1, (W - 1) // stride + 1
    if out is None:
        out = torch.empty((B, N, Ho, Wo), device=x.device, dtype=x.dtype, memory_format=torch.channels_last)
    if bias is not None and (bias.dtype != torch.float32 or bias.numel() != N or not bias.is_contiguous()):
        raise EdgelineError("conv3x3: bias must be a contiguous fp32 vector of N elements")
    check(_lib.lib().el_conv3x3_fwd(x.data_ptr(), _i64(x.stride()), C, wpk.data_ptr(), bias.data_ptr() if bias is not None else None, out.data_ptr(),
                                    _i64(out.stride()), B, H, W, N, int(stride), int(act), _dt(x), _stream()), "el_conv3x3_fwd")
    return out


def conv3x3_mma_ok(C: int, N: int, stride: int = 1) -> bool:
    """el_conv3x3_mma_fwd covers this site (C_in 16 / 32, N a multiple of 8 up to 64, stride 1 / 2)."""
    return bool(_lib.lib().el_conv3x3_mma_ok(int(C), int(N), int(stride)))


def conv3x3_mma(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor | None = None, act: int = ACT_NONE, stride: int = 1,
                out: torch.Tensor | None = None) -> torch.Tensor:
    """Dense 3x3 conv (padding 1, stride 1 / 2) + bias + activation for narrow channel counts: `weight` is the plain fp32 (N, C, 3, 3) tensor
    (contiguous, on the device); NHWC 16-bit activations; mma.sync from one haloed shared-memory tile (csrc/conv3x3_mma.cu)."""
    _need_cuda(x, weight)
    B, C, H, W = x.shape
    N = weight.shape[0]
    if tuple(weight.shape) != (N, C, 3, 3) or weight.dtype != torch.float32 or not weight.is_contiguous():
        raise EdgelineError("conv3x3_mma: weight must be a contiguous fp32 (N, C, 3, 3) tensor")
    if bias is not None and (bias.dtype != torch.float32 or bias.numel() != N or not bias.is_contiguous()):
        raise EdgelineError("conv3x3_mma: bias must be a contiguous fp32 vector of N elements")
    if out is None:
        out = torch.empty((B, N, (H - 1) // stride + 1, (W - 1) // stride + 1), device=x.device, dtype=x.dtype, memory_format=torch.channels_last)
    check(_lib.lib().el_conv3x3_mma_fwd(x.data_ptr(), _i64(x.stride()), C, weight.data_ptr(), bias.data_ptr() if bias is not None else None, out.data_ptr(),
                                        _i64(out.stride()), B, H, W, N, int(stride), int(act), _dt(x), _stream()), "el_conv3x3_mma_fwd")
    return out


def conv3x3_halo_ok(C: int, N: int) -> bool:
    """el_conv3x3_halo_fwd covers this (C_in, N) site (C_in a multiple of 64, the nine weight tiles resident in shared memory)."""
    return bool(_lib.lib().el_conv3x3_halo_ok(int(C), int(N)))


def pack_conv3x3_halo_weight(weight: torch.Tensor, dtype=torch.bfloat16) -> torch.Tensor:
    """Conv2d weight (N, C, 3, 3) -> resident UMMA B tiles of el_conv3x3_halo_fwd: [tap = ky*3+kx][64-channel chunk] K-major SW128 tiles of
    ceil16(N) rows (the packing of a 1x1 conv over nine C-channel sources with one output-channel tile)."""
    N, C, kh, kw = weight.shape
    if (kh, kw) != (3, 3) or C % 64:
        raise EdgelineError("pack_conv3x3_halo_weight: need a (N, C, 3, 3) weight with C a multiple of 64")
    return pack_pw_weight(weight.detach().permute(0, 2, 3, 1).reshape(N, 9 * C), [C] * 9, dtype, n_tile=-(-N // 16) * 16)


def conv3x3_halo(x: torch.Tensor, wpk: torch.Tensor, N: int, bias: torch.Tensor | None = None, act: int = ACT_NONE, out: torch.Tensor | None = None) -> torch.Tensor:
    """Dense 3x3 conv (padding 1, stride 1, wide C_in) + bias + activation from ONE haloed TMA tile per K chunk (NHWC 16-bit activations)."""
    _need_cuda(x, wpk)
    B, C, H, W = x.shape
    if out is None:
        out = torch.empty((B, N, H, W), device=x.device, dtype=x.dtype, memory_format=torch.channels_last)
    if bias is not None and (bias.dtype != torch.float32 or bias.numel() != N or not bias.is_contiguous()):
        raise EdgelineError("conv3x3_halo: bias must be a contiguous fp32 vector of N elements")
    check(_lib.lib().el_conv3x3_halo_fwd(x.data_ptr(), _i64(x.stride()), C, wpk.data_ptr(), bias.data_ptr() if bias is not None else None, out.data_ptr(),
                                         _i64(out.stride()), B, H, W, N, int(act), _dt(x), _stream()), "el_conv3x3_halo_fwd")
    return out


def upsample2x_cat(x: torch.Tensor, skip: torch.Tensor) -> torch.Tensor:
    """cat[nearest-2x(x), skip] along channels in one pass (NHWC activations)."""
    _need_cuda(x, skip)
    B, C1, h, w = x.shape
    C2, H, W = skip.shape[1:]
    if (H, W) != (2 * h, 2 * w) or skip.shape[0] != B or skip.dtype != x.dtype:
        raise EdgelineError("upsample2x_cat: skip must be (B, C2, 2h, 2w) of the same dtype")
    out = torch.empty((B, C1 + C2, H, W), device=x.device, dtype=x.dtype, memory_format=torch.channels_last)
    check(_lib.lib().el_upsample2x_cat_fwd(x.data_ptr(), _i64(x.stride()), skip.data_ptr(), _i64(skip.stride()), out.data_ptr(),
                                           _i64(out.stride()), B, C1, C2, H, W, _dt(x), _stream()), "el_upsample2x_cat_fwd")
    return out


def sppf_pool(x: torch.Tensor) -> torch.Tensor:
    """cat[x, m(x), m(m(x)), m(m(m(x)))] with m = MaxPool2d(5, 1, 2), in one kernel (NHWC, maps up to ~56x56)."""
    _need_cuda(x)
    B, C, H, W = x.shape
    out = torch.empty((B, 4 * C, H, W), device=x.device, dtype=x.dtype, memory_format=torch.channels_last)
    check(_lib.lib().el_sppf_pool_fwd(x.data_ptr(), _i64(x.stride()), out.data_ptr(), _i64(out.stride()), B, C, H, W, _dt(x), _stream()),
          "el_sppf_pool_fwd")
    return out
