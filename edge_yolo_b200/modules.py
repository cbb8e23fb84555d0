"""Host-side mirror of the reference's operator interface for the EdgeLine hot path.

Two things live here:

1. `*_forward` functions: the B200 bodies of the reference's module forwards.  They only rely on
   attribute names the reference classes also have (`f_ll`, `f_h`, `fuse`, `alpha`, `gamma`, `qkv`,
   `proj`, `cv2`, `cv3`, `reg_conf`, `stride`, ...), so `install.py` can bind them onto the
   reference's own classes (drop-in behind the ultralytics API), and the standalone classes
   below use the very same functions.
2. Standalone `nn.Module`s with the reference's constructor signatures and state-dict keys
    so that EdgeLine-YOLO can be built and benchmarked on a machine without
   ultralytics.  Convolutions / BatchNorm / SiLU stay on PyTorch + cuDNN (north_star).

Reference paths are relative to /root/reference/ultralytics/.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops

# ------------------------------------------------------------------------------------------
# shared forward bodies (bound onto reference classes by install.py)
# ------------------------------------------------------------------------------------------


def _ver(p: torch.Tensor) -> int:
    """`p._version`, or -1 for inference tensors (no version counter; they cannot be modified in place outside inference mode, and
    the caches below also key on the storage pointer).  The reference's validator loads and fuses checkpoints under
    `torch.inference_mode` (engine/validator.py:108, utils/torch_utils.py:62-72), so its parameters are inference tensors."""
    try:
        return p._version
    except RuntimeError:
        return -1


def dwt_forward(self, x: torch.Tensor):
    """`_PywtDWT2D.forward` (nn/modules/block.py:3619-3642): returns (LL, LH, HL, HH)."""
    if getattr(self, "wave_name", "haar") != "haar":
        raise NotImplementedError("edge_yolo_b200 implements the Haar wavelet only (the only one any EdgeLine yaml uses)")
    buf = ops.dwt_haar(x)
    B = x.shape[0]
    return buf[:B], buf[B : 2 * B], buf[2 * B : 3 * B], buf[3 * B :]


def wavelet_enhancer_forward(self, b: torch.Tensor, inplace: bool = False, out2: torch.Tensor | None = None) -> torch.Tensor:
    """`_WaveletEnhancer.forward` (block.py:3685-3710).

    DWT split -> f_ll / shared f_h convs (cuDNN) -> fused upsample*w+concat kernel -> fuse conv ->
    gated residual kernel.  In eval mode the three high bands go through `f_h` as one 3B batch
    (BatchNorm uses running statistics, so this is exactly three separate calls); in training
    mode they are three calls like the reference, because batch statistics differ per call.
    """
    B = b.shape[0]
    buf = ops.dwt_haar(b)
    LLp = self.f_ll(buf[:B])
    if self.training:
        LHp, HLp, HHp = self.f_h(buf[B : 2 * B]), self.f_h(buf[2 * B : 3 * B]), self.f_h(buf[3 * B :])
    else:
        hp = self.f_h(buf[B:])
        LHp, HLp, HHp = hp[:B], hp[B : 2 * B], hp[2 * B :]
    cat = ops.wave_merge(b, LLp, LHp, HLp, HHp, self.alpha)
    y = self.fuse(cat)
    return ops.gated_residual(b, y, self.gamma, inplace=inplace and not torch.is_grad_enabled(), out2=out2)


def wavelet_enhancer_engine_forward(self, b: torch.Tensor) -> torch.Tensor:
    """`_WaveletEnhancer.forward` (block.py:3685-3710) for the inference engine (in place on b where that is race-free): DWT split -> f_ll (tcgen05 1x1) /
    shared f_h (cuDNN 3x3) -> bands-only upsample*w kernel -> ONE GEMM for `fuse` that reads [b | upsampled bands] in place
    (no concat of b) and applies bias + SiLU + the gated residual b + tanh(gamma) * y in its epilogue."""
    fuse = self.fuse
    B, c, H, W = b.shape
    n_tiles = _pw_n_tiles(fuse.conv, (c, 2 * c), B * H * W) if c % 16 == 0 else 99  # of the GEMM that really runs: K = [b | 2c upsampled bands]
    if not (hasattr(fuse, "el_bias") and b.dtype in (torch.bfloat16, torch.float16) and _pixel_linear(b) and c % 16 == 0
            and H % 2 == 0 and W % 2 == 0 and _pw_ok(fuse.conv, [b]) and n_tiles <= 4):
        return wavelet_enhancer_forward(self, b, inplace=True)
    buf = ops.dwt_haar(b)
    LLp = self.f_ll(buf[:B])
    hp = self.f_h(buf[B:])
    a32 = self.__dict__.get("el_alpha32")  # fp32 copy of alpha for the kernel, cached until alpha is modified (a 16-bit model would
    if a32 is None or a32[0] != (_ver(self.alpha), self.alpha.data_ptr()) or a32[1].device != self.alpha.device:  # otherwise launch a cast kernel per call)
        a32 = self.el_alpha32 = ((_ver(self.alpha), self.alpha.data_ptr()), self.alpha.detach().float().contiguous())
    U = ops.wave_merge_bands(LLp, hp[:B], hp[B : 2 * B], hp[2 * B :], a32[1], H, W)
    gate = self.__dict__.get("el_gate")
    if gate is None or gate[0] != (_ver(self.gamma), self.gamma.data_ptr()):  # tanh(gamma) as a host scalar, cached until gamma is modified
        gate = self.el_gate = ((_ver(self.gamma), self.gamma.data_ptr()), float(torch.tanh(self.gamma.detach().float())))
    # In place over b only when ONE CTA column owns all output channels of a pixel tile.  With N split over several CTAs, CTA (m, 0) would
    # store its channels of b while CTA (m, 1) may still be loading all of b as its A operand (el_pwconv_fwd refuses that aliasing).
    out = b if n_tiles == 1 else torch.empty_like(b)
    return pw_apply(fuse.conv, [b, U], _bias_on(fuse, b), fuse.el_act, out=out, residual=b, res_scale=gate[1])


def linear_attention_forward(self, x: torch.Tensor) -> torch.Tensor:
    """`LinearAttention.forward` (block.py:3360-3373): qkv conv -> fused attention core -> proj conv."""
    return self.proj(ops.linear_attention(self.qkv(x), self.num_heads))


def _dgqp_weights(self):
    """fp32 views of the DGQP heads `reg_conf[i]` = Conv2d(20,64,1) ReLU Conv2d(64,1,1) Sigmoid (head.py:847-854)."""
    key = tuple((p.data_ptr(), _ver(p), p.dtype, p.device) for seq in self.reg_conf for p in seq.parameters())
    cache = getattr(self, "_el_dgqp", None)
    if cache is None or cache[0] != key:
        ws = []
        for seq in self.reg_conf:
            c1, c2 = seq[0], seq[2]
            if c1.weight.shape[1] != 20 or c1.weight.shape[0] != 64 or c2.weight.shape[:2] != (1, 64):
                raise NotImplementedError("edge_yolo_b200 decode supports reg_topk=4, add_mean=True, reg_channels=64 (the yaml's values)")
            ws.append(tuple(t.detach().float().reshape(-1).contiguous() for t in (c1.weight, c1.bias, c2.weight, c2.bias)))
        cache = (key, ws)
        self._el_dgqp = cache
    return cache[1]


def gfl_head_forward(self, x):
    """`GFLHeadv2_uniH.forward` (nn/modules/head.py:880-908).

    Training: returns the per-level cat(box, cls) list like the reference (whose quality maps are
    computed and dropped on this path).  Eval: one fused decode kernel produces
    y (B, 4+nc, A) in fp32; returns `y` when exporting, else `(y, x)`.
    """
    def tower_in(i):
        xi = x[i]
        for name in ("stem", "dat", "pos_cls", "pos_reg", "cit_cls", "cit_reg"):  # Identity placeholders in the reference
            mods = getattr(self, name, None)
            if mods:
                xi = mods[i](xi)
        return xi

    def towers(i):
        xi = tower_in(i)
        return self.cv2[i](xi), self.cv3[i](xi)

    if getattr(self, "el_level_streams", False) and x[0].is_cuda and not self.training:
        # engine: the box and class towers of the three pyramid levels are six independent chains until the decode, and the
        # 40x40 / 20x20 ones are short kernels that cannot fill the GPU -- run them as parallel branches (fork / join on side
        # streams; captured into the CUDA graph as a fork-join)
        main = torch.cuda.current_stream(x[0].device)
        side = self.__dict__.setdefault("_el_streams", [torch.cuda.Stream(device=x[0].device) for _ in range(2 * self.nl - 1)])
        fork = torch.cuda.Event()
        fork.record(main)
        boxes, clss = [None] * self.nl, [None] * self.nl
        joins = []
        jobs = [(i, t) for i in range(self.nl) for t in (0, 1)][1:]  # (level, tower); job (0, box tower) stays on the main stream
        for st, (i, t) in zip(side, jobs):
            st.wait_event(fork)
            with torch.cuda.stream(st):
                if t == 0:
                    boxes[i] = self.cv2[i](tower_in(i))
                else:
                    clss[i] = self.cv3[i](tower_in(i))
                ev = torch.cuda.Event()
                ev.record(st)
                joins.append(ev)
        boxes[0] = self.cv2[0](tower_in(0))
        for ev in joins:
            main.wait_event(ev)
    else:
        boxes, clss = [], []
        for i in range(self.nl):
            bi, ci = towers(i)
            boxes.append(bi)
            clss.append(ci)
    if self.training:
        for i in range(self.nl):
            x[i] = torch.cat((boxes[i], clss[i]), 1)
        return x
    if getattr(self, "export", False) and getattr(self, "format", None) in {"tflite", "edgetpu", "imx", "saved_model", "pb", "tfjs"}:
        raise NotImplementedError("edge_yolo_b200: export formats are out of scope (no multi-backend dispatch)")
    bias = getattr(self, "el_head_bias", None)  # engine fuse: biases of the towers' last convs, added inside the decode kernels
    if bias is not None and bias[0][0].device != boxes[0].device:
        bias = self.el_head_bias = tuple([t.to(boxes[0].device) for t in side] for side in bias)
    det = getattr(self, "el_detect", None)
    if det is not None:  # engine path (Predictor): fused decode + NMS, returns (rows (B, max_det, 6), counts (B))
        split = getattr(self, "el_detect_split", None)
        if split is None:
            return ops.gfl_detect(boxes, clss, _dgqp_weights(self), [float(s) for s in self.stride], bias=bias, **det)
        # pipelined engine: this call only decodes + emits candidate keys into the caller's workspace; `split["finish"]()` runs the
        # sort + sweep on the same buffers (captured as a second CUDA graph that overlaps the next batch's forward)
        args = (boxes, clss, _dgqp_weights(self), [float(s) for s in self.stride])
        kw = dict(det, bias=bias, workspace=split["workspace"], out=split["out"], cnt=split["cnt"])
        if split.get("defer_decode"):  # the decode + emit kernel moves into the second graph as well
            split["finish"] = lambda: ops.gfl_detect(*args, stages=7, **kw)
            return split["out"], split["cnt"]
        split["finish"] = lambda: ops.gfl_detect(*args, stages=6, **kw)
        return ops.gfl_detect(*args, stages=1, **kw)
    y = ops.gfl_decode(boxes, clss, _dgqp_weights(self), [float(s) for s in self.stride], bias=bias)
    if getattr(self, "export", False):
        return y
    if getattr(self, "el_skip_feats", False):  # fast predict path: nobody reads the raw maps
        return y, None
    for i in range(self.nl):
        if bias is not None:  # engine-fused towers: put the stripped biases back into the maps handed to the caller
            x[i] = torch.cat((boxes[i] + bias[0][i].view(1, -1, 1, 1).to(boxes[i].dtype), clss[i] + bias[1][i].view(1, -1, 1, 1).to(clss[i].dtype)), 1)
        else:
            x[i] = torch.cat((boxes[i], clss[i]), 1)
    return y, x


# ------------------------------------------------------------------------------------------
# inference-engine forwards (bound by EdgeLineYOLO.fuse(engine=True); eval only, NHWC activations)
#
# Same maths as the reference modules after BaseModel.fuse() (nn/tasks.py:214-242), but every conv
# epilogue (folded-BN bias + SiLU, shortcut add) is one el_bias_act kernel that can write straight into
# a channel slice of the block's concat buffer, so torch.cat / broadcast-add / SiLU launches vanish.
# ------------------------------------------------------------------------------------------


USE_DWCONV = True  # el_dwconv_fwd in the engine graph (model._dw_eligible picks the sites where it wins)
DWCONV_K3 = True   # also the bare k = 3 depthwise of DSConv (on par with PyTorch in isolation; ours chains with PDL)
DWCONV_ANY_C = True  # k = 3 depthwise with any multiple of 8 channels >= 16 (the 80-channel class towers) on dwconv3_tma_kernel


def _bias_on(self, x):
    b = self.el_bias
    if b is not None and b.device != x.device:
        b = self.el_bias = b.to(x.device)
    return b


def _dw_on(self, x):
    w = getattr(self, "el_dw", None)
    if w is not None and w.device != x.device:
        w = self.el_dw = w.to(x.device)
    return w


def _pixel_linear(t: torch.Tensor) -> bool:
    B, C, H, W = t.shape
    sn, sc, sh, sw = t.stride()
    return sc == 1 and C % 8 == 0 and sw % 8 == 0 and (W == 1 or sh == W * sw) and (B == 1 or H * W == 1 or sn == H * W * sw) \
        and t.data_ptr() % 16 == 0


def _pw_ok(conv: nn.Conv2d, srcs, *others) -> bool:
    """el_pwconv_fwd applies: 1x1 / stride 1 / dense conv, 16-bit NHWC (or channel-slice) operands, <= 4 sources, no autograd."""
    if not (conv.kernel_size == (1, 1) and conv.stride == (1, 1) and conv.groups == 1 and conv.padding in ((0, 0), 0) and conv.out_channels % 8 == 0
            and len(srcs) <= 4 and srcs[0].is_cuda and srcs[0].dtype in (torch.bfloat16, torch.float16) and not torch.is_grad_enabled()
            and all(_pixel_linear(t) for t in srcs) and all(t is None or _pixel_linear(t) for t in others)):
        return False
    # the weight block of one output-channel tile stays resident in shared memory (<= 128 KB): very wide K splits N into many tiles,
    # each of which re-reads the activations -- beyond 4 tiles cuDNN is the better choice
    x0 = srcs[0]
    return _pw_n_tiles(conv, tuple(t.shape[1] for t in srcs), x0.shape[0] * x0.shape[2] * x0.shape[3]) <= 4


def _pw_n_tiles(conv: nn.Conv2d, src_channels: tuple, M: int) -> int:
    """Output-channel tiles el_pwconv_fwd uses for this (source split, pixel count): the el_pwconv_tile rule, cached on the conv."""
    key = (tuple(src_channels), M)
    cache = conv.__dict__.setdefault("el_pw_tiles", {})
    n_tiles = cache.get(key)
    if n_tiles is None:
        row_bytes = sum(2 * bw for _, _, bw in ops._pw_chunks(key[0]))
        n_tile = ops._lib.lib().el_pwconv_tile(conv.out_channels, row_bytes, key[1])
        n_tiles = cache[key] = (-(-conv.out_channels // n_tile) if n_tile > 0 else 99)
    return n_tiles


USE_CONV3X3 = True  # el_conv3x3_fwd for the dense 3x3 convs of the engine graph


def _conv3_ok(conv: nn.Conv2d, x: torch.Tensor, out=None) -> bool:
    """el_conv3x3_fwd applies: dense 3x3, padding 1, stride 1 / 2, 16-bit NHWC, channel counts in 8s -- and pays off: measured on B200
    (tools/prof_conv.py c3) the nine tap-shifted TMA boxes re-read the input 9x from L2, which beats cuDNN + el_bias_act for
    C_in <= 32 (HBM-bound sites: 16->32 s2 @320^2 89 vs 185 us, 32->64 s2 @160^2 50 vs 70 us) and loses to cuDNN's smem-reusing
    implicit GEMM from C_in = 64 up, so wider convs stay on cuDNN."""
    if not (USE_CONV3X3 and conv.kernel_size == (3, 3) and conv.padding == (1, 1) and conv.stride in ((1, 1), (2, 2)) and conv.dilation == (1, 1)
            and conv.groups == 1 and x.is_cuda and x.dtype in (torch.bfloat16, torch.float16) and not torch.is_grad_enabled()):
        return False
    B, C, H, W = x.shape
    N = conv.out_channels
    if C % 8 or N % 8 or C > 32 or x.stride(1) != 1 or any(s % 8 for i, s in enumerate(x.stride()) if i != 1) or x.data_ptr() % 16:
        return False
    if out is not None and (out.stride(1) != 1 or any(s % 8 for i, s in enumerate(out.stride()) if i != 1) or out.data_ptr() % 16):
        return False
    n_tile, n_tiles = ops.conv3x3_tiles(N, C, B, H, W, conv.stride[0])
    return 0 < n_tiles <= 3


USE_CONV3X3_MMA = True  # el_conv3x3_mma_fwd for the narrow (C_in 16 / 32, N <= 64) dense 3x3 convs, stride 1 / 2 (f_h on the early maps, layer 1)


def _conv3_mma_ok(conv: nn.Conv2d, x: torch.Tensor, out=None) -> bool:
    """el_conv3x3_mma_fwd applies: dense 3x3, padding 1, stride 1 / 2, 16-bit NHWC, C_in 16 / 32, N <= 64."""
    if not (USE_CONV3X3_MMA and conv.kernel_size == (3, 3) and conv.padding == (1, 1) and conv.stride in ((1, 1), (2, 2)) and conv.dilation == (1, 1)
            and conv.groups == 1 and x.is_cuda and x.dtype in (torch.bfloat16, torch.float16) and not torch.is_grad_enabled()):
        return False
    if conv.stride != (1, 1):  # measured (B = 64, 16 -> 32 stride 2 at 320 x 320): 117 us against 89 us on el_conv3x3_fwd -- layer 1 stays there
        return False
    if not ops.conv3x3_mma_ok(x.shape[1], conv.out_channels, conv.stride[0]) or x.stride(1) != 1 or any(s % 8 for i, s in enumerate(x.stride()) if i != 1) or x.data_ptr() % 16:
        return False
    if out is not None and (out.stride(1) != 1 or any(s % 2 for i, s in enumerate(out.stride()) if i != 1) or out.data_ptr() % 4):
        return False
    return True


USE_CONV3X3_HALO = True  # el_conv3x3_halo_fwd for the wide (C_in % 64 == 0) stride-1 dense 3x3 convs of the engine graph


def _conv3_halo_ok(conv: nn.Conv2d, x: torch.Tensor, out=None) -> bool:
    """el_conv3x3_halo_fwd applies: dense 3x3, padding 1, stride 1, 16-bit NHWC, C_in a multiple of 64, weights resident."""
    if not (USE_CONV3X3_HALO and conv.kernel_size == (3, 3) and conv.padding == (1, 1) and conv.stride == (1, 1) and conv.dilation == (1, 1)
            and conv.groups == 1 and x.is_cuda and x.dtype in (torch.bfloat16, torch.float16) and not torch.is_grad_enabled()):
        return False
    C, N = x.shape[1], conv.out_channels
    if C % 64 or N % 8 or x.stride(1) != 1 or any(s % 8 for i, s in enumerate(x.stride()) if i != 1) or x.data_ptr() % 16:
        return False
    if out is not None and (out.stride(1) != 1 or any(s % 8 for i, s in enumerate(out.stride()) if i != 1) or out.data_ptr() % 16):
        return False
    return ops.conv3x3_halo_ok(C, N)


def pw_apply(conv: nn.Conv2d, srcs, bias, act, out=None, residual=None, out2=None, res_scale=1.0, up_addend=None):
    """1x1 conv over the channel concatenation of `srcs` + bias + activation (+ residual) as one tcgen05 GEMM (ops.pwconv).
    Weight tiles are packed once per (conv, source split) and cached on the conv module."""
    cache = conv.__dict__.setdefault("el_wpk", {})
    x0 = srcs[0]
    M = x0.shape[0] * x0.shape[2] * x0.shape[3]
    # the tile split depends on M only for small maps; weight version + storage pointer: an in-place update or a load_state_dict after
    # fuse(engine=True) repacks instead of silently using stale tiles
    key = (tuple(t.shape[1] for t in srcs), x0.dtype, x0.device, M if M < (1 << 16) else 0, _ver(conv.weight), conv.weight.data_ptr())
    wpk = cache.get(key)
    if wpk is None:
        for k in [k for k in cache if k[:4] == key[:4]]:
            del cache[k]  # superseded packing of the same site
        wpk = cache[key] = ops.pack_pw_weight(conv.weight, key[0], key[1], M).to(key[2])
    return ops.pwconv(srcs, wpk, conv.out_channels, bias=bias, act=act, residual=residual, out=out, out2=out2, res_scale=res_scale,
                      up_addend=up_addend)


def conv2d_pw_forward(self, x):
    """Bare nn.Conv2d 1x1 (the last conv of the Detect towers, head.py:59-74; its bias lives in the decode kernel)."""
    if _pw_ok(self, [x]):
        b = self.__dict__.get("el_bias32")
        if b is None and self.bias is not None:
            b = self.el_bias32 = self.bias.detach().float().contiguous()
        return pw_apply(self, [x], b, ops.ACT_NONE)
    return F.conv2d(x, self.weight, self.bias)


class UpCat(tuple):
    """(x_low, skip): lazy cat[nearest-2x(x_low), skip] handed from an engine-mode Concat to the cv1 of the following block."""


def upcat_cv1(cv1, uc: UpCat, out, out2):
    """cv1(cat[Upsample(x_low), skip]) for a fused 1x1 Conv: a 1x1 conv commutes with nearest upsampling, so the x_low part of the
    weights is applied at LOW resolution (a quarter of the pixels) and its result enters the skip part's GEMM as a pre-activation
    addend that the epilogue reads at (y/2, x/2).  Neither the upsampled nor the concatenated tensor exists."""
    x_low, skip = uc
    conv = cv1.conv
    C1 = x_low.shape[1]
    halves = conv.__dict__.get("el_up_halves")
    if halves is not None and halves[2] != (_ver(conv.weight), conv.weight.data_ptr()):
        halves = None  # cv1's weight was modified after the split was cached
    if halves is None:
        lo, sk = nn.Conv2d(C1, conv.out_channels, 1, bias=False), nn.Conv2d(skip.shape[1], conv.out_channels, 1, bias=False)
        lo.weight, sk.weight = nn.Parameter(conv.weight[:, :C1].detach().clone(), False), nn.Parameter(conv.weight[:, C1:].detach().clone(), False)
        halves = conv.el_up_halves = (lo, sk, (_ver(conv.weight), conv.weight.data_ptr()))
    z = pw_apply(halves[0], [x_low], None, ops.ACT_NONE)
    return pw_apply(halves[1], [skip], _bias_on(cv1, skip), cv1.el_act, out=out, out2=out2, up_addend=z)


def _as_list(x):
    return list(x) if isinstance(x, (list, tuple)) else [x]


def conv_engine_forward(self, x, out=None, residual=None, out2=None):
    """Conv.forward_fuse (conv.py:58-60).  1x1 convs: one tcgen05 GEMM with the bias + activation [+ residual] epilogue and,
    for a list input, the concat folded into the K loop; other convs: cuDNN (no bias) -> el_bias_act epilogue."""
    srcs = _as_list(x)
    if _pw_ok(self.conv, srcs, out, residual, out2):
        return pw_apply(self.conv, srcs, _bias_on(self, srcs[0]), self.el_act, out=out, residual=residual, out2=out2)
    x = srcs[0] if len(srcs) == 1 else torch.cat(srcs, 1)
    if residual is None and out2 is None and _conv3_halo_ok(self.conv, x, out):
        cache = self.conv.__dict__.setdefault("el_wpk", {})
        key = ("3x3halo", x.dtype, x.device, 0, _ver(self.conv.weight), self.conv.weight.data_ptr())
        wpk = cache.get(key)
        if wpk is None:
            wpk = cache[key] = ops.pack_conv3x3_halo_weight(self.conv.weight, x.dtype).to(x.device)
        return ops.conv3x3_halo(x, wpk, self.conv.out_channels, bias=_bias_on(self, x), act=self.el_act, out=out)
    if residual is None and out2 is None and _conv3_mma_ok(self.conv, x, out):
        w32 = self.conv.__dict__.get("el_w32")
        if w32 is None or w32[0] != (_ver(self.conv.weight), self.conv.weight.data_ptr(), x.device):
            w32 = self.conv.el_w32 = ((_ver(self.conv.weight), self.conv.weight.data_ptr(), x.device), self.conv.weight.detach().float().contiguous().to(x.device))
        return ops.conv3x3_mma(x, w32[1], bias=_bias_on(self, x), act=self.el_act, stride=self.conv.stride[0], out=out)
    if residual is None and out2 is None and _conv3_ok(self.conv, x, out):
        cache = self.conv.__dict__.setdefault("el_wpk", {})
        B, C, H, W = x.shape
        s = self.conv.stride[0]
        M = B * ((H - 1) // s + 1) * ((W - 1) // s + 1)
        key = ("3x3", x.dtype, x.device, M if M < (1 << 16) else 0)
        wpk = cache.get(key)
        if wpk is None:
            wpk = cache[key] = ops.pack_conv3x3_weight(self.conv.weight, x.dtype, M).to(x.device)
        return ops.conv3x3(x, wpk, self.conv.out_channels, bias=_bias_on(self, x), act=self.el_act, stride=s, out=out)
    return ops.bias_act(self.conv(x), _bias_on(self, x), self.el_act, residual=residual, out=out, out2=out2)


def dwconv_engine_forward(self, x, out=None, residual=None, out2=None):
    """DWConv.forward_fuse (conv.py:124-130): depthwise conv + folded-BN bias + activation in one kernel."""
    if residual is not None or out2 is not None:
        return conv_engine_forward(self, x, out=out, residual=residual, out2=out2)
    return ops.dwconv(x, _dw_on(self, x), self.el_k, bias=_bias_on(self, x), act=self.el_act, out=out)


USE_DSCONV3 = True  # el_dsconv3_fwd: depthwise 3x3 -> pointwise 1x1 in one kernel (DSConv k = 3; DWConv -> Conv pairs of the class towers)
DSCONV3_MIN_C = 32  # narrower sites stay on the two-kernel path
DSCONV3_MAX_HW = int(__import__("os").environ.get("EL_DS3_MAX_HW", 1600))  # 16 / 32-channel sites: only maps up to 40 x 40, where one launch instead of two is
# what pays; from 64 channels up the fused kernel wins or ties at every map (tools/prof_dsconv.py, profiles/r02f_prof_dsconv.jsonl)
DSCONV3_WIDE_C = int(__import__("os").environ.get("EL_DS3_WIDE_C", 64))


def _ds3_ok(dw: nn.Conv2d, pw: nn.Conv2d, x: torch.Tensor, out=None) -> bool:
    """el_dsconv3_fwd applies: depthwise 3x3 / stride 1 / padding 1 into a dense 1x1 conv, 16-bit NHWC views, no autograd."""
    C, N = dw.in_channels, pw.out_channels
    if not (USE_DSCONV3 and C >= DSCONV3_MIN_C and (C >= DSCONV3_WIDE_C or x.shape[2] * x.shape[3] <= DSCONV3_MAX_HW) and dw.kernel_size == (3, 3) and dw.stride == (1, 1) and dw.padding == (1, 1) and dw.dilation == (1, 1)
            and dw.groups == C == dw.out_channels and pw.kernel_size == (1, 1) and pw.stride == (1, 1) and pw.groups == 1 and pw.padding in ((0, 0), 0)
            and pw.in_channels == C and x.is_cuda and x.dtype in (torch.bfloat16, torch.float16) and not torch.is_grad_enabled()):
        return False
    for t in (x, out):
        if t is not None and (t.stride(1) != 1 or any(st % 8 for i, st in enumerate(t.stride()) if i != 1) or t.data_ptr() % 16):
            return False
    return ops.dsconv3_ok(C, N)


def ds3_apply(dw_packed, pw: nn.Conv2d, x, bias, act, dw_bias=None, dw_act=ops.ACT_NONE, out=None):
    """Fused depthwise 3x3 -> 1x1 (ops.dsconv3); the pointwise weight tile is packed once per (dtype, device, weight version) and cached on the conv."""
    key = (x.dtype, x.device, _ver(pw.weight), pw.weight.data_ptr())
    c = pw.__dict__.get("el_ds3_wpk")
    if c is None or c[0] != key:
        c = pw.el_ds3_wpk = (key, ops.pack_dsconv3_weight(pw.weight, x.dtype).to(x.device))
    return ops.dsconv3(x, dw_packed, c[1], pw.out_channels, bias=bias, act=act, dw_bias=dw_bias, dw_act=dw_act, out=out)


def dwpw_engine_forward(self, x):
    """Sequential(DWConv(x, x, 3), Conv(x, c, 1)) of the class towers (head.py:66-71) as one kernel: depthwise + bias + SiLU feeds the
    1x1 GEMM from shared memory."""
    dwm, pwm = self[0], self[1]
    w = _dw_on(dwm, x)
    if w is not None and dwm.el_k == 3 and _ds3_ok(dwm.conv, pwm.conv, x):
        return ds3_apply(w, pwm.conv, x, _bias_on(pwm, x), pwm.el_act, dw_bias=_bias_on(dwm, x), dw_act=dwm.el_act)
    return pwm(dwm(x))


def dsconv_engine_forward(self, x, out=None, residual=None, out2=None):
    """DSConv.forward (conv.py:100-104) with its BatchNorm folded into the pointwise conv."""
    w = _dw_on(self, x)
    if w is not None and self.el_k == 3 and residual is None and out2 is None and _ds3_ok(self.dw, self.pw, x, out):
        return ds3_apply(w, self.pw, x, _bias_on(self, x), ops.ACT_SILU, out=out)
    d = ops.dwconv(x, w, self.el_k) if w is not None else self.dw(x)
    if _pw_ok(self.pw, [d], out, residual, out2):
        return pw_apply(self.pw, [d], _bias_on(self, x), ops.ACT_SILU, out=out, residual=residual, out2=out2)
    return ops.bias_act(self.pw(d), _bias_on(self, x), ops.ACT_SILU, residual=residual, out=out, out2=out2)


def dsbottleneck_engine_forward(self, x, out=None):
    """DSBottleneck.forward (block.py:1500-1503): the shortcut add rides in cv2's epilogue."""
    return self.cv2(self.cv1(x), out=out, residual=x if self.add else None)


def _c3_cv12(self, x):
    """cv1 and cv2 of a C3-style block read the same input: ONE GEMM over the stacked output channels with a split store (one launch
    less per block on the 20 x 20 / 40 x 40 maps where a launch costs what the GEMM does).  Returns (conv, bias, act) or None."""
    c1, c2 = self.cv1, self.cv2
    if not (isinstance(x, torch.Tensor) and hasattr(c1, "el_bias") and hasattr(c2, "el_bias") and c1.el_act == c2.el_act
            and c1.conv.kernel_size == c2.conv.kernel_size == (1, 1) and c1.conv.out_channels % 16 == 0 and c1.conv.in_channels == c2.conv.in_channels
            and c1.conv.bias is None and c2.conv.bias is None):
        return None
    key = (_ver(c1.conv.weight), c1.conv.weight.data_ptr(), _ver(c2.conv.weight), c2.conv.weight.data_ptr(), x.device)
    c = self.__dict__.get("el_cv12")
    if c is None or c[0] != key:
        conv = nn.Conv2d(c1.conv.in_channels, c1.conv.out_channels + c2.conv.out_channels, 1, bias=False).to(device=x.device, dtype=c1.conv.weight.dtype)
        with torch.no_grad():
            conv.weight.copy_(torch.cat([c1.conv.weight, c2.conv.weight], 0))
        bias = torch.cat([_bias_on(c1, x), _bias_on(c2, x)]).contiguous()
        c = (key, conv.requires_grad_(False), bias)
        self.__dict__["el_cv12"] = c  # not a registered submodule: the state dict keeps the reference's keys
    return c[1], c[2], c1.el_act


def dsc3k_engine_forward(self, x, out=None):
    """C3.forward (block.py:394-396) for DSC3k: cv1 | cv2 as one GEMM with a split store, cv3 reads both branches in place (concat folded
    into its K loop)."""
    m12 = _c3_cv12(self, x)
    if m12 is not None and _pw_ok(m12[0], [x]):
        B, _, H, W = x.shape
        cur = torch.empty((B, self.cv1.conv.out_channels, H, W), device=x.device, dtype=x.dtype, memory_format=torch.channels_last)
        other = torch.empty((B, self.cv2.conv.out_channels, H, W), device=x.device, dtype=x.dtype, memory_format=torch.channels_last)
        pw_apply(m12[0], [x], m12[1], m12[2], out=cur, out2=other)
    else:
        cur, other = self.cv1(x), self.cv2(x)
    for blk in self.m:
        cur = blk(cur)
    return self.cv3([cur, other], out=out)


def dsc3k2_wavelet_engine_forward(self, x):
    """DSC3K2_Wavelet.forward (block.py:3783-3788): cv1 writes its two halves a | b as dense tensors (chunk(2, 1) is the
    epilogue's split), the enhancer updates b in place, and cv2 reads a, b', m1(b'), ... in place: no concat buffer.
    `x` may be the list of a lazily concatenated input (Concat in engine mode)."""
    x0 = x[1] if isinstance(x, UpCat) else _as_list(x)[0]
    B, _, H, W = x0.shape
    c = self.c
    a = torch.empty((B, c, H, W), device=x0.device, dtype=x0.dtype, memory_format=torch.channels_last)
    b = torch.empty_like(a)
    if isinstance(x, UpCat):
        upcat_cv1(self.cv1, x, a, b)
    else:
        self.cv1(x, out=a, out2=b)
    cur = wavelet_enhancer_engine_forward(self.wave, b)
    ys = [a, cur]
    for blk in self.m:
        cur = blk(cur)
        ys.append(cur)
    return self.cv2(ys)


def psablock_engine_forward(self, x, out=None):
    """PSABlock_LinearAttention.forward (block.py:3446-3449) with both residual adds fused into GEMM epilogues."""
    att = self.attn
    if "el_bias32" not in att.qkv.__dict__:
        att.qkv.el_bias32 = att.qkv.bias.detach().float().contiguous() if att.qkv.bias is not None else None
        att.proj.el_bias32 = att.proj.bias.detach().float().contiguous() if att.proj.bias is not None else None
    if _pw_ok(att.qkv, [x]) and _pw_ok(att.proj, [x]):
        y = ops.linear_attention(pw_apply(att.qkv, [x], att.qkv.el_bias32, ops.ACT_NONE), att.num_heads)
        x = pw_apply(att.proj, [y], att.proj.el_bias32, ops.ACT_NONE, residual=x)
    else:
        y = ops.linear_attention(att.qkv(x), att.num_heads)
        x = ops.bias_act(F.conv2d(y, att.proj.weight), att.proj.el_bias32, ops.ACT_NONE, residual=x)
    return self.ffn[1](self.ffn[0](x), out=out, residual=x)


def c2psa_engine_forward(self, x):
    """C2PSA_LinearAttention.forward (block.py:3489-3497) without split / cat copies."""
    B, _, H, W = x.shape
    c = self.c
    a = torch.empty((B, c, H, W), device=x.device, dtype=x.dtype, memory_format=torch.channels_last)
    cur = torch.empty_like(a)
    self.cv1(x, out=a, out2=cur)  # pass-through half and attention branch as two dense tensors
    for blk in self.m:
        cur = blk(cur)
    return self.cv2([a, cur])


def sppf_engine_forward(self, x):
    """SPPF.forward (block.py:219-223): the three chained max-pools and the concat are one kernel."""
    y = self.cv1(x)
    if y.shape[2] * y.shape[3] * 64 > 200 * 1024 or self.m.kernel_size != 5:  # large maps keep the PyTorch pooling
        ys = [y]
        for _ in range(3):
            ys.append(self.m(ys[-1]))
        return self.cv2(torch.cat(ys, 1))
    return self.cv2(ops.sppf_pool(y))


def concat_engine_forward(self, x):
    """Concat (conv.py Concat) whose first input is the low-resolution map of the preceding nn.Upsample."""
    if getattr(self, "el_upsample_first", False):
        if getattr(self, "el_up_lazy", False) and x[0].dtype in (torch.bfloat16, torch.float16) and _pixel_linear(x[0]) and _pixel_linear(x[1]) \
                and x[1].shape[2] % 2 == 0 and x[1].shape[3] % 2 == 0:
            return UpCat((x[0], x[1]))  # consumed by the next block's cv1: the upsampled / concatenated tensor is never built
        return ops.upsample2x_cat(x[0], x[1])
    if getattr(self, "el_lazy", False):  # the only consumer is a block whose cv1 folds the concat into its K loop
        return list(x)
    return torch.cat(x, self.d)


# ------------------------------------------------------------------------------------------
# standalone modules (same names, constructor signatures and state-dict keys as the reference)
# ------------------------------------------------------------------------------------------


def autopad(k, p=None, d=1):
    if d > 1:
        k = d * (k - 1) + 1
    return k // 2 if p is None else p


class Conv(nn.Module):
    """conv + BatchNorm + SiLU (nn/modules/conv.py:41-60); keys `conv.weight`, `bn.*`."""

    default_act = nn.SiLU()

    def __init__(self, c1, c2, k=1, s=1, p=None, g=1, d=1, act=True):
        super().__init__()
        self.conv = nn.Conv2d(c1, c2, k, s, autopad(k, p, d), groups=g, dilation=d, bias=False)
        self.bn = nn.BatchNorm2d(c2)
        self.act = self.default_act if act is True else act if isinstance(act, nn.Module) else nn.Identity()

    def forward(self, x):
        return self.act(self.bn(self.conv(x)))

    def forward_fuse(self, x):
        return self.act(self.conv(x))


class DWConv(Conv):
    """Depth-wise Conv (conv.py:124-129)."""

    def __init__(self, c1, c2, k=1, s=1, d=1, act=True):
        super().__init__(c1, c2, k, s, g=math.gcd(c1, c2), d=d, act=act)


class _ScaleModule(nn.Module):
    """conv.py:448-461: a learnable per-channel scale (state-dict key `weight`)."""

    def __init__(self, dims, init_scale=1.0, init_bias=0):
        super().__init__()
        self.dims = dims
        self.weight = nn.Parameter(torch.ones(*dims) * init_scale)
        self.bias = None

    def forward(self, x):
        return torch.mul(self.weight, x)


class WTConv2d(nn.Module):
    """conv.py:463-598: depthwise conv + multi-level wavelet branch.  Same constructor, parameter names and state-dict
    keys (`wt_filter`, `iwt_filter`, `base_conv.*`, `base_scale.weight`, `wavelet_convs.i.weight`, `wavelet_scale.i.weight`).  The
    analysis / synthesis (the reference's grouped stride-2 conv and transposed conv with a repeated filter bank) run on
    `el_dwt_haar_fwd` / `el_dwt_haar_bwd` directly in the module's (B, C, 4, h, w) coefficient layout; the depthwise convolutions stay
    on cuDNN on this API path.  Only the Haar bank (`db1` / `haar`, the only one the reference's yamls use) is implemented."""

    def __init__(self, in_channels, out_channels, kernel_size=5, stride=1, bias=True, wt_levels=1, wt_type="db1"):
        super().__init__()
        assert in_channels == out_channels, "WTConv2d needs in_channels == out_channels"
        if wt_type not in ("db1", "haar"):
            raise NotImplementedError(f"WTConv2d: only the Haar filter bank (db1) has a kernel, got {wt_type!r}")
        self.in_channels, self.wt_levels, self.stride, self.dilation = in_channels, wt_levels, stride, 1
        s = torch.tensor(2.0 ** -0.5, dtype=torch.float32)
        lo, hi = torch.stack((s, s)), torch.stack((s, -s))  # reversed decomposition taps = reconstruction taps for Haar
        bank = torch.stack([lo.unsqueeze(0) * lo.unsqueeze(1), lo.unsqueeze(0) * hi.unsqueeze(1),
                            hi.unsqueeze(0) * lo.unsqueeze(1), hi.unsqueeze(0) * hi.unsqueeze(1)], 0)[:, None].repeat(in_channels, 1, 1, 1)
        self.wt_filter = nn.Parameter(bank.clone(), requires_grad=False)   # kept for state-dict compatibility; the kernels hold the taps
        self.iwt_filter = nn.Parameter(bank.clone(), requires_grad=False)
        pad = autopad(kernel_size, None, 1)
        self.base_conv = nn.Conv2d(in_channels, in_channels, kernel_size, 1, pad, 1, groups=in_channels, bias=bias)
        self.base_scale = _ScaleModule([1, in_channels, 1, 1])
        self.wavelet_convs = nn.ModuleList(
            [nn.Conv2d(in_channels * 4, in_channels * 4, kernel_size, 1, pad, 1, groups=in_channels * 4, bias=False) for _ in range(wt_levels)])
        self.wavelet_scale = nn.ModuleList([_ScaleModule([1, in_channels * 4, 1, 1], init_scale=0.1) for _ in range(wt_levels)])
        self.do_stride = nn.AvgPool2d(kernel_size=1, stride=stride) if stride > 1 else None

    def forward(self, x):
        lls, highs, shapes = [], [], []
        cur = x
        for i in range(self.wt_levels):
            shapes.append(cur.shape)
            if cur.shape[2] % 2 or cur.shape[3] % 2:  # odd sizes: zero pad right / bottom (conv.py:556-558)
                cur = F.pad(cur, (0, cur.shape[3] % 2, 0, cur.shape[2] % 2))
            coeffs = ops.wavelet_2d_transform(cur.contiguous())              # (B, C, 4, h, w)
            b, c, _, hh, ww = coeffs.shape
            cur = coeffs[:, :, 0]                                             # raw LL feeds the next level
            tag = self.wavelet_scale[i](self.wavelet_convs[i](coeffs.reshape(b, c * 4, hh, ww))).reshape(b, c, 4, hh, ww)
            lls.append(tag[:, :, 0])
            highs.append(tag[:, :, 1:4])
        nxt = 0
        for i in range(self.wt_levels - 1, -1, -1):
            ll, hi, shp = lls.pop() + nxt, highs.pop(), shapes.pop()
            nxt = ops.inverse_2d_wavelet_transform(torch.cat([ll.unsqueeze(2), hi], 2))[:, :, : shp[2], : shp[3]]
        y = self.base_scale(self.base_conv(x)) + nxt
        return self.do_stride(y) if self.do_stride is not None else y


def haar_dwt2d_forward(self, x: torch.Tensor):
    """`HaarDWT2D.forward` (nn/modules/block.py:247-259): (LL, LH, HL, HH) with LH = vertical and HL = horizontal difference, i.e. the
    second and third band of `el_dwt_haar_fwd` swapped.  The kernel's taps are float32(2^-1/2)^2 = 0.49999997 where this
    module's are 0.5: 6e-8 relative, inside the 1e-5 contract."""
    if x.shape[-1] % 2 or x.shape[-2] % 2:
        raise ValueError("HaarDWT2D needs even H and W (the reference's view(B, C, 4, H // 2, W // 2) fails otherwise)")
    buf = ops.dwt_haar(x)
    B = x.shape[0]
    return buf[:B], buf[2 * B : 3 * B], buf[B : 2 * B], buf[3 * B :]


def ihaar_dwt2d_forward(self, LL, LH, HL, HH):
    """Haar synthesis, the inverse of `HaarDWT2D` (the forward the reference's `IHaarDWT2D`, block.py:2714-2750, was meant to have: that
    class cannot be constructed in the reference, Q5, and its un-reachable forward feeds a band-major cat into a transposed conv whose
    groups expect channel-interleaved bands, which mixes channels).
    The four bands are centre-cropped to their common size like the reference (:2729-2735) and go through `el_dwt_haar_bwd`."""
    Hm, Wm = min(t.shape[-2] for t in (LL, LH, HL, HH)), min(t.shape[-1] for t in (LL, LH, HL, HH))

    def crop(t):
        dh, dw = (t.shape[-2] - Hm) // 2, (t.shape[-1] - Wm) // 2
        return t[..., dh : dh + Hm, dw : dw + Wm]

    fmt = torch.channels_last if ops._channels_last(LL) else torch.contiguous_format
    bands = torch.cat([crop(LL), crop(HL), crop(LH), crop(HH)], 0).contiguous(memory_format=fmt)  # kernel order: LL, horizontal, vertical, HH
    return ops.idwt_haar(bands, 2 * Hm, 2 * Wm)


_HAAR_2X2 = (("ll", [[0.5, 0.5], [0.5, 0.5]]), ("lh", [[0.5, 0.5], [-0.5, -0.5]]), ("hl", [[0.5, -0.5], [0.5, -0.5]]), ("hh", [[0.5, -0.5], [-0.5, 0.5]]))


def ihaar_dwt2d_init(self):
    """`IHaarDWT2D.__init__` as block.py:2715-2732 defines it before the stray second `__init__(self, dim, num_heads)` (:2755) overrides it:
    no arguments; buffers `ll`, `lh`, `hl`, `hh`, `recon_ll`, `recon_h`.  A plain function (no zero-argument super()) so that install() can
    bind it onto the reference's own class."""
    nn.Module.__init__(self)
    for name, k in _HAAR_2X2:
        self.register_buffer(name, torch.tensor(k, dtype=torch.float32))
    self.register_buffer("recon_ll", torch.tensor([[0.5, 0.5], [0.5, 0.5]], dtype=torch.float32).view(1, 1, 2, 2))
    self.register_buffer("recon_h", torch.tensor([[0.5, -0.5], [0.5, -0.5]], dtype=torch.float32).view(1, 1, 2, 2))


class HaarDWT2D(nn.Module):
    """block.py:225-259; the four 2 x 2 filters stay registered as buffers (`ll`, `lh`, `hl`, `hh`) for state-dict compatibility."""

    def __init__(self):
        super().__init__()
        for name, k in _HAAR_2X2:
            self.register_buffer(name, torch.tensor(k, dtype=torch.float32))

    forward = haar_dwt2d_forward


class IHaarDWT2D(nn.Module):
    """block.py:2714-2750, as it was meant to be (see ihaar_dwt2d_init / ihaar_dwt2d_forward)."""

    __init__ = ihaar_dwt2d_init
    forward = ihaar_dwt2d_forward


class WaveletMixerMultiLevel(nn.Module):
    """Two-level Haar mixer, block.py:2600-2660: same constructor, parameter names (`f_ll1`, `f_lh1`, `f_hl1`, `f_hh1`, `f_ll2_head`,
    `dw_weight`, `f_ll2_tail`, `f_h2`) and forward; both analyses and both syntheses run on the DWT kernels (the reference's own
    version cannot be built because its `IHaarDWT2D` is broken, Q5).  Needs H and W divisible by 4."""

    def __init__(self, c, use_dilated=True, k=5, d=3):
        super().__init__()
        self.dwt, self.idwt = HaarDWT2D(), IHaarDWT2D()
        self.f_ll1, self.f_lh1, self.f_hl1, self.f_hh1 = Conv(c, c, 1, 1), Conv(c, c, 3, 1), Conv(c, c, 3, 1), Conv(c, c, 3, 1)
        self.use_dilated, self.k, self.d = use_dilated, k, d
        self.f_ll2_head = Conv(c, c, 1, 1)
        self.dw_weight = nn.Parameter(torch.empty(c, 1, k, k))
        nn.init.kaiming_uniform_(self.dw_weight, a=math.sqrt(5))
        self.dw_bias = None
        self.f_ll2_tail = Conv(c, c, 1, 1)
        self.f_h2 = Conv(c, c, 3, 1)

    def _depthwise_dynamic(self, x):
        """block.py:2629-2641: the dilation shrinks on small maps so that the dilated kernel never exceeds the input."""
        d = 1
        if self.use_dilated:
            d = min(self.d, max(1, (min(x.shape[-2:]) - 1) // (self.k - 1)))
        return F.conv2d(x, self.dw_weight, self.dw_bias, stride=1, padding=((self.k - 1) * d) // 2, dilation=d, groups=x.shape[1])

    def forward(self, x):
        LL1, LH1, HL1, HH1 = self.dwt(x)
        LL1, LH1, HL1, HH1 = self.f_ll1(LL1), self.f_lh1(LH1), self.f_hl1(HL1), self.f_hh1(HH1)
        LL2, LH2, HL2, HH2 = self.dwt(LL1)
        LL2 = self.f_ll2_tail(self._depthwise_dynamic(self.f_ll2_head(LL2)))
        LH2, HL2, HH2 = self.f_h2(LH2), self.f_h2(HL2), self.f_h2(HH2)
        return self.idwt(self.idwt(LL2, LH2, HL2, HH2), LH1, HL1, HH1)


class DSConv(nn.Module):
    """Depthwise-separable conv with its own BatchNorm (conv.py:87-104); keys `dw`, `pw`, `bn`."""

    def __init__(self, c_in, c_out, k=3, s=1, p=None, d=1, bias=False):
        super().__init__()
        p = (d * (k - 1)) // 2 if p is None else p
        self.dw = nn.Conv2d(c_in, c_in, k, s, p, dilation=d, groups=c_in, bias=bias)
        self.pw = nn.Conv2d(c_in, c_out, 1, 1, 0, bias=bias)
        self.bn = nn.BatchNorm2d(c_out)
        self.act = nn.SiLU()

    def forward(self, x):
        return self.act(self.bn(self.pw(self.dw(x))))


class Concat(nn.Module):
    def __init__(self, dimension=1):
        super().__init__()
        self.d = dimension

    def forward(self, x):
        return torch.cat(x, self.d)


class SPPF(nn.Module):
    """nn/modules/block.py:204-223."""

    def __init__(self, c1, c2, k=5):
        super().__init__()
        c_ = c1 // 2
        self.cv1 = Conv(c1, c_, 1, 1)
        self.cv2 = Conv(c_ * 4, c2, 1, 1)
        self.m = nn.MaxPool2d(kernel_size=k, stride=1, padding=k // 2)

    def forward(self, x):
        y = [self.cv1(x)]
        for _ in range(3):
            y.append(self.m(y[-1]))
        return self.cv2(torch.cat(y, 1))


class DSBottleneck(nn.Module):
    """block.py:1467-1503."""

    def __init__(self, c1, c2, shortcut=True, e=0.5, k1=3, k2=5, d2=1):
        super().__init__()
        c_ = int(c2 * e)
        self.cv1 = DSConv(c1, c_, k1, s=1, p=None, d=1)
        self.cv2 = DSConv(c_, c2, k2, s=1, p=None, d=d2)
        self.add = shortcut and c1 == c2

    def forward(self, x):
        y = self.cv2(self.cv1(x))
        return x + y if self.add else y


class DSC3k(nn.Module):
    """C3 shell around DSBottlenecks (block.py:1506-1562, parent C3 at :382-396)."""

    def __init__(self, c1, c2, n=1, shortcut=True, g=1, e=0.5, k1=3, k2=5, d2=1):
        super().__init__()
        c_ = int(c2 * e)
        self.cv1 = Conv(c1, c_, 1, 1)
        self.cv2 = Conv(c1, c_, 1, 1)
        self.cv3 = Conv(2 * c_, c2, 1)
        self.m = nn.Sequential(*(DSBottleneck(c_, c_, shortcut=shortcut, e=1.0, k1=k1, k2=k2, d2=d2) for _ in range(n)))

    def forward(self, x):
        return self.cv3(torch.cat((self.m(self.cv1(x)), self.cv2(x)), 1))


class _PywtDWT2D(nn.Module):
    """Haar analysis; no parameters and no persistent buffers (block.py:3582-3642)."""

    def __init__(self, wave: str = "haar", mode: str = "symmetric"):
        super().__init__()
        if wave != "haar":
            raise NotImplementedError("edge_yolo_b200 implements the Haar wavelet only")
        self.wave_name, self.mode = wave, mode

    forward = dwt_forward


class _WaveletEnhancer(nn.Module):
    """block.py:3645-3710; keys `alpha`, `gamma`, `f_ll.*`, `f_h.*`, `fuse.*`."""

    def __init__(self, c: int, use_ds: bool = False, alpha0=(0.5, 0.2, 0.2, 0.1), wave: str = "haar", mode: str = "symmetric"):
        super().__init__()
        self.c = c
        self.dwt = _PywtDWT2D(wave=wave, mode=mode)
        self.f_ll = Conv(c, c // 2, k=1, s=1)
        self.f_h = (DSConv if use_ds else Conv)(c, c // 2, k=3, s=1)
        self.fuse = Conv(3 * c, c, k=1, s=1)
        self.alpha = nn.Parameter(torch.tensor(alpha0, dtype=torch.float32))
        self.gamma = nn.Parameter(torch.tensor(0.0))

    forward = wavelet_enhancer_forward


class DSC3K2_Wavelet(nn.Module):
    """C2f shell whose stacked branch is wavelet-enhanced (block.py:3749-3788)."""

    def __init__(self, c1, c2, n=1, dsc3k=False, e=0.5, g=1, shortcut=True, k1=3, k2=7, d2=1, **kwargs):
        super().__init__()
        use_ds = bool(kwargs.get("use_ds", False))
        self.c = int(c2 * e)
        self.cv1 = Conv(c1, 2 * self.c, 1, 1)
        self.cv2 = Conv((2 + n) * self.c, c2, 1, 1)
        if dsc3k:
            self.m = nn.ModuleList(DSC3k(self.c, self.c, n=2, shortcut=shortcut, g=g) for _ in range(n))
        else:
            self.m = nn.ModuleList(DSBottleneck(self.c, self.c, shortcut=shortcut, e=1.0, k1=k1, k2=k2, d2=d2) for _ in range(n))
        self.wave = _WaveletEnhancer(self.c, use_ds=use_ds, wave=kwargs.get("wave", "haar"), mode=kwargs.get("mode", "symmetric"))

    def forward(self, x):
        y = list(self.cv1(x).chunk(2, 1))
        y[1] = self.wave(y[1])
        for m in self.m:
            y.append(m(y[-1]))
        return self.cv2(torch.cat(y, 1))


class LinearAttention(nn.Module):
    """block.py:3348-3373; keys `qkv.weight`, `qkv.bias`, `proj.weight`; extra kwargs are dropped."""

    def __init__(self, dim, num_heads, attn_ratio=None, qkv_bias=False, proj_bias=True, **kwargs):
        super().__init__()
        assert dim % num_heads == 0, "LinearAttention: dim must be divisible by num_heads"
        self.dim, self.num_heads, self.head_dim = dim, num_heads, dim // num_heads
        self.qkv = nn.Conv2d(dim, 3 * dim, kernel_size=1, bias=qkv_bias)
        self.proj = nn.Conv2d(dim, dim, kernel_size=1, bias=proj_bias)

    forward = linear_attention_forward


class PSABlock_LinearAttention(nn.Module):
    """block.py:3412-3449."""

    def __init__(self, dim, attn_ratio=0.5, num_heads=None, mlp_ratio=2.0, qkv_bias=True, proj_bias=False, fmap="elu", eps=1e-6):
        super().__init__()
        self.attn = LinearAttention(dim=dim, num_heads=num_heads, attn_ratio=attn_ratio, qkv_bias=qkv_bias, proj_bias=proj_bias, fmap=fmap, eps=eps)
        hidden = int(dim * mlp_ratio)
        self.ffn = nn.Sequential(Conv(dim, hidden, k=1, s=1, act=True), Conv(hidden, dim, k=1, s=1, act=False))

    def forward(self, x):
        x = x + self.attn(x)
        return x + self.ffn(x)


class C2PSA_LinearAttention(nn.Module):
    """block.py:3452-3497."""

    def __init__(self, c1, c2, n=1, e=0.5, attn_ratio=0.5, num_heads=None, mlp_ratio=2.0, fmap="elu"):
        super().__init__()
        assert c1 == c2, "C2PSA_LinearAttention requires c1 == c2"
        self.c = int(c1 * e)
        heads = max(1, (self.c // 64) if num_heads is None else num_heads)
        assert self.c % heads == 0, f"branch channels {self.c} must be divisible by num_heads {heads}"
        self.cv1 = Conv(c1, 2 * self.c, k=1, s=1)
        self.m = nn.Sequential(*[PSABlock_LinearAttention(dim=self.c, attn_ratio=attn_ratio, num_heads=heads, mlp_ratio=mlp_ratio, fmap=fmap)
                                 for _ in range(n)])
        self.cv2 = Conv(2 * self.c, c1, k=1, s=1)

    def forward(self, x):
        a, b = torch.split(self.cv1(x), (self.c, self.c), dim=1)
        return self.cv2(torch.cat((a, self.m(b)), dim=1))


class DFL(nn.Module):
    """Frozen arange 1x1 conv (block.py:72-91).  Kept for the `dfl.conv.weight` state-dict key; the
    integral itself runs inside the fused decode kernel."""

    def __init__(self, c1=16):
        super().__init__()
        self.conv = nn.Conv2d(c1, 1, 1, bias=False).requires_grad_(False)
        self.conv.weight.data[:] = torch.arange(c1, dtype=torch.float).view(1, c1, 1, 1)
        self.c1 = c1


class GFLHeadv2_uniH(nn.Module):
    """Detect towers + DGQP quality heads (head.py:38-189, 194-345, 827-908)."""

    dynamic = False
    export = False
    format = None
    end2end = False
    max_det = 300
    shape = None
    anchors = torch.empty(0)
    strides = torch.empty(0)
    legacy = False

    def __init__(self, nc=80, ch=(), reg_topk=4, add_mean=True, reg_channels=64, use_dat=False, use_cit=False, use_poscnn=False):
        super().__init__()
        self.nc, self.nl, self.reg_max = nc, len(ch), 16
        self.no = nc + self.reg_max * 4
        self.stride = torch.zeros(self.nl)
        c2, c3 = max((16, ch[0] // 4, self.reg_max * 4)), max(ch[0], min(self.nc, 100))
        self.cv2 = nn.ModuleList(nn.Sequential(Conv(x, c2, 3), Conv(c2, c2, 3), nn.Conv2d(c2, 4 * self.reg_max, 1)) for x in ch)
        self.cv3 = nn.ModuleList(
            nn.Sequential(nn.Sequential(DWConv(x, x, 3), Conv(x, c3, 1)), nn.Sequential(DWConv(c3, c3, 3), Conv(c3, c3, 1)), nn.Conv2d(c3, self.nc, 1))
            for x in ch)
        self.dfl = DFL(self.reg_max)
        self.reg_topk, self.add_mean, self.reg_channels = reg_topk, add_mean, reg_channels
        self.apply_quality_in_inference = True
        in_stat = 4 * (reg_topk + (1 if add_mean else 0))
        self.reg_conf = nn.ModuleList(
            nn.Sequential(nn.Conv2d(in_stat, reg_channels, 1, bias=True), nn.ReLU(inplace=True), nn.Conv2d(reg_channels, 1, 1, bias=True), nn.Sigmoid())
            for _ in ch)
        self.stem = nn.ModuleList(nn.Identity() for _ in ch)
        self.dat = self.pos_cls = self.pos_reg = self.cit_cls = self.cit_reg = None
        self._qualities = None

    forward = gfl_head_forward

    def bias_init(self):
        """head.py:150-161 (not called by the reference's DetectionModel; used for the 'trained-like' regime)."""
        for a, b, s in zip(self.cv2, self.cv3, self.stride):
            a[-1].bias.data[:] = 1.0
            b[-1].bias.data[: self.nc] = math.log(5 / self.nc / (640 / float(s)) ** 2)
