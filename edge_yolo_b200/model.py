"""EdgeLine-YOLO graph builder (the `yolo11-test.yaml` topology) without ultralytics.

Mirrors what `parse_model` (nn/tasks.py:958-1147) builds from
`cfg/models/11/yolo11-test.yaml`: same layer indices, channel arithmetic and therefore the same
`model.{i}.*` state-dict keys, so reference checkpoints load unchanged.  Used by bench / smoke /
tests on machines where the reference package is absent; with ultralytics present use
`edge_yolo_b200.install()` instead and keep `YOLO(cfg, task="detect")`.
"""
from __future__ import annotations

import math
import types

import torch
import torch.nn as nn

from . import modules as M
from .modules import C2PSA_LinearAttention, Concat, Conv, DSC3K2_Wavelet, DSConv, GFLHeadv2_uniH, SPPF

# [depth, width, max_channels]  (yolo11-test.yaml:9-15)
SCALES = {"n": (0.50, 0.25, 1024), "s": (0.50, 0.50, 1024), "m": (0.50, 1.00, 512), "l": (1.00, 1.00, 512), "x": (1.00, 1.50, 512)}

# (from, repeats, module, args)  (yolo11-test.yaml:18-50)
_LAYERS = [
    (-1, 1, "Conv", [64, 3, 2]),
    (-1, 1, "Conv", [128, 3, 2]),
    (-1, 2, "DSC3K2_Wavelet", [256, False, 0.25]),
    (-1, 1, "Conv", [256, 3, 2]),
    (-1, 2, "DSC3K2_Wavelet", [512, False, 0.25]),
    (-1, 1, "Conv", [512, 3, 2]),
    (-1, 2, "DSC3K2_Wavelet", [512, True]),
    (-1, 1, "Conv", [1024, 3, 2]),
    (-1, 2, "DSC3K2_Wavelet", [1024, True]),
    (-1, 1, "SPPF", [1024, 5]),
    (-1, 2, "C2PSA_LinearAttention", [1024]),
    (-1, 1, "Upsample", [None, 2, "nearest"]),
    ([-1, 6], 1, "Concat", [1]),
    (-1, 2, "DSC3K2_Wavelet", [512, False]),
    (-1, 1, "Upsample", [None, 2, "nearest"]),
    ([-1, 4], 1, "Concat", [1]),
    (-1, 2, "DSC3K2_Wavelet", [256, False]),
    (-1, 1, "Conv", [256, 3, 2]),
    ([-1, 13], 1, "Concat", [1]),
    (-1, 2, "DSC3K2_Wavelet", [512, False]),
    (-1, 1, "Conv", [512, 3, 2]),
    ([-1, 10], 1, "Concat", [1]),
    (-1, 2, "DSC3K2_Wavelet", [1024, True]),
    ([16, 19, 22], 1, "GFLHeadv2_uniH", ["nc"]),
]


def make_divisible(x, divisor):
    return math.ceil(x / divisor) * divisor


class EdgeLineYOLO(nn.Module):
    """Detection model: `forward(x)` -> list of raw maps (train) or `(y, maps)` (eval), like DetectionModel."""

    def __init__(self, scale: str = "n", nc: int = 80, ch: int = 3):
        super().__init__()
        depth, width, max_ch = SCALES[scale]
        self.scale, self.nc = scale, nc
        chans, layers, self.save = [ch], [], set()
        for i, (f, n, name, args) in enumerate(_LAYERS):
            args = list(args)
            n = max(round(n * depth), 1) if n > 1 else n
            c1 = chans[f] if isinstance(f, int) else None
            if name in ("Conv", "SPPF", "DSC3K2_Wavelet", "C2PSA_LinearAttention"):
                c2 = make_divisible(min(args[0], max_ch) * width, 8)
                args = [c1, c2, *args[1:]]
                if name in ("DSC3K2_Wavelet", "C2PSA_LinearAttention"):
                    args.insert(2, n)  # repeats become a constructor argument (tasks.py:1067-1068)
                    if name == "DSC3K2_Wavelet" and scale in "lx":
                        args[3] = True  # tasks.py:1069-1072
                m = {"Conv": Conv, "SPPF": SPPF, "DSC3K2_Wavelet": DSC3K2_Wavelet, "C2PSA_LinearAttention": C2PSA_LinearAttention}[name](*args)
            elif name == "Upsample":
                m, c2 = nn.Upsample(*args), c1
            elif name == "Concat":
                m, c2 = Concat(*args), sum(chans[x] for x in f)
            else:  # head
                m, c2 = GFLHeadv2_uniH(nc, [chans[x] for x in f]), None
            m.i, m.f = i, f
            layers.append(m)
            self.save.update(x % i for x in ([f] if isinstance(f, int) else f) if x != -1)
            if i == 0:
                chans = []
            chans.append(c2)
        self.model = nn.Sequential(*layers)
        head = self.model[-1]
        head.stride = torch.tensor([8.0, 16.0, 32.0])  # what the reference's 256x256 probe measures (tasks.py:361-363)
        self.stride = head.stride
        for m in self.modules():  # initialize_weights, utils/torch_utils.py:416-418
            if isinstance(m, nn.BatchNorm2d):
                m.eps, m.momentum = 1e-3, 0.03
            elif isinstance(m, (nn.SiLU, nn.ReLU)):
                m.inplace = True

    def forward(self, x, stem_out=None):
        """`stem_out`: precomputed output of layer 0 (the engine's fused uint8 stem, ops.stem_conv_u8); `x` is then unused."""
        outs = []
        for m in self.model:
            if stem_out is not None and m.i == 0:
                x = stem_out
                outs.append(x if 0 in self.save else None)
                continue
            if m.f != -1:
                x = outs[m.f] if isinstance(m.f, int) else [x if j == -1 else outs[j] for j in m.f]
            if getattr(m, "el_fused_into_next", False):  # nn.Upsample folded into the following Concat (engine mode)
                outs.append(None)
                continue
            x = m(x)
            outs.append(x if m.i in self.save else None)
        return x

    @torch.no_grad()
    def fuse(self, dsconv: bool = False, engine: bool = False):
        """Fold BatchNorm into the preceding conv.  Like BaseModel.fuse (tasks.py:214-242) only `Conv`
        (incl. DWConv) is folded by default; `dsconv=True` additionally folds DSConv's BN into its
        pointwise conv (identical maths in eval mode, one kernel less per DSConv).

        `engine=True` (implies dsconv; eval / NHWC only) additionally moves every folded bias + activation
        (+ shortcut add) into one `el_bias_act` epilogue kernel, removes the torch.cat / split copies of the
        C2f-style blocks and fuses nn.Upsample + Concat pairs (see modules.*_engine_forward)."""
        dsconv = dsconv or engine
        acts = {nn.SiLU: M.ops.ACT_SILU, nn.Identity: M.ops.ACT_NONE, nn.ReLU: M.ops.ACT_RELU}
        for m in self.modules():
            if isinstance(m, Conv) and hasattr(m, "bn"):
                m.conv = _fold(m.conv, m.bn)
                del m.bn
                m.forward = m.forward_fuse
                if engine and type(m.act) in acts:
                    m.el_bias, m.el_act = m.conv.bias.detach().float().clone(), acts[type(m.act)]
                    m.conv.bias = None
                    m.forward = types.MethodType(M.conv_engine_forward, m)
                    if _dw_eligible(m.conv, epilogue=True):  # DWConv: depthwise + bias + activation in one kernel
                        m.el_dw, m.el_k = M.ops.pack_dw_weight(m.conv.weight), m.conv.kernel_size[0]
                        m.forward = types.MethodType(M.dwconv_engine_forward, m)
            elif dsconv and isinstance(m, DSConv) and isinstance(m.bn, nn.BatchNorm2d):
                m.pw = _fold(m.pw, m.bn)
                m.bn = nn.Identity()
                if engine:
                    m.el_bias = m.pw.bias.detach().float().clone()
                    m.pw.bias = None
                    m.forward = types.MethodType(M.dsconv_engine_forward, m)
                    if _dw_eligible(m.dw) and m.dw.bias is None:
                        m.el_dw, m.el_k = M.ops.pack_dw_weight(m.dw.weight), m.dw.kernel_size[0]
        if engine:
            binds = {M.DSBottleneck: M.dsbottleneck_engine_forward, M.DSC3k: M.dsc3k_engine_forward,
                     M.DSC3K2_Wavelet: M.dsc3k2_wavelet_engine_forward, M.PSABlock_LinearAttention: M.psablock_engine_forward,
                     M.C2PSA_LinearAttention: M.c2psa_engine_forward, M.SPPF: M.sppf_engine_forward}
            for m in self.modules():
                if type(m) in binds:
                    m.forward = types.MethodType(binds[type(m)], m)
            for m in self.modules():  # DWConv(k = 3) -> Conv(1x1) pairs (class towers): one fused kernel
                if (isinstance(m, nn.Sequential) and len(m) == 2 and all(isinstance(t, Conv) and hasattr(t, "el_bias") for t in m)
                        and hasattr(m[0], "el_dw") and m[0].el_k == 3 and m[1].conv.kernel_size == (1, 1)):
                    m.forward = types.MethodType(M.dwpw_engine_forward, m)
            head = self.model[-1]
            if isinstance(head, GFLHeadv2_uniH):  # the last 1x1 convs of both towers lose their bias; the decode kernels add it
                box_b, cls_b = [], []
                for seq, dst in ((head.cv2, box_b), (head.cv3, cls_b)):
                    for tower in seq:
                        dst.append(tower[-1].bias.detach().float().clone())
                        tower[-1].bias = None
                        if tower[-1].kernel_size == (1, 1):
                            tower[-1].forward = types.MethodType(M.conv2d_pw_forward, tower[-1])
                head.el_head_bias = (box_b, cls_b)
            layers = list(self.model)
            for up, cat in zip(layers, layers[1:]):
                if (isinstance(up, nn.Upsample) and isinstance(cat, Concat) and up.scale_factor == 2 and up.mode == "nearest"
                        and isinstance(cat.f, list) and cat.f[0] == -1 and cat.d == 1 and up.i not in self.save):
                    up.el_fused_into_next, cat.el_upsample_first = True, True
                    cat.forward = types.MethodType(M.concat_engine_forward, cat)
                    nxt = layers[cat.i + 1] if cat.i + 1 < len(layers) else None
                    if (isinstance(nxt, M.DSC3K2_Wavelet) and nxt.f == -1 and cat.i not in self.save and hasattr(nxt.cv1, "el_bias")
                            and nxt.cv1.conv.kernel_size == (1, 1) and nxt.c % 16 == 0 and len(cat.f) == 2):
                        cat.el_up_lazy = True  # Upsample + Concat + cv1 collapse into a low-resolution GEMM + an addend epilogue
            for cat, nxt in zip(layers, layers[1:]):  # Concat -> DSC3K2_Wavelet: cv1 reads the parts in place
                if (isinstance(cat, Concat) and not getattr(cat, "el_upsample_first", False) and cat.d == 1 and cat.i not in self.save
                        and isinstance(nxt, M.DSC3K2_Wavelet) and nxt.f == -1 and isinstance(cat.f, list) and len(cat.f) <= 4):
                    cat.el_lazy = True
                    cat.forward = types.MethodType(M.concat_engine_forward, cat)
            self.el_engine_fused = True
        return self

    def load_state_dict(self, state_dict, strict: bool = True, assign: bool = False):
        """An engine-fused model holds state that no state dict carries (folded-BN biases as fp32 side tensors, packed depthwise
        filters, the head biases handed to the decode kernel): loading weights into it would silently pair new weights with old
        biases.  Load the checkpoint first, then `fuse(engine=True)`."""
        if getattr(self, "el_engine_fused", False):
            raise RuntimeError("EdgeLineYOLO: load_state_dict after fuse(engine=True) is not supported -- load the checkpoint before fusing")
        return super().load_state_dict(state_dict, strict=strict, assign=assign)


def _dw_eligible(conv: nn.Conv2d, epilogue: bool = False) -> bool:
    """Depthwise, stride 1, 'same' padding, k in {3,5,7}: what el_dwconv_fwd implements.  Measured on B200 (tools/prof_conv.py):
    it beats PyTorch's depthwise kernel for k = 7 and whenever the bias + activation epilogue rides along (DWConv); for the bare
    k = 3 depthwise of DSConv the two are on par in isolation (modules.DWCONV_K3 picks ours: it chains with the neighbouring
    GEMMs through programmatic dependent launch)."""
    k = conv.kernel_size[0]
    C = conv.in_channels
    if not M.USE_DWCONV or not (epilogue or k >= 5 or M.DWCONV_K3):
        return False
    # The kernels take any multiple of 8 channels (vectors blocked by their largest divisor <= 8).  Round 1 kept the sites whose vector count
    # is not a power of two (the 80-channel class towers: blocks of 5) on PyTorch's depthwise kernel + el_bias_act, which was ~1 % faster over
    # the graph than the cp.async tile kernel; the persistent TMA kernel takes any block width since round 2 (modules.DWCONV_ANY_C).
    C8 = C // 8
    return (C % 8 == 0 and (C8 & (C8 - 1) == 0 or C % 64 == 0 or (M.DWCONV_ANY_C and k == 3 and C >= 16)) and conv.groups == C == conv.out_channels and C > 1 and conv.kernel_size == (k, k)
            and k in (3, 5, 7) and conv.stride == (1, 1) and conv.dilation == (1, 1) and conv.padding == (k // 2, k // 2))


def _fold(conv: nn.Conv2d, bn: nn.BatchNorm2d) -> nn.Conv2d:
    fused = nn.Conv2d(conv.in_channels, conv.out_channels, conv.kernel_size, conv.stride, conv.padding, conv.dilation, conv.groups, bias=True)
    fused = fused.to(conv.weight.device, conv.weight.dtype).requires_grad_(False)
    scale = bn.weight / torch.sqrt(bn.running_var + bn.eps)
    fused.weight.copy_(conv.weight * scale.view(-1, 1, 1, 1))
    bias = conv.bias if conv.bias is not None else torch.zeros_like(bn.running_mean)
    fused.bias.copy_((bias - bn.running_mean) * scale + bn.bias)
    return fused
