"""TEST INFRASTRUCTURE ONLY -- CPU oracle of the whole EdgeLine-YOLO forward + NMS.

Builds the product's graph (`edge_yolo_b200.model.EdgeLineYOLO`: plain PyTorch convs) on the CPU in
fp32 and swaps every hot-path forward for the oracle restatement in `hotpath.py`, so the result
is "reference semantics on the host cores".  Checked against the real reference with shared
weights in tests/test_reference_model.py (dev container only).  Used by the tests, `smoke()` and
the `cpu_baseline` / `--impl reference` legs of bench.py -- never by the product.
"""
from __future__ import annotations

import types

import torch

from edge_yolo_b200 import modules as M
from edge_yolo_b200.model import EdgeLineYOLO

from . import hotpath as O


def _enhancer_forward(self, b):
    LL, LH, HL, HH = O.dwt_haar(b)
    LLp, LHp, HLp, HHp = self.f_ll(LL), self.f_h(LH), self.f_h(HL), self.f_h(HH)
    y = self.fuse(O.wave_merge(b, LLp, LHp, HLp, HHp, self.alpha))
    return O.gated_residual(b, y, self.gamma)


def _attention_forward(self, x):
    return self.proj(O.linear_attention_core(self.qkv(x), self.num_heads))


def _head_forward(self, x):
    boxes = [self.cv2[i](x[i]) for i in range(self.nl)]
    clss = [self.cv3[i](x[i]) for i in range(self.nl)]
    feats = [torch.cat((b, c), 1) for b, c in zip(boxes, clss)]
    if self.training:
        return feats
    quals = [O.dgqp_quality(boxes[i], self.reg_conf[i][0].weight, self.reg_conf[i][0].bias, self.reg_conf[i][2].weight, self.reg_conf[i][2].bias)
             for i in range(self.nl)]
    return O.gfl_decode(boxes, clss, quals, [float(s) for s in self.stride]), feats


def to_oracle(model: EdgeLineYOLO) -> EdgeLineYOLO:
    """Rebind the hot-path forwards of a (CPU, fp32) product model to the oracle ops, in place."""
    for m in model.modules():
        if isinstance(m, M._WaveletEnhancer):
            m.forward = types.MethodType(_enhancer_forward, m)
        elif isinstance(m, M.LinearAttention):
            m.forward = types.MethodType(_attention_forward, m)
        elif isinstance(m, M.GFLHeadv2_uniH):
            m.forward = types.MethodType(_head_forward, m)
    return model


def build(scale="n", nc=80, seed=0, gamma=0.5) -> EdgeLineYOLO:
    """Seeded random-init EdgeLine-YOLO on the CPU with oracle ops; gamma != 0 so the wavelet branch is live."""
    torch.manual_seed(seed)
    model = EdgeLineYOLO(scale, nc).float().eval()
    with torch.no_grad():
        for m in model.modules():
            if isinstance(m, M._WaveletEnhancer):
                m.gamma.fill_(gamma)
    return to_oracle(model)


@torch.no_grad()
def predict(model, images, conf=0.25, iou=0.7, max_det=300, multi_label=False):
    """images (B,3,H,W) float in [0,1] -> list of (k,6) arrays (reference predict defaults: cfg/default.yaml iou 0.7, max_det 300)."""
    y, _ = model(images.float())
    outs, _ = O.non_max_suppression(y.numpy(), conf_thres=conf, iou_thres=iou, max_det=max_det, multi_label=multi_label)
    return outs
