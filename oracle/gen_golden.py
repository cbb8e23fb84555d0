"""TEST INFRASTRUCTURE ONLY -- generate `tests/golden/*.npz` from the unmodified reference.

Run in the dev container (needs `/root/reference`):  `python -m oracle.gen_golden`.
Every fixture stores seeded inputs and the outputs of the reference's *own* functions
(cited below), so the oracle restatement and the CUDA kernels can be pinned on machines
where the reference is absent (the GPU box).  Fixtures are kept small (a few hundred KB in
total); parity at BASELINE sizes is tested oracle-vs-CUDA.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_loader  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _np(t):
    return t.detach().cpu().numpy()


def _save(name, **arrays):
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **arrays)
    print(f"{name}: {os.path.getsize(path) / 1024:.1f} KiB  keys={list(arrays)}")


def _randomise_bn(module, g):
    for m in module.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.running_mean.copy_(torch.randn(m.running_mean.shape, generator=g) * 0.2)
            m.running_var.copy_(torch.rand(m.running_var.shape, generator=g) + 0.5)
            m.weight.data.copy_(torch.rand(m.weight.shape, generator=g) + 0.5)
            m.bias.data.copy_(torch.randn(m.bias.shape, generator=g) * 0.1)


def gen_dwt(block):
    """`_PywtDWT2D.forward`, nn/modules/block.py:3619-3642."""
    g = torch.Generator().manual_seed(1)
    dwt = block._PywtDWT2D()
    out = {}
    cases = {
        "even": torch.randn(2, 6, 10, 12, generator=g),
        "odd": torch.randn(1, 3, 5, 7, generator=g),
        "slice": torch.randn(2, 8, 6, 8, generator=g)[:, 4:],  # channel-slice view like chunk(2,1)
    }
    for k, x in cases.items():
        LL, LH, HL, HH = dwt(x)
        out[f"{k}_x"] = _np(x)
        out[f"{k}_bands"] = _np(torch.stack([LL, LH, HL, HH], 0))
    out["weight"] = _np(dwt.weight)
    _save("dwt", **out)


def gen_enhancer(block):
    """`_WaveletEnhancer.forward`, block.py:3685-3710 (merge input to `fuse` hooked)."""
    out = {}
    for tag, (c, shape) in {"even": (8, (2, 8, 12, 10)), "odd": (4, (1, 4, 7, 9))}.items():
        g = torch.Generator().manual_seed(2)
        torch.manual_seed(2)
        enh = block._WaveletEnhancer(c).eval()
        _randomise_bn(enh, g)
        with torch.no_grad():
            enh.gamma.fill_(0.5)  # gamma=0 makes the branch an identity
            enh.alpha.copy_(torch.tensor([0.7, -0.3, 0.2, 0.1]))
        grabbed = {}
        h1 = enh.fuse.register_forward_hook(lambda m, i, o: grabbed.update(cat=i[0].detach(), fused=o.detach()))
        h2 = enh.f_ll.register_forward_hook(lambda m, i, o: grabbed.update(LLp=o.detach()))
        hp = []
        h3 = enh.f_h.register_forward_hook(lambda m, i, o: hp.append(o.detach()))
        x = torch.randn(*shape, generator=g)
        with torch.no_grad():
            y = enh(x)
        for h in (h1, h2, h3):
            h.remove()
        out.update({f"{tag}_x": _np(x), f"{tag}_y": _np(y), f"{tag}_cat": _np(grabbed["cat"]), f"{tag}_fused": _np(grabbed["fused"]),
                    f"{tag}_LLp": _np(grabbed["LLp"]), f"{tag}_LHp": _np(hp[0]), f"{tag}_HLp": _np(hp[1]), f"{tag}_HHp": _np(hp[2]),
                    f"{tag}_alpha": _np(enh.alpha), f"{tag}_gamma": _np(enh.gamma)})
        for k, v in enh.state_dict().items():
            out[f"{tag}_sd_{k}"] = _np(v)
    _save("enhancer", **out)


def gen_attention(block):
    """`LinearAttention.forward`, block.py:3360-3373 (qkv conv output and proj input hooked)."""
    out = {}
    for tag, (dim, heads, hw) in {"h2": (128, 2, (5, 6)), "h1": (64, 1, (7, 11))}.items():
        torch.manual_seed(3)
        g = torch.Generator().manual_seed(3)
        att = block.LinearAttention(dim, heads, qkv_bias=True, proj_bias=False).eval()
        grabbed = {}
        h1 = att.qkv.register_forward_hook(lambda m, i, o: grabbed.update(qkv=o.detach()))
        h2 = att.proj.register_forward_hook(lambda m, i, o: grabbed.update(y=i[0].detach()))
        x = torch.randn(2, dim, *hw, generator=g) * 2.0
        with torch.no_grad():
            att(x)
        h1.remove(), h2.remove()
        out.update({f"{tag}_qkv": _np(grabbed["qkv"]), f"{tag}_y": _np(grabbed["y"]), f"{tag}_heads": np.int64(heads)})
    _save("attention", **out)


def gen_head(head_mod):
    """`GFLHeadv2_uniH.forward` eval path, head.py:880-908 -> 227-243 + 301-345."""
    torch.manual_seed(4)
    g = torch.Generator().manual_seed(4)
    nc, ch = 5, (16, 32, 64)
    head = head_mod.GFLHeadv2_uniH(nc=nc, ch=ch).eval()
    head.stride = torch.tensor([8.0, 16.0, 32.0])
    _randomise_bn(head, g)
    with torch.no_grad():
        for seq in head.reg_conf:
            seq[0].weight.mul_(4.0)  # spread q away from 0.5
    feats = [torch.randn(2, c, s, s + 1, generator=g) for c, s in zip(ch, (8, 4, 2))]
    with torch.no_grad():
        y, x = head([f.clone() for f in feats])
        quals = [head.reg_conf[i](_stat(x[i][:, :64])) for i in range(3)]
    out = {"y": _np(y), "nc": np.int64(nc), "strides": np.array([8.0, 16.0, 32.0], np.float32)}
    for i in range(3):
        out[f"x{i}"] = _np(x[i])
        out[f"q{i}"] = _np(quals[i])
        out[f"w1_{i}"] = _np(head.reg_conf[i][0].weight)
        out[f"b1_{i}"] = _np(head.reg_conf[i][0].bias)
        out[f"w2_{i}"] = _np(head.reg_conf[i][2].weight)
        out[f"b2_{i}"] = _np(head.reg_conf[i][2].bias)
    _save("head", **out)


def _stat(box):
    """Statistic tensor exactly as head.py:232-240 builds it (reference ops, used to expose q)."""
    B, _, H, W = box.shape
    prob = box.view(B, 4, 16, H, W).softmax(dim=2)
    topk = torch.topk(prob, k=4, dim=2).values
    return torch.cat([topk, prob.mean(dim=2, keepdim=True)], dim=2).view(B, -1, H, W)


def _pred(g, B, nc, A, score_scale=1.0, quant=None, img=640.0):
    cxy = torch.rand(B, 2, A, generator=g) * img
    wh = torch.rand(B, 2, A, generator=g) * img * 0.3 + 2.0
    sc = torch.rand(B, nc, A, generator=g) * score_scale
    if quant:
        sc = torch.round(sc * quant) / quant
    return torch.cat((cxy, wh, sc), 1)


def gen_nms(ops):
    """`non_max_suppression`, utils/ops.py:167-316, and `torchvision.ops.nms` at ops.py:296."""
    import torchvision

    g = torch.Generator().manual_seed(5)
    out, cases = {}, {}
    cases["single"] = (_pred(g, 3, 4, 300), dict(conf_thres=0.25, iou_thres=0.45))
    cases["multi"] = (_pred(g, 2, 3, 200), dict(conf_thres=0.001, iou_thres=0.7, multi_label=True))
    cases["ties"] = (_pred(g, 2, 3, 400, quant=8), dict(conf_thres=0.1, iou_thres=0.5, multi_label=True))
    p = _pred(g, 2, 80, 150, score_scale=0.3)
    p[:, 4 + 69 :, :] *= 3.0  # make high class ids (>=69: 1/16-px offset grid) win
    cases["highcls"] = (p, dict(conf_thres=0.25, iou_thres=0.6))
    p = _pred(g, 3, 4, 120)
    p[1, 4:] = 0.01  # image 1 has no candidate
    cases["empty"] = (p, dict(conf_thres=0.25, iou_thres=0.45))
    cases["classes"] = (_pred(g, 2, 6, 200), dict(conf_thres=0.2, iou_thres=0.5, classes=[1, 4]))
    cases["agnostic"] = (_pred(g, 2, 6, 200), dict(conf_thres=0.2, iou_thres=0.5, agnostic=True, multi_label=True))
    cases["maxnms"] = (_pred(g, 2, 3, 300), dict(conf_thres=0.05, iou_thres=0.7, multi_label=True, max_nms=50, max_det=20))
    # boxes whose IoU is exactly the threshold (0.5): 2x1 box vs its left 1x1 half... use (0,0,2,2) vs (0,0,2,1)
    p = torch.zeros(1, 5, 4)
    p[0, :4, 0] = torch.tensor([1.0, 1.0, 2.0, 2.0])  # xywh -> (0,0,2,2)
    p[0, :4, 1] = torch.tensor([1.0, 0.5, 2.0, 1.0])  # (0,0,2,1): IoU 0.5 exactly -> kept
    p[0, :4, 2] = torch.tensor([1.0, 0.75, 2.0, 1.5])  # (0,0,2,1.5): IoU .75 -> dropped
    p[0, :4, 3] = torch.tensor([1.0, 1.0, 2.0, 2.0])  # exact duplicate -> dropped
    p[0, 4] = torch.tensor([0.9, 0.8, 0.7, 0.6])
    cases["exact"] = (p, dict(conf_thres=0.25, iou_thres=0.5))
    for name, (pred, kw) in cases.items():
        res = ops.non_max_suppression(pred.clone(), max_time_img=1e9, **kw)
        out[f"{name}_pred"] = _np(pred)
        out[f"{name}_n"] = np.array([r.shape[0] for r in res], np.int64)
        out[f"{name}_out"] = _np(torch.cat(res, 0)) if sum(r.shape[0] for r in res) else np.zeros((0, 6), np.float32)
        out[f"{name}_kw"] = np.array(repr(kw))
    # raw torchvision boundary: heavily tied scores, boxes on a coarse grid (many exact-IoU ties)
    n = 600
    xy = torch.randint(0, 40, (n, 2), generator=g).float() * 4
    wh = torch.randint(1, 12, (n, 2), generator=g).float() * 4
    boxes = torch.cat((xy, xy + wh), 1)
    scores = torch.randint(0, 20, (n,), generator=g).float() / 20
    out["tv_boxes"], out["tv_scores"] = _np(boxes), _np(scores)
    for thr in (0.3, 0.5, 0.7):
        out[f"tv_keep_{int(thr * 10)}"] = _np(torchvision.ops.nms(boxes, scores, thr))
    _save("nms", **out)


def gen_losses(loss_mod):
    """`quality_focal_loss` loss.py:22-70, `distribution_focal_loss` :88-137, `DFLoss` :209-224."""
    g = torch.Generator().manual_seed(6)
    pred = (torch.randn(64, 7, generator=g) * 3).requires_grad_()
    target = torch.zeros(64, 7)
    rows = torch.randint(0, 64, (20,), generator=g)
    target[rows, torch.randint(0, 7, (20,), generator=g)] = torch.rand(20, generator=g)
    loss = loss_mod.quality_focal_loss(pred, target, beta=2.0)
    (grad,) = torch.autograd.grad(loss.sum(), pred)
    out = {"qfl_pred": _np(pred), "qfl_target": _np(target), "qfl_loss": _np(loss), "qfl_grad": _np(grad)}
    loss15 = loss_mod.quality_focal_loss(pred, target, beta=1.5)
    out["qfl_loss_b15"] = _np(loss15)
    out["qfl_grad_b15"] = _np(torch.autograd.grad(loss15.sum(), pred)[0])

    pd = (torch.randn(40, 16, generator=g) * 2).requires_grad_()
    tgt = torch.rand(10, 4, generator=g) * 17 - 1  # exercises both clamps
    tgt[0, 0], tgt[0, 1] = 3.0, 14.99
    l = loss_mod.DFLoss(16)(pd, tgt.clone())
    out.update({"dfl_pred": _np(pd), "dfl_target": _np(tgt), "dfl_loss": _np(l), "dfl_grad": _np(torch.autograd.grad(l.sum(), pd)[0])})
    l2 = loss_mod.distribution_focal_loss(pd.view(10, 4, 16), tgt.clone())
    out["dfl_fn_loss"] = _np(l2)
    _save("losses", **out)


def gen_detection_loss(loss_mod, head_mod):
    """`v8DetectionLoss.__call__`, utils/loss.py:345-420 (BCE branch: the uniH head leaves `_qualities` None),
    including TaskAlignedAssigner (utils/tal.py:14-295) and BboxLoss (loss.py:227-249)."""
    from types import SimpleNamespace

    torch.manual_seed(7)
    g = torch.Generator().manual_seed(7)
    nc = 5
    head = head_mod.GFLHeadv2_uniH(nc=nc, ch=(16, 32, 64))
    head.stride = torch.tensor([8.0, 16.0, 32.0])

    class Wrapper(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.model = torch.nn.ModuleList([head])
            self.args = SimpleNamespace(box=7.5, cls=0.5, dfl=1.5)

    crit = loss_mod.v8DetectionLoss(Wrapper())
    out = {"nc": np.int64(nc)}
    for tag, B, sizes, boxes in (
        ("a", 2, (8, 4, 2), [[0, 1, .30, .35, .30, .40], [0, 3, .70, .60, .40, .50], [0, 1, .50, .50, .90, .90], [1, 0, .25, .75, .35, .30], [1, 4, .60, .30, .50, .45]]),
        ("b", 3, (6, 3, 2), [[0, 2, .5, .5, .6, .6], [2, 2, .4, .4, .5, .3], [2, 0, .45, .42, .5, .35], [2, 1, .8, .8, .3, .3]]),  # image 1 has no target
    ):
        feats = [(torch.randn(B, 64 + nc, s, s, generator=g) * 1.5).requires_grad_() for s in sizes]
        t = torch.tensor(boxes, dtype=torch.float32)
        batch = {"batch_idx": t[:, 0], "cls": t[:, 1], "bboxes": t[:, 2:]}
        total, items = crit([f for f in feats], batch)
        grads = torch.autograd.grad(total, feats)
        out[f"{tag}_targets"] = _np(t)
        out[f"{tag}_total"] = _np(total)
        out[f"{tag}_items"] = _np(items)
        for i, (f, gr) in enumerate(zip(feats, grads)):
            out[f"{tag}_feat{i}"] = _np(f)
            out[f"{tag}_grad{i}"] = _np(gr)
    _save("detection_loss", **out)


def main():
    ref_loader.load()
    from ultralytics.nn.modules import block, head
    from ultralytics.utils import loss, ops

    os.makedirs(OUT, exist_ok=True)
    gen_dwt(block)
    gen_enhancer(block)
    gen_attention(block)
    gen_head(head)
    gen_nms(ops)
    gen_losses(loss)
    gen_detection_loss(loss, head)


if __name__ == "__main__":
    main()
