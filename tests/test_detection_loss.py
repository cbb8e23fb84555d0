"""a12: v8DetectionLoss mirror (host orchestration + TaskAlignedAssigner restatement) against the reference-generated golden.
The CPU test swaps the CUDA DFL kernel for a torch expression (the orchestration is what is under test there);
the GPU test runs the real kernels."""
from types import SimpleNamespace

import numpy as np
import pytest
import torch


class _Head(torch.nn.Module):
    def __init__(self, nc):
        super().__init__()
        self.nc, self.reg_max = nc, 16
        self.stride = torch.tensor([8.0, 16.0, 32.0])
        self.p = torch.nn.Parameter(torch.zeros(1))


class _Model(torch.nn.Module):
    def __init__(self, nc):
        super().__init__()
        self.model = torch.nn.ModuleList([_Head(nc)])
        self.args = SimpleNamespace(box=7.5, cls=0.5, dfl=1.5)


def _torch_dfl(pred, target):
    t = target.clamp(0, 14.99)
    tl = t.long()
    wl = (tl + 1).float() - t
    lp = torch.log_softmax(pred.view(-1, 4, 16).float(), -1)
    ce = -(lp.gather(2, tl.unsqueeze(-1)).squeeze(-1) * wl + lp.gather(2, (tl + 1).unsqueeze(-1)).squeeze(-1) * (1 - wl))
    return ce.mean(-1, keepdim=True)


def _run(golden, tag, device, patch_dfl, fused_tal=False):
    from edge_yolo_b200.detection_loss import v8DetectionLoss

    g = golden("detection_loss")
    model = _Model(int(g["nc"])).to(device)
    crit = v8DetectionLoss(model)
    if not fused_tal:  # the torch-op formulation of the assigner is test infrastructure (oracle/tal_torch.py): swapped in through the plain attribute
        from oracle.tal_torch import TorchTaskAlignedAssigner

        crit.assigner = TorchTaskAlignedAssigner(topk=crit.assigner.topk, num_classes=crit.nc, alpha=crit.assigner.alpha, beta=crit.assigner.beta)
    if patch_dfl:
        crit.bbox_loss.dfl = _torch_dfl
    feats = [torch.from_numpy(g[f"{tag}_feat{i}"]).to(device).requires_grad_() for i in range(3)]
    t = torch.from_numpy(g[f"{tag}_targets"])
    batch = {"batch_idx": t[:, 0], "cls": t[:, 1], "bboxes": t[:, 2:]}
    total, items = crit(feats, batch)
    grads = torch.autograd.grad(total, feats)
    np.testing.assert_allclose(items.cpu().numpy(), g[f"{tag}_items"], rtol=2e-5, atol=1e-6)
    np.testing.assert_allclose(float(total), float(g[f"{tag}_total"]), rtol=2e-5)
    for i, gr in enumerate(grads):
        ref = g[f"{tag}_grad{i}"]
        np.testing.assert_allclose(gr.cpu().numpy(), ref, rtol=1e-4, atol=1e-5 * float(np.abs(ref).max()))


@pytest.mark.parametrize("tag", ["a", "b"])
def test_detection_loss_orchestration_cpu(golden, tag):
    _run(golden, tag, "cpu", patch_dfl=True)


@pytest.mark.gpu
@pytest.mark.parametrize("fused_tal", [False, True], ids=["tal_torch_ops", "tal_fused_kernel"])
@pytest.mark.parametrize("tag", ["a", "b"])
def test_detection_loss_gpu(golden, tag, fused_tal):
    """Loss items, total and gradients of the reference (golden), with the assigner as device-side torch ops and as `el_tal_assign`."""
    _run(golden, tag, "cuda", patch_dfl=False, fused_tal=fused_tal)


def _tal_case(B, imgsz, nc, M, seed, device):
    """Trained-like assigner inputs: predicted boxes scattered around their cells, sparse class probabilities, M ground truths per
    image of which the last ones are padding (zero boxes, mask_gt = 0), one image without any target."""
    from edge_yolo_b200.detection_loss import make_anchors

    g = torch.Generator().manual_seed(seed)
    feats = [torch.empty(1, 1, imgsz // s, imgsz // s) for s in (8, 16, 32)]
    pts, st = make_anchors(feats, (8, 16, 32))
    anchors = pts * st
    A = anchors.shape[0]
    ltrb = torch.rand(B, A, 4, generator=g) * 6 * st
    boxes = torch.cat((anchors - ltrb[..., :2], anchors + ltrb[..., 2:]), -1)
    scores = torch.sigmoid(torch.randn(B, A, nc, generator=g) * 2 - 3)
    c = torch.rand(B, M, 2, generator=g) * 0.8 + 0.1
    wh = torch.rand(B, M, 2, generator=g) * 0.35 + 0.05
    gt_boxes = torch.cat((c - wh / 2, c + wh / 2), -1) * imgsz
    gt_labels = torch.randint(0, nc, (B, M, 1), generator=g).float()
    n_valid = torch.randint(1, M + 1, (B,), generator=g)
    n_valid[-1] = 0
    valid = (torch.arange(M)[None, :] < n_valid[:, None]).unsqueeze(-1)
    gt_boxes = gt_boxes * valid
    gt_labels = gt_labels * valid
    return [t.to(device) for t in (scores, boxes, anchors, gt_labels, gt_boxes, valid)]


@pytest.mark.gpu
@pytest.mark.parametrize("B,imgsz,nc,M,topk", [(4, 640, 80, 8, 10), (2, 1280, 10, 5, 13), (3, 160, 3, 1, 10), (3, 128, 80, 12, 10), (3, 256, 4, 16, 13),
                                             (2, 1600, 2, 3, 10),    # 52 500 anchors: the metric row does not fit shared memory
                                             (2, 160, 1203, 6, 10), (2, 128, 365, 4, 10)])  # LVIS / Objects365-sized heads: target rows wider than a CTA
def test_tal_fused_kernel_matches_torch_formulation(B, imgsz, nc, M, topk):
    """`el_tal_assign` (utils/tal.py:14-295 in three kernels) against the torch-op formulation that `test_detection_loss_gpu` pins to the
    reference: identical target scores, and identical labels / boxes / ground-truth indices wherever an anchor carries a non-zero target.
    (Anchors selected with a metric of exactly zero are an open tie in torch.topk; their targets are zero either way.)"""
    from edge_yolo_b200.detection_loss import TaskAlignedAssigner

    args = _tal_case(B, imgsz, nc, M, seed=B * 1000 + M, device="cuda")
    from oracle.tal_torch import TorchTaskAlignedAssigner

    want = TorchTaskAlignedAssigner(topk=topk, num_classes=nc, alpha=0.5, beta=6.0)(*args)
    got = TaskAlignedAssigner(topk=topk, num_classes=nc, alpha=0.5, beta=6.0)(*args)
    l_w, b_w, s_w, fg_w, gi_w = [t.cpu() for t in want]
    l_g, b_g, s_g, fg_g, gi_g = [t.cpu() for t in got]
    assert s_g.shape == s_w.shape and s_g.dtype == s_w.dtype and fg_g.dtype == torch.bool and l_g.dtype == torch.int64 and gi_g.dtype == torch.int64
    np.testing.assert_allclose(s_g.numpy(), s_w.numpy(), rtol=1e-5, atol=1e-8)
    hot = s_w.sum(-1) > 0
    assert int(hot.sum()) > 0 and not bool(hot[-1].any())            # something was assigned; the image without targets stays empty
    assert bool(fg_g[hot].all()) and bool(fg_w[hot].all())
    assert torch.equal(l_g[hot], l_w[hot]) and torch.equal(gi_g[hot], gi_w[hot]) and torch.equal(b_g[hot], b_w[hot])
    bg = ~fg_g & ~fg_w                                               # background anchors carry ground truth 0 like the reference
    assert torch.equal(l_g[bg], l_w[bg]) and torch.equal(b_g[bg], b_w[bg]) and int(gi_g[bg].abs().sum()) == 0
    assert not bool(fg_g[-1].any())


def test_tal_cabi_argument_validation_without_gpu():
    """Workspace query and argument checks of `el_tal_assign` run before any launch (no GPU needed)."""
    import ctypes

    from edge_yolo_b200 import _lib

    L = _lib.lib()
    n = ctypes.c_size_t()
    assert L.el_tal_workspace_bytes(64, 8, 8400, ctypes.byref(n)) == 0
    assert n.value >= 64 * 8 * 8400 * 9 + 64 * 8 * 8           # metric + overlap planes (fp32), flag plane (bytes), per-gt maxima
    assert L.el_tal_workspace_bytes(0, 8, 8400, ctypes.byref(n)) == 1 and L.el_tal_workspace_bytes(1, 1, 1, None) == 1
    null = [None] * 6
    assert L.el_tal_assign(*null, 1, 1, 1, 1, 10, 0.5, 6.0, 1e-9, None, 0, *([None] * 5), None) == 1   # EL_ERR_ARG


@pytest.mark.gpu
def test_detection_loss_qfl_switch_gpu(golden):
    """`use_qfl=True` (the reference's documented one-line switch, loss.py:404-407) runs the QFL kernel and stays finite / differentiable."""
    from edge_yolo_b200.detection_loss import v8DetectionLoss

    g = golden("detection_loss")
    crit = v8DetectionLoss(_Model(int(g["nc"])).cuda(), use_qfl=True)
    feats = [torch.from_numpy(g[f"a_feat{i}"]).cuda().requires_grad_() for i in range(3)]
    t = torch.from_numpy(g["a_targets"])
    total, items = crit(feats, {"batch_idx": t[:, 0], "cls": t[:, 1], "bboxes": t[:, 2:]})
    total.backward()
    assert torch.isfinite(total) and all(torch.isfinite(f.grad).all() for f in feats)
    assert float(items[1]) != float(g["a_items"][1])  # a different class term than BCE


def _tal_golden(golden, tag, device, fused, rtol=1e-5, atol=1e-9):
    """tests/golden/tal.npz: inputs and outputs of the unmodified reference's TaskAlignedAssigner.forward (oracle/gen_golden_tal.py; seeds with
    >= 0.1 % margin on every discrete decision, so the whole assignment must be reproduced exactly and only the scores carry a tolerance)."""
    from edge_yolo_b200.detection_loss import TaskAlignedAssigner

    g = golden("tal")
    B, imgsz, nc, M, topk = (int(v) for v in g[f"{tag}_cfg"])
    alpha, beta = (float(v) for v in g[f"{tag}_alpha_beta"])
    args = [torch.from_numpy(g[f"{tag}_{k}"]).to(device) for k in ("scores", "boxes", "anchors", "gt_labels", "gt_boxes", "gt_valid")]
    from oracle.tal_torch import TorchTaskAlignedAssigner

    cls = TaskAlignedAssigner if fused else TorchTaskAlignedAssigner
    labels, tboxes, tscores, fg, gt_idx = cls(topk=topk, num_classes=nc, alpha=alpha, beta=beta)(*args)
    assert int(g[f"{tag}_out_fg"].sum()) > 0
    np.testing.assert_array_equal(fg.cpu().numpy(), g[f"{tag}_out_fg"])
    np.testing.assert_array_equal(gt_idx.cpu().numpy(), g[f"{tag}_out_gt_idx"])
    np.testing.assert_array_equal(labels.cpu().numpy(), g[f"{tag}_out_labels"])
    np.testing.assert_array_equal(tboxes.cpu().numpy(), g[f"{tag}_out_tboxes"])
    np.testing.assert_allclose(tscores.cpu().numpy(), g[f"{tag}_out_tscores"], rtol=rtol, atol=atol)


@pytest.mark.parametrize("tag", ["multi", "sparse"])
def test_tal_torch_formulation_matches_reference_golden(golden, tag):
    _tal_golden(golden, tag, "cpu", fused=False)


@pytest.mark.gpu
@pytest.mark.parametrize("fused", [False, True], ids=["torch_ops", "fused_kernel"])
@pytest.mark.parametrize("tag", ["multi", "sparse"])
def test_tal_matches_reference_golden_gpu(golden, tag, fused):
    # libdevice powf / atanf against the CPU's: the same tolerance as the loss golden above (the assignment itself is compared exactly)
    _tal_golden(golden, tag, "cuda", fused=fused, rtol=2e-5, atol=1e-8)
