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


def _run(golden, tag, device, patch_dfl):
    from edge_yolo_b200.detection_loss import v8DetectionLoss

    g = golden("detection_loss")
    model = _Model(int(g["nc"])).to(device)
    crit = v8DetectionLoss(model)
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
@pytest.mark.parametrize("tag", ["a", "b"])
def test_detection_loss_gpu(golden, tag):
    _run(golden, tag, "cuda", patch_dfl=False)


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
