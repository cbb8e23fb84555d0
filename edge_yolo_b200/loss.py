"""GFL losses with the reference's signatures (utils/loss.py:22-150, 201-224), CUDA forward + backward."""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops


def quality_focal_loss(pred, target, beta: float = 2.0, reduction: str = "none"):
    """utils/loss.py:22-70."""
    if target.shape != pred.shape:
        target = target.expand_as(pred)
    return ops.quality_focal_loss(pred, target, beta, reduction)


class QualityFocalLoss(nn.Module):
    """utils/loss.py:73-85."""

    def __init__(self, beta: float = 2.0, reduction: str = "none"):
        super().__init__()
        self.beta, self.reduction = beta, reduction

    def forward(self, pred, target):
        return quality_focal_loss(pred, target, beta=self.beta, reduction=self.reduction)


def distribution_focal_loss(pred, target, reduction: str = "none"):
    """utils/loss.py:88-137: per-side loss (no mean over the 4 sides).  target (...,) in bins, pred (..., 16).
    Clamps `target` in place like the reference."""
    reg_max = pred.size(-1)
    if reg_max != 16:
        raise NotImplementedError("edge_yolo_b200 DFL kernels are compiled for reg_max=16")
    target.clamp_(0, reg_max - 1 - 0.01)
    loss = ops.dfl_side_loss(pred, target)  # el_dfl_side_fwd / _bwd: the DFL row kernel without the mean over a row's 4 sides
    if reduction == "mean":
        return loss.mean()
    if reduction == "sum":
        return loss.sum()
    return loss


class DistributionFocalLoss(nn.Module):
    """utils/loss.py:140-150."""

    def __init__(self, reduction: str = "none"):
        super().__init__()
        self.reduction = reduction

    def forward(self, pred, target):
        return distribution_focal_loss(pred, target, reduction=self.reduction)


class DFLoss(nn.Module):
    """utils/loss.py:201-224: pred_dist (4n, reg_max), target (n, 4) -> (n, 1).  Clamps `target` in place like the reference."""

    def __init__(self, reg_max=16) -> None:
        super().__init__()
        self.reg_max = reg_max

    def __call__(self, pred_dist, target):
        if self.reg_max != 16:
            raise NotImplementedError("edge_yolo_b200 DFL kernels are compiled for reg_max=16")
        target.clamp_(0, self.reg_max - 1 - 0.01)
        return ops.dfl_loss(pred_dist, target)
