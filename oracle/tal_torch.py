"""TEST INFRASTRUCTURE ONLY -- the task-aligned assigner (utils/tal.py:14-295) as ~25 device-agnostic torch ops.

This is the formulation `el_tal_assign` (edge_yolo_b200/csrc/tal.cu) was developed against; it is pinned to the reference's own
`TaskAlignedAssigner.forward` by tests/golden/tal.npz and, in the dev container, against the live reference (tests/test_reference_model.py).
It lived in the product package in round 1 (`TaskAlignedAssigner(fused=False)`); it runs on CPU tensors too, i.e. it was a fallback inside the
product, so it moved here.  Used by the tests (A/B against the kernel; CPU host-orchestration test of v8DetectionLoss through
`crit.assigner = TorchTaskAlignedAssigner(...)`) and by tools/prof_loss.py as the eager-torch timing arm.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from edge_yolo_b200.detection_loss import bbox_ciou


class TorchTaskAlignedAssigner:
    def __init__(self, topk=13, num_classes=80, alpha=1.0, beta=6.0, eps=1e-9):
        self.topk, self.num_classes, self.alpha, self.beta, self.eps = topk, num_classes, alpha, beta, eps

    @torch.no_grad()
    def __call__(self, scores, boxes, anchors, gt_labels, gt_boxes, gt_valid):
        """scores (B,A,nc) probabilities, boxes (B,A,4) xyxy px, anchors (A,2) px, gt_labels (B,M,1), gt_boxes (B,M,4),
        gt_valid (B,M,1) -> (labels (B,A), boxes (B,A,4), scores (B,A,nc), fg (B,A) bool, gt index (B,A))."""
        B, A, nc = scores.shape
        M = gt_boxes.shape[1]
        if M == 0:
            z = torch.zeros_like(scores[..., 0])
            return torch.full_like(z, self.num_classes), torch.zeros_like(boxes), torch.zeros_like(scores), z, z
        valid = gt_valid.bool()                                                   # (B,M,1)
        # anchor centre strictly inside the ground-truth box (tal.py:241-262)
        lt = anchors.view(1, 1, A, 2) - gt_boxes[..., None, :2]
        rb = gt_boxes[..., None, 2:] - anchors.view(1, 1, A, 2)
        inside = torch.cat((lt, rb), -1).amin(-1) > 1e-9                          # (B,M,A)
        cand = inside & valid
        # per (gt, anchor): class probability of the gt's class and CIoU, only where the anchor is a candidate (tal.py:132-155)
        lab = gt_labels.squeeze(-1).long().clamp(0, nc - 1)                       # (B,M)
        cls_score = scores.gather(2, lab.unsqueeze(1).expand(B, A, M)).permute(0, 2, 1)   # (B,M,A)
        iou = bbox_ciou(gt_boxes.unsqueeze(2), boxes.unsqueeze(1)).squeeze(-1).clamp_(0)
        zero = torch.zeros((), dtype=iou.dtype, device=iou.device)
        iou = torch.where(cand, iou, zero)
        cls_score = torch.where(cand, cls_score, zero.to(cls_score.dtype))
        metric = cls_score.pow(self.alpha) * iou.pow(self.beta)
        # top-k anchors per ground truth; padded ground truths select nothing (tal.py:157-190, :126-128)
        top = metric.topk(self.topk, dim=-1).indices                              # (B,M,k)
        in_top = torch.zeros_like(metric, dtype=torch.bool).scatter_(2, top, True) & valid
        pos = in_top & inside                                                     # (B,M,A)
        # an anchor claimed by several ground truths goes to the one it overlaps most (tal.py:265-295)
        n_claims = pos.sum(1)                                                     # (B,A)
        best = F.one_hot(iou.argmax(1), M).permute(0, 2, 1).bool()                # (B,M,A)
        pos = torch.where((n_claims > 1).unsqueeze(1), best, pos)
        fg = pos.any(1)
        gt_idx = pos.float().argmax(1)                                            # (B,A); 0 where background
        # targets (tal.py:192-238)
        labels = lab.gather(1, gt_idx)
        tboxes = gt_boxes.gather(1, gt_idx.unsqueeze(-1).expand(B, A, 4))
        tscores = F.one_hot(labels, nc).to(scores.dtype) * fg.unsqueeze(-1)
        # soft labels: metric normalised per ground truth to its best overlap (tal.py:110-116)
        posf = pos.to(metric.dtype)
        metric = metric * posf
        best_metric = metric.amax(-1, keepdim=True)
        best_iou = (iou * posf).amax(-1, keepdim=True)
        norm = (metric * best_iou / (best_metric + self.eps)).amax(1).unsqueeze(-1)
        return labels, tboxes, tscores * norm, fg, gt_idx
