"""TEST INFRASTRUCTURE ONLY -- CPU restatement (numpy fp32, closed form) of the reference's task-aligned assigner.

Follows `TaskAlignedAssigner.forward` of ultralytics/utils/tal.py (no reference code is imported or called):
  select_candidates_in_gts   tal.py:241-262   anchor centre strictly inside the box: min(ltrb deltas) > 1e-9
  get_box_metrics            tal.py:132-155   metric = score[gt class]^alpha * clamp(CIoU(gt, pred), 0)^beta on the candidates
  bbox_iou (CIoU branch)     utils/metrics.py:74-134, xywh=False, eps=1e-7
  select_topk_candidates     tal.py:157-190   the `topk` largest metrics of every valid ground truth
  get_pos_mask               tal.py:120-130   positives = top-k AND inside AND valid
  select_highest_overlaps    tal.py:265-295   multi-claimed anchors go to the ground truth of largest overlap (first maximum)
  get_targets                tal.py:192-238   labels / boxes of the assigned ground truth (ground truth 0 for background), one-hot scores
  normalisation              tal.py:110-116   scores *= metric * best_overlap_of_gt / (best_metric_of_gt + eps)
Written per (image, ground truth) with explicit fp32 rounding after every operation, in the reference's evaluation order.
Ties inside the top-k (possible only between zero metrics) go to the lower anchor index; the reference leaves them to torch.topk.
Pinned to reference outputs by tests/test_oracle_golden.py::test_task_aligned_assigner (tests/golden/tal.npz)."""
from __future__ import annotations

import numpy as np

f32 = np.float32


def ciou(gt: np.ndarray, pred: np.ndarray) -> np.ndarray:
    """gt (4,), pred (n, 4) xyxy -> (n,) complete IoU; every intermediate rounded to fp32 like the reference's eager ops."""
    eps = f32(1e-7)
    ax1, ay1, ax2, ay2 = (f32(v) for v in gt)
    bx1, by1, bx2, by2 = (pred[:, i].astype(f32) for i in range(4))
    aw, ah = f32(ax2 - ax1), f32(f32(ay2 - ay1) + eps)
    bw, bh = bx2 - bx1, (by2 - by1) + eps
    iw = np.maximum(np.minimum(ax2, bx2) - np.maximum(ax1, bx1), f32(0))
    ih = np.maximum(np.minimum(ay2, by2) - np.maximum(ay1, by1), f32(0))
    inter = iw * ih
    union = ((f32(aw * ah) + bw * bh) - inter) + eps
    iou = inter / union
    cw = np.maximum(ax2, bx2) - np.minimum(ax1, bx1)
    ch = np.maximum(ay2, by2) - np.minimum(ay1, by1)
    c2 = (cw * cw + ch * ch) + eps
    dx = ((bx1 + bx2) - ax1) - ax2
    dy = ((by1 + by2) - ay1) - ay2
    rho2 = (dx * dx + dy * dy) / f32(4)
    da = np.arctan(bw / bh).astype(f32) - f32(np.arctan(f32(aw / ah)))
    v = f32(4 / np.pi ** 2) * (da * da)
    alpha = v / ((v - iou) + f32(1 + 1e-7))
    return (iou - (rho2 / c2 + v * alpha)).astype(f32)


def _pow(x: np.ndarray, p: float) -> np.ndarray:
    if p == 1.0:
        return x
    if p == 0.5:
        return np.sqrt(x).astype(f32)
    return np.power(x, f32(p)).astype(f32)


def task_aligned_assign(scores, boxes, anchors, gt_labels, gt_boxes, gt_valid, topk=10, alpha=0.5, beta=6.0, eps=1e-9):
    """scores (B,A,nc), boxes (B,A,4), anchors (A,2), gt_labels (B,M,1) or (B,M), gt_boxes (B,M,4), gt_valid (B,M,1) or (B,M)
    -> labels (B,A) int64, target boxes (B,A,4), target scores (B,A,nc), fg (B,A) bool, gt index (B,A) int64."""
    scores, boxes, anchors, gt_boxes = (np.asarray(t, dtype=f32) for t in (scores, boxes, anchors, gt_boxes))
    B, A, nc = scores.shape
    M = gt_boxes.shape[1]
    lab = np.clip(np.asarray(gt_labels).reshape(B, M).astype(np.int64), 0, nc - 1)
    valid = np.asarray(gt_valid).reshape(B, M).astype(bool)
    metric = np.zeros((B, M, A), f32)
    overlap = np.zeros((B, M, A), f32)
    pos = np.zeros((B, M, A), bool)
    for b in range(B):
        for m in range(M):
            if not valid[b, m]:
                continue
            g = gt_boxes[b, m]
            d = np.minimum(np.minimum(anchors[:, 0] - g[0], anchors[:, 1] - g[1]), np.minimum(g[2] - anchors[:, 0], g[3] - anchors[:, 1]))
            inside = d > f32(1e-9)
            idx = np.nonzero(inside)[0]
            if idx.size:
                u = np.maximum(ciou(g, boxes[b, idx]), f32(0))
                overlap[b, m, idx] = u
                metric[b, m, idx] = _pow(scores[b, idx, lab[b, m]], alpha) * _pow(u, beta)
            order = np.argsort(-metric[b, m], kind="stable")[: min(topk, A)]  # descending, ties -> lower anchor index
            pos[b, m, order] = inside[order]
    claims = pos.sum(1)                                             # (B, A)
    first = np.where(claims > 0, pos.argmax(1), 0)
    gt_idx = np.where(claims > 1, overlap.argmax(1), first).astype(np.int64)
    fg = claims > 0
    bi = np.arange(B)[:, None]
    labels = lab[bi, gt_idx]
    tboxes = gt_boxes[bi, gt_idx]
    ai = np.arange(A)[None, :]
    m_pos = np.where(fg, metric[bi, gt_idx, ai], f32(0))            # metric / overlap of every anchor towards its assigned ground truth
    o_pos = np.where(fg, overlap[bi, gt_idx, ai], f32(0))
    best_m = np.zeros((B, M), f32)
    best_o = np.zeros((B, M), f32)
    np.maximum.at(best_m, (np.broadcast_to(bi, gt_idx.shape)[fg], gt_idx[fg]), m_pos[fg])
    np.maximum.at(best_o, (np.broadcast_to(bi, gt_idx.shape)[fg], gt_idx[fg]), o_pos[fg])
    norm = (m_pos * best_o[bi, gt_idx]) / (best_m[bi, gt_idx] + f32(eps))
    tscores = np.zeros((B, A, nc), f32)
    bb, aa = np.nonzero(fg)
    tscores[bb, aa, labels[bb, aa]] = norm[bb, aa]
    return labels, tboxes, tscores, fg, gt_idx
