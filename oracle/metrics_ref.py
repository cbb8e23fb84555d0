"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's detection metric (mAP@0.5:0.95).

Follows, in numpy:
  box_iou                      ultralytics/utils/metrics.py:55-71     (pairwise IoU of xyxy boxes, eps 1e-7)
  match_predictions            ultralytics/engine/validator.py:222-262 (non-scipy branch: IoU-descending greedy, one
                               detection per label and one label per detection, per IoU threshold)
  compute_ap / ap_per_class    ultralytics/utils/metrics.py:505-623   (confidence-descending cumulative TP/FP, precision
                               envelope, 101-point interpolated area; mean over classes that have labels)
  IoU thresholds               ultralytics/models/yolo/detect/val.py:37 (linspace(0.5, 0.95, 10))
Only the numbers that enter mAP are reproduced (no P/R/F1 curves, no plots).  Used by tests/test_map_parity.py to check
the north star's "mAP within 0.1 points" criterion; never imported by the product.
"""
from __future__ import annotations

import numpy as np

IOUV = np.linspace(0.5, 0.95, 10)


def box_iou(a: np.ndarray, b: np.ndarray, eps: float = 1e-7) -> np.ndarray:
    """(N,4) x (M,4) xyxy -> (N,M)."""
    a = a.astype(np.float32)[:, None, :]
    b = b.astype(np.float32)[None, :, :]
    wh = np.clip(np.minimum(a[..., 2:], b[..., 2:]) - np.maximum(a[..., :2], b[..., :2]), 0, None)
    inter = wh[..., 0] * wh[..., 1]
    area_a = (a[..., 2] - a[..., 0]) * (a[..., 3] - a[..., 1])
    area_b = (b[..., 2] - b[..., 0]) * (b[..., 3] - b[..., 1])
    return inter / (area_a + area_b - inter + eps)


def match_predictions(pred_cls: np.ndarray, true_cls: np.ndarray, iou_lab_det: np.ndarray) -> np.ndarray:
    """iou_lab_det: (labels, detections).  Returns the (detections, 10) correct matrix."""
    correct = np.zeros((pred_cls.shape[0], IOUV.shape[0]), dtype=bool)
    iou = iou_lab_det * (true_cls[:, None] == pred_cls[None, :])
    for i, thr in enumerate(IOUV.tolist()):
        lab, det = np.nonzero(iou >= thr)
        if lab.size == 0:
            continue
        m = np.stack([lab, det], 1)
        if m.shape[0] > 1:
            m = m[iou[m[:, 0], m[:, 1]].argsort()[::-1]]             # best IoU first
            m = m[np.unique(m[:, 1], return_index=True)[1]]           # each detection once
            m = m[np.unique(m[:, 0], return_index=True)[1]]           # each label once
        correct[m[:, 1], i] = True
    return correct


def _ap(recall: np.ndarray, precision: np.ndarray) -> float:
    mrec = np.concatenate(([0.0], recall, [1.0]))
    mpre = np.concatenate(([1.0], precision, [0.0]))
    mpre = np.maximum.accumulate(mpre[::-1])[::-1]
    x = np.linspace(0, 1, 101)
    y = np.interp(x, mrec, mpre)
    return float(((y[1:] + y[:-1]) * 0.5 * np.diff(x)).sum())         # trapezoid rule (np.trapz)


def mean_ap(tp: np.ndarray, conf: np.ndarray, pred_cls: np.ndarray, target_cls: np.ndarray, eps: float = 1e-16):
    """tp (D,10) bool, conf (D,), pred_cls (D,), target_cls (L,) -> (mAP50-95, mAP50) over the classes that have labels."""
    order = np.argsort(-conf)
    tp, conf, pred_cls = tp[order], conf[order], pred_cls[order]
    classes, n_lab = np.unique(target_cls, return_counts=True)
    ap = np.zeros((classes.shape[0], tp.shape[1]))
    for ci, c in enumerate(classes):
        sel = pred_cls == c
        if sel.sum() == 0 or n_lab[ci] == 0:
            continue
        tpc = tp[sel].cumsum(0)
        fpc = (1 - tp[sel]).cumsum(0)
        recall = tpc / (n_lab[ci] + eps)
        precision = tpc / (tpc + fpc)
        for j in range(tp.shape[1]):
            ap[ci, j] = _ap(recall[:, j], precision[:, j])
    return float(ap.mean()), float(ap[:, 0].mean())


def evaluate(dets_per_image, labels_per_image):
    """dets: list of (k,6) [x1,y1,x2,y2,conf,cls]; labels: list of (m,5) [cls,x1,y1,x2,y2] -> (mAP50-95, mAP50).
    Mirrors DetectionValidator.update_metrics (models/yolo/detect/val.py:125-170) for the box task."""
    tps, confs, pcls, tcls = [], [], [], []
    for det, lab in zip(dets_per_image, labels_per_image):
        det = np.asarray(det, dtype=np.float32).reshape(-1, 6)
        lab = np.asarray(lab, dtype=np.float32).reshape(-1, 5)
        tcls.append(lab[:, 0])
        if det.shape[0] == 0:
            continue
        if lab.shape[0]:
            tp = match_predictions(det[:, 5], lab[:, 0], box_iou(lab[:, 1:], det[:, :4]))
        else:
            tp = np.zeros((det.shape[0], IOUV.shape[0]), dtype=bool)
        tps.append(tp)
        confs.append(det[:, 4])
        pcls.append(det[:, 5])
    if not tps:
        return 0.0, 0.0
    return mean_ap(np.concatenate(tps), np.concatenate(confs), np.concatenate(pcls), np.concatenate(tcls))


def synthetic_case(seed, n_img=6, nc=5):
    """Seeded detections / labels with plenty of near-misses, duplicates and class confusions."""
    rng = np.random.default_rng(seed)
    dets, labs = [], []
    for _ in range(n_img):
        m = int(rng.integers(0, 7))
        xy = rng.uniform(0, 200, (m, 2))
        wh = rng.uniform(20, 80, (m, 2))
        lab = np.concatenate([rng.integers(0, nc, (m, 1)).astype(np.float64), xy, xy + wh], 1)
        k = int(rng.integers(0, 25))
        if m and k:
            src = rng.integers(0, m, k)
            box = lab[src, 1:] + rng.normal(0, 6, (k, 4))
            cls = np.where(rng.random(k) < 0.8, lab[src, 0], rng.integers(0, nc, k))
        else:
            box = np.concatenate([rng.uniform(0, 200, (k, 2)), rng.uniform(200, 280, (k, 2))], 1)
            cls = rng.integers(0, nc, k).astype(np.float64)
        det = np.concatenate([box, rng.random((k, 1)), cls.reshape(-1, 1)], 1)
        dets.append(det.astype(np.float32))
        labs.append(lab.astype(np.float32))
    return dets, labs
