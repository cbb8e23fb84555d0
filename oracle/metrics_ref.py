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


def smooth(y: np.ndarray, f: float = 0.05) -> np.ndarray:
    """Box filter of fraction f, utils/metrics.py:447-452."""
    nf = round(len(y) * f * 2) // 2 + 1
    p = np.ones(nf // 2)
    yp = np.concatenate((p * y[0], y, p * y[-1]), 0)
    return np.convolve(yp, np.ones(nf) / nf, mode="valid")


def ap_per_class(tp: np.ndarray, conf: np.ndarray, pred_cls: np.ndarray, target_cls: np.ndarray, eps: float = 1e-16):
    """Restatement of `ap_per_class` (utils/metrics.py:537-623, plot=False) with the same 12-tuple.  The only liberty: a STABLE
    argsort by confidence (the reference's default quicksort leaves the order of exactly tied confidences open)."""
    order = np.argsort(-conf, kind="stable")
    tp, conf, pred_cls = tp[order], conf[order], pred_cls[order]
    classes, nt = np.unique(target_cls, return_counts=True)
    nc = classes.shape[0]
    x, prec_values = np.linspace(0, 1, 1000), []
    ap, p_curve, r_curve = np.zeros((nc, tp.shape[1])), np.zeros((nc, 1000)), np.zeros((nc, 1000))
    for ci, c in enumerate(classes):
        sel = pred_cls == c
        if sel.sum() == 0 or nt[ci] == 0:
            continue
        fpc = (1 - tp[sel]).cumsum(0)
        tpc = tp[sel].cumsum(0)
        recall = tpc / (nt[ci] + eps)
        r_curve[ci] = np.interp(-x, -conf[sel], recall[:, 0], left=0)
        precision = tpc / (tpc + fpc)
        p_curve[ci] = np.interp(-x, -conf[sel], precision[:, 0], left=1)
        for j in range(tp.shape[1]):
            mrec = np.concatenate(([0.0], recall[:, j], [1.0]))
            mpre = np.concatenate(([1.0], precision[:, j], [0.0]))
            mpre = np.flip(np.maximum.accumulate(np.flip(mpre)))
            xs = np.linspace(0, 1, 101)
            y = np.interp(xs, mrec, mpre)
            ap[ci, j] = ((y[1:] + y[:-1]) * np.diff(xs) / 2.0).sum()
            if j == 0:
                prec_values.append(np.interp(x, mrec, mpre))
    prec_values = np.array(prec_values)
    f1_curve = 2 * p_curve * r_curve / (p_curve + r_curve + eps)
    i = smooth(f1_curve.mean(0), 0.1).argmax()
    p, r, f1 = p_curve[:, i], r_curve[:, i], f1_curve[:, i]
    tpo = (r * nt).round()
    fpo = (tpo / (p + eps) - tpo).round()
    return tpo, fpo, p, r, f1, ap, classes.astype(int), p_curve, r_curve, f1_curve, x, prec_values


def scale_boxes(img1_shape, boxes: np.ndarray, img0_shape, ratio_pad=None, padding=True, xywh=False) -> np.ndarray:
    """`scale_boxes` + `clip_boxes` (utils/ops.py:92-127, 319-338) on a float32 copy, torch's fp32 arithmetic (subtract, divide, clamp)."""
    b = np.array(boxes, dtype=np.float32, copy=True)
    if ratio_pad is None:
        gain = min(img1_shape[0] / img0_shape[0], img1_shape[1] / img0_shape[1])
        pad = (round((img1_shape[1] - img0_shape[1] * gain) / 2 - 0.1), round((img1_shape[0] - img0_shape[0] * gain) / 2 - 0.1))
    else:
        gain, pad = ratio_pad[0][0], ratio_pad[1]
    if padding:
        b[..., 0] -= np.float32(pad[0])
        b[..., 1] -= np.float32(pad[1])
        if not xywh:
            b[..., 2] -= np.float32(pad[0])
            b[..., 3] -= np.float32(pad[1])
    b[..., :4] /= np.float32(gain)
    b[..., [0, 2]] = b[..., [0, 2]].clip(0, np.float32(img0_shape[1]))
    b[..., [1, 3]] = b[..., [1, 3]].clip(0, np.float32(img0_shape[0]))
    return b


def ap_case(seed: int, n_det: int = 4000, n_lab: int = 900, nc: int = 7, ties: bool = False):
    """Seeded validator statistics: confidences in [0, 1), correlated TP flags that fall off with the IoU threshold, a class without
    predictions and a predicted class without labels; `ties` quantises the confidences (many exact ties)."""
    rng = np.random.default_rng(seed)
    conf = rng.random(n_det).astype(np.float32)
    if ties:
        conf = (np.round(conf * 50) / 50).astype(np.float32)
    pred_cls = rng.integers(0, nc + 1, n_det).astype(np.float32)          # class nc has no labels
    pred_cls[pred_cls == 2] = 3                                             # class 2 has labels but no predictions
    target_cls = rng.integers(0, nc, n_lab).astype(np.float32)
    quality = rng.random(n_det) * (0.4 + 0.6 * conf)
    thr = np.linspace(0.15, 0.75, 10)
    tp = quality[:, None] > thr[None, :]
    return tp, conf, pred_cls, target_cls
