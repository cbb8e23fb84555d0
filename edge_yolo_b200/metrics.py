"""Validator metrics on the device: drop-ins for `ultralytics.utils.metrics.box_iou`
(utils/metrics.py:55-71), the non-scipy branch of `DetectionValidator.match_predictions` (engine/validator.py:222-262),
`ap_per_class` / `compute_ap` (utils/metrics.py:505-623) and `scale_boxes` / `clip_boxes` (utils/ops.py:92-127, 319-338).
The reference copies the IoU matrix to the host and loops over the ten thresholds in numpy for every image, and computes the
precision / recall curves of the whole validation set in numpy; here each step is a kernel (or a short chain of kernels) and
results stay on the device until the caller asks for them.  No CPU fallback."""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from . import _lib
from .ops import EdgelineError, _need_cuda, _stream, check

_IOUV = {}


def box_iou(box1: torch.Tensor, box2: torch.Tensor, eps: float = 1e-7) -> torch.Tensor:
    """(N,4) x (M,4) xyxy -> (N,M) fp32: inter / (area1 + area2 - inter + eps)."""
    _need_cuda(box1, box2)
    if box1.ndim != 2 or box2.ndim != 2 or box1.shape[1] != 4 or box2.shape[1] != 4:
        raise EdgelineError("box_iou: (N,4) and (M,4) boxes expected")
    a = box1.float() if box1.dtype != torch.float32 else box1
    b = box2.float() if box2.dtype != torch.float32 else box2
    if a.stride(1) != 1:
        a = a.contiguous()
    if b.stride(1) != 1:
        b = b.contiguous()
    N, M = a.shape[0], b.shape[0]
    out = torch.empty((N, M), device=a.device, dtype=torch.float32)
    if N and M:
        check(_lib.lib().el_box_iou(a.data_ptr(), a.stride(0), b.data_ptr(), b.stride(0), out.data_ptr(), N, M, float(eps), _stream()), "el_box_iou")
    return out


def match_predictions(pred_classes: torch.Tensor, true_classes: torch.Tensor, iou: torch.Tensor, iouv: torch.Tensor | None = None) -> torch.Tensor:
    """pred_classes (D,), true_classes (L,), iou (L, D) -> correct (D, T) bool on the device; iouv defaults to linspace(0.5, 0.95, 10)
    (models/yolo/detect/val.py:37)."""
    _need_cuda(pred_classes, true_classes, iou)
    D, L = pred_classes.shape[0], true_classes.shape[0]
    if tuple(iou.shape) != (L, D):
        raise EdgelineError(f"match_predictions: iou must be (labels, detections) = ({L}, {D}), got {tuple(iou.shape)}")
    dev = pred_classes.device
    if iouv is None:
        iouv = _IOUV.get(dev)
        if iouv is None:
            iouv = _IOUV[dev] = torch.linspace(0.5, 0.95, 10, device=dev)
    iouv = iouv.to(device=dev, dtype=torch.float32).contiguous()
    T = iouv.numel()
    correct = torch.zeros((D, T), device=dev, dtype=torch.uint8)
    if D:
        # keep the converted operands alive until the launch is enqueued (a temporary freed before the next allocation would alias it)
        iou_c, pc, tc = iou.float().contiguous(), pred_classes.float().contiguous(), true_classes.float().contiguous()
        check(_lib.lib().el_match_predictions(iou_c.data_ptr(), pc.data_ptr(), tc.data_ptr(), iouv.data_ptr(), L, D, T, correct.data_ptr(), _stream()),
              "el_match_predictions")
    return correct.bool()


def scale_boxes(img1_shape, boxes, img0_shape, ratio_pad=None, padding=True, xywh=False):
    """`ultralytics.utils.ops.scale_boxes` (utils/ops.py:92-127) + `clip_boxes` (:319-338) for CUDA tensors of (..., >= 4) fp32 rows,
    in place like the reference (the predictor passes `pred[:, :4]`, a row-strided view): one kernel, no temporaries.  gain and pad
    are the reference's host arithmetic, including Python's round-half-even."""
    if not isinstance(boxes, torch.Tensor) or not boxes.is_cuda:
        raise EdgelineError("scale_boxes: CUDA tensor expected (no CPU fallback)")
    if boxes.dtype != torch.float32:
        raise EdgelineError(f"scale_boxes: fp32 boxes expected, got {boxes.dtype}")
    if boxes.shape[-1] < 4:
        raise EdgelineError("scale_boxes: rows of at least 4 values expected")
    if ratio_pad is None:  # calculate from img0_shape
        gain = min(img1_shape[0] / img0_shape[0], img1_shape[1] / img0_shape[1])  # gain  = old / new
        pad = (round((img1_shape[1] - img0_shape[1] * gain) / 2 - 0.1), round((img1_shape[0] - img0_shape[0] * gain) / 2 - 0.1))  # wh padding
    else:
        gain = ratio_pad[0][0]
        pad = ratio_pad[1]
    n = boxes.numel() // boxes.shape[-1] if boxes.numel() else 0
    if n:
        rows = boxes if boxes.ndim == 2 else boxes.reshape(-1, boxes.shape[-1])  # a view for every layout the reference is called with
        if rows.data_ptr() != boxes.data_ptr() or rows.stride(1) != 1:
            raise EdgelineError("scale_boxes: rows must be unit-stride views of the caller's storage (in-place contract)")
        check(_lib.lib().el_scale_boxes(rows.data_ptr(), n, rows.stride(0) if n > 1 else rows.shape[1], float(pad[0]), float(pad[1]), float(gain),
                                        int(bool(padding)), int(bool(xywh)), float(img0_shape[1]), float(img0_shape[0]), _stream()), "el_scale_boxes")
    return boxes


def smooth(y, f=0.05):
    """Box filter of fraction f (utils/metrics.py:447-452); host, O(len(y))."""
    nf = round(len(y) * f * 2) // 2 + 1  # number of filter elements (must be odd)
    p = np.ones(nf // 2)  # ones padding
    yp = np.concatenate((p * y[0], y, p * y[-1]), 0)  # y padded
    return np.convolve(yp, np.ones(nf) / nf, mode="valid")  # y-smoothed


def ap_per_class_device(tp: torch.Tensor, conf: torch.Tensor, pred_cls: torch.Tensor, classes: torch.Tensor, n_labels: torch.Tensor, eps: float = 1e-16):
    """Device part of `ap_per_class`: tp (N, T) bool / uint8, conf (N), pred_cls (N) on the GPU; classes (nc) ascending unique target
    classes and their label counts.  Returns float64 device tensors (ap (nc, T), p_curve, r_curve, prec_values (nc, 1000)) and
    n_pred (nc) int32.  No synchronisation."""
    _need_cuda(tp, conf, pred_cls, classes, n_labels)
    N, T = tp.shape
    nc = classes.numel()
    dev = tp.device
    tp8 = tp.to(torch.uint8).contiguous()
    cf, pc = conf.float().contiguous(), pred_cls.float().contiguous()
    cls_f, nl = classes.float().contiguous(), n_labels.to(torch.int64).contiguous()
    L = _lib.lib()
    need = ctypes.c_size_t()
    check(L.el_ap_per_class_workspace_bytes(N, T, nc, ctypes.byref(need)), "el_ap_per_class_workspace_bytes")
    ws = torch.empty(max(need.value, 1), device=dev, dtype=torch.uint8)
    ap = torch.empty((nc, T), device=dev, dtype=torch.float64)
    p_curve, r_curve, prec = (torch.empty((nc, 1000), device=dev, dtype=torch.float64) for _ in range(3))
    n_pred = torch.empty((nc,), device=dev, dtype=torch.int32)
    check(L.el_ap_per_class(tp8.data_ptr(), cf.data_ptr(), pc.data_ptr(), N, T, cls_f.data_ptr(), nl.data_ptr(), nc, float(eps), ws.data_ptr(), need.value,
                            ap.data_ptr(), p_curve.data_ptr(), r_curve.data_ptr(), prec.data_ptr(), n_pred.data_ptr(), _stream()), "el_ap_per_class")
    return ap, p_curve, r_curve, prec, n_pred


def ap_per_class(tp, conf, pred_cls, target_cls, plot=False, on_plot=None, save_dir=None, names={}, eps=1e-16, prefix="", device=None):
    """`ultralytics.utils.metrics.ap_per_class` (utils/metrics.py:537-623) with the reference's signature and return tuple (numpy arrays,
    what `DetMetrics.process` consumes).  Inputs may be numpy arrays (what the validator passes, `val.py:196`) or tensors; the sort by
    confidence, the per-class cumulative sums, the interpolations and the AP integrals run on the GPU (`el_ap_per_class`), the
    O(nc x 1000) tail is the reference's numpy arithmetic on the returned curves.  Plotting is not reproduced (`plot=True` raises)."""
    if plot:
        raise EdgelineError("ap_per_class: plotting is outside the device path; call the reference's plot helpers on the returned curves")
    dev = torch.device(device) if device is not None else (tp.device if isinstance(tp, torch.Tensor) and tp.is_cuda else torch.device("cuda", torch.cuda.current_device()))
    to_dev = lambda a, dt: (a if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(a))).to(device=dev, dtype=dt, non_blocking=True)
    tcls = target_cls.detach().cpu().numpy() if isinstance(target_cls, torch.Tensor) else np.asarray(target_cls)
    unique_classes, nt = np.unique(tcls, return_counts=True)  # host: the labels arrive as a host array and nc is tiny
    nc = unique_classes.shape[0]
    tp_t = to_dev(tp, torch.uint8)
    T = tp_t.shape[1]
    x = np.linspace(0, 1, 1000)
    if nc == 0 or tp_t.shape[0] == 0:
        ap, p_curve, r_curve = np.zeros((nc, T)), np.zeros((nc, 1000)), np.zeros((nc, 1000))
        prec_values = np.array([])
    else:
        ap_d, p_d, r_d, prec_d, np_d = ap_per_class_device(tp_t, to_dev(conf, torch.float32), to_dev(pred_cls, torch.float32),
                                                           to_dev(unique_classes, torch.float32), to_dev(nt, torch.int64), eps)
        ap, p_curve, r_curve, prec_all, n_pred = (t.cpu().numpy() for t in (ap_d, p_d, r_d, prec_d, np_d))
        valid = (n_pred > 0) & (nt > 0)  # the reference appends a row only for classes with predictions and labels
        prec_values = prec_all[valid] if valid.any() else np.array([])
    # ---- tail, metrics.py:602-622
    f1_curve = 2 * p_curve * r_curve / (p_curve + r_curve + eps)
    i = smooth(f1_curve.mean(0), 0.1).argmax() if nc else 0  # max F1 index
    p, r, f1 = p_curve[:, i], r_curve[:, i], f1_curve[:, i]  # max-F1 precision, recall, F1 values
    tp_out = (r * nt).round()  # true positives
    fp_out = (tp_out / (p + eps) - tp_out).round()  # false positives
    return tp_out, fp_out, p, r, f1, ap, unique_classes.astype(int), p_curve, r_curve, f1_curve, x, prec_values
