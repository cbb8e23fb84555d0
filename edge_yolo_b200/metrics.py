"""Validator metrics on the device: drop-ins for `ultralytics.utils.metrics.box_iou`
(utils/metrics.py:55-71) and the non-scipy branch of `DetectionValidator.match_predictions` (engine/validator.py:222-262).
The reference copies the IoU matrix to the host and loops over the ten thresholds in numpy for every image; here both steps
are one kernel each and the result stays on the device (no synchronisation).  No CPU fallback."""
from __future__ import annotations

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
