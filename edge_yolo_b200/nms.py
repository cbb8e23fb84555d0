"""`non_max_suppression` with the reference's signature (utils/ops.py:167-316), running entirely on the GPU."""
from __future__ import annotations

import torch

from . import ops


def non_max_suppression(prediction, conf_thres=0.25, iou_thres=0.45, classes=None, agnostic=False, multi_label=False, labels=(),
                        max_det=300, nc=0, max_time_img=0.05, max_nms=30000, max_wh=7680, in_place=True, rotated=False):
    """Drop-in for `ultralytics.utils.ops.non_max_suppression` on detection outputs.

    Same arguments, same return type (list of (k, 6) tensors [x1, y1, x2, y2, conf, cls] on the input
    device).  Differences, all deliberate: one host sync per *batch* (to slice the per-image counts)
    instead of several per image; no wall-clock early exit (`max_time_img` is ignored);
    `prediction` is never modified (`in_place` is ignored).  Mask channels (`nm > 0`), apriori
    `labels` and `rotated` boxes are outside the EdgeLine detection path and raise.
    """
    assert 0 <= conf_thres <= 1, f"Invalid Confidence threshold {conf_thres}, valid values are between 0.0 and 1.0"
    assert 0 <= iou_thres <= 1, f"Invalid IoU {iou_thres}, valid values are between 0.0 and 1.0"
    if isinstance(prediction, (list, tuple)):
        prediction = prediction[0]
    if prediction.shape[-1] == 6:  # end-to-end heads already emit (B, N, 6) rows (ops.py:224-228)
        output = [pred[pred[:, 4] > conf_thres][:max_det] for pred in prediction]
        if classes is not None:
            cls_t = torch.tensor(classes, device=prediction.device)
            output = [pred[(pred[:, 5:6] == cls_t).any(1)] for pred in output]
        return output
    if rotated or (labels and any(len(l) for l in labels)):
        raise NotImplementedError("edge_yolo_b200 NMS covers axis-aligned detection without apriori labels")
    nc = nc or (prediction.shape[1] - 4)
    if prediction.shape[1] - nc - 4 != 0:
        raise NotImplementedError("edge_yolo_b200 NMS covers detection heads (no mask channels)")
    out, cnt = ops.nms_batched(prediction, conf_thres, iou_thres, multi_label=multi_label, agnostic=agnostic, classes=classes,
                               max_det=max_det, max_nms=max_nms, max_wh=float(max_wh))
    counts = cnt.tolist()  # the only device->host sync of the call
    return [out[i, :n] for i, n in enumerate(counts)]
