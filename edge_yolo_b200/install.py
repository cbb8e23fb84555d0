"""Drop-in installation behind the unchanged ultralytics API (the reference fork).

`install()` rebinds the *bodies* of the reference's hot-path functions to the CUDA implementations; class
identities, constructor signatures, parameter names and therefore state-dict / pickle compatibility are
untouched, so `YOLO(cfg, task="detect").predict / val`, `ultralytics.nn.modules.*` and
`ultralytics.utils.ops.non_max_suppression` keep working as before:

    import edge_yolo_b200.install as el; el.install()
    model = YOLO("yolo11n-test.yaml", task="detect")      # reference quirk Q1: task must be given

What is replaced (reference path -> ours):
    nn/modules/block.py  _PywtDWT2D.forward          -> modules.dwt_forward
                         _WaveletEnhancer.forward    -> modules.wavelet_enhancer_forward
                         LinearAttention.forward     -> modules.linear_attention_forward
    nn/modules/head.py   GFLHeadv2_uniH.forward      -> modules.gfl_head_forward
    nn/modules/conv.py   WTConv2d.forward            -> modules.WTConv2d.forward (Haar analysis / synthesis on the DWT kernels)
    nn/modules/block.py  HaarDWT2D.forward           -> modules.haar_dwt2d_forward
                         IHaarDWT2D.__init__/forward -> modules.IHaarDWT2D (the reference's class is broken, Q5: installing REPAIRS it, so that
                                                        WaveletMixerMultiLevel / C3AW_MLM of the fork become constructible and run on the DWT kernels)
    utils/ops.py         non_max_suppression         -> nms.non_max_suppression
    utils/loss.py        quality_focal_loss, QualityFocalLoss.forward, distribution_focal_loss, DFLoss.__call__
                                                      -> loss.*
                         v8DetectionLoss (also the name bound in nn/tasks.py) -> detection_loss.v8DetectionLoss
    (metrics=True)
    utils/metrics.py     box_iou (also the name imported by models/yolo/detect/val.py) -> metrics.box_iou
                         ap_per_class (called by DetMetrics.process, metrics.py:803)   -> metrics.ap_per_class (plot=True: the reference's)
    engine/validator.py  BaseValidator.match_predictions (non-scipy branch)           -> metrics.match_predictions
    utils/ops.py         scale_boxes (fp32 CUDA tensors; other callers keep the reference's) -> metrics.scale_boxes
There is no CPU fallback: after install() these functions need CUDA tensors.
"""
from __future__ import annotations

import importlib

_ORIGINALS: dict = {}


def _swap(obj, name, new):
    key = (obj, name)
    if key not in _ORIGINALS:
        _ORIGINALS[key] = getattr(obj, name)
    setattr(obj, name, new)


def _validator_match_predictions(self, pred_classes, true_classes, iou, use_scipy=False):
    """`BaseValidator.match_predictions` (engine/validator.py:222-262); the scipy branch stays the reference's."""
    from . import metrics

    if use_scipy:
        orig = next(v for (obj, name), v in _ORIGINALS.items() if name == "match_predictions")
        return orig(self, pred_classes, true_classes, iou, use_scipy=True)
    return metrics.match_predictions(pred_classes, true_classes, iou, self.iouv)


def _ap_per_class(tp, conf, pred_cls, target_cls, plot=False, **kw):
    """`ap_per_class` (utils/metrics.py:537): the curves on the GPU; a plotting request is the reference's own function (matplotlib)."""
    from . import metrics

    if plot:
        orig = next(v for (obj, name), v in _ORIGINALS.items() if name == "ap_per_class")
        return orig(tp, conf, pred_cls, target_cls, plot=True, **kw)
    return metrics.ap_per_class(tp, conf, pred_cls, target_cls, plot=False, **kw)


def _scale_boxes(img1_shape, boxes, img0_shape, ratio_pad=None, padding=True, xywh=False):
    """`scale_boxes` (utils/ops.py:92): the library's kernel for the tensors it is defined for (fp32, CUDA, unit-stride rows -- the
    predictor's `pred[:, :4]`); numpy arrays and CPU tensors (plotting, dataset utilities) are not this library's inputs and keep
    the reference's code."""
    import torch

    from . import metrics

    if isinstance(boxes, torch.Tensor) and boxes.is_cuda and boxes.dtype == torch.float32 and boxes.ndim == 2 and boxes.stride(1) == 1:
        return metrics.scale_boxes(img1_shape, boxes, img0_shape, ratio_pad=ratio_pad, padding=padding, xywh=xywh)
    orig = next(v for (obj, name), v in _ORIGINALS.items() if name == "scale_boxes")
    return orig(img1_shape, boxes, img0_shape, ratio_pad=ratio_pad, padding=padding, xywh=xywh)


def install(nms: bool = True, modules: bool = True, losses: bool = True, criterion: bool = True, metrics: bool = False):
    """Patch the imported `ultralytics` package in place.  Returns the list of patched names."""
    from . import _lib, detection_loss as el_det, loss as el_loss, modules as M, nms as el_nms

    _lib.lib()  # fail now, loudly, if the CUDA library has not been built
    block = importlib.import_module("ultralytics.nn.modules.block")
    head = importlib.import_module("ultralytics.nn.modules.head")
    uops = importlib.import_module("ultralytics.utils.ops")
    uloss = importlib.import_module("ultralytics.utils.loss")
    done = []
    if modules:
        _swap(block._PywtDWT2D, "forward", M.dwt_forward)
        _swap(block._WaveletEnhancer, "forward", M.wavelet_enhancer_forward)
        _swap(block.LinearAttention, "forward", M.linear_attention_forward)
        _swap(head.GFLHeadv2_uniH, "forward", M.gfl_head_forward)
        conv = importlib.import_module("ultralytics.nn.modules.conv")
        _swap(conv.WTConv2d, "forward", M.WTConv2d.forward)
        _swap(block.HaarDWT2D, "forward", M.haar_dwt2d_forward)
        _swap(block.IHaarDWT2D, "__init__", M.ihaar_dwt2d_init)
        _swap(block.IHaarDWT2D, "forward", M.ihaar_dwt2d_forward)
        done += ["block._PywtDWT2D.forward", "block._WaveletEnhancer.forward", "block.LinearAttention.forward", "head.GFLHeadv2_uniH.forward",
                 "conv.WTConv2d.forward", "block.HaarDWT2D.forward", "block.IHaarDWT2D.__init__", "block.IHaarDWT2D.forward"]
    if nms:
        _swap(uops, "non_max_suppression", el_nms.non_max_suppression)
        done.append("utils.ops.non_max_suppression")
    if losses:
        _swap(uloss, "quality_focal_loss", el_loss.quality_focal_loss)
        _swap(uloss, "distribution_focal_loss", el_loss.distribution_focal_loss)
        _swap(uloss.QualityFocalLoss, "forward", el_loss.QualityFocalLoss.forward)
        _swap(uloss.DistributionFocalLoss, "forward", el_loss.DistributionFocalLoss.forward)
        _swap(uloss.DFLoss, "__call__", el_loss.DFLoss.__call__)
        done += ["utils.loss.quality_focal_loss", "utils.loss.distribution_focal_loss", "utils.loss.QualityFocalLoss.forward",
                 "utils.loss.DistributionFocalLoss.forward", "utils.loss.DFLoss.__call__"]
    if criterion:  # DetectionModel.init_criterion resolves the name in nn/tasks.py at call time (tasks.py:411-413)
        tasks = importlib.import_module("ultralytics.nn.tasks")
        _swap(uloss, "v8DetectionLoss", el_det.v8DetectionLoss)
        _swap(tasks, "v8DetectionLoss", el_det.v8DetectionLoss)
        done += ["utils.loss.v8DetectionLoss", "nn.tasks.v8DetectionLoss"]
    if metrics:  # validator metrics on the device (opt-in: `val` then needs its predictions and labels on the GPU, which they are)
        from . import metrics as el_metrics

        umetrics = importlib.import_module("ultralytics.utils.metrics")
        dval = importlib.import_module("ultralytics.models.yolo.detect.val")
        validator = importlib.import_module("ultralytics.engine.validator")
        _swap(umetrics, "box_iou", el_metrics.box_iou)
        _swap(dval, "box_iou", el_metrics.box_iou)
        _swap(validator.BaseValidator, "match_predictions", _validator_match_predictions)
        _swap(umetrics, "ap_per_class", _ap_per_class)
        _swap(uops, "scale_boxes", _scale_boxes)
        done += ["utils.metrics.box_iou", "models.yolo.detect.val.box_iou", "engine.validator.BaseValidator.match_predictions",
                 "utils.metrics.ap_per_class", "utils.ops.scale_boxes"]
    return done


def uninstall():
    """Restore everything `install()` replaced."""
    for (obj, name), orig in list(_ORIGINALS.items()):
        setattr(obj, name, orig)
    _ORIGINALS.clear()
