"""North-star criterion: end-to-end mAP@0.5:0.95 on a FIXED SYNTHETIC validation set within 0.1 points of the reference.

There are no datasets or checkpoints offline, so the set is built the only way that gives a meaningful, non-zero mAP at
random init: seeded synthetic images, and as ground truth the reference semantics' own confident detections on them
(CPU fp32 oracle model, predict mode, top boxes per image).  The reference arm is then the same oracle run in *validation*
mode (conf 0.001, multi_label, iou 0.7, max_det 300 -- `DetectionValidator` defaults, models/yolo/detect/val.py,
cfg/default.yaml); the product arms are (1) the API path: fp32 model on the CUDA kernels + `non_max_suppression`, and
(2) the inference engine: bf16 NHWC CUDA-graph Predictor from uint8.  mAP is computed by oracle/metrics_ref.py, which is
pinned to the reference's validator code (tests/golden/metrics.json, tests/test_reference_model.py).
"""
import copy

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

B, S, NC = 8, 256, 80


def _val_images():
    g = torch.Generator().manual_seed(77)
    low = torch.rand(B, 3, S // 16, S // 16, generator=g)
    img = torch.nn.functional.interpolate(low, size=(S, S), mode="bilinear", align_corners=False)
    img = (img + 0.15 * torch.rand(B, 3, S, S, generator=g)).clamp_(0, 1)
    u8 = (img * 255).round().to(torch.uint8).permute(0, 2, 3, 1).contiguous()  # HWC uint8, what the predictor ingests
    return u8


def test_map_within_a_tenth_of_a_point():
    from edge_yolo_b200 import modules as M
    from edge_yolo_b200.engine import Predictor, build_model
    from edge_yolo_b200.nms import non_max_suppression
    from oracle import metrics_ref, model_ref

    u8 = _val_images()
    x = u8.permute(0, 3, 1, 2).float() / 255
    ref = model_ref.build("n", NC, seed=0)

    # ground truth: the reference's 12 most confident predict-mode detections per image
    labels = []
    for d in model_ref.predict(ref, x, conf=0.25, iou=0.7, max_det=300):
        d = np.asarray(d, dtype=np.float32).reshape(-1, 6)[:12]
        labels.append(np.concatenate([d[:, 5:6], d[:, :4]], 1))
    assert sum(l.shape[0] for l in labels) >= 4 * B, "the synthetic set must carry labels"

    val = dict(conf=0.001, iou=0.7, max_det=300, multi_label=True)
    map_ref, map50_ref = metrics_ref.evaluate(model_ref.predict(ref, x, **val), labels)
    assert map_ref > 0.05, f"degenerate validation set (reference mAP {map_ref})"

    # (1) API path, fp32: product CUDA forwards inside the same module graph + the CUDA NMS
    dev = copy.deepcopy(ref)
    for m in dev.modules():
        if isinstance(m, (M._WaveletEnhancer, M.LinearAttention, M.GFLHeadv2_uniH)):
            del m.forward
    dev = dev.to("cuda").eval()
    with torch.no_grad():
        y, _ = dev(x.to("cuda"))
        dets = non_max_suppression(y, conf_thres=val["conf"], iou_thres=val["iou"], max_det=val["max_det"], multi_label=True)
    map_api, map50_api = metrics_ref.evaluate([d.cpu().numpy() for d in dets], labels)

    # (2) inference engine, bf16 NHWC, uint8 in (same seeded weights)
    eng = build_model("n", NC, seed=0, device="cuda")
    pred = Predictor(eng, batch=B, imgsz=S, conf=val["conf"], iou=val["iou"], max_det=val["max_det"], multi_label=True)
    map_eng, map50_eng = metrics_ref.evaluate([d.numpy() for d in pred.predict(u8.pin_memory())], labels)

    print(f"\nmAP50-95  reference {100 * map_ref:.3f}  api-fp32 {100 * map_api:.3f}  engine-bf16 {100 * map_eng:.3f}   "
          f"mAP50  {100 * map50_ref:.3f} / {100 * map50_api:.3f} / {100 * map50_eng:.3f}")
    # 0.1 points = 0.001 absolute
    assert abs(map_api - map_ref) <= 1e-3, (map_api, map_ref)
    assert abs(map_eng - map_ref) <= 1e-3, (map_eng, map_ref)
