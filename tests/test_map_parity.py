"""North-star criterion: end-to-end mAP@0.5:0.95 on a FIXED SYNTHETIC validation set within 0.1 points of the reference.

There are no datasets or checkpoints offline, so both are made here:
  * task: colour / shape coded rectangles and ellipses on a noisy background, 8 classes (tools/synth_data.py);
  * checkpoint: tests/golden/edgeline_n_synth.pt -- EdgeLine-YOLO-n trained for 150 s on one B200 THROUGH THE PRODUCT'S
    TRAINING PATH (tools/train_synth.py: CUDA forward + backward kernels of DWT / merge / gated residual / linear attention,
    v8DetectionLoss on the DFL kernel + TaskAlignedAssigner, AdamW; log in profiles/r01f_train_synth.log, held-out mAP50-95 93.6);
  * validation set: 256 seeded images (4 batches of 64) with their true labels, evaluated as one set.
Arms, all with the same weights: the reference semantics (CPU fp32 oracle model + oracle NMS, validation settings of
`DetectionValidator`: conf 0.001, multi_label, iou 0.7, max_det 300 -- models/yolo/detect/val.py, cfg/default.yaml) against
(1) the API path: fp32 model on the CUDA kernels + `non_max_suppression`, and (2) the inference engine: bf16 NHWC CUDA-graph
Predictor fed uint8.  mAP comes from oracle/metrics_ref.py, pinned to the reference's validator code
(tests/golden/metrics.json, tests/test_reference_model.py).
"""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

B, S, NB = 64, 256, 4   # 4 batches of 64 images
CKPT = os.path.join(os.path.dirname(__file__), "golden", "edgeline_n_synth.pt")


def test_map_within_a_tenth_of_a_point():
    from edge_yolo_b200.engine import Predictor
    from edge_yolo_b200.model import EdgeLineYOLO
    from edge_yolo_b200.nms import non_max_suppression
    from oracle import metrics_ref, model_ref
    from tools import synth_data

    torch.backends.cudnn.allow_tf32 = False  # the API arm is the fp32 contract; cuDNN's default TF32 convs are 1e-3 off
    torch.backends.cuda.matmul.allow_tf32 = False
    NC = synth_data.NC
    state = {k: (v.float() if v.is_floating_point() else v) for k, v in torch.load(CKPT, map_location="cpu").items()}
    val = dict(conf=0.001, iou=0.7, max_det=300, multi_label=True)
    ref = model_ref.build("n", NC, seed=0)  # reference semantics on the CPU
    ref.load_state_dict(state, strict=True)
    api = EdgeLineYOLO("n", NC).eval()      # (1) API path, fp32: the product's CUDA forwards inside the module graph + the CUDA NMS
    api.load_state_dict(state, strict=True)
    api = api.to("cuda")
    eng = EdgeLineYOLO("n", NC).eval()      # (2) inference engine: bf16 NHWC, fused graph, uint8 in
    eng.load_state_dict(state, strict=True)
    eng = eng.fuse(engine=True).to(device="cuda", dtype=torch.bfloat16).to(memory_format=torch.channels_last)
    pred = Predictor(eng, batch=B, imgsz=S, conf=val["conf"], iou=val["iou"], max_det=val["max_det"], multi_label=True)

    labels, d_ref, d_api, d_eng = [], [], [], []
    for i in range(NB):
        x, targets = synth_data.synth_batch(B, S, torch.Generator().manual_seed(1000 + i), "cpu")
        labels += synth_data.labels_xyxy(targets, B, S)
        u8 = (x * 255).round().to(torch.uint8).permute(0, 2, 3, 1).contiguous()  # HWC uint8, what the predictor ingests
        d_ref += list(model_ref.predict(ref, x, **val))
        with torch.no_grad():
            y, _ = api(x.to("cuda"))
            dets = non_max_suppression(y, conf_thres=val["conf"], iou_thres=val["iou"], max_det=val["max_det"], multi_label=True)
        d_api += [d.cpu().numpy() for d in dets]
        d_eng += [d.numpy().copy() for d in pred.predict(u8.pin_memory())]
    map_ref, map50_ref = metrics_ref.evaluate(d_ref, labels)
    assert map_ref > 0.5, f"checkpoint / validation set mismatch (reference mAP {map_ref})"
    map_api, map50_api = metrics_ref.evaluate(d_api, labels)
    map_eng, map50_eng = metrics_ref.evaluate(d_eng, labels)

    print(f"\nmAP50-95  reference {100 * map_ref:.3f}  api-fp32 {100 * map_api:.3f}  engine-bf16 {100 * map_eng:.3f}   "
          f"mAP50  {100 * map50_ref:.3f} / {100 * map50_api:.3f} / {100 * map50_eng:.3f}")
    # 0.1 points = 0.001 absolute
    assert abs(map_api - map_ref) <= 1e-3, (map_api, map_ref)
    assert abs(map_eng - map_ref) <= 1e-3, (map_eng, map_ref)
