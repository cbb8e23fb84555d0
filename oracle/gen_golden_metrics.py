"""Writes tests/golden/metrics.json: mAP50-95 / mAP50 of the UNMODIFIED reference (box_iou + DetectionValidator.match_predictions
+ ap_per_class) on metrics_ref.synthetic_case(seed).  Dev container only (needs /root/reference):  python -m oracle.gen_golden_metrics"""
import json
import os

import numpy as np
import torch

from . import metrics_ref, ref_loader


def main():
    ref_loader.load()
    from ultralytics.models.yolo.detect.val import DetectionValidator
    from ultralytics.utils.metrics import ap_per_class, box_iou

    v = DetectionValidator.__new__(DetectionValidator)
    v.iouv = torch.linspace(0.5, 0.95, 10)
    out = {}
    for seed in range(6):
        dets, labs = metrics_ref.synthetic_case(seed)
        tps, confs, pcls, tcls = [], [], [], []
        for det, lab in zip(dets, labs):
            tcls.append(lab[:, 0])
            if det.shape[0] == 0:
                continue
            d, l_ = torch.from_numpy(det), torch.from_numpy(lab)
            tp = v.match_predictions(d[:, 5], l_[:, 0], box_iou(l_[:, 1:], d[:, :4])).numpy() if lab.shape[0] else np.zeros((det.shape[0], 10), bool)
            tps.append(tp); confs.append(det[:, 4]); pcls.append(det[:, 5])
        ap = ap_per_class(np.concatenate(tps), np.concatenate(confs), np.concatenate(pcls), np.concatenate(tcls))[5]
        out[str(seed)] = {"map": float(ap.mean()), "map50": float(ap[:, 0].mean())}
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "metrics.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    print(path, out)


if __name__ == "__main__":
    main()
