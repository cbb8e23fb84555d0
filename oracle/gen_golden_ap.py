"""Writes tests/golden/ap_per_class.npz: outputs of the UNMODIFIED reference's `ap_per_class` (utils/metrics.py:537-623) and
`scale_boxes` (utils/ops.py:92-127) on seeded inputs (metrics_ref.ap_case; no exact confidence ties, so the reference's unstable
argsort is well defined).  Dev container only (needs /root/reference):  python -m oracle.gen_golden_ap"""
import os

import numpy as np
import torch

from . import metrics_ref, ref_loader

NAMES = ["tp", "fp", "p", "r", "f1", "ap", "classes", "p_curve", "r_curve", "f1_curve", "x", "prec_values"]
SCALE_CASES = [((640, 640), (480, 640), None, True, False), ((384, 640), (720, 1280), None, True, False), ((640, 480), (1000, 750), None, True, True),
               ((640, 640), (333, 500), ((1.28, 1.28), (0.0, 106.5)), True, False), ((320, 320), (640, 640), None, False, False)]


def main():
    ref_loader.load()
    from ultralytics.utils import ops as uops
    from ultralytics.utils.metrics import ap_per_class

    out = {}
    for seed, n_det in [(0, 600), (1, 900), (2, 1)]:
        tp, conf, pc, tc = metrics_ref.ap_case(seed, n_det=n_det, n_lab=200)
        assert np.unique(conf).shape[0] == conf.shape[0]
        res = ap_per_class(tp, conf, pc, tc)
        for name, v in zip(NAMES, res):
            out[f"s{seed}_{name}"] = np.asarray(v)
    rng = np.random.default_rng(5)
    for i, (s1, s0, rp, padding, xywh) in enumerate(SCALE_CASES):
        bx = (rng.random((40, 6)) * 760 - 60).astype(np.float32)
        t = torch.from_numpy(bx.copy())
        uops.scale_boxes(s1, t[:, :4], s0, ratio_pad=rp, padding=padding, xywh=xywh)
        out[f"scale{i}_in"] = bx
        out[f"scale{i}_out"] = t.numpy()
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "ap_per_class.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path))


if __name__ == "__main__":
    main()
