"""box_iou + match_predictions kernels against the metric oracle (oracle/metrics_ref.py, pinned to the reference's validator
code by tests/golden/metrics.json and tests/test_reference_model.py).  Bit-exact IoU, identical correct matrices."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _check(det, lab):
    from edge_yolo_b200 import metrics
    from oracle import metrics_ref

    d, l_ = torch.from_numpy(det).cuda(), torch.from_numpy(lab).cuda()
    iou = metrics.box_iou(l_[:, 1:], d[:, :4])
    want_iou = metrics_ref.box_iou(lab[:, 1:], det[:, :4])
    assert iou.shape == want_iou.shape
    assert iou.cpu().numpy().tobytes() == want_iou.astype(np.float32).tobytes()
    got = metrics.match_predictions(d[:, 5], l_[:, 0], iou)
    want = metrics_ref.match_predictions(det[:, 5], lab[:, 0], want_iou) if lab.shape[0] else np.zeros((det.shape[0], 10), bool)
    assert got.dtype == torch.bool and tuple(got.shape) == (det.shape[0], 10)
    np.testing.assert_array_equal(got.cpu().numpy(), want)
    return int(want.sum())


@pytest.mark.parametrize("seed", range(6))
def test_match_predictions_synthetic_cases(seed):
    from oracle import metrics_ref

    hits = 0
    for det, lab in zip(*metrics_ref.synthetic_case(seed)):
        if det.shape[0]:
            hits += _check(det, lab)
    assert hits > 0


def test_match_predictions_full_size_and_edges():
    """max_det detections against many labels of few classes (lots of competing pairs), plus the empty cases."""
    rng = np.random.default_rng(7)
    L, D = 120, 300
    xy = rng.uniform(0, 600, (L, 2)); wh = rng.uniform(10, 60, (L, 2))
    lab = np.concatenate([rng.integers(0, 3, (L, 1)).astype(np.float64), xy, xy + wh], 1).astype(np.float32)
    src = rng.integers(0, L, D)
    det = np.concatenate([lab[src, 1:] + rng.normal(0, 3, (D, 4)), rng.random((D, 1)), lab[src, :1]], 1).astype(np.float32)
    assert _check(det, lab) > 50
    _check(det, lab[:0])                       # no labels: nothing is correct
    from edge_yolo_b200 import metrics

    empty = metrics.match_predictions(torch.zeros(0, device="cuda"), torch.zeros(4, device="cuda"), torch.zeros(4, 0, device="cuda"))
    assert tuple(empty.shape) == (0, 10)
    # mAP computed from the device path equals the oracle's on the same detections
    from oracle import metrics_ref

    dets, labs = metrics_ref.synthetic_case(3)
    tps, confs, pcls, tcls = [], [], [], []
    for dt, lb in zip(dets, labs):
        tcls.append(lb[:, 0])
        if dt.shape[0] == 0:
            continue
        d, l_ = torch.from_numpy(dt).cuda(), torch.from_numpy(lb).cuda()
        tp = metrics.match_predictions(d[:, 5], l_[:, 0], metrics.box_iou(l_[:, 1:], d[:, :4])).cpu().numpy() if lb.shape[0] else np.zeros((dt.shape[0], 10), bool)
        tps.append(tp); confs.append(dt[:, 4]); pcls.append(dt[:, 5])
    got = metrics_ref.mean_ap(np.concatenate(tps), np.concatenate(confs), np.concatenate(pcls), np.concatenate(tcls))
    assert got == metrics_ref.evaluate(dets, labs)


# ------------------------------------------------------------------------------------------ ap_per_class / scale_boxes (f-4)
AP_NAMES = ["tp", "fp", "p", "r", "f1", "ap", "classes", "p_curve", "r_curve", "f1_curve", "x", "prec_values"]


def _golden_ap():
    import os

    return np.load(os.path.join(os.path.dirname(__file__), "golden", "ap_per_class.npz"))


@pytest.mark.parametrize("seed,n_det", [(0, 600), (1, 900), (2, 1)])
def test_ap_per_class_reference_golden(seed, n_det):
    """el_ap_per_class behind the reference's signature against outputs of the reference's own ap_per_class (float64, 1e-9)."""
    from edge_yolo_b200 import metrics
    from oracle import metrics_ref

    g = _golden_ap()
    got = metrics.ap_per_class(*metrics_ref.ap_case(seed, n_det=n_det, n_lab=200))
    assert len(got) == 12
    for name, v in zip(AP_NAMES, got):
        np.testing.assert_allclose(np.asarray(v, dtype=float), g[f"s{seed}_{name}"].astype(float), rtol=1e-9, atol=1e-12, err_msg=name)


@pytest.mark.parametrize("n_det,n_lab,nc,ties", [(20000, 3000, 7, False), (70000, 9000, 80, False), (5000, 800, 5, True), (300, 1, 3, False), (257, 40, 1, False)])
def test_ap_per_class_vs_oracle(n_det, n_lab, nc, ties):
    """Sizes that cross the single-CTA sort (16384 keys) and the 256-wide tiles; exact confidence ties (stable order, as the oracle)."""
    from edge_yolo_b200 import metrics
    from oracle import metrics_ref

    tp, conf, pc, tc = metrics_ref.ap_case(n_det + nc, n_det=n_det, n_lab=n_lab, nc=nc, ties=ties)
    want = metrics_ref.ap_per_class(tp, conf, pc, tc)
    got = metrics.ap_per_class(torch.from_numpy(tp).cuda(), torch.from_numpy(conf).cuda(), torch.from_numpy(pc).cuda(), tc)
    for name, u, v in zip(AP_NAMES, got, want):
        np.testing.assert_allclose(np.asarray(u, dtype=float), np.asarray(v, dtype=float), rtol=1e-9, atol=1e-12, err_msg=name)
    assert n_lab < 100 or float(np.asarray(want[5]).mean()) > 0.01


def test_ap_per_class_empty_inputs():
    from edge_yolo_b200 import metrics

    out = metrics.ap_per_class(np.zeros((0, 10), bool), np.zeros(0, np.float32), np.zeros(0, np.float32), np.array([1.0, 1.0, 4.0], np.float32))
    assert out[5].shape == (2, 10) and float(out[5].sum()) == 0.0 and list(out[6]) == [1, 4]


def test_scale_boxes_reference_golden():
    """In place on the row-strided view the predictor passes (pred[:, :4]); bit-exact against the reference's outputs."""
    from edge_yolo_b200 import metrics
    from oracle import gen_golden_ap

    g = _golden_ap()
    for i, (s1, s0, rp, padding, xywh) in enumerate(gen_golden_ap.SCALE_CASES):
        pred = torch.from_numpy(g[f"scale{i}_in"].copy()).cuda()
        ret = metrics.scale_boxes(s1, pred[:, :4], s0, ratio_pad=rp, padding=padding, xywh=xywh)
        assert ret.data_ptr() == pred.data_ptr()
        assert pred.cpu().numpy().tobytes() == g[f"scale{i}_out"].tobytes(), i
    with pytest.raises(Exception):
        metrics.scale_boxes((640, 640), torch.zeros(3, 4), (480, 640))  # CPU tensor: no fallback
