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
