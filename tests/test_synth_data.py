"""The synthetic detection task behind the mAP parity test and the trained checkpoint (tools/synth_data.py): seeded generation is
reproducible, labels describe the drawn objects, images are exactly representable as uint8."""
import hashlib

import numpy as np
import torch

from tools import synth_data


def test_seeded_generation_is_reproducible():
    x, t = synth_data.synth_batch(4, 64, torch.Generator().manual_seed(1000), "cpu")
    x2, t2 = synth_data.synth_batch(4, 64, torch.Generator().manual_seed(1000), "cpu")
    assert torch.equal(x, x2) and torch.equal(t, t2)
    assert tuple(x.shape) == (4, 3, 64, 64) and t.shape[1] == 6
    # pinned: the validation set of tests/test_map_parity.py must not drift with the torch version
    digest = hashlib.sha256((x * 255).round().to(torch.uint8).numpy().tobytes()).hexdigest()[:16]
    assert digest == "c16984f662faf719", digest
    np.testing.assert_allclose(t[0].numpy(), [0.0, 6.0, 0.328621, 0.755239, 0.395373, 0.305334], atol=1e-6)


def test_images_are_uint8_exact_and_labels_in_range():
    x, t = synth_data.synth_batch(8, 96, torch.Generator().manual_seed(3), "cpu")
    assert float(x.min()) >= 0 and float(x.max()) <= 1
    assert torch.equal((x * 255).round() / 255, x)
    assert int(t[:, 0].max()) < 8 and int(t[:, 1].max()) < synth_data.NC and float(t[:, 1].min()) >= 0
    cx, cy, w, h = t[:, 2], t[:, 3], t[:, 4], t[:, 5]
    assert bool(((cx - w / 2) >= -1e-6).all() and ((cx + w / 2) <= 1 + 1e-6).all() and ((cy - h / 2) >= -1e-6).all() and ((cy + h / 2) <= 1 + 1e-6).all())
    counts = torch.bincount(t[:, 0].long(), minlength=8)
    assert int(counts.min()) >= 1 and int(counts.max()) <= synth_data.KMAX
    labs = synth_data.labels_xyxy(t, 8, 96)
    assert sum(l.shape[0] for l in labs) == t.shape[0] and all(l.shape[1] == 5 for l in labs)


def test_last_object_is_drawn_in_its_class_colour():
    """The top-most object of every image shows its class colour at its centre (rectangles and ellipses both cover their centre)."""
    x, t = synth_data.synth_batch(16, 128, torch.Generator().manual_seed(9), "cpu")
    hues = synth_data._HUES
    for b in range(16):
        r = t[t[:, 0] == b][-1]
        px = x[b, :, int(r[3] * 128), int(r[2] * 128)]
        hue = hues[int(r[1]) % 4]
        ratio = px / hue
        assert float(ratio.max() - ratio.min()) < 0.02 and 0.7 <= float(ratio.mean()) <= 1.01
