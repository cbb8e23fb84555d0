"""TEST INFRASTRUCTURE: helpers that drive the UNMODIFIED reference package through its own public API
(`YOLO(cfg, task="detect").predict / val / train`, engine/model.py:501, 609, 742) with and without
`edge_yolo_b200.install()` applied.  The reference is found by oracle/ref_loader.py (dev container: /root/reference;
GPU box: baseline/_ref, shipped with the gpurun snapshot).
"""
from __future__ import annotations

import os

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CKPT = os.path.join(ROOT, "tests", "golden", "edgeline_n_synth.pt")


def reference():
    from oracle import ref_loader

    return ref_loader.load(), ref_loader


def model_yaml(tmp_dir: str, scale: str = "n", nc: int | None = None) -> str:
    _, rl = reference()
    return rl.model_yaml(scale, nc, tmp_dir)


def synth_checkpoint_state():
    """The product-trained EdgeLine-YOLO-n checkpoint of tests/test_map_parity.py (8 classes; keys = the reference's `model.{i}.*`)."""
    return {k: (v.float() if v.is_floating_point() else v) for k, v in torch.load(CKPT, map_location="cpu").items()}


def build_yolo(tmp_dir: str, trained: bool = True, scale: str = "n", nc: int = 8, gamma: float | None = None, seed: int = 0):
    """`YOLO(cfg, task="detect")` of the reference (Q1: the task must be given).  trained=True loads the synthetic-task checkpoint."""
    reference()
    from ultralytics import YOLO

    torch.manual_seed(seed)
    m = YOLO(model_yaml(tmp_dir, scale, nc), task="detect", verbose=False)
    if trained:
        missing, unexpected = m.model.load_state_dict(synth_checkpoint_state(), strict=False)
        assert not unexpected and all("num_batches_tracked" in k or k.endswith("dwt.weight") for k in missing), (missing, unexpected)
    if gamma is not None:  # gamma = 0 (the reference's init) makes every wavelet branch an identity
        with torch.no_grad():
            for name, p in m.model.named_parameters():
                if name.endswith("wave.gamma"):
                    p.fill_(gamma)
    return m


def write_dataset(root: str, n_images: int, S: int, seed: int = 1000, sizes=None) -> str:
    """Synthetic detection set on disk in the layout `check_det_dataset` expects (data/utils.py:280-389): lossless PNGs + YOLO txt labels.
    `sizes`: optional list of (h, w) crops so that `rect=True` validation sees several non-square shapes."""
    import cv2

    from tools import synth_data

    os.makedirs(os.path.join(root, "images", "val"), exist_ok=True)
    os.makedirs(os.path.join(root, "labels", "val"), exist_ok=True)
    x, t = synth_data.synth_batch(n_images, S, torch.Generator().manual_seed(seed), "cpu")
    for b in range(n_images):
        img = (x[b] * 255).round().to(torch.uint8).permute(1, 2, 0).numpy()[:, :, ::-1]
        cv2.imwrite(os.path.join(root, "images", "val", f"{b:05d}.png"), img)
        rows = t[t[:, 0] == b][:, 1:]
        with open(os.path.join(root, "labels", "val", f"{b:05d}.txt"), "w") as f:
            for r in rows:
                f.write("%d %.6f %.6f %.6f %.6f\n" % (int(r[0]), *r[1:].tolist()))
    path = os.path.join(root, "data.yaml")
    with open(path, "w") as f:
        f.write(f"path: {root}\ntrain: images/val\nval: images/val\nnames:\n" + "".join(f"  {i}: c{i}\n" for i in range(synth_data.NC)))
    return path


class installed:
    """Context manager: edge_yolo_b200.install(**kw) on entry, uninstall() on exit; `.launches` = library kernels launched inside."""

    def __init__(self, **kw):
        self.kw = kw

    def __enter__(self):
        import edge_yolo_b200.install as el
        from edge_yolo_b200 import _lib

        self._el, self._lib = el, _lib.lib()
        self._n0 = self._lib.el_launch_count()
        self.names = el.install(**self.kw)
        return self

    @property
    def launches(self) -> int:
        return int(self._lib.el_launch_count() - self._n0)

    def __exit__(self, *exc):
        self._el.uninstall()
