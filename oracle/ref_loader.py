"""TEST INFRASTRUCTURE ONLY -- never imported by the product package.

Imports the *unmodified* reference `ultralytics` package.  Two locations are searched:

  1. `/root/reference` (the read-only mount of the dev container) -- used to generate the golden
     vectors (`gen_golden.py`);
  2. `<repo>/baseline/_ref` -- the same package pip-installed by `__graft_entry__.build()`
     (`pip install --no-index --no-deps --target baseline/_ref`, byte-identical .py / .yaml files).
     That directory is git-ignored but ships to the GPU box with the gpurun snapshot, which is what lets
     the `-m gpu` tests run `install()` behind the real `YOLO(cfg).predict / val / train` next to the
     uninstalled reference on the same B200, and `bench.py --impl reference` time the unmodified reference.

Nothing in the product package imports this module.

The reference needs three packages that are absent from this image
(SURVEY.md section 8c): matplotlib (utils/__init__.py:23), pywt (nn/modules/block.py:12,
conv.py:5) and thop (nn/tasks.py:10).  A meta-path finder hands out empty stub
modules for them; `pywt.Wavelet("haar")` returns PyWavelets' published Haar taps
(dec_lo=[s,s], dec_hi=[-s,s], s=1/sqrt(2)) because `_PywtDWT2D.__init__`
(block.py:3597-3599) reads them.
"""
import importlib.abc
import importlib.machinery
import math
import os
import sys
import types

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_CANDIDATES = [os.environ.get("EDGELINE_REFERENCE_ROOT"), "/root/reference", os.path.join(_REPO, "baseline", "_ref")]
REFERENCE_ROOT = next((p for p in _CANDIDATES if p and os.path.isdir(os.path.join(p, "ultralytics"))), "/root/reference")
CFG_DIR = os.path.join(REFERENCE_ROOT, "ultralytics", "cfg", "models", "11")
_STUBBED = ("matplotlib", "pywt", "thop", "seaborn")


class _Wavelet:
    def __init__(self, name):
        if name not in ("haar", "db1"):  # PyWavelets: db1 is the Haar wavelet (WTConv2d's default wt_type, conv.py:487)
            raise ValueError("stub pywt only provides the Haar wavelet")
        s = 1.0 / math.sqrt(2.0)
        self.dec_lo, self.dec_hi = [s, s], [-s, s]
        self.rec_lo, self.rec_hi = [s, s], [s, -s]


class _StubLoader(importlib.abc.Loader):
    def create_module(self, spec):
        m = types.ModuleType(spec.name)
        m.__path__ = []  # behave like a package so `import matplotlib.pyplot` resolves
        if spec.name == "pywt":
            m.Wavelet = _Wavelet
        if spec.name == "thop":
            m.profile = lambda *a, **k: (0.0, 0.0)
        if spec.name == "matplotlib.font_manager":  # utils/checks.py:325 (check_font) scans the system fonts before it tries to download
            m.findSystemFonts = lambda *a, **k: []
        return m

    def exec_module(self, module):
        pass


class _StubFinder(importlib.abc.MetaPathFinder):
    def find_spec(self, fullname, path=None, target=None):
        if fullname.split(".")[0] in _STUBBED:
            return importlib.machinery.ModuleSpec(fullname, _StubLoader(), is_package=True)
        return None


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "ultralytics"))


def load():
    """Return the imported reference `ultralytics` package (CPU)."""
    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    os.environ.setdefault("YOLO_CONFIG_DIR", "/tmp/edgeline_yolo_cfg")
    os.environ.setdefault("OMP_NUM_THREADS", str(os.cpu_count() or 1))  # ultralytics/__init__.py:7-9 forces 1
    os.makedirs(os.environ["YOLO_CONFIG_DIR"], exist_ok=True)
    for font in ("Arial.ttf", "Arial.Unicode.ttf"):  # check_det_dataset -> check_font would try to download them (no network); plots stay off
        open(os.path.join(os.environ["YOLO_CONFIG_DIR"], font), "ab").close()
    sys.dont_write_bytecode = True  # the reference mount is read-only
    if not any(isinstance(f, _StubFinder) for f in sys.meta_path):
        sys.meta_path.append(_StubFinder())
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import ultralytics  # noqa: E402

    return ultralytics


def model_yaml(scale: str = "n", nc: int | None = None, tmp_dir: str | None = None) -> str:
    """Path that `YOLO(...)` / `DetectionModel(...)` accept for EdgeLine-YOLO-<scale> (cfg/models/11/yolo11-test.yaml).  With `nc`, a copy of
    the reference's yaml whose only change is the class count: yaml_model_load strips the scale letter from the file name and opens
    `yolo11-test.yaml` beside it (nn/tasks.py:1150-1180), so the copy is written under that name into `tmp_dir`."""
    import re

    if nc is None:
        return os.path.join(CFG_DIR, f"yolo11{scale}-test.yaml")
    with open(os.path.join(CFG_DIR, "yolo11-test.yaml")) as f:
        text = f.read()
    text, n = re.subn(r"(?m)^nc:\s*\d+", f"nc: {nc}", text, count=1)
    assert n == 1, "nc line not found in the reference yaml"
    tmp_dir = tmp_dir or os.path.join(os.environ.get("YOLO_CONFIG_DIR", "/tmp/edgeline_yolo_cfg"), f"nc{nc}")
    os.makedirs(tmp_dir, exist_ok=True)
    with open(os.path.join(tmp_dir, "yolo11-test.yaml"), "w") as f:
        f.write(text)
    return os.path.join(tmp_dir, f"yolo11{scale}-test.yaml")
