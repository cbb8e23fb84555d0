"""Helper of test_bench_contract.py (run as a script): executes bench.py's product arm on a machine WITHOUT a GPU by replacing the
device-facing pieces (CUDA events, the engine's Predictor, process-group init) with host stand-ins.  It checks bench.py's control flow and
the JSON contract only -- no number it prints means anything."""
import importlib.util
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.argv = ["bench.py"] + sys.argv[1:]
import torch  # noqa: E402

spec = importlib.util.spec_from_file_location("bench_under_test", os.path.join(ROOT, "bench.py"))
bench = importlib.util.module_from_spec(spec)
spec.loader.exec_module(bench)


class FakeEvent:
    def __init__(self, enable_timing=False):
        self.t = None

    def record(self, *a):
        self.t = time.perf_counter()

    def synchronize(self):
        pass

    def elapsed_time(self, other):
        return (other.t - self.t) * 1e3


class FakePredictor:
    launches_per_step = 7

    def __init__(self, model, batch, imgsz, use_graph=True, conf=0.25, iou=0.7, max_det=300, multi_label=False):
        self.host_out = torch.zeros(batch, max_det, 6)
        self.host_cnt = torch.zeros(batch, dtype=torch.int32)

    def predict_u8(self, host_u8):
        time.sleep(0.002)
        return self.host_out, self.host_cnt

    def step_device(self):
        time.sleep(0.001)

    def load_resident(self, src):
        pass

    def drain(self):
        pass

    def predict_many(self, batches, consume=None):
        for i, _ in enumerate(batches):
            time.sleep(0.001)
            if consume is not None:
                consume(i, self.host_out, self.host_cnt + 3)
        return len(batches)


torch.cuda.Event = FakeEvent
torch.cuda.is_available = lambda: True
torch.cuda.set_device = lambda d: None
_real_device = torch.device
_real_empty = torch.empty
torch.empty = lambda *a, **k: _real_empty(*a, **{kk: (vv if kk != 'device' else 'cpu') for kk, vv in k.items()})
torch.cuda.synchronize = lambda *a, **k: None
torch.Tensor.pin_memory = lambda self: self

import edge_yolo_b200.dist as eld  # noqa: E402
import edge_yolo_b200.engine as engine  # noqa: E402

engine.Predictor = FakePredictor
engine.build_model = lambda *a, **k: None
eld.init = lambda dev: (0, 1)
bench._REAL_STDOUT = 1
bench.product_arm(bench.parse())
