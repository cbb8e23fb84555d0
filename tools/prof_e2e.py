"""Where does the end-to-end step go?  Times the Predictor's pieces with CUDA events: H2D alone, graph alone, serial, pipelined."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from edge_yolo_b200.engine import Predictor, build_model  # noqa: E402

dev = torch.device("cuda", 0)
torch.backends.cudnn.benchmark = True
model = build_model("n", 80, seed=0, device=dev)
pred = Predictor(model, 64, 640)
host = torch.randint(0, 256, (64, 640, 640, 3), dtype=torch.uint8).pin_memory()
pred.predict_u8(host)


def ev():
    return torch.cuda.Event(enable_timing=True)


def timed(fn, n=10):
    fn()
    torch.cuda.synchronize()
    e0, e1 = ev(), ev()
    t0 = time.perf_counter()
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    e1.synchronize()
    return e0.elapsed_time(e1) / n, (time.perf_counter() - t0) * 1e3 / n


print("H2D 78.6 MB            : %.3f ms (wall %.3f)" % timed(lambda: pred.u8.copy_(host, non_blocking=True)))
print("step from u8 (drained) : %.3f ms (wall %.3f)" % timed(lambda: (pred.step_device(), pred.drain())))
print("step from u8 (pipelined): %.3f ms (wall %.3f)" % timed(lambda: pred.step_device(), 20))
print("graph from x           : %.3f ms (wall %.3f)" % timed(lambda: pred.graph_from_x.replay()))
print("D2H rows               : %.3f ms (wall %.3f)" % timed(lambda: pred.host_out.copy_(pred.out, non_blocking=True)))
print("predict_u8 (serial)    : %.3f ms (wall %.3f)" % timed(lambda: pred.predict_u8(host)))
for n in (8, 32):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    pred.predict_many([host] * n)
    torch.cuda.synchronize()
    print(f"predict_many x{n:3d}      : {(time.perf_counter() - t0) * 1e3 / n:.3f} ms / batch")
# copy and compute concurrently on two streams, no dependencies: is the overlap itself slow?
side = torch.cuda.Stream()
stage = torch.empty_like(pred.u8)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(16):
    with torch.cuda.stream(side):
        stage.copy_(host, non_blocking=True)
    pred.step_device()
torch.cuda.synchronize()
print("independent copy||graph: %.3f ms / iter" % ((time.perf_counter() - t0) * 1e3 / 16))
