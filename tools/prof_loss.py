"""Loss kernels at the size of BASELINE configs[4] (B = 64, 8400 anchors, 80 classes, fp32 logits): CUDA-event time per launch with the
operands rotated through > L2, and achieved GB/s against the algorithmic bytes of SURVEY.md section 8(d).  python tools/prof_loss.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from edge_yolo_b200 import _lib  # noqa: E402
from edge_yolo_b200.ops import _dt, _stream, check  # noqa: E402

L = _lib.lib()
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
n = 64 * 8400 * 80
R = 4
preds = [torch.randn(n, device=dev, generator=g) for _ in range(R)]
tgts = [(torch.rand(n, device=dev, generator=g) > 0.99).float() * torch.rand(n, device=dev, generator=g) for _ in range(R)]
loss = torch.empty(n, device=dev)
grad = torch.empty(n, device=dev)
gout = torch.ones(n, device=dev)
total = torch.empty((), device=dev)
part = torch.empty(L.el_qfl_partials(n), device=dev)
one = torch.ones((), device=dev)


def timeit(fn):
    for i in range(R):
        fn(i)
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda._sleep(2_000_000)
        e0.record()
        for i in range(R):
            fn(i)
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3 / R)
    return sorted(ts)[2]


def report(name, us, nbytes):
    print(f"{name:34s} {us:8.1f} us  {nbytes / 1e6:7.1f} MB  {nbytes / us / 1e3:6.0f} GB/s", flush=True)


report("el_qfl_fwd (elementwise loss out)", timeit(lambda i: check(L.el_qfl_fwd(preds[i].data_ptr(), tgts[i].data_ptr(), n, 2.0, _dt(preds[i]), loss.data_ptr(), None, None, _stream()), "qfl")), 3 * n * 4)
report("el_qfl_fwd (sum reduction)", timeit(lambda i: check(L.el_qfl_fwd(preds[i].data_ptr(), tgts[i].data_ptr(), n, 2.0, _dt(preds[i]), None, total.data_ptr(), part.data_ptr(), _stream()), "qfl")), 2 * n * 4)
report("el_qfl_bwd (of the sum)", timeit(lambda i: check(L.el_qfl_bwd(preds[i].data_ptr(), tgts[i].data_ptr(), n, 2.0, _dt(preds[i]), None, one.data_ptr(), grad.data_ptr(), _stream()), "qfl")), 3 * n * 4)
rows = 64 * 8 * 10  # ~10 foreground anchors per box, 8 boxes per image
pd = [torch.randn(rows * 4, 16, device=dev, generator=g) for _ in range(R)]
td = [torch.rand(rows, 4, device=dev, generator=g) * 14 for _ in range(R)]
dl = torch.empty(rows, 1, device=dev)
dg = torch.empty(rows * 4, 16, device=dev)
go = torch.ones(rows, 1, device=dev)
report("el_dfl_fwd (5120 fg anchors)", timeit(lambda i: check(L.el_dfl_fwd(pd[i].data_ptr(), td[i].data_ptr(), rows, _dt(pd[i]), dl.data_ptr(), _stream()), "dfl")), rows * 4 * (16 * 4 + 4) + rows * 4)
report("el_dfl_bwd", timeit(lambda i: check(L.el_dfl_bwd(pd[i].data_ptr(), td[i].data_ptr(), rows, _dt(pd[i]), go.data_ptr(), dg.data_ptr(), _stream()), "dfl")), rows * 4 * (2 * 16 * 4 + 4))

# ---- task-aligned assigner at the same size (B = 64, 8 boxes per image, 8400 anchors, 80 classes): el_tal_assign (three kernels) against
# the torch-op formulation.  Algorithmic bytes: read scores' gathered column + boxes + anchors per (gt, anchor) candidate at worst, write
# target scores (B*A*nc*4) + boxes / labels / fg / index (B*A*37): the dense target-score write dominates.
from edge_yolo_b200.detection_loss import TaskAlignedAssigner, make_anchors  # noqa: E402

B, M, nc = 64, 8, 80
cpu_g = torch.Generator().manual_seed(1)
pts, st = make_anchors([torch.empty(1, 1, 640 // s, 640 // s) for s in (8, 16, 32)], (8, 16, 32))
anchors = (pts * st).to(dev)
A = anchors.shape[0]
cases = []
for _ in range(R):
    ltrb = torch.rand(B, A, 4, generator=cpu_g) * 6 * st
    boxes = torch.cat((pts * st - ltrb[..., :2], pts * st + ltrb[..., 2:]), -1).to(dev)
    scores = torch.sigmoid(torch.randn(B, A, nc, generator=cpu_g) * 2 - 3).to(dev)
    c, wh = torch.rand(B, M, 2, generator=cpu_g) * 0.8 + 0.1, torch.rand(B, M, 2, generator=cpu_g) * 0.35 + 0.05
    gt_boxes = (torch.cat((c - wh / 2, c + wh / 2), -1) * 640).to(dev)
    gt_labels = torch.randint(0, nc, (B, M, 1), generator=cpu_g).float().to(dev)
    cases.append((scores, boxes, anchors, gt_labels, gt_boxes, torch.ones(B, M, 1, dtype=torch.bool, device=dev)))
tal_bytes = B * A * (nc * 4 + 37) + B * A * 16 + B * M * A // 8 * (16 + 4)
from oracle.tal_torch import TorchTaskAlignedAssigner  # noqa: E402

fused, eager = TaskAlignedAssigner(10, nc, 0.5, 6.0), TorchTaskAlignedAssigner(10, nc, 0.5, 6.0)
report("el_tal_assign (3 kernels)", timeit(lambda i: fused(*cases[i])), tal_bytes)
report("TAL as device-side torch ops", timeit(lambda i: eager(*cases[i])), tal_bytes)
