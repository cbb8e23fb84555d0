#!/usr/bin/env python
"""EdgeLine-YOLO hot-path benchmark (contract: see the task brief).

  python bench.py [--gpus N --steps K --warmup W]      product arm (CUDA kernels through the C ABI)
  python bench.py --impl reference ...                 reference arm: the UNMODIFIED reference (baseline/_ref, `YOLO(cfg).predict`)
                                                       on the host cores (falls back to the CPU oracle port if it is absent)

Workload (BASELINE.json configs[1]): EdgeLine-YOLO-n inference, synthetic 640x640, batch 64 per GPU, bf16,
8400 anchors, 80 classes; one step = preprocess + model forward + decode + NMS (predict defaults:
conf 0.25, iou 0.7, max_det 300) over one batch.  N > 1: one process per GPU (torchrun), each with its
own replica and its own batch shard -- weak scaling, no data-path collective.

Besides the contract keys the line carries, measured in the same run: `sustained` (>= 2 s of back-to-back steps with the clocks
seen), `extra_configs` (BASELINE.json configs[2] s@1280 stress NMS, configs[3] m batch 512 sharded 512/N = strong scaling,
configs[4] DDP training step with its all-reduce share) and, at N = 1, `reference_eager_gpu` (the unmodified reference running
eagerly on the same B200 through `YOLO.predict`).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

UNIT = "images/s"


def metric(a):
    """BASELINE.json's metric, named for the configuration actually run (default: configs[1], EdgeLine-YOLO-n at 640x640)."""
    return f"images/sec (EdgeLine-YOLO-{a.scale}, {a.imgsz}x{a.imgsz}, bf16)"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="edgeline", choices=["edgeline", "reference"])
    ap.add_argument("--scale", default="n")
    ap.add_argument("--batch", type=int, default=64, help="images per GPU per step")
    ap.add_argument("--imgsz", type=int, default=640)
    ap.add_argument("--nc", type=int, default=80)
    ap.add_argument("--conf", type=float, default=0.25, help="NMS confidence threshold (BASELINE configs[2]: 0.001)")
    ap.add_argument("--iou", type=float, default=0.7)
    ap.add_argument("--max-det", type=int, default=300)
    ap.add_argument("--multi-label", action="store_true", help="validator-style NMS: every (anchor, class) pair above --conf is a candidate")
    ap.add_argument("--cpu-sample", type=int, default=None,
                    help="images per CPU-baseline step (default: 16 for the n model at 640x640 -- the batch at which the host cores are busiest -- else 4)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-profile", action="store_true", help="skip the per-kernel roofline pass")
    ap.add_argument("--no-graph", action="store_true", help="eager launches instead of CUDA-graph replay (for ncu)")
    ap.add_argument("--profile-step", action="store_true",
                    help="run warm-ups, then ONE eager step between cudaProfilerStart/Stop and exit (for `ncu --profile-from-start off`)")
    ap.add_argument("--no-cudnn-benchmark", action="store_true", help="skip cuDNN autotuning (keeps ncu launch lists short)")
    ap.add_argument("--no-extras", action="store_true", help="skip the extra_configs legs (configs[2], [3], [4])")
    ap.add_argument("--extras", default="c2,c3,c4", help="which extra legs to run: c2 = s@1280 stress, c3 = m batch 512 strong scaling, c4 = DDP train step")
    ap.add_argument("--sustained-seconds", type=float, default=2.5, help="length of the sustained leg (0 = skip)")
    ap.add_argument("--no-ref-gpu", action="store_true", help="skip the unmodified reference's eager run on the GPU")
    ap.add_argument("--ref-gpu-batch", type=int, default=None, help="batch of the reference's eager GPU run (default: --batch)")
    a = ap.parse_args()
    if a.cpu_sample is None:
        a.cpu_sample = 16 if (a.scale, a.imgsz) == ("n", 640) else 4
    return a


def nms_settings(a):
    return dict(conf=a.conf, iou=a.iou, max_det=a.max_det, multi_label=a.multi_label)


def workload(a):
    nms = "predict defaults (conf .25, iou .7, max_det 300)" if (a.conf, a.iou, a.max_det, a.multi_label) == (0.25, 0.7, 300, False) else \
        f"NMS conf {a.conf:g}, iou {a.iou:g}, max_det {a.max_det}, {'multi_label (validator / stress regime)' if a.multi_label else 'single label'}"
    return {"workload": f"EdgeLine-YOLO-{a.scale} inference, synthetic {a.imgsz}x{a.imgsz}, batch {a.batch}/GPU, nc={a.nc}, "
                        f"{(a.imgsz // 8) ** 2 + (a.imgsz // 16) ** 2 + (a.imgsz // 32) ** 2} anchors, {nms}, "
                        "random-init weights (wave.gamma=0.5, no bias_init)",
            "batch_per_gpu": a.batch, "imgsz": a.imgsz, "scale": a.scale, "nc": a.nc, "parallelism": f"batch-sharded replicas x{a.gpus}",
            "input": "uint8 HWC batch (what the predictor's preprocess consumes); `value`: resident in HBM, `e2e`: pinned host memory",
            "l2_policy": "every step streams ~8 GB of activations through the 126 MB L2, so nothing of the 78.6 MB input batch or of one step's "
                         "intermediates survives to the next step; the per-kernel pass rotates input copies totalling more than L2"}


# ----------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,power.draw"

    def __init__(self, index: int, period: float = 0.1, autostart: bool = True):
        self.index, self.rows, self.stop, self.period, self.autostart = index, [], threading.Event(), period, autostart
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.started = False

    def start(self):
        """Begin sampling (idempotent).  The 20-step leg calls this right AFTER its steps are enqueued: the samples still fall inside the timed
        region (the GPU needs ~50 ms for them), but the process spawn of the first nvidia-smi query can no longer hold up the thread that is
        enqueueing -- a 2 ms host hiccup is 4 % of that region (seen once: 25.0k against 26.0k img/s with the sustained leg unchanged)."""
        if not self.started:
            self.started = True
            self.thread.start()

    def _run(self):
        while not self.stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self.stop.wait(self.period)

    def __enter__(self):
        if self.autostart:
            self.start()
        return self

    def __exit__(self, *exc):
        self.stop.set()
        if self.started:
            self.thread.join(timeout=6)

    def summary(self):
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        reasons = set()
        for r in self.rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        pw = []
        for r in self.rows:
            try:
                pw.append(float(r[6]))
            except (IndexError, ValueError):
                pass
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_min_mhz": sm[0] if sm else None, "sm_max_mhz": mx[0] if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "power_w_max": max(pw) if pw else None}


# ------------------------------------------------------------------------------ CPU baseline / reference arm
def cpu_oracle_rate(a, steps, warmup):
    """Images/s of the CPU oracle port (reference semantics on the host cores), all torch threads.  Fallback of the reference arm when
    baseline/_ref is absent."""
    import torch

    from oracle import model_ref

    torch.set_num_threads(os.cpu_count() or 1)
    model = model_ref.build(a.scale, a.nc, seed=0)
    x = torch.rand(a.cpu_sample, 3, a.imgsz, a.imgsz, generator=torch.Generator().manual_seed(0))
    for _ in range(warmup):
        model_ref.predict(model, x, **nms_settings(a))
    t0 = time.perf_counter()
    for _ in range(steps):
        model_ref.predict(model, x, **nms_settings(a))
    dt = time.perf_counter() - t0
    return a.cpu_sample * steps / dt, dt / steps * 1e3, torch.get_num_threads()


def reference_available():
    from oracle import ref_loader

    return ref_loader.available()


def reference_rate(a, device, half, sample, steps, warmup):
    """Images/s of the UNMODIFIED reference package through its own public API and stock code path:
    `YOLO("yolo11<scale>-test.yaml", task="detect").predict(tensor, device=..., conf, iou, max_det, half)` (engine/model.py:501 ->
    engine/predictor.py:220 -> models/yolo/detect/predict.py:23-41).  `YOLO.predict` has no multi_label switch; the validator regime
    (BASELINE configs[2]) therefore runs what DetectionValidator does per batch (models/yolo/detect/val.py:92-103): the reference's fused
    DetectionModel forward + `ultralytics.utils.ops.non_max_suppression(multi_label=True)`.  Same seeded random-init regime as the
    product arm (wave.gamma = 0.5, no bias_init).  Returns (images/s, ms per step, torch threads, Results.speed of the last step)."""
    import torch

    from oracle import ref_loader

    ref_loader.load()  # sets OMP_NUM_THREADS = all cores before ultralytics is imported (Q10)
    import ultralytics.utils.ops as uops
    from ultralytics import YOLO

    on_gpu = str(device) != "cpu"
    if not on_gpu:
        # the stock package caps CPU inference at min(8, cores - 1) threads (utils/__init__.py NUM_THREADS, re-applied by select_device,
        # utils/torch_utils.py:170-175); the baseline is meant to use every host core, so the cap (a module constant, not code) is lifted
        import ultralytics.utils.torch_utils as tu

        tu.NUM_THREADS = os.cpu_count() or 1
        torch.set_num_threads(tu.NUM_THREADS)
    torch.manual_seed(0)
    m = YOLO(ref_loader.model_yaml(a.scale, None if a.nc == 80 else a.nc), task="detect", verbose=False)
    with torch.no_grad():
        for name, p in m.model.named_parameters():
            if name.endswith("wave.gamma"):
                p.fill_(0.5)
    x = torch.rand(sample, 3, a.imgsz, a.imgsz, generator=torch.Generator().manual_seed(0))
    speed = None
    if a.multi_label:
        dev = torch.device("cuda", int(device)) if on_gpu else torch.device("cpu")
        net = m.model.fuse(verbose=False).to(dev).eval()
        net = net.half() if half else net.float()

        def step():
            with torch.inference_mode():
                y = net(x.to(dev, torch.half if half else torch.float))
                return uops.non_max_suppression(y, a.conf, a.iou, multi_label=True, max_det=a.max_det)
    else:
        def step():
            nonlocal speed
            res = m.predict(x, device=device, conf=a.conf, iou=a.iou, max_det=a.max_det, half=half, verbose=False)
            speed = res[0].speed
            return res

    def sync():
        if on_gpu:
            torch.cuda.synchronize()

    for _ in range(warmup):
        step()
    sync()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    sync()
    dt = time.perf_counter() - t0
    return sample * steps / dt, dt / steps * 1e3, torch.get_num_threads(), speed


def reference_arm(a):
    """`--impl reference`: rank 0 times the reference's own implementation of the path on the host cores (all threads), on a bounded
    sample of the product arm's workload; the other ranks exit without work."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = max(1, a.steps), max(1, min(a.warmup, 2))
    sample = f"{a.cpu_sample} images per step (bounded sample of the batch-{a.batch} workload), preprocess + forward + decode + NMS, fp32"
    extra = {}
    if reference_available():
        rate, ms, threads, speed = reference_rate(a, "cpu", False, a.cpu_sample, steps, warmup)
        kind = "reference"
        extra["api"] = ("YOLO(cfg, task='detect').predict(tensor, device='cpu')" if not a.multi_label
                        else "DetectionModel.fuse() forward + ultralytics.utils.ops.non_max_suppression(multi_label=True) (what DetectionValidator runs per batch)")
        if speed:
            extra["results_speed_ms_per_image"] = speed
    else:
        rate, ms, threads = cpu_oracle_rate(a, steps, warmup)
        kind = "port"
        extra["api"] = "oracle/model_ref.py (baseline/_ref not found: run __graft_entry__.build() in the dev container)"
    line = {"impl": "reference", "metric": metric(a), "value": rate, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload(a),
            "cpu_baseline": dict({"value": rate, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample}, **extra),
            "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    if kind == "reference" and not a.no_ref_gpu:
        try:
            import torch

            if torch.cuda.is_available():
                gb = a.ref_gpu_batch or a.batch
                g_rate, g_ms, _, g_speed = reference_rate(a, 0, True, gb, 5, 2)
                line["reference_eager_gpu"] = {"value": g_rate, "unit": UNIT, "ms_per_step": g_ms, "batch": gb, "dtype": "fp16 (half=True, the reference's only 16-bit mode, Q11)",
                                               "api": extra["api"].replace("device='cpu'", "device=0, half=True"), "results_speed_ms_per_image": g_speed,
                                               "note": "the unmodified reference, whole model, eager PyTorch / cuDNN / torchvision on the same B200: the honest 'before' "
                                                       "of the product (the reference ships no GPU kernel of its own)"}
        except Exception as exc:  # the CPU number stands on its own
            line["reference_eager_gpu"] = {"error": f"{type(exc).__name__}: {exc}"[:300]}
    _emit(line)


def reference_subprocess(a, steps, warmup, gpu: bool):
    """Runs the reference arm in a fresh interpreter (ultralytics sets process-wide state at import: OMP threads, cv2 threads, cuBLAS
    workspace config, print options) and returns its parsed JSON line, or None."""
    cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", str(steps), "--warmup", str(warmup), "--scale", a.scale,
           "--batch", str(a.batch), "--imgsz", str(a.imgsz), "--nc", str(a.nc), "--conf", str(a.conf), "--iou", str(a.iou), "--max-det", str(a.max_det),
           "--cpu-sample", str(a.cpu_sample)] + (["--multi-label"] if a.multi_label else []) + ([] if gpu else ["--no-ref-gpu"])
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "MASTER_ADDR", "MASTER_PORT", "OMP_NUM_THREADS")}
    try:
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
        lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
        return json.loads(lines[-1]) if lines else {"error": (r.stderr or "no output")[-300:]}
    except Exception as exc:
        return {"error": f"{type(exc).__name__}: {exc}"[:300]}


# --------------------------------------------------------------------------- kernel roofline
def profile_kernels(pred, a, iters=10):
    """Times every hot-path kernel call of one forward in isolation (CUDA events on the launching stream, L2
    flushed before each timed launch, GPU kept busy so the events bracket only the kernel) and attaches the
    algorithmic bytes of SURVEY.md section 8(d)."""
    import torch

    from edge_yolo_b200 import ops

    calls = []
    nms_meta = {}
    e = lambda t: t.element_size()

    def record(name, fn, nbytes):
        def wrapped(*args, **kw):
            out = fn(*args, **kw)
            calls.append((name, fn, args, kw, nbytes(out, *args, **kw)))
            return out
        return wrapped

    shims = {
        "dwt_haar": lambda out, x: 2 * x.numel() * e(x),
        "wave_merge": lambda out, b, LLp, *r: int(4.5 * b.numel() * e(b)),
        "wave_merge_bands": lambda out, LLp, *r: int(2.5 * out.numel() // 2 * e(out)),
        "gated_residual": lambda out, b, y, g, **k: 3 * b.numel() * e(b),
        "linear_attention": lambda out, qkv, heads: (qkv.numel() + out.numel()) * e(qkv),
        "gfl_decode": lambda out, boxes, clss, *r, **k: sum(t.numel() * e(t) for t in boxes + clss) + out.numel() * 4,
        "bias_act": lambda _o, x, bias, act=1, residual=None, **k: (2 + (residual is not None)) * x.numel() * e(x),
        "pwconv": lambda _o, srcs, wpk, N, bias=None, act=0, residual=None, up_addend=None, **k: (
            sum(t.numel() for t in srcs) + srcs[0].numel() // srcs[0].shape[1] * N * (1 + (residual is not None))
            + (up_addend.numel() if up_addend is not None else 0)) * e(srcs[0]),
        "dwconv": lambda _o, x, *r, **k: 2 * x.numel() * e(x),
        "conv3x3": lambda _o, x, wpk, N, **k: (x.numel() + _o.numel()) * e(x),
        "conv3x3_halo": lambda _o, x, wpk, N, **k: (x.numel() + _o.numel()) * e(x),
        "conv3x3_mma": lambda _o, x, w, **k: (x.numel() + _o.numel()) * e(x),
        "dsconv3": lambda _o, x, *r, **k: (x.numel() + _o.numel()) * e(x),  # fused depthwise -> pointwise: the depthwise tensor does not exist
        "upsample2x_cat": lambda _o, x, skip: (x.numel() + skip.numel() + _o.numel()) * e(x),
        "sppf_pool": lambda _o, x: 5 * x.numel() * e(x),
        "nms_batched": lambda out, y, *r, **k: y.shape[0] * y.shape[2] * (y.shape[1] - 4) * 4 + out[0].numel() * 4,
        # fused decode + NMS: head maps in, xywh boxes (16 B / anchor) out, rows out
        "gfl_detect": lambda out, boxes, clss, *r, **k: sum(t.numel() * e(t) for t in boxes + clss)
        + sum(t.shape[0] * t.shape[2] * t.shape[3] for t in boxes) * 16 + out[0].numel() * 4,
    }
    saved = {k: getattr(ops, k) for k in shims}
    try:
        for k, nb in shims.items():
            setattr(ops, k, record(k, saved[k], nb))
        with torch.no_grad():
            pred._forward(False)
    finally:
        for k, v in saved.items():
            setattr(ops, k, v)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=pred.device)

    def clone_arg(v):  # same sizes AND strides (channel-slice views keep their pitch), fresh storage
        if isinstance(v, torch.Tensor):
            if not v.is_cuda or v.numel() == 0:
                return v
            c = torch.empty_strided(v.size(), v.stride(), dtype=v.dtype, device=v.device)
            c.copy_(v)
            return c
        if isinstance(v, (list, tuple)):
            return type(v)(clone_arg(t) for t in v)
        return v

    from edge_yolo_b200 import _lib

    # the fused detect call is a chain of three kernels groups; time them separately (emit is HBM-bound, sort / sweep latency-bound)
    staged = []
    for name, fn, args, kw, nbytes in calls:
        if name != "gfl_detect":
            staged.append((name, fn, args, kw, nbytes, 7))
            continue
        boxes, clss = args[0], args[1]
        nA = sum(t.shape[0] * t.shape[2] * t.shape[3] for t in boxes)
        kw = dict(kw, workspace=torch.empty(1 << 30, dtype=torch.uint8, device=pred.device))  # persistent across the staged calls
        out_full, cnt_full = fn(*args, **dict(kw, stages=7))
        torch.cuda.synchronize()
        Bn = boxes[0].shape[0]
        # candidates per image: the emit kernel's counters sit at the start of the workspace (NmsLayout.counts); the sort / sweep see at most max_nms
        n_cand = kw["workspace"][: 4 * Bn].view(torch.int32).clamp(max=kw.get("max_nms", 30000)).long()
        n_tot, k_tot = int(n_cand.sum()), int(cnt_full.sum())
        nms_meta.update(candidates_per_image=n_tot / Bn, kept_per_image=k_tot / Bn,
                        pair_tests_upper_bound=int((n_cand * cnt_full.long()).sum()))
        staged.append(("gfl_decode_emit", fn, args, kw, sum(t.numel() * e(t) for t in boxes + clss) + nA * 16 + 8 * n_tot, 1))
        # SURVEY 8(d) NMS bytes for THIS design (no n x n/64 bit mask exists): sort = read + write of the 8-byte keys; sweep = 8-byte key
        # + 16-byte gathered box per candidate + 24-byte row per kept detection
        staged.append(("nms_sort", fn, args, kw, 16 * n_tot, 2))
        staged.append(("nms_sweep", fn, args, kw, 24 * n_tot + 24 * k_tot + 4 * Bn, 4))

    per = {}
    with torch.no_grad():
        for name, fn, args, kw, nbytes, stages in staged:
            if stages != 7:
                fn(*args, **dict(kw, stages=7))  # leaves emitted + sorted keys in place for the partial runs
                kw = dict(kw, stages=stages)
            # R rotating copies of the inputs (>= 512 MB in total) so no launch finds its operands in L2; all R launches are
            # enqueued behind a ~1 ms spin so that they run back to back, bracketed by ONE event pair on the launching stream
            R = int(min(48, max(2, (512 << 20) // max(nbytes, 1)))) if stages in (7, 1) else 4
            sets = [(args, kw)] + ([(tuple(clone_arg(v) for v in args), kw) for _ in range(R - 1)] if stages in (7, 1) else [(args, kw)] * (R - 1))
            ts = []
            for it in range(iters // 2 + 1):
                flush.zero_()
                torch.cuda._sleep(6_000_000)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for aa, kk in sets:
                    fn(*aa, **kk)
                e1.record()
                e1.synchronize()
                if it >= 1:
                    ts.append(e0.elapsed_time(e1) * 1e-3 / R)
            del sets
            ts.sort()
            t = ts[len(ts) // 2]
            d = per.setdefault(name, {"launch_sites": 0, "bytes": 0, "seconds": 0.0, "sites": []})
            d["launch_sites"] += 1
            d["bytes"] += nbytes
            d["seconds"] += t
            first = args[0][0] if isinstance(args[0], (list, tuple)) else args[0]
            site = {"shape": list(first.shape), "MB": round(nbytes / 1e6, 2), "us": round(t * 1e6, 2), "gbs": round(nbytes / t / 1e9, 1)}
            same = [x for x in d["sites"] if x["shape"] == site["shape"] and x["MB"] == site["MB"]]
            if same:
                same[0]["count"] = same[0].get("count", 1) + 1
                same[0]["us"] = round((same[0]["us"] * (same[0]["count"] - 1) + site["us"]) / same[0]["count"], 2)
            else:
                d["sites"].append(site)
        if pred.stem is not None:  # the fused uint8 stem runs only on the from-uint8 path: timed here with rotating inputs
            R = 4
            srcs = [pred.u8.clone() for _ in range(R)]
            out0 = ops.stem_conv_u8(srcs[0], *pred.stem, dtype=pred.dtype)
            nbytes = srcs[0].numel() + out0.numel() * out0.element_size()
            ts = []
            for it in range(iters // 2 + 1):
                flush.zero_()
                torch.cuda._sleep(6_000_000)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for sx in srcs:
                    ops.stem_conv_u8(sx, *pred.stem, dtype=pred.dtype, out=out0)
                e1.record()
                e1.synchronize()
                if it >= 1:
                    ts.append(e0.elapsed_time(e1) * 1e-3 / R)
            ts.sort()
            t = ts[len(ts) // 2]
            per["stem_conv_u8"] = {"launch_sites": 1, "bytes": nbytes, "seconds": t,
                                   "sites": [{"shape": list(out0.shape), "MB": round(nbytes / 1e6, 2), "us": round(t * 1e6, 2), "gbs": round(nbytes / t / 1e9, 1)}]}
    for d in per.values():
        d["gbs"] = d["bytes"] / d["seconds"] / 1e9
        d["us"] = d["seconds"] * 1e6
        del d["seconds"]
    if "nms_sweep" in per and nms_meta:
        per["nms_sweep"].update(nms_meta, pair_tests_per_s_upper_bound=nms_meta["pair_tests_upper_bound"] / (per["nms_sweep"]["us"] * 1e-6),
                                bound="latency: greedy suppression is a serial chain per image (each keep decision depends on all earlier keeps); "
                                      "bytes = keys + gathered boxes + kept rows, so the HBM fraction is reported but is not what limits it")
    return per


def step_traffic(kernel):
    """(DRAM bytes read + written by all launches of `kernel` in one step, capture file) from the newest committed ncu capture of the
    default workload (profiles/r*_step_b64_time_dram.json: `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum`
    over `bench.py --profile-step`).  ncu cannot run inside the timed program, so this is labelled with its source file."""
    names = {"pwconv": ["pw::pwconv_tc_kernel"], "dwconv": ["el::dwconv_tile_kernel", "dwtc::dwconv_tc_kernel", "dwt::dwconv3_tma_kernel"],
             "bias_act": ["el::bias_act_tiled", "el::bias_act_flat"], "conv3x3_halo": ["c3::conv3x3_halo_kernel"], "conv3x3_mma": ["c3m::conv3x3_mma_kernel"], "dsconv3": ["ds::dsconv3_tc_kernel"], "linear_attention": ["ta::linattn_tma_kernel"],
             "stem_conv_u8": ["stemtc::stem_tc_kernel"], "wave_merge_bands": ["el::merge_fwd_x2"], "dwt_haar": ["el::dwt_fwd_tiled"],
             "gfl_decode_emit": ["el::gfl_decode_emit_kernel"], "nms_sweep": ["el::nms_sweep"], "upsample2x_cat": ["el::upsample2x_cat_tiled"]}
    import glob

    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_step_b64_time_dram.json")), reverse=True):
        try:
            with open(path) as f:
                prof = json.load(f)
            # pwconv_tc_kernel also runs the narrow 3x3 convs (el_conv3x3_fwd); the capture cannot tell the two apart
            total = sum(v["dram_bytes"] for k, v in prof.items() if any(k.startswith(n) for n in names.get(kernel, [])))
            if total:
                return total, "profiles/" + os.path.basename(path)
        except Exception:
            continue
    return None, None


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (burst copy)"
    except Exception:
        return 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"


# ------------------------------------------------------------------------------ product arm
def _free_cuda():
    import gc

    import torch

    gc.collect()
    torch.cuda.empty_cache()


def inference_leg(a, ctx, steps, warmup, full: bool):
    """One inference configuration on this rank's GPU: device-resident leg (`value`), end-to-end leg through the public API (pinned
    host uint8 in, pinned host detections out), and -- with `full` -- the sustained leg, the clocks and the Predictor itself (kept
    alive for the per-kernel pass).  Times are CUDA events on the launching stream, max over ranks."""
    import torch

    from edge_yolo_b200 import dist as eld
    from edge_yolo_b200.engine import Predictor, build_model

    dev, rank, world, local = ctx["dev"], ctx["rank"], ctx["world"], ctx["local"]
    model = build_model(a.scale, a.nc, seed=0, device=dev)
    pred = Predictor(model, a.batch, a.imgsz, use_graph=not a.no_graph, **nms_settings(a))
    gen = torch.Generator().manual_seed(1234 + rank)
    host_u8 = torch.randint(0, 256, (a.batch, a.imgsz, a.imgsz, 3), dtype=torch.uint8, generator=gen).pin_memory()
    pred.predict_u8(host_u8)  # also leaves a real batch in pred.x
    pred.load_resident(host_u8)  # `value`: the batch is resident in every input buffer of the pipelined graphs

    if a.profile_step and full:
        for _ in range(max(warmup, 2)):
            pred.step_device()
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        pred.step_device()
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        return None

    barrier = lambda: eld.barrier(dev)
    max_over_ranks = lambda v: eld.max_over_ranks(v, dev)

    # ---- leg 1: inputs resident in HBM, device-timed
    for _ in range(warmup):
        pred.step_device()
    pred.drain()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    res = {}
    with ClockSampler(local, autostart=False) as clk:
        e0.record()
        for _ in range(steps):
            pred.step_device()
        clk.start()   # everything is enqueued; the samples are taken while the GPU works through it
        pred.drain()  # the NMS graph of the last steps runs on the engine's side stream: the timed region ends when it has finished
        e1.record()
        e1.synchronize()
        barrier()
        ms_total = max_over_ranks(e0.elapsed_time(e1))
        # ---- leg 2: end to end through the public API: pinned host uint8 in, pinned host detections out
        for _ in range(min(warmup, 3)):
            pred.predict_u8(host_u8)
        pred.predict_many([host_u8] * 3)  # allocates the pipeline's pinned / staging buffers outside the timed region
        barrier()
        kept_per_step = []
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        g0.record()
        # public API, one call per step; the H2D of step i+1 overlaps the compute of step i (double-buffered staging)
        pred.predict_many([host_u8] * steps, consume=lambda i, rows, cnt: kept_per_step.append(int(cnt.sum())))
        g1.record()  # predict_many returns after the last batch's rows have reached pinned host memory: the device is idle here
        g1.synchronize()
        e2e_wall_s = time.perf_counter() - t0
        barrier()
        # device-timed (CUDA events around the whole call, max over ranks); the host's wall clock over the same region rides along
        e2e_s = max_over_ranks(g0.elapsed_time(g1) * 1e-3)
        e2e_wall_s = max_over_ranks(e2e_wall_s)
        t1 = time.perf_counter()
        for _ in range(steps):  # the same without pipelining: copy in, compute, copy out, one step at a time
            pred.predict_u8(host_u8)
        barrier()
        e2e_serial_s = max_over_ranks(time.perf_counter() - t1)
        if full:  # H2D alone: what one rank's pinned-host -> device copy path delivers while all ranks copy at once (the e2e limiter at N = 8)
            stage = torch.empty(host_u8.shape, dtype=torch.uint8, device=dev)
            barrier()
            h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            h0.record()
            for _ in range(10):
                stage.copy_(host_u8, non_blocking=True)
            h1.record()
            h1.synchronize()
            barrier()
            res["h2d_gbs_per_rank"] = 10 * host_u8.numel() / (max_over_ranks(h0.elapsed_time(h1)) * 1e-3) / 1e9
            del stage
    res["clocks"] = clk.summary()
    res.update(value=world * a.batch * steps / (ms_total * 1e-3), ms_per_step=ms_total / steps,
               e2e={"value": world * a.batch * steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": host_u8.numel(),
                    "d2h_bytes_per_step": pred.host_out.numel() * 4 + pred.host_cnt.numel() * 4, "ms_per_step": e2e_s / steps * 1e3,
                    "wall_ms_per_step": e2e_wall_s / steps * 1e3, "detections_per_step": kept_per_step[-1],
                    "unpipelined_value": world * a.batch * steps / e2e_serial_s},
               gpu_launches_per_step=pred.launches_per_step)
    if "h2d_gbs_per_rank" in res:
        res["e2e"]["h2d_gbs_per_rank"] = res.pop("h2d_gbs_per_rank")
    if full and a.sustained_seconds > 0:
        # ---- leg 3: sustained -- back-to-back steps for >= sustained_seconds (the 20-step legs above last ~60 ms: a burst at max clock)
        n = max(steps, int(1.25 * a.sustained_seconds / (res["ms_per_step"] * 1e-3)) + 1)  # margin: back-to-back steps run a little faster than the 20-step leg
        barrier()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with ClockSampler(local, period=0.05) as sclk:
            s0.record()
            for _ in range(n):
                pred.step_device()
            pred.drain()
            s1.record()
            s1.synchronize()
        barrier()
        sus_ms = max_over_ranks(s0.elapsed_time(s1))
        res["sustained"] = {"value": world * a.batch * n / (sus_ms * 1e-3), "unit": UNIT, "steps": n, "seconds": sus_ms * 1e-3,
                            "ms_per_step": sus_ms / n, "clocks": sclk.summary()}
    if full:
        res["pred"] = pred
    return res


def extra_configs(a, ctx, which):
    """BASELINE.json configs[2], [3], [4] measured in the same run, on all ranks (each leg: barrier, CUDA events, max over ranks)."""
    import copy

    import torch

    from edge_yolo_b200 import dist as eld

    out = {}
    world, rank, dev = ctx["world"], ctx["rank"], ctx["dev"]
    if "c2" in which:
        # configs[2]: EdgeLine-YOLO-s at 1280x1280, 10 classes, batch 32, conf 0.001 + multi_label (validator / NMS stress regime)
        b = copy.copy(a)
        b.scale, b.imgsz, b.batch, b.nc, b.conf, b.multi_label = "s", 1280, 32, 10, 0.001, True
        try:
            r = inference_leg(b, ctx, steps=5, warmup=3, full=False)
            out["configs[2]"] = {"metric": metric(b), "workload": workload(b)["workload"], "scaling": "weak", "value": r["value"], "unit": UNIT,
                                 "ms_per_step": r["ms_per_step"], "e2e": r["e2e"], "steps": 5, "gpu_launches_per_step": r["gpu_launches_per_step"]}
        except Exception as exc:
            out["configs[2]"] = {"error": f"{type(exc).__name__}: {exc}"[:300]}
        _free_cuda()
    if "c3" in which:
        # configs[3]: EdgeLine-YOLO-m, batch 512 in total, sharded 512 / N per GPU: STRONG scaling (total work fixed)
        b = copy.copy(a)
        lo, hi = eld.shard_bounds(512, world, rank)
        b.scale, b.imgsz, b.batch, b.nc, b.conf, b.multi_label = "m", 640, hi - lo, 80, 0.25, False
        try:
            r = inference_leg(b, ctx, steps=5, warmup=3, full=False)
            # inference_leg multiplies by world * its own batch; with an even split that is the 512-image total
            out["configs[3]"] = {"metric": metric(b), "workload": f"EdgeLine-YOLO-m inference, synthetic 640x640, batch 512 sharded {hi - lo}/GPU over {world} GPU(s), predict defaults",
                                 "scaling": "strong", "global_batch": 512, "batch_per_gpu": hi - lo, "value": r["value"], "unit": UNIT,
                                 "ms_per_step": r["ms_per_step"], "e2e": r["e2e"], "steps": 5, "gpu_launches_per_step": r["gpu_launches_per_step"]}
        except Exception as exc:
            out["configs[3]"] = {"error": f"{type(exc).__name__}: {exc}"[:300]}
        _free_cuda()
    if "c4" in which:
        try:
            out["configs[4]"] = train_leg(ctx, scale="s", batch=64, imgsz=640, nc=80, steps=5, warmup=3)
        except Exception as exc:
            out["configs[4]"] = {"error": f"{type(exc).__name__}: {exc}"[:300]}
        _free_cuda()
    return out


def train_leg(ctx, scale, batch, imgsz, nc, steps, warmup, amp=True, channels_last=True):
    """configs[4]: DDP training step (EdgeLine-YOLO-s, 64 images / GPU, 640x640, DFL + cls + box loss forward / backward, SGD step,
    NCCL all-reduce of the gradients over NVLink when N > 1)."""
    import torch

    from edge_yolo_b200 import dist as eld
    from edge_yolo_b200.train import TrainStep

    dev, rank, world, local = ctx["dev"], ctx["rank"], ctx["world"], ctx["local"]
    ts = TrainStep(scale, nc, dev, world, local, amp=amp, channels_last=channels_last)
    x, targets = ts.synth_batch(batch, imgsz, rank)
    ev = lambda: torch.cuda.Event(enable_timing=True)

    def timed(n, sync_grads=True):
        eld.barrier(dev)
        s0, s1 = ev(), ev()
        s0.record()
        for _ in range(n):
            items = ts.step(x, targets, sync_grads=sync_grads)
        s1.record()
        s1.synchronize()
        eld.barrier(dev)
        return eld.max_over_ranks(s0.elapsed_time(s1), dev) / n, items

    for _ in range(warmup):
        ts.step(x, targets)
    ms, items = timed(steps)
    phases = [0.0, 0.0, 0.0]
    for _ in range(3):  # phase split (forward / loss / backward + all-reduce + optimiser) with events inside the step
        es = [ev() for _ in range(4)]
        ts.step(x, targets, events=es)
        es[3].synchronize()
        for i in range(3):
            phases[i] += es[i].elapsed_time(es[i + 1]) / 3
    res = {"metric": f"training images/sec (EdgeLine-YOLO-{scale}, {imgsz}x{imgsz}, batch {batch}/GPU, {'bf16 autocast' if amp else 'fp32'}"
                     f"{', NHWC' if channels_last else ''})", "scaling": "weak", "value": world * batch / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms,
           "steps": steps, "forward_ms": phases[0], "loss_ms": phases[1], "backward_allreduce_step_ms": phases[2],
           "loss_items": [float(v) for v in items.tolist()], "peak_mem_GB": torch.cuda.max_memory_allocated(dev) / 1e9,
           "allreduce_payload_bytes": 4 * ts.n_params,
           "workload": "forward (CUDA DWT / merge / gated residual / attention kernels under autograd) + v8DetectionLoss (el_tal_assign, el_dfl, BCE) "
                       "+ backward (CUDA backward kernels) + DDP bucketed NCCL all-reduce + SGD step; 8 boxes / image, synthetic"}
    if world > 1:
        ms_nosync, _ = timed(steps, sync_grads=False)  # same step without the collective (DDP.no_sync): the difference is its EXPOSED cost
        res["ms_per_step_no_allreduce"] = ms_nosync
        res["allreduce_exposed_ms"] = ms - ms_nosync
        try:  # NCCL kernel time itself, from a CUPTI trace of two extra steps (outside every timed region)
            from torch.profiler import ProfilerActivity, profile

            with profile(activities=[ProfilerActivity.CUDA]) as prof:
                for _ in range(2):
                    ts.step(x, targets)
                torch.cuda.synchronize(dev)
            nccl_us = sum(e.device_time_total for e in prof.key_averages() if "nccl" in e.key.lower())
            res["allreduce_nccl_kernel_ms"] = nccl_us / 2 / 1e3
            res["allreduce_share_of_step"] = res["allreduce_nccl_kernel_ms"] / ms
        except Exception as exc:
            res["allreduce_nccl_kernel_ms"] = None
            res["allreduce_note"] = f"CUPTI trace unavailable: {type(exc).__name__}"
    return res


def product_arm(a):
    import torch

    from edge_yolo_b200 import _lib
    from edge_yolo_b200 import dist as eld

    _lib.lib()  # fail loudly if the extension is missing
    rank, world, local = eld.env_rank()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (product arm) needs a GPU; there is no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    pin_to_local_numa(local)
    eld.init(dev)
    torch.backends.cudnn.benchmark = not a.no_cudnn_benchmark
    ctx = {"dev": dev, "rank": rank, "world": world, "local": local}

    r = inference_leg(a, ctx, a.steps, a.warmup, full=True)
    if r is None:  # --profile-step
        eld.shutdown()
        return
    pred = r.pop("pred")
    line = {"metric": metric(a), "value": r["value"], "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic", "config": workload(a), "clocks": r["clocks"], "e2e": r["e2e"],
            "gpu_launches": (r["gpu_launches_per_step"] or 0) * a.steps, "gpu_launches_per_step": r["gpu_launches_per_step"]}
    line["e2e"]["note"] = ("CUDA events around the whole call, max over ranks (wall_ms_per_step = the host clock over the same region). Predictor.predict_many: "
                           "pinned uint8 batch H2D every step (copy stream, overlapped with the previous step's compute), forward graph + NMS graph (side stream, "
                           "overlapped with the next step's forward), rows+counts D2H every step; h2d_gbs_per_rank = this rank's pinned-host -> device copy "
                           "rate while all ranks copy at once")
    if "sustained" in r:
        line["sustained"] = r["sustained"]

    if rank == 0 and not a.no_profile:
        per = profile_kernels(pred, a)
        peak, peak_src = measured_peak()
        top = max(per, key=lambda k: per[k]["us"])
        line["kernels"] = per
        best = max(per[top]["sites"], key=lambda s_: s_["MB"])
        traffic, traffic_src = step_traffic(top) if (a.scale, a.batch, a.imgsz, a.nc) == ("n", 64, 640, 80) else (None, None)
        line["roofline"] = {"kernel": top, "bound": "hbm", "achieved": per[top]["gbs"], "peak": peak, "unit": "GB/s",
                            "frac": per[top]["gbs"] / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                            "launch_sites": per[top]["launch_sites"], "algorithmic_bytes": per[top]["bytes"],
                            "largest_site": {"shape": best["shape"], "MB": best["MB"], "us": best["us"], "achieved": best["gbs"], "frac": best["gbs"] / peak},
                            "note": "achieved = algorithmic bytes of ALL launch sites of this kernel in one step / their summed CUDA-event time "
                                    "(each site timed alone, inputs rotated through > L2); traffic = DRAM bytes of the same launches in one step "
                                    "from the committed ncu capture named in traffic_source (ncu cannot run inside the timed program), null for other configs"}
    del pred
    _free_cuda()
    if not a.no_extras:
        line["extra_configs"] = extra_configs(a, ctx, set(a.extras.split(",")))
    eld.shutdown()
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        cpu_steps = 16 if (a.scale, a.imgsz) == ("n", 640) else 4  # ~5-10 s of CPU work on the box's host cores either way
        ref = reference_subprocess(a, steps=cpu_steps, warmup=1, gpu=not a.no_ref_gpu)
        if ref and "cpu_baseline" in ref:
            line["cpu_baseline"] = dict(ref["cpu_baseline"], sample=f"{cpu_steps} steps x " + ref["cpu_baseline"]["sample"])
            if "reference_eager_gpu" in ref:
                line["reference_eager_gpu"] = ref["reference_eager_gpu"]
        else:
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "unavailable", "sample": str(ref)[:300]}
    if rank == 0:
        _emit(line)


def pin_to_local_numa(local_rank: int):
    """Bind this process (and therefore the pinned staging buffers it allocates next) to the CPUs of the GPU's NUMA node, so that at
    N = 8 the ranks' H2D copies do not all cross one memory controller.  Uses NVML's CPU affinity for the device; silently a no-op
    where NVML or sched_setaffinity is unavailable (single-node boxes)."""
    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = {64 * i + b for i, w in enumerate(words) for b in range(64) if (w >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
    except Exception:
        pass


def _emit(line: dict):
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


if __name__ == "__main__":
    args = parse()
    # stdout must carry exactly ONE JSON line: libraries (NCCL prints its version banner to stdout) are sent to stderr
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        reference_arm(args)
    else:
        product_arm(args)
