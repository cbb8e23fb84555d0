"""Host-side logic of the N > 1 path on CPU: world_size-2 gloo process groups (no GPU needed)."""
import json
import os
import subprocess
import sys

import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_bounds_cover_and_balance():
    from edge_yolo_b200.dist import shard_bounds

    for total in (0, 1, 7, 64, 512, 513):
        for world in (1, 2, 3, 4, 8):
            parts = [shard_bounds(total, world, r) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
            sizes = [hi - lo for lo, hi in parts]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    from edge_yolo_b200 import dist as eld

    r, w = eld.init(None)
    eld.barrier(None)
    m = eld.max_over_ranks(10.0 + rank, None)  # every rank must see the slowest rank's time
    lo, hi = eld.shard_bounds(9, w, r)
    q.put((r, w, m, lo, hi))
    eld.shutdown()


def test_gloo_world2_barrier_and_max_reduce():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert out == [(0, 2, 11.0, 0, 5), (1, 2, 11.0, 5, 9)]


def test_reference_arm_under_torchrun_prints_one_line():
    """`bench.py --impl reference` with 2 ranks: rank 0 runs the reference arm (the unmodified reference when it is present, else the CPU
    oracle port) and prints ONE JSON line, rank 1 exits 0 silently."""
    env = dict(os.environ, OMP_NUM_THREADS="2")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(29700 + os.getpid() % 200), os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
           "--steps", "1", "--warmup", "1", "--cpu-sample", "1", "--imgsz", "64"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 2 and d["cpu_baseline"]["kind"] in ("reference", "port") and d["value"] > 0
    assert d["e2e"]["h2d_bytes_per_step"] == 0
