"""Host -> device rate of one 78.6 MB uint8 batch per rank, all ranks copying at once, for three kinds of pinned host memory:
torch pin_memory (cudaHostAlloc), write-combined cudaHostAlloc, and a 2 MB-huge-page mmap registered with cudaHostRegister.
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/prof_h2d_modes.py"""
import ctypes
import json
import mmap
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
N = 64 * 640 * 640 * 3
dst = [torch.empty(N, dtype=torch.uint8, device=dev) for _ in range(2)]
cudart = ctypes.CDLL("libcudart.so.12")
keep = []


def buf_pinned():
    return torch.empty(N, dtype=torch.uint8).pin_memory()


def buf_wc():
    p = ctypes.c_void_p()
    rc = cudart.cudaHostAlloc(ctypes.byref(p), ctypes.c_size_t(N), ctypes.c_uint(0x04))  # cudaHostAllocWriteCombined
    assert rc == 0, rc
    arr = (ctypes.c_uint8 * N).from_address(p.value)
    keep.append(arr)
    return torch.frombuffer(arr, dtype=torch.uint8)


def buf_huge():
    size = (N + (2 << 20) - 1) & ~((2 << 20) - 1)
    m = mmap.mmap(-1, size + (2 << 20), flags=mmap.MAP_PRIVATE | mmap.MAP_ANONYMOUS)
    addr = ctypes.addressof(ctypes.c_char.from_buffer(m))
    off = (-addr) % (2 << 20)
    libc = ctypes.CDLL(None, use_errno=True)
    rc = libc.madvise(ctypes.c_void_p(addr + off), ctypes.c_size_t(size), 14)  # MADV_HUGEPAGE
    t = torch.frombuffer(m, dtype=torch.uint8, offset=off, count=N)
    t.fill_(1)  # touch: the pages are allocated (huge where the kernel grants them) before they are registered
    rc2 = cudart.cudaHostRegister(ctypes.c_void_p(addr + off), ctypes.c_size_t(size), ctypes.c_uint(0))
    keep.append(m)
    thp = [l for l in open("/proc/self/smaps_rollup") if "AnonHugePages" in l]
    return t, (rc, rc2, thp[0].split()[1] if thp else "?")


def rate(src, iters=10):
    st = torch.cuda.Stream(device=dev)
    with torch.cuda.stream(st):
        for _ in range(2):
            dst[0].copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier(device_ids=[local])
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(st):
        e0.record(st)
        for i in range(iters):
            dst[i & 1].copy_(src, non_blocking=True)
        e1.record(st)
    e1.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) * 1e-3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return iters * N / float(t.item()) / 1e9


out = {"world": world}
a = buf_pinned()
out["pin_memory_GBs_per_rank"] = rate(a)
try:
    b = buf_wc()
    out["write_combined_GBs_per_rank"] = rate(b)
except Exception as e:  # noqa: BLE001
    out["write_combined_error"] = repr(e)
try:
    c, info = buf_huge()
    out["hugepage_registered_GBs_per_rank"] = rate(c)
    out["hugepage_info_madvise_register_AnonHugePagesKB"] = info
except Exception as e:  # noqa: BLE001
    out["hugepage_error"] = repr(e)
out["pin_memory_again_GBs_per_rank"] = rate(a)
if rank == 0:
    print(json.dumps(out))
if world > 1:
    dist.barrier(device_ids=[local])
    dist.destroy_process_group()
