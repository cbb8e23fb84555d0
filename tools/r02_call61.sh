#!/usr/bin/env bash
# Round 2, call 61: host NUMA facts of the GPU box; dsconv3 with the conflict-free tap layout (parity + per-site times).
set -u
mkdir -p gpurun_out
{ echo "nodes online: $(cat /sys/devices/system/node/online 2>&1)"; grep -E "Cpus_allowed_list|Mems_allowed_list" /proc/self/status; nproc; lscpu | grep -E "Model name|Socket|NUMA|Thread|Core" ; free -g | head -2; for n in /sys/devices/system/node/node*; do echo "$n: $(cat $n/cpulist 2>/dev/null) $(grep MemTotal $n/meminfo 2>/dev/null)"; done; which numactl; python - <<'P'
import ctypes, os
libc = ctypes.CDLL(None, use_errno=True)
# get_mempolicy(mode, nodemask, maxnode, addr, flags)
mode = ctypes.c_int(); mask = (ctypes.c_ulong * 16)()
r = libc.syscall(239, ctypes.byref(mode), mask, 1024, 0, 0)
print("get_mempolicy rc", r, "errno", ctypes.get_errno(), "mode", mode.value, "mask", hex(mask[0]))
# try set_mempolicy(MPOL_INTERLEAVE=3, mask nodes 0-1)
m2 = (ctypes.c_ulong * 16)(); m2[0] = 0x3
r = libc.syscall(238, 3, m2, 1024)
print("set_mempolicy interleave {0,1} rc", r, "errno", ctypes.get_errno())
P
} > gpurun_out/c61_numa.txt 2>&1
timeout 600 python -m pytest tests -m gpu -q -x -k "dsconv3" > gpurun_out/c61_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/c61_pytest.log
timeout 300 python tools/prof_dsconv.py 64,80,128,256 > gpurun_out/c61_prof_dsconv.jsonl 2> gpurun_out/c61_prof_dsconv.err
true
