#!/usr/bin/env bash
# Round 2, call 47: ncu --set full (source counters) of the 3x3 halo kernel at 64 -> 64 @ 80 x 80, batch 64; full suite after the test fix.
set -u
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv3x3_halo_kernel -s 1 -c 1 -o gpurun_out/c47_halo -f python tools/prof_halo_one.py > gpurun_out/c47_ncu.log 2>&1
timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/c47_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/c47_pytest.log
true
