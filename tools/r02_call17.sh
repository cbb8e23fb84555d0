#!/usr/bin/env bash
# Round 2, call 17: ncu of the rewritten sweep in the bench regime.
set -u
mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"nms_sweep$" -c 1 -o gpurun_out/c17_sweep python tools/prof_detect.py --iters 1 > gpurun_out/c17_ncu.log 2>&1
true
