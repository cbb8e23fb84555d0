#!/usr/bin/env bash
# Round 2, call 56: ncu --set full of the fused depthwise -> pointwise kernel (tensor-core depthwise producer) at 64 -> 80 @ 80 x 80, batch 64.
set -u
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:dsconv3_tc_kernel -c 1 -o gpurun_out/c56_dsconv3 -f python tools/prof_dsconv.py 64 80 > gpurun_out/c56_ncu.log 2>&1
true
