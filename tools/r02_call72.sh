#!/usr/bin/env bash
# Round 2, call 72: ncu --set full of the final stem, merge, DWT, Toeplitz depthwise and NMS sweep kernels inside one eager step.
set -u
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 1 --no-extras --no-cpu-baseline --no-ref-gpu --sustained-seconds 0 --no-profile --profile-step --no-graph --no-cudnn-benchmark"
timeout 200 ncu --set full --clock-control none -k regex:"stem_tc_kernel|merge_fwd_x2p|dwt_fwd_tiled|dwconv_tc_kernel|nms_sweep$|sort_single|conv3x3_mma_kernel|dwconv3_tma_kernel|sppf_pool_kernel" -c 40 -f -o gpurun_out/c72_misc $B > gpurun_out/c72_ncu.log 2>&1
true
