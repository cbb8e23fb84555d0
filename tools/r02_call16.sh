#!/usr/bin/env bash
# Round 2, call 16: ncu of the rewritten sweep (bench regime) and of the decode-emit kernel.
set -u
mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"nms_sweep<" -c 1 -o gpurun_out/c16_sweep python tools/prof_detect.py --iters 1 > gpurun_out/c16_ncu.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"gfl_decode_emit" -c 1 -o gpurun_out/c16_emit python tools/prof_detect.py --iters 1 >> gpurun_out/c16_ncu.log 2>&1
true
