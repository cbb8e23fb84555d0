#!/usr/bin/env bash
# Round 2, call 14: pipelined merge kernel with PDL -- parity + timing; ncu of the NMS kernels in the bench regime (eager step).
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "merge or enhancer or predictor" > gpurun_out/c14_pytest_merge.log 2>&1; echo "rc=$?" >> gpurun_out/c14_pytest_merge.log
timeout 120 python tools/prof_merge.py > gpurun_out/c14_merge.json 2> gpurun_out/c14_merge.err
EL_PDL=0 timeout 120 python tools/prof_merge.py > gpurun_out/c14_merge_nopdl.json 2>> gpurun_out/c14_merge.err
timeout 120 python tools/prof_merge.py s > gpurun_out/c14_merge_s.json 2>> gpurun_out/c14_merge.err
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"nms_sweep|sort_single" -c 6 -o gpurun_out/c14_nms python bench.py --no-extras --no-cpu-baseline --no-ref-gpu --sustained-seconds 0 --steps 2 --warmup 3 --no-profile --no-graph > gpurun_out/c14_ncu_nms.log 2>&1
true
