#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "dwconv" > gpurun_out/c8_pytest_dw.log 2>&1; echo "rc=$?" >> gpurun_out/c8_pytest_dw.log
timeout 200 python tools/prof_dwconv.py > gpurun_out/c8_dw_new.json 2> gpurun_out/c8_dw_new.err
EL_DW_TMA=0 timeout 200 python tools/prof_dwconv.py > gpurun_out/c8_dw_old.json 2> gpurun_out/c8_dw_old.err
timeout 200 python tools/prof_conv3x3.py > gpurun_out/c8_conv3x3.json 2> gpurun_out/c8_conv3x3.err
timeout 200 python bench.py --no-extras --no-cpu-baseline > gpurun_out/c8_bench.json 2> gpurun_out/c8_bench.err
true
