#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 200 python tools/prof_conv3x3.py > gpurun_out/c9_conv3x3.json 2> gpurun_out/c9_conv3x3.err
timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "conv3x3_halo" > gpurun_out/c9_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/c9_pytest.log
timeout 200 python bench.py --no-extras --no-cpu-baseline > gpurun_out/c9_bench.json 2> gpurun_out/c9_bench.err
true
