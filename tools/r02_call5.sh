#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "conv3x3_halo" > gpurun_out/c5_pytest_halo.log 2>&1; echo "rc=$?" >> gpurun_out/c5_pytest_halo.log
timeout 200 python tools/prof_conv3x3.py > gpurun_out/c5_conv3x3.json 2> gpurun_out/c5_conv3x3.err
timeout 100 python tools/prof_attn.py > gpurun_out/c5_attn.json 2> gpurun_out/c5_attn.err
timeout 300 python -m pytest tests -m gpu -q -x --deselect tests/test_reference_api_gpu.py -k "not conv3x3_halo" > gpurun_out/c5_pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/c5_pytest_gpu.log
timeout 200 python bench.py --no-extras --no-cpu-baseline > gpurun_out/c5_bench.json 2> gpurun_out/c5_bench.err
true
