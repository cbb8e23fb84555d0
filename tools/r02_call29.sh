#!/usr/bin/env bash
# Round 2, call 29: vectorised DGQP weight staging: decode parity + timing (fused and dense kernels).
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -k "decode or detect or predictor or map or smoke" > gpurun_out/c29_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/c29_pytest.log
timeout 200 python tools/prof_detect.py > gpurun_out/c29_detect.jsonl 2> gpurun_out/c29_detect.err
timeout 200 python tools/prof_detect.py --stress >> gpurun_out/c29_detect.jsonl 2>> gpurun_out/c29_detect.err
timeout 200 python tools/bench_ref_eager.py > gpurun_out/c29_ref_eager.log 2>&1
true
