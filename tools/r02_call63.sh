#!/usr/bin/env bash
# Round 2, call 63: NMS sweep with 512 threads per image (co-residency with the forward graph): bit-exactness tests, detect-chain times, bench.
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "nms or detect or predictor" > gpurun_out/c63_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/c63_pytest.log
timeout 300 python bench.py --no-extras --no-cpu-baseline --no-ref-gpu --sustained-seconds 0 > gpurun_out/c63_bench.json 2> gpurun_out/c63_bench.err
true
