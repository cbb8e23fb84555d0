#!/usr/bin/env bash
# Round 2, call 41: narrow 3x3 conv on mma.sync (f_h of the early enhancers): parity, engine tests, bench with the per-kernel pass.
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -k "conv3x3 or predictor or map or smoke or enhancer" > gpurun_out/c41_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/c41_pytest.log
timeout 300 python bench.py --no-extras --no-cpu-baseline --no-ref-gpu --sustained-seconds 0 > gpurun_out/c41_bench.json 2> gpurun_out/c41_bench.err
true
