#!/usr/bin/env bash
# Round 2, call 52: narrow 3x3 mma kernel with a three-tile staging ring: parity, per-site times, bench.
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "conv3x3 or predictor or enhancer" > gpurun_out/c52_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/c52_pytest.log
timeout 200 python tools/prof_conv3x3_mma.py > gpurun_out/c52_prof_c3m.json 2> gpurun_out/c52_prof_c3m.err
timeout 300 python bench.py --no-extras --no-cpu-baseline --no-ref-gpu --sustained-seconds 0 --no-profile > gpurun_out/c52_bench.json 2> gpurun_out/c52_bench.err
true
