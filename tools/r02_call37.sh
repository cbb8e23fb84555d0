#!/usr/bin/env bash
# Round 2, call 37: stem kernel with incremental tile coordinates: parity + timing.
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q -k "stem or predictor or smoke or map" > gpurun_out/c37_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/c37_pytest.log
timeout 120 python tools/prof_stem_one.py > gpurun_out/c37_stem.log 2>&1
timeout 300 python bench.py --no-extras --no-cpu-baseline --no-ref-gpu --sustained-seconds 0 --no-profile > gpurun_out/c37_bench.json 2> gpurun_out/c37_bench.err
true
