#!/usr/bin/env bash
# Round 2, call 54: fused depthwise -> pointwise kernel at every site from 64 channels up: full suite, bench A/B.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/c54_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/c54_pytest.log
timeout 300 python bench.py --no-extras --no-cpu-baseline --no-ref-gpu --sustained-seconds 0 > gpurun_out/c54_bench.json 2> gpurun_out/c54_bench.err
EL_DS3_WIDE_C=100000 timeout 300 python bench.py --no-extras --no-cpu-baseline --no-ref-gpu --sustained-seconds 0 --no-profile > gpurun_out/c54_bench_small_only.json 2> gpurun_out/c54_bench_small_only.err
true
