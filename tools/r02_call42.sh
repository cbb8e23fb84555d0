#!/usr/bin/env bash
# Round 2, call 42: fused depthwise 3x3 -> pointwise kernel (el_dsconv3_fwd): parity, per-site A/B against the two-kernel path; any-C depthwise TMA kernel.
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "dsconv3 or dwconv" > gpurun_out/c42_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/c42_pytest.log
timeout 300 python tools/prof_dsconv.py > gpurun_out/c42_prof_dsconv.jsonl 2> gpurun_out/c42_prof_dsconv.err
timeout 300 python bench.py --no-extras --no-cpu-baseline --no-ref-gpu --sustained-seconds 0 > gpurun_out/c42_bench.json 2> gpurun_out/c42_bench.err
true
