#!/usr/bin/env bash
# Round 2, call 53: fused depthwise -> pointwise kernel with two epilogue groups: parity, per-site A/B.
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "dsconv3 or predictor or smoke" > gpurun_out/c53_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/c53_pytest.log
timeout 300 python tools/prof_dsconv.py > gpurun_out/c53_prof_dsconv.jsonl 2> gpurun_out/c53_prof_dsconv.err
true
