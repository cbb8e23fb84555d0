#!/usr/bin/env bash
# Round 2, call 15: rewritten NMS sweep (no fences / global stores on the chain, pair masks, band test) -- bit-exactness suite + timings in three regimes.
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -k "nms or detect or merge or predictor or map or api" > gpurun_out/c15_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/c15_pytest.log
timeout 200 python tools/prof_detect.py > gpurun_out/c15_detect.jsonl 2> gpurun_out/c15_detect.err
timeout 200 python tools/prof_detect.py --trained >> gpurun_out/c15_detect.jsonl 2>> gpurun_out/c15_detect.err
timeout 200 python tools/prof_detect.py --stress >> gpurun_out/c15_detect.jsonl 2>> gpurun_out/c15_detect.err
timeout 120 python tools/prof_merge.py > gpurun_out/c15_merge.json 2>> gpurun_out/c15_detect.err
true
