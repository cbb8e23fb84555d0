#!/usr/bin/env bash
# Round 2, call 18: sweep with branch-free IoU classes + float4 kept boxes: bit-exactness suite + timings.
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -k "nms or detect or predictor" > gpurun_out/c18_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/c18_pytest.log
timeout 200 python tools/prof_detect.py > gpurun_out/c18_detect.jsonl 2> gpurun_out/c18_detect.err
timeout 200 python tools/prof_detect.py --stress >> gpurun_out/c18_detect.jsonl 2>> gpurun_out/c18_detect.err
true
