#!/usr/bin/env bash
# Round 2, call 21: producer / consumer sweep: bit-exactness suite (repeated: the protocol is timing dependent), timings.
set -u
mkdir -p gpurun_out
for r in 1 2 3; do
timeout 600 python -m pytest tests -m gpu -q -k "nms or detect or predictor or map or decode or smoke" > gpurun_out/c21_pytest_$r.log 2>&1; echo "rc=$?" >> gpurun_out/c21_pytest_$r.log
done
timeout 200 python tools/prof_detect.py > gpurun_out/c21_detect.jsonl 2> gpurun_out/c21_detect.err
timeout 200 python tools/prof_detect.py --stress >> gpurun_out/c21_detect.jsonl 2>> gpurun_out/c21_detect.err
EL_NMS_PC=0 timeout 200 python tools/prof_detect.py >> gpurun_out/c21_detect.jsonl 2>> gpurun_out/c21_detect.err
true
