#!/usr/bin/env bash
# Round 2, call 48: which test file changes the outcome of test_half_mode_through_yolo_api
set -u
mkdir -p gpurun_out
: > gpurun_out/c48_bisect.log
for f in test_detection_loss test_install test_map_parity test_metrics_gpu test_gpu_parity; do
  echo "=== $f + half_mode" >> gpurun_out/c48_bisect.log
  timeout 300 python -m pytest tests/$f.py tests/test_reference_api_gpu.py::test_half_mode_through_yolo_api -m gpu -q -s 2>&1 | grep -E "fp16 model|passed|failed" >> gpurun_out/c48_bisect.log
done
true
