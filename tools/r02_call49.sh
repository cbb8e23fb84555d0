#!/usr/bin/env bash
# Round 2, call 49: which group of test_gpu_parity changes the outcome of test_half_mode_through_yolo_api
set -u
mkdir -p gpurun_out
: > gpurun_out/c49_bisect.log
for k in "dwt or merge or gated" "attention" "decode or detect" "nms" "qfl or dfl or loss or tal" "pwconv or conv3x3 or dwconv or dsconv3" "stem or sppf or bias_act or upsample or ingest" "predictor or engine or smoke or whole_model or model" "wtconv or haar or idwt"; do
  echo "=== $k" >> gpurun_out/c49_bisect.log
  timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_reference_api_gpu.py -m gpu -q -s -k "($k) or half_mode" 2>&1 | grep -E "fp16 model|passed|failed" >> gpurun_out/c49_bisect.log
done
true
