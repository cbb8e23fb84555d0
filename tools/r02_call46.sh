#!/usr/bin/env bash
# Round 2, call 46: order dependence of test_half_mode_through_yolo_api
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_reference_api_gpu.py -m gpu -q -s 2>&1 | grep -E "fp16 model|passed|failed|Assertion" > gpurun_out/c46_module.log
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_reference_api_gpu.py -m gpu -q -s -k "dsconv3 or dwconv or reference_api or half_mode or yolo_api or uninstalled or nms_rows or train" 2>&1 | grep -E "fp16 model|passed|failed|Assertion" > gpurun_out/c46_with_ds.log
timeout 600 python -m pytest tests -m gpu -q -s 2>&1 | grep -E "fp16 model|passed|failed|Assertion" > gpurun_out/c46_full.log
true
