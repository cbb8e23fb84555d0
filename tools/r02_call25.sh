#!/usr/bin/env bash
# Round 2, call 25: ap_per_class / scale_boxes kernels (f-4), plus the reference-API val test with metrics=True if present.
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_metrics_gpu.py -m gpu -q > gpurun_out/c25_pytest_metrics.log 2>&1; echo "rc=$?" >> gpurun_out/c25_pytest_metrics.log
timeout 600 python -m pytest tests/test_reference_api_gpu.py -m gpu -q > gpurun_out/c25_pytest_refapi.log 2>&1; echo "rc=$?" >> gpurun_out/c25_pytest_refapi.log
true
