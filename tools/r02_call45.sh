#!/usr/bin/env bash
# Round 2, call 45: is test_half_mode_through_yolo_api sensitive to the attention kernel / run-to-run?
set -u
mkdir -p gpurun_out
for i in 1 2; do timeout 300 python -m pytest tests/test_reference_api_gpu.py -m gpu -q -s -k half_mode 2>&1 | grep -E "fp16 model|passed|failed|Assertion" >> gpurun_out/c45_half.log; done
echo "--- EL_LINATTN_NO_TMA=1" >> gpurun_out/c45_half.log
EL_LINATTN_NO_TMA=1 timeout 300 python -m pytest tests/test_reference_api_gpu.py -m gpu -q -s -k half_mode 2>&1 | grep -E "fp16 model|passed|failed|Assertion" >> gpurun_out/c45_half.log
timeout 600 python -m pytest tests -m gpu -q --deselect tests/test_reference_api_gpu.py::test_half_mode_through_yolo_api > gpurun_out/c45_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/c45_pytest.log
true
