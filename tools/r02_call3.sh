#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 120 python tools/prof_attn.py > gpurun_out/c3_attn_new.json 2> gpurun_out/c3_attn_new.err; echo "rc=$?" >> gpurun_out/c3_attn_new.err
EL_LINATTN_NO_TMA=1 timeout 120 python tools/prof_attn.py > gpurun_out/c3_attn_old.json 2> gpurun_out/c3_attn_old.err
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "attention or decode_vs_oracle" > gpurun_out/c3_pytest_attn.log 2>&1; echo "rc=$?" >> gpurun_out/c3_pytest_attn.log
timeout 300 python -m pytest tests/test_reference_api_gpu.py -m gpu -q -s -k "half or predict" > gpurun_out/c3_pytest_refapi.log 2>&1; echo "rc=$?" >> gpurun_out/c3_pytest_refapi.log
true
