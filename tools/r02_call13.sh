#!/usr/bin/env bash
# Round 2, call 13: shared-memory staged merge kernel -- parity, A/B timing against the register kernel, ncu of both at the largest site.
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "merge" > gpurun_out/c13_pytest_merge.log 2>&1; echo "rc=$?" >> gpurun_out/c13_pytest_merge.log
timeout 120 python tools/prof_merge.py > gpurun_out/c13_merge_smem.json 2> gpurun_out/c13_merge_smem.err
EL_MERGE_SR=8 timeout 120 python tools/prof_merge.py > gpurun_out/c13_merge_smem_sr8.json 2>> gpurun_out/c13_merge_smem.err
EL_MERGE_RPC=8 timeout 120 python tools/prof_merge.py > gpurun_out/c13_merge_smem_rpc8.json 2>> gpurun_out/c13_merge_smem.err
EL_MERGE_SR=2 timeout 120 python tools/prof_merge.py > gpurun_out/c13_merge_smem_sr2.json
EL_MERGE_RPC=40 timeout 120 python tools/prof_merge.py > gpurun_out/c13_merge_smem_rpc40.json 2>> gpurun_out/c13_merge_smem.err 2>> gpurun_out/c13_merge_smem.err
timeout 120 python tools/prof_merge.py s > gpurun_out/c13_merge_smem_s.json 2>> gpurun_out/c13_merge_smem.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:merge_fwd -c 2 -o gpurun_out/c13_merge_smem python tools/prof_merge.py > gpurun_out/c13_ncu.log 2>&1
true
