#!/usr/bin/env bash
# Round 2, first GPU call: committed build's suite + the new reference-API tests, smoke, bench, the UMMA row-shift experiment, TAL ncu.
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q -x --deselect tests/test_reference_api_gpu.py > gpurun_out/c1_pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/c1_pytest_gpu.log
timeout 600 python -m pytest tests/test_reference_api_gpu.py -m gpu -q -s > gpurun_out/c1_pytest_refapi.log 2>&1; echo "rc=$?" >> gpurun_out/c1_pytest_refapi.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c1_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/c1_smoke.log
timeout 180 python bench.py > gpurun_out/c1_bench.json 2> gpurun_out/c1_bench.err
nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o /tmp/exp_umma_row_shift tools/exp_umma_row_shift.cu -lcuda \
  && timeout 60 /tmp/exp_umma_row_shift > gpurun_out/c1_exp_umma_row_shift.log 2>&1
timeout 90 ncu --set full --clock-control none --import-source on -k regex:tal_ -c 3 -f -o gpurun_out/c1_tal python tools/prof_loss.py \
  > gpurun_out/c1_ncu_tal.log 2>&1
nvidia-smi topo -m > gpurun_out/c1_topo.txt 2>&1; lscpu | head -30 >> gpurun_out/c1_topo.txt; numactl -H >> gpurun_out/c1_topo.txt 2>&1
true
