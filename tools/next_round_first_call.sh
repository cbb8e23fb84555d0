#!/usr/bin/env bash
# First GPU call of the next round (run through gpurun from the repo root, ~3 minutes of box time):
#   /usr/local/graft/bin/gpurun --timeout 420 -- 'bash tools/next_round_first_call.sh'
# 1. the GPU test suite and smoke() on the committed build, 2. the bench line, 3. the descriptor experiment that decides the design of the
# halo-tile 3x3 convolution, 4. the ncu capture of the rewritten assigner kernels that round 1 ran out of
# budget for.  Everything lands in gpurun_out/.
set -u
mkdir -p gpurun_out
timeout 120 python -m pytest tests -m gpu -q > gpurun_out/n0_pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/n0_pytest_gpu.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/n0_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/n0_smoke.log
timeout 120 python bench.py > gpurun_out/n0_bench.json 2> gpurun_out/n0_bench.err
nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o /tmp/exp_umma_row_shift tools/exp_umma_row_shift.cu \
  && timeout 60 /tmp/exp_umma_row_shift > gpurun_out/n0_exp_umma_row_shift.log 2>&1
timeout 60 ncu --set full --clock-control none --import-source on -k regex:tal_ -c 3 -f -o gpurun_out/n0_tal python tools/prof_loss.py \
  > gpurun_out/n0_ncu_tal.log 2>&1
true
