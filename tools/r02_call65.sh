#!/usr/bin/env bash
# Round 2, call 65: ncu --set full of the final linear-attention kernel, the task-aligned-assigner kernels, the decode kernel and the merge kernel
# (fresh captures of the final builds), plus the loss-kernel timings.
set -u
mkdir -p gpurun_out
timeout 200 ncu --set full --clock-control none --import-source on -k regex:linattn_tma_kernel -c 2 -f -o gpurun_out/c65_linattn python tools/prof_attn.py > gpurun_out/c65_ncu_linattn.log 2>&1
timeout 300 ncu --set full --clock-control none -k regex:"tal_" -c 6 -f -o gpurun_out/c65_tal python tools/prof_loss.py > gpurun_out/c65_ncu_tal.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:gfl_decode_emit_warp_kernel -c 1 -f -o gpurun_out/c65_decode python bench.py --steps 1 --warmup 1 --no-extras --no-cpu-baseline --no-ref-gpu --sustained-seconds 0 --no-profile --profile-step --no-graph --no-cudnn-benchmark > gpurun_out/c65_ncu_decode.log 2>&1
timeout 200 python tools/prof_loss.py > gpurun_out/c65_loss_kernels.log 2>&1
timeout 200 python tools/prof_attn.py > gpurun_out/c65_prof_attn.json 2>/dev/null
true
