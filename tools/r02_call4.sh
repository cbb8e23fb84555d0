#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 200 ncu --set full --clock-control none --import-source on -k regex:linattn_tma -c 2 -f -o gpurun_out/c4_attn python tools/prof_attn.py > gpurun_out/c4_ncu_attn.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:merge_fwd_x2 -c 8 -f -o gpurun_out/c4_merge python bench.py --steps 1 --warmup 1 --no-extras --no-cpu-baseline --no-profile --profile-step --no-graph > gpurun_out/c4_ncu_merge.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:gfl_decode_emit -c 2 -f -o gpurun_out/c4_decode python bench.py --steps 1 --warmup 1 --no-extras --no-cpu-baseline --no-profile --profile-step --no-graph > gpurun_out/c4_ncu_decode.log 2>&1
true
