#!/usr/bin/env bash
# Round 2, call 34: ncu of the uint8 stem kernel (instruction mix).
set -u
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"stem_tc" -c 1 -o gpurun_out/c34_stem python tools/prof_stem_one.py > gpurun_out/c34_ncu.log 2>&1
true
