#!/usr/bin/env bash
# Round 2, call 28: warp-autonomous decode emit kernel: parity + A/B timing + ncu.
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -k "decode or detect or predictor or map or smoke or api" > gpurun_out/c28_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/c28_pytest.log
timeout 200 python tools/prof_detect.py > gpurun_out/c28_detect.jsonl 2> gpurun_out/c28_detect.err
EL_DECODE_WARP=0 timeout 200 python tools/prof_detect.py >> gpurun_out/c28_detect.jsonl 2>> gpurun_out/c28_detect.err
timeout 200 python tools/prof_detect.py --stress >> gpurun_out/c28_detect.jsonl 2>> gpurun_out/c28_detect.err
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"gfl_decode_emit" -c 1 -o gpurun_out/c28_emit python tools/prof_detect.py --iters 1 > gpurun_out/c28_ncu.log 2>&1
true
