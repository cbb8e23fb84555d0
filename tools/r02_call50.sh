#!/usr/bin/env bash
# Round 2, call 50: 3x3 halo kernel with up to four epilogue groups (one TMEM accumulator each): parity, A/B against two groups, full suite, bench.
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "conv3x3" > gpurun_out/c50_pytest_c3.log 2>&1; echo "rc=$?" >> gpurun_out/c50_pytest_c3.log
timeout 200 python tools/prof_conv3x3.py > gpurun_out/c50_prof_conv3x3_g4.json 2> gpurun_out/c50_prof_conv3x3.err
EL_C3_GROUPS=2 timeout 200 python tools/prof_conv3x3.py > gpurun_out/c50_prof_conv3x3_g2.json 2>> gpurun_out/c50_prof_conv3x3.err
EL_C3_GROUPS=3 timeout 200 python tools/prof_conv3x3.py > gpurun_out/c50_prof_conv3x3_g3.json 2>> gpurun_out/c50_prof_conv3x3.err
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/c50_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/c50_pytest.log
timeout 300 python bench.py --no-extras --no-cpu-baseline --no-ref-gpu --sustained-seconds 0 > gpurun_out/c50_bench.json 2> gpurun_out/c50_bench.err
true
