#!/usr/bin/env bash
# Round 2, call 67: pwconv with 32-channel store boxes whenever they buy another resident CTA: parity, bench A/B.
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "pwconv or predictor or enhancer or smoke" > gpurun_out/c67_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/c67_pytest.log
timeout 300 python bench.py --no-extras --no-cpu-baseline --no-ref-gpu --sustained-seconds 0 > gpurun_out/c67_bench.json 2> gpurun_out/c67_bench.err
EL_PW_OB32=0 timeout 300 python bench.py --no-extras --no-cpu-baseline --no-ref-gpu --sustained-seconds 0 > gpurun_out/c67_bench_off.json 2> gpurun_out/c67_bench_off.err
true
