#!/usr/bin/env bash
# Round 2, call 51: pwconv with a second epilogue group at one-CTA-per-SM sites; cv1 | cv2 of DSC3k as one GEMM: parity, bench A/B.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/c51_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/c51_pytest.log
timeout 300 python bench.py --no-extras --no-cpu-baseline --no-ref-gpu --sustained-seconds 0 > gpurun_out/c51_bench.json 2> gpurun_out/c51_bench.err
EL_PW_GROUPS=1 timeout 300 python bench.py --no-extras --no-cpu-baseline --no-ref-gpu --sustained-seconds 0 > gpurun_out/c51_bench_g1.json 2> gpurun_out/c51_bench_g1.err
true
