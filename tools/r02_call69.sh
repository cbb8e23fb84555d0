#!/usr/bin/env bash
# Round 2, call 69: halo kernel parity with several tiles per CTA (all epilogue groups, several rounds).
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "dsconv3 or conv3x3_halo" > gpurun_out/c69_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/c69_pytest.log
true
