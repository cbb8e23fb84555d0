#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "two_epilogue_groups" > gpurun_out/c73_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/c73_pytest.log
true
