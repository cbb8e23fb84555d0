#!/usr/bin/env bash
# Round 2, call 68: parity of the two-epilogue-group pwconv path at the sizes that take it; full suite.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/c68_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/c68_pytest.log
true
