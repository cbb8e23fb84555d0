#!/usr/bin/env bash
# Round 2, call 55: predict_many without the device-to-device input copy (per-set input buffers), SPPF pool with whole-sector accesses: tests + bench.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/c55_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/c55_pytest.log
timeout 300 python bench.py --no-extras --no-cpu-baseline --no-ref-gpu --sustained-seconds 0 > gpurun_out/c55_bench.json 2> gpurun_out/c55_bench.err
timeout 100 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/c55_smoke.log 2>&1
true
