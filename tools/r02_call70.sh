#!/usr/bin/env bash
# Round 2, call 70: closing run -- full GPU suite, smoke (engine section first), bench (all legs).
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/c70_pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/c70_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/c70_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/c70_smoke.log
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/c70_smoke_launches.csv python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c70_smoke_ncu.log 2>&1
( time timeout 900 python bench.py > gpurun_out/c70_bench.json 2> gpurun_out/c70_bench.err ) 2> gpurun_out/c70_bench.time
true
