#!/usr/bin/env bash
# Round 2, call 10 (re-entry): full GPU suite, smoke, the driver's bench line and the reference arm on the restored build.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/c10_pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/c10_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/c10_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/c10_smoke.log
( time timeout 900 python bench.py > gpurun_out/c10_bench.json 2> gpurun_out/c10_bench.err ) 2> gpurun_out/c10_bench.time
( time timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/c10_bench_ref.json 2> gpurun_out/c10_bench_ref.err ) 2> gpurun_out/c10_bench_ref.time
true
