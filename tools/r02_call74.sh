#!/usr/bin/env bash
# Round 2, call 74: the driver's round-end sequence, verbatim flags.
set -u
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/ -x -q -m gpu > gpurun_out/c74_pytest.log 2>&1 ) 2> gpurun_out/c74_pytest.time; echo "rc=$?" >> gpurun_out/c74_pytest.log
( time timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c74_smoke.log 2>&1 ) 2> gpurun_out/c74_smoke.time
( time timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/c74_bench_ref.json 2> gpurun_out/c74_bench_ref.err ) 2> gpurun_out/c74_bench_ref.time
( time timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/c74_bench.json 2> gpurun_out/c74_bench.err ) 2> gpurun_out/c74_bench.time
true
