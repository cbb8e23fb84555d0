#!/usr/bin/env bash
# Round 2, call 33: record run -- full GPU suite, smoke, bench (all legs), reference arm, one eager step under ncu (time + DRAM bytes per launch).
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/c33_pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/c33_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/c33_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/c33_smoke.log
( time timeout 900 python bench.py > gpurun_out/c33_bench.json 2> gpurun_out/c33_bench.err ) 2> gpurun_out/c33_bench.time
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv \
  --log-file gpurun_out/c33_step.csv python bench.py --profile-step --no-graph --no-extras --no-cpu-baseline --no-ref-gpu --sustained-seconds 0 --no-cudnn-benchmark \
  > gpurun_out/c33_ncu_step.log 2>&1
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/c33_bench_ref.json 2> gpurun_out/c33_bench_ref.err
true
