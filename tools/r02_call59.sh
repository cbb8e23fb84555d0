#!/usr/bin/env bash
# Round 2, call 59: head towers of P3 / P4 started at their feature layer (parallel graph branches beside the bottom-up neck): tests + bench A/B.
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "predictor or smoke or map or whole_model or engine" > gpurun_out/c59_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/c59_pytest.log
timeout 300 python bench.py --no-extras --no-cpu-baseline --no-ref-gpu --sustained-seconds 0 --no-profile > gpurun_out/c59_bench.json 2> gpurun_out/c59_bench.err
timeout 300 python - > gpurun_out/c59_bench_off.json 2> gpurun_out/c59_bench_off.err <<'P'
import sys, runpy
import edge_yolo_b200.modules as M
M.HEAD_EARLY_LEVELS = False
sys.argv = ["bench.py", "--no-extras", "--no-cpu-baseline", "--no-ref-gpu", "--sustained-seconds", "0", "--no-profile"]
runpy.run_path("bench.py", run_name="__main__")
P
true
