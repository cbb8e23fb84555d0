#!/usr/bin/env bash
# Round 2, call 2: full GPU suite on the new build (stages argument, in-place fix, reference-API tests), new bench line, training variants.
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -s --deselect tests/test_reference_api_gpu.py > gpurun_out/c2_pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/c2_pytest_gpu.log
timeout 600 python -m pytest tests/test_reference_api_gpu.py -m gpu -q -s > gpurun_out/c2_pytest_refapi.log 2>&1; echo "rc=$?" >> gpurun_out/c2_pytest_refapi.log
( time timeout 900 python bench.py > gpurun_out/c2_bench.json 2> gpurun_out/c2_bench.err ) 2> gpurun_out/c2_bench.time
( time timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/c2_bench_ref.json 2> gpurun_out/c2_bench_ref.err ) 2> gpurun_out/c2_bench_ref.time
for v in "--no-amp --nchw" "--nchw" "--no-amp" ""; do
  timeout 300 python tools/bench_train.py --steps 5 $v >> gpurun_out/c2_train_variants.jsonl 2>> gpurun_out/c2_train_variants.err
done
true
