#!/usr/bin/env bash
# Round 2, call 71: bench line twice (value-leg stability with the clock sampler started after the enqueue).
set -u
mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/c71_bench_a.json 2> gpurun_out/c71_bench_a.err
timeout 300 python bench.py --no-extras --no-cpu-baseline --no-ref-gpu --no-profile > gpurun_out/c71_bench_b.json 2> gpurun_out/c71_bench_b.err
timeout 300 python bench.py --no-extras --no-cpu-baseline --no-ref-gpu --no-profile > gpurun_out/c71_bench_c.json 2> gpurun_out/c71_bench_c.err
true
