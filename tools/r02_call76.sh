#!/usr/bin/env bash
# Round 2, call 76: per-kernel table of EdgeLine-m (batch 128) and EdgeLine-s at 1280 (batch 32): where do the other scales spend their step?
set -u
mkdir -p gpurun_out
timeout 400 python bench.py --scale m --batch 128 --no-extras --no-cpu-baseline --no-ref-gpu --sustained-seconds 0 > gpurun_out/c76_bench_m.json 2> gpurun_out/c76_bench_m.err
timeout 400 python bench.py --scale s --imgsz 1280 --batch 32 --nc 10 --conf 0.001 --multi-label --no-extras --no-cpu-baseline --no-ref-gpu --sustained-seconds 0 > gpurun_out/c76_bench_s1280.json 2> gpurun_out/c76_bench_s1280.err
true
