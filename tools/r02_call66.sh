#!/usr/bin/env bash
# Round 2, call 66: four ranks under the driver's torchrun launch (final build): the bench line only.
set -u
mkdir -p gpurun_out
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29536 bench.py --gpus 4 --steps 20 --warmup 5 --no-extras --no-cpu-baseline --no-ref-gpu --no-profile > gpurun_out/c66_bench_4gpu.json 2> gpurun_out/c66_bench_4gpu.err
true
