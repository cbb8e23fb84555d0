#!/usr/bin/env bash
# Round 2, call 58: two ranks under the driver's torchrun launch (final build): bench line + reference arm.
set -u
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/c58_bench_2gpu.json 2> gpurun_out/c58_bench_2gpu.err
true
