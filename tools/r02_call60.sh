#!/usr/bin/env bash
# Round 2, call 60: eight ranks under the driver's torchrun launch (final build): the bench line only (no extras, no CPU leg).
set -u
mkdir -p gpurun_out
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 8 --steps 20 --warmup 5 --no-extras --no-cpu-baseline --no-ref-gpu --no-profile > gpurun_out/c60_bench_8gpu.json 2> gpurun_out/c60_bench_8gpu.err
nvidia-smi topo -m > gpurun_out/c60_topo.txt 2>&1
true
