#!/usr/bin/env bash
# Round 2, call 30 (2 GPUs): the driver's multi-rank launch of bench.py (torchrun, NCCL), both arms.
set -u
mkdir -p gpurun_out
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/c30_bench_2gpu.json 2> gpurun_out/c30_bench_2gpu.err ) 2> gpurun_out/c30_bench_2gpu.time
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/c30_bench_ref_2gpu.json 2> gpurun_out/c30_bench_ref_2gpu.err ) 2> gpurun_out/c30_bench_ref_2gpu.time
nvidia-smi topo -m > gpurun_out/c30_topo.txt 2>&1
true
