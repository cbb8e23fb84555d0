#!/usr/bin/env bash
# Round 2, call 62: host -> device rate per rank with eight ranks copying at once, three kinds of pinned host memory.
set -u
mkdir -p gpurun_out
cat /sys/kernel/mm/transparent_hugepage/enabled > gpurun_out/c62_thp.txt 2>&1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29535 tools/prof_h2d_modes.py > gpurun_out/c62_h2d_8.json 2> gpurun_out/c62_h2d_8.err
timeout 100 python tools/prof_h2d_modes.py > gpurun_out/c62_h2d_1.json 2> gpurun_out/c62_h2d_1.err
true
