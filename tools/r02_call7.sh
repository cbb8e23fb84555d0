#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 200 ncu --set full --clock-control none --import-source on -k regex:conv3x3_halo -c 1 -s 2 -f -o gpurun_out/c7_halo python tools/prof_conv3x3_one.py > gpurun_out/c7_ncu_halo.log 2>&1
timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "wavelet_mixer or haar or dfl" > gpurun_out/c7_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/c7_pytest.log
true
