#!/usr/bin/env bash
# Round 2, call 64: record run of the final build -- full GPU suite, smoke, bench (all legs), reference arm, one eager step under ncu (time + DRAM
# bytes per launch), ncu --set full of the dominant kernel's largest sites, the fused depthwise -> pointwise kernel and the 3x3 halo kernel.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/c64_pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/c64_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/c64_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/c64_smoke.log
( time timeout 900 python bench.py > gpurun_out/c64_bench.json 2> gpurun_out/c64_bench.err ) 2> gpurun_out/c64_bench.time
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv \
  --log-file gpurun_out/c64_step.csv python bench.py --profile-step --no-graph --no-extras --no-cpu-baseline --no-ref-gpu --sustained-seconds 0 --no-cudnn-benchmark \
  > gpurun_out/c64_ncu_step.log 2>&1
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/c64_bench_ref.json 2> gpurun_out/c64_bench_ref.err
timeout 200 ncu --set full --clock-control none --import-source on -k regex:pwconv_tc_kernel -c 12 -f -o gpurun_out/c64_pwconv python bench.py --steps 1 --warmup 1 --no-extras --no-cpu-baseline --no-ref-gpu --sustained-seconds 0 --no-profile --profile-step --no-graph --no-cudnn-benchmark > gpurun_out/c64_ncu_pwconv.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:dsconv3_tc_kernel -c 1 -f -o gpurun_out/c64_dsconv3 python tools/prof_dsconv.py 64 80 > gpurun_out/c64_ncu_dsconv3.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:conv3x3_halo_kernel -s 1 -c 1 -f -o gpurun_out/c64_halo python tools/prof_halo_one.py > gpurun_out/c64_ncu_halo.log 2>&1
true
