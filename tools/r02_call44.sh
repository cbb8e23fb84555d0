#!/usr/bin/env bash
# Round 2, call 44: warp-uniform TMA / MMA issue loops (halo 3x3, pwconv, linattn_tma, dsconv3): full GPU suite, per-kernel profiles, bench with
# the fused depthwise -> pointwise kernel on maps <= 40 x 40, and with it off.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/c44_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/c44_pytest.log
timeout 200 python tools/prof_conv3x3.py > gpurun_out/c44_prof_conv3x3.json 2> gpurun_out/c44_prof_conv3x3.err
timeout 200 python tools/prof_dsconv.py > gpurun_out/c44_prof_dsconv.jsonl 2> gpurun_out/c44_prof_dsconv.err
timeout 200 python tools/prof_attn.py > gpurun_out/c44_prof_attn.json 2> gpurun_out/c44_prof_attn.err
timeout 300 python bench.py --no-extras --no-cpu-baseline --no-ref-gpu --sustained-seconds 0 > gpurun_out/c44_bench.json 2> gpurun_out/c44_bench.err
EL_DS3_MAX_HW=0 timeout 300 python bench.py --no-extras --no-cpu-baseline --no-ref-gpu --sustained-seconds 0 --no-profile > gpurun_out/c44_bench_nods3.json 2> gpurun_out/c44_bench_nods3.err
EL_DS3_MAX_HW=100000 timeout 300 python bench.py --no-extras --no-cpu-baseline --no-ref-gpu --sustained-seconds 0 --no-profile > gpurun_out/c44_bench_allds3.json 2> gpurun_out/c44_bench_allds3.err
true
