#!/usr/bin/env bash
# Round 2, call 20: sweep with two walkers per step: parity, timing, trace.
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -k "nms or detect or predictor" > gpurun_out/c20_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/c20_pytest.log
timeout 200 python tools/prof_detect.py > gpurun_out/c20_detect.jsonl 2> gpurun_out/c20_detect.err
timeout 200 python tools/prof_detect.py --stress >> gpurun_out/c20_detect.jsonl 2>> gpurun_out/c20_detect.err
cp edge_yolo_b200/libedgeline_b200.so /tmp/lib_backup.so
cp tools/_trace/libedgeline_b200_trace.so edge_yolo_b200/libedgeline_b200.so
timeout 200 python tools/prof_detect.py --iters 1 > gpurun_out/c20_trace.log 2>&1
cp /tmp/lib_backup.so edge_yolo_b200/libedgeline_b200.so
true
