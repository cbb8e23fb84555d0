#!/usr/bin/env bash
# Round 2, call 19: clock64 trace of the sweep's serial chain (traced build of nms.cu swapped in on the box only).
set -u
mkdir -p gpurun_out
cp edge_yolo_b200/libedgeline_b200.so /tmp/lib_backup.so
cp tools/_trace/libedgeline_b200_trace.so edge_yolo_b200/libedgeline_b200.so
timeout 200 python tools/prof_detect.py --iters 1 > gpurun_out/c19_trace.log 2>&1
timeout 200 python tools/prof_detect.py --iters 1 --stress > gpurun_out/c19_trace_stress.log 2>&1
cp /tmp/lib_backup.so edge_yolo_b200/libedgeline_b200.so
true
