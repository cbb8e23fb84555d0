#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 300 python tools/prof_e2e.py > gpurun_out/c75_prof_e2e.log 2>&1
true
