#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/gpu_c4_check.py check > gpurun_out/p2_check.log 2>&1; echo "check rc=$?"; tail -30 gpurun_out/p2_check.log
timeout 300 python scripts/gpu_c4_check.py time > gpurun_out/p2_time.log 2>&1; echo "time rc=$?"; tail -8 gpurun_out/p2_time.log
