#!/bin/bash
# cta_group::2 tiles vs one-CTA tiles of the joint's two big bf16-output GEMMs: correctness, timing, then one ncu pass
# (4 launches: one-CTA logits+LSE, one-CTA d-hidden, pair logits+LSE, pair d-hidden)
mkdir -p gpurun_out
timeout 200 python scripts/gpu_pair_check.py check 2>&1 | tail -4
timeout 150 python scripts/gpu_pair_check.py time 2>&1 | tail -3
timeout 600 ncu --set full --import-source on --clock-control none -k regex:gemm_tc_kernel -c 4 -f -o gpurun_out/prof_r2_gemm_pair python scripts/gpu_pair_check.py once > gpurun_out/ncu_pair.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/ncu_pair.log
