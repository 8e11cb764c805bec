#!/bin/bash
# ncu --set full of the LSTM input-projection GEMM (the "LSTM gate GEMM" of the north star: xg = X W_ih^T, M = B*T = 32000, N = 4H = 4096,
# K = 1024) and of its dgrad / wgrad counterparts
mkdir -p gpurun_out
cat > /tmp/gemm_one.py <<'PY'
import os, sys, torch
sys.path.insert(0, os.getcwd())
from edgedict_b200 import ops
torch.manual_seed(0)
M, N, K = 32000, 4096, 1024
x = torch.randn(M, K, device="cuda").bfloat16()
w = (torch.randn(N, K, device="cuda") / 32).bfloat16()
b = torch.zeros(N, device="cuda")
dg = torch.randn(M, N, device="cuda").bfloat16()
for _ in range(3):
    ops.gemm_bf16(x, 0, w, 0, M, N, K, bias=b)            # xg = X W_ih^T (+ bias), fp32 out
    ops.gemm_bf16(dg, 0, w, 1, M, K, N)                   # dx = dG W_ih
    ops.gemm_bf16(dg, 1, x, 1, N, K, M)                   # dW_ih = dG^T X
torch.cuda.synchronize()
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 6 -c 3 -f -o gpurun_out/prof_r2_gemm_lstm python /tmp/gemm_one.py > gpurun_out/ncu_r2_gemm.log 2>&1
echo "exit $?"; tail -2 gpurun_out/ncu_r2_gemm.log
