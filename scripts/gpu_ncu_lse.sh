#!/bin/bash
# one full-set ncu capture of the LSE-epilogue joint GEMM (optional fused path) on a reduced cell count
mkdir -p gpurun_out
cat > /tmp/lse_one.py <<'PY'
import os, sys, torch
sys.path.insert(0, os.getcwd())
from edgedict_b200 import ops
B, T, U, J, V = 8, 250, 129, 640, 1024
torch.manual_seed(0)
hid = torch.randn(B * T * U, J, device="cuda").bfloat16()
w2 = (torch.randn(V, J, device="cuda") / 25).bfloat16()
b2 = torch.randn(V, device="cuda") / 10
labels = torch.randint(1, V, (B, U - 1), device="cuda", dtype=torch.int32)
xl = torch.full((B,), T, device="cuda", dtype=torch.int32)
yl = torch.full((B,), U - 1, device="cuda", dtype=torch.int32)
for _ in range(3):
    out = ops.joint_logits_lse(hid, w2, b2, labels, xl, yl, B, T, U, 0)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    ops.joint_logits_lse(hid, w2, b2, labels, xl, yl, B, T, U, 0)
e1.record(); torch.cuda.synchronize()
print("lse gemm ms", e0.elapsed_time(e1) / 5, "cells", B * T * U)
lo = torch.empty(B * T * U, V, device="cuda")
for _ in range(2):
    ops.gemm_bf16(hid, 0, w2, 0, B * T * U, V, J, bias=b2, out=lo)
e0.record()
for _ in range(5):
    ops.gemm_bf16(hid, 0, w2, 0, B * T * U, V, J, bias=b2, out=lo)
e1.record(); torch.cuda.synchronize()
print("plain gemm ms", e0.elapsed_time(e1) / 5)
PY
python /tmp/lse_one.py
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 2 -c 1 -f \
    -o gpurun_out/prof_lse_gemm python /tmp/lse_one.py > gpurun_out/ncu_lse.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/ncu_lse.log
