#!/bin/bash
# full-set ncu captures of the joint/loss kernels of the bf16 training step on a reduced cell count (ncu replays each
# kernel ~40 times; the full E6D2 shape would take minutes per kernel)
mkdir -p gpurun_out
cat > /tmp/jl_one.py <<'PY'
import os, sys, torch
sys.path.insert(0, os.getcwd())
from edgedict_b200 import functional as Fn
B, T, U, E, D, J, V = 8, 250, 129, 640, 256, 640, 1024
torch.manual_seed(0)
dev = "cuda"
ins = [torch.randn(B, T, E, device=dev), torch.randn(B, U, D, device=dev), torch.randn(J, E + D, device=dev) / 30,
       torch.zeros(J, device=dev), torch.randn(V, J, device=dev) / 25, torch.zeros(V, device=dev)]
ins = [t.requires_grad_(True) for t in ins]
labels = torch.randint(1, V, (B, U - 1), device=dev, dtype=torch.int32)
xl = torch.full((B,), T, device=dev, dtype=torch.int32)
yl = torch.full((B,), U - 1, device=dev, dtype=torch.int32)
for _ in range(2):
    loss, _ = Fn.JointLoss.apply(*ins, labels, xl, yl, 0, "bf16")
    loss.backward()
torch.cuda.synchronize()
PY
# per JointLoss fwd+bwd: 9 tcgen05 GEMM launches (fwd: e-proj, d-proj, logits+LSE; bwd: dW2, d-hidden+tanh', ...):
# launches 11..13 are the second iteration's logits+LSE, dW2 and d-hidden GEMMs
cap() {  # name skip count
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:$1 -s $2 -c $3 -f \
      -o gpurun_out/prof_r1f_$1 python /tmp/jl_one.py > gpurun_out/ncu_r1f_$1.log 2>&1
  echo "ncu $1 exit $?"
}
cap gemm_tc_kernel 11 3
cap rnnt_grad_bf16x8 1 1
cap joint_hidden_fwd_bf16 1 1
cap colsum_bf16_vec 2 1
