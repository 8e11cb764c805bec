#!/bin/bash
# one full-set ncu capture of the BPTT kernel (L2-reduce variant: ncu cannot replay cooperative+cluster launches)
mkdir -p gpurun_out
export EDGEDICT_LSTM_CLUSTER=0
cat > /tmp/bwd_one.py <<'PY'
import os, sys, torch
sys.path.insert(0, os.getcwd())
from edgedict_b200 import ops
B, T, H = 32, int(os.environ.get("TT", "96")), 1024
torch.manual_seed(0)
xg = torch.randn(B, T, 4 * H, device="cuda")
whh16 = (torch.randn(4 * H, H, device="cuda") / 32).bfloat16()
y, y16, hT, cT, gates, cseq = ops.lstm_tc_fwd(xg, whh16, None, None, True)
dy = torch.randn_like(y)
for _ in range(2):
    ops.lstm_tc_bwd(dy, gates, cseq, None, whh16.t().contiguous(), None, None)
torch.cuda.synchronize()
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lstm_tc_bwd -s 1 -c 1 -f \
    -o gpurun_out/prof_bwd_r1b python /tmp/bwd_one.py > gpurun_out/ncu_bwd_r1b.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/ncu_bwd_r1b.log
