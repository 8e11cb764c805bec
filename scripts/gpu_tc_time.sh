#!/bin/bash
# timing experiments on the tensor-core LSTM forward kernel (debug flags; results are NOT valid numerics)
mkdir -p gpurun_out; rm -f gpurun_out/tc_time.log
for dbg in ${TC_DBGS:-0 1 2 3 4 8 16 24}; do
EDGEDICT_TC_DBG=$dbg timeout 120 python - >> gpurun_out/tc_time.log 2>&1 <<'PY'
import os, sys, numpy as np, torch
sys.path.insert(0, os.getcwd())
from edgedict_b200 import ops
B, T, H = 32, 500, 1024
torch.manual_seed(0)
xg = torch.randn(B, T, 4 * H, device="cuda")
whh16 = (torch.randn(4 * H, H, device="cuda") / 32).bfloat16()
for save in (True,):
    for _ in range(2):
        ops.lstm_tc_fwd(xg, whh16, None, None, save)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        ops.lstm_tc_fwd(xg, whh16, None, None, save)
    e1.record(); torch.cuda.synchronize()
    print("dbg", os.environ["EDGEDICT_TC_DBG"], "save", save, "%.2f us/step" % (e0.elapsed_time(e1) / 5 / T * 1e3), flush=True)
PY
done
cat gpurun_out/tc_time.log
