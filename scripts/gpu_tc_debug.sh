#!/bin/bash
# Stage-by-stage bring-up of the tensor-core LSTM kernels with short timeouts (gpurun).
mkdir -p gpurun_out
python - > gpurun_out/tc_probe.log 2>&1 <<'PY'
from edgedict_b200._lib import lib
L = lib()
for H in (64, 256, 320, 512, 1024):
    print("H", H, "co-resident clusters (size: max/needed):", {cs: (L.eb_lstm_tc_max_clusters(H, cs), H // (8 * cs)) for cs in (8, 4, 2)})
PY
cat gpurun_out/tc_probe.log
for mode in ${TC_MODES:-0 2 4 8}; do
  for case in "4 6 32 64" "32 9 64 256" "32 4 64 1024" "32 500 64 1024"; do
    set -- $case
    EDGEDICT_LSTM_CLUSTER=$mode timeout 90 python - $1 $2 $3 $4 >> gpurun_out/tc_debug.log 2>&1 <<'PY'
import sys, os, time, numpy as np, torch
sys.path.insert(0, os.getcwd())
from edgedict_b200 import functional as Fn, ops
from oracle import model_torch as mt
B, T, I, H = map(int, sys.argv[1:5])
k = 1.0 / np.sqrt(H)
rb = lambda t: t.bfloat16().float()
torch.manual_seed(0)
w_ih, w_hh = rb((torch.rand(4*H, I)*2-1)*k), rb((torch.rand(4*H, H)*2-1)*k)
b_ih, b_hh = (torch.rand(4*H)*2-1)*k, (torch.rand(4*H)*2-1)*k
x = rb(torch.randn(B, T, I))
dy = torch.randn(B, T, H)
ins = [t.clone().cuda().requires_grad_(True) for t in (x, w_ih, w_hh, b_ih, b_hh)]
t0 = time.time()
y, hT, cT = Fn.LSTMLayer.apply(ins[0], None, None, ins[1], ins[2], ins[3], ins[4], "bf16")
torch.cuda.synchronize()
print("mode", os.environ["EDGEDICT_LSTM_CLUSTER"], (B, T, I, H), "fwd ok %.3fs" % (time.time()-t0), flush=True)
(y * dy.cuda()).sum().backward()
torch.cuda.synchronize()
print("   bwd ok %.3fs" % (time.time()-t0), flush=True)
if T >= 100:
    for name, fn in (("fwd", lambda: Fn.LSTMLayer.apply(ins[0], None, None, ins[1], ins[2], ins[3], ins[4], "bf16")),):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ops.PROF.reset(); ops.PROF.enabled = True
        for _ in range(3):
            yy, _, _ = fn()
            (yy * dy.cuda()).sum().backward()
        torch.cuda.synchronize()
        for k, v in ops.PROF.summary().items():
            if k.startswith("lstm"):
                print("   %s: %.3f ms/call -> %.2f us/step" % (k, v["ms"] / v["calls"], v["ms"] / v["calls"] / T * 1e3), flush=True)
        ops.PROF.enabled = False
if T <= 10:
    ref = [t.double().requires_grad_(True) for t in (x, w_ih, w_hh, b_ih, b_hh)]
    yr, _, _ = mt.lstm_layer(ref[0], torch.zeros(B, H).double(), torch.zeros(B, H).double(), *ref[1:])
    (yr * dy.double()).sum().backward()
    e = lambda a, b: float((a.double().cpu() - b).abs().max() / (b.abs().max() + 1e-30))
    print("   rel err y %.2e dx %.2e dw_ih %.2e dw_hh %.2e db %.2e" % (e(y.detach(), yr.detach()), e(ins[0].grad, ref[0].grad),
          e(ins[1].grad, ref[1].grad), e(ins[2].grad, ref[2].grad), e(ins[3].grad, ref[3].grad)), flush=True)
PY
    echo "mode $mode case $case exit $?" >> gpurun_out/tc_debug.log
  done
done
tail -40 gpurun_out/tc_debug.log
