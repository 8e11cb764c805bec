"""Probe: do two persistent tensor-core LSTM kernels (different layers) overlap when launched on two streams?"""
import os, sys, torch
sys.path.insert(0, os.getcwd())
from edgedict_b200 import ops
B, T, H = 32, 500, 1024
torch.manual_seed(0)
dev = "cuda"
def mk():
    xg = torch.randn(B, T, 4 * H, device=dev)
    whh16 = (torch.randn(4 * H, H, device=dev) / 32).bfloat16()
    whhT16 = whh16.t().contiguous()
    return xg, whh16, whhT16
sets = [mk(), mk()]
streams = [torch.cuda.Stream(), torch.cuda.Stream()]
saved = []
for xg, w, wT in sets:
    y, y16, hT, cT, gates, cseq = ops.lstm_tc_fwd(xg, w, None, None, True)
    saved.append((torch.randn_like(y), gates, cseq, wT))
torch.cuda.synchronize()
def run(kind, n):
    evs = []
    main = torch.cuda.current_stream()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n):
        s = streams[i]
        s.wait_stream(main)
        with torch.cuda.stream(s):
            for _ in range(3):
                if kind == "fwd":
                    ops.lstm_tc_fwd(sets[i][0], sets[i][1], None, None, True)
                else:
                    dy, gates, cseq, wT = saved[i]
                    ops.lstm_tc_bwd(dy, gates, cseq, None, wT, None, None)
    for i in range(n):
        main.wait_stream(streams[i])
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 3 / T * 1e3
for kind in (sys.argv[1:] or ("fwd", "bwd")):
    for n in (1, 2):
        run(kind, n)
        print(kind, "streams", n, "%.2f us per step (wall, per layer-step of ONE stream)" % run(kind, n), flush=True)
# mixed: fwd on one stream, bwd on the other
