"""Kernel timeline of one training step (E6D2 bench config) from torch.profiler / CUPTI: every kernel's stream, start and
duration, written as CSV to gpurun_out/timeline.csv (there is no nsys in the image).  The per-kernel events of bench.py
cannot show idle gaps or overlap; this can."""
import csv
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
import bench
from edgedict_b200.optim import FlatAdam
from edgedict_b200.rnnt.models import Transducer

dev = torch.device("cuda", 0)
torch.manual_seed(10)
model = Transducer(**bench.E6D2).to(dev)
model.set_precision("bf16")
opt = FlatAdam(model, lr=5e-4)
B, T, U, V = bench.B, bench.T, bench.U, bench.V
g = torch.Generator(device=dev).manual_seed(10)
xs = torch.randn(B, T, 240, device=dev, generator=g)
ys = torch.randint(4, V, (B, U), device=dev, dtype=torch.int32, generator=g)
xlen = torch.full((B,), T, dtype=torch.int32)
ylen = torch.full((B,), U, dtype=torch.int32)


def step():
    opt.zero_grad()
    loss = model(xs, ys, xlen, ylen)
    loss.backward()
    opt.step(grad_scale=1.0)
    return loss


for _ in range(3):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(2):
        step()
    torch.cuda.synchronize()
rows = []
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        rows.append((ev.time_range.start, ev.time_range.end - ev.time_range.start, getattr(ev, "device_index", 0),
                     getattr(ev, "stream", -1) if hasattr(ev, "stream") else -1, ev.name[:90]))
rows.sort()
os.makedirs("gpurun_out", exist_ok=True)
with open("gpurun_out/timeline.csv", "w", newline="") as fh:
    w = csv.writer(fh)
    w.writerow(["start_us", "dur_us", "device", "stream", "name"])
    for r in rows:
        w.writerow(r)
print("kernels:", len(rows), "span ms:", (rows[-1][0] + rows[-1][1] - rows[0][0]) / 1e3 if rows else 0)
try:
    prof.export_chrome_trace("gpurun_out/timeline_trace.json")
except Exception as e:
    print("trace export failed:", e)
