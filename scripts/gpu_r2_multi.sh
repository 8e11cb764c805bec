#!/bin/bash
# N-GPU check of the data-parallel path: weak scaling (default line, with the strong_scaling leg inside) and --scaling strong
N=${1:-2}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_${N}gpu_weak.json 2> gpurun_out/bench_${N}gpu_weak.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 3 --warmup 3 --scaling strong > gpurun_out/bench_${N}gpu_strong.json 2> gpurun_out/bench_${N}gpu_strong.err
python - $N <<'PY'
import json, sys
n = sys.argv[1]
for kind in ("weak", "strong"):
    txt = open("gpurun_out/bench_%sgpu_%s.json" % (n, kind)).read()
    d = json.loads([l for l in txt.splitlines() if l.startswith("{")][-1])
    print(kind, "N=%s" % n, d["value"], d["ms_per_step"], d.get("strong_scaling"))
PY
