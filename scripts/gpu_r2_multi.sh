#!/bin/bash
# 2-GPU check of the data-parallel path: weak scaling (default line, with the strong_scaling leg inside) and --scaling strong
N=${1:-2}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_${N}gpu_weak.json 2> gpurun_out/bench_${N}gpu_weak.err
python -c "
import json
d=json.loads([l for l in open("gpurun_out/bench_${N}gpu_weak.json").read().splitlines() if l.startswith("{")][-1]); print('weak N=$N', d['value'], d['ms_per_step'], d.get('strong_scaling'))"
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 3 --warmup 3 --scaling strong > gpurun_out/bench_${N}gpu_strong.json 2> gpurun_out/bench_${N}gpu_strong.err
python -c "
import json
d=json.loads([l for l in open("gpurun_out/bench_${N}gpu_strong.json").read().splitlines() if l.startswith("{")][-1]); print('strong N=$N', d['value'], d['ms_per_step'], d['config'])"
tail -3 gpurun_out/bench_${N}gpu_weak.err gpurun_out/bench_${N}gpu_strong.err
