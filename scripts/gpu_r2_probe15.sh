#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q -x 2>&1 | tail -2
for i in 1 2; do
python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/p15_bench_$i.json 2>gpurun_out/p15_bench_$i.err; python -c "
import json
d=json.load(open('gpurun_out/p15_bench_$i.json')); print('run $i:', d['ms_per_step'], d['e2e']['ms_per_step'], {k:round(v['ms_per_step'],2) for k,v in d['kernels'].items() if v['ms_per_step']>0.5})"; done
