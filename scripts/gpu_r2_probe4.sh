#!/bin/bash
mkdir -p gpurun_out
timeout 200 python scripts/gpu_c4_check.py trace_bwd 2>&1 | tail -14
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/p4_bench.json 2> gpurun_out/p4_bench.err; python -c "
import json
d=json.load(open('gpurun_out/p4_bench.json')); print('bench:', d['ms_per_step'], {k:v['ms_per_step'] for k,v in d['kernels'].items() if v['ms_per_step']>0.5}, 'launches', d['gpu_launches']); print(d.get('parity_probe'), d.get('strong_scaling'))
"; tail -3 gpurun_out/p4_bench.err
