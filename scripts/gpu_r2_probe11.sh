#!/bin/bash
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
for i in 1 2; do
python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/p11_bench_$i.json 2>gpurun_out/p11_bench_$i.err; python -c "
import json
d=json.load(open('gpurun_out/p11_bench_$i.json')); print('TMEM-A + 4-warp issue run $i:', d['ms_per_step'], d['e2e']['ms_per_step'], {k:v['ms_per_step'] for k,v in d['kernels'].items() if v['ms_per_step']>0.5})"; done
EDGEDICT_C4_TMEMA=0 EDGEDICT_C4_NACC=1 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/p11_bench_old.json 2>/dev/null; python -c "
import json
d=json.load(open('gpurun_out/p11_bench_old.json')); print('smem A, 1 acc:', d['ms_per_step'], {k:v['ms_per_step'] for k,v in d['kernels'].items() if v['ms_per_step']>0.5})"
