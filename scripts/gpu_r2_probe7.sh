#!/bin/bash
for w in 0 20 40; do
EDGEDICT_WGRAD_CTAS=$w python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/p7_bench_w$w.json 2>/dev/null; python -c "
import json,sys
d=json.load(open('gpurun_out/p7_bench_w$w.json')); print('wgrad ctas $w:', d['ms_per_step'], {k:v['ms_per_step'] for k,v in d['kernels'].items() if v['ms_per_step']>0.5})
"; done
