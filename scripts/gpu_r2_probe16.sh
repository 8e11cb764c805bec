#!/bin/bash
timeout 250 python scripts/gpu_pair_check.py check 2>&1 | tail -5
timeout 150 python scripts/gpu_pair_check.py time 2>&1 | tail -2
for cfg in "4096" "64" "4096"; do
EDGEDICT_GEMM_PAIR_NKB=$cfg python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/p16_bench_$cfg.json 2>gpurun_out/p16_bench_$cfg.err; python -c "
import json
d=json.load(open('gpurun_out/p16_bench_$cfg.json')); print('nkb $cfg:', d['ms_per_step'], d['e2e']['ms_per_step'], {k:round(v['ms_per_step'],2) for k,v in d['kernels'].items() if v['ms_per_step']>0.5})"; done
