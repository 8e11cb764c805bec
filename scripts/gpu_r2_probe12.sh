#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_parity_bf16.py -q -x 2>&1 | tail -3
for cfg in "1 4" "0 1"; do set -- $cfg
EDGEDICT_C4_TMEMA=$1 EDGEDICT_C4_NACC=$2 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/p12_bench_$1$2.json 2>gpurun_out/p12_bench_$1$2.err; python -c "
import json
d=json.load(open('gpurun_out/p12_bench_$1$2.json')); print('TMEMA=$1 NACC=$2:', d['ms_per_step'], d['e2e']['ms_per_step'], {k:v['ms_per_step'] for k,v in d['kernels'].items() if v['ms_per_step']>0.5})"; done
for ch in 4 8; do
EDGEDICT_WAVEFRONT_CHUNKS=$ch python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/p12_bench_c$ch.json 2>/dev/null; python -c "
import json
d=json.load(open('gpurun_out/p12_bench_c$ch.json')); print('chunks $ch:', d['ms_per_step'], {k:v['ms_per_step'] for k,v in d['kernels'].items() if 'lstm' in k})"; done
