#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -12
for ch in 4 6 8; do
EDGEDICT_WAVEFRONT_CHUNKS=$ch python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/p5_bench_c$ch.json 2> gpurun_out/p5_bench_c$ch.err; python -c "
import json,sys
d=json.load(open('gpurun_out/p5_bench_c$ch.json')); print('chunks $ch:', d['ms_per_step'], {k:v['ms_per_step'] for k,v in d['kernels'].items() if v['ms_per_step']>0.5}, 'launches', d['gpu_launches'])
"; done
EDGEDICT_PREDICTOR_STREAM=0 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/p5_bench_nopred.json 2>/dev/null; python -c "
import json
d=json.load(open('gpurun_out/p5_bench_nopred.json')); print('no predictor stream:', d['ms_per_step'])"
timeout 300 python scripts/bench_stream.py 2>&1 | tail -3
