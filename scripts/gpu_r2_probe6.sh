#!/bin/bash
mkdir -p gpurun_out
timeout 200 python scripts/gpu_c4_check.py trace 2>&1 | tail -14
timeout 200 python scripts/gpu_c4_check.py trace_bwd 2>&1 | tail -13
timeout 200 python scripts/gpu_c4_check.py time 2>&1 | grep fwd
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
EDGEDICT_WAVEFRONT_CHUNKS=6 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/p6_bench.json 2> gpurun_out/p6_bench.err; python -c "
import json,sys
d=json.load(open('gpurun_out/p6_bench.json')); print('bench chunks 6:', d['ms_per_step'], {k:v['ms_per_step'] for k,v in d['kernels'].items() if v['ms_per_step']>0.5}, 'launches', d['gpu_launches'])
"
timeout 300 python scripts/bench_stream.py 2>&1 | tail -2
