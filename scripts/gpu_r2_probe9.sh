#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_parity_bf16.py -q -x 2>&1 | tail -3
for d in 1 0; do
EDGEDICT_DEFER_JOINT_WGRAD=$d EDGEDICT_WAVEFRONT_CHUNKS=6 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/p9_bench_d$d.json 2>gpurun_out/p9_bench_d$d.err; python -c "
import json
d=json.load(open('gpurun_out/p9_bench_d$d.json')); print('defer joint wgrad $d:', d['ms_per_step'], d['e2e']['ms_per_step'], {k:v['ms_per_step'] for k,v in d['kernels'].items() if v['ms_per_step']>0.5})"; done
tail -3 gpurun_out/p9_bench_d1.err
