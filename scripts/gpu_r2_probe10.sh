#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_parity_bf16.py tests/test_gpu_ops.py -q -x 2>&1 | tail -3
for i in 1 2; do
python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/p10_bench_$i.json 2>gpurun_out/p10_bench_$i.err; python -c "
import json
d=json.load(open('gpurun_out/p10_bench_$i.json')); print('pipelined bwd run $i:', d['ms_per_step'], d['e2e']['ms_per_step'], {k:v['ms_per_step'] for k,v in d['kernels'].items() if v['ms_per_step']>0.5})"; done
tail -3 gpurun_out/p10_bench_1.err
