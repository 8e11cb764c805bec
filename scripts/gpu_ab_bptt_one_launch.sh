#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q -x 2>&1 | tail -2
for rep in 1 2; do for one in 1 0; do
EDGEDICT_BPTT_ONE_LAUNCH=$one python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/p18.json 2>gpurun_out/p18.err; python -c "
import json
d=json.load(open('gpurun_out/p18.json')); print('one_launch $one:', d['ms_per_step'], d['e2e']['ms_per_step'], d['gpu_launches'], {k:round(v['ms_per_step'],2) for k,v in d['kernels'].items() if k in ('lstm_tc_bwd','lstm_tc_fwd')}, d['roofline']['us_per_timestep'])"; done; done
timeout 600 python -m pytest tests/test_gpu_parity_bf16.py tests/test_gpu_model.py -m gpu -q -x 2>&1 | tail -2
