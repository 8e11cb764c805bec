#!/bin/bash
for rep in 1 2; do for side in 1 0; do
EDGEDICT_JOINT_WGRAD_SIDE=$side python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/p19.json 2>gpurun_out/p19.err; python -c "
import json
d=json.load(open('gpurun_out/p19.json')); print('wgrad_side $side:', d['ms_per_step'], d['e2e']['ms_per_step'], d['loss_first'], d['loss_last'])"; done; done
timeout 600 python -m pytest tests/test_gpu_parity_bf16.py tests/test_gpu_model.py -m gpu -q -x 2>&1 | tail -2
