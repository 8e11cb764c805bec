#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -k "dpre or joint" 2>&1 | tail -2
for rep in 1 2; do for f in 1 0; do
EDGEDICT_DPRE_FUSED=$f python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/p20.json 2>gpurun_out/p20.err; python -c "
import json
d=json.load(open('gpurun_out/p20.json')); print('dpre_fused $f:', d['ms_per_step'], d['e2e']['ms_per_step'], d['loss_first'])"; done; done
timeout 600 python -m pytest tests/test_gpu_parity_bf16.py tests/test_gpu_model.py -m gpu -q -x 2>&1 | tail -2
