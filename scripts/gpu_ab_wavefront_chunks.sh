#!/bin/bash
# chunks of the forward wavefront, re-measured now that BPTT is one launch per layer
for c in 6 5 7 8 6; do
EDGEDICT_WAVEFRONT_CHUNKS=$c timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/p24.json 2>gpurun_out/p24.err; python -c "
import json
d=json.load(open('gpurun_out/p24.json')); print('chunks $c:', d['ms_per_step'], d['e2e']['ms_per_step'], d['gpu_launches'], {k:round(v['ms_per_step'],2) for k,v in d['kernels'].items() if k in ('lstm_tc_bwd','lstm_tc_fwd','gemm_bf16_nt')})"; done
