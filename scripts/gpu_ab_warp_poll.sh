#!/bin/bash
# per-warp polling of the grid-barrier counter: 0 = one poller + block barrier, 1 = forward kernel, 2 = BPTT kernel, 3 = both
for rep in 1 2; do for w in 0 1 2 3; do
EDGEDICT_LSTM_WPOLL=$w python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/p21.json 2>gpurun_out/p21.err; python -c "
import json
d=json.load(open('gpurun_out/p21.json')); print('wpoll $w:', d['ms_per_step'], d['e2e']['ms_per_step'], {k:round(v['ms_per_step'],2) for k,v in d['kernels'].items() if k in ('lstm_tc_bwd','lstm_tc_fwd')}, d['loss_first'])"; done; done
