#!/bin/bash
mkdir -p gpurun_out
timeout 200 python scripts/gpu_c4_check.py trace_bwd 2>&1 | tail -14
timeout 200 python scripts/gpu_c4_check.py check 2>&1 | grep -v "bwd" | tail -14
timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_model.py tests/test_gpu_parity_bf16.py tests/test_gpu_optim.py tests/test_gpu_stream.py -x -q 2>&1 | tail -5
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/p3_bench.json 2> gpurun_out/p3_bench.err; python -c "
import json
d=json.load(open('gpurun_out/p3_bench.json')); print('c4 fwd:', d['ms_per_step'], {k:v['ms_per_step'] for k,v in d['kernels'].items() if v['ms_per_step']>0.5}); print(d.get('parity_probe'), d.get('strong_scaling'))
"; tail -3 gpurun_out/p3_bench.err
EDGEDICT_LSTM_C4=0 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/p3_bench_old.json 2> gpurun_out/p3_bench_old.err; python -c "
import json
d=json.load(open('gpurun_out/p3_bench_old.json')); print('old fwd:', d['ms_per_step'], {k:v['ms_per_step'] for k,v in d['kernels'].items() if v['ms_per_step']>0.5})
"
