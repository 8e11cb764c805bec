#!/bin/bash
# round-2 probe 1: bf16 bench-mode parity (measured errors), BPTT REMAP map, GPU test suite, bench baseline
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity_bf16.py -x -q -s > gpurun_out/p1_parity.log 2>&1; echo "parity rc=$?"
tail -40 gpurun_out/p1_parity.log
python scripts/gpu_tc_overlap.py fwd bwd > gpurun_out/p1_overlap_default.log 2>&1
EDGEDICT_LSTM_BWD_REMAP=1 python scripts/gpu_tc_overlap.py bwd > gpurun_out/p1_overlap_remap.log 2>&1
cat gpurun_out/p1_overlap_default.log gpurun_out/p1_overlap_remap.log
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/p1_gputests.log 2>&1; tail -5 gpurun_out/p1_gputests.log
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/p1_bench.json 2> gpurun_out/p1_bench.err; tail -c 1500 gpurun_out/p1_bench.json
EDGEDICT_LSTM_BWD_REMAP=1 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/p1_bench_remap.json 2> gpurun_out/p1_bench_remap.err; python -c "
import json
for f in ('gpurun_out/p1_bench.json','gpurun_out/p1_bench_remap.json'):
    d=json.load(open(f)); print(f, d['ms_per_step'], {k:v['ms_per_step'] for k,v in d['kernels'].items() if v['ms_per_step']>0.5})
"
