#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
for i in 1 2 3 4; do
python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/pf.json 2>gpurun_out/pf.err; python -c "
import json
d=json.load(open('gpurun_out/pf.json')); print('run $i:', d['ms_per_step'], d['e2e']['ms_per_step'], d['clocks'])"; done
python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; tail -2 gpurun_out/final_bench.err
python - <<'PY'
import json
txt=open('gpurun_out/final_bench.json').read()
d=json.loads([l for l in txt.splitlines() if l.startswith('{')][-1])
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'launches', d['gpu_launches'], d['clocks'])
print('cpu_baseline', d.get('cpu_baseline',{}).get('value'), 'strong', d.get('strong_scaling',{}).get('ms_per_optimizer_step'))
PY
