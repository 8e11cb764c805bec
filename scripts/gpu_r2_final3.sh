#!/bin/bash
# final tree: smoke(), the default bench line (with its cpu_baseline leg), the CUPTI kernel timeline
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
( time python bench.py ) > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; tail -4 gpurun_out/final_bench.err
python - <<'PY'
import json
txt=open('gpurun_out/final_bench.json').read()
d=json.loads([l for l in txt.splitlines() if l.startswith('{')][-1])
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'launches', d['gpu_launches'])
print('roofline', d['roofline'])
print('cpu_baseline', d.get('cpu_baseline'))
print('parity_probe', d.get('parity_probe'), 'strong', d.get('strong_scaling'))
print({k:(round(v['ms_per_step'],2), v['calls_per_step']) for k,v in d['kernels'].items()})
PY
timeout 400 python scripts/gpu_timeline.py 2>&1 | tail -2; gzip -f gpurun_out/timeline_trace.json
