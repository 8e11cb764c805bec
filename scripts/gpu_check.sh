#!/bin/bash
# Runs on the B200 box through gpurun: GPU parity tests (one pytest process per file so that a hang in
# one file cannot take the others down), smoke(), a short bench.  Logs land in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt; free -g | head -2 >> gpurun_out/gpu.txt
for f in tests/test_gpu_loss.py tests/test_gpu_ops.py tests/test_gpu_model.py ${EXTRA_TESTS}; do
  n=$(basename $f .py)
  timeout 900 python -m pytest $f -m gpu -q --timeout=240 --timeout-method=thread -p no:cacheprovider > gpurun_out/$n.log 2>&1
  echo "== $f exit $?" | tee -a gpurun_out/summary.txt
  tail -5 gpurun_out/$n.log
done
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "== smoke exit $?" | tee -a gpurun_out/summary.txt
tail -3 gpurun_out/smoke.log
if [ -z "$SKIP_BENCH" ]; then
  timeout 900 python bench.py --steps ${BENCH_STEPS:-3} --warmup 3 ${BENCH_ARGS} > gpurun_out/bench.log 2>&1; echo "== bench exit $?" | tee -a gpurun_out/summary.txt
  tail -c 3000 gpurun_out/bench.log
fi
