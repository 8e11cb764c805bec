#!/bin/bash
# ncu passes (B200 box, via gpurun).  1: launch list of a short bench (kernel shares of the step);
# 2: full-set captures of the kernels named in $NCU_KERNELS (regex), one launch each.
mkdir -p gpurun_out
export EB_BENCH_MIN_WARMUP=1
export EDGEDICT_LSTM_CLUSTER=0   # ncu cannot replay cooperative+cluster launches: profile the L2-reduce variant
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv \
    --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
echo "== launch list exit $?" | tee -a gpurun_out/summary.txt
for k in ${NCU_KERNELS}; do
  timeout 1200 ncu --set full --clock-control none --import-source on -k regex:$k -s ${NCU_SKIP:-1} -c 1 -f \
      -o gpurun_out/prof_$k python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_$k.log 2>&1
  echo "== ncu $k exit $?" | tee -a gpurun_out/summary.txt
done
