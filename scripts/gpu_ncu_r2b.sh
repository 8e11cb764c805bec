#!/bin/bash
# ncu cannot keep all CTAs of a cooperative CLUSTER launch co-resident (the kernels trap in their barrier watchdog), so
# the launch list and the recurrent-kernel metrics are taken with the non-cluster variants (EDGEDICT_LSTM_C4=0
# EDGEDICT_LSTM_CLUSTER=0); the production cluster kernels are characterised by the clock64 stage traces instead.
mkdir -p gpurun_out
export EDGEDICT_LSTM_C4=0 EDGEDICT_LSTM_CLUSTER=0
MET=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,l1tex__m_xbar2l1tex_read_bytes.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct,sm__inst_executed.sum.per_cycle_active,launch__registers_per_thread,launch__grid_size,launch__block_size
cat > /tmp/rec_two.py <<'PY'
import os, sys, torch
sys.path.insert(0, os.getcwd())
from edgedict_b200 import ops
B, T, H = 32, 250, 1024
torch.manual_seed(0)
dev = "cuda"
xg = torch.randn(B, T, 4 * H, device=dev)
w = (torch.randn(4 * H, H, device=dev) / 32).bfloat16()
wT = w.t().contiguous()
for _ in range(2):
    y, y16, hT, cT, gates, cseq = ops.lstm_tc_fwd(xg, w, None, None, True)
    dy = torch.randn_like(y)
    ops.lstm_tc_bwd(dy, gates, cseq, None, wT, None, None)
torch.cuda.synchronize()
PY
timeout 600 ncu --replay-mode application --metrics $MET --clock-control none -k regex:lstm_tc_bwd_kernel -s 1 -c 1 -f -o gpurun_out/prof_r2_lstm_tc_bwd_noncluster python /tmp/rec_two.py > gpurun_out/ncu_r2_bwd.log 2>&1
echo "ncu bwd exit $?"; tail -2 gpurun_out/ncu_r2_bwd.log
EB_BENCH_MIN_WARMUP=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launches_bench.log 2>&1
echo "launch list exit $?"; tail -2 gpurun_out/ncu_launches_bench.log | cut -c1-300; wc -l gpurun_out/launches.csv
