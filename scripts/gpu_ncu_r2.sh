#!/bin/bash
# round-2 ncu evidence: (1) launch list of one training step of the final tree, (2) --set full of the streaming decode kernel,
# (3) the cluster / cooperative recurrent kernels (kernel replay cannot re-launch them: application replay, few metrics)
mkdir -p gpurun_out
MET=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,l1tex__m_xbar2l1tex_read_bytes.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct,sm__inst_executed.sum.per_cycle_active,launch__registers_per_thread,launch__grid_size,launch__block_size
cat > /tmp/rec_one.py <<'PY'
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
    y, hp, hT, cT, gates, cseq = ops.lstm_c4_fwd(xg, w, None, None, True, std_saves=True)
    dy = torch.randn_like(y)
    ops.lstm_tc_bwd(dy, gates, cseq, None, wT, None, None)
    ops.lstm_tc_fwd(xg, w, None, None, True)
torch.cuda.synchronize()
PY
for k in lstm_c4_fwd_kernel lstm_tc_bwd_kernel lstm_tc_fwd_kernel; do
  timeout 600 ncu --replay-mode application --metrics $MET --clock-control none -k regex:$k -s 1 -c 1 -f -o gpurun_out/prof_r2_$k python /tmp/rec_one.py > gpurun_out/ncu_r2_$k.log 2>&1
  echo "ncu $k exit $?"; tail -2 gpurun_out/ncu_r2_$k.log
done
cat > /tmp/dec_one.py <<'PY'
import os, sys, torch
sys.path.insert(0, os.getcwd())
from edgedict_b200.rnnt.models import Transducer
from edgedict_b200.stream_engine import StreamEngine
LARGE = dict(vocab_embed_size=64, vocab_size=1024, input_size=240, enc_hidden_size=1024, enc_layers=6, enc_dropout=0.0,
             enc_proj_size=640, dec_hidden_size=512, dec_layers=2, dec_dropout=0.1, dec_proj_size=640, joint_size=640)
torch.manual_seed(10)
m = Transducer(output_loss=False, **LARGE).eval().cuda()
eng = StreamEngine(m, 64, 2)
x = torch.randn(64, 2, 240, device="cuda")
for _ in range(4):
    eng.step(x)
torch.cuda.synchronize()
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:decode_program -s 3 -c 1 -f -o gpurun_out/prof_r2_decode_program python /tmp/dec_one.py > gpurun_out/ncu_r2_decode.log 2>&1
echo "ncu decode exit $?"; tail -2 gpurun_out/ncu_r2_decode.log
EB_BENCH_MIN_WARMUP=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launches_bench.log 2>&1
echo "launch list exit $?"; tail -2 gpurun_out/ncu_launches_bench.log; wc -l gpurun_out/launches.csv
