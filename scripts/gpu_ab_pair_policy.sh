#!/bin/bash
# (EDGEDICT_GEMM_PAIR_LSE / EDGEDICT_GEMM_PAIR_NKB existed only for this A/B; the committed policy is in gemm_tc.cu: logits+LSE and
#  d-hidden on pair tiles automatically, split-K weight gradients only with eb_gemm_pair_mode(1))
# same-box A/B of the pair-tile policy: (LSE pair, dW2 pair) in {0,1}^2, twice
for rep in 1 2; do for cfg in "1 4096" "0 4096" "1 99999999" "0 99999999"; do
set -- $cfg
EDGEDICT_GEMM_PAIR_LSE=$1 EDGEDICT_GEMM_PAIR_NKB=$2 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/p17.json 2>gpurun_out/p17.err; python -c "
import json
d=json.load(open('gpurun_out/p17.json')); print('lse_pair $1 wgrad_nkb $2:', d['ms_per_step'], {k:round(v['ms_per_step'],2) for k,v in d['kernels'].items() if k in ('joint_logits_lse','gemm_bf16_tn','gemm_bf16_nn')})"; done; done
