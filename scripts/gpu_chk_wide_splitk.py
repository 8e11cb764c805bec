import os, sys, torch
sys.path.insert(0, os.getcwd())
from edgedict_b200 import ops
torch.manual_seed(0)
for (M, N, K) in [(640, 1024, 16384), (640, 1024, 100000 - 96)]:
    a = (torch.randn(K, M, device="cuda") / 8).bfloat16()
    b = (torch.randn(K, N, device="cuda") / 8).bfloat16()
    out = ops.gemm_bf16(a, 1, b, 1, M, N, K)
    ref = a.float().t() @ b.float()
    err = (out - ref).abs().max().item() / ref.abs().max().item()
    acc = torch.full((M, N), 1.5, device="cuda")
    ops.gemm_bf16(a, 1, b, 1, M, N, K, out=acc, accumulate=True)
    err2 = (acc - 1.5 - ref).abs().max().item() / ref.abs().max().item()
    print(M, N, K, "rel err", err, err2, flush=True)
