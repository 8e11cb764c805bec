"""Microbenchmark of the tcgen05 GEMM on the shapes of the E6D2 step (run on the B200 box)."""
import os, sys, json, torch
sys.path.insert(0, os.getcwd())
from edgedict_b200 import ops

PEAK = 1351.3
shapes = [
    ("joint logits  nt fp32-out", 0, 0, 2064000, 1024, 640, False),
    ("joint dhidden nn bf16-out", 0, 1, 2064000, 640, 1024, True),
    ("joint dW2     tn split-K ", 1, 1, 1024, 640, 2064000, False),
    ("joint dW2^T   tn split-K ", 1, 1, 640, 1024, 2064000, False),
    ("lstm xg       nt fp32-out", 0, 0, 32000, 4096, 1024, False),
    ("lstm dx       nn fp32-out", 0, 1, 32000, 1024, 4096, False),
    ("lstm dW       tn         ", 1, 1, 4096, 1024, 32000, False),
    ("lstm xg T/2   nt fp32-out", 0, 0, 16000, 4096, 1024, False),
    ("square 8192   nt bf16-out", 0, 0, 8192, 8192, 8192, True),
]
for name, amn, bmn, M, N, K, o16 in shapes:
    A = torch.randn((K, M) if amn else (M, K), device="cuda").bfloat16() if M * K < 3e9 else None
    B = torch.randn((K, N) if bmn else (N, K), device="cuda").bfloat16()
    out = torch.empty(M, N, dtype=torch.bfloat16 if o16 else torch.float32, device="cuda")
    for _ in range(2):
        ops.gemm_bf16(A, amn, B, bmn, M, N, K, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 5
    e0.record()
    for _ in range(n):
        ops.gemm_bf16(A, amn, B, bmn, M, N, K, out=out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    tf = 2.0 * M * N * K / ms / 1e9
    gb = (2.0 * (M * K + N * K) + out.element_size() * M * N) / ms / 1e6
    print("%-28s M=%-8d N=%-5d K=%-8d %8.3f ms  %7.1f TFLOP/s (%4.1f%% of %.0f)  %7.1f GB/s" % (name, M, N, K, ms, tf, 100 * tf / PEAK, PEAK, gb), flush=True)
    del A, B, out
