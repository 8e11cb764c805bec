"""cta_group::2 GEMM tiles (gemm_tc.cu, PAIR_) against the one-CTA tiles and torch: the three bf16-output products of the
joint -- logits + softmax statistics, d-hidden with tanh', and a plain nt product with bias -- at shapes with odd row-block
counts and a half-empty last column tile, then timed at the E6D2 joint's sizes.  Usage: gpu_pair_check.py [check|time]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from edgedict_b200 import ops

bf16, f32 = torch.bfloat16, torch.float32
dev = "cuda"


def ev_time(fn, reps=5):
    fn(); torch.cuda.synchronize()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    ts = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]


def both(fn):
    ops.gemm_pair_mode(0)
    ref = fn()
    ops.gemm_pair_mode(1)
    got = fn()
    ops.gemm_pair_mode(-1)
    torch.cuda.synchronize()
    return ref, got


def check():
    torch.manual_seed(0)
    worst = 0.0
    for (M, N, K) in [(256 * 5, 512, 320), (128 * 7 + 40, 640, 192), (128 * 301, 1024, 640), (3000, 256, 64)]:
        A = (torch.randn(M, K, device=dev) * 0.5).to(bf16)
        W = (torch.randn(N, K, device=dev) * 0.1).to(bf16)
        bias = torch.randn(N, device=dev)
        ref, got = both(lambda: ops.gemm_bf16(A, 0, W, 0, M, N, K, bias=bias, out_bf16=True))
        want = (A.float() @ W.float().t() + bias)
        e = float((got.float() - want).abs().max() / want.abs().max())
        same = bool((ref == got).all())
        print("nt  M=%d N=%d K=%d: pair vs torch %.2e, bit-identical to one-CTA tiles: %s" % (M, N, K, e, same))
        assert e < 1e-2 and same
        # accumulate into a bf16 output
        base = torch.randn(M, N, device=dev).to(bf16)
        ref, got = both(lambda: ops.gemm_bf16(A, 0, W, 0, M, N, K, out=base.clone(), accumulate=True))
        assert bool((ref == got).all()), "accumulate"
        # nn with tanh' (B MN-major [K, N])
        Wn = (torch.randn(K, N, device=dev) * 0.1).to(bf16)
        hid = torch.tanh(torch.randn(M, N, device=dev)).to(bf16)
        ref, got = both(lambda: ops.gemm_bf16_dtanh(A, Wn, 1, hid, M, N, K))
        want = (A.float() @ Wn.float()) * (1 - hid.float() ** 2)
        e = float((got.float() - want).abs().max() / want.abs().max())
        same = bool((ref == got).all())
        print("nn' M=%d N=%d K=%d: pair vs torch %.2e, bit-identical: %s" % (M, N, K, e, same))
        assert e < 1e-2 and same
        worst = max(worst, e)
    # split-K weight gradient, both operands MN-major (dW2 = dlogits^T hid): fp32 atomics, so not bit-identical
    for (M, N, K) in [(1024, 640, 64 * 700 + 24), (512, 256, 64 * 130), (320, 384, 64 * 97)]:
        dy = (torch.randn(K, M, device=dev) * 0.1).to(bf16)
        x = torch.randn(K, N, device=dev).to(bf16)
        ref, got = both(lambda: ops.gemm_bf16(dy, 1, x, 1, M, N, K))
        want = dy.float().t() @ x.float()
        e0 = float((ref - want).abs().max() / want.abs().max())
        e1 = float((got - want).abs().max() / want.abs().max())
        print("tn  M=%d N=%d K=%d: one-CTA vs torch %.2e, pair vs torch %.2e" % (M, N, K, e0, e1))
        assert e1 < 2e-5
        acc0 = torch.randn(M, N, device=dev)
        ref, got = both(lambda: ops.gemm_bf16(dy, 1, x, 1, M, N, K, out=acc0.clone(), accumulate=True))
        assert float((got - (want + acc0)).abs().max() / want.abs().max()) < 2e-5
    # logits + LSE: ragged lattice, V = 512 and 1024
    for (B, T, U, V, J) in [(3, 37, 9, 512, 128), (2, 150, 33, 1024, 640), (5, 41, 7, 256, 64)]:
        hid = torch.tanh(torch.randn(B, T, U, J, device=dev)).to(bf16)
        w2 = (torch.randn(V, J, device=dev) * 0.2).to(bf16)
        b2 = torch.randn(V, device=dev)
        labels = torch.randint(1, V, (B, U - 1), dtype=torch.int32, device=dev)
        xlen = torch.randint(T // 2, T + 1, (B,), dtype=torch.int32, device=dev); xlen[0] = T
        ylen = torch.randint(1, U, (B,), dtype=torch.int32, device=dev); ylen[0] = U - 1
        (l0, w0), (l1, w1) = both(lambda: ops.joint_logits_lse(hid, w2, b2, labels, xlen, ylen, B, T, U, 0))
        n = B * T * U
        same = bool((l0 == l1).all())
        # statistics only defined on valid cells
        t_ok = torch.arange(T, device=dev)[None, :, None] < xlen[:, None, None]
        u_ok = torch.arange(U, device=dev)[None, None, :] <= ylen[:, None, None]
        ok = (t_ok & u_ok).reshape(-1)
        s0, s1 = w0.view(f32)[:3 * n].view(3, n)[:, ok], w1.view(f32)[:3 * n].view(3, n)[:, ok]
        ds = float((s0 - s1).abs().max())
        logits = hid.float().view(n, J) @ w2.float().t() + b2
        den = -torch.logsumexp(logits, dim=1)
        e = float((w1.view(f32)[:n][ok] - den[ok]).abs().max())
        print("lse B=%d T=%d U=%d V=%d J=%d: logits bit-identical %s, statistics max |pair - one-CTA| %.2e, denom vs torch %.2e"
              % (B, T, U, V, J, same, ds, e))
        assert same and ds < 1e-5 and e < 2e-3
    print("pair check ok")


def time_(once=False):
    B, T, U, V, J = 32, 250, 129, 1024, 640
    M = B * T * U
    hid = torch.tanh(torch.randn(B, T, U, J, device=dev)).to(bf16)
    w2 = (torch.randn(V, J, device=dev) * 0.05).to(bf16)
    b2 = torch.randn(V, device=dev)
    labels = torch.randint(1, V, (B, U - 1), dtype=torch.int32, device=dev)
    xlen = torch.full((B,), T, dtype=torch.int32, device=dev)
    ylen = torch.full((B,), U - 1, dtype=torch.int32, device=dev)
    dlog = (torch.randn(M, V, device=dev) * 0.01).to(bf16)
    if once:          # under ncu: one launch per kernel and mode (one-CTA lse, d-hidden; pair lse, d-hidden)
        for mode in (0, 1):
            ops.gemm_pair_mode(mode)
            ops.joint_logits_lse(hid, w2, b2, labels, xlen, ylen, B, T, U, 0)
            ops.gemm_bf16_dtanh(dlog, w2, 1, hid.view(M, J), M, J, V)
        torch.cuda.synchronize()
        return
    for mode in (0, 1):
        ops.gemm_pair_mode(mode)
        t1 = ev_time(lambda: ops.joint_logits_lse(hid, w2, b2, labels, xlen, ylen, B, T, U, 0))
        t2 = ev_time(lambda: ops.gemm_bf16_dtanh(dlog, w2, 1, hid.view(M, J), M, J, V))
        t3 = ev_time(lambda: ops.gemm_bf16(dlog, 1, hid.view(M, J), 1, V, J, M))
        fl = 2.0 * M * V * J
        print("pair_mode %d: logits+LSE %.3f ms (%.0f TFLOP/s)   d-hidden(tanh') %.3f ms (%.0f TFLOP/s)   dW2 %.3f ms (%.0f TFLOP/s)"
              % (mode, t1, fl / t1 / 1e9, t2, fl / t2 / 1e9, t3, fl / t3 / 1e9))
    ops.gemm_pair_mode(-1)


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "check"
    if what in ("check", "all"):
        check()
    if what in ("time", "all"):
        time_()
    if what == "once":
        time_(once=True)
