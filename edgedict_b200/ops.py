"""Tensor-level wrappers over the C-ABI: torch supplies device memory and the current stream,
nothing else.  Every function requires CUDA tensors and raises otherwise (no fallback)."""
import torch

from ._lib import lib
from ._lib import check as _check


def _s():
    return torch.cuda.current_stream().cuda_stream


def _p(t):
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("edgedict_b200 ops need CUDA tensors (got a %s tensor); there is no CPU path" % t.device)
    return t.data_ptr()


def _need(t, dtype=None, name="tensor"):
    if not t.is_cuda:
        raise RuntimeError("edgedict_b200 ops need CUDA tensors (%s is on %s); there is no CPU path" % (name, t.device))
    if dtype is not None and t.dtype != dtype:
        raise TypeError("%s must be %s, got %s" % (name, dtype, t.dtype))
    if not t.is_contiguous():
        raise ValueError("%s must be contiguous" % name)
    return t


f32, bf16 = torch.float32, torch.bfloat16


# ---- optional per-kernel instrumentation (bench.py) ----------------------------------------------
class _Prof:
    """When enabled, every wrapped C-ABI call is bracketed by CUDA events on the launching stream
    and counted; bench.py reads `summary()` after a synchronize.  Disabled: zero overhead."""

    def __init__(self):
        self.enabled = False
        self.events = []          # (name, start, end, bytes, flops)
        self.launches = 0

    def reset(self):
        self.events, self.launches = [], 0

    def summary(self, base=None):
        """Per-name totals.  `ms` is the UNION of the call intervals (calls of one name overlap when two layers'
        kernels run on two streams in the wavefront schedule) when a `base` event recorded before the first call
        is given, else their sum; `ms_sum` is always the plain sum."""
        out, spans = {}, {}
        for name, a, b, nbytes, flops in self.events:
            d = out.setdefault(name, dict(ms=0.0, ms_sum=0.0, calls=0, bytes=0.0, flops=0.0))
            d["ms_sum"] += a.elapsed_time(b)
            d["calls"] += 1
            d["bytes"] += nbytes
            d["flops"] += flops
            if base is not None:
                spans.setdefault(name, []).append((base.elapsed_time(a), base.elapsed_time(b)))
        for name, d in out.items():
            if base is None:
                d["ms"] = d["ms_sum"]
                continue
            tot, cur_s, cur_e = 0.0, None, None
            for s0, e0 in sorted(spans[name]):
                if cur_e is None or s0 > cur_e:
                    if cur_e is not None:
                        tot += cur_e - cur_s
                    cur_s, cur_e = s0, e0
                else:
                    cur_e = max(cur_e, e0)
            if cur_e is not None:
                tot += cur_e - cur_s
            d["ms"] = tot
        return out


PROF = _Prof()


def check(rc, what):
    PROF.launches += 1            # every C-ABI call launches at least one of our kernels
    _check(rc, what)


class _timed:
    def __init__(self, name, kernels=1, nbytes=0.0, flops=0.0):
        self.name, self.k, self.nbytes, self.flops = name, kernels, nbytes, flops

    def __enter__(self):
        if PROF.enabled:
            self.a = torch.cuda.Event(enable_timing=True)
            self.a.record()
        return self

    def __exit__(self, *exc):
        PROF.launches += self.k - 1     # extra kernels beyond the one `check` counts
        if PROF.enabled:
            b = torch.cuda.Event(enable_timing=True)
            b.record()
            PROF.events.append((self.name, self.a, b, self.nbytes, self.flops))
        return False


def cast_bf16(x, out=None):
    _need(x, f32, "x")
    y = out if out is not None else torch.empty(x.shape, dtype=bf16, device=x.device)
    n = x.numel()
    if n:
        check(lib().eb_cast_bf16(_p(x), _p(y), n, _s()), "eb_cast_bf16")
    return y


def gemm_f32(A, sam, sak, B, sbk, sbn, M, N, K, bias=None, out=None, beta=0.0, alpha=1.0):
    """C[M,N] = alpha*A'B' + beta*C + bias; strided views of fp32 storage (see eb_gemm_f32)."""
    if out is None:
        out = torch.empty(M, N, dtype=f32, device=A.device)
    with _timed("gemm_f32", 1, 0.0, 2.0 * M * N * K):
        check(lib().eb_gemm_f32(_p(A), sam, sak, _p(B), sbk, sbn, _p(out), N, _p(bias), M, N, K, alpha, beta, _s()),
              "eb_gemm_f32")
    return out


GEMM_CORESIDENT = 1      # include/edgedict_b200.h EB_GEMM_CORESIDENT


def gemm_pair_mode(mode):
    """-1 automatic / 0 never / 1 whenever legal: cta_group::2 tiles in the bf16-output GEMMs (eb_gemm_pair_mode);
    returns the previous mode."""
    return int(lib().eb_gemm_pair_mode(int(mode)))


def gemm_bf16(A, a_mn, B, b_mn, M, N, K, bias=None, out=None, out_bf16=False, accumulate=False, tag=None, flags=0):
    if out is None:
        out = torch.empty(M, N, dtype=bf16 if out_bf16 else f32, device=A.device)
    name = tag or ("gemm_bf16_%s%s" % ("t" if a_mn else "n", "n" if b_mn else "t"))
    with _timed(name, 1, 2.0 * (M * K + N * K) + out.element_size() * M * N, 2.0 * M * N * K):
        check(lib().eb_gemm_bf16_ex(_p(A), int(a_mn), _p(B), int(b_mn), _p(out), int(out.dtype == bf16), _p(bias),
                                    int(accumulate), M, N, K, int(flags), _s()), "eb_gemm_bf16")
    return out


# ---- the three GEMM shapes of a Linear layer, dispatched on precision ---------------------------
def mm_nt(x, w, bias=None, precision="fp32", x16=None, w16=None, out_bf16=False):
    """y[M,N] = x[M,K] @ w[N,K]^T + bias.  w may be a column-slice view (fp32 mode uses strides)."""
    M, K = x.shape
    N = w.shape[0]
    if precision == "bf16":
        x16 = x16 if x16 is not None else (x if x.dtype == bf16 else cast_bf16(x))
        w16 = w16 if w16 is not None else cast_bf16(w.contiguous())
        return gemm_bf16(x16, 0, w16, 0, M, N, K, bias=bias, out_bf16=out_bf16)
    _need(x, f32, "x")
    return gemm_f32(x, K, 1, w, w.stride(1), w.stride(0), M, N, K, bias=bias)


def mm_nn(dy, w, precision="fp32", dy16=None, w16=None, out_bf16=False):
    """dx[M,K] = dy[M,N] @ w[N,K]."""
    M, N = dy.shape
    K = w.shape[1]
    if precision == "bf16":
        dy16 = dy16 if dy16 is not None else (dy if dy.dtype == bf16 else cast_bf16(dy))
        w16 = w16 if w16 is not None else cast_bf16(w.contiguous())
        return gemm_bf16(dy16, 0, w16, 1, M, K, N, out_bf16=out_bf16)
    _need(dy, f32, "dy")
    return gemm_f32(dy, N, 1, w, w.stride(0), w.stride(1), M, K, N)


def mm_tn(dy, x, precision="fp32", dy16=None, x16=None, out=None, accumulate=False):
    """dw[N,K] (+)= dy[M,N]^T @ x[M,K]  (contraction over the rows)."""
    M, N = dy.shape
    K = x.shape[1]
    if precision == "bf16":
        dy16 = dy16 if dy16 is not None else (dy if dy.dtype == bf16 else cast_bf16(dy))
        x16 = x16 if x16 is not None else (x if x.dtype == bf16 else cast_bf16(x))
        return gemm_bf16(dy16, 1, x16, 1, N, K, M, out=out, accumulate=accumulate)
    _need(dy, f32, "dy")
    _need(x, f32, "x")
    return gemm_f32(dy, 1, N, x, K, 1, N, K, M, out=out, beta=1.0 if accumulate else 0.0)


def colsum(x, out=None):
    rows, N = x.shape
    if out is None:
        out = torch.zeros(N, dtype=f32, device=x.device)
    check(lib().eb_colsum(_p(x), int(x.dtype == bf16), _p(out), rows, N, _s()), "eb_colsum")
    return out


# ---- LayerNorm / TimeReduction / Embedding -------------------------------------------------------
def layernorm_fwd(x, res, gamma, beta, eps=1e-5, want_bf16=False, out=None):
    """out = (y, y16 | None, mean, rstd) preallocated (contiguous slices of larger buffers), or None."""
    H = x.shape[-1]
    rows = x.numel() // H
    if out is not None:
        y, y16, mean, rstd = out
    else:
        y = torch.empty_like(x)
        y16 = torch.empty(x.shape, dtype=bf16, device=x.device) if want_bf16 else None
        mean = torch.empty(rows, dtype=f32, device=x.device)
        rstd = torch.empty(rows, dtype=f32, device=x.device)
    check(lib().eb_layernorm_fwd(_p(x), _p(res), _p(gamma), _p(beta), _p(y), _p(y16), _p(mean), _p(rstd),
                                 rows, H, eps, _s()), "eb_layernorm_fwd")
    return y, y16, mean, rstd


def layernorm_bwd(dy, x, res, gamma, mean, rstd):
    H = x.shape[-1]
    rows = x.numel() // H
    dz = torch.empty_like(x)
    dgamma = torch.zeros(H, dtype=f32, device=x.device)
    dbeta = torch.zeros(H, dtype=f32, device=x.device)
    check(lib().eb_layernorm_bwd(_p(dy), _p(x), _p(res), _p(gamma), _p(mean), _p(rstd), _p(dz), _p(dgamma),
                                 _p(dbeta), rows, H, _s()), "eb_layernorm_bwd")
    return dz, dgamma, dbeta


def time_reduce_fwd(x, want_bf16=False, out=None):
    B, T, H = x.shape
    if out is not None:
        y, y16 = out
    else:
        y = torch.empty(B, (T + 1) // 2, H, dtype=f32, device=x.device)
        y16 = torch.empty(y.shape, dtype=bf16, device=x.device) if want_bf16 else None
    check(lib().eb_time_reduce_fwd(_p(x), _p(y), _p(y16), B, T, H, _s()), "eb_time_reduce_fwd")
    return y, y16


def time_reduce_bwd(dy, T, out=None):
    B, _, H = dy.shape
    dx = out if out is not None else torch.empty(B, T, H, dtype=f32, device=dy.device)
    check(lib().eb_time_reduce_bwd(_p(dy), _p(dx), B, T, H, _s()), "eb_time_reduce_bwd")
    return dx


def embedding_fwd(ids, W, prepend_bos, bos):
    B, U = ids.shape
    E = W.shape[1]
    out = torch.empty(B, U + (1 if prepend_bos else 0), E, dtype=f32, device=W.device)
    if out.numel():
        check(lib().eb_embedding_fwd(_p(ids), int(ids.dtype == torch.int64), _p(W), _p(out), None, B, U, E,
                                     int(prepend_bos), bos, _s()), "eb_embedding_fwd")
    return out


def embedding_bwd(ids, dout, V, prepend_bos, bos, pad):
    B, U = ids.shape
    E = dout.shape[-1]
    dW = torch.zeros(V, E, dtype=f32, device=dout.device)
    if dout.numel():
        check(lib().eb_embedding_bwd(_p(ids), int(ids.dtype == torch.int64), _p(dout), _p(dW), B, U, E,
                                     int(prepend_bos), bos, pad, _s()), "eb_embedding_bwd")
    return dW


# ---- LSTM recurrent part -------------------------------------------------------------------------
_scratch = {}


def _lstm_scratch(B, H, device):
    key = (B, H, device, _s())        # per stream: layers on different streams (predictor / encoder) run concurrently
    t = _scratch.get(key)
    if t is None:
        n = lib().eb_lstm_scratch_bytes(B, H)
        if n == 0:
            raise ValueError("LSTM hidden size %d not supported by the persistent kernel" % H)
        t = torch.zeros(n, dtype=torch.uint8, device=device)
        _scratch[key] = t
    return t


def lstm_seq_fwd(xg, whh, h0, c0, save):
    B, T, H4 = xg.shape
    H = H4 // 4
    dev = xg.device
    y = torch.empty(B, T, H, dtype=f32, device=dev)
    hT = torch.empty(B, H, dtype=f32, device=dev)
    cT = torch.empty(B, H, dtype=f32, device=dev)
    gates = torch.empty(B, T, H4, dtype=f32, device=dev) if save else None
    cseq = torch.empty(B, T, H, dtype=f32, device=dev) if save else None
    with _timed("lstm_seq_fwd", 1, 0.0, 2.0 * B * T * 4 * H * H):
        check(lib().eb_lstm_seq_fwd(_p(xg), _p(whh), _p(h0), _p(c0), _p(y), _p(hT), _p(cT), _p(gates), _p(cseq),
                                    _p(_lstm_scratch(B, H, dev)), B, T, H, _s()), "eb_lstm_seq_fwd")
    return y, hT, cT, gates, cseq


def lstm_seq_bwd(dy, gates, cseq, c0, whh, dhT, dcT):
    """Returns (dgates -- written IN PLACE over `gates` --, dh0, dc0)."""
    B, T, H = dy.shape
    dev = dy.device
    dh0 = torch.empty(B, H, dtype=f32, device=dev)
    dc0 = torch.empty(B, H, dtype=f32, device=dev)
    with _timed("lstm_seq_bwd", 1, 0.0, 2.0 * B * T * 4 * H * H):
        check(lib().eb_lstm_seq_bwd(_p(dy), _p(gates), _p(cseq), _p(c0), _p(whh), _p(dhT), _p(dcT), _p(gates),
                                    _p(dh0), _p(dc0), _p(_lstm_scratch(B, H, dev)), B, T, H, _s()), "eb_lstm_seq_bwd")
    return gates, dh0, dc0


def lstm_tc_supported(B, H):
    return bool(lib().eb_lstm_tc_supported(B, H))


def _lstm_tc_scratch(B, H, device):
    key = ("tc", H, device, _s())     # one exchange buffer + barrier block per stream: kernels of two layers overlap
    t = _scratch.get(key)
    if t is None:
        t = torch.zeros(lib().eb_lstm_tc_scratch_bytes(B, H), dtype=torch.uint8, device=device)
        _scratch[key] = t
    return t


def lstm_tc_fwd(xg, whh16, h0, c0, save, out=None):
    """out = (y, y16, hT, cT, gates | None, cseq | None) preallocated, or None."""
    B, T, H4 = xg.shape
    H = H4 // 4
    dev = xg.device
    if out is not None:
        y, y16, hT, cT, gates, cseq = out
    else:
        y = torch.empty(B, T, H, dtype=f32, device=dev)
        y16 = torch.empty(B, T, H, dtype=bf16, device=dev)
        hT = torch.empty(B, H, dtype=f32, device=dev)
        cT = torch.empty(B, H, dtype=f32, device=dev)
        gates = torch.empty(B, T, H4, dtype=f32, device=dev) if save else None
        cseq = torch.empty(B, T, H, dtype=f32, device=dev) if save else None
    with _timed("lstm_tc_fwd", 1, 0.0, 2.0 * B * T * 4 * H * H):
        check(lib().eb_lstm_tc_fwd(_p(xg), _p(whh16), _p(h0), _p(c0), _p(y), _p(y16), _p(hT), _p(cT), _p(gates),
                                   _p(cseq), _p(_lstm_tc_scratch(B, H, dev)), B, T, H, _s()), "eb_lstm_tc_fwd")
    return y, y16, hT, cT, gates, cseq


def lstm_tc_bwd(dy, gates, cseq, c0, whhT16, dhT, dcT, out=None):
    B, T, H = dy.shape
    dev = dy.device
    dg16 = out if out is not None else torch.empty(B, T, 4 * H, dtype=bf16, device=dev)
    dh0 = torch.empty(B, H, dtype=f32, device=dev)
    dc0 = torch.empty(B, H, dtype=f32, device=dev)
    with _timed("lstm_tc_bwd", 1, 0.0, 2.0 * B * T * 4 * H * H):
        check(lib().eb_lstm_tc_bwd(_p(dy), _p(gates), _p(cseq), _p(c0), _p(whhT16), _p(dhT), _p(dcT), _p(dg16),
                                   _p(dh0), _p(dc0), _p(_lstm_tc_scratch(B, H, dev)), B, T, H, _s()), "eb_lstm_tc_bwd")
    return dg16, dh0, dc0


def lstm_tc_bwd_chunks(dy, gates, cseq, whhT16, lens, B, out):
    """BPTT of one layer over chunk-major buffers (functional._Chunks: rows = B * sum(lens)) in ONE launch; zero initial and
    final-state gradients.  dy [rows,H] fp32, gates [rows,4H], cseq [rows,H], out = dg16 [rows,4H] bf16."""
    import ctypes
    H = dy.shape[1]
    dev = dy.device
    dh0 = torch.empty(B, H, dtype=f32, device=dev)
    dc0 = torch.empty(B, H, dtype=f32, device=dev)
    arr = (ctypes.c_int * len(lens))(*[int(n) for n in lens])
    with _timed("lstm_tc_bwd", 1, 0.0, 2.0 * B * sum(lens) * 4 * H * H):
        check(lib().eb_lstm_tc_bwd_chunks(_p(dy), _p(gates), _p(cseq), None, _p(whhT16), None, None, _p(out), _p(dh0),
                                          _p(dc0), _p(_lstm_tc_scratch(B, H, dev)), B, arr, len(lens), H, _s()),
              "eb_lstm_tc_bwd_chunks")
    return out, dh0, dc0


# ---- cluster / tcgen05 recurrent kernels (csrc/lstm_c4.cu) ------------------------------------------
_c4_ok = {}


def lstm_c4_supported(B, H):
    """True when eb_lstm_c4_* can run this layer: H % 256 == 0, H <= 1024 and all clusters co-resident."""
    v = _c4_ok.get(H)
    if v is None:
        v = _c4_ok[H] = bool(lib().eb_lstm_c4_supported(B, H))
    return v


def _lstm_c4_scratch(B, H, device):
    key = ("c4", H, device, _s())     # one exchange buffer + barrier block per stream: kernels of two layers overlap
    t = _scratch.get(key)
    if t is None:
        t = torch.zeros(lib().eb_lstm_c4_scratch_bytes(B, H), dtype=torch.uint8, device=device)
        _scratch[key] = t
    return t


def lstm_c4_save_buffers(B, T, H, device):
    """(gsave, csave): CTA-private layout of the forward saves (bf16 gates, fp32 cell states)."""
    return (torch.empty(lib().eb_lstm_c4_gsave_bytes(B, T, H), dtype=torch.uint8, device=device),
            torch.empty(lib().eb_lstm_c4_csave_bytes(B, T, H), dtype=torch.uint8, device=device))


C4_BWD = __import__("os").environ.get("EDGEDICT_LSTM_C4_BWD", "0") != "0"   # BPTT through lstm_c4 (else lstm_tc's kernel)


def lstm_c4_fwd(xg, whh16, h0, c0, save, out=None, std_saves=None):
    """out = (y, hprev16, hT, cT) preallocated, or None.  Returns (y, hprev16, hT, cT, gsave | None, csave | None);
    hprev16[:, t] = bf16(h_{t-1}) (frame 0 = h0) is the operand of the dW_hh GEMM.  The saves for backward are
    either the kernels' own CTA-private layout (gsave, csave: consumed by lstm_c4_bwd) or, with
    std_saves = (gates [B,T,4H] fp32, cseq [B,T,H] fp32) (or True to allocate them), the layout of lstm_tc_bwd."""
    B, T, H4 = xg.shape
    H = H4 // 4
    dev = xg.device
    if out is not None:
        y, hprev16, hT, cT = out
    else:
        y = torch.empty(B, T, H, dtype=f32, device=dev)
        hprev16 = torch.empty(B, T, H, dtype=bf16, device=dev) if save else None
        hT = torch.empty(B, H, dtype=f32, device=dev)
        cT = torch.empty(B, H, dtype=f32, device=dev)
    gsave = csave = gstd = cstd = None
    if save and std_saves is not None and std_saves is not False:
        gstd, cstd = std_saves if isinstance(std_saves, tuple) else (torch.empty(B, T, H4, dtype=f32, device=dev),
                                                                     torch.empty(B, T, H, dtype=f32, device=dev))
    elif save:
        gsave, csave = lstm_c4_save_buffers(B, T, H, dev)
    with _timed("lstm_tc_fwd", 1, 0.0, 2.0 * B * T * 4 * H * H):
        check(lib().eb_lstm_c4_fwd(_p(xg), _p(whh16), _p(h0), _p(c0), _p(y), _p(hprev16), _p(hT), _p(cT), _p(gsave),
                                   _p(csave), _p(gstd), _p(cstd), _p(_lstm_c4_scratch(B, H, dev)), B, T, H, _s()),
              "eb_lstm_c4_fwd")
    if gstd is not None:
        return y, hprev16, hT, cT, gstd, cstd
    return y, hprev16, hT, cT, gsave, csave


def lstm_c4_bwd(dy, gsave, csave, c0, whhT16, dhT, dcT, out=None):
    B, T, H = dy.shape
    dev = dy.device
    dg16 = out if out is not None else torch.empty(B, T, 4 * H, dtype=bf16, device=dev)
    dh0 = torch.empty(B, H, dtype=f32, device=dev)
    dc0 = torch.empty(B, H, dtype=f32, device=dev)
    with _timed("lstm_tc_bwd", 1, 0.0, 2.0 * B * T * 4 * H * H):
        check(lib().eb_lstm_c4_bwd(_p(dy), _p(gsave), _p(csave), _p(c0), _p(whhT16), _p(dhT), _p(dcT), _p(dg16),
                                   _p(dh0), _p(dc0), _p(_lstm_c4_scratch(B, H, dev)), B, T, H, _s()), "eb_lstm_c4_bwd")
    return dg16, dh0, dc0


def transpose_to_bf16(x):
    """x [rows, cols] (fp32 or bf16) -> bf16 [cols, rows]."""
    rows, cols = x.shape
    y = torch.empty(cols, rows, dtype=bf16, device=x.device)
    check(lib().eb_transpose_to_bf16(_p(x), int(x.dtype == bf16), _p(y), rows, cols, _s()), "eb_transpose_to_bf16")
    return y


# ---- joint + loss ----------------------------------------------------------------------------------
def joint_hidden_fwd(ep, dp, want_bf16):
    B, T, J = ep.shape
    U = dp.shape[1]
    hid = torch.empty(B, T, U, J, dtype=bf16 if want_bf16 else f32, device=ep.device)
    check(lib().eb_joint_hidden_fwd(_p(ep), _p(dp), _p(hid), int(want_bf16), B, T, U, J, _s()), "eb_joint_hidden_fwd")
    return hid


def joint_hidden_bwd(dhid, hid):
    """dhid is overwritten with d(pre-activation); returns (dep [B,T,J], ddp [B,U,J])."""
    B, T, U, J = hid.shape
    dep = torch.empty(B, T, J, dtype=f32, device=hid.device)
    ddp = torch.empty(B, U, J, dtype=f32, device=hid.device)
    check(lib().eb_joint_hidden_bwd(_p(dhid), _p(hid), int(hid.dtype == bf16), _p(dep), _p(ddp), B, T, U, J, _s()),
          "eb_joint_hidden_bwd")
    return dep, ddp


def gemm_bf16_dtanh(A16, B16, b_mn, hid16, M, N, K):
    """bf16 [M,N] = (A16 [M,K] @ B) * (1 - hid16^2)  (B16: [N,K] if not b_mn else [K,N])."""
    out = torch.empty(M, N, dtype=bf16, device=A16.device)
    with _timed("gemm_bf16_n%s" % ("n" if b_mn else "t"), 1, 2.0 * (M * K + N * K) + 4.0 * M * N, 2.0 * M * N * K):
        check(lib().eb_gemm_bf16_dtanh(_p(A16), 0, _p(B16), int(b_mn), _p(out), _p(hid16), M, N, K, _s()),
              "eb_gemm_bf16_dtanh")
    return out


def joint_dpre_reduce(dpre16):
    """dpre16 [B,T,U,J] bf16 -> (dep [B,T,J], ddp [B,U,J]) fp32."""
    B, T, U, J = dpre16.shape
    dep = torch.empty(B, T, J, dtype=f32, device=dpre16.device)
    ddp = torch.empty(B, U, J, dtype=f32, device=dpre16.device)
    check(lib().eb_joint_dpre_reduce(_p(dpre16), _p(dep), _p(ddp), B, T, U, J, _s()), "eb_joint_dpre_reduce")
    PROF.launches += 1
    return dep, ddp


def rnnt_workspace(B, T, U, dtype, device):
    n = lib().eb_rnnt_workspace_bytes(B, T, U, 8 if dtype == torch.float64 else 4)
    return torch.empty(n, dtype=torch.uint8, device=device)


def rnnt_loss_fwd(logits, labels, xlen, ylen, blank, need_beta=True):
    B, T, U, V = logits.shape
    ds = 8 if logits.dtype == torch.float64 else 4
    ws = rnnt_workspace(B, T, U, logits.dtype, logits.device)
    costs = torch.empty(B, dtype=logits.dtype, device=logits.device)
    with _timed("rnnt_loss_fwd", 3, float(ds) * B * T * U * V, 0.0):
        check(lib().eb_rnnt_loss_fwd(_p(logits), _p(labels), _p(xlen), _p(ylen), B, T, U, V, blank, ds, _p(ws),
                                     _p(costs), int(need_beta), _s()), "eb_rnnt_loss_fwd")
    return costs, ws


def rnnt_loss_bwd(logits, labels, xlen, ylen, blank, ws, gscale, host_scale, out=None, out_bf16=False):
    B, T, U, V = logits.shape
    ds = 8 if logits.dtype == torch.float64 else 4
    if out is None:
        out = torch.empty(logits.shape, dtype=bf16 if out_bf16 else logits.dtype, device=logits.device)
    per_batch = int(gscale is not None and gscale.numel() > 1)
    with _timed("rnnt_loss_bwd", 1, float(ds + out.element_size()) * B * T * U * V, 0.0):
        check(lib().eb_rnnt_loss_bwd(_p(logits), _p(out), int(out.dtype == bf16), _p(labels), _p(xlen), _p(ylen), B, T,
                                     U, V, blank, ds, _p(ws), _p(gscale), per_batch, float(host_scale), _s()),
              "eb_rnnt_loss_bwd")
    return out


def joint_logits_lse(hid16, w2_16, b2, labels, xlen, ylen, B, T, U, blank):
    """bf16 logits [B,T,U,V] + loss workspace with denom / log p(blank) / log p(label) filled in."""
    V, J = w2_16.shape
    logits16 = torch.empty(B, T, U, V, dtype=bf16, device=hid16.device)
    ws = rnnt_workspace(B, T, U, f32, hid16.device)
    n = B * T * U
    wsf = ws.view(f32)
    N = float(n) * V
    with _timed("joint_logits_lse", 1, 2.0 * n * J + 2.0 * N, 2.0 * N * J):
        check(lib().eb_joint_logits_lse(_p(hid16), _p(w2_16), _p(b2), _p(logits16), _p(labels), _p(xlen), _p(ylen),
                                        _p(wsf[0:n]), _p(wsf[n:2 * n]), _p(wsf[2 * n:3 * n]), B, T, U, V, J, blank,
                                        _s()), "eb_joint_logits_lse")
    return logits16, ws


def rnnt_lattice(xlen, ylen, B, T, U, ws, need_beta=True):
    costs = torch.empty(B, dtype=f32, device=ws.device)
    with _timed("rnnt_loss_fwd", 2, 0.0, 0.0):
        check(lib().eb_rnnt_loss_lattice(_p(xlen), _p(ylen), B, T, U, _p(ws), _p(costs), int(need_beta), _s()),
              "eb_rnnt_loss_lattice")
    return costs


def rnnt_loss_bwd_bf16(logits16, labels, xlen, ylen, blank, ws, gscale, host_scale):
    """In place: logits16 becomes d loss / d logits (bf16)."""
    B, T, U, V = logits16.shape
    per_batch = int(gscale is not None and gscale.numel() > 1)
    with _timed("rnnt_loss_bwd", 1, 4.0 * B * T * U * V, 0.0):
        check(lib().eb_rnnt_loss_bwd_bf16(_p(logits16), _p(logits16), _p(labels), _p(xlen), _p(ylen), B, T, U, V, blank,
                                          _p(ws), _p(gscale), per_batch, float(host_scale), _s()), "eb_rnnt_loss_bwd_bf16")
    return logits16


def adam_step(p, g, m, v, lr, beta1, beta2, eps, weight_decay, step, grad_scale=1.0):
    check(lib().eb_adam_step(_p(p), _p(g), _p(m), _p(v), p.numel(), lr, beta1, beta2, eps, weight_decay, step,
                             grad_scale, _s()), "eb_adam_step")


def adam_step_ex(p, g, m, v, lr, beta1, beta2, eps, weight_decay, step, grad_scale=1.0, sumsq=None, max_norm=0.0,
                 adamw=False):
    check(lib().eb_adam_step_ex(_p(p), _p(g), _p(m), _p(v), p.numel(), lr, beta1, beta2, eps, weight_decay, step,
                                grad_scale, _p(sumsq), float(max_norm or 0.0), int(adamw), _s()), "eb_adam_step_ex")


def sumsq(x, out):
    check(lib().eb_sumsq(_p(x), x.numel(), _p(out), _s()), "eb_sumsq")


# ---- log-mel front end ------------------------------------------------------------
def logmel_frontend(x, basis, fbT, n_fft, hop, n_mels, n_stack, preemph, take_log=True, pad_to_divisible=True):
    """x [B, L] fp32 waveform -> [B, T, n_mels*n_stack] log-mel features (see csrc/frontend.cu).
    basis [n_fft, 2*nbins] = window * (cos | -sin); fbT [nbins, n_mels]."""
    _need(x, f32, "x")
    B, L = x.shape
    nb = n_fft // 2 + 1
    pad = n_fft // 2
    R = -(-(L + 2 * pad) // hop)                       # frame slots per utterance (>= the 1 + L//hop real frames)
    Lp = R * hop
    F = 1 + L // hop
    seq_len = -(-L // hop)
    dev = x.device
    xp = torch.zeros(B * Lp + n_fft, dtype=f32, device=dev)      # tail: the last slots' windows stay in bounds
    check(lib().eb_fe_preemph_pad(_p(x), _p(xp), B, L, Lp, pad, float(preemph or 0.0), int(preemph is not None), _s()),
          "eb_fe_preemph_pad")
    rows = B * R
    spec = gemm_f32(xp, hop, 1, basis, 2 * nb, 1, rows, 2 * nb, n_fft)
    power = torch.empty(rows, nb, dtype=f32, device=dev)
    check(lib().eb_fe_power(_p(spec), _p(power), rows, nb, _s()), "eb_fe_power")
    mel = gemm_f32(power, nb, 1, fbT, n_mels, 1, rows, n_mels, nb)
    Fs = F if pad_to_divisible else F - F % n_stack
    T = -(-Fs // n_stack)
    out = torch.empty(B, T, n_mels * n_stack, dtype=f32, device=dev)
    check(lib().eb_fe_log_stack(_p(mel), _p(out), B, R, Fs, seq_len, n_mels, n_stack, T, int(take_log), _s()),
          "eb_fe_log_stack")
    return out


def fe_mask(x, spans, axis, fill=0.0):
    """SpecAugment masking in place: x [B, D1, D2] fp32, spans int32 [B, nmask, 2] along axis 1 or 2."""
    _need(x, f32, "x")
    B, D1, D2 = x.shape
    check(lib().eb_fe_mask(_p(x), _p(spans), B, D1, D2, spans.shape[1], axis, float(fill), _s()), "eb_fe_mask")
    return x
