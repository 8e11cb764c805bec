"""Autograd glue: each Function's forward/backward is a short sequence of C-ABI calls
(edgedict_b200/ops.py).  torch.autograd only stitches them together -- it never differentiates
through a torch op on the hot path.

Reference semantics per Function are cited next to each class (paths under /root/reference).
"""
import torch

from . import ops

import os

f32, bf16 = torch.float32, torch.bfloat16
# bf16 mode: the joint's output GEMM also emits the softmax statistics of the loss from its fp32 TMEM accumulators
# and writes bf16 logits; the 8 GB denominator pass disappears and the gradient pass runs in place on 4 GB
# (EDGEDICT_FUSE_LSE=0 selects the unfused fp32-logits path).  Measured at E6D2: joint GEMM + loss 4.8 ms against
# 7.2 ms unfused, once the epilogue used ex2.approx.ftz, unguarded full chunks and eight epilogue warps.
FUSE_JOINT_LSE = os.environ.get("EDGEDICT_FUSE_LSE", "1") != "0"


def _c(t):
    return t if t.is_contiguous() else t.contiguous()


class Linear(torch.autograd.Function):
    """nn.Linear: y = x W^T + b  (rnnt/models.py:129,148,163-167)."""

    @staticmethod
    def forward(ctx, x, w, b, precision):
        shp = x.shape
        x2 = _c(x).view(-1, shp[-1])
        x16 = ops.cast_bf16(x2) if precision == "bf16" else None
        y = ops.mm_nt(x2, w, b, precision, x16=x16)
        ctx.save_for_backward(x2 if x16 is None else x16, w)
        ctx.precision, ctx.has_b, ctx.shp = precision, b is not None, shp
        return y.view(*shp[:-1], w.shape[0])

    @staticmethod
    def backward(ctx, dy):
        xs, w = ctx.saved_tensors
        p = ctx.precision
        dy2 = _c(dy).view(-1, dy.shape[-1])
        dy16 = ops.cast_bf16(dy2) if p == "bf16" else None
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = ops.mm_nn(dy2, w, p, dy16=dy16).view(ctx.shp)
        if ctx.needs_input_grad[1]:
            dw = ops.mm_tn(dy2, xs, p, dy16=dy16, x16=xs if p == "bf16" else None)
        if ctx.has_b and ctx.needs_input_grad[2]:
            db = ops.colsum(dy2)
        return dx, dw, db, None


class LayerNormRes(torch.autograd.Function):
    """y = LayerNorm(x + res) (res optional): nn.LayerNorm + the residual add of
    ResLayerNormLSTM.forward (rnnt/models.py:47,66-70,124)."""

    @staticmethod
    def forward(ctx, x, res, gamma, beta, eps):
        x = _c(x)
        res = _c(res) if res is not None else None
        y, _, mean, rstd = ops.layernorm_fwd(x, res, gamma, beta, eps)
        ctx.save_for_backward(x, res, gamma, mean, rstd)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, res, gamma, mean, rstd = ctx.saved_tensors
        dz, dgamma, dbeta = ops.layernorm_bwd(_c(dy), x, res, gamma, mean, rstd)
        return dz, (dz if res is not None else None), dgamma, dbeta, None


class TimeReduce(torch.autograd.Function):
    """TimeReduction(2): zero-pad to even length, mean of frame pairs (rnnt/models.py:21-29)."""

    @staticmethod
    def forward(ctx, x):
        ctx.T = x.shape[1]
        y, _ = ops.time_reduce_fwd(_c(x))
        return y

    @staticmethod
    def backward(ctx, dy):
        return ops.time_reduce_bwd(_c(dy), ctx.T)


class Embedding(torch.autograd.Function):
    """nn.Embedding(padding_idx=PAD) with the optional BOS prepend of Decoder.forward
    (rnnt/models.py:150-153)."""

    @staticmethod
    def forward(ctx, ids, w, prepend_bos, bos, pad):
        ids = _c(ids)
        if ids.dtype not in (torch.int32, torch.int64):
            ids = ids.long()
        ctx.save_for_backward(ids)
        ctx.meta = (w.shape[0], prepend_bos, bos, pad)
        return ops.embedding_fwd(ids, w, prepend_bos, bos)

    @staticmethod
    def backward(ctx, dout):
        (ids,) = ctx.saved_tensors
        V, prepend_bos, bos, pad = ctx.meta
        return None, ops.embedding_bwd(ids, _c(dout), V, prepend_bos, bos, pad), None, None, None


class LSTMLayer(torch.autograd.Function):
    """One unidirectional batch_first nn.LSTM layer (rnnt/models.py:45-46,64-65,145-147):
    bulk input GEMM + persistent recurrent kernel; backward = BPTT kernel + three bulk GEMMs."""

    @staticmethod
    def forward(ctx, x, h0, c0, w_ih, w_hh, b_ih, b_hh, precision):
        B, T, I = x.shape
        H = w_hh.shape[1]
        x2 = _c(x).view(B * T, I)
        x16 = ops.cast_bf16(x2) if precision == "bf16" else None
        bias = b_ih + b_hh
        xg = ops.mm_nt(x2, w_ih, bias, precision, x16=x16).view(B, T, 4 * H)
        h0c = _c(h0) if h0 is not None else None
        c0c = _c(c0) if c0 is not None else None
        need = any(ctx.needs_input_grad)     # (grad mode is off inside Function.forward)
        tc = precision == "bf16" and ops.lstm_tc_supported(B, H)
        c4 = precision == "bf16" and ops.lstm_c4_supported(B, H)
        if c4:
            # cluster / tcgen05 kernels: saves in their CTA-private layout, y16 IS the h_{t-1}-shifted copy
            y, y16, hT, cT, gates, cseq = ops.lstm_c4_fwd(xg, ops.cast_bf16(_c(w_hh)), h0c, c0c, need,
                                                          std_saves=None if ops.C4_BWD else True)
        elif tc:
            y, y16, hT, cT, gates, cseq = ops.lstm_tc_fwd(xg, ops.cast_bf16(_c(w_hh)), h0c, c0c, need)
        else:
            y, hT, cT, gates, cseq = ops.lstm_seq_fwd(xg, _c(w_hh), h0c, c0c, need)
            y16 = None
        if need:
            ctx.save_for_backward(x2 if x16 is None else x16, h0c, c0c, w_ih, w_hh, y16 if (tc or c4) else y, gates, cseq)
            ctx.precision, ctx.dims, ctx.tc, ctx.c4, ctx.c4_bwd = precision, (B, T, I, H), tc, c4, c4 and ops.C4_BWD
        return y, hT, cT

    @staticmethod
    def backward(ctx, dy, dhT, dcT):
        xs, h0, c0, w_ih, w_hh, y, gates, cseq = ctx.saved_tensors
        B, T, I, H = ctx.dims
        p = ctx.precision
        dy = _c(dy) if dy is not None else torch.zeros(B, T, H, dtype=f32, device=y.device)
        dhT = _c(dhT) if dhT is not None else None
        dcT = _c(dcT) if dcT is not None else None
        if getattr(ctx, "consumed", False):
            raise RuntimeError("LSTMLayer.backward ran twice on the same graph: the saved gates are overwritten in "
                               "place by the first pass (retain_graph is not supported by this node)")
        ctx.consumed = True
        if ctx.c4_bwd:
            dg16, dh0, dc0 = ops.lstm_c4_bwd(dy, gates, cseq, c0, ops.transpose_to_bf16(_c(w_hh)), dhT, dcT)
            dg2 = dg16.view(B * T, 4 * H)
        elif ctx.tc or ctx.c4:
            dg16, dh0, dc0 = ops.lstm_tc_bwd(dy, gates, cseq, c0, ops.transpose_to_bf16(_c(w_hh)), dhT, dcT)
            dg2 = dg16.view(B * T, 4 * H)
        else:
            dg, dh0, dc0 = ops.lstm_seq_bwd(dy, gates, cseq, c0, _c(w_hh), dhT, dcT)
            dg2 = dg.view(B * T, 4 * H)
            dg16 = ops.cast_bf16(dg2) if p == "bf16" else None
        dg16 = dg16.view(B * T, 4 * H) if dg16 is not None else None
        if ctx.c4:
            hp2 = y.view(B * T, H)                           # the kernel already wrote h_{t-1} for every step
        else:
            # h_{t-1} for every step: y shifted right by one frame, h0 (or zeros) in front
            hprev = torch.empty_like(y)
            hprev[:, 1:] = y[:, :-1]
            if h0 is not None:
                hprev[:, 0] = h0
            else:
                hprev[:, 0].zero_()
            hp2 = hprev.view(B * T, H)
        dx = ops.mm_nn(dg2, w_ih, p, dy16=dg16).view(B, T, I) if ctx.needs_input_grad[0] else None
        dw_ih = ops.mm_tn(dg2, xs, p, dy16=dg16, x16=xs if p == "bf16" else None)
        dw_hh = ops.mm_tn(dg2, hp2, p, dy16=dg16, x16=hp2 if hp2.dtype == bf16 else None)
        db = ops.colsum(dg2)
        return (dx, dh0 if ctx.needs_input_grad[1] else None, dc0 if ctx.needs_input_grad[2] else None,
                dw_ih, dw_hh, db, db.clone(), None)


# ---- layer-wavefront schedule of the LSTM stack ------------------------------------------------------------
# The recurrence of one layer is a chain of T grid-synchronous steps that is bound by exchange/barrier LATENCY
# (DESIGN.md section 7): one layer's persistent kernel uses ~1 of 4 issue slots of 128 SMs.  Layer l+1 at frame t
# only needs layer l up to frame t, so the time axis is cut into chunks and layer l+1 runs chunk c on a second
# stream while layer l runs chunk c+1: two persistent kernels (compiled for <= 128 registers, ~100 KB of shared
# memory each) are co-resident on every SM and hide each other's latency.  The per-chunk input GEMM uses the
# co-resident tile configuration (EB_GEMM_CORESIDENT) so that it fits next to the other layer's recurrent CTA.
WAVEFRONT_CHUNKS = int(os.environ.get("EDGEDICT_WAVEFRONT_CHUNKS", "6"))     # 0 disables; 4/6/8 measured: 54.43/54.09/54.40 ms
_wave_streams = {}


def _side_streams(device):
    st = _wave_streams.get(device)
    if st is None:
        st = _wave_streams[device] = (torch.cuda.Stream(device), torch.cuda.Stream(device))
    return st


_aux_streams = {}


def _wave_aux_streams(device):
    """Four more streams of the forward wavefront: input-projection GEMMs (two) and LayerNorm / TimeReduction (two)."""
    st = _aux_streams.get(device)
    if st is None:
        st = _aux_streams[device] = tuple(torch.cuda.Stream(device) for _ in range(4))
    return st


def wavefront_plan(T, reductions, max_chunks=None):
    """Chunk lengths per layer for the wavefront schedule, or None when T is too short to cut.
    Chunk boundaries are multiples of 2^(number of time reductions) so that TimeReduction pairs never straddle
    a boundary; only the last chunk may be ragged (its zero padding is the reference's padding of the last
    frame, rnnt/models.py:24-27).  Returns a list (len L+1) of per-chunk lengths on each layer's time axis."""
    max_chunks = WAVEFRONT_CHUNKS if max_chunks is None else max_chunks
    if max_chunks < 2:
        return None
    gran = 1 << sum(1 for r in reductions if r)
    step = -(-T // (max_chunks * gran)) * gran
    if step < 16 * gran:                     # too short to amortise the per-chunk launches
        step = 16 * gran
    lens = [min(step, T - t0) for t0 in range(0, T, step)]
    if len(lens) < 2:
        return None
    plan = [lens]
    for r in reductions:
        lens = [(n + 1) // 2 for n in lens] if r else lens
        plan.append(lens)
    return plan


class _Chunks:
    """Chunk-major storage: the [B, Tc, D] blocks of all chunks back to back in one [B*T, D] buffer.  Row-wise
    kernels (LayerNorm, casts, the bulk GEMMs of the backward pass) run on the flat buffer, time-ordered ones
    (the recurrence, TimeReduction) on one block."""

    def __init__(self, B, lens):
        self.B, self.lens = B, lens
        self.off = [0]
        for n in lens:
            self.off.append(self.off[-1] + n)
        self.rows = B * self.off[-1]

    def new(self, D, dtype, device):
        return torch.empty(self.rows, D, dtype=dtype, device=device) if D else torch.empty(self.rows, dtype=dtype, device=device)

    def blk(self, buf, c):
        a, b = self.B * self.off[c], self.B * self.off[c + 1]
        v = buf[a:b]
        return v.view(self.B, self.lens[c], buf.shape[1]) if buf.dim() == 2 else v

    def scatter(self, x):
        """[B, T, D] -> chunk-major flat buffer."""
        out = self.new(x.shape[2], x.dtype, x.device)
        for c in range(len(self.lens)):
            self.blk(out, c).copy_(x[:, self.off[c]:self.off[c + 1]])
        return out

    def gather(self, buf):
        """chunk-major flat buffer -> [B, T, D]."""
        out = torch.empty(self.B, self.off[-1], buf.shape[1], dtype=buf.dtype, device=buf.device)
        for c in range(len(self.lens)):
            out[:, self.off[c]:self.off[c + 1]].copy_(self.blk(buf, c))
        return out


BPTT_ONE_LAUNCH = __import__("os").environ.get("EDGEDICT_BPTT_ONE_LAUNCH", "1") != "0"   # one BPTT launch per layer (else per chunk)


class LSTMStack(torch.autograd.Function):
    """ResLayerNormLSTM.forward (rnnt/models.py:57-75) for all layers at once, bf16 tensor-core mode, zero
    initial state: per layer nn.LSTM -> LayerNorm(y + x) (no residual for layer 0) -> optional TimeReduction,
    executed as a layer wavefront over time chunks on two streams (see above).  Numerically identical to the
    layer-by-layer Functions (same kernels, same per-row / per-step arithmetic); the backward pass reuses their
    kernels on the chunk-major buffers, the BPTT kernel once per chunk with the (dh, dc) carry.
    args: x [B,T,I] fp32, cfg = (reductions, eps, plan), then per layer w_ih, w_hh, b_ih, b_hh, ln_w, ln_b.
    returns y [B,T',H], h_T [L,B,H], c_T [L,B,H] (final states: not differentiated through)."""

    collect = None      # tests: set to a list to receive every layer's output [B,T_l,H] (parity of the per-layer activations)

    @staticmethod
    def forward(ctx, x, cfg, *params):
        reductions, eps, plan = cfg
        L = len(reductions)
        B, T, I0 = x.shape
        dev = x.device
        H = params[1].shape[1]
        C = len(plan[0])
        need = any(ctx.needs_input_grad)
        c4 = ops.lstm_c4_supported(B, H)
        c4b = c4 and ops.C4_BWD                              # saves in lstm_c4's own layout, BPTT through lstm_c4
        ck = [_Chunks(B, lens) for lens in plan]
        P = [params[6 * l:6 * l + 6] for l in range(L)]
        wih16 = [ops.cast_bf16(_c(p[0])) for p in P]
        whh16 = [ops.cast_bf16(_c(p[1])) for p in P]
        bias = [p[2] + p[3] for p in P]
        x16 = [ck[0].scatter(ops.cast_bf16(_c(x)))] + [None] * L
        xs = [None] * (L + 1)                                # fp32 layer inputs (LayerNorm residuals), l >= 1
        y, y16, gates, cseq, mean, rstd = ([None] * L for _ in range(6))
        z = [None] * L
        for l in range(L):
            k, kn = ck[l], ck[l + 1]
            y[l] = k.new(H, f32, dev)
            y16[l] = k.new(H, bf16, dev) if (need or not c4) else None     # c4: holds h_{t-1} (the dW_hh operand)
            gates[l] = k.new(4 * H, f32, dev) if (need and not c4b) else None
            cseq[l] = k.new(H, f32, dev) if (need and not c4b) else None
            mean[l], rstd[l] = k.new(0, f32, dev), k.new(0, f32, dev)
            xs[l + 1] = kn.new(H, f32, dev)
            z[l] = k.new(H, f32, dev) if reductions[l] else xs[l + 1]
            x16[l + 1] = kn.new(H, bf16, dev) if l + 1 < L else None
        hT = torch.empty(L, C, B, H, dtype=f32, device=dev)
        cT = torch.empty(L, C, B, H, dtype=f32, device=dev)
        # Streams: the recurrence of layer l on rec[l % 2]; its input projection (xg = x W_ih^T + b, per chunk) on gem[l % 2]
        # and its LayerNorm / TimeReduction on nrm[l % 2].  The GEMM of chunk c+1 therefore runs AHEAD, under the recurrence of
        # chunk c, instead of queueing behind it (in one stream per layer parity the ~0.15 ms of GEMM + LayerNorm per chunk
        # were 20 % of the stream's time with no recurrence running); xg is double-buffered per layer.
        xgbuf = [[torch.empty(B * max(plan[l]), 4 * H, dtype=f32, device=dev) for _ in range(2)] for l in range(L)]
        c4saves = [[None] * C for _ in range(L)]               # c4: per (layer, chunk) saves in the kernels' layout

        main = torch.cuda.current_stream(dev)
        rec = _side_streams(dev)
        aux = _wave_aux_streams(dev)
        gem, nrm = aux[:2], aux[2:]
        ev = lambda: torch.cuda.Event()
        prep_done = [[ev() for _ in range(C)] for _ in range(L)]
        rec_done = [[ev() for _ in range(C)] for _ in range(L)]
        done = [[ev() for _ in range(C)] for _ in range(L)]
        for s_ in (*rec, *aux):
            s_.wait_stream(main)
        for d in range(L + C - 1):
            for l in range(max(0, d - C + 1), min(L, d + 1)):
                c = d - l
                k, kn = ck[l], ck[l + 1]
                Tc = k.lens[c]
                xg = xgbuf[l][c % 2][:B * Tc]
                with torch.cuda.stream(gem[l % 2]):
                    if l > 0:
                        gem[l % 2].wait_event(done[l - 1][c])
                    if c >= 2:
                        gem[l % 2].wait_event(rec_done[l][c - 2])          # the xg buffer is free again
                    xin = k.blk(x16[l], c)
                    ops.gemm_bf16(xin.view(B * Tc, xin.shape[2]), 0, wih16[l], 0, B * Tc, 4 * H, xin.shape[2],
                                  bias=bias[l], out=xg, flags=ops.GEMM_CORESIDENT)
                    prep_done[l][c].record(gem[l % 2])
                with torch.cuda.stream(rec[l % 2]):
                    rec[l % 2].wait_event(prep_done[l][c])
                    if c4:
                        r = ops.lstm_c4_fwd(xg.view(B, Tc, 4 * H), whh16[l], hT[l, c - 1] if c else None,
                                            cT[l, c - 1] if c else None, need,
                                            out=(k.blk(y[l], c), k.blk(y16[l], c) if need else None, hT[l, c], cT[l, c]),
                                            std_saves=None if (c4b or not need) else (k.blk(gates[l], c), k.blk(cseq[l], c)))
                        if need and c4b:
                            c4saves[l][c] = (r[4], r[5])
                            r[4].record_stream(main)
                            r[5].record_stream(main)
                    else:
                        ops.lstm_tc_fwd(xg.view(B, Tc, 4 * H), whh16[l], hT[l, c - 1] if c else None,
                                        cT[l, c - 1] if c else None, need,
                                        out=(k.blk(y[l], c), k.blk(y16[l], c), hT[l, c], cT[l, c],
                                             k.blk(gates[l], c) if need else None, k.blk(cseq[l], c) if need else None))
                    rec_done[l][c].record(rec[l % 2])
                with torch.cuda.stream(nrm[l % 2]):
                    nrm[l % 2].wait_event(rec_done[l][c])
                    res = k.blk(xs[l], c) if l else None
                    nx16 = kn.blk(x16[l + 1], c) if x16[l + 1] is not None else None
                    if reductions[l]:
                        ops.layernorm_fwd(k.blk(y[l], c), res, P[l][4], P[l][5], eps[l],
                                          out=(k.blk(z[l], c), None, k.blk(mean[l], c), k.blk(rstd[l], c)))
                        ops.time_reduce_fwd(k.blk(z[l], c), out=(kn.blk(xs[l + 1], c), nx16))
                    else:
                        ops.layernorm_fwd(k.blk(y[l], c), res, P[l][4], P[l][5], eps[l],
                                          out=(kn.blk(xs[l + 1], c), nx16, k.blk(mean[l], c), k.blk(rstd[l], c)))
                    done[l][c].record(nrm[l % 2])
        for s_ in (*rec, *aux):
            main.wait_stream(s_)
        out = ck[L].gather(xs[L])
        if LSTMStack.collect is not None:
            LSTMStack.collect.extend(ck[l + 1].gather(xs[l + 1]) for l in range(L))
        hT_last, cT_last = hT[:, C - 1].contiguous(), cT[:, C - 1].contiguous()
        if need:
            if c4b:
                gates = [torch.empty(0, device=dev)] * L
                cseq = [torch.empty(0, device=dev)] * L
            flat_saves = [t for row in c4saves for pair in row for t in pair] if c4b else []
            ctx.save_for_backward(*params, hT, cT, *x16[:L], *xs[1:L], *y, *y16, *gates, *cseq, *mean, *rstd, *flat_saves)
            ctx.cfg, ctx.dims, ctx.c4, ctx.c4b = cfg, (B, T, I0, H, L, C), c4, c4b
        ctx.mark_non_differentiable(hT_last, cT_last)
        return out, hT_last, cT_last

    @staticmethod
    def backward(ctx, dout, _dh, _dc):
        reductions, eps, plan = ctx.cfg
        B, T, I0, H, L, C = ctx.dims
        sv = list(ctx.saved_tensors)
        params, sv = sv[:6 * L], sv[6 * L:]
        hT, cT = sv[0], sv[1]
        sv = sv[2:]
        x16, sv = sv[:L], sv[L:]
        xs, sv = [None] + sv[:L - 1], sv[L - 1:]
        y, y16, gates, cseq, mean, rstd = (sv[i * L:(i + 1) * L] for i in range(6))
        c4, c4b = ctx.c4, ctx.c4b
        flat_saves = sv[6 * L:]
        if getattr(ctx, "consumed", False):
            raise RuntimeError("LSTMStack.backward ran twice on the same graph (retain_graph is not supported)")
        ctx.consumed = True
        P = [params[6 * l:6 * l + 6] for l in range(L)]
        ck = [_Chunks(B, lens) for lens in plan]
        dev = dout.device
        main = torch.cuda.current_stream(dev)
        side = _side_streams(dev)[0]
        side.wait_stream(main)
        g = ck[L].scatter(_c(dout))                           # d xs[L], chunk-major
        grads = [None] * (6 * L)
        for l in range(L - 1, -1, -1):
            k, kn = ck[l], ck[l + 1]
            if reductions[l]:
                gz = k.new(H, f32, dev)
                for c in range(C):
                    ops.time_reduce_bwd(kn.blk(g, c), k.lens[c], out=k.blk(gz, c))
                g = gz
            dz, dgamma, dbeta = ops.layernorm_bwd(g, y[l], xs[l], P[l][4], mean[l], rstd[l])
            whhT16 = ops.transpose_to_bf16(_c(P[l][1]))
            dg16 = k.new(4 * H, bf16, dev)
            dh = dc = None
            if c4b:
                hprev = y16[l]                                # written by the forward kernel
                for c in range(C - 1, -1, -1):
                    gs, cs = flat_saves[2 * (l * C + c)], flat_saves[2 * (l * C + c) + 1]
                    _, dh, dc = ops.lstm_c4_bwd(k.blk(dz, c), gs, cs, cT[l, c - 1] if c else None, whhT16, dh, dc,
                                                out=k.blk(dg16, c))
            else:
                hprev = y16[l] if c4 else k.new(H, bf16, dev)  # c4 forward: h_{t-1} already written by the kernel
                one_launch = c4 and BPTT_ONE_LAUNCH and C <= 8
                if one_launch:                                 # the kernel walks the chunk-major buffers itself
                    ops.lstm_tc_bwd_chunks(dz, gates[l], cseq[l], whhT16, k.lens, B, dg16)
                for c in (() if one_launch else range(C - 1, -1, -1)):
                    _, dh, dc = ops.lstm_tc_bwd(k.blk(dz, c), k.blk(gates[l], c), k.blk(cseq[l], c),
                                                cT[l, c - 1] if c else None, whhT16, dh, dc, out=k.blk(dg16, c))
                    if c4:
                        continue
                    hp, yc = k.blk(hprev, c), k.blk(y16[l], c)
                    hp[:, 1:] = yc[:, :-1]
                    if c:
                        hp[:, 0] = k.blk(y16[l], c - 1)[:, -1]
                    else:
                        hp[:, 0].zero_()
            # critical path to the next layer: d x = dG W_ih (+ dz: the LayerNorm residual branch, accumulated by the
            # GEMM's epilogue straight into dz)
            wih16 = ops.cast_bf16(_c(P[l][0]))
            M, I_l = dg16.shape[0], P[l][0].shape[1]
            if l > 0:
                g = ops.gemm_bf16(dg16, 0, wih16, 1, M, I_l, 4 * H, out=dz, accumulate=True, tag="gemm_bf16_nn")
            elif ctx.needs_input_grad[0]:
                g = ops.gemm_bf16(dg16, 0, wih16, 1, M, I_l, 4 * H, tag="gemm_bf16_nn")
            else:
                g = None
            # off the critical path: the weight / bias gradients of this layer run on a side stream under the BPTT
            # recurrence of the next layer (which leaves 20 SMs idle and the tensor pipe nearly so)
            ready = torch.cuda.Event()
            ready.record(main)
            with torch.cuda.stream(side):
                side.wait_event(ready)
                grads[6 * l + 0] = ops.mm_tn(dg16, x16[l], "bf16", dy16=dg16, x16=x16[l])
                grads[6 * l + 1] = ops.mm_tn(dg16, hprev, "bf16", dy16=dg16, x16=hprev)
                db = ops.colsum(dg16)
                grads[6 * l + 2], grads[6 * l + 3] = db, db.clone()
                for t_ in (dg16, hprev):
                    t_.record_stream(side)
                for t_ in grads[6 * l:6 * l + 4]:
                    t_.record_stream(main)
            grads[6 * l + 4], grads[6 * l + 5] = dgamma, dbeta
        main.wait_stream(side)
        dxin = ck[0].gather(g) if g is not None else None
        return (dxin, None, *grads)


def _joint_pre(h_enc, h_dec, w1, b1, precision):
    """ep = W1[:, :E] h_enc + b1, dp = W1[:, E:] h_dec  -- exact split of Linear(cat[e, d])."""
    B, T, E = h_enc.shape
    U, Dd = h_dec.shape[1], h_dec.shape[2]
    J = w1.shape[0]
    he2, hd2 = _c(h_enc).view(B * T, E), _c(h_dec).view(B * U, Dd)
    ep = ops.mm_nt(he2, w1[:, :E], b1, precision).view(B, T, J)
    dp = ops.mm_nt(hd2, w1[:, E:], None, precision).view(B, U, J)
    return he2, hd2, ep, dp


JOINT_WGRAD_SIDE = __import__("os").environ.get("EDGEDICT_JOINT_WGRAD_SIDE", "1") != "0"


def _joint_bwd(ctx_p, dlog2, hid, he2, hd2, w1, w2, dims, db2=None):
    """Shared backward of the joint given d logits [N_cells, V] (fp32 or bf16); db2 may already have
    been accumulated by the fused loss-gradient kernel."""
    B, T, U, E, Dd, J, V = dims
    p = ctx_p
    dl16 = dlog2 if dlog2.dtype == bf16 else (ops.cast_bf16(dlog2) if p == "bf16" else None)
    hid2 = hid.view(B * T * U, J)

    def out_layer_grads(db2):
        if db2 is None:
            db2 = ops.colsum(dlog2)
        if p == "bf16" and V % 256 == 0:
            # compute dW2^T = hidden^T dlogits ([J, V]: 256-wide tcgen05 tiles divide V, not J) and flip it
            dw2 = ops.mm_tn(hid2, dl16, p, dy16=hid2, x16=dl16).t().contiguous()
        else:
            dw2 = ops.mm_tn(dlog2, hid2, p, dy16=dl16, x16=hid2 if p == "bf16" else None)
        return dw2, db2

    side = None
    if JOINT_WGRAD_SIDE and p == "bf16" and dlog2.is_cuda:
        # the output layer's weight / bias gradients (a 4.2 GB column sum + the split-K dW2 GEMM, 2.5 ms at E6D2) are off
        # the critical path to the encoder: they run on a side stream under the d-hidden GEMM and the two d-pre reductions
        main = torch.cuda.current_stream(dlog2.device)
        side = _side_streams(dlog2.device)[1]
        side.wait_stream(main)
        with torch.cuda.stream(side):
            dw2, db2 = out_layer_grads(db2)
        for t_ in (dlog2, dl16, hid2):
            if t_ is not None:
                t_.record_stream(side)
    else:
        dw2, db2 = out_layer_grads(db2)
    if p == "bf16" and J % 8 == 0 and hid.dtype == bf16:
        # tanh' applied in the d-hidden GEMM's epilogue: d(pre-activation) leaves the GEMM, then two pure reductions
        dpre = ops.gemm_bf16_dtanh(dl16, ops.cast_bf16(w2.contiguous()), True, hid2, B * T * U, J, V)
        dep, ddp = ops.joint_dpre_reduce(dpre.view(B, T, U, J))
    else:
        dhid = ops.mm_nn(dlog2, w2, p, dy16=dl16, out_bf16=(p == "bf16"))
        dep, ddp = ops.joint_hidden_bwd(dhid.view(B, T, U, J), hid)
    dep2, ddp2 = dep.view(B * T, J), ddp.view(B * U, J)
    w1e, w1d = w1[:, :E], w1[:, E:]
    dhe = ops.mm_nn(dep2, w1e, p).view(B, T, E)
    dhd = ops.mm_nn(ddp2, w1d, p).view(B, U, Dd)
    dw1 = torch.empty_like(w1)
    dw1[:, :E] = ops.mm_tn(dep2, he2, p)
    dw1[:, E:] = ops.mm_tn(ddp2, hd2, p)
    db1 = ops.colsum(dep2)
    if side is not None:
        main.wait_stream(side)
        dw2.record_stream(main)
        db2.record_stream(main)
    return dhe, dhd, dw1, db1, dw2, db2


class JointLogits(torch.autograd.Function):
    """Joint.forward on [B,T,E] x [B,U,D] -> logits [B,T,U,V] (rnnt/models.py:169-179)."""

    @staticmethod
    def forward(ctx, h_enc, h_dec, w1, b1, w2, b2, precision):
        B, T, E = h_enc.shape
        U, Dd = h_dec.shape[1], h_dec.shape[2]
        J, V = w1.shape[0], w2.shape[0]
        he2, hd2, ep, dp = _joint_pre(h_enc, h_dec, w1, b1, precision)
        hid = ops.joint_hidden_fwd(ep, dp, precision == "bf16")
        logits = ops.mm_nt(hid.view(B * T * U, J), w2, b2, precision, x16=hid.view(B * T * U, J) if precision == "bf16" else None)
        ctx.save_for_backward(hid, he2, hd2, w1, w2)
        ctx.precision, ctx.dims = precision, (B, T, U, E, Dd, J, V)
        return logits.view(B, T, U, V)

    @staticmethod
    def backward(ctx, dlogits):
        hid, he2, hd2, w1, w2 = ctx.saved_tensors
        B, T, U, E, Dd, J, V = ctx.dims
        dl2 = _c(dlogits).view(B * T * U, V)
        dhe, dhd, dw1, db1, dw2, db2 = _joint_bwd(ctx.precision, dl2, hid, he2, hd2, w1, w2, ctx.dims)
        return dhe, dhd, dw1, db1, dw2, db2, None


class RNNTLossFn(torch.autograd.Function):
    """warprnnt_pytorch._RNNT (pytorch_binding/warprnnt_pytorch/__init__.py:10-50) on CUDA logits:
    costs stay on the device, the gradient kernel runs in backward with the upstream gradient
    folded in (the reference computes it eagerly and rescales it in a second pass)."""

    @staticmethod
    def forward(ctx, acts, labels, act_lens, label_lens, blank, reduction):
        acts = _c(acts)
        costs, ws = ops.rnnt_loss_fwd(acts, labels, act_lens, label_lens, blank, need_beta=True)
        B = acts.shape[0]
        ctx.save_for_backward(acts, labels, act_lens, label_lens, ws)
        ctx.blank, ctx.reduction, ctx.B = blank, reduction, B
        if reduction in ("sum", "mean"):
            costs = costs.sum().unsqueeze(-1)
            if reduction == "mean":
                costs = costs / B
        return costs

    @staticmethod
    def backward(ctx, go):
        acts, labels, act_lens, label_lens, ws = ctx.saved_tensors
        scale = 1.0 / ctx.B if ctx.reduction == "mean" else 1.0
        g = _c(go.to(acts.dtype)).view(-1)
        grads = ops.rnnt_loss_bwd(acts, labels, act_lens, label_lens, ctx.blank, ws, g, scale)
        return grads, None, None, None, None, None


class JointLoss(torch.autograd.Function):
    """Transducer.forward's joint + loss (rnnt/models.py:234-239) as one autograd node: logits are
    produced, consumed by the loss, and their gradient is written IN PLACE over them (fp32 mode)
    or straight to bf16 (bf16 mode) -- no autograd copy of the 8 GB tensor is ever made."""

    @staticmethod
    def forward(ctx, h_enc, h_dec, w1, b1, w2, b2, labels, act_lens, label_lens, blank, precision):
        B, T, E = h_enc.shape
        U, Dd = h_dec.shape[1], h_dec.shape[2]
        J, V = w1.shape[0], w2.shape[0]
        he2, hd2, ep, dp = _joint_pre(h_enc, h_dec, w1, b1, precision)
        hid = ops.joint_hidden_fwd(ep, dp, precision == "bf16")
        hid2 = hid.view(B * T * U, J)
        fused = precision == "bf16" and FUSE_JOINT_LSE and J % 8 == 0 and U <= 1024
        if fused:
            # bf16 mode: the logits GEMM epilogue also produces the softmax statistics (fp32, from the
            # TMEM accumulators) and writes bf16 logits; the denominator pass over 8 GB disappears
            b2a = b2 if (b2.is_contiguous() and b2.data_ptr() % 16 == 0) else b2.clone()
            logits, ws = ops.joint_logits_lse(hid2, ops.cast_bf16(w2.contiguous()), b2a, labels, act_lens,
                                              label_lens, B, T, U, blank)
            costs = ops.rnnt_lattice(act_lens, label_lens, B, T, U, ws)
        else:
            logits = ops.mm_nt(hid2, w2, b2, precision, x16=hid2 if precision == "bf16" else None).view(B, T, U, V)
            costs, ws = ops.rnnt_loss_fwd(logits, labels, act_lens, label_lens, blank, need_beta=True)
        ctx.save_for_backward(hid, he2, hd2, w1, w2, logits, labels, act_lens, label_lens, ws)
        ctx.precision, ctx.dims, ctx.blank = precision, (B, T, U, E, Dd, J, V), blank
        ctx.mark_non_differentiable(costs)
        loss = costs.sum().unsqueeze(-1) / B
        ctx.costs = costs
        return loss, costs

    @staticmethod
    def backward(ctx, go, _gc):
        hid, he2, hd2, w1, w2, logits, labels, act_lens, label_lens, ws = ctx.saved_tensors
        if getattr(ctx, "consumed", False):
            raise RuntimeError("JointLoss.backward ran twice on the same graph: the gradient is written in place over "
                               "the saved logits (retain_graph is not supported by this node)")
        ctx.consumed = True
        B, T, U, E, Dd, J, V = ctx.dims
        p = ctx.precision
        g = _c(go.to(f32)).view(-1)
        # (a variant of the gradient kernel that also accumulated the bias gradient in registers was
        #  measured 2.5x slower -- occupancy -- than this kernel plus a separate column-sum pass)
        if logits.dtype == bf16:
            dl = ops.rnnt_loss_bwd_bf16(logits, labels, act_lens, label_lens, ctx.blank, ws, g, 1.0 / B)
        elif p == "bf16":
            dl = ops.rnnt_loss_bwd(logits, labels, act_lens, label_lens, ctx.blank, ws, g, 1.0 / B, out_bf16=True)
        else:
            dl = ops.rnnt_loss_bwd(logits, labels, act_lens, label_lens, ctx.blank, ws, g, 1.0 / B, out=logits)
        dhe, dhd, dw1, db1, dw2, db2 = _joint_bwd(p, dl.view(B * T * U, V), hid, he2, hd2, w1, w2, ctx.dims)
        return dhe, dhd, dw1, db1, dw2, db2, None, None, None, None, None
