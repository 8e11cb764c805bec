"""Functional torch restatement of the reference path with autograd -- TEST INFRASTRUCTURE.

Used (a) as the gradient oracle for the CUDA backward kernels (fp32 or fp64, CPU),
(b) as the CPU baseline arm of bench.py (``--impl reference`` / ``cpu_baseline``): with
``fast=True`` the LSTM layers go through ``torch._VF.lstm`` -- the same ATen/oneDNN kernel
the reference's ``nn.LSTM`` modules dispatch to (rnnt/models.py:45-46,145-147) -- and the loss
goes through the reference's own compiled CPU library (oracle/_ref) when it is present.

All functions take ``sd``: {reference state_dict key -> tensor}.  Cited lines as in
oracle/model_np.py.  Pinned by tests/test_oracle_model.py against tests/golden/*.npz.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

from . import loss as _loss

NUL, PAD, BOS, UNK = 0, 1, 2, 3


def _n(sd, stem):
    n = 0
    while (stem % n) in sd:
        n += 1
    return n


def lstm_layer(x, h0, c0, w_ih, w_hh, b_ih, b_hh, fast=False):
    """One unidirectional batch_first LSTM layer.  x [B,T,I]; h0,c0 [B,H]."""
    if fast:
        y, h, c = torch._VF.lstm(x, (h0[None], c0[None]), [w_ih, w_hh, b_ih, b_hh],
                                 True, 1, 0.0, False, False, True)
        return y, h[0], c[0]
    H = w_hh.shape[1]
    xg = x @ w_ih.t() + (b_ih + b_hh)
    h, c, ys = h0, c0, []
    for t in range(x.shape[1]):
        g = xg[:, t] + h @ w_hh.t()
        i, f, gg, o = g[:, :H].sigmoid(), g[:, H:2 * H].sigmoid(), g[:, 2 * H:3 * H].tanh(), g[:, 3 * H:].sigmoid()
        c = f * c + i * gg
        h = o * c.tanh()
        ys.append(h)
    return torch.stack(ys, 1), h, c


def time_reduction(x, factor=2):
    B, T, H = x.shape
    pad = (factor - T % factor) % factor
    if pad:
        x = F.pad(x, [0, 0, 0, pad])
    return x.reshape(B, -1, factor, H).mean(2)


def encoder(sd, xs, hiddens=None, time_reductions=(1,), pre="encoder.", fast=False, collect=None):
    L = _n(sd, pre + "lstm.lstms.%d.weight_ih_l0")
    H = sd[pre + "lstm.lstms.0.weight_hh_l0"].shape[1]
    B = xs.shape[0]
    if hiddens is None:
        hs = xs.new_zeros(L, B, H)
        cs = xs.new_zeros(L, B, H)
    else:
        hs, cs = hiddens
    x = F.layer_norm(xs, (xs.shape[-1],), sd[pre + "norm.weight"], sd[pre + "norm.bias"], 1e-5)
    nh, nc = [], []
    for i in range(L):
        p = pre + "lstm.lstms.%d." % i
        y, h, c = lstm_layer(x, hs[i], cs[i], sd[p + "weight_ih_l0"], sd[p + "weight_hh_l0"],
                             sd[p + "bias_ih_l0"], sd[p + "bias_hh_l0"], fast)
        x = y if i == 0 else x + y
        q = pre + "lstm.projs.%d.0." % i
        x = F.layer_norm(x, (H,), sd[q + "weight"], sd[q + "bias"], 1e-5)
        if i in time_reductions:
            x = time_reduction(x)
        if collect is not None:
            collect.append(x)
        nh.append(h)
        nc.append(c)
    if (pre + "proj.weight") in sd:
        x = F.linear(x, sd[pre + "proj.weight"], sd[pre + "proj.bias"])
    return x, (torch.stack(nh), torch.stack(nc))


def decoder(sd, ys, hidden=None, pre="decoder.", fast=False):
    L = _n(sd, pre + "lstm.weight_ih_l%d")
    H = sd[pre + "lstm.weight_hh_l0"].shape[1]
    B = ys.shape[0]
    w = sd[pre + "embed.weight"]
    if hidden is None:
        ys = F.pad(ys, [1, 0, 0, 0], value=BOS)
        h0 = w.new_zeros(L, B, H)
        c0 = w.new_zeros(L, B, H)
    else:
        h0, c0 = hidden
    x = F.embedding(ys.long(), w, padding_idx=PAD)
    nh, nc = [], []
    for k in range(L):
        x, h, c = lstm_layer(x, h0[k], c0[k], sd[pre + "lstm.weight_ih_l%d" % k],
                             sd[pre + "lstm.weight_hh_l%d" % k], sd[pre + "lstm.bias_ih_l%d" % k],
                             sd[pre + "lstm.bias_hh_l%d" % k], fast)
        nh.append(h)
        nc.append(c)
    x = F.linear(x, sd[pre + "proj.weight"], sd[pre + "proj.bias"])
    return x, (torch.stack(nh), torch.stack(nc))


def joint(sd, h_enc, h_dec, pre="joint."):
    if h_enc.dim() == 3:
        T, U = h_enc.shape[1], h_dec.shape[1]
        h_enc = h_enc[:, :, None, :].expand(-1, -1, U, -1)
        h_dec = h_dec[:, None, :, :].expand(-1, T, -1, -1)
    z = torch.cat([h_enc, h_dec], -1)
    z = torch.tanh(F.linear(z, sd[pre + "joint.0.weight"], sd[pre + "joint.0.bias"]))
    return F.linear(z, sd[pre + "joint.2.weight"], sd[pre + "joint.2.bias"])


def scale_length(T_out, xlen):
    scale = (xlen.max().float() / T_out).ceil()
    return (xlen / scale).ceil().int()


class _LossFn(torch.autograd.Function):
    """warprnnt_pytorch._RNNT on CPU (``__init__.py:10-50``): input log-probs, 'mean' = /B."""

    @staticmethod
    def forward(ctx, log_probs, labels, act_lens, label_lens, blank, reduction, use_ref):
        npdt = np.float32 if log_probs.dtype == torch.float32 else np.float64
        fn = _loss.ref_cpu if (use_ref and _loss.have_ref()) else _loss.logprobs
        costs, grads = fn(log_probs.detach().numpy(), labels.numpy(), act_lens.numpy(),
                          label_lens.numpy(), blank=blank, want_grads=True, dtype=npdt)
        costs = torch.from_numpy(costs)
        grads = torch.from_numpy(grads)
        B = log_probs.shape[0]
        if reduction in ("sum", "mean"):
            costs = costs.sum().unsqueeze(-1)
            if reduction == "mean":
                costs = costs / B
                grads = grads / B
        ctx.grads = grads
        return costs

    @staticmethod
    def backward(ctx, go):
        return ctx.grads * go.view(-1, 1, 1, 1), None, None, None, None, None, None


def rnnt_loss(logits, labels, act_lens, label_lens, blank=0, reduction="mean", use_ref=True):
    lp = F.log_softmax(logits, -1)
    return _LossFn.apply(lp, labels.int(), act_lens.int(), label_lens.int(), blank, reduction, use_ref)


def transducer_loss(sd, xs, ys, xlen, ylen, blank=NUL, time_reductions=(1,), fast=False,
                    use_ref=True, keep=None):
    """Transducer.forward with output_loss=True (rnnt/models.py:228-241)."""
    xs = xs[:, :int(xlen.max())].contiguous()
    ys = ys[:, :int(ylen.max())].contiguous()
    h_enc, _ = encoder(sd, xs, None, time_reductions, fast=fast)
    h_dec, _ = decoder(sd, ys, None, fast=fast)
    logits = joint(sd, h_enc, h_dec)
    if keep is not None:
        keep.update(h_enc=h_enc, h_dec=h_dec, logits=logits)
    xl = scale_length(logits.shape[1], xlen)
    return rnnt_loss(logits, ys.int(), xl, ylen.int(), blank, "mean", use_ref)


@torch.no_grad()
def greedy_decode(sd, xs, xlen, blank=NUL, time_reductions=(1,), fast=False):
    h_enc, _ = encoder(sd, xs, None, time_reductions, fast=fast)
    B = xs.shape[0]
    h_dec, (hp, cp) = decoder(sd, torch.zeros(B, 0, dtype=torch.long), None, fast=fast)
    h_dec, hp, cp = h_dec.clone(), hp.clone(), cp.clone()
    seq, lps = [], []
    for i in range(h_enc.shape[1]):
        lp = F.log_softmax(joint(sd, h_enc[:, i], h_dec[:, 0]), 1)
        p, pred = lp.max(1)
        seq.append(pred)
        lps.append(p)
        nd, (hn, cn) = decoder(sd, pred[:, None], (hp, cp), fast=fast)
        m = pred != blank
        h_dec[m] = nd[m]
        hp[:, m] = hn[:, m]
        cp[:, m] = cn[:, m]
    seq = torch.stack(seq, 1)
    return [s[:int(n)].numpy() for s, n in zip(seq, xlen)], -torch.stack(lps, 1).sum(1)


@torch.no_grad()
def beam_search(sd, xs, xlen=None, W=4, blank=NUL, merge=True, time_reductions=(1,)):
    """Restatement of edgedict_b200's time-synchronous beam (Transducer.beam_search; SURVEY 8(f) N4 -- the reference
    holds no beam search in rnnt/, so the pin is W = 1 == greedy_decode plus this independent CPU restatement)."""
    h_enc_all, _ = encoder(sd, xs, None, time_reductions)
    outs, nlps = [], []
    for b in range(xs.shape[0]):
        Tn = h_enc_all.shape[1]
        frames = Tn if xlen is None else min(Tn, int(scale_length(Tn, xlen)[b]))
        dec_x, (dh, dc) = decoder(sd, torch.zeros(1, 0, dtype=torch.long), None)
        hyps = [dict(seq=[], lp=torch.zeros(()), x=dec_x[0, 0], h=dh[:, 0], c=dc[:, 0])]
        for t in range(frames):
            cand = []
            for qi, hy in enumerate(hyps):
                lp = F.log_softmax(joint(sd, h_enc_all[b, t][None], hy["x"][None])[0], 0) + hy["lp"]
                cand += [(float(lp[k]), qi, k, lp[k]) for k in range(lp.shape[0])]
            cand.sort(key=lambda c: (-c[0], c[1], c[2]))
            new, seen = [], {}
            for _, qi, k, lpk in cand[:W]:
                hy = hyps[qi]
                seq = hy["seq"] + ([k] if k != blank else [])
                key = tuple(seq)
                if merge and key in seen:
                    seen[key]["lp"] = torch.logaddexp(seen[key]["lp"], lpk)
                    continue
                nh = dict(seq=seq, lp=lpk, x=hy["x"], h=hy["h"], c=hy["c"])
                if k != blank:
                    nx, (h2, c2) = decoder(sd, torch.full((1, 1), k), (hy["h"][:, None], hy["c"][:, None]))
                    nh.update(x=nx[0, 0], h=h2[:, 0], c=c2[:, 0])
                seen[key] = nh
                new.append(nh)
            hyps = new
        best = max(hyps, key=lambda h: float(h["lp"]))
        outs.append(best["seq"])
        nlps.append(-best["lp"])
    return outs, torch.stack(nlps)


def encoder_gru(sd, xs, hiddens=None, time_reductions=(1,), pre="encoder."):
    """rnnt/models.py:77-116 + :131-136 (Encoder with ResLayerNormGRU): explicit GRU cell loop, gate order r|z|n."""
    L = _n(sd, pre + "lstm.lstms.%d.weight_ih_l0")
    H = sd[pre + "lstm.lstms.0.weight_hh_l0"].shape[1]
    B = xs.shape[0]
    hs = xs.new_zeros(L, B, H) if hiddens is None else hiddens
    x = F.layer_norm(xs, (xs.shape[-1],), sd[pre + "norm.weight"], sd[pre + "norm.bias"], 1e-5)
    nh = []
    for i in range(L):
        p = pre + "lstm.lstms.%d." % i
        w_ih, w_hh, b_ih, b_hh = (sd[p + k] for k in ("weight_ih_l0", "weight_hh_l0", "bias_ih_l0", "bias_hh_l0"))
        h, ys = hs[i], []
        for t in range(x.shape[1]):
            gi, gh = x[:, t] @ w_ih.t() + b_ih, h @ w_hh.t() + b_hh
            r = torch.sigmoid(gi[:, :H] + gh[:, :H])
            z = torch.sigmoid(gi[:, H:2 * H] + gh[:, H:2 * H])
            n = torch.tanh(gi[:, 2 * H:] + r * gh[:, 2 * H:])
            h = (1 - z) * n + z * h
            ys.append(h)
        y = torch.stack(ys, 1)
        x = y if i == 0 else x + y
        q = pre + "lstm.projs.%d.0." % i
        x = F.layer_norm(x, (H,), sd[q + "weight"], sd[q + "bias"], 1e-5)
        if i in time_reductions:
            x = time_reduction(x)
        nh.append(h)
    if (pre + "proj.weight") in sd:
        x = F.linear(x, sd[pre + "proj.weight"], sd[pre + "proj.bias"])
    return x, torch.stack(nh)


class StreamState:
    def __init__(self, sd, fast=False):
        w = sd["decoder.embed.weight"]
        L = _n(sd, "encoder.lstm.lstms.%d.weight_ih_l0")
        H = sd["encoder.lstm.lstms.0.weight_hh_l0"].shape[1]
        Ld = _n(sd, "decoder.lstm.weight_ih_l%d")
        Hd = sd["decoder.lstm.weight_hh_l0"].shape[1]
        self.enc_h = w.new_zeros(L, 1, H)
        self.enc_c = w.new_zeros(L, 1, H)
        with torch.no_grad():
            self.dec_x, (self.dec_h, self.dec_c) = decoder(
                sd, torch.full((1, 1), BOS), (w.new_zeros(Ld, 1, Hd), w.new_zeros(Ld, 1, Hd)), fast=fast)


@torch.no_grad()
def stream_decode(sd, st, chunk, unk_id=UNK, time_reductions=(1,), fast=False):
    """rnnt/stream.py:93-120 on one synthetic log-mel chunk [1,n,F]; returns emitted ids."""
    enc, (st.enc_h, st.enc_c) = encoder(sd, chunk, (st.enc_h, st.enc_c), time_reductions, fast=fast)
    out = []
    for k in range(enc.shape[1]):
        prob = joint(sd, enc[:, k], st.dec_x[:, 0])
        pred = int(prob.argmax(-1))
        if pred == unk_id:
            prob[:, pred] = 0
            pred = int(prob.argmax(-1))
        if pred != NUL:
            st.dec_x, (st.dec_h, st.dec_c) = decoder(sd, torch.full((1, 1), pred), (st.dec_h, st.dec_c), fast=fast)
            out.append(pred)
    return out
