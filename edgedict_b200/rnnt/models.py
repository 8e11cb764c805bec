"""B200-native mirror of the reference's ``rnnt/models.py`` hot path.

Same class names, constructor arguments, ``forward`` signatures, return values and
``state_dict`` keys as /root/reference/rnnt/models.py:16-269, so ``cli/train.py``,
``cli/baseline.py``, ``cli/lightning.py`` and ``rnnt/stream.py`` can import this module
unchanged.  The torch ``nn.LSTM`` / ``nn.LayerNorm`` / ``nn.Linear`` / ``nn.Embedding`` objects
below are PARAMETER CONTAINERS ONLY (identical default initialisation and checkpoint keys); their
``forward`` is never called -- all arithmetic goes through the C-ABI of libedgedict_b200.so
(edgedict_b200/functional.py).  CUDA tensors are mandatory: there is no CPU fallback.

precision: "fp32" (default; parity mode, CUDA-core GEMMs, matches torch-CPU fp32 to ~1e-5) or
"bf16" (tcgen05 tensor-core GEMMs with fp32 accumulation; also selected automatically inside
``torch.autocast('cuda')``, the modern spelling of the reference's apex-O1 switch).
"""
import os

import torch
from torch import nn

from .. import functional as Fn
from .. import ops
from .tokenizer import NUL, BOS, PAD

_DEFAULT_PRECISION = os.environ.get("EDGEDICT_PRECISION", "fp32")
_PREDICTOR_STREAM = os.environ.get("EDGEDICT_PREDICTOR_STREAM", "1") != "0"
_pred_streams = {}


def _predictor_stream(device):
    s = _pred_streams.get(device)
    if s is None:
        s = _pred_streams[device] = torch.cuda.Stream(device)
    return s


def _precision(module):
    if torch.is_autocast_enabled():
        return "bf16"
    return getattr(module, "precision", _DEFAULT_PRECISION)


def _set_precision(root, precision):
    assert precision in ("fp32", "bf16")
    for m in root.modules():
        m.precision = precision
    return root


class TimeReduction(nn.Module):
    """rnnt/models.py:16-29 (only reduction_factor == 2 is used by the reference configs)."""

    def __init__(self, reduction_factor=2):
        super().__init__()
        if reduction_factor != 2:
            raise NotImplementedError("edgedict_b200 implements the reference's factor-2 reduction")
        self.reduction_factor = reduction_factor

    def forward(self, xs):
        return Fn.TimeReduce.apply(xs)


class ResLayerNormLSTM(nn.Module):
    """rnnt/models.py:32-75.  ``lstms.{i}`` / ``projs.{i}.0`` keep the reference's key layout."""

    def __init__(self, input_size, hidden_size, num_layers, dropout=0, time_reductions=[1],
                 reduction_factor=2):
        super().__init__()
        self.hidden_size = hidden_size
        self.lstms = nn.ModuleList()
        self.projs = nn.ModuleList()
        self.time_reductions = set(time_reductions)
        for i in range(num_layers):
            self.lstms.append(nn.LSTM(input_size, hidden_size, 1, batch_first=True))
            stack = [nn.LayerNorm(hidden_size)]
            if i in self.time_reductions:
                stack.append(TimeReduction(reduction_factor))
            if dropout > 0:
                stack.append(nn.Dropout(dropout))
            self.projs.append(nn.Sequential(*stack))
            input_size = hidden_size

    def _wavefront_cfg(self, xs, p):
        """(cfg, params) of the layer-wavefront schedule (functional.LSTMStack) when it applies: bf16 tensor-core
        mode, zero initial state, no active dropout, a sequence long enough to be cut into chunks."""
        if p != "bf16" or len(self.lstms) < 2 or not xs.is_cuda:
            return None
        B, T = xs.shape[0], xs.shape[1]
        if not ops.lstm_tc_supported(B, self.hidden_size):
            return None
        reductions, eps, params = [], [], []
        for cell, post in zip(self.lstms, self.projs):
            extras = list(post)[1:]
            if any(isinstance(m, nn.Dropout) and self.training and m.p > 0 for m in extras):
                return None
            reductions.append(any(isinstance(m, TimeReduction) for m in extras))
            eps.append(post[0].eps)
            params += [cell.weight_ih_l0, cell.weight_hh_l0, cell.bias_ih_l0, cell.bias_hh_l0,
                       post[0].weight, post[0].bias]
        plan = Fn.wavefront_plan(T, reductions)
        if plan is None:
            return None
        return (tuple(reductions), tuple(eps), plan), params

    def forward(self, xs, hiddens=None):
        p = _precision(self)
        if hiddens is None:
            wf = self._wavefront_cfg(xs, p)
            if wf is not None:
                out, hT, cT = Fn.LSTMStack.apply(xs, wf[0], *wf[1])
                return out, (hT, cT)
        hs, cs = (None, None) if hiddens is None else hiddens
        out_h, out_c = [], []
        for i, (cell, post) in enumerate(zip(self.lstms, self.projs)):
            h0 = None if hs is None else hs[i]
            c0 = None if cs is None else cs[i]
            y, hT, cT = Fn.LSTMLayer.apply(xs, h0, c0, cell.weight_ih_l0, cell.weight_hh_l0,
                                           cell.bias_ih_l0, cell.bias_hh_l0, p)
            ln = post[0]
            xs = Fn.LayerNormRes.apply(y, xs if i != 0 else None, ln.weight, ln.bias, ln.eps)
            for extra in list(post)[1:]:
                xs = extra(xs)
            out_h.append(hT)
            out_c.append(cT)
        return xs, (torch.stack(out_h, 0), torch.stack(out_c, 0))


class ResLayerNormGRU(nn.Module):
    """rnnt/models.py:77-116, the GRU encoder variant (`module_type='GRU'`, only cli/lightning.py:63 selects it).
    SURVEY 8(a) a22: kept as a TORCH FALLBACK -- nn.GRU / nn.LayerNorm run through ATen (cuDNN on a GPU), not through
    this library's kernels; same constructor, ``state_dict`` keys (`lstms.{i}`, `projs.{i}.0`) and return value
    (xs, hs [L, B, H]) as the reference, so a GRU checkpoint loads and decodes."""

    def __init__(self, input_size, hidden_size, num_layers, dropout=0, time_reductions=[1], reduction_factor=2):
        super().__init__()
        self.hidden_size = hidden_size
        self.lstms = nn.ModuleList()
        self.projs = nn.ModuleList()
        self.time_reductions = set(time_reductions)
        for i in range(num_layers):
            self.lstms.append(nn.GRU(input_size, hidden_size, 1, batch_first=True))
            proj = [nn.LayerNorm(hidden_size)]
            if i in self.time_reductions:
                proj.append(TimeReduction(reduction_factor))
            if dropout > 0:
                proj.append(nn.Dropout(dropout))
            input_size = hidden_size
            self.projs.append(nn.Sequential(*proj))

    @staticmethod
    def _time_reduce(x, factor=2):
        B, T, H = x.shape
        pad = (factor - T % factor) % factor
        if pad:
            x = nn.functional.pad(x, [0, 0, 0, pad])
        return x.reshape(B, -1, factor, H).mean(2)

    def forward(self, xs, hiddens=None):
        hs = xs.new_zeros(len(self.lstms), xs.shape[0], self.hidden_size) if hiddens is None else hiddens
        new_hs = []
        for i, (gru, proj) in enumerate(zip(self.lstms, self.projs)):
            ys, h = gru(xs, hs[i, None].contiguous())
            xs = ys if i == 0 else xs + ys
            for m in proj:                                   # torch modules, except the parameter-free reduction
                xs = self._time_reduce(xs, m.reduction_factor) if isinstance(m, TimeReduction) else m(xs)
            new_hs.append(h)
        return xs, torch.cat(new_hs, dim=0)


class Encoder(nn.Module):
    """rnnt/models.py:119-136.  ``module`` defaults to the LSTM stack (the reference's default argument is the GRU
    variant, but every caller that matters passes the LSTM one; ResLayerNormGRU above is the torch fallback)."""

    def __init__(self, input_size, hidden_size, num_layers, dropout, proj_size,
                 module=ResLayerNormLSTM, time_reductions=[1], has_proj=True):
        super().__init__()
        self.norm = nn.LayerNorm(input_size)
        self.lstm = module(input_size, hidden_size, num_layers, dropout=dropout,
                           time_reductions=time_reductions)
        self.has_proj = has_proj
        if has_proj:
            self.proj = nn.Linear(hidden_size, proj_size)

    def forward(self, xs, hiddens=None):
        xs = Fn.LayerNormRes.apply(xs, None, self.norm.weight, self.norm.bias, self.norm.eps)
        xs, hiddens = self.lstm(xs, hiddens)
        if self.has_proj:
            xs = Fn.Linear.apply(xs, self.proj.weight, self.proj.bias, _precision(self))
        return xs, hiddens


class Decoder(nn.Module):
    """Prediction network, rnnt/models.py:139-157."""

    def __init__(self, vocab_embed_size, vocab_size, hidden_size, num_layers, dropout=0,
                 proj_size=None):
        super().__init__()
        self.embed = nn.Embedding(vocab_size, vocab_embed_size, padding_idx=PAD)
        self.lstm = nn.LSTM(vocab_embed_size, hidden_size, num_layers, batch_first=True,
                            dropout=dropout)
        self.proj = nn.Linear(hidden_size, proj_size)
        self.dropout = dropout

    def forward(self, ys, hidden=None):
        p = _precision(self)
        prime = hidden is None
        xs = Fn.Embedding.apply(ys, self.embed.weight, prime, BOS, PAD)
        L = self.lstm.num_layers
        hs, cs = (None, None) if prime else hidden
        out_h, out_c = [], []
        for k in range(L):
            w = [getattr(self.lstm, n % k) for n in ("weight_ih_l%d", "weight_hh_l%d", "bias_ih_l%d", "bias_hh_l%d")]
            xs, hT, cT = Fn.LSTMLayer.apply(xs, None if hs is None else hs[k], None if cs is None else cs[k],
                                            w[0], w[1], w[2], w[3], p)
            if self.dropout > 0 and self.training and k < L - 1:
                xs = nn.functional.dropout(xs, self.dropout, True)      # nn.LSTM inter-layer dropout
            out_h.append(hT)
            out_c.append(cT)
        ys = Fn.Linear.apply(xs, self.proj.weight, self.proj.bias, p)
        return ys, (torch.stack(out_h, 0), torch.stack(out_c, 0))


class Joint(nn.Module):
    """rnnt/models.py:160-179; ``joint.0`` / ``joint.2`` key layout kept (Tanh at index 1)."""

    def __init__(self, input_size, hidden_size, vocab_size):
        super().__init__()
        self.joint = nn.Sequential(nn.Linear(input_size, hidden_size), nn.Tanh(),
                                   nn.Linear(hidden_size, vocab_size))

    def forward(self, h_enc, h_dec):
        two_d = h_enc.dim() == 2 and h_dec.dim() == 2
        if two_d:
            h_enc, h_dec = h_enc[:, None, :], h_dec[:, None, :]
        elif not (h_enc.dim() == 3 and h_dec.dim() == 3):
            raise AssertionError("Joint expects [B,T,E]/[B,U,D] or [B,E]/[B,D]")
        l0, l2 = self.joint[0], self.joint[2]
        if l0.weight.shape[1] != h_enc.shape[-1] + h_dec.shape[-1]:
            raise ValueError("joint input size mismatch")
        out = Fn.JointLogits.apply(h_enc, h_dec, l0.weight, l0.bias, l2.weight, l2.bias, _precision(self))
        return out[:, 0, 0] if two_d else out


class Transducer(nn.Module):
    """rnnt/models.py:182-269."""

    def __init__(self, vocab_embed_size, vocab_size, input_size, enc_hidden_size, enc_layers,
                 enc_dropout, enc_proj_size, dec_hidden_size, dec_layers, dec_dropout, dec_proj_size,
                 joint_size, enc_time_reductions=[1], blank=NUL, module_type='LSTM', output_loss=True):
        super().__init__()
        self.blank = blank
        if module_type not in ['GRU', 'LSTM']:
            raise ValueError('Unsupported module type')
        # rnnt/models.py:196-205: the GRU variant runs as a torch fallback, the LSTM one on this engine
        self.encoder = Encoder(input_size=input_size, hidden_size=enc_hidden_size, num_layers=enc_layers,
                               dropout=enc_dropout, proj_size=enc_proj_size,
                               time_reductions=enc_time_reductions,
                               module=ResLayerNormGRU if module_type == 'GRU' else ResLayerNormLSTM)
        self.decoder = Decoder(vocab_embed_size=vocab_embed_size, vocab_size=vocab_size,
                               hidden_size=dec_hidden_size, num_layers=dec_layers, dropout=dec_dropout,
                               proj_size=dec_proj_size)
        self.joint = Joint(input_size=enc_proj_size + dec_proj_size, hidden_size=joint_size,
                           vocab_size=vocab_size)
        self.output_loss = output_loss
        if output_loss:
            from ..warprnnt_pytorch import RNNTLoss
            self.loss_fn = RNNTLoss(blank=blank)
        self.last_costs = None

    def set_precision(self, precision):
        return _set_precision(self, precision)

    def scale_length(self, logits, xlen):
        return scale_length(logits.shape[1], xlen)

    def forward(self, xs, ys, xlen, ylen):
        xs = xs[:, :int(xlen.max())].contiguous()
        ys = ys[:, :int(ylen.max())].contiguous()
        if xs.is_cuda and _PREDICTOR_STREAM:
            # The prediction network (2 x 129 recurrent steps) is independent of the encoder until the joint: it runs on
            # a side stream under the encoder's recurrence (its kernels use 32 of the 148 SMs at H_d = 256); autograd
            # replays each node's backward on the stream of its forward, so the backward passes overlap the same way.
            main = torch.cuda.current_stream(xs.device)
            side = _predictor_stream(xs.device)
            side.wait_stream(main)
            ys.record_stream(side)
            with torch.cuda.stream(side):
                h_dec, _ = self.decoder(ys)
            h_enc, _ = self.encoder(xs)
            main.wait_stream(side)
            h_dec.record_stream(main)
        else:
            h_enc, _ = self.encoder(xs)
            h_dec, _ = self.decoder(ys)
        if not self.output_loss:
            return self.joint(h_enc, h_dec)
        xl = _lens_to_device(_i32(scale_length(h_enc.shape[1], xlen)), h_enc.device)
        yl = _lens_to_device(_i32(ylen), h_enc.device)
        l0, l2 = self.joint.joint[0], self.joint.joint[2]
        loss, costs = Fn.JointLoss.apply(h_enc, h_dec, l0.weight, l0.bias, l2.weight, l2.bias,
                                         _i32(ys), xl, yl, self.blank, _precision(self))
        self.last_costs = costs
        return loss

    @torch.no_grad()
    def greedy_decode(self, xs, xlen):
        """rnnt/models.py:243-269: at most one symbol per encoder frame; returns (list of id arrays
        incl. blanks, truncated by the UNSCALED xlen as the reference does, -sum log p).  The T'
        per-frame iterations run device-side in one persistent kernel (stream_engine.GreedyEngine)."""
        from ..stream_engine import GreedyEngine, param_fingerprint
        h_enc, _ = self.encoder(xs)
        B, T = h_enc.shape[0], h_enc.shape[1]
        # the phase program bakes raw weight pointers: re-homed parameters (FlatAdam, .to(), .float()) rebuild it
        key = (B, T, h_enc.device, param_fingerprint(self))
        cache = self.__dict__.setdefault("_greedy_engines", {})
        eng = cache.get(key)
        if eng is None:
            cache.clear()                                  # one resident program is enough
            eng = cache[key] = GreedyEngine(self, B, T, blank=self.blank)
        ids, logp = eng.run(h_enc)
        ids = ids.cpu().numpy()
        out = [ids[i, :int(n)].astype("int64") for i, n in enumerate(xlen)]
        return out, -logp.clone()


    @torch.no_grad()
    def beam_search(self, xs, xlen=None, W=4, merge=True):
        """beam decode.  The reference has no beam search in rnnt/ (north_star mentions one); its
        legacy v0 stack holds a batch-1 Graves-style search (models.py:121-202, with no-op `sorted(...)` calls and a
        removed `volatile=` API).  This is a time-synchronous beam under the SAME emission constraint as
        `greedy_decode` (at most one symbol per encoder frame, rnnt/models.py:243-269): per frame every hypothesis is
        scored against the whole vocabulary in one joint call, the W best continuations survive (hypotheses that
        reach the same token sequence are merged by log-add when `merge`), and only the survivors that emitted a
        non-blank take a predictor step (one batched call).  W = 1 reproduces `greedy_decode` token for token.
        xs [B,T,F] -> (list of non-blank id lists, -log p [B]); utterances are searched one at a time like the
        reference's beam."""
        h_enc_all, _ = self.encoder(xs)
        B, Tn = h_enc_all.shape[0], h_enc_all.shape[1]
        dev = h_enc_all.device
        outs, nlps = [], []
        for b in range(B):
            frames = Tn if xlen is None else min(Tn, int(scale_length(Tn, xlen)[b]))
            dec_x, (dh, dc) = self.decoder(torch.zeros(1, 0, dtype=torch.long, device=dev))     # BOS prime
            dec_x = dec_x[:, 0]                                                                   # [n, D]
            seqs, logp = [[]], torch.zeros(1, device=dev)
            for t in range(frames):
                n = dec_x.shape[0]
                he = h_enc_all[b, t][None].expand(n, -1).contiguous()
                logits = self.joint(he, dec_x.contiguous())                                      # [n, V]
                lp = torch.log_softmax(logits.float(), 1) + logp[:, None]                        # [n, V]
                V = lp.shape[1]
                top, idx = lp.reshape(-1).topk(min(W, n * V))                                   # ties: lowest index first
                par, tok = (idx // V).tolist(), (idx % V).tolist()
                new_seqs = [seqs[q] + ([k] if k != self.blank else []) for q, k in zip(par, tok)]
                keep, merged_lp, seen = [], [], {}
                for i, sq in enumerate(new_seqs):
                    key = tuple(sq)
                    if merge and key in seen:
                        j = seen[key]
                        merged_lp[j] = torch.logaddexp(merged_lp[j], top[i])
                        continue
                    seen[key] = len(keep)
                    keep.append(i)
                    merged_lp.append(top[i])
                par_t = torch.tensor([par[i] for i in keep], device=dev)
                tok_t = torch.tensor([tok[i] for i in keep], device=dev)
                seqs = [new_seqs[i] for i in keep]
                logp = torch.stack(merged_lp)
                dec_x, dh, dc = dec_x[par_t], dh[:, par_t], dc[:, par_t]
                nb = (tok_t != self.blank).nonzero()[:, 0]
                if nb.numel():
                    nx, (nh, nc) = self.decoder(tok_t[nb][:, None], (dh[:, nb].contiguous(), dc[:, nb].contiguous()))
                    dec_x, dh, dc = dec_x.clone(), dh.clone(), dc.clone()
                    dec_x[nb], dh[:, nb], dc[:, nb] = nx[:, 0], nh, nc
            best = int(logp.argmax())
            outs.append(seqs[best])
            nlps.append(-logp[best])
        return outs, torch.stack(nlps)


def _i32(t):
    return t.to(torch.int32).contiguous()


def _lens_to_device(t, device):
    """Host length vector -> device without stalling the host: a copy from PAGEABLE memory synchronises the stream first
    (the host then sits behind the whole encoder forward, 0.2 - 0.4 ms of idle GPU per step in the kernel timeline); a pinned
    staging tensor + non_blocking copy does not (the caching host allocator keeps the block until the copy has run)."""
    if t.device == device or device.type != "cuda":
        return t.to(device)
    staged = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
    staged.copy_(t)
    return staged.to(device, non_blocking=True)


def scale_length(T_out, xlen):
    """Transducer.scale_length (rnnt/models.py:223-226) on the host lengths."""
    scale = (xlen.max().float() / T_out).ceil()
    return (xlen / scale).ceil().int()


def convert_lightning2normal(checkpoint):
    """rnnt/models.py:368-380: unwrap a Lightning checkpoint; when its keys carry the ``model.``
    prefix, drop it and re-wrap as {'model': state_dict}."""
    if 'state_dict' not in checkpoint:
        return checkpoint
    sd = checkpoint['state_dict']
    if 'model.' in next(iter(sd.keys())):
        return {'model': {k.replace('model.', ''): v for k, v in sd.items()}}
    return sd
