"""Host side of the persistent streaming-decode kernel (csrc/decode.cu).

``StreamEngine(transducer, n_streams, frames_per_chunk)`` owns the per-stream recurrent state on
the device and a *phase program* (array of ``EbPhase``) built once; ``step(chunk)`` copies the
chunk's log-mel frames into a fixed input buffer and launches ONE cooperative kernel that runs the
stateful encoder, then for every encoder output frame joint -> argmax (with the ``<unk>`` rule) ->
masked predictor step, for all streams at once.  Semantics per stream are exactly those of
PytorchStreamDecoder.reset/decode (reference rnnt/stream.py:78-120): at most one symbol per
encoder frame, argmax over raw logits, predictor advanced only on non-blank.
"""
import ctypes as C

import torch

from ._lib import lib, check
from .rnnt.tokenizer import NUL, BOS, UNK

PH_LN, PH_PAIR, PH_LSTM, PH_LINEAR, PH_ARGMAX, PH_COPY = range(6)
F_TANH, F_EMBED, F_MASKED, F_LOGP = 1, 2, 4, 8


class EbPhase(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("type", "S", "K1", "K2", "N", "flags", "ldx1", "ldx2", "ldw1", "ldw2",
                                         "ldy", "aux", "aux2", "hist_ld", "hist_col", "pad_")] + \
               [(n, C.c_void_p) for n in ("x1", "x2", "w1", "w2", "b1", "b2", "y", "y2", "c", "tok_in", "tok_out",
                                          "hist")]


def _ptr(t, off=0):
    return None if t is None else t.data_ptr() + off * t.element_size()


def param_fingerprint(module):
    """Identity of the parameter storage a phase program was built over: the programs bake raw device pointers, so a
    re-homed parameter (FlatAdam bucket, .to(), .float(), p.data = ...) must trigger a rebuild."""
    return tuple(p.data_ptr() for p in module.parameters())


class StreamEngine:
    STATE = ("enc_h", "enc_c", "dec_h", "dec_c", "dec_x", "tok")

    def __init__(self, transducer, n_streams, frames_per_chunk, unk_id=UNK, blank=NUL, max_ctas=0, state=None):
        assert C.sizeof(EbPhase) == lib().eb_decode_phase_size(), "EbPhase layout mismatch"
        enc, dec, joint = transducer.encoder, transducer.decoder, transducer.joint.joint
        self.dev = enc.norm.weight.device
        if self.dev.type != "cuda":
            raise RuntimeError("StreamEngine needs the model on a CUDA device")
        f32 = torch.float32
        S, n = n_streams, frames_per_chunk
        self.S, self.n, self.blank, self.unk, self.max_ctas = S, n, blank, unk_id, max_ctas
        lstms = list(enc.lstm.lstms)
        L = len(lstms)
        H = enc.lstm.hidden_size
        F = enc.norm.weight.shape[0]
        reductions = enc.lstm.time_reductions
        z = lambda *shape: torch.zeros(*shape, dtype=f32, device=self.dev)
        self.xin, self.a0 = z(S, n, F), z(S, n, F)
        self.enc_h, self.enc_c, self.enc_htmp = z(L, S, H), z(L, S, H), z(L, S, H)
        self._keep = [p.detach() for p in transducer.parameters()]      # weights are read in place
        self.fingerprint = param_fingerprint(transducer)
        prog = []

        def ph(**kw):
            p = EbPhase()
            for k, v in kw.items():
                setattr(p, k, v)
            prog.append(p)

        ph(type=PH_LN, S=S * n, N=F, x1=_ptr(self.xin), ldx1=F, w1=_ptr(enc.norm.weight), b1=_ptr(enc.norm.bias),
           y=_ptr(self.a0), ldy=F)
        X, I, ni = self.a0, F, n
        self._bufs = []
        for i, (cell, post) in enumerate(zip(lstms, enc.lstm.projs)):
            yL, zL = z(S, ni, H), z(S, ni, H)
            self._bufs += [yL, zL]
            for t in range(ni):
                ph(type=PH_LSTM, S=S, N=H, K1=I, K2=H, x1=_ptr(X, t * I), ldx1=ni * I,
                   x2=_ptr(self.enc_h[i]) if t == 0 else _ptr(yL, (t - 1) * H), ldx2=H if t == 0 else ni * H,
                   w1=_ptr(cell.weight_ih_l0), ldw1=I, w2=_ptr(cell.weight_hh_l0), ldw2=H,
                   b1=_ptr(cell.bias_ih_l0), b2=_ptr(cell.bias_hh_l0), c=_ptr(self.enc_c[i]),
                   y=_ptr(yL, t * H), ldy=ni * H, y2=_ptr(self.enc_htmp[i]) if t == ni - 1 else None)
            ln = post[0]
            ph(type=PH_LN, S=S * ni, N=H, x1=_ptr(yL), ldx1=H, x2=_ptr(X) if i > 0 else None, ldx2=H,
               w1=_ptr(ln.weight), b1=_ptr(ln.bias), y=_ptr(zL), ldy=H)
            X, I = zL, H
            if i in reductions:
                if ni % 2:
                    raise ValueError("streaming chunks must hold an even number of frames before each time "
                                     "reduction (cli/export_onnx.py:20-21 asserts the same)")
                zr = z(S, ni // 2, H)
                self._bufs.append(zr)
                ph(type=PH_PAIR, S=S, N=H, aux=ni, x1=_ptr(zL), y=_ptr(zr))
                X, ni = zr, ni // 2
        self.n_out = ni
        E = enc.proj.weight.shape[0] if enc.has_proj else H
        if enc.has_proj:
            self.enc_out = z(S, ni, E)
            ph(type=PH_LINEAR, S=S * ni, N=E, K1=H, x1=_ptr(X), ldx1=H, w1=_ptr(enc.proj.weight), ldw1=H,
               b1=_ptr(enc.proj.bias), y=_ptr(self.enc_out), ldy=E)
        else:
            self.enc_out = X
        # ---- predictor + joint state
        Ld, Hd = dec.lstm.num_layers, dec.lstm.hidden_size
        Em = dec.embed.weight.shape[1]
        D = dec.proj.weight.shape[0]
        J, V = joint[0].weight.shape[0], joint[2].weight.shape[0]
        assert joint[0].weight.shape[1] == E + D
        self.dec_h, self.dec_c, self.dec_htmp = z(Ld, S, Hd), z(Ld, S, Hd), z(Ld, S, Hd)
        self.dec_x, self.hidden, self.logits = z(S, D), z(S, J), z(S, V)
        self.tok = torch.zeros(S, dtype=torch.int32, device=self.dev)
        self.hist = torch.zeros(S, max(ni, 1), dtype=torch.int32, device=self.dev)

        def predictor_phases(masked):
            fl = F_EMBED | (F_MASKED if masked else 0)
            for k in range(Ld):
                w = [getattr(dec.lstm, s % k) for s in ("weight_ih_l%d", "weight_hh_l%d", "bias_ih_l%d", "bias_hh_l%d")]
                ph(type=PH_LSTM, S=S, N=Hd, K1=Em if k == 0 else Hd, K2=Hd, flags=fl if k == 0 else (fl & F_MASKED),
                   x1=_ptr(dec.embed.weight) if k == 0 else _ptr(self.dec_htmp[k - 1]), ldx1=Em if k == 0 else Hd,
                   x2=_ptr(self.dec_h[k]), ldx2=Hd, w1=_ptr(w[0]), ldw1=w[0].shape[1], w2=_ptr(w[1]), ldw2=Hd,
                   b1=_ptr(w[2]), b2=_ptr(w[3]), c=_ptr(self.dec_c[k]), y=_ptr(self.dec_htmp[k]), ldy=Hd,
                   tok_in=_ptr(self.tok), aux=blank)
            ph(type=PH_COPY, S=Ld * S, N=Hd, x1=_ptr(self.dec_htmp), y=_ptr(self.dec_h))
            ph(type=PH_LINEAR, S=S, N=D, K1=Hd, x1=_ptr(self.dec_h[Ld - 1]), ldx1=Hd, w1=_ptr(dec.proj.weight), ldw1=Hd,
               b1=_ptr(dec.proj.bias), y=_ptr(self.dec_x), ldy=D)

        w1 = joint[0].weight
        for k in range(ni):
            ph(type=PH_LINEAR, S=S, N=J, flags=F_TANH, K1=E, x1=_ptr(self.enc_out, k * E), ldx1=ni * E, w1=_ptr(w1),
               ldw1=E + D, K2=D, x2=_ptr(self.dec_x), ldx2=D, w2=_ptr(w1, E), ldw2=E + D, b1=_ptr(joint[0].bias),
               y=_ptr(self.hidden), ldy=J)
            ph(type=PH_LINEAR, S=S, N=V, K1=J, x1=_ptr(self.hidden), ldx1=J, w1=_ptr(joint[2].weight), ldw1=J,
               b1=_ptr(joint[2].bias), y=_ptr(self.logits), ldy=V)
            ph(type=PH_ARGMAX, S=S, N=V, x1=_ptr(self.logits), ldx1=V, aux=blank, aux2=unk_id, tok_out=_ptr(self.tok),
               hist=_ptr(self.hist), hist_ld=self.hist.shape[1], hist_col=k)
            predictor_phases(masked=True)
        ph(type=PH_COPY, S=L * S, N=H, x1=_ptr(self.enc_htmp), y=_ptr(self.enc_h))
        self.n_chunk_phases = len(prog)
        chunk_prog = prog
        prog = []
        predictor_phases(masked=False)                  # priming program: tok = BOS from a zero state
        self.n_prime_phases = len(prog)
        self._chunk = self._upload(chunk_prog)
        self._prime = self._upload(prog)
        self._bar = torch.zeros(64, dtype=torch.int32, device=self.dev)
        if state is None:
            self.reset()
        else:
            self.load_state(state)                      # a rebuilt program continues the utterance (rnnt/stream.py:97-98)

    def state(self):
        """The recurrent state of every stream (what PytorchStreamDecoder carries between chunks, rnnt/stream.py:78-91)."""
        return {k: getattr(self, k).clone() for k in self.STATE}

    @torch.no_grad()
    def load_state(self, st):
        for k in self.STATE:
            getattr(self, k).copy_(st[k])

    def _upload(self, prog):
        arr = (EbPhase * len(prog))(*prog)
        raw = bytes(arr)
        t = torch.frombuffer(bytearray(raw), dtype=torch.uint8).to(self.dev)
        return t

    def _run(self, prog, nphase):
        check(lib().eb_decode_run(prog.data_ptr(), nphase, self._bar.data_ptr(), self.max_ctas,
                                  torch.cuda.current_stream().cuda_stream), "eb_decode_run")

    @torch.no_grad()
    def reset(self):
        """PytorchStreamDecoder.reset (rnnt/stream.py:78-91) for every stream."""
        for t in (self.enc_h, self.enc_c, self.dec_h, self.dec_c):
            t.zero_()
        self.tok.fill_(BOS)
        self._run(self._prime, self.n_prime_phases)

    @torch.no_grad()
    def step(self, chunk):
        """chunk [S, n, F] log-mel frames (device or pinned host) -> int32 [S, n_out] token ids
        (blank = 0 means 'no symbol for this frame')."""
        self.xin.copy_(chunk, non_blocking=True)
        self._run(self._chunk, self.n_chunk_phases)
        return self.hist


class GreedyEngine:
    """Device-side batched greedy decode (reference Transducer.greedy_decode, rnnt/models.py:243-269):
    the T' per-frame iterations (joint -> log_softmax/argmax -> predictor step for every row -> keep
    the new state only where the prediction is non-blank) run inside ONE cooperative kernel launch as
    a phase program over the encoder output, instead of T' Python iterations of ~10 launches each."""

    def __init__(self, transducer, batch, t_out, blank=NUL, max_ctas=0):
        dec, joint = transducer.decoder, transducer.joint.joint
        self.dev = dec.embed.weight.device
        f32 = torch.float32
        B, T = batch, t_out
        self.B, self.T, self.blank, self.max_ctas = B, T, blank, max_ctas
        z = lambda *shape: torch.zeros(*shape, dtype=f32, device=self.dev)
        Ld, Hd = dec.lstm.num_layers, dec.lstm.hidden_size
        Em, D = dec.embed.weight.shape[1], dec.proj.weight.shape[0]
        J, V = joint[0].weight.shape[0], joint[2].weight.shape[0]
        E = joint[0].weight.shape[1] - D
        self.h_enc = z(B, T, E)
        self.dec_h, self.dec_c, self.dec_htmp = z(Ld, B, Hd), z(Ld, B, Hd), z(Ld, B, Hd)
        self.dec_x, self.hidden, self.logits = z(B, D), z(B, J), z(B, V)
        self.tok = torch.zeros(B, dtype=torch.int32, device=self.dev)
        self.hist = torch.zeros(B, T, dtype=torch.int32, device=self.dev)
        self.logp = z(B)
        self._keep = [p.detach() for p in transducer.parameters()]
        prog = []

        def ph(**kw):
            q = EbPhase()
            for k, v in kw.items():
                setattr(q, k, v)
            prog.append(q)

        def predictor(masked):
            fl = F_EMBED | (F_MASKED if masked else 0)
            for k in range(Ld):
                w = [getattr(dec.lstm, n % k) for n in ("weight_ih_l%d", "weight_hh_l%d", "bias_ih_l%d", "bias_hh_l%d")]
                ph(type=PH_LSTM, S=B, N=Hd, K1=Em if k == 0 else Hd, K2=Hd, flags=fl if k == 0 else (fl & F_MASKED),
                   x1=_ptr(dec.embed.weight) if k == 0 else _ptr(self.dec_htmp[k - 1]), ldx1=Em if k == 0 else Hd,
                   x2=_ptr(self.dec_h[k]), ldx2=Hd, w1=_ptr(w[0]), ldw1=w[0].shape[1], w2=_ptr(w[1]), ldw2=Hd,
                   b1=_ptr(w[2]), b2=_ptr(w[3]), c=_ptr(self.dec_c[k]), y=_ptr(self.dec_htmp[k]), ldy=Hd,
                   tok_in=_ptr(self.tok), aux=blank)
            ph(type=PH_COPY, S=Ld * B, N=Hd, x1=_ptr(self.dec_htmp), y=_ptr(self.dec_h))
            ph(type=PH_LINEAR, S=B, N=D, K1=Hd, x1=_ptr(self.dec_h[Ld - 1]), ldx1=Hd, w1=_ptr(dec.proj.weight), ldw1=Hd,
               b1=_ptr(dec.proj.bias), y=_ptr(self.dec_x), ldy=D)

        predictor(masked=False)                          # prime with BOS from the zero state
        w1 = joint[0].weight
        for k in range(T):
            ph(type=PH_LINEAR, S=B, N=J, flags=F_TANH, K1=E, x1=_ptr(self.h_enc, k * E), ldx1=T * E, w1=_ptr(w1),
               ldw1=E + D, K2=D, x2=_ptr(self.dec_x), ldx2=D, w2=_ptr(w1, E), ldw2=E + D, b1=_ptr(joint[0].bias),
               y=_ptr(self.hidden), ldy=J)
            ph(type=PH_LINEAR, S=B, N=V, K1=J, x1=_ptr(self.hidden), ldx1=J, w1=_ptr(joint[2].weight), ldw1=J,
               b1=_ptr(joint[2].bias), y=_ptr(self.logits), ldy=V)
            ph(type=PH_ARGMAX, S=B, N=V, flags=F_LOGP, x1=_ptr(self.logits), ldx1=V, aux=blank, aux2=-1,
               tok_out=_ptr(self.tok), hist=_ptr(self.hist), hist_ld=T, hist_col=k, y=_ptr(self.logp))
            predictor(masked=True)
        self.nphase = len(prog)
        arr = (EbPhase * len(prog))(*prog)
        self._prog = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(self.dev)
        self._bar = torch.zeros(64, dtype=torch.int32, device=self.dev)

    @torch.no_grad()
    def run(self, h_enc):
        """h_enc [B, T', E] -> (ids int32 [B, T'] incl. blanks, sum of log p [B])."""
        self.h_enc.copy_(h_enc)
        for t in (self.dec_h, self.dec_c, self.logp):
            t.zero_()
        self.tok.fill_(BOS)
        check(lib().eb_decode_run(self._prog.data_ptr(), self.nphase, self._bar.data_ptr(), self.max_ctas,
                                  torch.cuda.current_stream().cuda_stream), "eb_decode_run")
        return self.hist, self.logp
