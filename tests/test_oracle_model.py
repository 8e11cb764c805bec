"""Pins the model restatements (oracle/model_np.py, oracle/model_torch.py) against the golden
fixtures produced by the reference itself (tests/golden/make_golden.py)."""
import numpy as np
import torch

from oracle import model_np as mnp
from oracle import model_torch as mt
from tests.util import load_tiny, load_e4d1, to_t, rel_err


def test_np_forward_tiny():
    z, cfg, sd, _ = load_tiny()
    acts = []
    h_enc, (eh, ec) = mnp.encoder(sd, z["xs"], collect=acts)
    for i, a in enumerate(acts):
        assert rel_err(a, z["layer_act.%d" % i]) < 1e-5
    assert rel_err(h_enc, z["h_enc"]) < 1e-5
    assert rel_err(eh, z["enc_h"]) < 1e-5 and rel_err(ec, z["enc_c"]) < 1e-5
    h_dec, (dh, dc) = mnp.decoder(sd, z["ys"])
    assert rel_err(h_dec, z["h_dec"]) < 1e-5
    assert rel_err(dh, z["dec_h"]) < 1e-5 and rel_err(dc, z["dec_c"]) < 1e-5
    logits = mnp.joint(sd, h_enc, h_dec)
    assert rel_err(logits, z["logits"]) < 1e-5
    assert (mnp.scale_length(logits.shape[1], z["xlen"]) == z["xlen_scaled"]).all()


def test_np_greedy_and_stream_tiny():
    z, cfg, sd, _ = load_tiny()
    ids, nlp = mnp.greedy_decode(sd, z["xs"], z["xlen"])
    for i, row in zip(ids, z["greedy_ids"]):
        assert (i == row[:len(i)]).all()
    assert rel_err(nlp, z["greedy_nlp"]) < 1e-5
    st = mnp.StreamState(sd)
    for ch, tok in zip(z["stream_chunks"], z["stream_tokens"]):
        out = mnp.stream_decode(sd, st, ch[None])
        assert (out[0] if out else -1) == tok


def test_torch_forward_backward_tiny():
    z, cfg, sd, pg = load_tiny()
    for dtype, tol in ((torch.float32, 2e-5), (torch.float64, 2e-5)):
        tsd = {k: v.requires_grad_(True) for k, v in to_t(sd, dtype).items()}
        keep = {}
        loss = mt.transducer_loss(tsd, torch.as_tensor(z["xs"]).to(dtype), torch.as_tensor(z["ys"]),
                                  torch.as_tensor(z["xlen"]), torch.as_tensor(z["ylen"]), keep=keep)
        assert rel_err(keep["logits"].detach(), z["logits"]) < tol
        assert rel_err(loss.detach(), z["loss"]) < tol
        keep["logits"].retain_grad()
        loss.backward()
        for k, g in pg.items():
            assert rel_err(tsd[k].grad, g) < 5e-4, k


def test_torch_fast_lstm_equals_loop():
    z, cfg, sd, _ = load_tiny()
    tsd = to_t(sd)
    xs = torch.as_tensor(z["xs"])
    a, _ = mt.encoder(tsd, xs, fast=True)
    b, _ = mt.encoder(tsd, xs, fast=False)
    assert rel_err(a, b) < 1e-5


def test_torch_greedy_and_stream_tiny():
    z, cfg, sd, _ = load_tiny()
    tsd = to_t(sd)
    ids, nlp = mt.greedy_decode(tsd, torch.as_tensor(z["xs"]), z["xlen"])
    for i, row in zip(ids, z["greedy_ids"]):
        assert (i == row[:len(i)]).all()
    st = mt.StreamState(tsd)
    for ch, tok in zip(z["stream_chunks"], z["stream_tokens"]):
        out = mt.stream_decode(tsd, st, torch.as_tensor(ch[None]))
        assert (out[0] if out else -1) == tok
