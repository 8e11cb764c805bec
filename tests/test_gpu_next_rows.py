"""SURVEY 8(f) "next" rows and 8(a) a22: beam search (N4), SpecAugment masks (N2), GRU encoder variant (a22)."""
import random

import numpy as np
import pytest
import torch

from tests.util import load_tiny, rel_err

pytestmark = pytest.mark.gpu


def _tiny():
    from edgedict_b200.rnnt.models import Transducer
    z, cfg, sd, _ = load_tiny()
    m = Transducer(output_loss=False, **cfg)
    m.load_state_dict({k: torch.as_tensor(v) for k, v in sd.items()})
    return m.cuda().eval(), z, {k: torch.as_tensor(v) for k, v in sd.items()}


def test_beam_width_one_is_greedy_decode():
    """The only pin the reference offers for a beam search: W = 1 == Transducer.greedy_decode."""
    m, z, _ = _tiny()
    xs, xlen = torch.as_tensor(z["xs"]).cuda(), torch.as_tensor(z["xlen"])
    ids, nlp = m.greedy_decode(xs, xlen)
    full = torch.full_like(xlen, int(xlen.max()))
    ids_full, nlp_full = m.greedy_decode(xs, full)
    seqs, blp = m.beam_search(xs, None, W=1)
    for got, want in zip(seqs, ids_full):
        assert got == [int(t) for t in want if t != 0]
    assert rel_err(blp.cpu(), nlp_full.cpu()) < 1e-4


@pytest.mark.parametrize("W,merge", [(2, True), (4, True), (3, False)])
def test_beam_search_matches_cpu_restatement(W, merge):
    from oracle import model_torch as mt
    m, z, sd = _tiny()
    xs = torch.as_tensor(z["xs"])
    want, wlp = mt.beam_search(sd, xs, None, W=W, merge=merge)
    got, glp = m.beam_search(xs.cuda(), None, W=W, merge=merge)
    assert got == want
    assert rel_err(glp.cpu(), wlp) < 1e-4
    # a wider beam never scores worse than greedy
    _, g1 = m.beam_search(xs.cuda(), None, W=1)
    assert (glp <= g1 + 1e-4).all()


@pytest.mark.parametrize("cls,axis", [("TimeMasking", 2), ("FrequencyMasking", 1)])
def test_specaugment_masks_match_reference_loop(cls, axis):
    """rnnt/transforms.py:53-147 restated: same python `random` call order -> same spans -> same masked features."""
    from edgedict_b200.rnnt import features as ft
    x = torch.randn(5, 24, 37)
    mod = getattr(ft, cls)(max_width=6, num_masks=3)
    random.seed(11)
    got = mod(x.cuda()).cpu()
    random.seed(11)
    mask = torch.zeros(x.shape, dtype=torch.bool)
    for i in range(x.shape[0]):
        for _ in range(3):
            start = random.randrange(0, x.shape[axis])
            end = start + random.randrange(0, 6)
            if axis == 2:
                mask[i, :, start:end] = True
            else:
                mask[i, start:end, :] = True
    want = x.masked_fill(mask, 0.0)
    assert torch.equal(got, want) and mask.any()
    tr, te, size = ft.build_transform("logfbank", 16, downsample=2, T_mask=5, T_num_mask=2, F_mask=3, F_num_mask=1)
    assert size == 32 and len(tr) == len(te) + 2 and isinstance(tr[-1], ft.FrequencyMasking)


def test_gru_encoder_variant_torch_fallback():
    """SURVEY 8(a) a22 (rnnt/models.py:77-116): module_type='GRU' builds and runs (torch fallback), keeps the
    reference's state_dict keys and matches an explicit GRU cell loop."""
    from edgedict_b200.rnnt.models import Transducer
    from oracle import model_torch as mt
    torch.manual_seed(3)
    cfg = dict(vocab_embed_size=16, vocab_size=64, input_size=24, enc_hidden_size=48, enc_layers=3, enc_dropout=0,
               enc_proj_size=40, dec_hidden_size=32, dec_layers=1, dec_dropout=0, dec_proj_size=24, joint_size=56)
    m = Transducer(module_type="GRU", output_loss=False, **cfg).cuda().eval()
    keys = set(m.state_dict().keys())
    assert {"encoder.lstm.lstms.0.weight_ih_l0", "encoder.lstm.projs.1.0.weight", "encoder.proj.bias"} <= keys
    assert m.encoder.lstm.lstms[0].weight_ih_l0.shape[0] == 3 * 48
    xs = torch.randn(2, 11, 24)
    with torch.no_grad():
        out, hs = m.encoder(xs.cuda())
    sd = {k: v.detach().cpu() for k, v in m.state_dict().items()}
    ref, rh = mt.encoder_gru(sd, xs)
    assert out.shape == ref.shape == (2, 6, 40) and hs.shape == (3, 2, 48)
    # torch fallback on a GPU = cuDNN's GRU, which runs its GEMMs in TF32 by default: 1e-3, not fp32 round-off
    assert rel_err(out.cpu(), ref) < 1e-3 and rel_err(hs.cpu(), rh) < 1e-3
    ys = torch.randint(4, 64, (2, 5), dtype=torch.int32).cuda()
    logits = m(xs.cuda(), ys, torch.tensor([11, 11]), torch.tensor([5, 5]))
    assert logits.shape == (2, 6, 6, 64)
