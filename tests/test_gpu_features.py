"""Log-mel front end (csrc/frontend.cu through edgedict_b200.rnnt.features) against the numpy oracle
(oracle/features_np.py, itself pinned to torch.stft and the Slaney filterbank in tests/test_oracle_features.py).
Tolerance: absolute 2e-3 on the log-mel values (fp32 direct DFT vs fp32 FFT; |log| <= ~12 here), i.e. ~1e-3
relative on the power spectrum wherever it is not vanishing."""
import numpy as np
import pytest
import torch

from oracle import features_np as F

pytestmark = pytest.mark.gpu


def _wave(B, L, seed):
    g = torch.Generator().manual_seed(seed)
    t = torch.arange(L) / 16000.0
    x = 0.1 * torch.randn(B, L, generator=g)
    for b in range(B):
        x[b] += 0.5 * torch.sin(2 * np.pi * (220.0 * (b + 1)) * t) + 0.2 * torch.sin(2 * np.pi * 3100.0 * t + b)
    return x


@pytest.mark.parametrize("B,L,win,hop,n_filt", [(3, 4800, 400, 200, 80), (2, 16000, 320, 160, 64), (1, 3217, 400, 200, 80),
                                                (5, 9000, 400, 320, 80)])
def test_filterbank_features_match_oracle(B, L, win, hop, n_filt):
    from edgedict_b200.rnnt.features import FilterbankFeatures
    x = _wave(B, L, L)
    m = FilterbankFeatures(win_length=win, hop_length=hop, n_filt=n_filt, dither=0).cuda()
    got = m(x.clone().cuda()).cpu().numpy()
    want = F.filterbank_features(x.numpy(), win_length=win, hop_length=hop, n_filt=n_filt)
    assert got.shape == want.shape == (B, n_filt, 1 + L // hop)
    assert np.abs(got - want).max() < 2e-3
    if L % hop == 0:
        assert (got[:, :, -1] == 0).all()                       # masked frame (features.py:160-164)


@pytest.mark.parametrize("L,ds,divisible", [(30000, 3, True), (30100, 3, True), (30100, 3, False), (8000, 2, True)])
def test_fused_frontend_and_downsample_module(L, ds, divisible):
    from edgedict_b200.rnnt.features import LogMelFrontend, build_transform
    x = _wave(2, L, 7)
    fe = LogMelFrontend(80, downsample=ds, pad_to_divisible=divisible, dither=0).cuda()
    got = fe(x.clone().cuda()).cpu().numpy()
    f = F.filterbank_features(x.numpy(), win_length=400, hop_length=200, n_filt=80)
    want = F.downsample(f, ds, divisible).transpose(0, 2, 1)
    assert got.shape == want.shape and got.shape[2] == 80 * ds == fe.input_size
    assert np.abs(got - want).max() < 2e-3
    # the unfused reference composition (FilterbankFeatures -> Downsample) gives the same tensor in [B, C, T] layout
    _, test_tf, size = build_transform("logfbank", 80, downsample=ds, pad_to_divisible=divisible)
    test_tf = test_tf.cuda()
    test_tf[0].dither = 0
    ref_layout = test_tf(x.clone().cuda())
    assert size == 80 * ds and tuple(ref_layout.shape) == (2, 80 * ds, got.shape[1])
    assert np.abs(ref_layout.transpose(1, 2).cpu().numpy() - got).max() < 1e-6


def test_frontend_feeds_the_encoder():
    """waveform -> features -> Transducer.greedy_decode runs end to end on the device (E6D2 feature geometry)."""
    from edgedict_b200.rnnt.features import LogMelFrontend
    from edgedict_b200.rnnt.models import Transducer
    torch.manual_seed(1)
    m = Transducer(vocab_embed_size=16, vocab_size=64, input_size=240, enc_hidden_size=64, enc_layers=2, enc_dropout=0,
                   enc_proj_size=32, dec_hidden_size=32, dec_layers=1, dec_dropout=0, dec_proj_size=32, joint_size=48).cuda()
    fe = LogMelFrontend(80, downsample=3, dither=0).cuda()
    xs = fe(_wave(2, 24000, 5).cuda())
    assert tuple(xs.shape) == (2, 41, 240)
    ids, nlp = m.greedy_decode(xs, torch.tensor([41, 41]))
    assert len(ids) == 2 and torch.isfinite(nlp).all()
