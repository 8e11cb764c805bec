"""Pins of the front-end oracle (oracle/features_np.py) against the third-party arithmetic the reference calls:
torch.stft (rnnt/features.py:121-124) and the Slaney mel filterbank (librosa.filters.mel, rnnt/features.py:76-80;
librosa is absent, torchaudio's independent implementation of the same published formula is the pin)."""
import numpy as np
import pytest
import torch

from oracle import features_np as F


@pytest.mark.parametrize("L,n_fft,hop,win", [(4000, 512, 160, 320), (3217, 512, 200, 400), (1600, 256, 80, 256)])
def test_stft_power_matches_torch_stft(L, n_fft, hop, win):
    g = torch.Generator().manual_seed(L)
    x = torch.randn(2, L, generator=g, dtype=torch.float64)
    w = torch.hann_window(win, periodic=False, dtype=torch.float64)
    ref = torch.stft(x, n_fft=n_fft, hop_length=hop, win_length=win, window=w, center=True, pad_mode="reflect",
                     normalized=False, onesided=True, return_complex=True).abs().pow(2).numpy()
    got = F.stft_power(x.numpy(), n_fft, hop, win, F.hann_window(win))
    assert got.shape == ref.shape == (2, 1 + n_fft // 2, 1 + L // hop)
    assert np.abs(got - ref).max() <= 1e-9 * ref.max()
    assert np.abs(F.hann_window(win) - w.numpy()).max() < 1e-15


@pytest.mark.parametrize("n_mels,n_fft,sr", [(80, 512, 16000), (64, 512, 16000), (40, 400, 8000)])
def test_slaney_filterbank_matches_torchaudio(n_mels, n_fft, sr):
    ta = pytest.importorskip("torchaudio")
    ref = ta.functional.melscale_fbanks(1 + n_fft // 2, 0.0, sr / 2.0, n_mels, sr, norm="slaney", mel_scale="slaney")
    got = F.slaney_mel_filterbank(sr, n_fft, n_mels)
    assert got.dtype == np.float32 and got.shape == (n_mels, 1 + n_fft // 2)
    assert np.abs(got - ref.numpy().T).max() < 1e-5 * np.abs(got).max()      # torchaudio evaluates the ramps in fp32
    assert (got >= 0).all() and (got.sum(1) > 0).all()


def test_filterbank_features_against_torch_pipeline():
    """The whole of FilterbankFeatures.forward re-derived with torch ops (the reference's own call sequence,
    features.py:131-164, with the filterbank from the oracle)."""
    g = torch.Generator().manual_seed(3)
    x = torch.randn(3, 4800, generator=g)
    win, hop, n_fft, n_filt = 400, 200, 512, 80
    xp = torch.cat([x[:, :1], x[:, 1:] - 0.97 * x[:, :-1]], dim=1)
    s = torch.stft(xp, n_fft=n_fft, hop_length=hop, win_length=win, window=torch.hann_window(win, periodic=False),
                   return_complex=True)
    p = s.real.pow(2) + s.imag.pow(2)
    fb = torch.tensor(F.slaney_mel_filterbank(16000, n_fft, n_filt)).unsqueeze(0)
    y = torch.log(torch.matmul(fb, p) + 1e-20)
    seq_len = int(np.ceil(4800 / hop))
    y[:, :, seq_len:] = 0
    got = F.filterbank_features(x.numpy(), win_length=win, hop_length=hop, n_fft=n_fft, n_filt=n_filt)
    assert got.shape == (3, 80, 25)
    assert np.abs(got - y.numpy()).max() < 2e-4
    assert (got[:, :, 24] == 0).all()                           # L % hop == 0: the last centred frame is masked


def test_downsample_matches_reference_reshape():
    feat = np.arange(2 * 5 * 7, dtype=np.float32).reshape(2, 5, 7)
    d = F.downsample(feat, 3)
    assert d.shape == (2, 15, 3)
    t = torch.tensor(feat).transpose(1, 2)
    t = torch.nn.functional.pad(t, [0, 0, 0, 2, 0, 0]).reshape(2, -1, 15).transpose(1, 2)
    assert (d == t.numpy()).all()
    assert F.downsample(feat, 3, pad_to_divisible=False).shape == (2, 15, 2)
    x = np.random.default_rng(0).standard_normal((2, 3000)).astype(np.float32)
    assert F.logmel_frontend(x).shape == (2, 6, 240)
