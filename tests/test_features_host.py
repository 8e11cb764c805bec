"""Host-side checks of the front-end mirror (no GPU): tables, buffer names, loud failure without CUDA."""
import numpy as np
import pytest
import torch

from oracle import features_np as F


def test_tables_match_the_oracle_and_reference_buffer_names():
    from edgedict_b200.rnnt.features import FilterbankFeatures, mel_filterbank
    m = FilterbankFeatures(n_filt=80, win_length=400, hop_length=200)
    assert list(m.state_dict().keys()) == ["fb", "window"]               # rnnt/features.py:84-85
    assert tuple(m.fb.shape) == (1, 80, 257) and tuple(m.window.shape) == (400,)
    assert np.abs(m.fb[0].numpy() - F.slaney_mel_filterbank(16000, 512, 80)).max() < 1e-9
    assert np.abs(m.window.numpy() - F.hann_window(400)).max() < 1e-6
    # DFT basis: frame @ basis == rfft(frame * centred window)
    frame = np.random.default_rng(0).standard_normal(512)
    w = np.zeros(512); w[56:456] = F.hann_window(400)
    spec = np.fft.rfft(frame * w)
    got = frame @ m.dft_basis.numpy().astype(np.float64)
    assert np.abs(got[:257] - spec.real).max() < 1e-4 and np.abs(got[257:] - spec.imag).max() < 1e-4
    assert m.max_length % 16 == 0 or m.max_length > 0
    assert np.abs(mel_filterbank(8000, 256, 40) - F.slaney_mel_filterbank(8000, 256, 40)).max() < 1e-9


def test_front_end_has_no_cpu_path():
    from edgedict_b200.rnnt.features import FilterbankFeatures, Downsample
    m = FilterbankFeatures(dither=0)
    with pytest.raises(RuntimeError, match="CUDA"):
        m(torch.randn(1, 4000))
    d = Downsample(3)(torch.arange(2 * 5 * 7, dtype=torch.float32).reshape(2, 5, 7))   # pure reshape: device agnostic
    assert np.array_equal(d.numpy(), F.downsample(np.arange(70, dtype=np.float32).reshape(2, 5, 7), 3))
