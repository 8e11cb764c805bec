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


def test_wavefront_plan_and_chunk_major_buffers():
    """Host logic of the layer-wavefront schedule (functional.wavefront_plan / _Chunks): chunk boundaries never
    split a TimeReduction pair, lengths add up on every layer, scatter/gather are inverse permutations."""
    from edgedict_b200 import functional as Fn
    for T, red in [(1000, [0, 1, 0, 0, 0, 0]), (999, [0, 1, 0]), (37, [0, 1]), (133, [1, 0, 1, 0]), (64, [0, 0])]:
        plan = Fn.wavefront_plan(T, [bool(r) for r in red], max_chunks=4)
        if plan is None:
            assert T < 2 * 16 * (1 << sum(red)) or T <= 16 * (1 << sum(red))
            continue
        assert len(plan) == len(red) + 1 and sum(plan[0]) == T and 2 <= len(plan[0]) <= 4 + 1
        gran = 1 << sum(red)
        assert all(n % gran == 0 for n in plan[0][:-1])                   # only the last chunk may be ragged
        Tl = T
        for l, r in enumerate(red):
            assert sum(plan[l]) == Tl
            Tl = (Tl + 1) // 2 if r else Tl
            assert plan[l + 1] == ([(n + 1) // 2 for n in plan[l]] if r else plan[l])
        assert sum(plan[-1]) == Tl                                         # == the unchunked output length
    assert Fn.wavefront_plan(1000, [False, True], max_chunks=0) is None    # schedule disabled
    assert Fn.wavefront_plan(20, [False, True]) is None                    # too short to cut
    ck = Fn._Chunks(3, [4, 4, 2])
    x = torch.arange(3 * 10 * 5, dtype=torch.float32).reshape(3, 10, 5)
    flat = ck.scatter(x)
    assert flat.shape == (30, 5) and torch.equal(ck.blk(flat, 1), x[:, 4:8]) and torch.equal(ck.gather(flat), x)
    v = ck.new(0, torch.float32, "cpu")
    assert v.shape == (30,) and ck.blk(v, 2).shape == (6,)
