"""Shared helpers for the tests (golden loading, seeded E4D1 inputs)."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_tiny():
    z = np.load(os.path.join(GOLDEN, "tiny.npz"))
    cfg = {k[4:]: int(z[k]) for k in z.files if k.startswith("cfg_")}
    sd = {k[3:]: z[k] for k in z.files if k.startswith("sd.")}
    pg = {k[6:]: z[k] for k in z.files if k.startswith("pgrad.")}
    return z, cfg, sd, pg


def load_e4d1():
    return np.load(os.path.join(GOLDEN, "e4d1.npz"))


E4D1_CFG = dict(vocab_embed_size=64, vocab_size=1024, input_size=240, enc_hidden_size=320,
                enc_layers=4, enc_dropout=0, enc_proj_size=320, dec_hidden_size=320, dec_layers=1,
                dec_dropout=0, dec_proj_size=320, joint_size=320)


def e4d1_inputs():
    """SURVEY.md 8(d) recipe: torch.manual_seed(0); randn(2,200,240); randint(4,1024,(2,40))."""
    torch.manual_seed(0)
    xs = torch.randn(2, 200, 240)
    ys = torch.randint(4, 1024, (2, 40), dtype=torch.int32)
    return xs, ys


def to_t(sd, dtype=torch.float32, device="cpu"):
    return {k: torch.as_tensor(v).to(device=device, dtype=dtype) for k, v in sd.items()}


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-30))
