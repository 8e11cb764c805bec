"""GPU parity of the RNN-T loss kernels, called through the C-ABI, against the oracle
(oracle/rnnt_loss_oracle.c) and the reference's known-answer vectors."""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle import loss as ol
from tests.golden import loss_kat as K

pytestmark = pytest.mark.gpu


def _dev(a, dtype):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=dtype).cuda()


def _compat_call(acts, labels, tl, ul, blank=0, want_grads=True, dtype=np.float32):
    """Drive compute_rnnt_loss exactly like pytorch_binding/src/binding.cpp:84-154 does:
    device acts/labels/lengths/workspace, host costs, options struct by value."""
    from edgedict_b200._lib import lib
    L = lib()
    tdt = torch.float32 if dtype == np.float32 else torch.float64
    a = _dev(acts, tdt)
    B, T, U, V = a.shape
    lab, xl, yl = _dev(labels, torch.int32), _dev(tl, torch.int32), _dev(ul, torch.int32)
    sz = C.c_size_t(0)
    assert L.get_workspace_size(T, U, B, True, C.byref(sz), a.element_size()) == 0
    ws = torch.empty(sz.value, dtype=torch.uint8, device="cuda")
    grads = torch.full_like(a, 7.0) if want_grads else None      # poison: must be fully overwritten
    costs = np.zeros(B, dtype=dtype)
    opt = ol.RnntOptions(loc=1, num_threads=0, stream=torch.cuda.current_stream().cuda_stream,
                         blank_label=blank, maxT=T, maxU=U, batch_first=True)
    fn = L.compute_rnnt_loss if dtype == np.float32 else L.compute_rnnt_loss_fp64
    st = fn(C.c_void_p(a.data_ptr()), C.c_void_p(grads.data_ptr()) if want_grads else None,
            C.c_void_p(lab.data_ptr()), C.c_void_p(yl.data_ptr()), C.c_void_p(xl.data_ptr()), C.c_int(V), C.c_int(B),
            costs.ctypes.data_as(C.c_void_p), C.c_void_p(ws.data_ptr()), opt)
    assert st == 0, st
    return costs, (grads.cpu().numpy() if want_grads else None)


def test_kat_small_through_compat_abi():
    costs, grads = _compat_call(K.SMALL_ACTS, K.SMALL_LABELS, [2], [2])
    assert np.allclose(costs[0], K.SMALL_COST, rtol=1e-6)
    assert np.allclose(grads, K.SMALL_GRADS_LOGITS, atol=1e-6)


def test_kat_big_through_compat_abi_fp32_and_fp64():
    for dt, tol in ((np.float32, 1e-6), (np.float64, 1e-9)):
        costs, grads = _compat_call(K.BIG_ACTS, K.BIG_LABELS, [4, 4], [2, 2], dtype=dt)
        assert np.allclose(costs, K.BIG_COSTS, rtol=tol * 10)
        assert np.allclose(grads, K.BIG_GRADS_LOGITS, rtol=1e-3, atol=1e-6)


def test_compat_abi_errors():
    from edgedict_b200._lib import lib
    L = lib()
    opt = ol.RnntOptions(loc=0, num_threads=1, stream=None, blank_label=0, maxT=2, maxU=3, batch_first=True)
    a = torch.zeros(1, 2, 3, 5, device="cuda")
    i = torch.zeros(4, dtype=torch.int32, device="cuda")
    costs = np.zeros(1, np.float32)
    p = lambda t: C.c_void_p(t.data_ptr())
    # CPU location is not implemented in this build: must fail, not fall back
    assert L.compute_rnnt_loss(p(a), None, p(i), p(i), p(i), 5, 1, costs.ctypes.data_as(C.c_void_p), p(a), opt) == 3
    opt.loc = 1
    assert L.compute_rnnt_loss(None, None, p(i), p(i), p(i), 5, 1, costs.ctypes.data_as(C.c_void_p), p(a), opt) == 2


def test_forward_only_score():
    costs, grads = _compat_call(K.BIG_ACTS, K.BIG_LABELS, [4, 4], [2, 2], want_grads=False)
    assert grads is None and np.allclose(costs, K.BIG_COSTS, rtol=1e-5)


@pytest.mark.parametrize("B,T,U,V", [(1, 1, 1, 2), (3, 17, 6, 11), (2, 50, 16, 20), (4, 33, 9, 1024),
                                     (2, 7, 40, 130), (5, 10, 6, 5), (1, 50, 10, 15), (2, 64, 129, 64)])
def test_random_ragged_vs_oracle(B, T, U, V):
    rng = np.random.RandomState(B * 1000 + T + U)
    acts = (rng.randn(B, T, U, V) * 2).astype(np.float32)
    labels = rng.randint(1, V, size=(B, max(U - 1, 0))).astype(np.int32)     # U == 1: no labels at all
    tl = np.full(B, T, np.int32)
    ul = np.full(B, U - 1, np.int32)
    if B > 1:
        tl[1:] = rng.randint(1, T + 1, size=B - 1)
        ul[1:] = rng.randint(0, U, size=B - 1)
    c_o, g_o = ol.logits(acts, labels.reshape(B, U - 1), tl, ul, dtype=np.float64)
    c, g = _compat_call(acts, labels.reshape(B, U - 1), tl, ul)
    assert np.allclose(c, c_o, rtol=1e-5), (c, c_o)
    # fp32 lattice sums reach |alpha+beta| ~ 1e2, so one ulp of the exponent is ~1e-5 relative (the
    # reference's own fp32 CPU library deviates from fp64 by 3.5e-5 on these problems); |g| <= 1
    assert np.abs(g - g_o).max() < 1e-3
    pad = np.ones((B, T, U), bool)
    for b in range(B):
        pad[b, :tl[b], :ul[b] + 1] = False
    assert (g[pad] == 0).all()                         # padded cells are zero-filled


def test_inf_problem_is_finite():
    # warp-transducer/tests/test_gpu.cu:226-306
    rng = np.random.RandomState(0)
    acts = rng.uniform(0, 1, size=(1, 50, 10, 15)).astype(np.float32)
    labels = rng.randint(1, 15, size=(1, 9)).astype(np.int32)
    c, g = _compat_call(acts, labels, [50], [9])
    assert np.isfinite(c).all() and np.isfinite(g).all()


def test_module_api_reductions_and_backward_scaling():
    from edgedict_b200.warprnnt_pytorch import RNNTLoss
    rng = np.random.RandomState(5)
    B, T, U, V = 3, 9, 5, 12
    acts = rng.randn(B, T, U, V).astype(np.float32)
    labels = rng.randint(1, V, size=(B, U - 1)).astype(np.int32)
    tl = np.array([9, 7, 9], np.int32)
    ul = np.array([4, 4, 2], np.int32)
    c_o, g_o = ol.logits(acts, labels, tl, ul, dtype=np.float64)
    for red, cs, gs in (("none", c_o, g_o), ("sum", c_o.sum(), g_o), ("mean", c_o.sum() / B, g_o / B)):
        a = torch.tensor(acts, device="cuda", requires_grad=True)
        out = RNNTLoss(reduction=red)(a, _dev(labels, torch.int32), _dev(tl, torch.int32), _dev(ul, torch.int32))
        assert np.allclose(out.detach().cpu().numpy().reshape(-1), np.reshape(cs, -1), rtol=1e-5)
        (out.sum() * 2.5).backward()
        assert np.abs(a.grad.cpu().numpy() - 2.5 * gs).max() < 5e-5
    # per-utterance upstream gradients with reduction='none'
    a = torch.tensor(acts, device="cuda", requires_grad=True)
    out = RNNTLoss(reduction="none")(a, _dev(labels, torch.int32), _dev(tl, torch.int32), _dev(ul, torch.int32))
    w = torch.tensor([1.0, -2.0, 0.5], device="cuda")
    (out * w).sum().backward()
    assert np.abs(a.grad.cpu().numpy() - g_o * w.cpu().numpy()[:, None, None, None]).max() < 5e-5


def test_bf16_and_inplace_gradient_outputs():
    from edgedict_b200 import ops
    rng = np.random.RandomState(6)
    B, T, U, V = 2, 12, 7, 256
    acts = rng.randn(B, T, U, V).astype(np.float32)
    labels = rng.randint(1, V, size=(B, U - 1)).astype(np.int32)
    tl, ul = np.array([12, 10], np.int32), np.array([6, 3], np.int32)
    _, g_o = ol.logits(acts, labels, tl, ul, dtype=np.float64)
    a = _dev(acts, torch.float32)
    lab, xl, yl = _dev(labels, torch.int32), _dev(tl, torch.int32), _dev(ul, torch.int32)
    costs, ws = ops.rnnt_loss_fwd(a, lab, xl, yl, 0)
    g16 = ops.rnnt_loss_bwd(a, lab, xl, yl, 0, ws, None, 1.0, out_bf16=True)
    assert g16.dtype == torch.bfloat16
    assert np.abs(g16.float().cpu().numpy() - g_o).max() < 4e-3        # bf16 rounding of |g| <= 1
    ops.rnnt_loss_bwd(a, lab, xl, yl, 0, ws, None, 0.5, out=a)            # in place over the logits
    assert np.abs(a.cpu().numpy() - 0.5 * g_o).max() < 2e-5


def test_full_size_properties():
    """BASELINE size (B=32, T'=500, U+1=129, V=1024): the oracle would take minutes, so check the
    size-independent properties: forward and backward likelihoods agree (cpu_rnnt.h:167-170),
    every gradient row sums to zero (softmax shift invariance), padded cells are zero."""
    from edgedict_b200 import ops
    B, T, U, V = 32, 500, 129, 1024
    g = torch.Generator(device="cuda").manual_seed(1)
    a = torch.randn(B, T, U, V, device="cuda", generator=g)
    lab = torch.randint(1, V, (B, U - 1), device="cuda", dtype=torch.int32, generator=g)
    xl = torch.randint(T // 2, T + 1, (B,), device="cuda", dtype=torch.int32, generator=g)
    yl = torch.randint(U // 2, U, (B,), device="cuda", dtype=torch.int32, generator=g)
    xl[0], yl[0] = T, U - 1
    costs, ws = ops.rnnt_loss_fwd(a, lab, xl, yl, 0)
    n = B * T * U
    wsf = ws.view(torch.float32)
    ll_f, ll_b = wsf[5 * n:5 * n + B], wsf[5 * n + B:5 * n + 2 * B]
    assert torch.allclose(ll_f, ll_b, rtol=1e-5, atol=1e-2)
    assert torch.isfinite(costs).all() and (costs > 0).all()
    grads = ops.rnnt_loss_bwd(a, lab, xl, yl, 0, ws, None, 1.0)
    rs = grads.sum(-1)
    assert rs.abs().max() < 1e-3
    t_idx = torch.arange(T, device="cuda")[None, :, None]
    u_idx = torch.arange(U, device="cuda")[None, None, :]
    pad = (t_idx >= xl[:, None, None]) | (u_idx > yl[:, None, None])
    assert (grads[pad] == 0).all()
    # a sub-batch small enough for the oracle: first utterance truncated to 40 x 20 cells
    sub = a[:1, :40, :20].contiguous()
    c_s, _ = ops.rnnt_loss_fwd(sub, lab[:1, :19].contiguous(), torch.tensor([40], dtype=torch.int32, device="cuda"),
                               torch.tensor([19], dtype=torch.int32, device="cuda"), 0)
    c_o, _ = ol.logits(sub.cpu().numpy(), lab[:1, :19].cpu().numpy(), [40], [19], want_grads=False, dtype=np.float64)
    assert np.allclose(c_s.cpu().numpy(), c_o, rtol=1e-5)
