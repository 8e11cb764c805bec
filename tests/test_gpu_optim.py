"""the flat-bucket optimizer (edgedict_b200/optim.py) against torch.optim.Adam +
torch.nn.utils.clip_grad_norm_ (cli/baseline.py:141-156,239-245) and against the reference's AdamW update
(modules/optimizer.py:283-290, restated below)."""
import math

import pytest
import torch

from tests.util import rel_err

pytestmark = pytest.mark.gpu


def _net():
    torch.manual_seed(2)
    return torch.nn.Sequential(torch.nn.Linear(13, 7), torch.nn.Tanh(), torch.nn.Linear(7, 5))   # odd sizes: padded bucket


def _grads(step):
    g = torch.Generator().manual_seed(100 + step)
    return [torch.randn(7, 13, generator=g) * 3, torch.randn(7, generator=g), torch.randn(5, 7, generator=g) * 3,
            torch.randn(5, generator=g)]


@pytest.mark.parametrize("max_norm,wd", [(None, 0.0), (1.5, 0.0), (100.0, 1e-2)])
def test_flat_adam_bucket_clip_matches_torch(max_norm, wd):
    from edgedict_b200.optim import FlatAdam
    ref = _net()
    ours = _net().cuda()
    topt = torch.optim.Adam(ref.parameters(), lr=1e-2, weight_decay=wd)
    opt = FlatAdam(ours, lr=1e-2, weight_decay=wd)
    # re-homing: parameters and gradients are views of the two flat buckets
    lo, hi = opt.flat_params.data_ptr(), opt.flat_params.data_ptr() + 4 * opt.n
    assert all(lo <= p.data_ptr() < hi and p.data_ptr() % 16 == 0 for p in ours.parameters())
    for step in range(1, 5):
        opt.zero_grad()
        for p, q, g in zip(ref.parameters(), ours.parameters(), _grads(step)):
            p.grad = g.clone()
            q.grad.add_(g.cuda())                     # what autograd's accumulation does
        want_norm = math.sqrt(sum(float(g.double().pow(2).sum()) for g in _grads(step)))
        assert abs(float(opt.grad_norm()) - want_norm) < 1e-4 * want_norm
        if max_norm:
            torch.nn.utils.clip_grad_norm_(ref.parameters(), max_norm)
        topt.step()
        opt.step(max_norm=max_norm)
        for p, q in zip(ref.parameters(), ours.parameters()):
            assert rel_err(q.detach().cpu(), p.detach()) < 2e-5, step
    # loss-scale overflow: a non-finite gradient skips the update
    before = opt.flat_params.clone()
    list(ours.parameters())[0].grad[0, 0] = float("inf")
    opt.step(grad_scale=1.0 / 1024, check_overflow=True)
    assert torch.equal(before, opt.flat_params)


def test_flat_adamw_matches_reference_formula():
    from edgedict_b200.optim import FlatAdamW
    ours = _net().cuda()
    ref = [p.detach().clone().double() for p in _net().parameters()]
    lr, b1, b2, eps, wd = 3e-3, 0.9, 0.999, 1e-8, 1e-2
    opt = FlatAdamW(ours, lr=lr, betas=(b1, b2), eps=eps, weight_decay=wd)
    m = [torch.zeros_like(p) for p in ref]
    v = [torch.zeros_like(p) for p in ref]
    for step in range(1, 5):
        opt.zero_grad()
        gs = _grads(step)
        for q, g in zip(ours.parameters(), gs):
            q.grad.add_(g.cuda())
        opt.step()
        for p, mi, vi, g in zip(ref, m, v, gs):       # modules/optimizer.py:275-290
            g = g.double()
            mi.mul_(b1).add_(g, alpha=1 - b1)
            vi.mul_(b2).addcmul_(g, g, value=1 - b2)
            denom = vi.sqrt().add_(eps)
            step_size = lr * math.sqrt(1 - b2 ** step) / (1 - b1 ** step)
            p.add_(torch.mul(p, wd).addcdiv_(mi, denom), alpha=-step_size)
        for p, q in zip(ref, ours.parameters()):
            assert rel_err(q.detach().cpu(), p) < 2e-5, step
