"""Flat-bucket optimizer state for data-parallel training ("next" row N3).

All parameters of a module are re-homed into ONE contiguous fp32 buffer (``flat_params``) and
their gradients into another (``flat_grads``): the gradient all-reduce is a single NCCL call on
``flat_grads`` and the Adam update is a single launch of ``eb_adam_step`` over the bucket
(replaces torch.optim.Adam in cli/baseline.py:141-156,239-245; clip_grad_norm_ via ``eb_sumsq``).
"""
import math

import torch

from . import ops


class FlatAdam:
    """torch.optim.Adam over the flat bucket.  ``step(grad_scale, lr, max_norm)``: grad_scale multiplies every
    gradient (1 / loss_scale, 1 / accumulation steps); max_norm applies ``clip_grad_norm_`` semantics on the
    device (cli/baseline.py:239-245) with no host read of the norm; with clipping or loss scaling active a
    non-finite gradient norm skips the update (apex's overflow rule).  ``adamw=True`` selects the reference's own
    AdamW update (modules/optimizer.py:195-292)."""

    def __init__(self, module, lr=5e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, adamw=False):
        params = [p for p in module.parameters() if p.requires_grad]
        if not params:
            raise ValueError("no parameters")
        dev = params[0].device
        if dev.type != "cuda":
            raise RuntimeError("FlatAdam needs the module on a CUDA device")
        # every parameter starts on a 16-byte boundary of the bucket (vector loads, TMA-friendly)
        offs, n = [], 0
        for p in params:
            offs.append(n)
            n += (p.numel() + 3) // 4 * 4
        self.n = n
        self.flat_params = torch.zeros(n, dtype=torch.float32, device=dev)
        self.flat_grads = torch.zeros(n, dtype=torch.float32, device=dev)
        self.m = torch.zeros(n, dtype=torch.float32, device=dev)
        self.v = torch.zeros(n, dtype=torch.float32, device=dev)
        with torch.no_grad():
            for p, off in zip(params, offs):
                k = p.numel()
                self.flat_params[off:off + k].copy_(p.reshape(-1))
                p.data = self.flat_params[off:off + k].view(p.shape)
                p.grad = self.flat_grads[off:off + k].view(p.shape)
        self.params = params
        self.lr, self.betas, self.eps, self.weight_decay, self.adamw = lr, betas, eps, weight_decay, adamw
        self.step_count = 0
        self._norm = torch.zeros(1, dtype=torch.float32, device=dev)

    def zero_grad(self):
        self.flat_grads.zero_()
        for p in self.params:                       # autograd may have replaced .grad: re-home it
            if p.grad is None or p.grad.data_ptr() < self.flat_grads.data_ptr() or \
                    p.grad.data_ptr() >= self.flat_grads.data_ptr() + 4 * self.n:
                raise RuntimeError("parameter gradient left the flat bucket")

    def grad_norm(self):
        self._norm.zero_()
        ops.sumsq(self.flat_grads, self._norm)
        return self._norm.sqrt()

    def step(self, grad_scale=1.0, lr=None, max_norm=None, check_overflow=False):
        self.step_count += 1
        lr = self.lr if lr is None else lr
        if max_norm or check_overflow or self.adamw:
            ss = None
            if max_norm or check_overflow:
                self._norm.zero_()
                ops.sumsq(self.flat_grads, self._norm)
                ss = self._norm
            ops.adam_step_ex(self.flat_params, self.flat_grads, self.m, self.v, lr, self.betas[0], self.betas[1],
                             self.eps, self.weight_decay, self.step_count, grad_scale, ss, max_norm or 0.0, self.adamw)
        else:
            ops.adam_step(self.flat_params, self.flat_grads, self.m, self.v, lr, self.betas[0], self.betas[1],
                          self.eps, self.weight_decay, self.step_count, grad_scale)


class FlatAdamW(FlatAdam):
    """The reference's AdamW (modules/optimizer.py:195-292) over the flat bucket."""

    def __init__(self, module, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        super().__init__(module, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, adamw=True)
