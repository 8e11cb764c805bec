"""Data-parallel plumbing: one process per GPU (torchrun), replicated weights, the minibatch
sharded across ranks, ONE all-reduce per optimizer step on the flat gradient bucket.

Replaces torch.nn.DataParallel (cli/train.py:152-153, cli/baseline.py:158-159) and the
Lightning-DDP bucketed all-reduce (cli/lightning.py:325-332).  No other tensor crosses GPUs:
utterances are independent through encoder, predictor, joint and lattice.
"""
import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """Initialise torch.distributed from torchrun's environment; returns (rank, world, local_rank)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local


def shard_batch(global_batch, rank, world):
    """Contiguous, near-even split of ``global_batch`` utterances; returns (start, stop)."""
    base, rem = divmod(global_batch, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def allreduce_bucket(flat_grads, world=None, average=True):
    """Sum (or average) the flat gradient bucket across ranks in a single collective."""
    if not dist.is_initialized():
        return flat_grads
    world = world or dist.get_world_size()
    if world == 1:
        return flat_grads
    dist.all_reduce(flat_grads, op=dist.ReduceOp.SUM)
    if average:
        flat_grads.mul_(1.0 / world)
    return flat_grads


def broadcast_bucket(flat_params, src=0):
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.broadcast(flat_params, src=src)
    return flat_params


def max_over_ranks(value, device):
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
