"""The NCCL leg of the data-parallel path on real GPUs: two ranks, flat-bucket all-reduce + broadcast, and
one optimizer step whose parameters stay identical across ranks.  Needs two visible GPUs (skipped on a 1-GPU box; the
host logic is covered with gloo in tests/test_dist_cpu.py)."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    import torch.distributed as dist
    from edgedict_b200 import dist as ed
    from edgedict_b200.optim import FlatAdam
    r, w, local = ed.init_from_env("nccl")
    torch.cuda.set_device(local)
    torch.manual_seed(7 + rank)                       # different weights per rank before the broadcast
    net = torch.nn.Linear(33, 17).cuda()
    opt = FlatAdam(net, lr=1e-2)
    ed.broadcast_bucket(opt.flat_params)
    opt.zero_grad()
    for p in net.parameters():
        p.grad.add_(float(rank + 1))                  # "gradients" of this rank
    ed.allreduce_bucket(opt.flat_grads, w)            # mean over ranks = 1.5
    g = opt.flat_grads.clone()
    opt.step()
    q.put((rank, float(g.min()), float(g[:33 * 17].max()), float(opt.flat_params.double().sum())))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_nccl_bucket_allreduce_and_step():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=300) for _ in range(world))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    (_, lo0, hi0, s0), (_, lo1, hi1, s1) = res
    assert abs(hi0 - 1.5) < 1e-6 and abs(hi1 - 1.5) < 1e-6 and lo0 >= 0.0      # padding slots of the bucket stay 0
    assert s0 == s1                                                             # replicas stay bit-identical
