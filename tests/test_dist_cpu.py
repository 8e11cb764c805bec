"""N>1 host logic on CPU (gloo, world_size 2): batch sharding, the single flat-bucket all-reduce,
max-over-ranks timing reduction -- the multi-GPU path of SURVEY.md 8(e) minus the kernels."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from edgedict_b200 import dist as ed


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    r, w, _ = ed.init_from_env("gloo")
    assert (r, w) == (rank, world)
    s, e = ed.shard_batch(7, rank, world)
    bucket = torch.arange(10, dtype=torch.float32) * (rank + 1)          # "gradients" of this rank
    ed.allreduce_bucket(bucket, world, average=True)
    params = torch.full((4,), float(rank))
    ed.broadcast_bucket(params, src=0)
    mx = ed.max_over_ranks(10.0 + rank, torch.device("cpu"))
    q.put((rank, (s, e), bucket.tolist(), params.tolist(), mx))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_bucket_allreduce_and_sharding():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, s0, b0, p0, m0), (r1, s1, b1, p1, m1) = res
    assert s0 == (0, 4) and s1 == (4, 7)                                   # 7 utterances -> 4 + 3
    want = [i * 1.5 for i in range(10)]                                    # mean of 1x and 2x
    assert b0 == want and b1 == want
    assert p0 == [0.0] * 4 and p1 == [0.0] * 4                             # rank-0 weights everywhere
    assert m0 == 11.0 and m1 == 11.0


def test_shard_batch_covers_everything():
    for gb in (1, 7, 32, 256):
        for world in (1, 2, 4, 8):
            spans = [ed.shard_batch(gb, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == gb
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
