"""CPU, world_size 2 over gloo: the N>1 plumbing of the batch-sharded path (no data-path collective)."""
import os
import socket

import torch
import torch.multiprocessing as mp

from bcos_b200.utils import dist as D


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    r, lr, w = D.init("gloo")
    assert (r, w) == (rank, world)
    lo, hi = D.shard_range(10, r, w)
    mine = torch.arange(lo, hi, dtype=torch.float32) * 2.0          # "results" of this rank's images
    D.barrier()
    slowest = D.max_over_ranks(10.0 + rank)
    parts = D.gather_to_rank0(mine)
    if rank == 0:
        q.put((slowest, torch.cat(parts).tolist()))
    D.shutdown()


def test_two_rank_sharding_and_timing_reduction():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    slowest, gathered = q.get(timeout=120)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert slowest == 11.0
    assert gathered == [2.0 * i for i in range(10)]


def test_shard_range_covers_everything():
    for total in (1, 7, 256, 257):
        for world in (1, 2, 4, 8):
            spans = [D.shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
