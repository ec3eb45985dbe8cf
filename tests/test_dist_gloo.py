"""CPU, world_size = 2 over gloo: chain sharding and the observable gather (the path's only collective)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dmrg_b200.batched import gather_observables, shard


def test_shard_covers_everything():
    for n, w in ((4096, 8), (10, 3), (5, 8)):
        blocks = [shard(n, r, w) for r in range(w)]
        assert blocks[0][0] == 0 and blocks[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(blocks[:-1], blocks[1:]))
        assert max(b - a for a, b in blocks) - min(b - a for a, b in blocks) <= 1


def _worker(rank, world, port, n_total, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard(n_total, rank, world)
    local = torch.arange(lo, hi, dtype=torch.float64)[:, None] * torch.tensor([[1.0, -2.0, 0.5]], dtype=torch.float64)
    full = gather_observables(local)
    q.put((rank, full.numpy()))
    dist.barrier()
    dist.destroy_process_group()


def test_gather_observables_world2_ragged():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    n_total = 7                                           # ragged: 4 + 3
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_total, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    want = np.arange(n_total)[:, None] * np.array([[1.0, -2.0, 0.5]])
    for r in range(2):
        np.testing.assert_array_equal(got[r], want)
