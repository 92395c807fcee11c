"""world_size-2 gloo test of the multi-GPU host logic (sharding + the final moment all-reduce)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ip_mcmc_b200 import parallel, stats


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, data, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n_chains, d = data.shape[0], data.shape[2]
    lo, hi = parallel.shard(n_chains, rank, world)
    mine = data[lo:hi].reshape(-1, d)
    n = float(mine.shape[0])
    mean = mine.mean(0)
    m2 = ((mine - mean) ** 2).sum(0)
    counters = np.array([n, n / 2, 10.0 * (rank + 1), 1.0, 0.0, 0.0])
    pooled = torch.tensor(np.concatenate([[n], mean, m2, counters]))
    out = parallel.allreduce_pooled(pooled, d)
    ms = parallel.max_over_ranks(10.0 + rank, "cpu")
    if rank == 0:
        ret.put((out.numpy(), ms))
    dist.destroy_process_group()


def test_shard_covers_all_chains():
    for n, w in ((10, 3), (8, 8), (5, 8), (65536, 8)):
        blocks = [parallel.shard(n, r, w) for r in range(w)]
        assert blocks[0][0] == 0 and blocks[-1][1] == n
        assert all(blocks[i][1] == blocks[i + 1][0] for i in range(w - 1))
        sizes = [b - a for a, b in blocks]
        assert max(sizes) - min(sizes) <= 1


def test_allreduce_pooled_two_ranks():
    rng = np.random.default_rng(0)
    data = rng.standard_normal((7, 50, 3)) * [1, 2, 3] + [100, -5, 0.1]     # 7 chains: uneven shards
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, data, ret)) for r in range(2)]
    for p in procs:
        p.start()
    out, ms = ret.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    flat = data.reshape(-1, 3)
    assert out[0] == flat.shape[0]
    np.testing.assert_allclose(out[1:4], flat.mean(0), rtol=1e-13)
    np.testing.assert_allclose(out[4:7] / (out[0] - 1), flat.var(0, ddof=1), rtol=1e-12)
    assert out[7] == 350 and out[9] == 30.0 and ms == 11.0


def test_single_process_passthrough():
    p = torch.arange(13, dtype=torch.float64)
    assert torch.equal(parallel.allreduce_pooled(p, 3), p)
