"""world_size-2 gloo test of the N>1 host logic (sharding + the single all-gather), CPU only."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from flowchain_iccv2023_b200 import dist as fdist


def test_shard_ranges_cover_exactly():
    for n in (0, 1, 7, 64, 4096):
        for w in (1, 2, 3, 8):
            r = [fdist.shard_range(n, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[k][1] == r[k + 1][0] for k in range(w - 1))
            assert max(b - a for a, b in r) - min(b - a for a, b in r) <= 1
    assert fdist.plan_shards(64, 65536, 8)[0] == "agents"
    assert fdist.plan_shards(1, 10000, 2)[0] == "points"


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class FakeFlow:
    """Stands in for the CUDA flow: a deterministic per-(agent, point, step) function, so the test checks
    only the sharding/gather plumbing (the density kernels need a GPU)."""

    def grid_log_prob(self, base_pos, cond, grid, cond_traj=None, map_layout=False, **kw):
        a = base_pos[:, 0][:, None, None] * 100.0
        t = torch.arange(cond.shape[1], dtype=torch.float32)[None, :, None]
        g = grid[:, 0][None, None, :] + 10.0 * grid[:, 1][None, None, :]
        return (a + t + g).contiguous()


def _worker(rank, world, port, A, G, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        base_pos = torch.randn(A, 2)
        cond = torch.randn(A, 12, 16)
        grid = torch.randn(G, 2)
        full = FakeFlow().grid_log_prob(base_pos, cond, grid)
        out = fdist.sharded_grid_log_prob(FakeFlow(), base_pos, cond, grid)
        q.put((rank, bool(torch.equal(out, full)), tuple(out.shape)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("A,G", [(5, 9), (1, 11), (4, 6)])
def test_sharded_maps_world2(A, G):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, A, G, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok, shape in res:
        assert ok and shape == (A, 12, G), (rank, ok, shape)
