"""Multi-process host logic on the CPU (gloo, world_size 2): shard plans, the padded in-place
all-gather used for positions, and that sharded counter-based noise equals the unsharded stream.
No kernel is launched here; the GPU side of the same path is in test_gpu_parity.py / bench.py."""
from __future__ import annotations

import os
import socket

import numpy as np
import pytest

from turtlemd_b200.parallel import ShardPlan, replica_shard


def test_shard_plan_covers_everything_once():
    for n, world in ((1000188, 8), (7, 2), (5, 8), (0, 4), (64, 1), (1000188, 3)):
        plans = [ShardPlan(n, world, r) for r in range(world)]
        assert sum(p.count for p in plans) == n
        assert plans[0].padded >= n and plans[0].padded % world == 0
        covered = np.zeros(n, dtype=int)
        for p in plans:
            assert 0 <= p.first <= n and p.count <= p.shard
            covered[p.first: p.first + p.count] += 1
            assert p.bounds(p.rank) == (p.first, p.count)
        assert np.all(covered == 1)
        assert plans[0].all_counts() == [p.count for p in plans]
    assert replica_shard(1000, 4, 2) == (2000, 4000)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, dim, queue):
    import torch
    import torch.distributed as dist

    from turtlemd_b200.parallel import ShardPlan, all_gather_padded
    from turtlemd_b200.philox import PhiloxNormal

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        plan = ShardPlan(n, world, rank)
        # positions: every rank writes f(global index) into its own rows, then the in-place gather
        buf = torch.full((plan.padded * dim,), -1.0, dtype=torch.float64)
        own = torch.arange(plan.first * dim, (plan.first + plan.count) * dim, dtype=torch.float64) * 0.5 + 3.0
        buf[plan.first * dim: (plan.first + plan.count) * dim] = own
        all_gather_padded(buf, plan, dim)
        expect = torch.arange(n * dim, dtype=torch.float64) * 0.5 + 3.0
        ok_gather = bool(torch.equal(buf[: n * dim], expect))
        # noise: draws (first+i)*2*dim .. of step s come from the same stream whatever the sharding
        twin = PhiloxNormal(key=99, position=1000)
        mine = twin.peek(plan.count * 2 * dim, offset=plan.first * 2 * dim)
        parts = [None] * world
        dist.all_gather_object(parts, mine)
        whole = PhiloxNormal(key=99, position=1000).peek(n * 2 * dim)
        ok_noise = bool(np.array_equal(np.concatenate(parts), whole))
        # energy shares: all-reduce of per-rank (V, W) buffers
        vw = torch.full((10,), float(rank + 1), dtype=torch.float64)
        dist.all_reduce(vw)
        ok_reduce = bool(torch.all(vw == world * (world + 1) / 2))
        queue.put((rank, ok_gather, ok_noise, ok_reduce))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n,dim", [(1001, 3), (6, 1)])
def test_two_rank_gather_noise_and_reduce(n, dim):
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    queue = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, dim, queue)) for r in range(2)]
    for p in procs:
        p.start()
    results = [queue.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(r[0] for r in results) == [0, 1]
    assert all(r[1] and r[2] and r[3] for r in results), results
