"""Multi-process host logic on the CPU (gloo, world_size 2): shard plans, the padded in-place
all-gather used for positions, and that sharded counter-based noise equals the unsharded stream.
No kernel is launched here; the GPU side of the same path is in test_gpu_parity.py / bench.py."""
from __future__ import annotations

import os
import socket

import numpy as np
import pytest

from turtlemd_b200.parallel import ShardPlan, halo_ranges, replica_shard, slab_layer_plan


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


# ---- slab decomposition: halo plan and exchange -----------------------------------------------------------


def test_halo_ranges_form_a_ring():
    """What a rank sends up is what its upper neighbour receives from below, at the same slots;
    received ranges lie outside the own slab; the ring closes across the periodic boundary."""
    for n, world, halo in ((1000188, 8, 58000), (10976, 3, 3000), (4000, 2, 1500), (64, 4, 16)):
        plans = [ShardPlan(n, world, r) for r in range(world)]
        ranges = [halo_ranges(p, halo) for p in plans]
        for r, (p, mine) in enumerate(zip(plans, ranges)):
            up, down = (r + 1) % world, (r - 1) % world
            assert mine["send_up"] == (up, p.first + p.count - halo, p.first + p.count)
            assert mine["send_down"] == (down, p.first, p.first + halo)
            assert ranges[up]["recv_down"] == (r,) + mine["send_up"][1:]
            assert ranges[down]["recv_up"] == (r,) + mine["send_down"][1:]
            for name in ("recv_up", "recv_down"):
                _, a, b = mine[name]
                assert b - a == halo and 0 <= a and b <= n
                assert b <= p.first or a >= p.first + p.count
    with pytest.raises(NotImplementedError):
        halo_ranges(ShardPlan(100, 4, 0), 26)  # slabs of 25 slots


def _halo_worker(rank, world, port, n, halo, queue):
    import torch
    import torch.distributed as dist

    from turtlemd_b200.parallel import ShardPlan, halo_exchange, halo_ranges

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        plan = ShardPlan(n, world, rank)
        truth = torch.arange(plan.padded * 4, dtype=torch.float64).reshape(plan.padded, 4) * 0.25 + 1.0
        records = torch.full((plan.padded, 4), -7.0, dtype=torch.float64)
        records[plan.first: plan.first + plan.count] = truth[plan.first: plan.first + plan.count]
        halo_exchange(records, plan, halo)
        ok = True
        for name in ("recv_up", "recv_down"):
            _, a, b = halo_ranges(plan, halo)[name]
            ok = ok and bool(torch.equal(records[a:b], truth[a:b]))
        ok = ok and bool(torch.equal(records[plan.first: plan.first + plan.count],
                                     truth[plan.first: plan.first + plan.count]))
        # nothing else was touched
        mask = torch.ones(plan.padded, dtype=torch.bool)
        mask[plan.first: plan.first + plan.count] = False
        for name in ("recv_up", "recv_down"):
            _, a, b = halo_ranges(plan, halo)[name]
            mask[a:b] = False
        ok = ok and bool(torch.all(records[mask] == -7.0))
        # the rebuild decision: OR of the per-rank displacement flags
        flag = torch.tensor([1 if rank == world - 1 else 0], dtype=torch.int32)
        dist.all_reduce(flag, op=dist.ReduceOp.MAX)
        queue.put((rank, ok, int(flag.item())))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n,halo", [(1001, 200), (64, 32)])
def test_two_rank_halo_exchange(n, halo):
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    queue = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_halo_worker, args=(r, 2, port, n, halo, queue)) for r in range(2)]
    for p in procs:
        p.start()
    results = [queue.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(r[1] and r[2] == 1 for r in results), results


# ---- resident slab run: layer plan and the all-ranks verdict ------------------------------------------------------


def test_slab_layer_plan_invariants():
    """Every layer is owned exactly once, ghost layers are the periodic neighbours of a slab and never own layers, the
    two ghost layers of a rank differ; the benchmark systems: 38 layers -> 32 on eight ranks, 36 on four, 38 on two."""
    for layers, world in ((38, 8), (38, 4), (38, 2), (24, 8), (14, 2), (6, 3), (7, 2), (4, 2)):
        plan = slab_layer_plan(layers, world)
        nz = (layers // world) * world
        owned = np.zeros(nz, dtype=int)
        for z0, z1, below, above in plan:
            assert z1 - z0 == nz // world >= 2
            owned[z0:z1] += 1
            assert below == (z0 - 1) % nz and above == z1 % nz
            assert below != above and not (z0 <= below < z1) and not (z0 <= above < z1)
        assert np.all(owned == 1)
    assert [p[1] - p[0] for p in slab_layer_plan(38, 8)] == [4] * 8
    assert slab_layer_plan(38, 4)[-1][:2] == (27, 36) and slab_layer_plan(38, 2)[1] == (19, 38, 18, 0)
    for layers, world in ((15, 8), (3, 2), (38, 1)):
        with pytest.raises(NotImplementedError):
            slab_layer_plan(layers, world)


def _verdict_worker(rank, world, port, queue):
    import torch.distributed as dist

    from turtlemd_b200.parallel import all_ranks_agree

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        everyone_fine = all_ranks_agree(True)
        one_failed = all_ranks_agree(rank != 1)  # rank 1 could not map a peer's window: nobody may go on
        queue.put((rank, everyone_fine, one_failed))
    finally:
        dist.destroy_process_group()


def test_two_rank_verdict_is_the_same_everywhere():
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    queue = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_verdict_worker, args=(r, 2, port, queue)) for r in range(2)]
    for p in procs:
        p.start()
    results = [queue.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(r[1] is True and r[2] is False for r in results), results
