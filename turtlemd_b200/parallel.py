"""Multi-GPU execution of the hot path: one process per GPU, ``torch.distributed`` for the plumbing.

The reference is single-process; this module is what the B200 build adds
(SURVEY.md section 8e):

* :class:`AtomDecomposition` -- one large Lennard-Jones system on G GPUs.  Every rank keeps the
  whole position array (it is what the pair kernels read), owns a contiguous block of atoms
  ``[first, first + count)``, integrates only those, evaluates forces only for those rows
  (``tmd_lj_* (first, count)``: no Newton's-third-law traffic between ranks) and, once per step,
  all-gathers the new positions over NVLink (NCCL, in place, equal padded shards).  Energy and
  virial are per-rank shares that are all-reduced only when somebody looks.  A single-GPU run is
  the yardstick: forces agree to summation order (<= 1e-12).
* :func:`replica_shard` -- independent replica ensembles (the infretis use case): rank r holds
  replicas ``[r*n, (r+1)*n)`` of a global ensemble, no communication at all; the noise of replica
  i depends only on its global index (counter-based stream), so any sharding gives the same
  trajectories.

:class:`ShardPlan` is pure host logic and is what the CPU (gloo) tests exercise.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _device, _lib
from .integrators import LangevinInertia, VelocityVerlet, _is_device_stream
from .potentials.lennardjones import LennardJonesCut

_ALL = _lib.WANT_V | _lib.WANT_F | _lib.WANT_W


@dataclass(frozen=True)
class ShardPlan:
    """Contiguous, equal (padded) blocks of ``n`` units over ``world`` ranks."""

    n: int
    world: int
    rank: int

    @property
    def shard(self) -> int:
        """Units per rank in the padded layout (the all-gather needs equal counts)."""
        return -(-self.n // self.world)

    @property
    def padded(self) -> int:
        return self.shard * self.world

    @property
    def first(self) -> int:
        return min(self.rank * self.shard, self.n)

    @property
    def count(self) -> int:
        return max(0, min(self.shard, self.n - self.first))

    def bounds(self, rank: int) -> tuple[int, int]:
        other = ShardPlan(self.n, self.world, rank)
        return other.first, other.count

    def all_counts(self) -> list[int]:
        return [self.bounds(r)[1] for r in range(self.world)]


def replica_shard(n_per_rank: int, world: int, rank: int) -> tuple[int, int]:
    """(first, n_global) of this rank's block in a weak-scaling replica ensemble."""
    return rank * n_per_rank, n_per_rank * world


def all_gather_padded(buffer, plan: ShardPlan, width: int, group=None) -> None:
    """In-place all-gather of ``buffer`` (flat, ``plan.padded * width`` elements): every rank has
    written its own shard ``[rank*shard*width, (rank+1)*shard*width)``; afterwards all ranks hold
    all shards.  Works on CPU tensors with gloo (tests) and on CUDA tensors with NCCL."""
    import torch.distributed as dist

    lo = plan.rank * plan.shard * width
    mine = buffer[lo: lo + plan.shard * width]
    if buffer.is_cuda:
        dist.all_gather_into_tensor(buffer, mine, group=group)
    else:  # gloo has no in-place flat variant for CPU tensors: go through a list of views
        parts = [buffer[r * plan.shard * width: (r + 1) * plan.shard * width] for r in range(plan.world)]
        dist.all_gather(parts, mine.clone(), group=group)


class AtomDecomposition:
    """One LJ (+ any other device potential) system stepped by ``world`` GPUs.

    ``system`` is the full system on every rank (same arrays everywhere); ``integrator`` is a
    :class:`VelocityVerlet` or a :class:`LangevinInertia` driven by the counter-based stream.
    """

    def __init__(self, system, integrator, group=None, plan: ShardPlan | None = None):
        import torch.distributed as dist

        if not isinstance(integrator, (VelocityVerlet, LangevinInertia)):
            raise NotImplementedError("AtomDecomposition supports VelocityVerlet and LangevinInertia")
        if isinstance(integrator, LangevinInertia) and not _is_device_stream(integrator.rgen):
            raise NotImplementedError("sharded Langevin dynamics needs the counter-based PhiloxNormal stream")
        self.system, self.integrator, self.group = system, integrator, group
        part = system.particles
        self.n, self.dim = part._pos.shape
        if plan is not None:  # explicit plan: tests emulate several ranks in one process and exchange by hand
            self.world, self.rank, self.plan = plan.world, plan.rank, plan
            self._collectives = False
        else:
            self.world = dist.get_world_size(group) if dist.is_initialized() else 1
            self.rank = dist.get_rank(group) if dist.is_initialized() else 0
            self.plan = ShardPlan(self.n, self.world, self.rank)
            self._collectives = self.world > 1
        self._pos_pad = None
        self._graph = None
        self._graph_launches = 0

    # ---- set-up -----------------------------------------------------------------------------------
    def prepare(self) -> None:
        """Upload the state, re-home positions in a padded buffer and evaluate the initial forces."""
        _device.require_cuda()
        part, plan, dim = self.system.particles, self.plan, self.dim
        pos = part._pos.device()
        pad = _device.zeros(plan.padded * dim)
        pad[: self.n * dim].copy_(pos)
        self._pos_pad = pad
        part._pos._dev = pad[: self.n * dim]  # the kernels and the Mirror see the first n rows
        part._pos.dev_valid, part._pos.host_valid = True, True
        part._vel.device()
        part._force.device_uninitialised(self.n * dim).zero_()
        part._force.written_on_device(shape=part._pos.shape)
        if isinstance(self.integrator, LangevinInertia):
            self.integrator.initiate_parameters(self.system)
        self._evaluate(_ALL)

    # ---- pieces of a step ---------------------------------------------------------------------------
    def _own(self, tensor):
        lo = self.plan.first * self.dim
        return tensor[lo: lo + self.plan.count * self.dim]

    def _evaluate(self, flags: int) -> None:
        """All potentials for this rank's rows; vw holds this rank's share of V and W."""
        part = self.system.particles
        pos, force, vw = part._pos.device(), part._force.device(), part._vw_buffer()
        first = True
        for pot in self.system.potentials:
            if not isinstance(pot, LennardJonesCut):
                raise NotImplementedError("AtomDecomposition shards LennardJonesCut force fields only")
            pot._launch(self.system, pos, force if flags & _lib.WANT_F else None, vw,
                   flags | (0 if first else _lib.ACCUMULATE), first=self.plan.first, count=self.plan.count)
            first = False
        if flags & _lib.WANT_F:
            part._force.written_on_device()

    def _gather_positions(self) -> None:
        if self._collectives:
            all_gather_padded(self._pos_pad, self.plan, self.dim, self.group)

    def step(self) -> None:
        """One integration step: move the own atoms, exchange positions, forces for the own rows."""
        if self._graph is not None:
            self._graph.replay()
            _lib.add_launches(self._graph_launches)
            part = self.system.particles
            part._pos.written_on_device()
            part._vel.written_on_device()
            part._force.written_on_device()
            return
        self.step_move()
        self._gather_positions()
        self.step_force()

    def capture(self, warm_steps: int = 3) -> bool:
        """Record one step (kernels + the NCCL all-gather) in a CUDA graph and replay it from now on:
        at 8 GPUs a step is ~100 us of device work and the Python / launch path would dominate.
        Only for VelocityVerlet (the Langevin kernels take the stream offset as a launch parameter).
        Returns False, leaving eager stepping in place, when capture is not possible."""
        if not isinstance(self.integrator, VelocityVerlet) or self._graph is not None:
            return self._graph is not None
        torch = _device.torch()
        for _ in range(warm_steps):  # allocations, occupancy queries and the first list build happen here
            self.step()
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        before = _lib.launch_count()
        try:
            with torch.cuda.graph(graph):
                self.step_move()
                self._gather_positions()
                self.step_force()
        except Exception:  # noqa: BLE001 - capture is an optimisation; eager stepping stays correct
            torch.cuda.synchronize()
            return False
        self._graph_launches = _lib.launch_count() - before
        _lib.add_launches(-self._graph_launches)  # recording is not launching
        self._graph = graph
        return True

    def _step_args(self):
        part, plan = self.system.particles, self.plan
        pos, vel = self._own(part._pos.device()), self._own(part._vel.device())
        force = self._own(part._force.device())
        imass = part._imass_device()[plan.first: plan.first + plan.count]
        return pos, vel, force, imass

    def step_move(self) -> None:
        """Everything before the exchange: the position update of this rank's atoms."""
        plan, dim, integ = self.plan, self.dim, self.integrator
        pos, vel, force, imass = self._step_args()
        stream = _device.stream_ptr()
        if isinstance(integ, VelocityVerlet):
            if plan.count:
                _lib.call("tmd_vv_kick_drift", _device.dptr(pos), _device.dptr(vel), _device.dptr(force),
                          _device.dptr(imass), plan.count, dim, float(integ.timestep), stream)
        else:
            rng = _lib.Rng(integ.rgen.key, integ.rgen.position)
            integ.rgen.advance(2 * self.n * dim)  # the stream moves on by the GLOBAL draw count on every rank
            if plan.count:
                _lib.call("tmd_langevin_inertia_pre", _device.dptr(pos), _device.dptr(vel), _device.dptr(force),
                          _device.dptr(imass), plan.count, dim, C.byref(integ._consts), C.byref(rng), plan.first,
                          None, None, stream)
        self.system.particles._pos.written_on_device()

    def step_force(self) -> None:
        """Everything after the exchange: forces for the own rows and the closing velocity update."""
        plan, dim, integ = self.plan, self.dim, self.integrator
        pos, vel, force, imass = self._step_args()
        stream = _device.stream_ptr()
        if isinstance(integ, VelocityVerlet):
            self._evaluate(_ALL)
            if plan.count:
                _lib.call("tmd_vv_kick", _device.dptr(vel), _device.dptr(force), _device.dptr(imass),
                          plan.count, dim, float(integ.timestep), stream)
        else:
            self._evaluate(_lib.WANT_F | _lib.WANT_W)
            if plan.count:
                _lib.call("tmd_langevin_inertia_post", _device.dptr(vel), _device.dptr(force), _device.dptr(imass),
                          plan.count, dim, C.byref(integ._consts), stream)
            self._evaluate(_lib.WANT_V | _lib.V_STANDALONE)
        self.system.particles._vel.written_on_device()

    # ---- looking at the result ------------------------------------------------------------------------
    def reduce_observables(self) -> tuple[float, np.ndarray]:
        """All-reduce the (V, W) shares; returns (v_pot, virial) of the whole system on every rank."""
        import torch.distributed as dist

        part = self.system.particles
        vw = part._vw_buffer().clone()
        if self._collectives:
            dist.all_reduce(vw, group=self.group)
        host = vw.cpu().numpy()
        return float(host[0]), host[1:].reshape(3, 3)[: self.dim, : self.dim].copy()

    def gather_state(self):
        """(pos, vel, force) of the whole system as host arrays on every rank."""
        import torch.distributed as dist

        part, plan, dim = self.system.particles, self.plan, self.dim
        out = []
        for mirror in (part._pos, part._vel, part._force):
            dev = mirror.device()
            if self._collectives and mirror is not part._pos:
                pad = _device.zeros(plan.padded * dim)
                lo = plan.first * dim
                pad[lo: lo + plan.count * dim].copy_(dev[lo: lo + plan.count * dim])
                all_gather_padded(pad, plan, dim, self.group)
                dev = pad[: self.n * dim]
            out.append(dev.cpu().numpy().reshape(self.n, dim))
        del dist
        return tuple(out)
