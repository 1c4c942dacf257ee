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
* :class:`SlabDecomposition` -- the same system, but every rank owns a spatial SLAB: a contiguous
  range of the neighbour list's cell-order slots, with its state (records, velocities, forces) kept
  in that order.  The integrator kernels advance the records the pair kernel gathers from, so per
  step a rank only exchanges the records of its boundary slots with its two neighbours (a few
  hundred kB instead of the whole position array) and nothing else scales with N; the whole state
  is exchanged only when the lists are rebuilt (every ~10 steps), which also re-assigns the slabs.
* :func:`replica_shard` -- independent replica ensembles (the infretis use case): rank r holds
  replicas ``[r*n, (r+1)*n)`` of a global ensemble, no communication at all; the noise of replica
  i depends only on its global index (counter-based stream), so any sharding gives the same
  trajectories.

:class:`ShardPlan` is pure host logic and is what the CPU (gloo) tests exercise.
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass

import numpy as np

from . import _device, _lib
from .integrators import LangevinInertia, VelocityVerlet, _is_device_stream
from .potentials.lennardjones import LennardJonesCut
from .potentials.well import DoubleWell, DoubleWellPair

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
        self._step_counter = None  # device-side step counter of a captured Langevin step
        self.steps_done = 0        # integration steps taken since prepare()
        self._counter_base = 0

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
        """All potentials for this rank's rows; vw holds this rank's share of V and W.

        ``LennardJonesCut``: rows ``[first, first+count)`` of the pair sum (no Newton's-third-law traffic
        between ranks).  ``DoubleWellPair`` (well.py:262-286; a handful of active atoms, no cut-off): every
        rank evaluates all active pairs from the replicated positions -- the forces land in the replicated
        force array, of which a rank only ever reads its own rows -- and rank 0 alone contributes V and W to
        the all-reduced observables.  ``DoubleWell`` (one-body, well.py:65-81): the own rows only.
        Pair potentials first, in list order, so that the first one overwrites and the rest accumulate."""
        part = self.system.particles
        pos, force, vw = part._pos.device(), part._force.device(), part._vw_buffer()
        want_f = bool(flags & _lib.WANT_F)
        pots = self.system.potentials
        ordered = [p for p in pots if isinstance(p, LennardJonesCut)] + [p for p in pots if not isinstance(p, LennardJonesCut)]
        if not ordered or not isinstance(ordered[0], LennardJonesCut):
            raise NotImplementedError("AtomDecomposition needs at least one LennardJonesCut potential (it defines the rows)")
        first = True
        for pot in ordered:
            acc = 0 if first else _lib.ACCUMULATE
            if isinstance(pot, LennardJonesCut):
                pot._launch(self.system, pos, force if want_f else None, vw, flags | acc,
                            first=self.plan.first, count=self.plan.count)
            elif isinstance(pot, DoubleWellPair):
                mine = flags if self.rank == 0 else flags & ~(_lib.WANT_V | _lib.WANT_W)
                if mine & (_lib.WANT_V | _lib.WANT_F | _lib.WANT_W):
                    pot._launch(self.system, pos, force if want_f else None, vw, mine | acc)
            elif type(pot) is DoubleWell:
                lo, cnt, dim = self.plan.first * self.dim, self.plan.count, self.dim
                if cnt:
                    _lib.call("tmd_doublewell", _device.dptr(pos[lo: lo + cnt * dim]), cnt, dim, *pot._abc(), flags | acc,
                              _device.dptr(force[lo: lo + cnt * dim]) if want_f else None, _device.dptr(vw),
                              _device.stream_ptr())
            else:
                raise NotImplementedError(f"AtomDecomposition cannot shard {type(pot).__name__}")
            first = False
        if want_f:
            part._force.written_on_device()

    def _gather_positions(self) -> None:
        if self._collectives:
            all_gather_padded(self._pos_pad, self.plan, self.dim, self.group)

    def step(self) -> None:
        """One integration step: move the own atoms, exchange positions, forces for the own rows."""
        self.steps_done += 1
        if self._graph is not None:
            self._graph.replay()
            _lib.add_launches(self._graph_launches)
            if self._step_counter is not None:  # keep the host-side stream position in step with the device counter
                self.integrator.rgen.advance(2 * self.n * self.dim)
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
        LangevinInertia: the position in the noise stream is a device-side step counter that the recorded
        step itself advances (``tmd_rng.step_counter``), so every replay draws the noise of ITS step --
        the same numbers an eager run draws.  Returns False, leaving eager stepping in place, when capture
        is not possible."""
        if self._graph is not None:
            return True
        torch = _device.torch()
        for _ in range(warm_steps):  # allocations, occupancy queries and the first list build happen here
            self.step()
        torch.cuda.synchronize()
        langevin = isinstance(self.integrator, LangevinInertia)
        if langevin:
            self._step_counter = torch.zeros(1, dtype=torch.int64, device="cuda")
            self._counter_base = self.integrator.rgen.position
        graph = torch.cuda.CUDAGraph()
        before = _lib.launch_count()
        try:
            with torch.cuda.graph(graph):
                self.step_move()
                self._gather_positions()
                self.step_force()
                if langevin:
                    _lib.call("tmd_counter_add", _device.dptr(self._step_counter), 1, _device.stream_ptr())
        except Exception:  # noqa: BLE001 - capture is an optimisation; eager stepping stays correct
            torch.cuda.synchronize()
            self._step_counter = None
            if langevin:
                self.integrator.rgen.position = self._counter_base
            return False
        if langevin:  # recording advanced the host-side position once without drawing anything
            self.integrator.rgen.position = self._counter_base
        self._graph_launches = _lib.launch_count() - before
        _lib.add_launches(-self._graph_launches)  # recording is not launching
        self._graph = graph
        return True

    def release(self) -> None:
        """Drop the captured graph (do this before destroying the process group: a CUDA graph that
        holds NCCL kernels keeps the communicator busy)."""
        if self._graph is not None:
            _device.torch().cuda.synchronize()
            self._graph = None
            self._step_counter = None  # the host-side stream position was advanced with every replay

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
            if self._step_counter is not None:  # being recorded: position = base + counter * (draws per step)
                rng = _lib.Rng(integ.rgen.key, self._counter_base, _device.dptr(self._step_counter).value,
                               2 * self.n * dim)
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
            one_pass = self.system.energy_rides_with_forces()
            self._evaluate(_lib.WANT_F | _lib.WANT_W | (_lib.WANT_V if one_pass else 0))
            if plan.count:
                _lib.call("tmd_langevin_inertia_post", _device.dptr(vel), _device.dptr(force), _device.dptr(imass),
                          plan.count, dim, C.byref(integ._consts), stream)
            if not one_pass:
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


def halo_ranges(plan: ShardPlan, halo: int) -> dict[str, tuple[int, int, int]]:
    """Slot ranges of one halo exchange for ``plan.rank``: name -> (peer, begin, end).

    A rank sends the first ``halo`` slots it owns to the rank below and the last ``halo`` to the rank
    above, and receives the corresponding ranges of its neighbours at the SAME slot indices (records
    are stored by global slot on every rank).  Ranks form a ring: the slot order is periodic because
    the box is.  Pure host logic (tested on CPU)."""
    world, rank = plan.world, plan.rank
    prev, nxt = (rank - 1) % world, (rank + 1) % world
    lo, hi = plan.first, plan.first + plan.count
    nlo, ncnt = plan.bounds(nxt)
    plo, pcnt = plan.bounds(prev)
    if halo > min(plan.count, ncnt, pcnt):
        raise NotImplementedError(
            f"halo of {halo} slots exceeds a slab of {min(plan.count, ncnt, pcnt)}: too many ranks for this "
            "system, use AtomDecomposition")
    return {
        "send_down": (prev, lo, lo + halo),
        "send_up": (nxt, hi - halo, hi),
        "recv_up": (nxt, nlo, nlo + halo),
        "recv_down": (prev, plo + pcnt - halo, plo + pcnt),
    }


def halo_exchange(records, plan: ShardPlan, halo: int, group=None) -> None:
    """Exchange the boundary records with the two neighbouring ranks (NCCL send/recv on CUDA
    tensors, gloo on CPU tensors).  ``records`` is (padded, width); rows = global slots."""
    import torch.distributed as dist

    if plan.world == 1 or halo == 0:
        return
    r = halo_ranges(plan, halo)
    ops = []
    for name in ("send_down", "send_up"):
        peer, a, b = r[name]
        ops.append(dist.P2POp(dist.isend, records[a:b], peer, group))
    for name in ("recv_up", "recv_down"):  # the peer's send_down is my recv_up: same order on both sides
        peer, a, b = r[name]
        ops.append(dist.P2POp(dist.irecv, records[a:b], peer, group))
    for work in dist.batch_isend_irecv(ops):
        work.wait()


class SlabDecomposition:
    """One LJ system stepped by ``world`` GPUs, each owning a slab of cell-order slots.

    Per step and rank: one fused kick+drift(+displacement test) kernel over the own slots, a halo
    exchange of boundary records with the two neighbours, the list traversal for the own rows, the
    closing kick.  The displacement flags are OR-ed over the ranks (a 4-byte all-reduce) and read by
    the host; when set, every rank all-gathers records and velocities, re-bins the whole system
    (identical result everywhere), takes its new slot range and rebuilds its lists.
    ``world == 1`` is the sorted single-GPU run (no exchange).

    The trajectory equals the single-GPU one up to rounding: positions advance as records
    (``x - lattice vector``), forces sum in list order."""

    def __init__(self, system, integrator, group=None, plan: ShardPlan | None = None):
        import torch.distributed as dist

        if not isinstance(integrator, (VelocityVerlet, LangevinInertia)):
            raise NotImplementedError("SlabDecomposition supports VelocityVerlet and LangevinInertia")
        if isinstance(integrator, LangevinInertia) and not _is_device_stream(integrator.rgen):
            raise NotImplementedError("sharded Langevin dynamics needs the counter-based PhiloxNormal stream")
        pots = system.potentials
        if len(pots) != 1 or not isinstance(pots[0], LennardJonesCut):
            raise NotImplementedError("SlabDecomposition steps systems with a single LennardJonesCut potential")
        self.system, self.integrator, self.group, self.pot = system, integrator, group, pots[0]
        self.n, self.dim = system.particles._pos.shape
        if plan is not None:  # tests emulate several ranks in one process and exchange by hand
            self.world, self.rank, self.plan = plan.world, plan.rank, plan
            self._collectives = False
        else:
            self.world = dist.get_world_size(group) if dist.is_initialized() else 1
            self.rank = dist.get_rank(group) if dist.is_initialized() else 0
            self.plan = ShardPlan(self.n, self.world, self.rank)
            self._collectives = self.world > 1
        self.halo = 0
        self.rebuilds = 0
        self._graphs = None
        self._graph_launches = {}
        self.steps_done = 0  # integration steps taken since prepare()

    # ---- set-up -----------------------------------------------------------------------------------
    def prepare(self) -> None:
        """Upload the state, build the lists for this rank's slab, evaluate the initial forces."""
        _device.require_cuda()
        torch = _device.torch()
        part, plan, dim, n = self.system.particles, self.plan, self.dim, self.n
        if plan.count == 0:
            raise NotImplementedError("more ranks than atoms")
        self.records = _device.zeros((plan.padded, 4))
        self.vel_s = _device.zeros(plan.padded * dim)
        self.force_s = _device.zeros(plan.padded * dim)
        self.imass_s = _device.zeros(plan.padded)
        self._flag = _device.zeros(1, dtype="int32")
        self._flag_host = torch.zeros(1, dtype=torch.int32, pin_memory=True)
        self._reach_host = torch.zeros(8, dtype=torch.int32, pin_memory=True)
        types, table, tidx = self.pot._tables(part)
        self._table, self._ntypes = table, len(types)
        self._tidx = part._static_device(f"lj_tidx_{id(self.pot)}", tidx, np.int32) if len(types) > 1 else None
        self._box = self.system.box.to_abi()
        if not isinstance(self._box, _lib.Box):
            raise NotImplementedError("SlabDecomposition needs an orthogonal, fully periodic Box")
        self._half_skin2 = 0.25 * float(self.pot.skin) ** 2
        # a private list handle: the potential's own one may already hold lists (and must stay usable by
        # System.evaluate); this one is bound to the slab's record buffer and dies with the decomposition
        self._nl = C.c_void_p()
        _lib.call("tmd_nlist_create", C.byref(self._nl), float(self.pot.skin))
        _lib.call("tmd_nlist_bind_records", self._nl, _device.dptr(self.records))
        part._pos.device()
        part._vel.device()
        part._force.device_uninitialised(n * dim).zero_()
        part._force.written_on_device(shape=part._pos.shape)
        if isinstance(self.integrator, LangevinInertia):
            self.integrator.initiate_parameters(self.system)
        self._reach_event = None
        self._build()
        if self.world > 1:
            self._agree_on_halo(first=True)
        self._force(_ALL)

    def _view(self):
        view = _lib.NlistBuffers()
        _lib.call("tmd_nlist_view", self._nl, C.byref(view))
        return view

    def _build(self) -> None:
        """Bin all atoms of the atom-ordered position array, lists for the own slots, sort the state."""
        part, plan, dim = self.system.particles, self.plan, self.dim
        stream = _device.stream_ptr()
        pos, vel = part._pos.device(), part._vel.device()
        _lib.call("tmd_nlist_build_slots", self._nl, _device.dptr(pos),
                  _device.dptr(self._tidx) if self._tidx is not None else None, self.n, C.byref(self._box),
                  _lib.hptr(self._table), self._ntypes, plan.first, plan.count, 0, stream)
        self._v = self._view()
        if self._v.records != self.records.data_ptr():
            raise RuntimeError("neighbour list does not use the bound record buffer")
        _lib.call("tmd_slab_sort", self._v.order, self.n, dim, _device.dptr(vel), _device.dptr(part._imass_device()),
                  None, _device.dptr(self.vel_s), _device.dptr(self.imass_s), None, None, stream)
        self.rebuilds += 1

    def _request_reach(self) -> None:
        """Asynchronous download of the flag words the build just enqueued will write."""
        _lib.call("tmd_memcpy_d2h_async", C.c_void_p(self._reach_host.data_ptr()), self._v.flags, 32,
                  _device.stream_ptr())
        self._reach_event = _device.torch().cuda.Event()
        self._reach_event.record()

    def _collect_reach(self) -> tuple[int, int] | None:
        """(slots beyond the own range that the lists reference, atoms in the fullest cell layer) of the
        build whose download was requested last; waits for it (a no-op when that was steps ago)."""
        if self._reach_event is None:
            return None
        self._reach_event.synchronize()
        self._reach_event = None
        return int(max(self._reach_host[4], self._reach_host[5])), int(self._reach_host[6])

    def needed_halo(self) -> int:
        """Slots beyond the own range that the current lists reference (synchronises)."""
        self._request_reach()
        return self._collect_reach()[0]

    def _agree_on_halo(self, first: bool) -> None:
        """Halo width in slots, the same on every rank.  The lists of a row in layer z (slowest
        direction) reference layers z-h .. z+h, and the slab boundary may sit anywhere inside a layer,
        so up to h+1 layers of the neighbour are needed: sized by the fullest layer (+25 %) at the
        first build (all ranks bin the whole system identically, so they agree without talking).
        Layer populations drift (a lattice start melts): every later build's reach is downloaded
        asynchronously and checked when it has arrived -- an overrun means forces were computed from
        stale halo records, which is an error, not something to paper over."""
        thinnest = min(self.plan.all_counts())
        if first:
            self._request_reach()
            need, fullest = self._collect_reach()
            if self._collectives:  # the same verdict on every rank: prepare() may fall back to another scheme
                import torch.distributed as dist

                torch = _device.torch()
                worst = torch.tensor([need], device="cuda", dtype=torch.int64)
                dist.all_reduce(worst, op=dist.ReduceOp.MAX, group=self.group)
                need = int(worst.item())
            self.halo = min(int(math.ceil((self._v.half_width + 1) * fullest * 1.25)) + 64, thinnest)
            halo_ranges(self.plan, self.halo)  # raises when a slab is thinner than the halo
            if need > self.halo:
                raise NotImplementedError(
                    f"the lists reach {need} slots beyond a slab, more than the {self.halo} a neighbour can "
                    "supply: too many ranks for this system, use AtomDecomposition")
            return
        previous = self._collect_reach()  # the numbers of the build before this one
        self._request_reach()             # this build's: looked at next time
        need = previous[0] if previous else 0
        if need > self.halo:
            raise RuntimeError(f"rank {self.rank}: the lists reached {need} slots beyond the slab but the halo "
                               f"exchange carries {self.halo}; forces since the previous rebuild are invalid")

    # ---- pieces of a step ---------------------------------------------------------------------------
    def _own(self, tensor, width):
        lo = self.plan.first * width
        return tensor[lo: lo + self.plan.count * width]

    def _force(self, flags: int) -> None:
        vw = self.system.particles._vw_buffer()
        _lib.call("tmd_nlist_force_slots", self._nl, flags,
                  _device.dptr(self.force_s) if flags & _lib.WANT_F else None, _device.dptr(vw), _device.stream_ptr())

    def step_move(self) -> None:
        """Kick + drift of the own slots (records advance in place) and the displacement test."""
        plan, dim, integ, v = self.plan, self.dim, self.integrator, self._v
        stream = _device.stream_ptr()
        self._flag.zero_()
        args = (_device.dptr(self.records), _device.dptr(self.vel_s), _device.dptr(self.force_s),
                _device.dptr(self.imass_s), v.build_records)
        if isinstance(integ, VelocityVerlet):
            _lib.call("tmd_slab_vv_kick_drift", *args, plan.first, plan.count, dim, float(integ.timestep),
                      self._half_skin2, _device.dptr(self._flag), stream)
        else:
            rng = _lib.Rng(integ.rgen.key, integ.rgen.position)
            integ.rgen.advance(2 * self.n * dim)  # the stream moves on by the GLOBAL draw count on every rank
            _lib.call("tmd_slab_langevin_inertia_pre", *args, v.order, plan.first, plan.count, dim,
                      C.byref(integ._consts), C.byref(rng), self._half_skin2, _device.dptr(self._flag), stream)

    def step_force(self) -> None:
        """Forces for the own rows from the current records and the closing velocity update."""
        plan, dim, integ = self.plan, self.dim, self.integrator
        vel, force = self._own(self.vel_s, dim), self._own(self.force_s, dim)
        imass = self._own(self.imass_s, 1)
        stream = _device.stream_ptr()
        if isinstance(integ, VelocityVerlet):
            self._force(_ALL)
            _lib.call("tmd_vv_kick", _device.dptr(vel), _device.dptr(force), _device.dptr(imass), plan.count, dim,
                      float(integ.timestep), stream)
        else:
            self._force(_ALL)  # force() and potential() at the same positions (integrators.py:280-282): one traversal
            _lib.call("tmd_langevin_inertia_post", _device.dptr(vel), _device.dptr(force), _device.dptr(imass),
                      plan.count, dim, C.byref(integ._consts), stream)

    def moved_too_far(self) -> bool:
        """This rank's displacement flag of the last :meth:`step_move` (synchronises)."""
        torch = _device.torch()
        self._flag_host.copy_(self._flag)
        torch.cuda.current_stream().synchronize()
        return bool(self._flag_host[0])

    def unsort(self, with_force: bool = False) -> None:
        """Slot-ordered state of ALL slots -> the atom-ordered device arrays of ``system.particles``."""
        part, dim, v = self.system.particles, self.dim, self._v
        pos, vel = part._pos.device_uninitialised(self.n * dim), part._vel.device_uninitialised(self.n * dim)
        force = part._force.device_uninitialised(self.n * dim) if with_force else None
        _lib.call("tmd_slab_unsort", _device.dptr(self.records), _device.dptr(self.vel_s),
                  _device.dptr(self.force_s) if with_force else None, v.order, v.image, self.n, dim,
                  _device.dptr(pos), _device.dptr(vel), _device.dptr(force) if with_force else None, None,
                  _device.stream_ptr())
        part._pos.written_on_device()
        part._vel.written_on_device()
        if with_force:
            part._force.written_on_device()

    def rebuild(self) -> None:
        """After the FULL exchange of records and velocities: new cell order, new slabs, new lists."""
        self.unsort()
        self._build()
        if self.world > 1:
            self._agree_on_halo(first=False)

    def _exchange_full(self, with_force: bool = False) -> None:
        if not self._collectives:
            return
        plan, dim = self.plan, self.dim
        all_gather_padded(self.records.view(-1), plan, 4, self.group)
        all_gather_padded(self.vel_s, plan, dim, self.group)
        if with_force:
            all_gather_padded(self.force_s, plan, dim, self.group)

    def step(self) -> None:
        """One integration step (all ranks call this together)."""
        import torch.distributed as dist

        torch = _device.torch()
        self.steps_done += 1
        g = self._graphs
        if g is not None:
            g["move"].replay()
            _lib.add_launches(self._graph_launches["move"])
            torch.cuda.current_stream().synchronize()
            if self._flag_host[0]:
                self._exchange_full()
                self.rebuild()
                g["force"].replay()
                _lib.add_launches(self._graph_launches["force"])
            else:
                g["halo_force"].replay()
                _lib.add_launches(self._graph_launches["halo_force"])
            return
        self.step_move()
        if self._collectives:
            dist.all_reduce(self._flag, op=dist.ReduceOp.MAX, group=self.group)
        if self.moved_too_far():
            self._exchange_full()
            self.rebuild()
        elif self._collectives:
            halo_exchange(self.records, self.plan, self.halo, self.group)
        self.step_force()

    def capture(self, warm_steps: int = 3) -> bool:
        """Record the three pieces of a step in CUDA graphs (move + flag all-reduce + flag download;
        halo exchange + forces + kick; forces + kick after a rebuild).  VelocityVerlet only (the
        Langevin kernels take the stream offset as a launch parameter)."""
        if not isinstance(self.integrator, VelocityVerlet) or self._graphs is not None:
            return self._graphs is not None
        torch = _device.torch()
        for _ in range(warm_steps):
            self.step()
        torch.cuda.synchronize()
        self._graphs, self._graph_launches = {}, {}
        try:
            for name in ("move", "halo_force", "force"):
                self._record(name)
        except Exception:  # noqa: BLE001 - capture is an optimisation; eager stepping stays correct
            torch.cuda.synchronize()
            self._graphs = None
            return False
        return True  # recording executes nothing: the state is that of the last eager step

    def release(self) -> None:
        """Drop the captured graphs (do this before destroying the process group: a CUDA graph that
        holds NCCL kernels keeps the communicator busy)."""
        if self._graphs is not None:
            _device.torch().cuda.synchronize()
            self._graphs = None
        if getattr(self, "_nl", None):
            _device.torch().cuda.synchronize()
            _lib.call("tmd_nlist_destroy", self._nl)
            self._nl = None

    def __del__(self):
        try:
            self.release()
        except Exception:  # noqa: BLE001 - interpreter shutdown
            pass

    def _piece(self, name: str) -> None:
        import torch.distributed as dist

        if name == "move":
            self.step_move()
            if self._collectives:
                dist.all_reduce(self._flag, op=dist.ReduceOp.MAX, group=self.group)
            self._flag_host.copy_(self._flag, non_blocking=True)
            return
        if name == "halo_force" and self._collectives:
            halo_exchange(self.records, self.plan, self.halo, self.group)
        self.step_force()

    def _record(self, name: str) -> None:
        torch = _device.torch()
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        before = _lib.launch_count()
        with torch.cuda.graph(graph):
            self._piece(name)
        self._graph_launches[name] = _lib.launch_count() - before
        _lib.add_launches(-self._graph_launches[name])  # recording is not launching
        self._graphs[name] = graph

    # ---- looking at the result ------------------------------------------------------------------------
    def reduce_observables(self) -> tuple[float, np.ndarray]:
        """All-reduce the (V, W) shares; returns (v_pot, virial) of the whole system on every rank."""
        import torch.distributed as dist

        vw = self.system.particles._vw_buffer().clone()
        if self._collectives:
            dist.all_reduce(vw, group=self.group)
        host = vw.cpu().numpy()
        return float(host[0]), host[1:].reshape(3, 3)[: self.dim, : self.dim].copy()

    def gather_state(self):
        """(pos, vel, force) of the whole system in atom order as host arrays on every rank."""
        self._exchange_full(with_force=True)
        self.unsort(with_force=True)
        part = self.system.particles
        return tuple(m.device().cpu().numpy().reshape(self.n, self.dim) for m in (part._pos, part._vel, part._force))



# ---- resident slab run (csrc/slabrun.cuh) -------------------------------------------------------------------------

SR_PHASES = dict(MIGRATE=0, B0_SIGNAL=1, B0_WAIT=2, COLLECT=3, B1_SIGNAL=4, B1_WAIT=5, BUILD=6, B2_SIGNAL=7, B2_WAIT=8,
                 TRAVERSE=9, B3_SIGNAL=10, B3_WAIT=11, PRE=12, ADVANCE=13)
SR_ERRORS = {1: "a peer did not arrive at a barrier", 2: "an atom left its slab by more than one layer between two rebuilds",
             4: "migrant inbox full", 8: "ghost inbox full", 16: "own-atom capacity exceeded",
             32: "own + ghost capacity exceeded", 64: "a ghost layer does not match its owner's layer"}


def slab_layer_plan(layers: int, world: int) -> list[tuple[int, int, int, int]]:
    """Cell layers per rank of a resident slab run: ``(z0, z1, below, above)`` for every rank -- own layers ``[z0, z1)``,
    the two ghost layers.  ``layers`` (cells of edge rcut + skin that fit along z) is rounded down to a multiple of
    ``world`` (thicker cells instead of uneven slabs); two layers per rank are the minimum, so that own and ghost layers
    never coincide.  The rule of ``sr_make_grid`` (csrc/slabrun.cuh), as pure host logic."""
    if world < 2:
        raise NotImplementedError("a slab run needs at least two ranks")
    nz = (layers // world) * world
    if nz < 2 * world or nz < 3:
        raise NotImplementedError(f"{layers} cell layers along z are too few for {world} ranks (two per rank are needed)")
    per = nz // world
    return [(r * per, (r + 1) * per, (r * per - 1) % nz, ((r + 1) * per) % nz) for r in range(world)]


def all_ranks_agree(ok: bool, group=None, device=None) -> bool:
    """True on every rank only if ``ok`` is true on every rank (an all-reduce of one word): the verdict after a step
    that may fail on one rank only -- a rank that went on alone would wait for its peers forever."""
    import torch
    import torch.distributed as dist

    word = torch.tensor([1 if ok else 0], dtype=torch.int32, device=device)
    dist.all_reduce(word, op=dist.ReduceOp.MIN, group=group)
    return bool(word.item())


def _sr_error_text(bits: int) -> str:
    return "; ".join(text for bit, text in SR_ERRORS.items() if bits & bit) or "ok"


class _SlabRunRank:
    """One rank's ``tmd_slabrun`` handle plus the replicated device arrays it loads from / exports to."""

    def __init__(self, system, integrator, world: int, rank: int):
        if not isinstance(integrator, (VelocityVerlet, LangevinInertia)):
            raise NotImplementedError("the resident slab run integrates with VelocityVerlet or LangevinInertia")
        if isinstance(integrator, LangevinInertia) and not _is_device_stream(integrator.rgen):
            raise NotImplementedError("sharded Langevin dynamics needs the counter-based PhiloxNormal stream")
        pots = system.potentials
        bond = None
        if len(pots) == 2 and isinstance(pots[0], LennardJonesCut) and isinstance(pots[1], DoubleWellPair) \
                and isinstance(integrator, VelocityVerlet):
            bond = pots[1]  # System sums the potentials in order: Lennard-Jones, then the pair double well
        elif len(pots) != 1 or not isinstance(pots[0], LennardJonesCut):
            raise NotImplementedError("the resident slab run steps one LennardJonesCut potential, optionally followed by a "
                                      "DoubleWellPair under VelocityVerlet")
        _device.require_cuda()
        self.system, self.integrator, self.pot = system, integrator, pots[0]
        self.world, self.rank = world, rank
        part = system.particles
        self.n, self.dim = part._pos.shape
        if self.dim != 3:
            raise NotImplementedError("the resident slab run needs a three-dimensional system")
        box = system.box.to_abi()
        if not isinstance(box, _lib.Box):
            raise NotImplementedError("the resident slab run needs an orthogonal, fully periodic Box")
        types, table, tidx = self.pot._tables(part)
        self._table = np.ascontiguousarray(table, dtype=np.float64)
        self._tidx = part._static_device(f"lj_tidx_{id(self.pot)}", tidx, np.int32) if len(types) > 1 else None
        # the same verdict on every rank before anything is allocated: enough cell layers along z for this many ranks?
        reach = math.sqrt(float(np.nanmax(self._table[5]))) + float(self.pot.skin)
        self.layer_plan = slab_layer_plan(min(1024, int(float(box.length[2]) / (reach * (1.0 + 1e-6)))), world)
        self.handle = C.c_void_p()
        _lib.call("tmd_slabrun_create", C.byref(self.handle), world, rank, self.n, C.byref(box), _lib.hptr(self._table),
                  len(types), float(self.pot.skin), float(integrator.timestep))
        ptr, size = C.c_void_p(), C.c_size_t()
        _lib.call("tmd_slabrun_window", self.handle, C.byref(ptr), C.byref(size))
        self.window, self.window_bytes = ptr.value, size.value
        self._vw = _device.zeros(10)
        if bond is not None:
            idx0, idx1, _, _ = bond._active(part)
            same = bond.types[0] == bond.types[1]
            if (len(idx0) if same else len(idx0) + len(idx1)) > 128 or len(idx0) == 0:
                self.destroy()
                raise NotImplementedError("the resident slab run takes between 1 and 128 atoms of the bonded types")
            g0 = np.ascontiguousarray(idx0, dtype=np.int32)
            g1 = np.ascontiguousarray(idx1, dtype=np.int32)
            _lib.call("tmd_slabrun_set_dwpair", self.handle, _lib.hptr(g0), len(g0), _lib.hptr(g1) if len(g1) else None,
                      len(g1), 1 if same else 0, *bond._scalars(), C.byref(box))
        self.draws_per_step = 0
        if isinstance(integrator, LangevinInertia):
            integrator.initiate_parameters(system)
            _lib.call("tmd_slabrun_set_langevin", self.handle, C.byref(integrator._consts), integrator.rgen.key,
                      integrator.rgen.position)
            self.draws_per_step = 2 * self.n * 3  # the stream moves on by the GLOBAL draw count on every rank

    def connect(self, windows: list[int]) -> None:
        arr = (C.c_void_p * self.world)(*windows)
        _lib.call("tmd_slabrun_connect", self.handle, arr)

    def load(self) -> None:
        """Slabs, ghost layers and lists from the replicated arrays of ``system.particles`` (forces must be valid)."""
        part = self.system.particles
        pos, vel, force = part._pos.device(), part._vel.device(), part._force.device()
        _lib.call("tmd_slabrun_load", self.handle, _device.dptr(pos), _device.dptr(vel), _device.dptr(force),
                  _device.dptr(part._imass_device()), _device.dptr(self._tidx) if self._tidx is not None else None,
                  _device.stream_ptr())

    def phase(self, name: str, last: int = 0) -> None:
        _lib.call("tmd_slabrun_phase", self.handle, SR_PHASES[name], last, _device.stream_ptr())

    def export(self):
        """(pos, vel, force) torch tensors of n x 3 with this rank's atoms filled in and zeros elsewhere, vw share."""
        pos, vel, force = (_device.zeros(self.n * 3) for _ in range(3))
        _lib.call("tmd_slabrun_export", self.handle, _device.dptr(pos), _device.dptr(vel), _device.dptr(force),
                  _device.dptr(self._vw), _device.stream_ptr())
        return pos, vel, force, self._vw

    def observables(self):
        """This rank's share of (V, W) from the last traversal, as the device buffer vw[10]."""
        _lib.call("tmd_slabrun_export", self.handle, None, None, None, _device.dptr(self._vw), _device.stream_ptr())
        return self._vw

    def status(self) -> dict:
        err, n_own, builds = C.c_int32(), C.c_int64(), C.c_int64()
        layers = (C.c_int32 * 2)()
        _lib.call("tmd_slabrun_status", self.handle, C.byref(err), C.byref(n_own), C.byref(builds), layers)
        return {"err": err.value, "error": _sr_error_text(err.value), "n_own": n_own.value, "builds": builds.value,
                "layers_per_rank": layers[0], "layers": layers[1]}

    def destroy(self) -> None:
        if self.handle:
            _lib.call("tmd_slabrun_destroy", self.handle)
            self.handle = C.c_void_p()


class SlabRun:
    """One LJ system stepped by ``world`` GPUs, each owning a slab of cell layers, without leaving the device:
    per step ONE traversal kernel (forces + kick + kick + drift + the records of the next step, the boundary
    layers stored straight into the neighbours' memory) and ONE barrier kernel; list rebuilds, atom migration
    and ghost layers are gated launches decided on the device.  See ``csrc/slabrun.cuh``.

    ``system`` is the full system on every rank with valid forces (``system.potential_and_force()``).
    ``run(k)`` = k VelocityVerlet steps (integrators.py:84-94); afterwards ``gather_state()`` returns the whole
    system.  The trajectory equals the single-GPU one up to summation order."""

    def __init__(self, system, integrator, group=None):
        import torch.distributed as dist

        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.system = system
        self.r = _SlabRunRank(system, integrator, self.world, self.rank)
        self.n = self.r.n
        self._peers = []
        self.steps_done = 0

    def prepare(self) -> None:
        import torch.distributed as dist

        torch = _device.torch()
        failure = None
        try:
            handle = (C.c_ubyte * 64)()
            _lib.call("tmd_ipc_get_handle", C.c_void_p(self.r.window), handle)
            blob = bytes(handle)
        except Exception as exc:  # noqa: BLE001 - reported to every rank below
            failure, blob = exc, b""
        everyone = [None] * self.world
        dist.all_gather_object(everyone, blob, group=self.group)
        if failure is None and all(everyone):
            try:
                windows = []
                for rank, peer in enumerate(everyone):
                    if rank == self.rank:
                        windows.append(self.r.window)
                        continue
                    ptr = C.c_void_p()
                    _lib.call("tmd_ipc_open_handle", (C.c_ubyte * 64).from_buffer_copy(peer), C.byref(ptr))
                    self._peers.append(ptr)
                    windows.append(ptr.value)
                self.r.connect(windows)
                self.r.load()
                _device.synchronize()
            except Exception as exc:  # noqa: BLE001
                failure = exc
        elif failure is None:
            failure = RuntimeError("a peer could not export its window")
        # one verdict for all ranks (a rank that went on alone would wait for the others forever): after this
        # all-reduce every rank's lists exist, or every rank raises
        if not all_ranks_agree(failure is None, self.group, device="cuda"):
            raise NotImplementedError(
                f"rank {self.rank}: the resident slab run could not be set up (peer memory over CUDA IPC is required): "
                f"{failure if failure is not None else 'a peer failed'}")

    def capture(self) -> bool:
        """Ask :meth:`run` to replay pairs of steps from the library's step graph (recorded at the first run: two
        steps whose rebuild launches sit in conditional IF nodes decided on the device)."""
        self.use_graph = True
        return True

    use_graph = False

    def run(self, steps: int) -> None:
        """``steps`` VelocityVerlet steps; velocities and positions are at the same time afterwards."""
        if steps < 1:
            return
        replayed = C.c_int64(0)
        _lib.call("tmd_slabrun_run", self.r.handle, steps, 1 if self.use_graph else 0, C.byref(replayed),
                  _device.stream_ptr())
        _lib.add_launches(replayed.value)
        self.steps_done += steps
        if self.r.draws_per_step:
            self.r.integrator.rgen.advance(steps * self.r.draws_per_step)

    def status(self) -> dict:
        return self.r.status()

    def reduce_observables(self):
        import torch.distributed as dist

        vw = self.r.observables().clone()
        dist.all_reduce(vw, group=self.group)
        host = vw.cpu().numpy()
        return float(host[0]), host[1:].reshape(3, 3).copy()

    def gather_state(self):
        """(pos, vel, force) of the whole system in atom order as host arrays on every rank."""
        import torch.distributed as dist

        pos, vel, force, vw = self.r.export()
        for t in (pos, vel, force):
            dist.all_reduce(t, group=self.group)  # every atom is owned by exactly one rank, the others hold zeros
        status = self.status()
        if status["err"]:
            raise RuntimeError(f"rank {self.rank}: slab run failed: {status['error']}")
        return tuple(t.cpu().numpy().reshape(self.n, 3) for t in (pos, vel, force))

    def release(self) -> None:
        _device.torch().cuda.synchronize()
        for ptr in self._peers:
            _lib.call("tmd_ipc_close_handle", ptr)
        self._peers = []
        self.r.destroy()


class SlabRunEmulation:
    """``world`` slab-run ranks on ONE GPU, stepped phase by phase in one stream (the signal halves of a barrier
    for all ranks, then the wait halves): the data path of :class:`SlabRun` -- migration, ghost layers, halo
    stores into the neighbours' buffers, gate words -- without a second device.  For tests."""

    def __init__(self, systems, integrator):
        self.world = len(systems)
        self.ranks = [_SlabRunRank(system, integrator, self.world, r) for r, system in enumerate(systems)]
        windows = [r.window for r in self.ranks]
        for r in self.ranks:
            r.connect(windows)
            r.load()
        self.n = self.ranks[0].n
        self._open = False

    def _all(self, name: str, last: int = 0) -> None:
        for r in self.ranks:
            r.phase(name, last)

    def _start(self) -> None:
        self._all("B2_SIGNAL", 1)
        self._all("B2_WAIT", 1)
        self._all("PRE")
        self._all("B3_SIGNAL")
        self._all("B3_WAIT")
        self._all("ADVANCE")

    def _step(self, last: int) -> None:
        for name in ("MIGRATE", "B0_SIGNAL", "B0_WAIT", "COLLECT", "B1_SIGNAL", "B1_WAIT", "BUILD", "B2_SIGNAL", "B2_WAIT"):
            self._all(name)
        self._all("TRAVERSE", last)
        self._all("B3_SIGNAL")
        self._all("B3_WAIT", last)
        self._all("ADVANCE", last)

    def run(self, steps: int) -> None:
        if steps < 1:
            return
        self._start()
        for _ in range(steps - 1):
            self._step(0)
        self._step(1)
        if self.ranks[0].draws_per_step:  # one integrator object (one stream) for all emulated ranks
            self.ranks[0].integrator.rgen.advance(steps * self.ranks[0].draws_per_step)

    def gather_state(self):
        total = None
        vw = np.zeros(10)
        for r in self.ranks:
            pos, vel, force, share = r.export()
            parts = [t.cpu().numpy().reshape(self.n, 3) for t in (pos, vel, force)]
            total = parts if total is None else [a + b for a, b in zip(total, parts)]
            vw += share.cpu().numpy()
            status = r.status()
            if status["err"]:
                raise RuntimeError(f"emulated rank {r.rank}: slab run failed: {status['error']}")
        return total[0], total[1], total[2], float(vw[0]), vw[1:].reshape(3, 3).copy()

    def status(self):
        return [r.status() for r in self.ranks]

    def release(self) -> None:
        _device.synchronize()
        for r in self.ranks:
            r.destroy()
