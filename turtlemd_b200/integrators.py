"""Time integrators (API of ``turtlemd/integrators.py``) over the CUDA kernels.

Every ``integration_step`` enqueues element-wise kernels on the device-resident
particle arrays and calls :meth:`System.evaluate` for the force field; nothing
is copied to the host unless the caller reads ``particles.pos`` etc.

Random numbers (``LangevinOverdamped``, ``LangevinInertia``):

* ``rgen`` is a :class:`turtlemd_b200.philox.PhiloxNormal`: the normals are generated
  inside the kernel from the counter-based stream; the twin object is advanced
  by the number of draws the reference would have made.  This is the fast path
  (and the only one that shards over GPUs); ask for it explicitly.
* ``rgen`` is a NumPy ``Generator`` over ``PCG64`` -- ``numpy.random.default_rng(s)``, the
  reference's default for ``LangevinInertia(rgen=None, seed=s)`` (integrators.py:216): the
  generator is CONTINUED ON THE DEVICE (:mod:`turtlemd_b200.npstream`, ``tmd_npstream_*``): the
  very normals ``rgen.normal`` would return, bit for bit, produced where the kernels consume them;
  the host generator is moved to the position the device reached.  A stock script retraces the
  reference's trajectory under the same seed without a host draw or an upload per step.
* any other object with ``.normal(loc, scale, size)`` (other bit generators, the reference
  tests' ``FakeRandomGenerator``, a monkeypatched ``multivariate_normal``): the draws are made on
  the host exactly like the reference makes them (same calls, same order, through the
  module-level ``multivariate_normal`` looked up at call time), uploaded, and the same kernels
  consume them.
"""
from __future__ import annotations

import ctypes as C
import logging
from abc import ABC, abstractmethod

import numpy as np
from numpy.random import Generator

from . import _device, _lib, npstream
from .philox import PhiloxNormal
from .potentials.well import DoubleWell, DoubleWellPair

LOGGER = logging.getLogger(__name__)
LOGGER.addHandler(logging.NullHandler())

_ALL = _lib.WANT_V | _lib.WANT_F | _lib.WANT_W


def _arrays(system):
    """Device tensors (pos, vel, force, imass) and (n, dim) of a system."""
    _device.require_cuda()
    part = system.particles
    n, dim = part._pos.shape
    return part, n, dim, part._pos.device(), part._vel.device(), part._force.device(), part._imass_device()


class MDIntegrator(ABC):
    """Base class (integrators.py:15-33)."""

    def __init__(self, timestep: float, description: str, dynamics: str):
        self.timestep = timestep
        self.description = description
        self.dynamics = dynamics

    @abstractmethod
    def integration_step(self, system):
        """Advance the system one time step."""

    def __call__(self, system):
        return self.integration_step(system)


class Verlet(MDIntegrator):
    """Position Verlet (integrators.py:36-68); ``previous_pos`` lives on the integrator."""

    def __init__(self, timestep: float):
        super().__init__(timestep=timestep, description="Verlet integrator", dynamics="NVE")
        self.timestep_squared = self.timestep**2
        self.half_idt = 0.5 / self.timestep
        self._prev = None

    @property
    def previous_pos(self):
        return None if self._prev is None else self._prev.host()

    @previous_pos.setter
    def previous_pos(self, value):
        self._prev = None if value is None else _device.Mirror(np.asarray(value))

    def integration_step(self, system):
        part, n, dim, pos, vel, force, imass = _arrays(system)
        first = self._prev is None
        if first:
            self._prev = _device.Mirror(np.zeros(part._pos.shape))
            prev = self._prev.device_uninitialised(n * dim)
        else:
            prev = self._prev.device()
        _lib.call("tmd_verlet_step", _device.dptr(pos), _device.dptr(vel), _device.dptr(prev),
                  _device.dptr(force), _device.dptr(imass), n, dim, float(self.timestep),
                  1 if first else 0, _device.stream_ptr())
        part._pos.written_on_device()
        part._vel.written_on_device()
        self._prev.written_on_device()
        system.evaluate(_ALL)


class VelocityVerlet(MDIntegrator):
    """Velocity Verlet (integrators.py:71-94)."""

    def __init__(self, timestep: float):
        super().__init__(timestep=timestep, description="Velocity Verlet integrator", dynamics="NVE")
        self.half_timestep = self.timestep * 0.5

    def integration_step(self, system):
        part, n, dim, pos, vel, force, imass = _arrays(system)
        dt, stream = float(self.timestep), _device.stream_ptr()
        _lib.call("tmd_vv_kick_drift", _device.dptr(pos), _device.dptr(vel), _device.dptr(force),
                  _device.dptr(imass), n, dim, dt, stream)
        part._pos.written_on_device()
        part._vel.written_on_device()
        part._pos.prefetch_host()  # a caller that reads pos after every step gets it under the force evaluation
        system.evaluate(_ALL)
        _lib.call("tmd_vv_kick", _device.dptr(vel), _device.dptr(part._force.device()),
                  _device.dptr(imass), n, dim, dt, stream)
        part._vel.written_on_device()

    def run(self, system, steps: int):
        """``steps`` integration steps without touching the host in between.

        The last half kick of one step and the first half kick + drift of the next are fused into
        one pass (both roundings kept, so the trajectory is bit-identical to ``steps`` single calls).
        """
        if steps <= 0:
            return
        if steps >= 2 and self._run_resident(system, steps):
            return
        part, n, dim, pos, vel, force, imass = _arrays(system)
        dt, stream = float(self.timestep), _device.stream_ptr()
        args = (_device.dptr(pos), _device.dptr(vel), _device.dptr(force), _device.dptr(imass), n, dim, dt, stream)
        _lib.call("tmd_vv_kick_drift", *args)
        part._pos.written_on_device()
        part._vel.written_on_device()
        for _ in range(steps - 1):
            system.evaluate(_ALL)
            _lib.call("tmd_vv_kick_kick_drift", _device.dptr(pos), _device.dptr(vel),
                      _device.dptr(part._force.device()), _device.dptr(imass), n, dim, dt, stream)
            part._pos.written_on_device()
            part._vel.written_on_device()
        system.evaluate(_ALL)
        _lib.call("tmd_vv_kick", _device.dptr(vel), _device.dptr(part._force.device()),
                  _device.dptr(imass), n, dim, dt, stream)
        part._vel.written_on_device()


def _vv_run_resident(self, system, steps: int) -> bool:
    """The fused force + kick + drift path (``tmd_nlist_run_vv``): the whole run is ONE library call, the state
    lives in the neighbour list's cell order and each step is a single traversal whose epilogue integrates.
    Taken when the force field is one ``LennardJonesCut`` on its neighbour-list path; bit-identical to the
    step-by-step path (same lists, same roundings).  Returns False when the system does not qualify."""
    pots = system.potentials
    if not RESIDENT_RUNS or not pots or not hasattr(pots[0], "_resident_args"):
        return False
    bond = None
    if len(pots) == 2 and isinstance(pots[1], DoubleWellPair):
        bond = pots[1]  # System sums the potentials in order: Lennard-Jones, then the pair double well
    elif len(pots) != 1:
        return False
    _device.require_cuda()
    res = pots[0]._resident_args(system)
    if res is None:
        return False
    handle, tidx, n, box_abi, table, ntyp = res
    part, n, dim, pos, vel, force, imass = _arrays(system)
    vw = part._vw_buffer()
    args = (handle, _device.dptr(pos), _device.dptr(vel), _device.dptr(force), _device.dptr(imass),
            _device.dptr(tidx) if tidx is not None else None, n, C.byref(box_abi), _lib.hptr(table), ntyp,
            float(self.timestep), int(steps))
    if bond is not None:
        idx0, idx1, dev0, dev1 = bond._active(part)
        same = bond.types[0] == bond.types[1]
        if (len(idx0) if same else len(idx0) + len(idx1)) > 128:
            return False  # the resident run evaluates the bonded atoms in one block
        rwidth, width2, height = bond._scalars()
        term = _lib.DwPairTerm(dev0.data_ptr() if dev0 is not None else None, len(idx0),
                               dev1.data_ptr() if dev1 is not None else None, len(idx1), 1 if same else 0,
                               rwidth, width2, height)
        _lib.call("tmd_nlist_run_vv_dwpair", *args, C.byref(term), _device.dptr(vw), _device.stream_ptr())
    else:
        _lib.call("tmd_nlist_run_vv", *args, _device.dptr(vw), _device.stream_ptr())
    part._pos.written_on_device()
    part._vel.written_on_device()
    part._force.written_on_device(shape=part._pos.shape)
    part._vw_fresh.update(("v_pot", "virial"))
    return True


RESIDENT_RUNS = True  # module switch (tests compare the fused run with the step-by-step path)
VelocityVerlet._run_resident = _vv_run_resident


def _is_device_stream(rgen) -> bool:
    return isinstance(rgen, PhiloxNormal)


class LangevinOverdamped(MDIntegrator):
    """Overdamped Langevin / Brownian dynamics (integrators.py:97-163)."""

    def __init__(self, timestep: float, gamma: float, rgen: Generator, beta: float):
        super().__init__(timestep=timestep, description="Langevin overdamped integrator", dynamics="stochastic")
        self.gamma = gamma
        self.sigma = 0.0
        self.rgen = rgen
        self.bddt = 0.0
        self.beta = beta
        self._initiate = True
        self._np_stream = None

    def _numpy_stream(self):
        """The device continuation of ``rgen`` when it is a NumPy PCG64 generator, else None."""
        if not npstream.usable(self.rgen):
            return None
        if self._np_stream is None or self._np_stream.rgen is not self.rgen:
            self._np_stream = npstream.NumpyStream(self.rgen)
        return self._np_stream

    def initiate_parameters(self, system):
        """sigma and dt/(m gamma) per particle (integrators.py:137-149); the kernel re-derives them from imass."""
        if self._initiate:
            imass = system.particles.imass
            self.sigma = np.sqrt(2.0 * self.timestep * imass / (self.beta * self.gamma))
            self.bddt = self.timestep * imass / self.gamma
            self._initiate = False

    def integration_step(self, system):
        self.run(system, 1)

    def run(self, system, steps: int):
        """``steps`` steps without touching the host in between.  The reference evaluates ``potential()`` after
        every step (integrators.py:163), but a later step overwrites it before anybody can look: only the last
        step's evaluation is made -- steps + 1 passes over the force field instead of 2 steps.  Same
        trajectory and same final ``force`` / ``virial`` / ``v_pot`` as ``steps`` single calls."""
        if steps <= 0:
            return
        stream = None if _is_device_stream(self.rgen) else self._numpy_stream()
        if stream is not None:
            stream.begin()
        try:
            for k in range(steps):
                self._step(system, last=k == steps - 1, stream=stream)
        finally:
            if stream is not None:
                stream.end()

    def _step(self, system, last: bool, stream=None):
        self.initiate_parameters(system)
        system.evaluate(_lib.WANT_F | _lib.WANT_W)
        part, n, dim, pos, vel, force, imass = _arrays(system)
        rng, noise = None, None
        if stream is not None:
            # rgen.normal(loc=0, scale=sigma, size=vel.shape) (integrators.py:156-160) = sigma * the next n dim
            # standard normals of the generator, which the device produces itself
            normals = stream.normals(n * dim)
            _lib.call("tmd_langevin_overdamped_normals", _device.dptr(pos), _device.dptr(vel), _device.dptr(force),
                      _device.dptr(imass), n, dim, float(self.timestep), float(self.gamma), float(self.beta),
                      _device.dptr(normals), _device.stream_ptr())
        else:
            if _is_device_stream(self.rgen):
                rng = _lib.Rng(self.rgen.key, self.rgen.position)
                self.rgen.advance(n * dim)
            else:
                rands = self.rgen.normal(loc=0.0, scale=self.sigma, size=part._vel.shape)
                noise = _device.to_device(np.asarray(rands, dtype=np.float64))
            _lib.call("tmd_langevin_overdamped", _device.dptr(pos), _device.dptr(vel), _device.dptr(force),
                      _device.dptr(imass), n, dim, float(self.timestep), float(self.gamma), float(self.beta),
                      C.byref(rng) if rng is not None else None, 0,
                      _device.dptr(noise) if noise is not None else None, _device.stream_ptr())
        part._pos.written_on_device()
        part._vel.written_on_device()
        if last:
            system.evaluate(_lib.WANT_V | _lib.V_STANDALONE)


class LangevinParameter:
    """Constants of the Langevin integrator (integrators.py:166-177).

    ``mean`` / ``cov`` / ``cho`` are per-particle lists in the reference; here they are built on
    first access from stacked arrays (a million-element Python list is not something the device
    path should pay for).
    """

    def __init__(self, c0: float = 0.0, a1: float = 0.0, a2=None, b1=None, b2=None, mean=None, cov=None, cho=None):
        # the reference's dataclass fields, in its order (integrators.py:170-177)
        self.c0 = c0
        self.a1 = a1
        self.a2 = np.zeros(1) if a2 is None else a2
        self.b1 = np.zeros(1) if b1 is None else b1
        self.b2 = np.zeros(1) if b2 is None else b2
        self._cov = None
        self._lists = None
        if mean is not None or cov is not None or cho is not None:
            self._lists = (list(mean or []), list(cov or []), list(cho or []))

    def __repr__(self) -> str:
        return (f"LangevinParameter(c0={self.c0!r}, a1={self.a1!r}, a2={self.a2!r}, b1={self.b1!r}, "
                f"b2={self.b2!r}, mean=[...{len(self.mean)}], cov=[...], cho=[...])")

    def _materialise(self):
        if self._lists is None:
            if self._cov is None:
                self._lists = ([], [], [])
            else:
                cho = np.linalg.cholesky(self._cov)
                self._lists = ([np.zeros(2) for _ in self._cov], list(self._cov), list(cho))
        return self._lists

    @property
    def mean(self):
        return self._materialise()[0]

    @property
    def cov(self):
        return self._materialise()[1]

    @property
    def cho(self):
        return self._materialise()[2]


class LangevinInertia(MDIntegrator):
    """Langevin dynamics with inertia (integrators.py:180-297)."""

    def __init__(self, timestep: float, gamma: float, beta: float, rgen: Generator | None = None, seed: int = 0):
        super().__init__(timestep=timestep, description="Langevin overdamped integrator", dynamics="stochastic")
        self.gamma = gamma
        # the reference's default (integrators.py:216): PCG64 + ziggurat, drawn on the host in the reference's order, so
        # that a stock script retraces the reference under the same seed; pass rgen=PhiloxNormal(key) for in-kernel noise
        self.rgen = np.random.default_rng(seed) if rgen is None else rgen
        self.beta = beta
        self._initiate = True
        self.param = LangevinParameter()
        self._consts = None
        self._np_stream = None
        self._chol_dev = None

    def _numpy_stream(self):
        """The device continuation of ``rgen`` when it is a NumPy PCG64 generator drawn from the way the reference
        draws (stock ``draw_random_numbers`` and ``multivariate_normal``), else None."""
        import sys

        if (type(self).draw_random_numbers is not LangevinInertia.draw_random_numbers
                or sys.modules[__name__].multivariate_normal is not multivariate_normal
                or not npstream.usable(self.rgen)):
            return None
        if self._np_stream is None or self._np_stream.rgen is not self.rgen:
            self._np_stream = npstream.NumpyStream(self.rgen)
        return self._np_stream

    def _cholesky_device(self):
        """(l00, l10, l11) per particle of the reference's ``np.linalg.cholesky(cov_matrix)`` (integrators.py:261), on
        the device; one factorisation per distinct mass."""
        if self._chol_dev is None:
            cov = self.param._cov
            uniq, inverse = np.unique(cov.reshape(len(cov), 4), axis=0, return_inverse=True)
            cho = np.linalg.cholesky(uniq.reshape(-1, 2, 2))
            packed = np.stack([cho[:, 0, 0], cho[:, 1, 0], cho[:, 1, 1]], axis=1)[inverse.reshape(-1)]
            self._chol_dev = _device.to_device(np.ascontiguousarray(packed, dtype=np.float64))
        return self._chol_dev

    def initiate_parameters(self, system):
        """c0, a1, a2, b1, b2 and the noise covariance per particle (integrators.py:223-263)."""
        if not self._initiate:
            return
        dt = self.timestep
        gamma_dt = self.gamma * dt
        exp_gdt = np.exp(-gamma_dt)
        c_0 = exp_gdt
        c_1 = (1.0 - c_0) / gamma_dt
        c_2 = (1.0 - c_1) / gamma_dt
        imasses = system.particles.imass
        par = self.param
        par.c0 = c_0
        par.a1 = c_1 * dt
        par.a2 = c_2 * dt**2 * imasses
        par.b1 = (c_1 - c_2) * dt * imasses
        par.b2 = c_2 * dt * imasses
        kfac = 2.0 - (3.0 - 4.0 * exp_gdt + exp_gdt**2) / gamma_dt
        im = imasses[:, 0]
        sig_r2 = (dt * im / (self.beta * self.gamma)) * kfac
        sig_v2 = (1.0 - exp_gdt**2) * im / self.beta
        cov_rv = (im / (self.beta * self.gamma)) * (1.0 - exp_gdt) ** 2
        cov = np.empty((len(im), 2, 2))
        cov[:, 0, 0] = sig_r2
        cov[:, 0, 1] = cov_rv
        cov[:, 1, 0] = cov_rv
        cov[:, 1, 1] = sig_v2
        par._cov = cov
        par._lists = None
        consts = _lib.LangevinConsts()
        consts.c0 = c_0
        consts.a1 = par.a1
        consts.a2f = c_2 * dt**2
        consts.b1f = (c_1 - c_2) * dt
        consts.b2f = c_2 * dt
        consts.dt = dt
        consts.beta = self.beta
        consts.beta_gamma = self.beta * self.gamma
        consts.kfac = kfac
        consts.one_minus_e2 = 1.0 - exp_gdt**2
        consts.one_minus_e_sq = (1.0 - exp_gdt) ** 2
        self._consts = consts
        self._chol_dev = None
        self._initiate = False

    def draw_random_numbers(self, system):
        """Host draw of (pos_rand, vel_rand), particle-major like the reference (integrators.py:284-297)."""
        self.initiate_parameters(system)
        particles = system.particles
        shape = particles._pos.shape
        dim = particles.dim
        pos_rand = np.zeros(shape)
        vel_rand = np.zeros(shape)
        import sys

        draw = sys.modules[__name__].multivariate_normal  # looked up at call time: tests monkeypatch it
        if draw is multivariate_normal and isinstance(self.rgen, Generator):
            # The reference makes one rgen.normal(0, 1, 2 dim) call per particle, in particle order; for a NumPy
            # Generator that is the same stream as ONE call of shape (n, dim, 2) (probed), and
            # mean + norm @ cho.T per particle is the batched product below: no Python loop over a million particles.
            n = shape[0]
            norm = self.rgen.normal(loc=0.0, scale=1.0, size=(n, dim, 2))
            cho = np.asarray(self.param.cho, dtype=float).reshape(n, 2, 2)
            mean = np.asarray(self.param.mean, dtype=float).reshape(n, 1, 2)
            randxv = mean + np.matmul(norm, np.transpose(cho, (0, 2, 1)))
            return np.ascontiguousarray(randxv[:, :, 0]), np.ascontiguousarray(randxv[:, :, 1])
        for i, (meani, choi) in enumerate(zip(self.param.mean, self.param.cho)):
            randxv = draw(self.rgen, meani, choi, dim)
            pos_rand[i] = randxv[:, 0]
            vel_rand[i] = randxv[:, 1]
        return pos_rand, vel_rand

    def _fusable(self, system) -> bool:
        pots = system.potentials
        return _is_device_stream(self.rgen) and len(pots) == 1 and type(pots[0]) is DoubleWell

    def integration_step(self, system):
        self.initiate_parameters(system)
        if self._fusable(system):
            return self.run(system, 1)
        if not _is_device_stream(self.rgen) and self._numpy_stream() is not None:
            return self._run_numpy_stream(system, 1)
        part, n, dim, pos, vel, force, imass = _arrays(system)
        stream = _device.stream_ptr()
        rng, npos, nvel = None, None, None
        if _is_device_stream(self.rgen):
            rng = _lib.Rng(self.rgen.key, self.rgen.position)
            self.rgen.advance(2 * n * dim)
        else:
            pos_rand, vel_rand = self.draw_random_numbers(system)
            npos, nvel = _device.to_device(pos_rand), _device.to_device(vel_rand)
        _lib.call("tmd_langevin_inertia_pre", _device.dptr(pos), _device.dptr(vel), _device.dptr(force),
                  _device.dptr(imass), n, dim, C.byref(self._consts),
                  C.byref(rng) if rng is not None else None, 0,
                  _device.dptr(npos) if npos is not None else None,
                  _device.dptr(nvel) if nvel is not None else None, stream)
        part._pos.written_on_device()
        part._vel.written_on_device()
        part._pos.prefetch_host()
        # force() and, after the velocity update, potential() -- at the same positions
        # (integrators.py:280-282): one traversal when that gives the same numbers
        one_pass = system.energy_rides_with_forces()
        system.evaluate(_lib.WANT_F | _lib.WANT_W | (_lib.WANT_V if one_pass else 0))
        _lib.call("tmd_langevin_inertia_post", _device.dptr(vel), _device.dptr(part._force.device()),
                  _device.dptr(imass), n, dim, C.byref(self._consts), stream)
        part._vel.written_on_device()
        if not one_pass:
            system.evaluate(_lib.WANT_V | _lib.V_STANDALONE)

    def run(self, system, steps: int, first: int = 0, n_global: int | None = None):
        """``steps`` steps in ONE kernel for a DoubleWell-only system with the in-kernel stream.

        ``first`` / ``n_global``: this process holds replicas ``first .. first+n-1`` of an ensemble
        of ``n_global`` (replica sharding over GPUs); the noise of a replica depends only on its
        global index and the step, never on the sharding.
        """
        self.initiate_parameters(system)
        if not self._fusable(system):
            if first != 0 or n_global is not None:
                # a shard of a larger ensemble on a force field the fused kernel does not cover: the general
                # path addresses its noise from draw 0 of this process's arrays, so every shard would see the
                # SAME noise.  Refuse rather than correlate replicas across GPUs.
                raise NotImplementedError(
                    "LangevinInertia.run(first=, n_global=) shards DoubleWell-only replica ensembles driven by the "
                    "counter-based PhiloxNormal stream; for other force fields use parallel.AtomDecomposition")
            if steps > 1 and _is_device_stream(self.rgen):
                if self._run_resident(system, steps):
                    return None
                return self._run_general(system, steps)
            if steps > 0 and not _is_device_stream(self.rgen) and self._numpy_stream() is not None:
                return self._run_numpy_stream(system, steps)
            for _ in range(steps):
                self.integration_step(system)
            return
        part, n, dim, pos, vel, force, imass = _arrays(system)
        n_global = n if n_global is None else n_global
        pot = system.potentials[0]
        vw = part._vw_buffer()
        rng = _lib.Rng(self.rgen.key, self.rgen.position)
        _lib.call("tmd_langevin_inertia_doublewell", _device.dptr(pos), _device.dptr(vel), _device.dptr(force),
                  _device.dptr(imass), n, dim, *pot._abc(), C.byref(self._consts), C.byref(rng), first,
                  n_global, steps, _device.dptr(vw), _device.stream_ptr())
        self.rgen.advance(2 * n_global * dim * steps)
        part._pos.written_on_device()
        part._vel.written_on_device()
        part._force.written_on_device()
        part.virial = np.zeros((dim, dim))  # DoubleWell.force returns a zero virial (well.py:80-81)
        part._vw_fresh.add("v_pot")


def _langevin_run_general(self, system, steps: int):
    """``steps`` steps of any device force field without touching the host in between: the closing velocity
    update of a step and the opening update of the next are one pass (``tmd_langevin_inertia_post_pre``, same
    roundings), and ``potential()`` -- which the reference evaluates after every step (integrators.py:282) but
    whose value only the last step leaves behind -- is evaluated once, at the end.  Same trajectory, same
    final ``v_pot`` / ``virial`` as ``steps`` single calls."""
    part, n, dim, pos, vel, force, imass = _arrays(system)
    stream = _device.stream_ptr()
    one_pass = system.energy_rides_with_forces()

    def rng_now():
        rng = _lib.Rng(self.rgen.key, self.rgen.position)
        self.rgen.advance(2 * n * dim)
        return rng

    _lib.call("tmd_langevin_inertia_pre", _device.dptr(pos), _device.dptr(vel), _device.dptr(force),
              _device.dptr(imass), n, dim, C.byref(self._consts), C.byref(rng_now()), 0, None, None, stream)
    part._pos.written_on_device()
    part._vel.written_on_device()
    for _ in range(steps - 1):
        system.evaluate(_lib.WANT_F | _lib.WANT_W)
        _lib.call("tmd_langevin_inertia_post_pre", _device.dptr(pos), _device.dptr(vel),
                  _device.dptr(part._force.device()), _device.dptr(imass), n, dim, C.byref(self._consts),
                  C.byref(rng_now()), 0, stream)
        part._pos.written_on_device()
        part._vel.written_on_device()
    system.evaluate(_lib.WANT_F | _lib.WANT_W | (_lib.WANT_V if one_pass else 0))
    _lib.call("tmd_langevin_inertia_post", _device.dptr(vel), _device.dptr(part._force.device()),
              _device.dptr(imass), n, dim, C.byref(self._consts), stream)
    part._vel.written_on_device()
    if not one_pass:
        system.evaluate(_lib.WANT_V | _lib.V_STANDALONE)


LangevinInertia._run_general = _langevin_run_general


def _langevin_run_numpy_stream(self, system, steps: int):
    """``steps`` steps under the reference's own generator continued on the device (:mod:`turtlemd_b200.npstream`): per
    step the next 2 n dim standard normals of ``rgen`` -- what the reference's n calls of ``multivariate_normal`` consume,
    in its order (integrators.py:284-297, :312) -- are generated in HBM and one pass applies the closing update of the
    previous step and the opening update of this one.  Same structure as :func:`_langevin_run_general`; the host
    generator is moved to the device's position once, at the end."""
    part, n, dim, pos, vel, force, imass = _arrays(system)
    stream = _device.stream_ptr()
    one_pass = system.energy_rides_with_forces()
    ns = self._numpy_stream()
    chol = self._cholesky_device()
    ns.begin()
    try:
        for k in range(steps):
            if k > 0:
                system.evaluate(_lib.WANT_F | _lib.WANT_W)
            normals = ns.normals(2 * n * dim)
            _lib.call("tmd_langevin_inertia_pre_normals", _device.dptr(pos), _device.dptr(vel),
                      _device.dptr(part._force.device()), _device.dptr(imass), n, dim, C.byref(self._consts),
                      _device.dptr(normals), _device.dptr(chol), 1 if k > 0 else 0, stream)
            part._pos.written_on_device()
            part._vel.written_on_device()
        part._pos.prefetch_host()
        system.evaluate(_lib.WANT_F | _lib.WANT_W | (_lib.WANT_V if one_pass else 0))
        _lib.call("tmd_langevin_inertia_post", _device.dptr(vel), _device.dptr(part._force.device()),
                  _device.dptr(imass), n, dim, C.byref(self._consts), stream)
        part._vel.written_on_device()
        if not one_pass:
            system.evaluate(_lib.WANT_V | _lib.V_STANDALONE)
    finally:
        ns.end()


LangevinInertia._run_numpy_stream = _langevin_run_numpy_stream


def _langevin_run_resident(self, system, steps: int) -> bool:
    """The resident path (``tmd_nlist_run_langevin``): slot-ordered state, per step one traversal that leaves the forces by
    slot and ONE pass over the state; taken when the force field is one ``LennardJonesCut`` on its neighbour-list path.
    Bit-identical to the step-by-step path (same lists, same roundings, noise addressed by atom index)."""
    pots = system.potentials
    if len(pots) != 1 or not hasattr(pots[0], "_resident_args") or not (RESIDENT_RUNS and RESIDENT_LANGEVIN):
        return False
    _device.require_cuda()
    res = pots[0]._resident_args(system)
    if res is None:
        return False
    handle, tidx, n, box_abi, table, ntyp = res
    part, n, dim, pos, vel, force, imass = _arrays(system)
    vw = part._vw_buffer()
    rng = _lib.Rng(self.rgen.key, self.rgen.position)
    _lib.call("tmd_nlist_run_langevin", handle, _device.dptr(pos), _device.dptr(vel), _device.dptr(force),
              _device.dptr(imass), _device.dptr(tidx) if tidx is not None else None, n, C.byref(box_abi),
              _lib.hptr(table), ntyp, C.byref(self._consts), C.byref(rng), int(steps), _device.dptr(vw),
              _device.stream_ptr())
    self.rgen.advance(2 * n * dim * steps)
    part._pos.written_on_device()
    part._vel.written_on_device()
    part._force.written_on_device(shape=part._pos.shape)
    part._vw_fresh.update(("v_pot", "virial"))
    return True


# Measured on B200 at N = 1M (profiles/r02ab_launches_langevin.txt): the slot-ordered pass draws its six normals per atom
# from three Philox blocks (93 us) where the element-wise kernels share every block between two coordinates, and that
# outweighs the gather / check pass it saves: 0.489 ms per step against 0.470 for the step-by-step path.  Off by default;
# the multi-GPU slab run (parallel.SlabRun) uses the same pass, where it is 12 us of a 124 us step.
RESIDENT_LANGEVIN = False
LangevinInertia._run_resident = _langevin_run_resident


def multivariate_normal(rgen: Generator, mean: np.ndarray, cho: np.ndarray, dim: int) -> np.ndarray:
    """``dim`` draws from N(mean, cho cho^T) given the Cholesky factor (integrators.py:300-315)."""
    norm = rgen.normal(loc=0.0, scale=1.0, size=2 * dim)
    norm = norm.reshape(dim, 2)
    return np.array([mean] * dim) + norm @ cho.T
