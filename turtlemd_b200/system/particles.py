"""The particle list (API of ``turtlemd/system/particles.py``) over device-resident arrays.

``pos``, ``vel`` and ``force`` are properties backed by a
:class:`turtlemd_b200._device.Mirror`: the NumPy array the reference API hands
out on one side, a flat float64 tensor in HBM on the other.  ``virial`` and
``v_pot`` are read lazily from the 10-double ``vw`` buffer the kernels reduce
into.  Nothing here computes on the hot path; the observables at the bottom are
the reference's host helpers (SURVEY.md section 8f ranks a device version next).
"""
from __future__ import annotations

import ctypes as C
from collections.abc import Iterator

import numpy as np

from .._device import Mirror


class Particles:
    """A collection of particles: positions, velocities, forces, masses, types."""

    def __init__(self, dim: int = 3):
        self.dim = dim
        self.empty()

    def empty(self):
        """Drop all particles (particles.py:37-48; arrays keep one row while npart == 0)."""
        self.npart = 0
        self._pos = Mirror(np.zeros((1, self.dim)))
        self._vel = Mirror(np.zeros((1, self.dim)))
        self._force = Mirror(np.zeros((1, self.dim)))
        self.mass = np.zeros((1, 1))
        self.imass = np.zeros((1, 1))
        self.name = np.array([], dtype=str)
        self.ptype = np.array([], dtype=int)
        self._virial = np.zeros((self.dim, self.dim))
        self._v_pot = None
        self._vw = None          # device double[10]: V, W(3x3)
        self._vw_fresh = set()   # which of {"v_pot", "virial"} live only on the device
        self._static = {}        # device copies of mass / imass / type indices, keyed by array identity
        self._generation = 0     # bumped by touch(): caches the potentials keep per ptype array check it

    # ---- the mirrored arrays ------------------------------------------------------------------
    @property
    def pos(self) -> np.ndarray:
        return self._pos.host()

    @pos.setter
    def pos(self, value):
        self._pos.set_host(value)

    @property
    def vel(self) -> np.ndarray:
        return self._vel.host()

    @vel.setter
    def vel(self, value):
        self._vel.set_host(value)

    @property
    def force(self) -> np.ndarray:
        return self._force.host()

    @force.setter
    def force(self, value):
        self._force.set_host(value)

    # ---- scalars reduced on the device ---------------------------------------------------------
    def _pull_vw(self):
        if self._vw_fresh:
            host = self._vw.cpu().numpy()
            if "v_pot" in self._vw_fresh:
                self._v_pot = float(host[0])  # a Python float: System.thermo divides it in a local
            if "virial" in self._vw_fresh:
                self._virial = host[1:].reshape(3, 3)[: self.dim, : self.dim].copy()
            self._vw_fresh.clear()

    @property
    def v_pot(self):
        self._pull_vw()
        return self._v_pot

    @v_pot.setter
    def v_pot(self, value):
        self._vw_fresh.discard("v_pot")
        self._v_pot = value

    @property
    def virial(self) -> np.ndarray:
        self._pull_vw()
        return self._virial

    @virial.setter
    def virial(self, value):
        self._vw_fresh.discard("virial")
        self._virial = value

    # ---- building the list ------------------------------------------------------------------------
    def add_particle(self, pos, vel=None, force=None, mass: float = 1.0, name: str = "?", ptype: int = 1):
        """Append one particle (particles.py:50-78)."""
        vel = np.zeros_like(pos) if vel is None else vel
        force = np.zeros_like(pos) if force is None else force
        if self.npart == 0:
            self.pos[0] = pos
            self.vel[0] = vel
            self.force[0] = force
            self.mass[0] = mass
        else:
            self.pos = np.vstack([self.pos, pos])
            self.vel = np.vstack([self.vel, vel])
            self.force = np.vstack([self.force, force])
            self.mass = np.vstack([self.mass, mass])
        self.name = np.append(self.name, name)
        self.ptype = np.append(self.ptype, ptype)
        self.imass = 1.0 / self.mass
        self.npart += 1

    @classmethod
    def from_arrays(cls, pos, vel=None, mass=None, ptype=None, name="?"):
        """Bulk constructor (not in the reference, whose ``add_particle`` is O(N^2) for N particles)."""
        pos = np.array(pos, dtype=float)
        npart, dim = pos.shape
        part = cls(dim=dim)
        part.pos = pos
        part.vel = np.zeros_like(pos) if vel is None else np.array(vel, dtype=float)
        part.force = np.zeros_like(pos)
        mass = np.ones(npart) if mass is None else np.asarray(mass, dtype=float)
        part.mass = np.array(np.broadcast_to(mass.reshape(-1, 1) if mass.ndim else mass, (npart, 1)), dtype=float)
        part.imass = 1.0 / part.mass
        part.ptype = np.zeros(npart, dtype=int) if ptype is None else np.array(ptype, dtype=int)
        part.name = np.array([name] * npart)
        part.npart = npart
        return part

    def __iter__(self) -> Iterator[dict]:
        pos, vel, force = self.pos, self.vel, self.force
        for i in range(len(pos)):
            yield {
                "pos": pos[i], "vel": vel[i], "force": force[i], "mass": self.mass[i],
                "imass": self.imass[i], "name": self.name[i], "type": self.ptype[i],
            }

    def __len__(self) -> int:
        return self.npart

    def __getitem__(self, key):
        """Sub-list; basic slices are views of the host arrays, as in the reference (particles.py:103-118)."""
        part = Particles(dim=self.dim)
        part.pos = self.pos[key]
        part.vel = self.vel[key]
        part.force = self.force[key]
        part.mass = self.mass[key]
        part.imass = self.imass[key]
        part.name = self.name[key]
        part.ptype = self.ptype[key]
        part.virial = self.virial = np.zeros((self.dim, self.dim))  # the parent's is reset too (:113-115)
        part.v_pot = None
        part.npart = len(part.name)
        return part

    def pairs(self) -> Iterator[tuple[int, int, int, int]]:
        """All pairs (i, j, type_i, type_j) with i < j (particles.py:120-132)."""
        ptype = self.ptype
        for i in range(len(ptype) - 1):
            for j in range(i + 1, len(ptype)):
                yield (i, j, ptype[i], ptype[j])

    # ---- device views for the kernels -------------------------------------------------------------
    def _vw_buffer(self):
        if self._vw is None:
            from .. import _device

            self._vw = _device.zeros(10)
        return self._vw

    def _static_device(self, tag: str, array: np.ndarray, dtype):
        """Device copy of an array that only changes by rebinding (mass, imass, type indices)."""
        from .. import _device

        entry = self._static.get(tag)
        if entry is None or entry[0] is not array or entry[1].numel() != array.size:
            entry = (array, _device.to_device(np.asarray(array).reshape(-1), dtype=dtype))
            self._static[tag] = entry
        return entry[1]

    def _imass_device(self):
        return self._static_device("imass", self.imass, np.float64)

    def _mass_device(self):
        return self._static_device("mass", self.mass, np.float64)

    def touch(self):
        """Forget cached device copies of mass / imass / ptype after editing them IN PLACE (also what the potentials
        derived from ptype: type tables, active-atom lists)."""
        self._static.clear()
        self._generation += 1


def _on_device_only(particles: Particles) -> bool:
    """The velocities were last written by a kernel and nobody has looked at them on the host since."""
    mirror = particles._vel
    return mirror.dev_valid and not mirror.host_valid and len(particles.mass) > 1


def _momentum_device(particles: Particles) -> np.ndarray:
    """sum_i m_i v_i reduced on the device (``tmd_momentum``); four doubles come back."""
    from .. import _device, _lib

    npart, dim = particles._vel.shape
    out = _device.zeros(4)
    _lib.call("tmd_momentum", _device.dptr(particles._vel.device()), _device.dptr(particles._mass_device()),
              npart, dim, _device.dptr(out), _device.stream_ptr())
    return out.cpu().numpy()[1: 1 + dim].copy()


def linear_momentum(particles: Particles) -> np.ndarray:
    """Total linear momentum (particles.py:135-137); reduced on the device when the velocities live there."""
    if _on_device_only(particles):
        return _momentum_device(particles)
    return np.sum(particles.vel * particles.mass, axis=0)


def zero_momentum(particles: Particles, dim: list[bool] | None = None):
    """Remove the centre-of-mass motion, optionally only along chosen axes (particles.py:140-153)."""
    on_device = _on_device_only(particles)
    mom = linear_momentum(particles)
    if dim is not None:
        for k, reset in enumerate(dim):
            if not reset:
                mom[k] = 0
    shift = mom / particles.mass.sum()
    if on_device:
        from .. import _device, _lib

        npart, ndim = particles._vel.shape
        shift = np.ascontiguousarray(shift, dtype=np.float64)
        _lib.call("tmd_vel_shift_scale", _device.dptr(particles._vel.device()), npart, ndim, _lib.hptr(shift),
                  0, 1.0, _device.stream_ptr())
        particles._vel.written_on_device()
        return
    particles.vel -= shift


def _kinetic_energy_device(particles: Particles) -> np.ndarray:
    """0.5 sum m v (x) v reduced on the device (``tmd_kinetic_tensor``); only 9 doubles come back."""
    import ctypes as C  # noqa: F401

    from .. import _device, _lib

    npart, dim = particles._vel.shape
    kin = _device.zeros(9)
    _lib.call("tmd_kinetic_tensor", _device.dptr(particles._vel.device()), _device.dptr(particles._mass_device()),
              npart, dim, _device.dptr(kin), _device.stream_ptr())
    return kin.cpu().numpy().reshape(3, 3)[:dim, :dim].copy()


def kinetic_energy(particles: Particles) -> tuple[np.ndarray, float]:
    """Kinetic energy tensor and its trace (particles.py:156-168).

    When the velocities live on the device (an integrator wrote them and nobody has read them on
    the host since) the reduction runs there and the velocity array is not downloaded.
    """
    if _on_device_only(particles):
        kin = _kinetic_energy_device(particles)
        return kin, kin.trace()
    vel = particles._vel.peek()
    mom = vel * particles.mass
    if len(particles.mass) == 1:
        kin = 0.5 * np.outer(mom, vel)
    else:
        kin = 0.5 * np.einsum("ij,ik->jk", mom, vel)
    return kin, kin.trace()


def kinetic_temperature(particles: Particles, boltzmann: float, dof=None, kin_tensor=None):
    """(mean temperature, per-axis temperature, kinetic tensor) (particles.py:171-199)."""
    ndof = particles.npart * np.ones(particles._vel.shape[1:])
    if kin_tensor is None:
        kin_tensor, _ = kinetic_energy(particles)
    if dof is not None:
        ndof = ndof - dof
    temperature = (2.0 * np.diagonal(kin_tensor) / ndof) / boltzmann
    return np.mean(temperature), temperature, kin_tensor


def pressure_tensor(particles: Particles, volume: float, kin_tensor=None):
    """(pressure tensor, scalar pressure) from virial + kinetic tensor (particles.py:202-224)."""
    if kin_tensor is None:
        kin_tensor, _ = kinetic_energy(particles)
    pressure = (particles.virial + 2.0 * kin_tensor) / volume
    return pressure, pressure.trace() / float(particles.dim)


def generate_maxwell_velocities(particles: Particles, rgen, temperature: float = 1.0,
                                boltzmann: float = 1.0, dof=None, momentum: bool = True):
    """Maxwell velocities at a target temperature (particles.py:227-269).

    With the counter-based stream (:class:`turtlemd_b200.philox.PhiloxNormal`) as ``rgen`` and a CUDA
    device present, all three steps run on the device -- draw (``tmd_maxwell_draw``, the same numbers
    ``rgen.normal(size=vel.shape)`` would hand out), momentum removal, rescaling -- and only the
    momentum and the kinetic tensor (13 doubles) visit the host.  Any other generator draws on the
    host exactly like the reference.
    """
    if _device_stream(rgen) and particles.npart > 1:
        from .. import _device, _lib

        npart, dim = particles._vel.shape
        vel = particles._vel.device_uninitialised(npart * dim)
        rng = _lib.Rng(rgen.key, rgen.position)
        rgen.advance(npart * dim)
        _lib.call("tmd_maxwell_draw", _device.dptr(vel), _device.dptr(particles._imass_device()), npart, dim,
                  C.byref(rng), 0, _device.stream_ptr())
        particles._vel.written_on_device()
        if momentum:
            zero_momentum(particles)
        temp_now, _, _ = kinetic_temperature(particles, boltzmann, dof=dof)
        scale = float(np.sqrt(temperature / temp_now))
        _lib.call("tmd_vel_shift_scale", _device.dptr(particles._vel.device()), npart, dim, None, 1, scale,
                  _device.stream_ptr())
        particles._vel.written_on_device()
        return
    vel = np.sqrt(particles.imass) * rgen.normal(loc=0.0, scale=1.0, size=particles.vel.shape)
    particles.vel = vel
    if momentum:
        zero_momentum(particles)
    temp_now, _, _ = kinetic_temperature(particles, boltzmann, dof=dof)
    particles.vel *= np.sqrt(temperature / temp_now)


def _device_stream(rgen) -> bool:
    """``rgen`` is the counter-based stream and there is a device to run it on."""
    from ..philox import PhiloxNormal

    if not isinstance(rgen, PhiloxNormal):
        return False
    from .. import _lib

    try:
        return _lib.device_count() > 0
    except Exception:  # noqa: BLE001 - no library / no driver: the host path of the twin still works
        return False
