"""Position-dependent model potentials (API of ``turtlemd/potentials/well.py``).

* :class:`DoubleWell` -- one-body ``a x^4 - b (x - c)^2`` over every coordinate
  (CUDA kernel K4, element-wise, HBM-bound);
* :class:`DoubleWellPair` -- double-well bond between atoms of two given types,
  no cut-off (CUDA kernel K5 over the active atoms only);
* :class:`RectangularWell` -- outside the hot path (no force in the reference
  either); kept as host NumPy for API completeness.
"""
from __future__ import annotations

import ctypes as C
import logging

import numpy as np

from .. import _device, _lib
from .potential import Potential

LOGGER = logging.getLogger(__name__)
LOGGER.addHandler(logging.NullHandler())


def _host_result(vw, dim, force):
    return float(vw[0]), force, vw[1:].reshape(3, 3)[:dim, :dim].copy()


def host_positions(particles) -> np.ndarray:
    """Contiguous float64 positions of any particle list (ours: without invalidating the device copy)."""
    mirror = getattr(particles, "_pos", None)
    pos = mirror.peek() if mirror is not None else particles.pos
    return np.ascontiguousarray(pos, dtype=np.float64)


class DoubleWell(Potential):
    r"""One-dimensional double well :math:`V = a x^4 - b (x - c)^2` (well.py:26-81)."""

    def __init__(self, a: float = 1.0, b: float = 1.0, c: float = 0.0, desc: str = "1D double well potential"):
        super().__init__(dim=1, desc=desc)
        self.params = {"a": a, "b": b, "c": c}

    def _abc(self):
        return float(self.params["a"]), float(self.params["b"]), float(self.params["c"])

    def _energy_is_pass_independent(self, system) -> bool:
        return True  # one formula for V whatever else is computed (well.py)

    def _launch(self, system, pos, force, vw, flags):
        particles = system.particles
        shape = particles._pos.shape
        _lib.call(
            "tmd_doublewell", _device.dptr(pos), shape[0], shape[1], *self._abc(), flags,
            _device.dptr(force) if force is not None else None, _device.dptr(vw), _device.stream_ptr(),
        )

    def _host_call(self, system, flags):
        _device.require_cuda()
        pos = host_positions(system.particles)
        force = np.zeros(pos.shape) if flags & _lib.WANT_F else None
        vw = np.zeros(10)
        _lib.call(
            "tmd_host_doublewell", _lib.hptr(pos), pos.shape[0], pos.shape[1], *self._abc(), flags,
            _lib.hptr(force) if force is not None else None, _lib.hptr(vw),
        )
        return vw, force

    def potential(self, system) -> float:
        """Sum over all coordinates (well.py:65-72)."""
        vw, _ = self._host_call(system, _lib.WANT_V)
        return float(vw[0])

    def force(self, system):
        """Forces; the virial is not computed by this potential and stays zero (well.py:74-81)."""
        _, force = self._host_call(system, _lib.WANT_F)
        return force, np.zeros_like(system.particles.virial)


class RectangularWell(Potential):
    """0 inside (left, right), ``largenumber`` outside (well.py:84-137).  Host only."""

    def __init__(self, left: float = 0.0, right: float = 1.0, largenumber: float = float("inf"),
                 desc: str = "1D Rectangular well potential"):
        super().__init__(dim=1, desc=desc)
        self.params = {"left": left, "right": right, "largenumber": largenumber}
        if left >= right:
            LOGGER.warning("Setting left >= right in RectangularWell potential!")

    def potential(self, system) -> float:
        pos = system.particles.pos
        inside = np.logical_and(pos > self.params["left"], pos < self.params["right"])
        return np.where(inside, 0.0, self.params["largenumber"]).sum()

    def force(self, system):
        LOGGER.warning("Calling force for %s is not implemented!", self.desc)
        return np.zeros_like(system.particles.force), np.zeros_like(system.particles.virial)


class DoubleWellPair(Potential):
    r"""Pair double well :math:`V = h (1 - (r - r_0 - w)^2 / w^2)^2` between two particle types (well.py:140-286)."""

    def __init__(self, types: tuple[int, int], dim: int = 3, desc: str = "A double well pair potential"):
        super().__init__(dim=dim, desc=desc)
        self.params = {"height": 0.0, "height4": 0.0, "rwidth": 0.0, "rzero": 0.0, "width": 0.0, "width2": 0.0}
        self.types = types
        self._cache = None

    def set_parameters(self, parameters: dict[str, float]):
        """Take rzero / width / height and derive width2, rwidth, height4 (well.py:195-212)."""
        for key, value in parameters.items():
            if key in self.params:
                self.params[key] = value
            else:
                LOGGER.warning('Ignored unknown parameter "%s"', key)
        self.params["width2"] = self.params["width"] ** 2
        self.params["rwidth"] = self.params["rzero"] + self.params["width"]
        self.params["height4"] = 4.0 * self.params["height"]

    def min_max(self) -> tuple[float, float, float]:
        """(first minimum, second minimum, barrier top) (well.py:214-228)."""
        rzero, width = self.params["rzero"], self.params["width"]
        return rzero, rzero + 2.0 * width, rzero + width

    def activate(self, itype: int, jtype: int) -> bool:
        """Is the pair (itype, jtype) one this potential acts on? (well.py:230-234)"""
        return (itype == self.types[0] and jtype == self.types[1]) or (
            jtype == self.types[0] and itype == self.types[1])

    def _potential_function(self, delr):
        """V(r); also used with arrays for plotting (well.py:250-260)."""
        return self.params["height"] * (1.0 - (((delr - self.params["rwidth"]) ** 2) / self.params["width2"])) ** 2

    def _scalars(self):
        return float(self.params["rwidth"]), float(self.params["width2"]), float(self.params["height"])

    def _active(self, particles):
        """Device index lists of the atoms of types[0] / types[1] (cached per ptype array)."""
        ptype = particles.ptype
        cache = self._cache
        generation = getattr(particles, "_generation", 0)  # Particles.touch(): ptype was edited in place
        if cache is None or cache[0] is not ptype or cache[1] != (len(ptype), generation):
            idx0 = np.where(np.asarray(ptype) == self.types[0])[0].astype(np.int32)
            idx1 = np.where(np.asarray(ptype) == self.types[1])[0].astype(np.int32)
            dev0 = _device.to_device(idx0) if len(idx0) else None
            dev1 = _device.to_device(idx1) if len(idx1) else None
            cache = self._cache = (ptype, (len(ptype), generation), idx0, idx1, dev0, dev1)
        return cache[2:]

    def _energy_is_pass_independent(self, system) -> bool:
        return True  # one formula for V whatever else is computed (well.py)

    def _launch(self, system, pos, force, vw, flags):
        particles = system.particles
        idx0, idx1, dev0, dev1 = self._active(particles)
        box_abi = system.box.to_abi()  # tmd_box, or tmd_cell for a TriclinicBox
        name = "tmd_dwpair_triclinic" if isinstance(box_abi, _lib.Cell) else "tmd_dwpair"
        _lib.call(
            name, _device.dptr(pos), particles._pos.shape[0], C.byref(box_abi),
            _device.dptr(dev0) if dev0 is not None else None, len(idx0),
            _device.dptr(dev1) if dev1 is not None else None, len(idx1),
            1 if self.types[0] == self.types[1] else 0, *self._scalars(), flags,
            _device.dptr(force) if force is not None else None, _device.dptr(vw), _device.stream_ptr(),
        )

    def _host_call(self, system, flags):
        _device.require_cuda()
        particles = system.particles
        pos = host_positions(particles)
        ptype = np.ascontiguousarray(particles.ptype, dtype=np.int64)
        force = np.zeros(pos.shape) if flags & _lib.WANT_F else None
        vw = np.zeros(10)
        box_abi = system.box.to_abi()
        name = "tmd_host_dwpair_triclinic" if isinstance(box_abi, _lib.Cell) else "tmd_host_dwpair"
        _lib.call(
            name, _lib.hptr(pos), _lib.hptr(ptype), pos.shape[0], C.byref(box_abi),
            int(self.types[0]), int(self.types[1]), *self._scalars(), flags,
            _lib.hptr(force) if force is not None else None, _lib.hptr(vw),
        )
        return _host_result(vw, pos.shape[1], force)

    def potential(self, system) -> float:
        """Sum of V(r) over the active pairs (well.py:236-248)."""
        return self._host_call(system, _lib.WANT_V)[0]

    def force(self, system):
        """Forces and virial of the active pairs (well.py:262-286)."""
        _, force, virial = self._host_call(system, _lib.WANT_F | _lib.WANT_W)
        return force, virial
