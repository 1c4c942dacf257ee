"""The system: box + particles + force field (API of ``turtlemd/system/system.py``).

``potential_and_force`` / ``force`` / ``potential`` launch every potential's
kernels into ONE device force buffer and ONE (V, W) buffer (first potential
overwrites, the others accumulate), which is the sum the reference builds with
NumPy at system.py:58-99.  The integrators call :meth:`System.evaluate`, which
leaves everything in HBM; the three public methods additionally hand the
results back as host arrays, like the reference.
"""
from __future__ import annotations

from typing import Any, TypedDict

import numpy as np

from .. import _device, _lib
from .particles import kinetic_energy, kinetic_temperature, pressure_tensor


class Thermo(TypedDict):
    ekin: None | float
    pressure: None | float
    temperature: None | float
    vpot: None | float


class System:
    """Box + particles + potentials."""

    def __init__(self, box, particles, potentials: list[Any] | None = None):
        self.box = box
        self.particles = particles
        self.potentials = list(potentials) if potentials is not None else []

    # ---- device path --------------------------------------------------------------------------
    def evaluate(self, flags: int) -> bool:
        """Run all potentials on the device for ``flags`` (TMD_WANT_V|F|W).

        Results stay in HBM: ``particles._force`` (if F was asked for) and the
        ``vw`` buffer.  Returns False when there is no potential (the reference
        then leaves force/virial untouched and sets v_pot = 0, system.py:70-75).
        """
        particles = self.particles
        if not self.potentials:
            if flags & _lib.WANT_V:
                particles.v_pot = 0
            return False
        _device.require_cuda()
        size = int(np.prod(particles._pos.shape))
        pos = particles._pos.device(read_only=True)  # potentials never write positions
        force = particles._force.device_uninitialised(size) if flags & _lib.WANT_F else None
        vw = particles._vw_buffer()
        first = True
        for pot in self.potentials:
            launch = getattr(pot, "_launch", None)
            if launch is None:
                self._foreign_potential(pot, flags, first, force, vw)
            else:
                launch(self, pos, force, vw, flags | (0 if first else _lib.ACCUMULATE))
            first = False
        if flags & _lib.WANT_F:
            particles._force.written_on_device(shape=particles._pos.shape)
        if flags & _lib.WANT_V:
            particles._vw_fresh.add("v_pot")
        if flags & _lib.WANT_W:
            particles._vw_fresh.add("virial")
        return True

    def energy_rides_with_forces(self) -> bool:
        """True when evaluating V in the same pass as F and W gives the very numbers a separate
        ``potential()`` pass would (every potential says so through ``_energy_is_pass_independent``).
        LangevinInertia evaluates ``force()`` and then ``potential()`` at the SAME positions
        (integrators.py:280-282); when this holds the second traversal is folded into the first."""
        if not self.potentials:
            return False
        for pot in self.potentials:
            check = getattr(pot, "_energy_is_pass_independent", None)
            if check is None or not check(self):
                return False
        return True

    def _foreign_potential(self, pot, flags, first, force, vw):
        """A user-defined host ``Potential`` in the list: call it and add its result on the device."""
        want_f = bool(flags & (_lib.WANT_F | _lib.WANT_W))
        vpot, frc, vir = 0.0, None, None
        if (flags & _lib.WANT_V) and want_f:
            vpot, frc, vir = pot.potential_and_force(self)
        elif want_f:
            frc, vir = pot.force(self)
        else:
            vpot = pot.potential(self)
        add = np.zeros(10)
        add[0] = vpot
        if vir is not None:
            block = np.zeros((3, 3))
            block[: vir.shape[0], : vir.shape[1]] = vir
            add[1:] = block.reshape(-1)
        mask = np.zeros(10)
        if flags & _lib.WANT_V:
            mask[0] = 1.0
        if flags & _lib.WANT_W:
            mask[1:] = 1.0
        add_dev = _device.to_device(add * mask)
        if first:
            keep = _device.to_device(1.0 - mask)
            vw.mul_(keep).add_(add_dev)
        else:
            vw.add_(add_dev)
        if flags & _lib.WANT_F:
            frc_dev = _device.to_device(np.asarray(frc, dtype=float))
            if first:
                force.copy_(frc_dev)
            else:
                force.add_(frc_dev)

    # ---- the reference API ----------------------------------------------------------------------
    def potential_and_force(self):
        """Energy, forces and virial (system.py:58-76)."""
        if not self.evaluate(_lib.WANT_V | _lib.WANT_F | _lib.WANT_W):
            return 0, None, None
        return self.particles.v_pot, self.particles.force, self.particles.virial

    def potential(self):
        """Energy only (system.py:78-82)."""
        self.evaluate(_lib.WANT_V | _lib.V_STANDALONE)
        return self.particles.v_pot

    def force(self):
        """Forces and virial (system.py:84-99)."""
        if not self.evaluate(_lib.WANT_F | _lib.WANT_W):
            return None, None
        return self.particles.force, self.particles.virial

    def thermo(self, boltzmann: float = 1.0) -> Thermo:
        """Per-particle energies, pressure and temperature (system.py:101-130)."""
        thermo: Thermo = {"ekin": None, "pressure": None, "temperature": None, "vpot": None}
        particles = self.particles
        if particles is None or particles.npart == 0:
            return thermo
        vpot = particles.v_pot
        if vpot is not None:
            vpot = vpot / particles.npart
        thermo["vpot"] = vpot
        kin_tensor, ekin = kinetic_energy(particles)
        thermo["ekin"] = ekin / particles.npart
        _, pressure = pressure_tensor(particles, self.box.volume(), kin_tensor=kin_tensor)
        thermo["pressure"] = pressure
        dof = getattr(self.box, "dof", None)
        temp, _, _ = kinetic_temperature(particles, boltzmann, dof=dof, kin_tensor=kin_tensor)
        thermo["temperature"] = float(temp)
        return thermo
