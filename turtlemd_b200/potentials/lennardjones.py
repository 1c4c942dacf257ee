"""Cut (and optionally shifted) Lennard-Jones 12-6 pair potential on the GPU.

API of ``turtlemd/potentials/lennardjones.py``.  ``set_parameters`` keeps the
reference's host logic (mixing rules, pair table, the six coefficient
dictionaries); evaluation is the CUDA library:

* ``potential`` / ``force`` / ``potential_and_force`` take the system's host
  arrays through the host-buffer entry point ``tmd_host_lj`` -- this is the
  plug-in boundary a reference maintainer would bind;
* ``_launch`` enqueues ``tmd_lj_allpairs`` / ``tmd_lj_cells`` / ``tmd_lj_nlist`` on
  device-resident arrays for ``System.evaluate`` and the integrators.

Extra, optional constructor arguments (reference scripts run unchanged without
them): ``method`` = "auto" | "allpairs" | "cells" | "nlist" and ``skin`` (the
Verlet-list margin of the "nlist" method, in length units).  "auto" picks the
neighbour list for device-resident loops and the stateless cell list for
one-off host calls when the box holds >= 3 cells per direction and there are at
least ``CELLS_MIN_ATOMS`` atoms; otherwise all-pairs.  All methods evaluate the
same pairs (the list is rebuilt on the device whenever an atom moved more than
skin/2), so the choice changes speed, not results.
"""
from __future__ import annotations

import ctypes as C
import itertools
import logging
from typing import Any

import numpy as np

from .. import _device, _lib
from .potential import Potential
from .well import host_positions

LOGGER = logging.getLogger(__name__)
LOGGER.addHandler(logging.NullHandler())

_METHODS = {"auto": 0, "allpairs": 1, "cells": 2, "nlist": 2}  # value = tmd_host_lj method code
DEFAULT_SKIN = 0.3
CELLS_MIN_ATOMS = 4096  # "auto": below this the all-pairs kernel is used even when cells would fit


class LennardJonesCut(Potential):
    """Lennard-Jones 6-12 potential with a cut-off, per type-pair parameters."""

    def __init__(self, dim: int = 3, shift: bool = True, mixing: str = "geometric",
                 desc: str = "Lennard-Jones pair potential", method: str = "auto",
                 skin: float = DEFAULT_SKIN):
        super().__init__(dim=dim, desc=desc)
        if method not in _METHODS:
            raise ValueError(f'Unknown method "{method}", expected one of {sorted(_METHODS)}')
        self.shift = shift
        self.mixing = mixing
        self.method = method
        self.skin = float(skin)
        self._nlist = None  # tmd_nlist handle (device-resident path)
        self._lj1: dict = {}
        self._lj2: dict = {}
        self._lj3: dict = {}
        self._lj4: dict = {}
        self._rcut2: dict = {}
        self._offset: dict = {}
        self.params = {}
        self._cache = None  # (ptype array, types, table, host tidx, particles generation)

    def set_parameters(self, parameters: dict[Any, dict[str, float]]):
        """Atom/pair parameters -> the six coefficient tables (lennardjones.py:87-116)."""
        self.params = {}
        self._cache = None
        pair_param = generate_pair_interactions(parameters, self.mixing)
        for pair, par in pair_param.items():
            eps, sig, rcut = par["epsilon"], par["sigma"], par["rcut"]
            self._lj1[pair] = 48.0 * eps * sig**12
            self._lj2[pair] = 24.0 * eps * sig**6
            self._lj3[pair] = 4.0 * eps * sig**12
            self._lj4[pair] = 4.0 * eps * sig**6
            self._rcut2[pair] = rcut**2
            vcut = 0.0
            if self.shift:
                try:
                    vcut = 4.0 * eps * ((sig / rcut) ** 12 - (sig / rcut) ** 6)
                except ZeroDivisionError:
                    vcut = 0.0
            self._offset[pair] = vcut
            self.params[pair] = par

    def __str__(self):
        """Table of all pair parameters (lennardjones.py:118-143)."""
        lines = [self.desc, "Potential parameters, Lennard-Jones:",
                 f"Shift potential: {'yes' if self.shift else 'no'}", "Pair parameters:",
                 "{:12s} {:>9s} {:>9s} {:>9s}".format("Atom/pair", "epsilon", "sigma", "cut-off")]
        for pair in sorted(self.params):
            par = self.params[pair]
            lines.append("{:12s} {:>9.4f} {:>9.4f} {:>9.4f}".format(
                f"{pair[0]}-{pair[1]}", par["epsilon"], par["sigma"], np.sqrt(self._rcut2[pair])))
        return "\n".join(lines)

    # ---- host-side preparation of what the kernels consume -----------------------------------
    def _tables(self, particles):
        """(types, dense (6,T,T) table, dense int64 type index per atom) for this particle list.

        Raises ``KeyError`` for a type pair that occurs in the system but has no parameters, as
        the reference's per-pair dictionary lookup does (lennardjones.py:291-294).
        """
        ptype = particles.ptype
        cache = self._cache
        generation = getattr(particles, "_generation", 0)  # Particles.touch(): ptype was edited in place
        if cache is not None and cache[0] is ptype and len(cache[3]) == len(ptype) and cache[4] == generation:
            return cache[1], cache[2], cache[3]
        types, tidx, counts = np.unique(np.asarray(ptype), return_inverse=True, return_counts=True)
        types = [int(t) for t in types]
        ntyp = len(types)
        if ntyp > _lib.MAX_TYPES:
            raise NotImplementedError(
                f"{ntyp} particle types: the CUDA path supports at most {_lib.MAX_TYPES}")
        table = np.full((6, ntyp, ntyp), np.nan)
        for a, ta in enumerate(types):
            for b, tb in enumerate(types):
                if ta == tb and counts[a] < 2:
                    continue  # a lone atom of this type never meets its own kind
                if (ta, tb) not in self._rcut2:
                    raise KeyError((ta, tb))
                key = (ta, tb)
                table[:, a, b] = (self._lj1[key], self._lj2[key], self._lj3[key], self._lj4[key],
                                  self._offset[key], self._rcut2[key])
        tidx = np.ascontiguousarray(tidx.reshape(-1), dtype=np.int64)
        self._cache = (ptype, types, table, tidx, generation)
        return types, table, tidx

    def __del__(self):
        handle, self._nlist = getattr(self, "_nlist", None), None
        if handle is not None:
            try:
                _lib.load().tmd_nlist_destroy(handle)
            except Exception:  # interpreter shutdown
                pass

    def _nlist_handle(self):
        if self._nlist is None:
            handle = C.c_void_p()
            _lib.call("tmd_nlist_create", C.byref(handle), self.skin)
            self._nlist = handle
        return self._nlist

    def nlist_stats(self) -> dict:
        """{"builds", "calls", "slow_rows"} of the neighbour list (synchronises); zeros if unused."""
        builds, calls, slow = C.c_int64(0), C.c_int64(0), C.c_int32(0)
        if self._nlist is not None:
            _lib.call("tmd_nlist_stats", self._nlist, C.byref(builds), C.byref(calls), C.byref(slow))
        return {"builds": builds.value, "calls": calls.value, "slow_rows": slow.value}

    def time_traversals(self, on: bool = True) -> None:
        """Measurement: CUDA events around every traversal launch of a resident run (``tmd_nlist_set_timing``;
        a timed run synchronises at its end).  Read with :meth:`traversal_time`."""
        if self._nlist is not None:
            _lib.call("tmd_nlist_set_timing", self._nlist, 1 if on else 0)

    def traversal_time(self) -> tuple[float, int]:
        """(milliseconds spent in traversal launches, their number) since :meth:`time_traversals`."""
        ms, count = C.c_double(0.0), C.c_int64(0)
        if self._nlist is not None:
            _lib.call("tmd_nlist_timing", self._nlist, C.byref(ms), C.byref(count))
        return ms.value, count.value

    def _kernel_for(self, npart, box_abi, table, ntyp) -> str:
        """Entry point for the device-resident path."""
        if self.method == "allpairs":
            return "tmd_lj_allpairs"
        if self.method == "cells":
            self._use_cells(npart, box_abi, table, ntyp)
            return "tmd_lj_cells"
        usable = C.c_int(0)
        _lib.call("tmd_lj_nlist_usable", npart, C.byref(box_abi), _lib.hptr(table), ntyp, self.skin,
                  C.byref(usable))
        if self.method == "nlist":
            if not usable.value:
                raise NotImplementedError(
                    "neighbour list needs >= 3 cells of edge rcut+skin in every direction and a fully periodic box")
            return "tmd_lj_nlist"
        if usable.value and npart >= CELLS_MIN_ATOMS:
            return "tmd_lj_nlist"
        return "tmd_lj_cells" if self._use_cells(npart, box_abi, table, ntyp) else "tmd_lj_allpairs"

    def _energy_is_pass_independent(self, system) -> bool:
        """The neighbour-list traversal forms r6inv the same way whether or not forces are asked for
        (it ignores TMD_V_STANDALONE), so V of a combined pass equals V of an energy-only pass bit for
        bit.  The all-pairs and cell kernels follow the reference's two formulas (lennardjones.py:172
        vs :254-255), which differ in the last place: there the two passes are kept."""
        particles = system.particles
        if particles.npart < 2:
            return True
        key = (particles.npart, tuple(np.asarray(system.box.length, dtype=float).tolist()), id(particles.ptype), self.method,
               getattr(particles, "_generation", 0))
        cached = getattr(self, "_pass_cache", None)
        if cached is None or cached[0] != key:
            types, table, _ = self._tables(particles)
            box_abi = system.box.to_abi()
            name = ("tmd_lj_triclinic" if isinstance(box_abi, _lib.Cell)
                    else self._kernel_for(particles.npart, box_abi, table, len(types)))
            cached = self._pass_cache = (key, name == "tmd_lj_nlist")
        return cached[1]

    def _use_cells(self, npart, box_abi, table, ntyp) -> bool:
        if self.method == "allpairs":
            return False
        usable = C.c_int(0)
        _lib.call("tmd_lj_cells_usable", npart, C.byref(box_abi), _lib.hptr(table), ntyp, C.byref(usable))
        if self.method in ("cells", "nlist"):
            if not usable.value:
                raise NotImplementedError(
                    "cell list needs >= 3 cells of edge rcut in every direction and a fully periodic box")
            return True
        return bool(usable.value) and npart >= CELLS_MIN_ATOMS

    # ---- device path ----------------------------------------------------------------------------------
    def _launch(self, system, pos, force, vw, flags, first=0, count=None):
        particles = system.particles
        npart = particles.npart
        if npart < 2:
            # no pair: zero contribution, but overwriting callers still need defined outputs
            if force is not None and not flags & _lib.ACCUMULATE:
                force.zero_()
            if not flags & _lib.ACCUMULATE:
                if flags & _lib.WANT_V:
                    vw[0:1].zero_()
                if flags & _lib.WANT_W:
                    vw[1:].zero_()
            return
        types, table, tidx = self._tables(particles)
        ntyp = len(types)
        box_abi = system.box.to_abi()
        tidx_dev = particles._static_device(f"lj_tidx_{id(self)}", tidx, np.int32) if ntyp > 1 else None
        count = npart if count is None else count
        if isinstance(box_abi, _lib.Cell):  # TriclinicBox: the all-pairs kernel with the exact image scan
            _lib.call(
                "tmd_lj_triclinic", _device.dptr(pos), _device.dptr(tidx_dev) if tidx_dev is not None else None,
                npart, C.byref(box_abi), _lib.hptr(table), ntyp, flags, first, count,
                _device.dptr(force) if force is not None else None, _device.dptr(vw), _device.stream_ptr(),
            )
            return
        name = self._kernel_for(npart, box_abi, table, ntyp)
        handle = (self._nlist_handle(),) if name == "tmd_lj_nlist" else ()
        _lib.call(
            name, *handle, _device.dptr(pos), _device.dptr(tidx_dev) if tidx_dev is not None else None, npart,
            C.byref(box_abi), _lib.hptr(table), ntyp, flags, first, count,
            _device.dptr(force) if force is not None else None, _device.dptr(vw), _device.stream_ptr(),
        )

    def _resident_args(self, system):
        """(handle, tidx, n, box, table, ntypes) for the fused resident run (``tmd_nlist_run_vv``) or None when
        this system does not take the neighbour-list path (small, not fully periodic, triclinic)."""
        particles = system.particles
        npart = particles.npart
        if npart < 2:
            return None
        box_abi = system.box.to_abi()
        if isinstance(box_abi, _lib.Cell):
            return None
        types, table, tidx = self._tables(particles)
        ntyp = len(types)
        if self._kernel_for(npart, box_abi, table, ntyp) != "tmd_lj_nlist":
            return None
        tidx_dev = particles._static_device(f"lj_tidx_{id(self)}", tidx, np.int32) if ntyp > 1 else None
        return self._nlist_handle(), tidx_dev, npart, box_abi, table, ntyp

    # ---- the plug-in boundary: host arrays in, host arrays out ----------------------------------------
    def _host_call(self, system, flags):
        particles = system.particles
        pos = host_positions(particles)
        dim = pos.shape[1]
        force = np.zeros(pos.shape) if flags & _lib.WANT_F else None
        vw = np.zeros(10)
        if particles.npart >= 2:
            _device.require_cuda()
            types, table, tidx = self._tables(particles)
            box_abi = system.box.to_abi()
            if isinstance(box_abi, _lib.Cell):
                _lib.call(
                    "tmd_host_lj_triclinic", _lib.hptr(pos), _lib.hptr(tidx), particles.npart, C.byref(box_abi),
                    _lib.hptr(table), len(types), flags, _lib.hptr(force) if force is not None else None,
                    _lib.hptr(vw),
                )
            else:
                _lib.call(
                    "tmd_host_lj", _lib.hptr(pos), _lib.hptr(tidx), particles.npart, C.byref(box_abi),
                    _lib.hptr(table), len(types), flags, _METHODS[self.method],
                    _lib.hptr(force) if force is not None else None, _lib.hptr(vw),
                )
        virial = vw[1:].reshape(3, 3)[:dim, :dim].copy()
        return float(vw[0]), force, virial

    def potential(self, system):
        """Potential energy (lennardjones.py:145-183)."""
        return self._host_call(system, _lib.WANT_V | _lib.V_STANDALONE)[0]

    def force(self, system):
        """Forces and virial (lennardjones.py:185-224)."""
        _, force, virial = self._host_call(system, _lib.WANT_F | _lib.WANT_W)
        return force, virial

    def potential_and_force(self, system):
        """Energy, forces and virial in one pass (lennardjones.py:226-273)."""
        return self._host_call(system, _lib.WANT_V | _lib.WANT_F | _lib.WANT_W)


def generate_pair_interactions(parameters: dict[Any, dict[str, float]], mixing: str):
    """Pair parameters from atom parameters; explicit ``(i, j)`` entries win (lennardjones.py:297-362).

    Like the reference, tuple keys are not filtered out of the "atom" list, so entries such as
    ``((0, 1), 0)`` appear in the result; only ``(int, int)`` keys are ever looked up.
    """
    atoms = list(parameters)
    pairs: dict = {}
    for atom_i, atom_j in itertools.product(atoms, atoms):
        if (atom_i, atom_j) in parameters:
            pairs[atom_i, atom_j] = dict(parameters[atom_i, atom_j])
            pairs[atom_j, atom_i] = pairs[atom_i, atom_j]
        elif (atom_j, atom_i) in parameters:
            pairs[atom_j, atom_i] = dict(parameters[atom_j, atom_i])
            pairs[atom_i, atom_j] = pairs[atom_j, atom_i]
        elif atom_i == atom_j:
            own = parameters[atom_i]
            pairs[atom_i, atom_i] = {"epsilon": own["epsilon"], "sigma": own["sigma"], "rcut": own["rcut"]}
        else:
            par_i, par_j = parameters[atom_i], parameters[atom_j]
            eps, sig, rcut = mix_parameters(
                par_i["epsilon"], par_i["sigma"], par_i["rcut"],
                par_j["epsilon"], par_j["sigma"], par_j["rcut"], mixing=mixing,
            )
            pairs[atom_i, atom_j] = {"epsilon": eps, "sigma": sig, "rcut": rcut}
            pairs[atom_j, atom_i] = pairs[atom_i, atom_j]
    return pairs


def mix_parameters(epsilon_i: float, sigma_i: float, rcut_i: float, epsilon_j: float, sigma_j: float,
                   rcut_j: float, mixing: str = "geometric") -> tuple[float, float, float]:
    """Mixed (epsilon, sigma, rcut): geometric, arithmetic or sixthpower (lennardjones.py:365-461)."""
    if mixing == "geometric":
        return (np.sqrt(epsilon_i * epsilon_j), np.sqrt(sigma_i * sigma_j), np.sqrt(rcut_i * rcut_j))
    if mixing == "arithmetic":
        return (np.sqrt(epsilon_i * epsilon_j), 0.5 * (sigma_i + sigma_j), 0.5 * (rcut_i + rcut_j))
    if mixing == "sixthpower":
        si3, sj3 = sigma_i**3, sigma_j**3
        si6, sj6 = si3**2, sj3**2
        avgs6 = 0.5 * (si6 + sj6)
        return (np.sqrt(epsilon_i * epsilon_j) * si3 * sj3 / avgs6, avgs6 ** (1.0 / 6.0),
                (0.5 * (rcut_i**6 + rcut_j**6)) ** (1.0 / 6.0))
    raise ValueError(f'Uknown mixing rule "{mixing}" requested!')
