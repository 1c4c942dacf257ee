"""GPU parity of the fused, device-resident VelocityVerlet run (``tmd_nlist_run_vv``: one traversal per step
whose epilogue does kick + kick + drift, state in the neighbour list's cell order, rebuilds decided on the
device) against the step-by-step path (``tmd_vv_kick_drift`` + ``tmd_lj_nlist`` + ``tmd_vv_kick``).

Both paths share the list handle, the build rule and every rounding, so the comparison is BIT-exact; the
step-by-step path is itself pinned to the reference-generated trajectories in ``test_gpu_parity.py``.
Reference: ``turtlemd/integrators.py:84-94`` over ``potentials/lennardjones.py:226-273``.
"""
from __future__ import annotations

import numpy as np
import pytest

from conftest import load_golden, rel_norm

pytestmark = pytest.mark.gpu

SINGLE = {0: {"sigma": 1.0, "epsilon": 1.0, "rcut": 2.5}}
TWO = {0: {"sigma": 1.0, "epsilon": 1.0, "rcut": 2.5}, 1: {"sigma": 1.1, "epsilon": 0.8, "rcut": 2.7}}


def _fluid(rep, dim=3, seed=3):
    from turtlemd_b200.tools import generate_lattice

    if dim == 3:
        pos, size = generate_lattice("fcc", [rep] * 3, density=0.8)
    else:
        pos, size = generate_lattice("sq", [rep] * 2, density=0.7)
    pos = pos + 0.05 * (2.0 * np.random.default_rng(seed).random(pos.shape) - 1.0)
    return pos, size


def _system(pos, size, vel, params, ptype=None, mass=None, skin=0.3):
    from turtlemd_b200.potentials import LennardJonesCut
    from turtlemd_b200.system import Box, Particles, System

    pot = LennardJonesCut(dim=pos.shape[1], shift=True, method="nlist", skin=skin)
    pot.set_parameters(params)
    part = Particles.from_arrays(pos, vel=vel, ptype=ptype, mass=mass)
    return System(Box(low=size[:, 0], high=size[:, 1]), part, [pot]), pot


def _run_both(pos, size, vel, params, chunks, dt=0.004, **kw):
    import turtlemd_b200.integrators as integrators

    out = {}
    for fused in (False, True):
        integrators.RESIDENT_RUNS = fused
        try:
            system, pot = _system(pos, size, vel, params, **kw)
            system.potential_and_force()
            integ = integrators.VelocityVerlet(dt)
            for k in chunks:
                integ.run(system, k)
            p = system.particles
            out[fused] = (p.pos.copy(), p.vel.copy(), p.force.copy(), p.v_pot, p.virial.copy(), pot.nlist_stats())
        finally:
            integrators.RESIDENT_RUNS = True
    return out[False], out[True]


def _assert_identical(a, b):
    for name, x, y in zip(("pos", "vel", "force"), a[:3], b[:3]):
        assert np.array_equal(x, y), (name, float(np.max(np.abs(x - y))))
    assert a[3] == b[3], (a[3], b[3])
    assert np.array_equal(a[4], b[4])
    assert a[5]["builds"] == b[5]["builds"] and a[5]["calls"] == b[5]["calls"], (a[5], b[5])


def test_fused_run_is_bit_identical_to_single_steps_one_type():
    """N = 4000 hot fluid, 60 steps in runs of 7 + 13 + 40: several device-decided rebuilds inside the runs."""
    pos, size = _fluid(10)
    vel = np.random.default_rng(4).standard_normal(pos.shape) * 1.2
    vel -= vel.mean(axis=0)
    a, b = _run_both(pos, size, vel, SINGLE, (7, 13, 40))
    _assert_identical(a, b)
    assert a[5]["builds"] >= 3, a[5]


def test_fused_run_two_types_masses_and_unwrapped_coordinates():
    """Type table in shared memory, per-atom masses, atoms that start several boxes away from the cell."""
    rng = np.random.default_rng(11)
    pos, size = _fluid(10)
    n = len(pos)
    length = size[:, 1] - size[:, 0]
    pos = pos + rng.integers(-2, 3, size=pos.shape) * length
    vel = rng.standard_normal(pos.shape)
    ptype = (rng.random(n) < 0.3).astype(int)
    mass = np.where(ptype == 1, 2.5, 1.0).reshape(-1, 1)
    a, b = _run_both(pos, size, vel, TWO, (5, 31), ptype=ptype, mass=mass)
    _assert_identical(a, b)
    assert a[5]["builds"] >= 2, a[5]


def test_fused_run_two_dimensions():
    pos, size = _fluid(70, dim=2)
    vel = np.random.default_rng(5).standard_normal(pos.shape) * 0.8
    a, b = _run_both(pos, size, vel, SINGLE, (25,), dt=0.003)
    _assert_identical(a, b)


def test_fused_run_on_rows_that_outgrew_the_list():
    """Half-empty box: most rows take the exact slow path (cells walked at traversal time)."""
    from turtlemd_b200.tools import generate_lattice

    pos, size = generate_lattice("fcc", [8, 8, 16], density=1.0)
    pos = pos + 0.04 * (2.0 * np.random.default_rng(8).random(pos.shape) - 1.0)
    size = size.copy()
    size[2, 1] *= 2.6
    vel = np.random.default_rng(9).standard_normal(pos.shape) * 0.3
    a, b = _run_both(pos, size, vel, SINGLE, (12,), dt=0.002, skin=0.4)
    _assert_identical(a, b)
    assert b[5]["slow_rows"] > 100, b[5]


def test_fused_run_follows_the_reference_trajectory():
    """The reference's own VelocityVerlet trajectory of a hot 864-atom fluid (tests/golden/make_golden_r2.py:
    frames 0 / 4 / 8 / 12 of an MDSimulation run) through the fused path, 4 steps per run() call; the list is
    rebuilt on the way.  Tolerance of the north star: 1e-8 on trajectories."""
    import turtlemd_b200.integrators as integrators

    g = load_golden("traj_vv_lj_864_hot.npz")
    size = np.stack([g["low"], g["high"]], axis=1)
    system, pot = _system(g["pos0"], size, g["vel0"], SINGLE)
    system.potential_and_force()
    integ = integrators.VelocityVerlet(float(g["dt"]))
    for k, frame in enumerate(g["frames"]):
        if frame:
            integ.run(system, int(frame - g["frames"][k - 1]))
        p = system.particles
        assert rel_norm(p.pos, g["pos"][k]) < 1e-8, frame
        assert rel_norm(p.vel, g["vel"][k]) < 1e-8, frame
        assert rel_norm(p.force, g["force"][k]) < 1e-8, frame
        assert abs(p.v_pot - g["v_pot"][k]) <= 1e-8 * abs(g["v_pot"][k]), frame
        assert rel_norm(p.virial, g["virial"][k]) < 1e-8, frame
    assert pot.nlist_stats()["builds"] >= 2


def test_fused_run_energy_conservation_at_benchmark_size():
    """N = 1 000 188 (BASELINE.json's metric size): 40 fused steps conserve energy like the step-by-step path and
    end on the same state bit for bit (one rebuild or more on the way)."""
    import turtlemd_b200.integrators as integrators
    from turtlemd_b200.system.particles import kinetic_energy

    pos, size = _fluid(63)
    vel = np.random.default_rng(1).standard_normal(pos.shape) * np.sqrt(0.8)
    vel -= vel.mean(axis=0)
    res = {}
    for fused in (False, True):
        integrators.RESIDENT_RUNS = fused
        try:
            system, pot = _system(pos, size, vel, SINGLE)
            system.potential_and_force()
            e0 = system.particles.v_pot + kinetic_energy(system.particles)[1]
            integrators.VelocityVerlet(0.005).run(system, 40)
            e1 = system.particles.v_pot + kinetic_energy(system.particles)[1]
            res[fused] = (system.particles.pos.copy(), system.particles.vel.copy(), e0, e1, pot.nlist_stats())
        finally:
            integrators.RESIDENT_RUNS = True
    assert np.array_equal(res[False][0], res[True][0]) and np.array_equal(res[False][1], res[True][1])
    e0, e1 = res[True][2], res[True][3]
    assert abs(e1 - e0) < 2e-4 * abs(e0), (e0, e1)
    assert res[True][4]["builds"] >= 2


# ---- LangevinInertia: the resident run (tmd_nlist_run_langevin) -------------------------------------------------


def _run_both_langevin(pos, size, vel, params, chunks, dt=0.004, **kw):
    import turtlemd_b200.integrators as integrators
    from turtlemd_b200.philox import PhiloxNormal

    out = {}
    for fused in (False, True):
        integrators.RESIDENT_RUNS = integrators.RESIDENT_LANGEVIN = fused
        try:
            system, pot = _system(pos, size, vel, params, **kw)
            system.potential_and_force()
            integ = integrators.LangevinInertia(timestep=dt, gamma=0.7, beta=1.0 / 2.0, rgen=PhiloxNormal(key=23))
            for k in chunks:
                integ.run(system, k)
            p = system.particles
            out[fused] = (p.pos.copy(), p.vel.copy(), p.force.copy(), p.v_pot, p.virial.copy(), pot.nlist_stats(),
                          integ.rgen.position)
        finally:
            integrators.RESIDENT_RUNS, integrators.RESIDENT_LANGEVIN = True, False
    return out[False], out[True]


def test_resident_langevin_run_is_bit_identical_to_the_step_by_step_path():
    """integrators.py:265-282 over lennardjones.py:226-273: one pass per step over slot-ordered state vs the
    element-wise kernels on atom-ordered arrays.  Same lists, same roundings, noise addressed by atom index: bit-identical,
    through several rebuilds and a second run that starts from the state the first one left."""
    pos, size = _fluid(12)
    vel = np.random.default_rng(5).standard_normal(pos.shape) * 1.4
    a, b = _run_both_langevin(pos, size, vel, SINGLE, [21, 14])
    _assert_identical(a, b)
    assert a[6] == b[6] == 35 * 2 * len(pos) * 3  # the stream moved on by 2 n dim normals per step
    assert b[5]["builds"] >= 3


def test_resident_langevin_run_two_types_masses_two_dimensions():
    pos, size = _fluid(12)
    rng = np.random.default_rng(8)
    vel = rng.standard_normal(pos.shape)
    ptype = (np.arange(len(pos)) % 4 == 0).astype(int)
    mass = 1.0 + 0.25 * (np.arange(len(pos)) % 3)
    a, b = _run_both_langevin(pos, size, vel, TWO, [17], ptype=ptype, mass=mass)
    _assert_identical(a, b)
    pos2, size2 = _fluid(48, dim=2)
    vel2 = rng.standard_normal(pos2.shape)
    a2, b2 = _run_both_langevin(pos2, size2, vel2, SINGLE, [15])
    _assert_identical(a2, b2)


# ---- a DoubleWellPair as the second potential of the fused run (tmd_nlist_run_vv_dwpair) -----------------------------


def _dimer_system(pos, size, vel, ptype, types, skin=0.3):
    from turtlemd_b200.potentials import DoubleWellPair, LennardJonesCut
    from turtlemd_b200.system import Box, Particles, System

    one = {"sigma": 1.0, "epsilon": 1.0, "rcut": 2.5}
    lj = LennardJonesCut(dim=pos.shape[1], shift=True, method="nlist", skin=skin)
    lj.set_parameters({0: dict(one), 1: dict(one), 2: dict(one)})
    bond = DoubleWellPair(types=types, dim=pos.shape[1])
    bond.set_parameters({"rzero": 2.0 ** (1.0 / 6.0), "height": 6.0, "width": 0.25})
    part = Particles.from_arrays(pos, vel=vel, ptype=ptype)
    return System(Box(low=size[:, 0], high=size[:, 1]), part, [lj, bond]), lj


@pytest.mark.parametrize("types", [(1, 1), (1, 2)])
def test_fused_run_with_a_bonded_pair_is_bit_identical_to_single_steps(types):
    """BASELINE.json configs[4] in small: LennardJonesCut + DoubleWellPair (well.py:230-286) under VelocityVerlet.  The fused
    run evaluates the bonded atoms from the slot-ordered state and adds their forces in the traversal's epilogue; forces,
    energy and virial are System.potential_and_force's sums (system.py:58-76), bit for bit, through list rebuilds."""
    import turtlemd_b200.integrators as integrators

    pos, size = _fluid(12)
    rng = np.random.default_rng(21)
    vel = rng.standard_normal(pos.shape) * 1.3
    ptype = np.zeros(len(pos), dtype=int)
    first = len(pos) // 2
    near = np.argsort(np.sum((pos - pos[first]) ** 2, axis=1))[1:4]
    ptype[first] = 1
    ptype[near[0]] = types[1]
    if types[0] != types[1]:
        ptype[near[1]] = 1  # two atoms of the first type, one of the second: two bonded pairs
    out = {}
    for fused in (False, True):
        integrators.RESIDENT_RUNS = fused
        try:
            system, lj = _dimer_system(pos, size, vel, ptype, types)
            system.potential_and_force()
            integ = integrators.VelocityVerlet(0.004)
            for k in (19, 12):
                integ.run(system, k)
            p = system.particles
            out[fused] = (p.pos.copy(), p.vel.copy(), p.force.copy(), p.v_pot, p.virial.copy(), lj.nlist_stats())
        finally:
            integrators.RESIDENT_RUNS = True
    _assert_identical(out[False], out[True])
    assert out[True][5]["builds"] >= 2


def test_fused_run_with_a_bonded_pair_follows_the_reference_trajectory():
    """traj_vv_dimer_864.npz (the reference's own run) through the fused path."""
    g = load_golden("traj_vv_dimer_864.npz")
    from turtlemd_b200.integrators import VelocityVerlet
    from turtlemd_b200.potentials import DoubleWellPair, LennardJonesCut
    from turtlemd_b200.system import Box, Particles, System

    one = {"sigma": 1.0, "epsilon": 1.0, "rcut": 2.5}
    lj = LennardJonesCut(dim=3, shift=True, method="nlist")
    lj.set_parameters({0: dict(one), 1: dict(one)})
    bond = DoubleWellPair(types=(1, 1))
    bond.set_parameters({"rzero": 2.0 ** (1.0 / 6.0), "height": 6.0, "width": 0.25})
    part = Particles.from_arrays(g["pos0"], vel=g["vel0"], mass=g["mass"], ptype=g["ptype"])
    system = System(Box(low=g["low"], high=g["high"]), part, [lj, bond])
    system.potential_and_force()
    integ = VelocityVerlet(float(g["dt"]))
    integ.run(system, int(g["frames"][-1]))
    p = system.particles
    assert rel_norm(p.pos, g["pos"][-1]) < 1e-8 and rel_norm(p.vel, g["vel"][-1]) < 1e-8
    assert rel_norm(p.force, g["force"][-1]) < 1e-8
    assert abs(p.v_pot - g["v_pot"][-1]) <= 1e-8 * abs(g["v_pot"][-1])
    assert rel_norm(p.virial, g["virial"][-1]) < 1e-8
