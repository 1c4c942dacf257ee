"""GPU parity of the resident multi-GPU slab run (``csrc/slabrun.cuh``, ``parallel.SlabRun``) against the
single-GPU run of the same steps.  Several ranks are emulated on ONE device, phase by phase
(``parallel.SlabRunEmulation``): atom migration between slabs, ghost layers, the halo stores into the
neighbours' record buffers and the device-side rebuild gates all run exactly as on real ranks; only the spin
of the barrier is never exercised (all signals of a barrier are enqueued before any wait).  Real ranks:
``tests/multi_gpu_check.py slabrun`` (two GPUs).  Reference: ``turtlemd/integrators.py:84-94`` over
``potentials/lennardjones.py:226-273``; the two runs differ in summation order only.
"""
from __future__ import annotations

import numpy as np
import pytest

from conftest import rel_norm

pytestmark = pytest.mark.gpu

SINGLE = {0: {"sigma": 1.0, "epsilon": 1.0, "rcut": 2.5}}
TWO = {0: {"sigma": 1.0, "epsilon": 1.0, "rcut": 2.5}, 1: {"sigma": 1.05, "epsilon": 0.9, "rcut": 2.5}}


def _make(rep, params, temperature, seed=3, masses=False):
    from turtlemd_b200.potentials import LennardJonesCut
    from turtlemd_b200.system import Box, Particles, System
    from turtlemd_b200.tools import generate_lattice

    pos, size = generate_lattice("fcc", [rep] * 3, density=0.8)
    rng = np.random.default_rng(seed)
    pos = pos + 0.05 * (2.0 * rng.random(pos.shape) - 1.0)
    pos += rng.integers(-2, 3, size=pos.shape) * (size[:, 1] - size[:, 0])  # unwrapped coordinates are legal input
    vel = rng.standard_normal(pos.shape) * np.sqrt(temperature)
    vel -= vel.mean(axis=0)
    ptype = (np.arange(len(pos)) % 5 == 0).astype(int) if len(params) > 1 else None
    mass = (1.0 + 0.5 * (np.arange(len(pos)) % 3)) if masses else None

    def build():
        pot = LennardJonesCut(dim=3, shift=True, method="nlist")
        pot.set_parameters(params)
        part = Particles.from_arrays(pos, vel=vel, ptype=ptype, mass=mass)
        system = System(Box(low=size[:, 0], high=size[:, 1]), part, [pot])
        system.potential_and_force()
        return system

    return build


def _compare(build, world, chunks, dt=0.004, langevin=False):
    from turtlemd_b200.integrators import LangevinInertia, VelocityVerlet
    from turtlemd_b200.parallel import SlabRunEmulation
    from turtlemd_b200.philox import PhiloxNormal

    def integrator():
        if langevin:
            return LangevinInertia(timestep=dt, gamma=0.5, beta=1.0 / 1.5, rgen=PhiloxNormal(key=17))
        return VelocityVerlet(dt)

    ref = build()
    integ = integrator()
    system = build()
    emu = SlabRunEmulation([system] * world, integrator())
    try:
        for k in chunks:
            integ.run(ref, k)
            emu.run(k)
        x, v, f, v_pot, virial = emu.gather_state()
        status = emu.status()
    finally:
        emu.release()
    q = ref.particles
    assert rel_norm(x, q.pos) < 1e-11
    assert rel_norm(v, q.vel) < 1e-9
    assert rel_norm(f, q.force) < 1e-9
    assert abs(v_pot - q.v_pot) <= 1e-10 * abs(q.v_pot)
    assert rel_norm(virial, q.virial) < 1e-9
    assert sum(s["n_own"] for s in status) == len(x)
    return status


@pytest.mark.parametrize("world", [2, 3])
def test_emulated_slab_ranks_follow_the_single_gpu_trajectory(world):
    """N = 6912 (L = 20.5: six cell layers along z), hot enough that the lists are rebuilt several times and atoms
    migrate between the slabs inside the run; two chunks, so a second start() follows a completed run."""
    status = _compare(_make(12, SINGLE, temperature=3.0), world, [25, 20])
    assert all(s["builds"] >= 2 for s in status), status
    assert all(s["layers"] % world == 0 and s["layers_per_rank"] >= 2 for s in status)


def test_emulated_slab_ranks_two_types_and_masses():
    _compare(_make(12, TWO, temperature=2.0, masses=True), 2, [30])


@pytest.mark.parametrize("world", [2, 3])
def test_emulated_slab_ranks_langevin_inertia(world):
    """BASELINE.json configs[3] in small: LangevinInertia on slabs.  The noise of a coordinate is addressed by its global
    atom index, so the decomposed run retraces the single-GPU ``LangevinInertia.run`` (integrators.py:265-282)."""
    status = _compare(_make(12, SINGLE, temperature=1.5, masses=True), world, [22, 15], langevin=True)
    assert all(s["builds"] >= 1 for s in status), status


def test_four_emulated_slab_ranks_two_types_langevin():
    """N = 23 328 (L = 30.8: eight cell layers along z, two per rank): four ranks, two particle types with different
    Lennard-Jones parameters and three masses under LangevinInertia -- every rank has a different neighbour above and below."""
    status = _compare(_make(18, TWO, temperature=1.2, masses=True), 4, [24], langevin=True)
    assert all(s["layers_per_rank"] == 2 and s["layers"] == 8 for s in status), status
    assert all(s["builds"] >= 1 for s in status)


@pytest.mark.parametrize("world,types", [(2, (1, 1)), (3, (1, 2))])
def test_emulated_slab_ranks_with_a_bonded_pair(world, types):
    """BASELINE.json configs[4] in small on slabs: LennardJonesCut + DoubleWellPair (well.py:230-286) under VelocityVerlet.
    The bonded atoms sit next to a slab boundary (the middle of the box is one for two ranks), so they are owned by
    different ranks or migrate during the run; their positions travel through the replicated table."""
    from turtlemd_b200.integrators import VelocityVerlet
    from turtlemd_b200.parallel import SlabRunEmulation
    from turtlemd_b200.potentials import DoubleWellPair, LennardJonesCut
    from turtlemd_b200.system import Box, Particles, System
    from turtlemd_b200.tools import generate_lattice

    pos, size = generate_lattice("fcc", [12] * 3, density=0.8)
    rng = np.random.default_rng(31)
    pos = pos + 0.05 * (2.0 * rng.random(pos.shape) - 1.0)
    vel = rng.standard_normal(pos.shape) * 1.4
    vel -= vel.mean(axis=0)
    length = size[:, 1] - size[:, 0]
    centre = np.argmin(np.sum((pos - 0.5 * length) ** 2, axis=1))  # an atom at the middle of the box
    near = np.argsort(np.sum((pos - pos[centre]) ** 2, axis=1))[1:4]
    ptype = np.zeros(len(pos), dtype=int)
    ptype[centre] = 1
    ptype[near[0]] = types[1]
    if types[0] != types[1]:
        ptype[near[1]] = 1
    one = {"sigma": 1.0, "epsilon": 1.0, "rcut": 2.5}

    def build():
        lj = LennardJonesCut(dim=3, shift=True, method="nlist")
        lj.set_parameters({0: dict(one), 1: dict(one), 2: dict(one)})
        bond = DoubleWellPair(types=types, dim=3)
        bond.set_parameters({"rzero": 2.0 ** (1.0 / 6.0), "height": 6.0, "width": 0.25})
        system = System(Box(low=size[:, 0], high=size[:, 1]), Particles.from_arrays(pos, vel=vel, ptype=ptype), [lj, bond])
        system.potential_and_force()
        return system

    ref = build()
    integ = VelocityVerlet(0.004)
    emu = SlabRunEmulation([build()] * world, VelocityVerlet(0.004))
    try:
        for k in (23, 14):
            integ.run(ref, k)
            emu.run(k)
        x, v, f, v_pot, virial = emu.gather_state()
        status = emu.status()
    finally:
        emu.release()
    q = ref.particles
    assert rel_norm(x, q.pos) < 1e-11 and rel_norm(v, q.vel) < 1e-9 and rel_norm(f, q.force) < 1e-9
    assert abs(v_pot - q.v_pot) <= 1e-10 * abs(q.v_pot)
    assert rel_norm(virial, q.virial) < 1e-9
    assert all(s["builds"] >= 1 for s in status)


def test_slab_run_status_and_error_text():
    from turtlemd_b200.integrators import VelocityVerlet
    from turtlemd_b200.parallel import SlabRunEmulation, _sr_error_text

    emu = SlabRunEmulation([_make(12, SINGLE, temperature=0.5)()] * 2, VelocityVerlet(0.004))
    try:
        emu.run(3)
        st = emu.status()
        assert [s["err"] for s in st] == [0, 0] and all(s["error"] == "ok" for s in st)
        assert sum(s["n_own"] for s in st) == emu.n
    finally:
        emu.release()
    assert "migrant inbox full" in _sr_error_text(4) and "capacity" in _sr_error_text(16)


def test_slab_run_refuses_too_many_ranks():
    """Two layers per rank are the minimum (own and ghost layers must not coincide)."""
    from turtlemd_b200.integrators import VelocityVerlet
    from turtlemd_b200.parallel import SlabRunEmulation

    system = _make(12, SINGLE, temperature=1.0)()
    with pytest.raises(NotImplementedError):
        SlabRunEmulation([system] * 4, VelocityVerlet(0.004))
