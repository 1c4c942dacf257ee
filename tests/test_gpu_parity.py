"""GPU parity: the CUDA path (through the C ABI) against golden fixtures and the CPU oracle.

Tolerances (BASELINE.json north_star): forces / potential energy / virial to 1e-10 relative
(norm-wise, max|a-b|/max|b|, see SURVEY.md section 4 on why not element-wise), 10-step
trajectories to 1e-8, RNG draws bit-exact.
"""
from __future__ import annotations

import glob

import numpy as np
import pytest

from conftest import GOLDEN, load_golden, rel_norm

pytestmark = pytest.mark.gpu

TOL = 1e-10
TRAJ_TOL = 1e-8

LJ_CASES = sorted(p.split("/")[-1] for p in glob.glob(str(GOLDEN / "lj_*.npz")))
SINGLE = {0: {"sigma": 1.0, "epsilon": 1.0, "rcut": 2.5}}


def _pkg():
    import turtlemd_b200.integrators as integrators
    from turtlemd_b200.potentials import DoubleWell, DoubleWellPair, LennardJonesCut
    from turtlemd_b200.simulation import MDSimulation
    from turtlemd_b200.system import Box, Particles, System

    return dict(integrators=integrators, DoubleWell=DoubleWell, DoubleWellPair=DoubleWellPair,
                LennardJonesCut=LennardJonesCut, MDSimulation=MDSimulation, Box=Box,
                Particles=Particles, System=System)


def _lj_params_for(name):
    two = {0: {"sigma": 1.0, "epsilon": 1.0, "rcut": 2.5}, 1: {"sigma": 1.2, "epsilon": 0.7, "rcut": 3.0}}
    three = dict(two)
    three[2] = {"sigma": 0.9, "epsilon": 1.3, "rcut": 2.2}
    three[(0, 2)] = {"sigma": 1.1, "epsilon": 0.5, "rcut": 2.0}
    if "two_types_geometric" in name:
        return two, dict(mixing="geometric", shift=True)
    if "two_types_arithmetic" in name:
        return two, dict(mixing="arithmetic", shift=False)
    if "three_types" in name:
        return three, dict(mixing="sixthpower", shift=True)
    if "wca" in name:
        return {0: {"sigma": 1.0, "epsilon": 1.0, "rcut": 2.0 ** (1.0 / 6.0)}}, dict(shift=True)
    return SINGLE, dict(shift=True)


def _lj_system(g, name, method="allpairs"):
    k = _pkg()
    params, kw = _lj_params_for(name)
    box = k["Box"](low=g["low"], high=g["high"], periodic=[bool(p) for p in g["periodic"]])
    part = k["Particles"].from_arrays(g["pos"], ptype=g["ptype"])
    pot = k["LennardJonesCut"](dim=g["pos"].shape[1], method=method, **kw)
    pot.set_parameters(params)
    return k["System"](box=box, particles=part, potentials=[pot]), pot


@pytest.mark.parametrize("name", LJ_CASES)
def test_lj_plugin_boundary_matches_reference(name):
    g = load_golden(name)
    system, pot = _lj_system(g, name)
    # the dense table the kernel consumes == the reference's dictionaries
    _, table, _ = pot._tables(system.particles)
    assert np.array_equal(np.nan_to_num(table, nan=-1.0), np.nan_to_num(g["table"], nan=-1.0))
    v, f, w = pot.potential_and_force(system)
    assert abs(v - g["v_pf"]) <= TOL * abs(g["v_pf"])
    assert rel_norm(f, g["f_pf"]) < TOL
    assert rel_norm(w, g["w_pf"]) < TOL
    f2, w2 = pot.force(system)
    assert rel_norm(f2, g["f_f"]) < TOL and rel_norm(w2, g["w_f"]) < TOL
    assert np.array_equal(f2, f)  # force() and potential_and_force() share the arithmetic
    assert abs(pot.potential(system) - g["v_p"]) <= TOL * abs(g["v_p"])


@pytest.mark.parametrize("name", ["lj_fcc_864.npz", "lj_two_types_geometric.npz", "lj_fcc_256_unwrapped.npz"])
def test_lj_resident_path_equals_plugin_path(name):
    g = load_golden(name)
    system, pot = _lj_system(g, name)
    v, f, w = pot.potential_and_force(system)
    v2, f2, w2 = system.potential_and_force()
    assert v2 == v and np.array_equal(f2, f) and np.array_equal(w2, w)
    assert isinstance(v2, float)
    f3, w3 = system.force()
    assert np.array_equal(f3, f) and np.array_equal(w3, w)
    assert system.particles.v_pot == v  # force() leaves v_pot alone
    vp = system.potential()
    assert abs(vp - g["v_p"]) <= TOL * abs(g["v_p"])
    assert np.array_equal(system.particles.virial, w)


def test_lj_deterministic_and_newton():
    g = load_golden("lj_fcc_4000.npz")
    system, pot = _lj_system(g, "lj_fcc_4000.npz")
    v, f, w = pot.potential_and_force(system)
    v2, f2, w2 = pot.potential_and_force(system)
    assert v == v2 and np.array_equal(f, f2) and np.array_equal(w, w2)  # no atomics: bit-reproducible
    assert np.max(np.abs(f.sum(axis=0))) < 1e-9 * np.max(np.abs(f))
    assert rel_norm(w, w.T) < 1e-12


def test_lj_reference_known_answers():
    """tests/potentials/test_lennardjones.py:16-43,186-220 of the reference."""
    k = _pkg()
    kat = load_golden("reference_kats.npz")
    part = k["Particles"](dim=3)
    for xyz in ([1.0, 1.0, 1.0], [1.5, 1.0, 1.0], [1.5, 1.5, 1.0]):
        part.add_particle(name="Ar", pos=np.array(xyz), mass=1.0, ptype=0)
    system = k["System"](k["Box"](high=np.array([10, 10, 10])), part)
    pot = k["LennardJonesCut"](dim=3, shift=True, mixing="geometric")
    pot.set_parameters(SINGLE)
    assert pot.potential(system) == pytest.approx(float(kat["lj_vpot"]))
    f, w = pot.force(system)
    assert f == pytest.approx(kat["lj_force"]) and w == pytest.approx(kat["lj_virial"])
    v, f, w = pot.potential_and_force(system)
    assert v == pytest.approx(float(kat["lj_vpot"])) and f == pytest.approx(kat["lj_force"])
    pot0 = k["LennardJonesCut"](dim=3, shift=True)
    pot0.set_parameters({0: {"sigma": 1.0, "epsilon": 1.0, "rcut": 0.0}})
    assert pot0.potential(system) == 0.0
    # a type pair without parameters raises KeyError like the reference's dict lookup
    part.add_particle(name="X", pos=np.array([3.0, 3.0, 3.0]), mass=1.0, ptype=5)
    with pytest.raises(KeyError):
        pot.potential(system)


@pytest.mark.parametrize("name", ["doublewell_1d_11.npz", "doublewell_3d_257.npz"])
def test_doublewell(name):
    k = _pkg()
    g = load_golden(name)
    a, b, c = (float(x) for x in g["abc"])
    dim = g["pos"].shape[1]
    system = k["System"](k["Box"](periodic=[False] * dim), k["Particles"].from_arrays(g["pos"]),
                         [k["DoubleWell"](a=a, b=b, c=c)])
    pot = system.potentials[0]
    v = pot.potential(system)
    f, w = pot.force(system)
    assert abs(v - g["v"]) <= TOL * abs(g["v"])
    assert rel_norm(f, g["f"]) < TOL and np.array_equal(w, g["w"])
    v2, f2, w2 = system.potential_and_force()
    assert v2 == v and np.array_equal(f2, f) and np.array_equal(w2, np.zeros((dim, dim)))


@pytest.mark.parametrize("types", ["11", "12", "02"])
def test_doublewell_pair(types):
    k = _pkg()
    g = load_golden(f"dwpair_108_types{types}.npz")
    system = k["System"](k["Box"](low=g["low"], high=g["high"]),
                         k["Particles"].from_arrays(g["pos"], ptype=g["ptype"]))
    pot = k["DoubleWellPair"](types=tuple(int(t) for t in g["types"]))
    pot.set_parameters({"rzero": float(g["rzero"]), "height": float(g["height"]), "width": float(g["width"])})
    v = pot.potential(system)
    f, w = pot.force(system)
    assert abs(v - g["v"]) <= TOL * abs(g["v"])
    assert rel_norm(f, g["f"]) < TOL and rel_norm(w, g["w"]) < TOL
    v2, f2, w2 = pot.potential_and_force(system)
    assert v2 == v and np.array_equal(f2, f)
    system.potentials = [pot]
    v3, f3, w3 = system.potential_and_force()
    assert v3 == v and np.array_equal(f3, f) and np.array_equal(w3, w)


def test_doublewell_pair_reference_known_answers():
    """tests/potentials/test_wellwca.py:12-45,87-110 of the reference."""
    k = _pkg()
    kat = load_golden("reference_kats.npz")

    def make(pos):
        part = k["Particles"](dim=3)
        for xyz, name, ptype in zip(pos, ("Ar1", "Ar2", "X", "X"), (0, 0, 1, 1)):
            part.add_particle(pos=xyz, mass=1.0, name=name, ptype=ptype)
        return k["System"](k["Box"](high=[10, 10, 10]), part)

    pot = k["DoubleWellPair"](types=(0, 0))
    pot.set_parameters({"rzero": 1.0, "width": 0.25, "height": 6.0})
    assert pot.min_max() == (1.0, 1.5, 1.25)
    sys1 = make([[1.0, 1.0, 1.0], [2.25, 1.0, 1.0], [1.0, 2.0, 1.0], [1.0, 3.25, 1.0]])
    assert pot.potential(sys1) == pytest.approx(6)
    sys2 = make([[1.0, 1.0, 1.0], [1.2, 1.2, 1.2], [1.0, 2.0, 1.0], [1.0, 3.25, 1.0]])
    f, w = pot.force(sys2)
    assert f[0] == pytest.approx(kat["dwp_force"]) and f[1] == pytest.approx(-kat["dwp_force"])
    assert np.all(f[2:] == 0.0) and w == pytest.approx(kat["dwp_virial"])
    v, f2, w2 = pot.potential_and_force(sys2)
    assert v == pytest.approx(float(kat["dwp_vpot"])) and f2 == pytest.approx(f)


# ---- trajectories --------------------------------------------------------------------------------


def _traj_system(g, potentials):
    k = _pkg()
    if "low" in g:
        box = k["Box"](low=g["low"], high=g["high"])
    else:
        box = k["Box"](periodic=[False] * g["pos0"].shape[1])
    part = k["Particles"].from_arrays(g["pos0"], vel=g["vel0"], mass=g["mass"], ptype=g["ptype"])
    return k["System"](box=box, particles=part, potentials=potentials)


def _lj(rc=2.5, types=(0,)):
    k = _pkg()
    pot = k["LennardJonesCut"](dim=3, shift=True)
    pot.set_parameters({t: {"sigma": 1.0, "epsilon": 1.0, "rcut": rc} for t in types})
    return pot


def _assert_frame(system, g, frame, tol=TRAJ_TOL):
    p = system.particles
    assert rel_norm(p.pos, g["pos"][frame]) < tol, f"pos, frame {frame}"
    assert rel_norm(p.vel, g["vel"][frame]) < tol, f"vel, frame {frame}"
    assert rel_norm(p.force, g["force"][frame]) < tol, f"force, frame {frame}"
    if not np.isnan(g["v_pot"][frame]):
        assert abs(p.v_pot - g["v_pot"][frame]) <= tol * max(1.0, abs(g["v_pot"][frame]))
    assert rel_norm(p.virial, g["virial"][frame]) < tol or np.max(np.abs(g["virial"][frame])) == 0.0


@pytest.mark.parametrize("name", ["traj_vv_lj_108.npz", "traj_vv_lj_256_masses.npz"])
def test_velocity_verlet_through_mdsimulation(name):
    k = _pkg()
    g = load_golden(name)
    system = _traj_system(g, [_lj()])
    sim = k["MDSimulation"](system, k["integrators"].VelocityVerlet(float(g["dt"])), 10)
    frames = 0
    for frame, sysi in enumerate(sim.run()):
        _assert_frame(sysi, g, frame)
        frames += 1
    assert frames == 11


def test_velocity_verlet_multi_step_run_is_bit_identical():
    k = _pkg()
    g = load_golden("traj_vv_lj_256_masses.npz")
    a = _traj_system(g, [_lj()])
    b = _traj_system(g, [_lj()])
    a.potential_and_force()
    b.potential_and_force()
    integ = k["integrators"].VelocityVerlet(float(g["dt"]))
    for _ in range(10):
        integ(a)
    integ.run(b, 10)
    assert np.array_equal(a.particles.pos, b.particles.pos)
    assert np.array_equal(a.particles.vel, b.particles.vel)
    _assert_frame(b, g, 10)


def test_verlet_trajectory():
    k = _pkg()
    g = load_golden("traj_verlet_lj_108.npz")
    system = _traj_system(g, [_lj()])
    integ = k["integrators"].Verlet(float(g["dt"]))
    assert integ.previous_pos is None
    sim = k["MDSimulation"](system, integ, 10)
    for frame, sysi in enumerate(sim.run()):
        _assert_frame(sysi, g, frame)
        if frame > 0:
            assert rel_norm(integ.previous_pos, g["pos"][frame - 1]) < TRAJ_TOL


def test_dimer_in_solvent_trajectory():
    k = _pkg()
    g = load_golden("traj_vv_dimer_108.npz")
    rc = float(g["rc"])
    dw = k["DoubleWellPair"](types=(1, 1))
    dw.set_parameters({"rzero": rc, "height": 6.0, "width": 0.25})
    system = _traj_system(g, [_lj(rc=rc, types=(0, 1)), dw])
    sim = k["MDSimulation"](system, k["integrators"].VelocityVerlet(float(g["dt"])), 10)
    for frame, sysi in enumerate(sim.run()):
        _assert_frame(sysi, g, frame)


@pytest.mark.parametrize("dim", [1, 2])
@pytest.mark.parametrize("kind", ["inertia", "overdamped"])
@pytest.mark.parametrize("draws", ["in_kernel", "host_twin"])
def test_langevin_doublewell_replicas(kind, dim, draws):
    """The reference, fed the Philox twin, produced the fixture; the kernels must retrace it."""
    k = _pkg()
    from oracle.philox_ref import OraclePhiloxNormal
    from turtlemd_b200.philox import PhiloxNormal

    g = load_golden(f"traj_langevin_{kind}_{dim}d_64.npz")
    a, b, c = (float(x) for x in g["abc"])
    system = _traj_system(g, [k["DoubleWell"](a=a, b=b, c=c)])
    system.potential_and_force()
    # in_kernel: PhiloxNormal selects the device stream.  host_twin: an object the package does
    # not know (the oracle's twin) -> draws happen on the host exactly like in the reference.
    rgen = PhiloxNormal(key=int(g["key"])) if draws == "in_kernel" else OraclePhiloxNormal(key=int(g["key"]))
    dt, gamma, beta = float(g["dt"]), float(g["gamma"]), float(g["beta"])
    I = k["integrators"]
    integ = (I.LangevinInertia(timestep=dt, gamma=gamma, beta=beta, rgen=rgen) if kind == "inertia"
             else I.LangevinOverdamped(timestep=dt, gamma=gamma, rgen=rgen, beta=beta))
    for frame in range(11):
        _assert_frame(system, g, frame)
        integ(system)
    assert rgen.position == 11 * 64 * dim * (2 if kind == "inertia" else 1)  # 11 steps were taken


@pytest.mark.parametrize("tag", ["3d_150", "3d_90_longcut", "2d_80"])
def test_lj_in_a_triclinic_box(tag):
    """LennardJonesCut with box = TriclinicBox (Mezei reduction + 3^dim image scan, box.py:271-296) against the
    reference: host-buffer entry points and the device-resident path, plus the host mirror of pbc_dist."""
    from oracle import turtle_oracle as to
    from turtlemd_b200.system import Particles, System, TriclinicBox

    k = _pkg()
    g = load_golden(f"triclinic_lj_{tag}.npz")
    alpha, beta, gamma = (None if np.isnan(a) else float(a) for a in g["angles"])
    box = TriclinicBox(high=list(g["high"]), alpha=alpha, beta=beta, gamma=gamma)
    assert np.array_equal(box.box_matrix, g["box_matrix"])
    cell = to.TriclinicCell(g["high"], alpha, beta, gamma)
    rng = np.random.default_rng(5)
    for d in rng.normal(scale=6.0, size=(50, len(g["high"]))):
        assert np.array_equal(box.pbc_dist(d), to.tri_pbc_dist(d, cell))
    params = {t: dict(v, rcut=v["rcut"] * float(g["rscale"]))
              for t, v in {0: {"sigma": 1.0, "epsilon": 1.0, "rcut": 2.5}, 1: {"sigma": 1.1, "epsilon": 0.8, "rcut": 2.2}}.items()}
    pot = k["LennardJonesCut"](dim=len(g["high"]), shift=True, mixing="geometric")
    pot.set_parameters(params)
    system = System(box, Particles.from_arrays(g["pos"], ptype=g["ptype"]), [pot])
    v, f, w = pot.potential_and_force(system)        # host buffers in, host buffers out
    assert abs(v - float(g["v"])) <= TOL * abs(float(g["v"]))
    assert rel_norm(f, g["f"]) < TOL and rel_norm(w, g["w"]) < TOL
    assert abs(pot.potential(system) - float(g["v_only"])) <= TOL * abs(float(g["v_only"]))
    f2, w2 = pot.force(system)
    assert rel_norm(f2, g["f"]) < TOL and rel_norm(w2, g["w_only"]) < TOL
    system.potential_and_force()                      # device-resident path (System.evaluate)
    assert rel_norm(system.particles.force, g["f"]) < TOL
    assert rel_norm(system.particles.virial, g["w"]) < TOL
    assert abs(system.particles.v_pot - float(g["v"])) <= TOL * abs(float(g["v"]))
    # size-independent properties: Newton's third law, symmetric virial; and the integrator runs on it
    fdev = system.particles.force
    assert np.all(np.abs(fdev.sum(axis=0)) < 1e-9 * np.abs(fdev).sum())
    assert rel_norm(system.particles.virial, system.particles.virial.T) < 1e-12
    integ = k["integrators"].VelocityVerlet(1e-4)
    for _ in range(3):
        integ(system)
    assert np.all(np.isfinite(system.particles.pos)) and np.isfinite(system.particles.v_pot)


@pytest.mark.parametrize("name", ["3d_90_longcut_types11", "3d_90_longcut_types01", "2d_80_types11", "2d_80_types01"])
def test_doublewell_pair_in_a_triclinic_box(name):
    """DoubleWellPair with box = TriclinicBox against the reference, host-buffer entry point and device path,
    alone and summed with LennardJonesCut by System (system.py:58-99)."""
    from turtlemd_b200.system import Particles, System, TriclinicBox

    k = _pkg()
    g = load_golden(f"triclinic_dwpair_{name}.npz")
    alpha, beta, gamma = (None if np.isnan(a) else float(a) for a in g["angles"])
    box = TriclinicBox(high=list(g["high"]), alpha=alpha, beta=beta, gamma=gamma)
    dw = k["DoubleWellPair"](types=tuple(int(t) for t in g["types"]), dim=len(g["high"]))
    dw.set_parameters({"rzero": float(g["rzero"]), "width": float(g["width"]), "height": float(g["height"])})
    system = System(box, Particles.from_arrays(g["pos"], ptype=g["ptype"]), [dw])
    assert abs(dw.potential(system) - float(g["v"])) <= TOL * abs(float(g["v"]))
    f, w = dw.force(system)
    assert rel_norm(f, g["f"]) < TOL and rel_norm(w, g["w"]) < TOL
    system.potential_and_force()
    assert rel_norm(system.particles.force, g["f"]) < TOL and rel_norm(system.particles.virial, g["w"]) < TOL
    # with the Lennard-Jones fluid of the same fixture family on top
    lj_tag = name.rsplit("_types", 1)[0]
    gl = load_golden(f"triclinic_lj_{lj_tag}.npz")
    params = {t: dict(v, rcut=v["rcut"] * float(gl["rscale"]))
              for t, v in {0: {"sigma": 1.0, "epsilon": 1.0, "rcut": 2.5}, 1: {"sigma": 1.1, "epsilon": 0.8, "rcut": 2.2}}.items()}
    lj = k["LennardJonesCut"](dim=len(g["high"]), shift=True, mixing="geometric")
    lj.set_parameters(params)
    both = System(box, Particles.from_arrays(g["pos"], ptype=g["ptype"]), [lj, dw])
    both.potential_and_force()
    assert rel_norm(both.particles.force, gl["f"] + g["f"]) < TOL
    assert rel_norm(both.particles.virial, gl["w"] + g["w"]) < TOL
    assert abs(both.particles.v_pot - float(gl["v"]) - float(g["v"])) <= TOL * (abs(float(gl["v"])) + abs(float(g["v"])))


@pytest.mark.parametrize("tag", ["3d_500", "1d_64", "2d_301"])
def test_maxwell_velocities_on_the_device(tag):
    """Velocity set-up with the in-kernel stream (draw, momentum removal, rescaling on the device) against
    the reference fed the Philox twin (system/particles.py:140-153, :227-269)."""
    from turtlemd_b200.philox import PhiloxNormal
    from turtlemd_b200.system import Particles
    from turtlemd_b200.system.particles import (generate_maxwell_velocities, kinetic_temperature,
                                                linear_momentum, zero_momentum)

    g = load_golden(f"maxwell_{tag}.npz")
    n, dim = g["vel"].shape
    dof = None if g["dof"].size == 0 else list(g["dof"])
    part = Particles.from_arrays(np.zeros((n, dim)), mass=g["mass"])
    rgen = PhiloxNormal(key=int(g["key"]))
    generate_maxwell_velocities(part, rgen, temperature=float(g["temperature"]), boltzmann=float(g["boltzmann"]),
                                dof=dof, momentum=bool(g["momentum"]))
    assert rgen.position == int(g["draws"])
    assert part._vel.dev_valid and not part._vel.host_valid  # nothing was drawn or reduced on the host
    temp, _, _ = kinetic_temperature(part, float(g["boltzmann"]), dof=dof)  # device reduction
    assert temp == pytest.approx(float(g["temperature"]), rel=1e-12)
    if bool(g["momentum"]):
        scale = np.abs(g["vel"] * g["mass"][:, None]).sum()
        assert np.all(np.abs(linear_momentum(part)) < 1e-13 * scale)
    assert rel_norm(part.vel, g["vel"]) < 1e-13
    # zero_momentum with an axis mask on velocities that live on the device only
    part.vel = g["vel"] + g["drift"]
    part._vel.device()
    part._vel.written_on_device()
    zero_momentum(part, dim=[bool(m) for m in g["mask"]])
    assert part._vel.dev_valid and not part._vel.host_valid
    assert rel_norm(part.vel, g["vel_masked"]) < 1e-13


def test_langevin_fused_multi_step_equals_single_steps():
    k = _pkg()
    from turtlemd_b200.philox import PhiloxNormal

    g = load_golden("traj_langevin_inertia_1d_64.npz")
    a, b, c = (float(x) for x in g["abc"])
    dt, gamma, beta = float(g["dt"]), float(g["gamma"]), float(g["beta"])
    system = _traj_system(g, [k["DoubleWell"](a=a, b=b, c=c)])
    system.potential_and_force()
    integ = k["integrators"].LangevinInertia(timestep=dt, gamma=gamma, beta=beta, rgen=PhiloxNormal(int(g["key"])))
    integ.run(system, 10)
    _assert_frame(system, g, 10)
    # sharded: two halves with global offsets reproduce the same replicas bit for bit
    full_pos = system.particles.pos.copy()
    # (0, 40) / (40, 24): two coordinates per thread sharing one Philox block; (0, 41): the same kernel with a
    # lone last coordinate; (41, 23): stream position not a multiple of four -> one coordinate per thread
    for first, count in ((0, 40), (40, 24), (0, 41), (41, 23)):
        sl = slice(first, first + count)
        sub = dict(pos0=g["pos0"][sl], vel0=g["vel0"][sl], mass=g["mass"][sl], ptype=g["ptype"][sl])
        shard = _traj_system(sub, [k["DoubleWell"](a=a, b=b, c=c)])
        shard.potential_and_force()
        integ = k["integrators"].LangevinInertia(timestep=dt, gamma=gamma, beta=beta,
                                                 rgen=PhiloxNormal(int(g["key"])))
        integ.run(shard, 10, first=first, n_global=64)
        assert np.array_equal(shard.particles.pos, full_pos[sl])


@pytest.mark.parametrize("method", ["allpairs", "nlist"])
def test_langevin_multi_step_run_on_a_pair_force_field_is_bit_identical(method):
    """LangevinInertia.run(system, k) on a Lennard-Jones system (one fused pass between force evaluations,
    potential() once at the end) leaves exactly the state of k integration_step calls."""
    from turtlemd_b200.philox import PhiloxNormal
    from turtlemd_b200.system import Box, Particles, System

    k = _pkg()
    g = load_golden("lj_fcc_4000.npz" if method == "nlist" else "lj_fcc_108.npz")
    vel0 = 0.3 * np.random.default_rng(8).standard_normal(g["pos"].shape)

    def make():
        pot = k["LennardJonesCut"](dim=3, shift=True, method=method)
        pot.set_parameters({0: {"sigma": 1.0, "epsilon": 1.0, "rcut": 2.5}})
        system = System(Box(low=g["low"], high=g["high"]), Particles.from_arrays(g["pos"], vel=vel0), [pot])
        system.potential_and_force()
        integ = k["integrators"].LangevinInertia(timestep=0.002, gamma=0.3, beta=1.25, rgen=PhiloxNormal(key=21))
        return system, integ

    a, integ_a = make()
    b, integ_b = make()
    integ_a.run(a, 6)
    for _ in range(6):
        integ_b(b)
    assert integ_a.rgen.position == integ_b.rgen.position
    for name in ("pos", "vel", "force"):
        assert np.array_equal(getattr(a.particles, name), getattr(b.particles, name)), name
    assert a.particles.v_pot == b.particles.v_pot
    assert np.array_equal(a.particles.virial, b.particles.virial)


def test_overdamped_multi_step_run_is_bit_identical():
    """LangevinOverdamped.run(system, k): the intermediate potential() evaluations nobody can observe are
    skipped; the state after k steps is that of k integration_step calls."""
    from turtlemd_b200.philox import PhiloxNormal
    from turtlemd_b200.system import Box, Particles, System

    k = _pkg()
    g = load_golden("lj_fcc_4000.npz")

    def make():
        pot = k["LennardJonesCut"](dim=3, shift=True, method="nlist")
        pot.set_parameters({0: {"sigma": 1.0, "epsilon": 1.0, "rcut": 2.5}})
        system = System(Box(low=g["low"], high=g["high"]), Particles.from_arrays(g["pos"]), [pot])
        system.potential_and_force()
        return system, k["integrators"].LangevinOverdamped(timestep=1e-4, gamma=0.5, rgen=PhiloxNormal(key=9), beta=1.25)

    a, integ_a = make()
    b, integ_b = make()
    integ_a.run(a, 5)
    for _ in range(5):
        integ_b(b)
    for name in ("pos", "vel", "force"):
        assert np.array_equal(getattr(a.particles, name), getattr(b.particles, name)), name
    assert a.particles.v_pot == b.particles.v_pot and np.array_equal(a.particles.virial, b.particles.virial)
    assert a.potentials[0].nlist_stats()["calls"] == 1 + 5 + 1 and b.potentials[0].nlist_stats()["calls"] == 1 + 10


def test_langevin_inertia_lj_trajectory():
    k = _pkg()
    from turtlemd_b200.philox import PhiloxNormal

    g = load_golden("traj_langevin_inertia_lj_108.npz")
    system = _traj_system(g, [_lj()])
    system.potential_and_force()
    integ = k["integrators"].LangevinInertia(timestep=float(g["dt"]), gamma=float(g["gamma"]),
                                             beta=float(g["beta"]), rgen=PhiloxNormal(int(g["key"])))
    for frame in range(11):
        _assert_frame(system, g, frame)
        integ(system)


def test_langevin_numpy_generator_host_draw_path():
    """A default_rng generator as ``rgen`` (continued on the device since round 2, drawn on the host before; the host-draw
    path itself is exercised by tests/test_npy_normal.py); compare with the oracle fed the same generator."""
    k = _pkg()
    from oracle import turtle_oracle as to

    g = load_golden("traj_langevin_inertia_1d_64.npz")
    a, b, c = (float(x) for x in g["abc"])
    dt, gamma, beta = float(g["dt"]), float(g["gamma"]), float(g["beta"])
    system = _traj_system(g, [k["DoubleWell"](a=a, b=b, c=c)])
    system.potential_and_force()
    integ = k["integrators"].LangevinInertia(timestep=dt, gamma=gamma, beta=beta, rgen=np.random.default_rng(123))
    ff = to.ForceField(np.array([np.inf]), np.zeros(1), [False], [("dw", (a, b, c))])
    st = to.State(g["pos0"], g["vel0"], g["mass"])
    to._eval_all(ff, st)
    stepper = to.InertiaStepper(dt, gamma, beta, np.random.default_rng(123))
    for _ in range(5):
        integ(system)
        stepper.step(st, ff)
    assert rel_norm(system.particles.pos, st.pos) < TRAJ_TOL
    assert rel_norm(system.particles.vel, st.vel) < TRAJ_TOL


def test_reference_integrator_series():
    """tests/test_integrators.py:410-562 of the reference (one particle, DoubleWell(1,2,0), dt=0.002)."""
    k = _pkg()
    I = k["integrators"]
    kat = load_golden("reference_kats.npz")

    class FakeRandom:
        def __init__(self, table, seed):
            self.table, self.seed = list(table), seed

        def normal(self, loc=0.0, scale=1.0, size=None):
            out = np.zeros(size)
            flat = out.reshape(-1)
            for n in range(flat.size):
                if self.seed >= len(self.table):
                    self.seed = 0
                flat[n] = self.table[self.seed]
                self.seed += 1
            return out

    def make():
        system = k["System"](box=k["Box"](periodic=[False]), particles=k["Particles"](dim=1),
                             potentials=[k["DoubleWell"](a=1.0, b=2.0, c=0.0)])
        system.particles.add_particle(pos=np.array([-1.0]), vel=np.array([0.78008020]), name="Ar", mass=1.0, ptype=0)
        return system

    system, integ = make(), I.VelocityVerlet(timestep=0.002)
    for i in range(21):
        assert system.particles.pos[0][0] == pytest.approx(kat["vv_pos"][i])
        assert system.particles.vel[0][0] == pytest.approx(kat["vv_vel"][i])
        integ(system)
    system, integ = make(), I.Verlet(timestep=0.002)
    prev = np.copy(system.particles.pos)
    for i in range(5):
        integ.integration_step(system)
        assert system.particles.pos[0][0] == pytest.approx(kat["verlet_pos"][i])
        assert system.particles.vel[0][0] == pytest.approx(kat["verlet_vel"][i])
        assert prev == pytest.approx(integ.previous_pos)
        prev = np.copy(system.particles.pos)
    system = make()
    integ = I.LangevinOverdamped(timestep=0.002, gamma=0.3, rgen=FakeRandom(kat["fake_table"], 1), beta=1.0)
    for i in range(51):
        assert system.particles.pos[0][0] == pytest.approx(kat["over_pos"][i])
        assert system.particles.vel[0][0] == pytest.approx(kat["over_vel"][i])
        integ(system)
    # LangevinInertia with the module-level multivariate_normal monkeypatched (tests/test_integrators.py:552-557)
    system = make()
    integ = I.LangevinInertia(timestep=0.002, gamma=0.3, rgen=FakeRandom(kat["fake_table"], 1), beta=1.0)

    def fake_mvn(rgen, mean, cho, dim):
        norm = rgen.normal(loc=0.0, scale=1.0, size=2 * dim).reshape(dim, 2)
        return 0.01 * (np.array([mean] * dim) + norm)

    original = I.multivariate_normal
    I.multivariate_normal = fake_mvn
    try:
        for i in range(51):
            assert system.particles.pos[0][0] == pytest.approx(kat["inertia_pos"][i])
            assert system.particles.vel[0][0] == pytest.approx(kat["inertia_vel"][i])
            integ(system)
    finally:
        I.multivariate_normal = original


def test_reference_md_trajectory_fixture():
    """tests/simulation/test_mdsimulation.py:98-141 of the reference: md-traj.npy, thermo series."""
    k = _pkg()
    from turtlemd_b200.tools import generate_lattice

    kat = load_golden("reference_kats.npz")
    xyz, size = generate_lattice("fcc", [3, 3, 3], density=0.9)
    box = k["Box"](low=size[:, 0], high=size[:, 1])
    part = k["Particles"](dim=box.dim)
    for pos in xyz:
        part.add_particle(pos=pos, mass=1.0, name="Ar", ptype=0)
    part.vel = kat["md_vel"]
    system = k["System"](box=box, particles=part, potentials=[_lj()])
    sim = k["MDSimulation"](system, k["integrators"].VelocityVerlet(timestep=0.002), 10)
    count = 0
    for i, sysi in enumerate(sim.run()):
        th = sysi.thermo(boltzmann=1)
        assert th["temperature"] == pytest.approx(kat["md_temperature"][i])
        assert th["pressure"] == pytest.approx(kat["md_pressure"][i])
        assert th["ekin"] == pytest.approx(kat["md_ekin"][i])
        assert th["vpot"] == pytest.approx(kat["md_vpot"][i])
        assert sysi.particles.pos == pytest.approx(kat["md_traj"][i][:, :3])
        assert sysi.particles.vel == pytest.approx(kat["md_traj"][i][:, 3:])
        count += 1
    assert count == 11 and sim.cycle["stepno"] == 10


def test_host_edits_are_seen_by_the_next_step():
    """In-place writes into handed-out arrays and rebinding both reach the device (hard part 1)."""
    k = _pkg()
    g = load_golden("traj_vv_lj_108.npz")
    a = _traj_system(g, [_lj()])
    b = _traj_system(g, [_lj()])
    integ = k["integrators"].VelocityVerlet(float(g["dt"]))
    for system in (a, b):
        system.potential_and_force()
        integ(system)
    a.particles.pos[3, 1] += 0.01          # in-place element write
    a.particles.vel = a.particles.vel * 0.5  # rebinding
    pos_b = b.particles.pos.copy()
    pos_b[3, 1] += 0.01
    b.particles.pos = pos_b
    b.particles.vel *= 0.5                  # in-place on the handed-out array
    for system in (a, b):
        system.potential_and_force()
        integ(system)
    assert np.array_equal(a.particles.pos, b.particles.pos)
    assert np.array_equal(a.particles.vel, b.particles.vel)
    assert not np.array_equal(a.particles.pos, g["pos"][2])


# ---- RNG ---------------------------------------------------------------------------------------------


def test_philox_device_stream_is_bit_exact():
    import ctypes as C

    from oracle import philox_ref as pr
    from turtlemd_b200 import _device, _lib
    from turtlemd_b200.philox import PhiloxNormal

    _device.require_cuda()
    for key, offset, count in ((0, 0, 4096), (2024, 7, 1001), (2**63 + 11, 10**12 + 1, 513)):
        out = _device.empty(count)
        rng = _lib.Rng(key, offset)
        _lib.call("tmd_philox_normal_fill", _device.dptr(out), count, C.byref(rng), _device.stream_ptr())
        dev = out.cpu().numpy()
        assert np.array_equal(dev, PhiloxNormal(key, offset).peek(count))      # device == host twin (same source)
        assert np.array_equal(dev, pr.normals(key, offset, count))             # == independent NumPy oracle
    raw = PhiloxNormal(99).raw(64)
    assert np.array_equal(raw, np.random.Philox(key=99).random_raw(64))        # == NumPy's bit generator


# ---- sizes of the benchmark configs --------------------------------------------------------------------


def _fluid(rep, seed=3):
    from turtlemd_b200.tools import generate_lattice

    pos, size = generate_lattice("fcc", [rep] * 3, density=0.8)
    pos = pos + 0.05 * (2.0 * np.random.default_rng(seed).random(pos.shape) - 1.0)
    return pos, size


def test_lj_16k_against_oracle():
    from oracle import turtle_oracle as to

    k = _pkg()
    pos, size = _fluid(16)
    system = k["System"](k["Box"](low=size[:, 0], high=size[:, 1]), k["Particles"].from_arrays(pos), [_lj()])
    v, f, w = system.potentials[0].potential_and_force(system)
    length = size[:, 1] - size[:, 0]
    table, _ = to.lj_tables(SINGLE)
    vo, fo, wo = to.lj_potential_and_force(pos, np.zeros(len(pos), dtype=int), length, 1.0 / length, [True] * 3, table)
    assert abs(v - vo) <= TOL * abs(vo) and rel_norm(f, fo) < TOL and rel_norm(w, wo) < TOL


def test_lj_32k_benchmark_size_properties():
    """Config #2 (N=32 000 all-pairs): sampled rows vs the oracle + size-independent properties."""
    from oracle import turtle_oracle as to

    k = _pkg()
    pos, size = _fluid(20)
    assert len(pos) == 32000
    length = size[:, 1] - size[:, 0]
    box = k["Box"](low=size[:, 0], high=size[:, 1])
    system = k["System"](box, k["Particles"].from_arrays(pos), [_lj()])
    v, f, w = system.potentials[0].potential_and_force(system)
    rows = np.random.default_rng(5).choice(len(pos), 200, replace=False)
    table, _ = to.lj_tables(SINGLE)
    fr, vr = to.lj_rows(pos, np.zeros(len(pos), dtype=int), length, 1.0 / length, [True] * 3, table, rows)
    assert rel_norm(f[rows], fr) < TOL
    assert np.max(np.abs(f.sum(axis=0))) < 1e-9 * np.max(np.abs(f))      # Newton's third law
    assert rel_norm(w, w.T) < 1e-12                                        # symmetric virial
    # translation by whole box vectors + a rigid shift must not change anything beyond rounding
    shifted = pos + np.array([3.0, -2.0, 1.0]) * length + 0.123
    system2 = k["System"](box, k["Particles"].from_arrays(shifted), [_lj()])
    v2, f2, w2 = system2.potentials[0].potential_and_force(system2)
    assert abs(v2 - v) <= 1e-9 * abs(v) and rel_norm(f2, f) < 1e-9


# ---- cell list (K2/K3) ----------------------------------------------------------------------------------


def test_lj_cells_against_reference_fixture():
    """N=4000 (6 cells per direction): the cell-list kernel against the reference's own numbers."""
    g = load_golden("lj_fcc_4000.npz")
    system, pot = _lj_system(g, "lj_fcc_4000.npz", method="cells")
    v, f, w = pot.potential_and_force(system)
    assert abs(v - g["v_pf"]) <= TOL * abs(g["v_pf"])
    assert rel_norm(f, g["f_pf"]) < TOL and rel_norm(w, g["w_pf"]) < TOL
    v2, f2, w2 = pot.potential_and_force(system)
    assert v2 == v and np.array_equal(f2, f) and np.array_equal(w2, w)  # atomics only rank atoms; sums are ordered
    vr, fr, wr = system.potential_and_force()
    assert vr == v and np.array_equal(fr, f)
    assert abs(system.potential() - g["v_pf"]) <= TOL * abs(g["v_pf"])


def test_lj_cells_small_box_is_refused():
    g = load_golden("lj_fcc_256.npz")
    system, pot = _lj_system(g, "lj_fcc_256.npz", method="cells")
    with pytest.raises(NotImplementedError):
        pot.potential_and_force(system)
    g = load_golden("lj_slab_256.npz")
    system, pot = _lj_system(g, "lj_slab_256.npz", method="auto")  # falls back to all-pairs
    v, f, w = pot.potential_and_force(system)
    assert rel_norm(f, g["f_pf"]) < TOL


@pytest.mark.parametrize("dim", [2, 3])
def test_lj_cells_equals_allpairs_multitype_unwrapped(dim):
    k = _pkg()
    from turtlemd_b200.tools import generate_lattice

    rng = np.random.default_rng(71 + dim)
    if dim == 3:
        pos, size = generate_lattice("fcc", [12, 12, 12], density=0.8)
    else:
        pos, size = generate_lattice("sq2", [64, 64], density=0.7)
    length = size[:, 1] - size[:, 0]
    pos = pos + 0.06 * (2.0 * rng.random(pos.shape) - 1.0) + rng.integers(-2, 3, size=pos.shape) * length
    ptype = rng.integers(0, 3, size=len(pos))
    params = {0: {"sigma": 1.0, "epsilon": 1.0, "rcut": 2.5}, 1: {"sigma": 1.1, "epsilon": 0.8, "rcut": 2.8},
              2: {"sigma": 0.9, "epsilon": 1.2, "rcut": 2.2}}
    out = {}
    for method in ("allpairs", "cells"):
        pot = k["LennardJonesCut"](dim=dim, shift=True, mixing="arithmetic", method=method)
        pot.set_parameters(params)
        system = k["System"](k["Box"](low=size[:, 0], high=size[:, 1]),
                             k["Particles"].from_arrays(pos, ptype=ptype), [pot])
        out[method] = pot.potential_and_force(system)
    (va, fa, wa), (vc, fc, wc) = out["allpairs"], out["cells"]
    assert abs(vc - va) <= TOL * abs(va) and rel_norm(fc, fa) < TOL and rel_norm(wc, wa) < TOL


# ---- Verlet neighbour list (the large-N path) ---------------------------------------------------------------


def test_lj_nlist_against_reference_fixture():
    """N=4000: the neighbour-list path (device-resident System methods) against the reference's numbers."""
    g = load_golden("lj_fcc_4000.npz")
    system, pot = _lj_system(g, "lj_fcc_4000.npz", method="nlist")
    v, f, w = system.potential_and_force()
    assert abs(v - g["v_pf"]) <= TOL * abs(g["v_pf"])
    assert rel_norm(f, g["f_pf"]) < TOL and rel_norm(w, g["w_pf"]) < TOL
    v2, f2, w2 = system.potential_and_force()  # second call reuses the list
    assert v2 == v and np.array_equal(f2, f) and np.array_equal(w2, w)
    f3, w3 = system.force()
    assert np.array_equal(f3, f) and rel_norm(w3, g["w_pf"]) < TOL
    assert abs(system.potential() - g["v_pf"]) <= TOL * abs(g["v_pf"])
    stats = pot.nlist_stats()
    assert stats["builds"] == 1 and stats["calls"] == 4 and stats["slow_rows"] == 0
    # host-array plug-in call of the same potential object (stateless cells kernel) agrees
    vh, fh, wh = pot.potential_and_force(system)
    assert abs(vh - v) <= 1e-13 * abs(v) and rel_norm(fh, f) < 1e-13 and rel_norm(wh, w) < 1e-12


def test_lj_nlist_types_with_identical_parameters_take_the_one_type_kernel():
    """Two particle types whose Lennard-Jones parameters coincide (a solute that differs from the solvent only in
    its bonded terms, examples/movie-doublewell-pair.py): same numbers as the reference's one-type fixture."""
    from turtlemd_b200.system import Box, Particles, System

    k = _pkg()
    g = load_golden("lj_fcc_4000.npz")
    ptype = np.zeros(len(g["pos"]), dtype=int)
    ptype[::7] = 1
    one = {"sigma": 1.0, "epsilon": 1.0, "rcut": 2.5}
    pot = k["LennardJonesCut"](dim=3, shift=True, method="nlist")
    pot.set_parameters({0: dict(one), 1: dict(one)})
    system = System(Box(low=g["low"], high=g["high"]), Particles.from_arrays(g["pos"], ptype=ptype), [pot])
    v, f, w = system.potential_and_force()
    assert abs(v - g["v_pf"]) <= TOL * abs(g["v_pf"])
    assert rel_norm(f, g["f_pf"]) < TOL and rel_norm(w, g["w_pf"]) < TOL
    # parameters that differ go through the table look-ups again
    pot2 = k["LennardJonesCut"](dim=3, shift=True, method="nlist")
    pot2.set_parameters({0: dict(one), 1: dict(one, epsilon=0.5)})
    system2 = System(Box(low=g["low"], high=g["high"]), Particles.from_arrays(g["pos"], ptype=ptype), [pot2])
    ap = k["LennardJonesCut"](dim=3, shift=True, method="allpairs")
    ap.set_parameters({0: dict(one), 1: dict(one, epsilon=0.5)})
    system3 = System(Box(low=g["low"], high=g["high"]), Particles.from_arrays(g["pos"], ptype=ptype), [ap])
    v2, f2, w2 = system2.potential_and_force()
    v3, f3, w3 = system3.potential_and_force()
    assert abs(v2 - v3) <= TOL * abs(v3) and rel_norm(f2, f3) < TOL and rel_norm(w2, w3) < TOL
    assert rel_norm(f2, f) > 1e-3  # and they are a different force field


def test_positions_downloaded_under_the_force_evaluation_are_the_step_positions():
    """A caller that re-binds page-locked host arrays and reads them back after every step (bench.py's e2e
    loop) gets the positions through an asynchronous download that overlaps the force evaluation; they must be
    the very positions a device-resident run produces."""
    from turtlemd_b200 import _device
    from turtlemd_b200.system import Box, Particles, System

    k = _pkg()
    g = load_golden("lj_fcc_4000.npz")
    vel0 = 0.3 * np.random.default_rng(4).standard_normal(g["pos"].shape)

    def make():
        pot = k["LennardJonesCut"](dim=3, shift=True, method="nlist")
        pot.set_parameters({0: {"sigma": 1.0, "epsilon": 1.0, "rcut": 2.5}})
        system = System(Box(low=g["low"], high=g["high"]), Particles.from_arrays(g["pos"], vel=vel0), [pot])
        system.potential_and_force()
        return system

    a, b = make(), make()
    integ_a, integ_b = k["integrators"].VelocityVerlet(0.004), k["integrators"].VelocityVerlet(0.004)
    hp, hv, hf = (_device.pinned_array(g["pos"].shape) for _ in range(3))
    np.copyto(hp, a.particles.pos)
    np.copyto(hv, a.particles.vel)
    np.copyto(hf, a.particles.force)
    prefetched = 0
    for _ in range(6):
        a.particles.pos, a.particles.vel, a.particles.force = hp, hv, hf
        integ_a(a)
        prefetched += a.particles._pos._prefetch is not None
        pos_a, vel_a, force_a = a.particles.pos, a.particles.vel, a.particles.force
        integ_b(b)
        assert np.array_equal(pos_a, b.particles._pos.peek())
        assert np.array_equal(vel_a, b.particles._vel.peek())
        assert np.array_equal(force_a, b.particles._force.peek())
    assert prefetched >= 4  # from the second step on the download rides under the force evaluation


def test_lj_nlist_small_box_is_refused():
    g = load_golden("lj_fcc_256.npz")
    system, pot = _lj_system(g, "lj_fcc_256.npz", method="nlist")
    with pytest.raises(NotImplementedError):
        system.potential_and_force()


@pytest.mark.parametrize("dim", [2, 3])
def test_lj_nlist_equals_allpairs_multitype_unwrapped(dim):
    """Three types with different cut-offs, coordinates several boxes away from the origin cell."""
    k = _pkg()
    from turtlemd_b200.tools import generate_lattice

    rng = np.random.default_rng(171 + dim)
    if dim == 3:
        pos, size = generate_lattice("fcc", [12, 12, 12], density=0.8)
    else:
        pos, size = generate_lattice("sq2", [64, 64], density=0.7)
    length = size[:, 1] - size[:, 0]
    pos = pos + 0.06 * (2.0 * rng.random(pos.shape) - 1.0) + rng.integers(-2, 3, size=pos.shape) * length
    ptype = rng.integers(0, 3, size=len(pos))
    params = {0: {"sigma": 1.0, "epsilon": 1.0, "rcut": 2.5}, 1: {"sigma": 1.1, "epsilon": 0.8, "rcut": 2.8},
              2: {"sigma": 0.9, "epsilon": 1.2, "rcut": 2.2}}
    out = {}
    for method in ("allpairs", "nlist"):
        pot = k["LennardJonesCut"](dim=dim, shift=True, mixing="arithmetic", method=method, skin=0.25)
        pot.set_parameters(params)
        system = k["System"](k["Box"](low=size[:, 0], high=size[:, 1]),
                             k["Particles"].from_arrays(pos, ptype=ptype), [pot])
        out[method] = system.potential_and_force()
    (va, fa, wa), (vc, fc, wc) = out["allpairs"], out["nlist"]
    assert abs(vc - va) <= TOL * abs(va) and rel_norm(fc, fa) < TOL and rel_norm(wc, wa) < TOL
    assert rel_norm(wc, wc.T) == 0.0


def test_lj_nlist_trajectory_rebuilds_and_matches_allpairs():
    """60 VelocityVerlet steps of a hot fluid: the list is rebuilt on the device several times and the
    trajectory stays on the all-pairs one (1e-8 after 60 steps; chaos amplifies the 1e-16 differences)."""
    k = _pkg()
    pos, size = _fluid(10)  # N = 4000
    vel = np.random.default_rng(4).standard_normal(pos.shape) * 1.2
    vel -= vel.mean(axis=0)
    out = {}
    for method in ("allpairs", "nlist"):
        pot = k["LennardJonesCut"](dim=3, shift=True, method=method, skin=0.3)
        pot.set_parameters(SINGLE)
        system = k["System"](k["Box"](low=size[:, 0], high=size[:, 1]),
                             k["Particles"].from_arrays(pos, vel=vel), [pot])
        system.potential_and_force()
        integ = k["integrators"].VelocityVerlet(0.004)
        for _ in range(6):
            integ.run(system, 10)
        out[method] = (system.particles.pos.copy(), system.particles.vel.copy(), system.particles.v_pot,
                       system.particles.virial.copy(), pot.nlist_stats())
    (xa, va, ea, wa, _), (xn, vn, en, wn, stats) = out["allpairs"], out["nlist"]
    assert rel_norm(xn, xa) < TRAJ_TOL and rel_norm(vn, va) < 1e-6
    assert abs(en - ea) <= 1e-8 * abs(ea) and rel_norm(wn, wa) < 1e-7
    assert stats["calls"] == 61 and 2 <= stats["builds"] < 40, stats


def test_lj_nlist_list_overflow_takes_the_exact_slow_path():
    """Half of the box empty, the other half at twice the mean density: most lists outgrow the capacity
    (sized from the mean density) and those rows are evaluated from the cells -- same numbers."""
    k = _pkg()
    from turtlemd_b200.tools import generate_lattice

    pos, size = generate_lattice("fcc", [8, 8, 16], density=1.0)
    pos = pos + 0.04 * (2.0 * np.random.default_rng(8).random(pos.shape) - 1.0)
    size = size.copy()
    size[2, 1] *= 2.6  # the lattice fills the lower 38 % of the box
    out = {}
    for method in ("allpairs", "nlist"):
        pot = k["LennardJonesCut"](dim=3, shift=True, method=method, skin=0.4)
        pot.set_parameters(SINGLE)
        system = k["System"](k["Box"](low=size[:, 0], high=size[:, 1]), k["Particles"].from_arrays(pos), [pot])
        out[method] = system.potential_and_force() + (pot.nlist_stats(),)
    (va, fa, wa, _), (vn, fn, wn, stats) = out["allpairs"], out["nlist"]
    assert stats["slow_rows"] > 100, stats
    assert abs(vn - va) <= TOL * abs(va) and rel_norm(fn, fa) < TOL and rel_norm(wn, wa) < TOL


def test_lj_nlist_follows_host_edits_and_box_changes():
    """Positions rewritten on the host (teleporting atoms) and a rescaled box both invalidate the list."""
    k = _pkg()
    pos, size = _fluid(10)
    pot = k["LennardJonesCut"](dim=3, shift=True, method="nlist")
    pot.set_parameters(SINGLE)
    ref = k["LennardJonesCut"](dim=3, shift=True, method="allpairs")
    ref.set_parameters(SINGLE)
    box = k["Box"](low=size[:, 0], high=size[:, 1])
    system = k["System"](box, k["Particles"].from_arrays(pos), [pot])
    system.potential_and_force()
    rng = np.random.default_rng(12)
    moved = pos.copy()
    moved[rng.choice(len(pos), 50, replace=False)] += rng.uniform(-3.0, 3.0, size=(50, 3))
    system.particles.pos = moved
    v, f, w = system.potential_and_force()
    vr, fr, wr = ref.potential_and_force(system)
    assert abs(v - vr) <= TOL * abs(vr) and rel_norm(f, fr) < TOL and rel_norm(w, wr) < TOL
    system.box = k["Box"](low=size[:, 0], high=size[:, 1] * 1.03)
    v, f, w = system.potential_and_force()
    vr, fr, wr = ref.potential_and_force(system)
    assert abs(v - vr) <= TOL * abs(vr) and rel_norm(f, fr) < TOL and rel_norm(w, wr) < TOL
    assert pot.nlist_stats()["builds"] == 3


def test_lj_row_slabs_add_up():
    """Atom decomposition: forces of a slab of rows + V/W shares sum to the whole (both kernels)."""
    import ctypes as C

    from turtlemd_b200 import _device, _lib

    g = load_golden("lj_fcc_4000.npz")
    for method, fn in (("allpairs", "tmd_lj_allpairs"), ("cells", "tmd_lj_cells"), ("nlist", "tmd_lj_nlist")):
        system, pot = _lj_system(g, "lj_fcc_4000.npz", method="cells" if method == "nlist" else method)
        v, f, w = pot.potential_and_force(system)
        _, table, _ = pot._tables(system.particles)
        box = system.box.to_abi()
        pos = _device.to_device(g["pos"])
        force = _device.zeros(g["pos"].size)
        vw_sum = np.zeros(10)
        for first, count in ((0, 1500), (1500, 1), (1501, 2499)):
            vw = _device.zeros(10)
            handle = ()
            if method == "nlist":  # one list per slab, as one rank of an atom decomposition keeps
                h = C.c_void_p()
                _lib.call("tmd_nlist_create", C.byref(h), 0.3)
                handle = (h,)
            _lib.call(fn, *handle, _device.dptr(pos), None, 4000, C.byref(box), _lib.hptr(table), 1, 7, first, count,
                      _device.dptr(force), _device.dptr(vw), _device.stream_ptr())
            vw_sum += vw.cpu().numpy()
            if handle:
                _lib.call("tmd_nlist_destroy", handle[0])
        assert rel_norm(force.cpu().numpy().reshape(-1, 3), f) < 1e-13
        assert abs(vw_sum[0] - v) <= 1e-12 * abs(v)
        assert rel_norm(vw_sum[1:].reshape(3, 3), w) < 1e-12


@pytest.mark.parametrize("method", ["cells", "nlist"])
def test_lj_1m_sampled_rows_and_properties(method):
    """Config #4 size (N = 1 000 188): sampled rows against the oracle, Newton, symmetric virial."""
    from oracle import turtle_oracle as to

    k = _pkg()
    pos, size = _fluid(63)
    assert len(pos) == 1000188
    length = size[:, 1] - size[:, 0]
    pot = k["LennardJonesCut"](dim=3, shift=True, method=method)
    pot.set_parameters(SINGLE)
    system = k["System"](k["Box"](low=size[:, 0], high=size[:, 1]), k["Particles"].from_arrays(pos), [pot])
    v, f, w = system.potential_and_force()
    rows = np.random.default_rng(9).choice(len(pos), 48, replace=False)
    table, _ = to.lj_tables(SINGLE)
    fr, vr = to.lj_rows(pos, np.zeros(len(pos), dtype=int), length, 1.0 / length, [True] * 3, table, rows)
    assert rel_norm(f[rows], fr) < TOL
    assert np.max(np.abs(f.sum(axis=0))) < 1e-8 * np.max(np.abs(f))
    assert rel_norm(w, w.T) < 1e-12
    # energy per atom of the jittered fluid is an intensive quantity: compare with the 32k system
    pos32, size32 = _fluid(20)
    system32 = k["System"](k["Box"](low=size32[:, 0], high=size32[:, 1]), k["Particles"].from_arrays(pos32), [_lj()])
    v32 = system32.potential()
    assert abs(v / len(pos) - v32 / len(pos32)) < 2e-2 * abs(v32 / len(pos32))
    # ten VelocityVerlet steps conserve energy and keep the centre of mass at rest
    system.particles.vel = np.random.default_rng(2).standard_normal(pos.shape) * 0.5
    system.particles.vel -= system.particles.vel.mean(axis=0)
    e0 = v + 0.5 * np.sum(system.particles.vel**2)
    k["integrators"].VelocityVerlet(0.002).run(system, 10)
    e1 = system.particles.v_pot + 0.5 * np.sum(system.particles.vel**2)
    assert abs(e1 - e0) < 2e-5 * abs(e0)


def test_langevin_inertia_on_the_list_path_folds_potential_into_the_force_pass():
    """LangevinInertia calls force() and then potential() at the same positions (integrators.py:280-282).
    On the neighbour-list path the energy is taken from the force traversal; it must be the number a
    separate potential() pass gives, bit for bit, and half the traversals must be launched."""
    from turtlemd_b200 import _lib
    from turtlemd_b200.philox import PhiloxNormal

    k = _pkg()
    pos, size = _fluid(10)

    def make(method):
        pot = k["LennardJonesCut"](dim=3, shift=True, method=method, skin=0.3)
        pot.set_parameters(SINGLE)
        system = k["System"](k["Box"](low=size[:, 0], high=size[:, 1]), k["Particles"].from_arrays(pos), [pot])
        integ = k["integrators"].LangevinInertia(timestep=0.004, gamma=0.3, beta=1.25, rgen=PhiloxNormal(key=5))
        system.potential_and_force()
        return system, integ

    system, integ = make("nlist")
    assert system.energy_rides_with_forces()
    integ(system)
    before = _lib.launch_count()
    for _ in range(5):
        integ(system)
    per_step = (_lib.launch_count() - before) / 5
    v_step = system.particles.v_pot
    assert v_step == system.potential()  # a fresh energy-only pass over the same positions
    other, integ2 = make("allpairs")
    assert not other.energy_rides_with_forces()  # two formulas there (lennardjones.py:172 vs :254-255): two passes
    for _ in range(6):
        integ2(other)
    assert abs(other.particles.v_pot - v_step) <= 1e-10 * abs(v_step)
    assert per_step <= 8  # pre, gather/check, 2 gated build kernels, traversal, reduce, post (was 12 with two passes)


# ---- atom decomposition (multi-GPU path, ranks emulated on one device) -------------------------------------


@pytest.mark.parametrize("kind", ["vv", "inertia"])
@pytest.mark.parametrize("method", ["allpairs", "nlist"])
def test_atom_decomposition_three_emulated_ranks_match_single_gpu(kind, method):
    """Three 'ranks' in one process, each with its own copy of the system and its own row block; the
    position all-gather is done by hand between the step halves.  Result = the single-GPU trajectory."""
    from turtlemd_b200.parallel import AtomDecomposition, ShardPlan
    from turtlemd_b200.philox import PhiloxNormal

    k = _pkg()
    pos, size = _fluid(10)  # N = 4000
    vel = np.random.default_rng(6).standard_normal(pos.shape) * 0.9
    mass = np.random.default_rng(7).uniform(0.8, 1.6, size=len(pos))
    world, steps = 3, 12

    def make():
        pot = k["LennardJonesCut"](dim=3, shift=True, method=method, skin=0.3)
        pot.set_parameters(SINGLE)
        system = k["System"](k["Box"](low=size[:, 0], high=size[:, 1]),
                             k["Particles"].from_arrays(pos, vel=vel, mass=mass), [pot])
        if kind == "vv":
            integ = k["integrators"].VelocityVerlet(0.004)
        else:
            integ = k["integrators"].LangevinInertia(timestep=0.004, gamma=0.3, beta=1.25, rgen=PhiloxNormal(key=21))
        return system, integ

    ref_system, ref_integ = make()
    ref_system.potential_and_force()
    for _ in range(steps):
        ref_integ(ref_system)

    ranks = []
    for r in range(world):
        system, integ = make()
        dec = AtomDecomposition(system, integ, plan=ShardPlan(len(pos), world, r))
        dec.prepare()
        ranks.append(dec)
    for _ in range(steps):
        for dec in ranks:
            dec.step_move()
        for dst in ranks:  # the all-gather: everybody receives everybody's own rows
            for src in ranks:
                lo, cnt = src.plan.first * 3, src.plan.count * 3
                dst._pos_pad[lo: lo + cnt].copy_(src._pos_pad[lo: lo + cnt])
        for dec in ranks:
            dec.step_force()
    x = ranks[0].system.particles._pos.device().cpu().numpy().reshape(-1, 3)
    assert rel_norm(x, ref_system.particles.pos) < 1e-11
    v_total, w_total = 0.0, np.zeros((3, 3))
    for dec in ranks:
        lo, cnt = dec.plan.first, dec.plan.count
        v_own = dec.system.particles._vel.device().cpu().numpy().reshape(-1, 3)[lo: lo + cnt]
        assert rel_norm(v_own, ref_system.particles.vel[lo: lo + cnt]) < 1e-10
        v_share, w_share = dec.reduce_observables()
        v_total += v_share
        w_total += w_share
    assert abs(v_total - ref_system.particles.v_pot) <= 1e-11 * abs(ref_system.particles.v_pot)
    if kind == "vv":
        assert rel_norm(w_total, ref_system.particles.virial) < 1e-10


# ---- slab decomposition (halo exchange instead of the position all-gather) ----------------------------------


def _exchange_by_hand(ranks, full):
    """What NCCL does between real ranks: copy own slot ranges (all of them, or the halo parts)."""
    from turtlemd_b200.parallel import halo_ranges

    for dst in ranks:
        for src in ranks:
            if src is dst:
                continue
            if full:
                lo, cnt = src.plan.first, src.plan.count
                dst.records[lo: lo + cnt].copy_(src.records[lo: lo + cnt])
                dst.vel_s[lo * 3: (lo + cnt) * 3].copy_(src.vel_s[lo * 3: (lo + cnt) * 3])
                dst.force_s[lo * 3: (lo + cnt) * 3].copy_(src.force_s[lo * 3: (lo + cnt) * 3])
        if not full:
            for name in ("recv_up", "recv_down"):
                peer, a, b = halo_ranges(dst.plan, dst.halo)[name]
                dst.records[a:b].copy_(ranks[peer].records[a:b])


@pytest.mark.parametrize("kind", ["vv", "inertia"])
@pytest.mark.parametrize("world", [1, 3])
def test_slab_decomposition_emulated_ranks_match_single_gpu(kind, world):
    """Ranks emulated in one process: slot-ordered state, halo exchange by hand, full exchange +
    re-binning when any rank's displacement flag is up.  Result = the single-GPU trajectory."""
    from turtlemd_b200.parallel import ShardPlan, SlabDecomposition
    from turtlemd_b200.philox import PhiloxNormal

    k = _pkg()
    pos, size = _fluid(14)  # N = 10976: 8 cell layers of ~1372 atoms
    pos = pos + 3.0 * (size[:, 1] - size[:, 0]) * np.random.default_rng(5).integers(-1, 2, size=pos.shape)  # unwrapped input
    vel = np.random.default_rng(6).standard_normal(pos.shape) * 0.9
    mass = np.random.default_rng(7).uniform(0.8, 1.6, size=len(pos))
    steps = 40

    def make():
        pot = k["LennardJonesCut"](dim=3, shift=True, method="nlist", skin=0.3)
        pot.set_parameters(SINGLE)
        system = k["System"](k["Box"](low=size[:, 0], high=size[:, 1]),
                             k["Particles"].from_arrays(pos, vel=vel, mass=mass), [pot])
        if kind == "vv":
            integ = k["integrators"].VelocityVerlet(0.004)
        else:
            integ = k["integrators"].LangevinInertia(timestep=0.004, gamma=0.3, beta=1.25, rgen=PhiloxNormal(key=21))
        return system, integ

    ref_system, ref_integ = make()
    ref_system.potential_and_force()
    for _ in range(steps):
        ref_integ(ref_system)

    ranks = []
    for r in range(world):
        system, integ = make()
        dec = SlabDecomposition(system, integ, plan=ShardPlan(len(pos), world, r))
        dec.prepare()
        ranks.append(dec)
    for _ in range(steps):
        for dec in ranks:
            dec.step_move()
        if any([dec.moved_too_far() for dec in ranks]):
            _exchange_by_hand(ranks, full=True)
            for dec in ranks:
                dec.rebuild()
        else:
            _exchange_by_hand(ranks, full=False)
        for dec in ranks:
            dec.step_force()
    assert ranks[0].rebuilds >= 3  # the initial build and at least two on the way
    _exchange_by_hand(ranks, full=True)
    v_total, w_total = 0.0, np.zeros((3, 3))
    for dec in ranks:
        x, v, f = dec.gather_state()
        assert rel_norm(x, ref_system.particles.pos) < 1e-11
        assert rel_norm(v, ref_system.particles.vel) < 1e-9
        assert rel_norm(f, ref_system.particles.force) < 1e-9
        v_share, w_share = dec.reduce_observables()
        v_total += v_share
        w_total += w_share
    assert abs(v_total - ref_system.particles.v_pot) <= 1e-10 * abs(ref_system.particles.v_pot)
    if kind == "vv":
        assert rel_norm(w_total, ref_system.particles.virial) < 1e-9


def test_slab_decomposition_refuses_thin_slabs():
    from turtlemd_b200.parallel import ShardPlan, SlabDecomposition

    k = _pkg()
    pos, size = _fluid(10)  # N = 4000: six cell layers, eight ranks would own less than a halo each
    pot = k["LennardJonesCut"](dim=3, shift=True, method="nlist", skin=0.3)
    pot.set_parameters(SINGLE)
    system = k["System"](k["Box"](low=size[:, 0], high=size[:, 1]), k["Particles"].from_arrays(pos), [pot])
    dec = SlabDecomposition(system, k["integrators"].VelocityVerlet(0.004), plan=ShardPlan(len(pos), 8, 3))
    with pytest.raises(NotImplementedError):
        dec.prepare()


@pytest.mark.parametrize("kind,graph,physics", [("slab", "", "vv"), ("slab", "graph", "vv"), ("atoms", "", "vv"),
                                                ("atoms", "graph", "vv"), ("atoms", "graph", "langevin"),
                                                ("atoms", "", "langevin"), ("atoms", "graph", "dimer"),
                                                ("slabrun", "", "vv"), ("slabrun", "graph", "vv"),
                                                ("slabrun", "graph", "langevin"), ("slabrun", "graph", "dimer")])
def test_two_real_ranks_match_single_gpu(kind, graph, physics):
    """Two processes, two GPUs, NCCL: tests/multi_gpu_check.py compares the decomposed trajectory with
    the single-GPU one.  Skipped on a one-GPU box (the emulated-rank tests above cover the logic)."""
    import socket
    import subprocess
    import sys

    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    with socket.socket() as sock:
        sock.bind(("127.0.0.1", 0))
        port = sock.getsockname()[1]
    script = str(GOLDEN.parent / "multi_gpu_check.py")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", str(port), script, kind, "24", "40", graph or "eager", physics]
    proc = subprocess.run(cmd, capture_output=True, text=True, timeout=400)
    assert proc.returncode == 0 and "MULTI_GPU_CHECK ok" in proc.stdout, proc.stdout[-3000:] + proc.stderr[-3000:]


# ---- observables ------------------------------------------------------------------------


def test_degenerate_sizes():
    """One and two particles, a type nobody has, a potential list that is empty: the reference's answers
    (zeros, no exception) through the device path and the host-buffer path."""
    from oracle import turtle_oracle as to
    from turtlemd_b200.system import Box, Particles, System

    k = _pkg()
    box = Box(low=[0.0, 0.0, 0.0], high=[6.0, 6.0, 6.0])
    one = {0: {"sigma": 1.0, "epsilon": 1.0, "rcut": 2.5}}
    length, ilength = to.box_lengths([0, 0, 0], [6, 6, 6], [True] * 3)
    table, _ = to.lj_tables(one)
    for pos in (np.array([[1.0, 2.0, 3.0]]), np.array([[1.0, 2.0, 3.0], [1.9, 2.4, 5.8]])):
        lj = k["LennardJonesCut"](dim=3, shift=True)
        lj.set_parameters(one)
        dw = k["DoubleWellPair"](types=(0, 0), dim=3)
        dw.set_parameters({"rzero": 1.0, "width": 0.25, "height": 6.0})
        system = System(box, Particles.from_arrays(pos), [lj, dw])
        v, f, w = system.potential_and_force()
        tidx = np.zeros(len(pos), dtype=int)
        v_lj, f_lj, w_lj = to.lj_potential_and_force(pos, tidx, length, ilength, [True] * 3, table)
        par = to.dwpair_params(1.0, 0.25, 6.0)
        v_dw = to.dwpair_potential(pos, tidx, length, ilength, [True] * 3, (0, 0), par)
        f_dw, w_dw = to.dwpair_force(pos, tidx, length, ilength, [True] * 3, (0, 0), par)
        assert v == pytest.approx(v_lj + v_dw, rel=1e-12, abs=1e-300)
        assert np.allclose(f, f_lj + f_dw, rtol=1e-12, atol=0) and np.allclose(w, w_lj + w_dw, rtol=1e-12, atol=0)
        assert np.allclose(lj.force(system)[0], f_lj, rtol=1e-12, atol=0)
        for integ in (k["integrators"].VelocityVerlet(0.001), k["integrators"].Verlet(0.001)):
            integ(system)
        assert np.all(np.isfinite(system.particles.pos))
    # a pair potential whose types are absent: zeros
    system = System(box, Particles.from_arrays(np.array([[1.0, 1.0, 1.0], [2.0, 2.0, 2.0]])),
                    [k["DoubleWellPair"](types=(5, 6), dim=3)])
    system.potentials[0].set_parameters({"rzero": 1.0, "width": 0.25, "height": 6.0})
    v, f, w = system.potential_and_force()
    assert v == 0.0 and not f.any() and not w.any()
    # one-dimensional replicas, a single one
    reps = System(Box(periodic=[False]), Particles.from_arrays(np.array([[-1.0]])), [k["DoubleWell"](a=1.0, b=2.0, c=0.0)])
    reps.potential_and_force()
    from turtlemd_b200.philox import PhiloxNormal

    k["integrators"].LangevinInertia(timestep=0.002, gamma=0.3, beta=1.0 / 0.07, rgen=PhiloxNormal(3)).run(reps, 7)
    assert np.isfinite(reps.particles.pos).all() and reps.particles.pos.shape == (1, 1)


def test_thermo_uses_the_device_reduction_and_matches_the_host_formula():
    """System.thermo() after device-resident steps: kinetic tensor reduced on the GPU (velocities are
    not downloaded), same numbers as the reference's NumPy expressions (system.py:101-130)."""
    k = _pkg()
    pos, size = _fluid(10)
    rng = np.random.default_rng(14)
    vel = rng.standard_normal(pos.shape)
    mass = rng.uniform(0.5, 2.0, size=len(pos))
    pot = _lj()
    system = k["System"](k["Box"](low=size[:, 0], high=size[:, 1]),
                         k["Particles"].from_arrays(pos, vel=vel, mass=mass), [pot])
    system.potential_and_force()
    k["integrators"].VelocityVerlet(0.002).run(system, 5)
    part = system.particles
    assert part._vel.dev_valid and not part._vel.host_valid
    thermo = system.thermo()
    assert not part._vel.host_valid  # thermo() did not pull the velocities
    v, m = part.vel, part.mass
    kin = 0.5 * np.einsum("ij,ik->jk", v * m, v)
    volume = float(np.prod(size[:, 1] - size[:, 0]))
    assert thermo["ekin"] == pytest.approx(kin.trace() / len(pos), rel=1e-13)
    assert thermo["temperature"] == pytest.approx(np.mean(2.0 * np.diagonal(kin) / (len(pos) - 1)), rel=1e-13)
    assert thermo["pressure"] == pytest.approx(((part.virial + 2.0 * kin) / volume).trace() / 3.0, rel=1e-12)
    assert thermo["vpot"] == pytest.approx(part.v_pot / len(pos), rel=1e-15)


def test_atom_decomposition_graph_replay_equals_eager():
    """The CUDA-graph replay of a step (kernels incl. the gated rebuild) is bit-identical to eager stepping."""
    from turtlemd_b200.parallel import AtomDecomposition, ShardPlan

    k = _pkg()
    pos, size = _fluid(10)
    vel = np.random.default_rng(16).standard_normal(pos.shape) * 1.1
    out = []
    for graphed in (False, True):
        pot = k["LennardJonesCut"](dim=3, shift=True, method="nlist", skin=0.3)
        pot.set_parameters(SINGLE)
        system = k["System"](k["Box"](low=size[:, 0], high=size[:, 1]), k["Particles"].from_arrays(pos, vel=vel), [pot])
        dec = AtomDecomposition(system, k["integrators"].VelocityVerlet(0.004), plan=ShardPlan(len(pos), 1, 0))
        dec.prepare()
        if graphed:
            assert dec.capture(warm_steps=3)
            done = 3
        else:
            done = 0
        for _ in range(40 - done):
            dec.step()
        out.append((system.particles.pos.copy(), system.particles.vel.copy(), pot.nlist_stats()))
    (xe, ve, se), (xg, vg, sg) = out
    assert np.array_equal(xe, xg) and np.array_equal(ve, vg)
    assert sg["builds"] == se["builds"] >= 3 and sg["calls"] == se["calls"]
