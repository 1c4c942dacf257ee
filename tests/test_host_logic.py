"""CPU tests: the C-ABI library loads and exports what the header declares, and the host logic
(box, particles, parameter tables, Langevin constants, driver loop, twin RNG) behaves like the
reference.  No compute call is made on the device here."""
from __future__ import annotations

import logging
import os
import pathlib
import re

import numpy as np
import pytest

from conftest import REPO, load_golden

from turtlemd_b200 import _lib
from turtlemd_b200.integrators import (
    LangevinInertia,
    LangevinOverdamped,
    MDIntegrator,
    VelocityVerlet,
    Verlet,
    multivariate_normal,
)
from turtlemd_b200.philox import PhiloxNormal
from turtlemd_b200.potentials import DoubleWell, DoubleWellPair, LennardJonesCut, RectangularWell
from turtlemd_b200.potentials.lennardjones import generate_pair_interactions, mix_parameters
from turtlemd_b200.potentials.potential import Potential
from turtlemd_b200.simulation import MDSimulation
from turtlemd_b200.simulation.stopping import MaxSteps, SoftExit, StopCondition
from turtlemd_b200.system import Box, Particles, System, TriclinicBox
from turtlemd_b200.system.box import guess_dimensionality
from turtlemd_b200.system.particles import (
    generate_maxwell_velocities,
    kinetic_energy,
    kinetic_temperature,
    linear_momentum,
    pressure_tensor,
    zero_momentum,
)
from turtlemd_b200.tools import generate_lattice, particles_from_xyz_file, read_xyz_file


# ---- the C ABI ------------------------------------------------------------------------------------


def _header_functions():
    text = (REPO / "include" / "turtlemd_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(tmd_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    names = _header_functions()
    assert len(names) >= 40
    for name in names:
        assert hasattr(lib, name), f"{name} is declared in include/turtlemd_b200.h but not exported"
    assert sorted(_lib.PROTOTYPES) == names  # the ctypes binding covers exactly the header
    assert lib.tmd_abi_version() == 2
    assert lib.tmd_status_string(3).decode().startswith("no CUDA device")


def test_struct_layouts_match_the_header():
    import ctypes as C

    assert C.sizeof(_lib.Box) == 4 * 4 + 8 * 6
    assert C.sizeof(_lib.Rng) == 32  # key, offset, step_counter, step_stride (ABI 2)
    assert C.sizeof(_lib.LangevinConsts) == 8 * 11
    out = _lib.LangevinConsts()
    _lib.call("tmd_langevin_inertia_constants", 0.002, 0.3, 1.0, C.byref(out))
    g = load_golden("langevin_constants.npz")
    assert out.c0 == pytest.approx(float(g["a_c0"]), rel=1e-15)
    assert out.a1 == pytest.approx(float(g["a_a1"]), rel=1e-15)


def test_product_fails_loudly_without_a_device():
    if _lib.device_count() > 0:
        pytest.skip("a CUDA device is present")
    part = Particles(dim=3)
    part.add_particle([1.0, 1.0, 1.0], ptype=0)
    part.add_particle([2.0, 1.0, 1.0], ptype=0)
    pot = LennardJonesCut()
    pot.set_parameters({0: {"sigma": 1, "epsilon": 1, "rcut": 2.5}})
    system = System(Box(high=[10, 10, 10]), part, [pot])
    for call in (system.potential_and_force, system.force, system.potential,
                 lambda: pot.force(system), lambda: VelocityVerlet(0.002)(system),
                 lambda: DoubleWell().potential(system)):
        with pytest.raises(_lib.NoDeviceError):
            call()
    import ctypes as C

    vw = np.zeros(10)
    pos = np.zeros((4, 1))
    assert _lib.load().tmd_host_doublewell(_lib.hptr(pos), 4, 1, 1.0, 1.0, 0.0, 1, None, _lib.hptr(vw)) == _lib.ERR_NO_DEVICE
    del C


def test_no_oracle_import_in_the_product():
    for path in (REPO / "turtlemd_b200").rglob("*.py"):
        assert "oracle" not in path.read_text(), f"{path} mentions the oracle"


# ---- box ------------------------------------------------------------------------------------------------


def test_box_minimum_image_and_wrap():
    g = load_golden("box_pbc.npz")
    box = Box(low=g["low"], high=g["high"], periodic=[bool(p) for p in g["periodic"]])
    assert np.array_equal(box.pbc_dist_matrix(g["dist"]), g["dist_matrix"])
    assert np.array_equal(np.array([box.pbc_dist(d) for d in g["dist"]]), g["dist_each"])
    assert np.array_equal(box.pbc_wrap(g["dist"]), g["wrapped"])
    abi = box.to_abi()
    assert abi.dim == 3 and list(abi.periodic) == [1, 0, 1] and abi.length[1] == 0.0
    assert abi.ilength[0] == box.ilength[0]


def test_box_construction(caplog):
    with caplog.at_level(logging.WARNING):
        box = Box(periodic=[False])
    assert "Set box low values" in caplog.text
    assert box.dim == 1 and np.isinf(box.length[0]) and box.ilength[0] == 0.0
    assert np.isinf(box.volume())
    assert Box(high=[2, 3, 4]).volume() == pytest.approx(24.0)
    assert list(Box(high=[1, 1], periodic=[True, False]).dof) == [1, 0]
    with pytest.raises(ValueError):
        guess_dimensionality(low=[0, 0], high=[1])
    with caplog.at_level(logging.WARNING):
        assert guess_dimensionality() == 1
    with pytest.raises(ValueError):
        TriclinicBox(high=[1, 1, 1])
    tri = TriclinicBox(high=[10.0, 10.0, 10.0], alpha=75, beta=60, gamma=70)
    assert tri.volume() > 0
    d = tri.pbc_dist(np.array([9.0, 8.0, 7.0]))
    assert np.dot(d, d) <= np.dot([9.0, 8.0, 7.0], [9.0, 8.0, 7.0])
    cell = tri.to_abi()  # the tmd_cell of the triclinic pair kernel
    assert cell.dim == 3 and cell.half_width == tri.shortest_width_half
    assert np.array_equal(np.array(cell.matrix[:]).reshape(3, 3), tri.box_matrix)
    assert np.array_equal(np.array(cell.translations[:]).reshape(27, 3), tri.translations)


# ---- particles -----------------------------------------------------------------------------------------


def test_particles_container():
    part = Particles(dim=2)
    assert part.npart == 0 and part.pos.shape == (1, 2) and part.v_pot is None
    part.add_particle(np.array([1.0, 2.0]), mass=2.0, name="A", ptype=0)
    part.add_particle(np.array([3.0, 4.0]), vel=np.array([1.0, 0.0]), mass=4.0, name="B", ptype=1)
    part.add_particle(np.array([5.0, 6.0]), name="C", ptype=1)
    assert len(part) == 3 and part.pos.shape == (3, 2)
    assert part.imass == pytest.approx(np.array([[0.5], [0.25], [1.0]]))
    assert list(part.pairs()) == [(0, 1, 0, 1), (0, 2, 0, 1), (1, 2, 1, 1)]
    items = list(part)
    assert items[1]["name"] == "B" and items[1]["type"] == 1 and items[1]["vel"][0] == 1.0
    part.virial = np.ones((2, 2))
    sub = part[1:]
    assert sub.npart == 2 and np.array_equal(part.virial, np.zeros((2, 2)))
    sub.pos[0, 0] = -7.0
    assert part.pos[1, 0] == -7.0  # basic slices are views
    one = part[[0]]
    one.pos[0, 0] = 99.0
    assert part.pos[0, 0] == 1.0   # fancy indices are copies
    part.vel = np.array([[1.0, 0.0], [0.0, 1.0], [1.0, 1.0]])
    assert linear_momentum(part) == pytest.approx([3.0, 5.0])
    kin, ekin = kinetic_energy(part)
    assert ekin == pytest.approx(0.5 * (2 * 1 + 4 * 1 + 1 * 2))
    zero_momentum(part)
    assert linear_momentum(part) == pytest.approx([0.0, 0.0], abs=1e-14)
    bulk = Particles.from_arrays(np.zeros((5, 3)), mass=np.arange(1, 6), ptype=[0, 1, 0, 1, 0])
    assert bulk.npart == 5 and bulk.mass.shape == (5, 1) and bulk.imass[4, 0] == 0.2


def test_maxwell_velocities_and_thermo():
    pos, size = generate_lattice("fcc", [3, 3, 3], density=0.9)
    part = Particles.from_arrays(pos)
    generate_maxwell_velocities(part, rgen=np.random.default_rng(1), temperature=2.0)
    temp, _, kin = kinetic_temperature(part, 1.0)
    assert temp == pytest.approx(2.0)
    assert linear_momentum(part) == pytest.approx(np.zeros(3), abs=1e-12)
    box = Box(low=size[:, 0], high=size[:, 1])
    system = System(box, part)  # no potentials: nothing needs the GPU
    assert system.potential_and_force() == (0, None, None)
    part.virial = np.zeros((3, 3))
    th = system.thermo()
    assert th["vpot"] == 0 and th["ekin"] == pytest.approx(kin.trace() / 108)
    _, pressure = pressure_tensor(part, box.volume(), kin_tensor=kin)
    assert th["pressure"] == pytest.approx(pressure)
    assert System(box, Particles(dim=3)).thermo()["ekin"] is None


@pytest.mark.parametrize("tag", ["3d_500", "1d_64", "2d_301"])
def test_maxwell_velocities_match_reference_fixture(tag):
    """generate_maxwell_velocities / zero_momentum against the reference fed the same stream
    (tests/golden/make_golden_maxwell.py; system/particles.py:140-153, :227-269)."""
    from oracle.philox_ref import OraclePhiloxNormal

    g = load_golden(f"maxwell_{tag}.npz")
    n, dim = g["vel"].shape
    dof = None if g["dof"].size == 0 else list(g["dof"])
    for make in (OraclePhiloxNormal, PhiloxNormal):  # unknown generator: host draws; the twin: device when there is one
        part = Particles.from_arrays(np.zeros((n, dim)), mass=g["mass"])
        rgen = make(key=int(g["key"]))
        generate_maxwell_velocities(part, rgen, temperature=float(g["temperature"]),
                                    boltzmann=float(g["boltzmann"]), dof=dof, momentum=bool(g["momentum"]))
        assert rgen.position == int(g["draws"])
        assert np.allclose(part.vel, g["vel"], rtol=1e-12, atol=1e-14)
    part.vel = g["vel"] + g["drift"]
    zero_momentum(part, dim=[bool(m) for m in g["mask"]])
    assert np.allclose(part.vel, g["vel_masked"], rtol=1e-13, atol=1e-15)


# ---- potentials: host-side parameter logic ----------------------------------------------------------


def test_lj_parameter_tables_match_reference():
    two = {0: {"sigma": 1.0, "epsilon": 1.0, "rcut": 2.5}, 1: {"sigma": 1.2, "epsilon": 0.7, "rcut": 3.0}}
    three = dict(two)
    three[2] = {"sigma": 0.9, "epsilon": 1.3, "rcut": 2.2}
    three[(0, 2)] = {"sigma": 1.1, "epsilon": 0.5, "rcut": 2.0}
    cases = (("lj_two_types_geometric.npz", two, dict(mixing="geometric")),
             ("lj_two_types_arithmetic_noshift.npz", two, dict(mixing="arithmetic", shift=False)),
             ("lj_three_types_sixthpower_explicit.npz", three, dict(mixing="sixthpower")))
    for name, params, kw in cases:
        g = load_golden(name)
        pot = LennardJonesCut(**kw)
        pot.set_parameters(params)
        part = Particles.from_arrays(g["pos"], ptype=g["ptype"])
        types, table, tidx = pot._tables(part)
        assert types == [int(t) for t in g["types"]]
        assert np.array_equal(table, g["table"])
        assert np.array_equal(np.array(types)[tidx], g["ptype"])
    assert ((0, 2), 0) in pot.params  # the reference's junk entries for tuple keys are reproduced
    assert pot.params[0, 2]["sigma"] == 1.1 and pot.params[2, 0] is pot.params[0, 2]
    with pytest.raises(TypeError):
        str(pot)  # mixed tuple/int keys cannot be sorted -- in the reference neither
    plain = LennardJonesCut(shift=False)
    plain.set_parameters(two)
    text = str(plain)
    assert "Shift potential: no" in text and "0-1" in text and "3.0000" in text
    with pytest.raises(ValueError):
        mix_parameters(1, 1, 1, 1, 1, 1, mixing="nope")
    assert mix_parameters(1.0, 2.0, 3.0, 4.0, 8.0, 12.0) == pytest.approx((2.0, 4.0, 6.0))
    assert mix_parameters(1.0, 2.0, 3.0, 4.0, 8.0, 12.0, "arithmetic") == pytest.approx((2.0, 5.0, 7.5))
    pairs = generate_pair_interactions({0: two[0], 1: two[1]}, "geometric")
    assert set(pairs) == {(0, 0), (0, 1), (1, 0), (1, 1)}
    lone = Particles.from_arrays(np.zeros((3, 3)), ptype=[0, 0, 7])
    with pytest.raises(KeyError):
        pot._tables(lone)
    with pytest.raises(ValueError):
        LennardJonesCut(method="magic")


def test_well_parameters(caplog):
    pot = DoubleWellPair(types=(0, 1))
    with caplog.at_level(logging.WARNING):
        pot.set_parameters({"rzero": 1.0, "width": 0.25, "height": 6.0, "extra": 1.0})
    assert 'Ignored unknown parameter "extra"' in caplog.text
    assert pot.params["width2"] == 0.0625 and pot.params["rwidth"] == 1.25 and pot.params["height4"] == 24.0
    assert pot.min_max() == (1.0, 1.5, 1.25)
    assert pot.activate(0, 1) and pot.activate(1, 0) and not pot.activate(0, 0) and not pot.activate(1, 1)
    assert pot._potential_function(np.array([1.0, 1.25, 1.5])) == pytest.approx([0.0, 6.0, 0.0])
    assert DoubleWell(a=1, b=2, c=3).params == {"a": 1, "b": 2, "c": 3}
    assert "Potential: 1D double well potential" in str(DoubleWell())
    with caplog.at_level(logging.WARNING):
        rect = RectangularWell(left=2.0, right=1.0)
    assert "left >= right" in caplog.text
    system = System(Box(periodic=[False]), Particles.from_arrays(np.array([[0.5], [3.0]])))
    assert RectangularWell().potential(system) == float("inf")
    with caplog.at_level(logging.WARNING):
        f, w = rect.force(system)
    assert np.all(f == 0) and "not implemented" in caplog.text

    class Custom(Potential):
        def potential(self, system):
            return 1.0

        def force(self, system):
            return np.zeros((2, 1)), np.zeros((1, 1))

    with caplog.at_level(logging.INFO):
        Custom().set_parameters({"a": 1})
    assert "ignoring the given parameters" in caplog.text
    assert Custom().potential_and_force(system)[0] == 1.0


# ---- integrators: host-side logic ---------------------------------------------------------------------


def test_langevin_parameters_match_reference():
    g = load_golden("langevin_constants.npz")
    part = Particles.from_arrays(np.zeros((4, 1)), mass=1.0 / g["imass"][:, 0])
    system = System(Box(periodic=[False]), part)
    for tag in "abc":
        dt, gamma, beta = (float(x) for x in g[f"{tag}_in"])
        integ = LangevinInertia(timestep=dt, gamma=gamma, beta=beta)
        integ.initiate_parameters(system)
        par = integ.param
        assert par.c0 == pytest.approx(float(g[f"{tag}_c0"]), rel=1e-15)
        assert par.a1 == pytest.approx(float(g[f"{tag}_a1"]), rel=1e-15)
        for key in ("a2", "b1", "b2"):
            assert getattr(par, key) == pytest.approx(g[f"{tag}_{key}"], rel=1e-14)
        assert np.array(par.cov) == pytest.approx(g[f"{tag}_cov"], rel=1e-14)
        assert np.array(par.cho) == pytest.approx(g[f"{tag}_cho"], rel=1e-14)
        assert len(par.mean) == 4 and np.all(par.mean[0] == 0)
        assert integ._consts.b2f * g["imass"][1, 0] == pytest.approx(par.b2[1, 0], rel=1e-15)
    # the reference's own parameter KATs (tests/test_integrators.py:305-353)
    one = System(Box(periodic=[False]), Particles.from_arrays(np.array([[-1.0]])))
    for tag, (dt, gamma) in (("1", (0.002, 0.3)), ("2", (0.0005, 0.5))):
        k = load_golden(f"reference_langevin_parameter{tag}.npz")
        integ = LangevinInertia(timestep=dt, gamma=gamma, beta=1.0)
        integ.initiate_parameters(one)
        assert integ.param.c0 == pytest.approx(float(k["c0"])) and integ.param.a1 == pytest.approx(float(k["a1"]))
        assert np.array(integ.param.cho) == pytest.approx(k["cho"])
        assert np.array(integ.param.cov) == pytest.approx(k["cov"])
    assert integ.description == "Langevin overdamped integrator"  # sic, as in the reference


def test_langevin_host_draw_equals_numpy_multivariate_normal():
    """tests/test_integrators.py:524-539 of the reference."""
    one = System(Box(periodic=[False]), Particles.from_arrays(np.array([[-1.0]])))
    integ = LangevinInertia(timestep=0.002, gamma=0.3, beta=1.0, rgen=np.random.default_rng(123))
    integ.initiate_parameters(one)
    pos, vel = integ.draw_random_numbers(one)
    pos2, vel2 = np.random.default_rng(123).multivariate_normal(
        integ.param.mean[0], integ.param.cov[0], method="cholesky")
    assert pos[0][0] == pytest.approx(pos2) and vel[0][0] == pytest.approx(vel2)
    xv = multivariate_normal(np.random.default_rng(1), np.zeros(2), np.eye(2), 3)
    assert xv.shape == (3, 2)


def test_integrator_constructors():
    over = LangevinOverdamped(timestep=0.002, gamma=0.3, rgen=np.random.default_rng(0), beta=2.0)
    part = Particles.from_arrays(np.zeros((2, 1)), mass=[1.0, 4.0])
    over.initiate_parameters(System(Box(periodic=[False]), part))
    assert over.sigma == pytest.approx(np.sqrt(2 * 0.002 * part.imass / (2.0 * 0.3)))
    assert over.bddt == pytest.approx(0.002 * part.imass / 0.3)
    vv = VelocityVerlet(0.004)
    assert vv.half_timestep == 0.002 and vv.dynamics == "NVE" and isinstance(vv, MDIntegrator)
    ver = Verlet(0.5)
    assert ver.timestep_squared == 0.25 and ver.half_idt == 1.0 and ver.previous_pos is None
    default = LangevinInertia(timestep=0.002, gamma=0.3, beta=1.0, seed=5)  # the reference's default (integrators.py:216)
    assert isinstance(default.rgen, np.random.Generator)
    assert default.rgen.normal() == np.random.default_rng(5).normal()
    fast = LangevinInertia(timestep=0.002, gamma=0.3, beta=1.0, rgen=PhiloxNormal(key=5))
    assert isinstance(fast.rgen, PhiloxNormal) and fast.rgen.key == 5


# ---- driver loop ---------------------------------------------------------------------------------------


class _CountingIntegrator(MDIntegrator):
    def __init__(self):
        super().__init__(timestep=1.0, description="counts", dynamics="none")
        self.calls = 0

    def integration_step(self, system):
        self.calls += 1


def test_mdsimulation_contract(tmp_path, caplog):
    system = System(Box(periodic=[False]), Particles.from_arrays(np.zeros((1, 1))))
    integ = _CountingIntegrator()
    sim = MDSimulation(system, integ, steps=10, start=5)
    assert sum(1 for _ in sim.run()) == 11 and integ.calls == 10  # step 0 + 10 integrated
    assert sim.cycle == {"step": 15, "steps": 10, "start": 5, "stepno": 10}
    assert sim.step() is system and integ.calls == 10            # stopped: nothing happens
    assert any(isinstance(c, SoftExit) for c in MDSimulation(system, integ, 3, stop_conditions=[MaxSteps()]).stop_conditions)

    class Never(StopCondition):
        def stop(self, simulation):
            return False

    sim = MDSimulation(system, _CountingIntegrator(), steps=3, stop_conditions=[Never()])
    sim.exe_dir = tmp_path
    gen = sim.run()
    for _ in range(6):
        next(gen)
    (tmp_path / "EXIT").write_text("stop")
    with caplog.at_level(logging.INFO):
        assert list(gen) == []
    assert "Found exit file" in caplog.text
    cwd = os.getcwd()
    try:
        os.chdir(tmp_path)
        assert MDSimulation(system, _CountingIntegrator(), steps=3).stop()  # default exe_dir is the cwd
    finally:
        os.chdir(cwd)
    batched = MDSimulation(system, _CountingIntegrator(), steps=10)
    assert [batched.cycle["stepno"] for _ in batched.run_batch(4)] == [0, 4, 8, 10]


# ---- tools ------------------------------------------------------------------------------------------------


def test_lattice_and_xyz(tmp_path):
    pos, size = generate_lattice("fcc", [2, 2, 2], lattice_constant=2.0)
    assert pos.shape == (32, 3) and np.array_equal(size, [[0, 4.0]] * 3)
    assert np.array_equal(pos[:5], [[0, 0, 0], [1, 1, 0], [0, 1, 1], [1, 0, 1], [0, 0, 2]])
    pos, size = generate_lattice("sq2", [3, 2], density=0.5)
    assert pos.shape == (12, 2) and size[0, 1] == pytest.approx(3 * 2.0) and size[1, 1] == pytest.approx(2 * 2.0)
    assert generate_lattice("sc")[0].shape == (1, 3)
    with pytest.raises(ValueError):
        generate_lattice("quasicrystal")
    xyz = tmp_path / "conf.xyz"
    xyz.write_text("2\nframe one\nAr 0 0 0 1 1 1\nKr 1 2 3 0 0 0\n2\nframe two\nAr 0 0 1\nKr 1 2 4\n")
    frames = list(read_xyz_file(xyz))
    assert len(frames) == 2 and frames[0].comment == "frame one" and frames[1].xyz[1, 2] == 4.0
    part = particles_from_xyz_file(xyz, dim=3, masses={"Kr": 83.8})
    assert part.npart == 2 and part.mass[1, 0] == 83.8 and part.vel[0, 0] == 1.0
    assert set(part.ptype) == {0, 1}
    bad = tmp_path / "bad.xyz"
    bad.write_text("two\n")
    with pytest.raises(ValueError):
        list(read_xyz_file(bad))


# ---- the host twin of the in-kernel noise stream -----------------------------------------------------


def test_philox_twin_matches_oracle_and_numpy():
    from oracle import philox_ref as pr

    for key in (0, 7, 2**64 - 3):
        twin = PhiloxNormal(key)
        assert np.array_equal(twin.raw(40, 5), np.random.Philox(key=key).random_raw(45)[5:])
        z = twin.normal(0.0, 1.0, size=(50, 3))
        assert np.array_equal(z.reshape(-1), pr.normals(key, 0, 150))
        assert twin.position == 150
        assert twin.normal() == pr.normals(key, 150, 1)[0]
        scaled = twin.normal(1.0, np.array([[2.0], [3.0]]), size=(2, 2))
        base = pr.normals(key, 151, 4).reshape(2, 2)
        assert np.array_equal(scaled, 1.0 + np.array([[2.0], [3.0]]) * base)


def test_langevin_run_refuses_to_shard_a_force_field_the_fused_kernel_does_not_cover():
    """first / n_global used to be dropped silently on the general path (identical noise on
    every shard).  Host logic only: the refusal comes before any kernel."""
    from turtlemd_b200.integrators import LangevinInertia
    from turtlemd_b200.philox import PhiloxNormal
    from turtlemd_b200.potentials import DoubleWell
    from turtlemd_b200.system import Box, Particles, System

    class Other(DoubleWell):  # a subclass is not the fused kernel's potential
        pass

    system = System(Box(periodic=[False]), Particles.from_arrays(np.full((8, 1), -1.0)), [Other(a=1.0, b=2.0, c=0.0)])
    integ = LangevinInertia(timestep=0.002, gamma=0.3, beta=1.0 / 0.07, rgen=PhiloxNormal(key=1))
    with pytest.raises(NotImplementedError):
        integ.run(system, 3, first=8, n_global=16)


def test_slab_run_phase_ids_match_the_header():
    """parallel.SR_PHASES drives tmd_slabrun_phase by number: the numbers are the header's enum."""
    from turtlemd_b200.parallel import SR_ERRORS, SR_PHASES, _sr_error_text

    text = (REPO / "include" / "turtlemd_b200.h").read_text()
    body = re.search(r"enum tmd_slabrun_phase_id \{(.*?)\};", text, flags=re.S).group(1)
    header = {name[len("TMD_SR_"):]: int(value) for name, value in re.findall(r"(TMD_SR_[A-Z0-9_]+)\s*=\s*(\d+)", body)}
    assert header == SR_PHASES
    assert _sr_error_text(0) == "ok" and "barrier" in _sr_error_text(1) and len(SR_ERRORS) == 7
    assert ";" in _sr_error_text(1 | 4)


def test_vectorised_host_draw_equals_the_per_particle_draws():
    """LangevinInertia.draw_random_numbers: one (n, dim, 2) call on a NumPy Generator + a batched product against the
    reference's loop of per-particle multivariate_normal calls (integrators.py:284-315): the same stream position afterwards
    and the same numbers (to the last bit or two: the 2 x 2 products may be contracted differently)."""
    import turtlemd_b200.integrators as integrators

    part = Particles.from_arrays(np.zeros((50, 3)), mass=1.0 + 0.5 * (np.arange(50) % 3))
    system = System(Box(high=[10, 10, 10]), part)
    fast = LangevinInertia(timestep=0.002, gamma=0.3, beta=2.0, seed=11)
    pos_f, vel_f = fast.draw_random_numbers(system)
    slow = LangevinInertia(timestep=0.002, gamma=0.3, beta=2.0, seed=11)
    original = integrators.multivariate_normal
    integrators.multivariate_normal = lambda rgen, mean, cho, dim: original(rgen, mean, cho, dim)  # not the original object
    try:
        pos_s, vel_s = slow.draw_random_numbers(system)
    finally:
        integrators.multivariate_normal = original
    assert pos_f.shape == pos_s.shape == (50, 3)
    assert np.allclose(pos_f, pos_s, rtol=1e-15, atol=0.0) and np.allclose(vel_f, vel_s, rtol=1e-14, atol=1e-18)
    assert fast.rgen.normal() == slow.rgen.normal()  # both streams stand at the same position


def test_touch_invalidates_what_the_potentials_derived_from_ptype():
    """ptype edited IN PLACE + Particles.touch(): the type table of LennardJonesCut is rebuilt (the caches
    were keyed on the identity of the ptype array alone)."""
    particles = Particles.from_arrays(np.zeros((4, 3)), ptype=np.zeros(4, dtype=int))
    pot = LennardJonesCut(dim=3, shift=True)
    pot.set_parameters({0: {"sigma": 1.0, "epsilon": 1.0, "rcut": 2.5}, 1: {"sigma": 1.2, "epsilon": 0.5, "rcut": 3.0}})
    types, table, tidx = pot._tables(particles)
    assert types == [0] and table.shape == (6, 1, 1)
    particles.ptype[2:] = 1
    assert pot._tables(particles)[0] == [0]  # same array, no touch(): the cached table (documented behaviour)
    particles.touch()
    types, table, tidx = pot._tables(particles)
    assert types == [0, 1] and table.shape == (6, 2, 2) and tidx.tolist() == [0, 0, 1, 1]
