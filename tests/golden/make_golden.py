"""Generate the golden fixtures in this directory from the live reference.

Run HERE (the build container), where the reference source tree is mounted
read-only at /root/reference:

    python tests/golden/make_golden.py

It imports the *reference itself* (TurtleMD 2023.3.1; the import needs a
one-line shim because ``turtlemd/version.py`` asks importlib.metadata for an
installed distribution), feeds it seeded inputs and stores inputs + outputs as
small ``.npz`` files.  The GPU box has no /root/reference: tests only read the
``.npz`` files.  Also collected here: the known-answer vectors and binary
fixtures of the reference's own tests for this path (file:line cited per key).
"""
from __future__ import annotations

import importlib.metadata as _md
import pathlib
import sys

import numpy as np

HERE = pathlib.Path(__file__).resolve().parent
REPO = HERE.parent.parent
REF = pathlib.Path("/root/reference")

_orig_version = _md.version
_md.version = lambda name: "2023.3.1" if name == "turtlemd" else _orig_version(name)
sys.path.insert(0, str(REF))
sys.path.insert(0, str(REPO))

import turtlemd.integrators as ref_integrators  # noqa: E402
from turtlemd.integrators import (  # noqa: E402
    LangevinInertia,
    LangevinOverdamped,
    VelocityVerlet,
    Verlet,
)
from turtlemd.potentials import DoubleWell, DoubleWellPair, LennardJonesCut  # noqa: E402
from turtlemd.simulation import MDSimulation  # noqa: E402
from turtlemd.system import Box, Particles, System  # noqa: E402
from turtlemd.system.particles import generate_maxwell_velocities  # noqa: E402
from turtlemd.tools import generate_lattice  # noqa: E402

from oracle.philox_ref import OraclePhiloxNormal  # noqa: E402  (duck-typed rgen twin)


def bulk_particles(pos, vel=None, mass=None, ptype=None):
    """Fill a reference ``Particles`` by array assignment (add_particle is O(N^2))."""
    pos = np.array(pos, dtype=float)
    n, dim = pos.shape
    part = Particles(dim=dim)
    part.pos = pos
    part.vel = np.zeros_like(pos) if vel is None else np.array(vel, dtype=float)
    part.force = np.zeros_like(pos)
    part.mass = np.ones((n, 1)) if mass is None else np.array(mass, dtype=float).reshape(n, 1)
    part.imass = 1.0 / part.mass
    part.ptype = np.zeros(n, dtype=int) if ptype is None else np.array(ptype, dtype=int)
    part.name = np.array(["Ar"] * n)
    part.npart = n
    return part


def lj_fluid(repeat, density=0.8, lattice="fcc", jitter=0.05, seed=3, unwrap=False):
    """Jittered lattice + box of BASELINE.md section 3."""
    pos, size = generate_lattice(lattice, repeat, density=density)
    rng = np.random.default_rng(seed)
    pos = pos + jitter * (2.0 * rng.random(pos.shape) - 1.0)
    if unwrap:  # legal input: integrators never wrap (integrators.py:90)
        pos = pos + rng.integers(-3, 4, size=pos.shape) * (size[:, 1] - size[:, 0])
    return pos, size


def save(name, **arrays):
    np.savez_compressed(HERE / name, **arrays)
    print("wrote", name, {k: np.shape(v) for k, v in arrays.items()})


def lj_case(name, pos, size, params, ptype=None, periodic=None, shift=True, mixing="geometric"):
    box = Box(low=size[:, 0], high=size[:, 1], periodic=periodic)
    part = bulk_particles(pos, ptype=ptype)
    pot = LennardJonesCut(dim=pos.shape[1], shift=shift, mixing=mixing)
    pot.set_parameters(params)
    system = System(box=box, particles=part, potentials=[pot])
    v_pf, f_pf, w_pf = pot.potential_and_force(system)
    f_f, w_f = pot.force(system)
    v_p = pot.potential(system)
    types = sorted({k[0] for k in pot.params if isinstance(k[0], (int, np.integer))})
    table = np.array(
        [
            [[d[a, b] for b in types] for a in types]
            for d in (pot._lj1, pot._lj2, pot._lj3, pot._lj4, pot._offset, pot._rcut2)
        ]
    )
    save(
        name,
        pos=part.pos, ptype=part.ptype, low=box.low, high=box.high,
        periodic=np.array(box.periodic), types=np.array(types), table=table,
        shift=np.array(shift), v_pf=v_pf, f_pf=f_pf, w_pf=w_pf, f_f=f_f, w_f=w_f, v_p=v_p,
    )
    return system, pot


def traj_record(system, frames):
    p = system.particles
    frames["pos"].append(p.pos.copy())
    frames["vel"].append(p.vel.copy())
    frames["force"].append(p.force.copy())
    frames["virial"].append(p.virial.copy())
    frames["v_pot"].append(np.nan if p.v_pot is None else float(p.v_pot))


def run_traj(name, system, integrator, steps, extra=None, first_eval=True):
    """Drive through MDSimulation like the reference's end-to-end test (test_mdsimulation.py:122-141)."""
    frames = {k: [] for k in ("pos", "vel", "force", "virial", "v_pot")}
    p = system.particles
    init = dict(pos0=p.pos.copy(), vel0=p.vel.copy(), mass=p.mass.copy(), ptype=p.ptype.copy())
    if first_eval:
        sim = MDSimulation(system, integrator, steps)
        for sysi in sim.run():
            traj_record(sysi, frames)
    else:  # bare integrator calls, as tests/test_integrators.py does
        traj_record(system, frames)
        for _ in range(steps):
            integrator(system)
            traj_record(system, frames)
    out = {k: np.array(v) for k, v in frames.items()}
    out.update(init)
    if extra:
        out.update(extra)
    save(name, **out)


def main():
    single = {0: {"sigma": 1.0, "epsilon": 1.0, "rcut": 2.5}}

    # ---- LJ force/energy/virial, jittered FCC, one type -------------------
    for rep in (3, 4, 6, 10):
        pos, size = lj_fluid([rep] * 3)
        lj_case(f"lj_fcc_{len(pos)}.npz", pos, size, single)

    # unwrapped coordinates (several box lengths away) are legal input
    pos, size = lj_fluid([4] * 3, unwrap=True)
    lj_case("lj_fcc_256_unwrapped.npz", pos, size, single)

    # L < 2*rc: "nearest image only" regime of the 108-atom fixture
    pos, size = lj_fluid([3] * 3, density=0.9, seed=11)
    lj_case("lj_fcc_108_smallbox.npz", pos, size, single)

    # two types, mixing rules, an explicit pair entry, no shift
    pos, size = lj_fluid([4] * 3, seed=5)
    rng = np.random.default_rng(17)
    ptype = rng.integers(0, 2, size=len(pos))
    two = {
        0: {"sigma": 1.0, "epsilon": 1.0, "rcut": 2.5},
        1: {"sigma": 1.2, "epsilon": 0.7, "rcut": 3.0},
    }
    lj_case("lj_two_types_geometric.npz", pos, size, two, ptype=ptype)
    lj_case("lj_two_types_arithmetic_noshift.npz", pos, size, two, ptype=ptype,
            shift=False, mixing="arithmetic")
    three = dict(two)
    three[2] = {"sigma": 0.9, "epsilon": 1.3, "rcut": 2.2}
    three[(0, 2)] = {"sigma": 1.1, "epsilon": 0.5, "rcut": 2.0}
    ptype3 = rng.integers(0, 3, size=len(pos))
    lj_case("lj_three_types_sixthpower_explicit.npz", pos, size, three, ptype=ptype3,
            mixing="sixthpower")

    # 2-D (sq2 lattice) and a slab with one open direction
    pos, size = lj_fluid([12, 12], lattice="sq2", density=0.7, seed=7)
    lj_case("lj_2d_288.npz", pos, size, single)
    pos, size = lj_fluid([4] * 3, seed=9)
    lj_case("lj_slab_256.npz", pos, size, single, periodic=[True, False, True])
    # WCA cut-off (examples/movie-doublewell-pair.py:41-44)
    wca = {0: {"sigma": 1.0, "epsilon": 1.0, "rcut": 2.0 ** (1.0 / 6.0)}}
    pos, size = lj_fluid([4] * 3, seed=13)
    lj_case("lj_wca_256.npz", pos, size, wca)

    # ---- DoubleWell (well.py:65-81) ----------------------------------------
    rng = np.random.default_rng(23)
    for dim, n in ((1, 11), (3, 257)):
        pos = 3.0 * (rng.random((n, dim)) - 0.5)
        box = Box(periodic=[False] * dim)
        part = bulk_particles(pos)
        pot = DoubleWell(a=1.0, b=2.0, c=0.3)
        system = System(box=box, particles=part, potentials=[pot])
        f, w = pot.force(system)
        save(f"doublewell_{dim}d_{n}.npz", pos=pos, abc=np.array([1.0, 2.0, 0.3]),
             v=pot.potential(system), f=f, w=w)

    # ---- DoubleWellPair (well.py:230-286) ------------------------------------
    pos, size = lj_fluid([3] * 3, seed=29)
    ptype = np.zeros(len(pos), dtype=int)
    ptype[[5, 17, 40, 77]] = 1
    ptype[[8, 60]] = 2
    for types in ((1, 1), (1, 2), (0, 2)):
        box = Box(low=size[:, 0], high=size[:, 1])
        part = bulk_particles(pos, ptype=ptype)
        pot = DoubleWellPair(types=types)
        pot.set_parameters({"rzero": 2.0 ** (1.0 / 6.0), "height": 6.0, "width": 0.25})
        system = System(box=box, particles=part, potentials=[pot])
        f, w = pot.force(system)
        save(f"dwpair_108_types{types[0]}{types[1]}.npz", pos=pos, ptype=ptype,
             low=box.low, high=box.high, periodic=np.array(box.periodic),
             types=np.array(types), rzero=2.0 ** (1.0 / 6.0), height=6.0, width=0.25,
             v=pot.potential(system), f=f, w=w)

    # ---- trajectories ----------------------------------------------------------
    # VelocityVerlet, LJ, 10 steps through MDSimulation (config #1 in miniature)
    pos, size = lj_fluid([3] * 3, density=0.9, seed=31)
    part = bulk_particles(pos)
    generate_maxwell_velocities(part, rgen=np.random.default_rng(1), temperature=0.8)
    pot = LennardJonesCut(dim=3, shift=True)
    pot.set_parameters(single)
    box = Box(low=size[:, 0], high=size[:, 1])
    system = System(box=box, particles=part, potentials=[pot])
    run_traj("traj_vv_lj_108.npz", system, VelocityVerlet(0.002), 10,
             extra=dict(low=box.low, high=box.high, dt=0.002))

    pos, size = lj_fluid([4] * 3, seed=37)
    rng = np.random.default_rng(41)
    mass = 0.5 + rng.random(len(pos)) * 2.0
    part = bulk_particles(pos, mass=mass)
    generate_maxwell_velocities(part, rgen=np.random.default_rng(2), temperature=0.8)
    pot = LennardJonesCut(dim=3, shift=True)
    pot.set_parameters(single)
    box = Box(low=size[:, 0], high=size[:, 1])
    system = System(box=box, particles=part, potentials=[pot])
    run_traj("traj_vv_lj_256_masses.npz", system, VelocityVerlet(0.005), 10,
             extra=dict(low=box.low, high=box.high, dt=0.005))

    # Verlet, same kind of system
    pos, size = lj_fluid([3] * 3, seed=43)
    part = bulk_particles(pos)
    generate_maxwell_velocities(part, rgen=np.random.default_rng(3), temperature=0.8)
    pot = LennardJonesCut(dim=3, shift=True)
    pot.set_parameters(single)
    box = Box(low=size[:, 0], high=size[:, 1])
    system = System(box=box, particles=part, potentials=[pot])
    run_traj("traj_verlet_lj_108.npz", system, Verlet(0.002), 10,
             extra=dict(low=box.low, high=box.high, dt=0.002))

    # dimer in WCA solvent: LJ + DoubleWellPair, VelocityVerlet (config #5 in miniature)
    pos, size = lj_fluid([3] * 3, density=0.7, seed=47)
    ptype = np.zeros(len(pos), dtype=int)
    ptype[[0, 1]] = 1
    part = bulk_particles(pos, ptype=ptype)
    generate_maxwell_velocities(part, rgen=np.random.default_rng(4), temperature=0.8)
    rc = 2.0 ** (1.0 / 6.0)
    lj = LennardJonesCut(dim=3, shift=True)
    lj.set_parameters({0: {"sigma": 1.0, "epsilon": 1.0, "rcut": rc},
                       1: {"sigma": 1.0, "epsilon": 1.0, "rcut": rc}})
    dw = DoubleWellPair(types=(1, 1))
    dw.set_parameters({"rzero": rc, "height": 6.0, "width": 0.25})
    box = Box(low=size[:, 0], high=size[:, 1])
    system = System(box=box, particles=part, potentials=[lj, dw])
    run_traj("traj_vv_dimer_108.npz", system, VelocityVerlet(0.002), 10,
             extra=dict(low=box.low, high=box.high, dt=0.002, rc=rc))

    # DoubleWell replicas, Langevin integrators driven by the Philox twin (config #3 in miniature)
    n = 64
    for dim in (1, 2):
        rng = np.random.default_rng(53 + dim)
        pos0 = -1.0 + 0.1 * rng.standard_normal((n, dim))
        vel0 = 0.3 * rng.standard_normal((n, dim))
        mass = np.where(np.arange(n) % 2 == 0, 1.0, 2.5)
        for tag, make in (
            ("inertia", lambda rg: LangevinInertia(timestep=0.002, gamma=0.3, beta=1.0 / 0.07, rgen=rg)),
            ("overdamped", lambda rg: LangevinOverdamped(timestep=0.002, gamma=0.3, rgen=rg, beta=1.0 / 0.07)),
        ):
            part = bulk_particles(pos0, vel=vel0, mass=mass)
            box = Box(periodic=[False] * dim)
            system = System(box=box, particles=part, potentials=[DoubleWell(a=1.0, b=2.0, c=0.0)])
            system.potential_and_force()
            integ = make(OraclePhiloxNormal(key=2024))
            run_traj(f"traj_langevin_{tag}_{dim}d_{n}.npz", system, integ, 10, first_eval=False,
                     extra=dict(dt=0.002, gamma=0.3, beta=1.0 / 0.07, key=2024, abc=np.array([1.0, 2.0, 0.0])))

    # LangevinInertia on an LJ fluid (config #4 in miniature)
    pos, size = lj_fluid([3] * 3, seed=59)
    part = bulk_particles(pos)
    generate_maxwell_velocities(part, rgen=np.random.default_rng(5), temperature=0.8)
    pot = LennardJonesCut(dim=3, shift=True)
    pot.set_parameters(single)
    box = Box(low=size[:, 0], high=size[:, 1])
    system = System(box=box, particles=part, potentials=[pot])
    system.potential_and_force()
    integ = LangevinInertia(timestep=0.005, gamma=0.3, beta=1.0 / 0.8, rgen=OraclePhiloxNormal(key=7))
    run_traj("traj_langevin_inertia_lj_108.npz", system, integ, 10, first_eval=False,
             extra=dict(low=box.low, high=box.high, dt=0.005, gamma=0.3, beta=1.0 / 0.8, key=7))

    # Langevin constants (integrators.py:223-263) for a spread of masses
    imass = 1.0 / np.array([[1.0], [2.5], [0.3], [39.948]])
    sysd = System(box=Box(periodic=[False]), particles=bulk_particles(np.zeros((4, 1)), mass=1.0 / imass))
    consts = {}
    for tag, (dt, gamma, beta) in {"a": (0.002, 0.3, 1.0), "b": (0.0005, 0.5, 1.0), "c": (0.005, 0.3, 1.25)}.items():
        integ = LangevinInertia(timestep=dt, gamma=gamma, beta=beta)
        integ.initiate_parameters(sysd)
        p = integ.param
        consts.update({
            f"{tag}_in": np.array([dt, gamma, beta]), f"{tag}_c0": p.c0, f"{tag}_a1": p.a1,
            f"{tag}_a2": p.a2, f"{tag}_b1": p.b1, f"{tag}_b2": p.b2,
            f"{tag}_cov": np.array(p.cov), f"{tag}_cho": np.array(p.cho),
        })
    save("langevin_constants.npz", imass=imass, **consts)

    # ---- the reference's own known answers for this path ---------------------
    import gzip

    sys.path.insert(0, str(REF / "tests"))
    import importlib.util

    def load(path, name):
        spec = importlib.util.spec_from_file_location(name, path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        return mod

    t_lj = load(REF / "tests/potentials/test_lennardjones.py", "ref_test_lj")
    t_ww = load(REF / "tests/potentials/test_wellwca.py", "ref_test_wellwca")
    t_in = load(REF / "tests/test_integrators.py", "ref_test_integrators")
    t_md = load(REF / "tests/simulation/test_mdsimulation.py", "ref_test_md")
    with gzip.open(REF / "tests/simulation/velocities.txt.gz") as fh:
        md_vel = np.loadtxt(fh)
    fake = t_in.FakeRandomGenerator(seed=1)
    save(
        "reference_kats.npz",
        # tests/potentials/test_lennardjones.py:16-26
        lj_vpot=t_lj.CORRECT_VPOT, lj_force=t_lj.CORRECT_FORCE, lj_virial=t_lj.CORRECT_VIRIAL,
        # tests/potentials/test_wellwca.py:12-14
        dwp_force=t_ww.CORRECT_FORCE, dwp_virial=t_ww.CORRECT_VIRIAL, dwp_vpot=t_ww.CORRECT_VPOT,
        # tests/test_integrators.py:19-302 (series from the external code "TISMOL")
        vv_pos=np.array(t_in.TISMOL_POS_VV), vv_vel=np.array(t_in.TISMOL_VEL_VV),
        verlet_pos=np.array(t_in.CORRECT_POS_V), verlet_vel=np.array(t_in.CORRECT_VEL_V),
        over_pos=np.array(t_in.TISMOL_POS_LANG_OVER), over_vel=np.array(t_in.TISMOL_VEL_LANG_OVER),
        inertia_pos=np.array(t_in.TISMOL_POS_LANG), inertia_vel=np.array(t_in.TISMOL_VEL_LANG),
        # tests/test_integrators.py:356-380: the 20-number table of FakeRandomGenerator
        fake_table=np.array(fake.rgen),
        # tests/simulation/md-traj.npy, velocities.txt.gz, test_mdsimulation.py:19-95
        md_traj=np.load(REF / "tests/simulation/md-traj.npy"), md_vel=md_vel,
        md_temperature=np.array(t_md.CORRECT_ENERGIES["temperature"]),
        md_pressure=np.array(t_md.CORRECT_ENERGIES["pressure"]),
        md_ekin=np.array(t_md.CORRECT_ENERGIES["ekin"]),
        md_vpot=np.array(t_md.CORRECT_ENERGIES["vpot"]),
    )
    # Langevin parameter KATs (tests/test_integrators.py:305-353)
    for tag, par in (("1", t_in.LANGEVIN_PARAMETER1), ("2", t_in.LANGEVIN_PARAMETER2)):
        save(f"reference_langevin_parameter{tag}.npz", c0=par.c0, a1=par.a1, a2=par.a2, b1=par.b1,
             b2=par.b2, cov=np.array(par.cov), cho=np.array(par.cho))

    # Box KATs (tests/system/test_box.py:146-178)
    box = Box(low=[0, 0, 0], high=[10, 10, 10], periodic=[True, False, True])
    rng = np.random.default_rng(61)
    dist = 30.0 * (rng.random((64, 3)) - 0.5)
    save("box_pbc.npz", low=box.low, high=box.high, periodic=np.array(box.periodic), dist=dist,
         dist_matrix=box.pbc_dist_matrix(dist), dist_each=np.array([box.pbc_dist(d) for d in dist]),
         wrapped=box.pbc_wrap(dist))


if __name__ == "__main__":
    main()
