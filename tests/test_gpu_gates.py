"""Parity gates the north star names beyond the fixture trajectories ("missing parity gates"):

* Langevin ensembles "matching in distribution": the in-kernel Philox stream against a sample the
  reference produced under its OWN default generator (``default_rng(seed)``: PCG64 + ziggurat,
  ``turtlemd/integrators.py:216,284-297``) -- two-sample Kolmogorov-Smirnov on x and v at a fixed time,
  plus <v^2> = 1/(beta m) to three standard errors; kinetic temperature of an LJ fluid under the thermostat.
* benchmark-size parity of BASELINE.json configs[4] (dimer in LJ solvent, N = 256 000: neighbour list
  one-type kernel + ``tmd_dwpair`` accumulating) and of ``lj256k``: sampled rows against the oracle plus the
  oracle's DoubleWellPair on the active subset (SURVEY.md 8c iv).
* config #5 in miniature on the neighbour-list path against a reference-generated trajectory.
* C-ABI entry points without a Python caller: ``tmd_host_vv_lj_step`` and the ``tmd_ipc_*`` round trip.
"""
from __future__ import annotations

import ctypes as C
import math
import os
import subprocess
import sys
import textwrap

import numpy as np
import pytest

from conftest import REPO, load_golden, rel_norm

pytestmark = pytest.mark.gpu

TOL = 1e-10        # forces / energy / virial, relative (north star)
TRAJ_TOL = 1e-8    # 10-step trajectories (north star)
ONE = {"sigma": 1.0, "epsilon": 1.0, "rcut": 2.5}


def _fluid(rep, seed=3):
    from turtlemd_b200.tools import generate_lattice

    pos, size = generate_lattice("fcc", [rep] * 3, density=0.8)
    return pos + 0.05 * (2.0 * np.random.default_rng(seed).random(pos.shape) - 1.0), size


def _ks_two_sample(a, b):
    """Two-sample Kolmogorov-Smirnov statistic D and its asymptotic p-value (Smirnov's series)."""
    a, b = np.sort(np.ravel(a)), np.sort(np.ravel(b))
    both = np.concatenate([a, b])
    cdf_a = np.searchsorted(a, both, side="right") / len(a)
    cdf_b = np.searchsorted(b, both, side="right") / len(b)
    d = float(np.max(np.abs(cdf_a - cdf_b)))
    en = math.sqrt(len(a) * len(b) / (len(a) + len(b)))
    lam = (en + 0.12 + 0.11 / en) * d
    p = 2.0 * sum((-1.0) ** (j - 1) * math.exp(-2.0 * j * j * lam * lam) for j in range(1, 101))
    return d, min(max(p, 0.0), 1.0)


# ---- (i) matching in distribution ---------------------------------------------------------------------


def test_langevin_ensemble_matches_the_reference_default_rng_in_distribution():
    """20 000 reference particles after 500 steps under default_rng(7) (fixture written by
    tests/golden/make_golden_r2.py) vs 100 000 replicas of the same set-up on the in-kernel stream."""
    from turtlemd_b200.integrators import LangevinInertia
    from turtlemd_b200.philox import PhiloxNormal
    from turtlemd_b200.potentials import DoubleWell
    from turtlemd_b200.system import Box, Particles, System

    g = load_golden("langevin_dist_default_rng.npz")
    n, steps = 100_000, int(g["steps"])
    system = System(Box(periodic=[False]), Particles.from_arrays(np.full((n, 1), float(g["x0"]))),
                    [DoubleWell(a=float(g["a"]), b=float(g["b"]), c=float(g["c"]))])
    system.potential_and_force()
    integ = LangevinInertia(timestep=float(g["dt"]), gamma=float(g["gamma"]), beta=float(g["beta"]),
                            rgen=PhiloxNormal(key=2024))
    integ.run(system, steps)
    x, v = system.particles.pos[:, 0], system.particles.vel[:, 0]
    for name, mine, ref in (("x", x, g["pos"]), ("v", v, g["vel"])):
        d, p = _ks_two_sample(mine, ref)
        assert p > 1e-3, (name, d, p)  # a wrong noise amplitude of 2 % gives p < 1e-6 at these sample sizes
    # second moment of v: the two ensembles agree within their standard errors (t = 1 is before equilibrium:
    # <v^2> is still growing towards 1/(beta m), so the two samples are compared with each other, not with it)
    ref_v = g["vel"][:, 0]
    se_v2 = math.sqrt(np.var(v**2) / n + np.var(ref_v**2) / len(ref_v))
    assert abs(np.mean(v**2) - np.mean(ref_v**2)) < 4.0 * se_v2, (np.mean(v**2), np.mean(ref_v**2), se_v2)
    # the first two moments of x agree between the two ensembles within their standard errors
    se_x = math.sqrt(np.var(x) / n + np.var(g["pos"]) / len(g["pos"]))
    assert abs(np.mean(x) - np.mean(g["pos"])) < 4.0 * se_x
    assert abs(np.std(x) - np.std(g["pos"])) < 0.03 * np.std(g["pos"])


def test_stock_langevin_script_retraces_the_reference_under_the_same_seed():
    """``LangevinInertia(timestep, gamma, beta, seed=7)`` exactly as a TurtleMD script writes it: the default generator is
    the reference's ``default_rng(seed)`` (integrators.py:216), continued on the device in the reference's order (:284-315;
    tmd_npstream_*, csrc/npy_normal.cu: bit-exact PCG64 + ziggurat), and the kernels do the rest.  20 000 particles x 500 steps against the reference's own run (fixture of
    tests/golden/make_golden_r2.py): the same trajectory, not merely the same distribution."""
    from turtlemd_b200.integrators import LangevinInertia
    from turtlemd_b200.potentials import DoubleWell
    from turtlemd_b200.system import Box, Particles, System

    g = load_golden("langevin_dist_default_rng.npz")
    n, steps = int(g["n"]), int(g["steps"])
    system = System(Box(periodic=[False]), Particles.from_arrays(np.full((n, 1), float(g["x0"]))),
                    [DoubleWell(a=float(g["a"]), b=float(g["b"]), c=float(g["c"]))])
    system.potential_and_force()
    integ = LangevinInertia(timestep=float(g["dt"]), gamma=float(g["gamma"]), beta=float(g["beta"]), seed=int(g["seed"]))
    for _ in range(steps):
        integ(system)
    assert rel_norm(system.particles.pos, g["pos"]) < 1e-9
    assert rel_norm(system.particles.vel, g["vel"]) < 1e-9


def test_stock_langevin_script_on_an_lj_fluid_retraces_the_reference():
    """BASELINE.json configs[3] in miniature as a stock script writes it: ``LangevinInertia(0.005, 0.3, 1 / 0.8, seed=3)`` over
    a 108-atom LennardJonesCut fluid, every frame of the reference's own 10-step run (positions, velocities, forces, virial,
    potential energy) to the 10-step tolerance of the north star."""
    from turtlemd_b200.integrators import LangevinInertia
    from turtlemd_b200.potentials import LennardJonesCut
    from turtlemd_b200.system import Box, Particles, System

    g = load_golden("traj_langevin_inertia_lj_108_default_rng.npz")
    pot = LennardJonesCut(dim=3, shift=True)
    pot.set_parameters({0: dict(ONE)})
    part = Particles.from_arrays(g["pos0"], vel=g["vel0"], mass=g["mass"], ptype=g["ptype"])
    system = System(Box(low=g["low"], high=g["high"]), part, [pot])
    system.potential_and_force()
    integ = LangevinInertia(timestep=float(g["dt"]), gamma=float(g["gamma"]), beta=float(g["beta"]), seed=int(g["seed"]))
    for frame in range(11):
        p = system.particles
        assert rel_norm(p.pos, g["pos"][frame]) < TRAJ_TOL, frame
        assert rel_norm(p.vel, g["vel"][frame]) < TRAJ_TOL, frame
        assert rel_norm(p.force, g["force"][frame]) < TRAJ_TOL, frame
        assert rel_norm(p.virial, g["virial"][frame]) < TRAJ_TOL, frame
        assert abs(p.v_pot - g["v_pot"][frame]) <= TRAJ_TOL * abs(g["v_pot"][frame]), frame
        integ(system)


def test_langevin_replicas_reach_equipartition():
    """<v^2> = 1/(beta m) once the ensemble has equilibrated (t = 20 = 6 / gamma): 100 000 replicas, three
    standard errors (sqrt(2/n) <v^2> for a Gaussian) plus 0.5 % for the finite time step."""
    from turtlemd_b200.integrators import LangevinInertia
    from turtlemd_b200.philox import PhiloxNormal
    from turtlemd_b200.potentials import DoubleWell
    from turtlemd_b200.system import Box, Particles, System

    n, beta = 100_000, 1.0 / 0.07
    system = System(Box(periodic=[False]), Particles.from_arrays(np.full((n, 1), -1.0)), [DoubleWell(a=1.0, b=2.0, c=0.0)])
    system.potential_and_force()
    integ = LangevinInertia(timestep=0.002, gamma=0.3, beta=beta, rgen=PhiloxNormal(key=99))
    integ.run(system, 10_000)
    v = system.particles.vel[:, 0]
    target = 1.0 / beta
    assert abs(np.mean(v**2) - target) < 3.0 * math.sqrt(2.0 / n) * target + 0.005 * target, (np.mean(v**2), target)
    assert abs(np.mean(v)) < 4.0 * math.sqrt(target / n)


def test_lj_fluid_under_langevin_inertia_reaches_the_bath_temperature():
    """N = 32 000 LJ fluid, gamma = 2: after 400 steps the kinetic temperature is the bath's (0.8) to 2 %
    (standard error of T at 96 000 degrees of freedom: 0.5 %)."""
    from turtlemd_b200.integrators import LangevinInertia
    from turtlemd_b200.philox import PhiloxNormal
    from turtlemd_b200.potentials import LennardJonesCut
    from turtlemd_b200.system import Box, Particles, System

    pos, size = _fluid(20)
    vel = np.zeros_like(pos)  # start cold: the thermostat has to supply all of it
    pot = LennardJonesCut(dim=3, shift=True, method="nlist")
    pot.set_parameters({0: dict(ONE)})
    system = System(Box(low=size[:, 0], high=size[:, 1]), Particles.from_arrays(pos, vel=vel), [pot])
    system.potential_and_force()
    integ = LangevinInertia(timestep=0.005, gamma=2.0, beta=1.0 / 0.8, rgen=PhiloxNormal(key=5))
    integ.run(system, 400)
    temps = []
    for _ in range(5):
        integ.run(system, 40)
        v = system.particles.vel
        temps.append(float(np.sum(v * v) / v.size))
    assert abs(np.mean(temps) - 0.8) < 0.02 * 0.8, temps


# ---- (ii), (iii) benchmark-size parity ---------------------------------------------------------------------


def _sampled_rows_check(system, pos, size, ptype, params, rows):
    from oracle import turtle_oracle as to

    length = size[:, 1] - size[:, 0]
    table, types = to.lj_tables(params)
    tidx = np.searchsorted(types, ptype)
    fr, _ = to.lj_rows(pos, tidx, length, 1.0 / length, [True] * 3, table, rows)
    return fr, length


def test_lj_256k_sampled_rows_against_oracle():
    from turtlemd_b200.potentials import LennardJonesCut
    from turtlemd_b200.system import Box, Particles, System

    pos, size = _fluid(40)
    assert len(pos) == 256_000
    pot = LennardJonesCut(dim=3, shift=True, method="nlist")
    pot.set_parameters({0: dict(ONE)})
    system = System(Box(low=size[:, 0], high=size[:, 1]), Particles.from_arrays(pos), [pot])
    v, f, w = system.potential_and_force()
    rows = np.random.default_rng(11).choice(len(pos), 64, replace=False)
    fr, _ = _sampled_rows_check(system, pos, size, np.zeros(len(pos), dtype=int), {0: dict(ONE)}, rows)
    assert rel_norm(f[rows], fr) < TOL
    assert np.max(np.abs(f.sum(axis=0))) < 1e-8 * np.max(np.abs(f))
    assert rel_norm(w, w.T) < 1e-12


def test_dimer_256k_benchmark_configuration_against_oracle():
    """bench.py's dimer256k system: LJ through the neighbour list's one-type kernel (two types, identical
    coefficients) and DoubleWellPair accumulating into the same force / virial / energy buffers."""
    from oracle import turtle_oracle as to
    from turtlemd_b200.potentials import DoubleWellPair, LennardJonesCut
    from turtlemd_b200.system import Box, Particles, System

    pos, size = _fluid(40)
    n = len(pos)
    ptype = np.zeros(n, dtype=int)
    first = n // 2
    near = int(np.argsort(np.sum((pos - pos[first]) ** 2, axis=1))[1])
    ptype[[first, near]] = 1
    params = {0: dict(ONE), 1: dict(ONE)}
    lj = LennardJonesCut(dim=3, shift=True, method="nlist")
    lj.set_parameters(params)
    bond = DoubleWellPair(types=(1, 1), dim=3)
    well = {"rzero": 2.0 ** (1.0 / 6.0), "height": 6.0, "width": 0.25}
    bond.set_parameters(well)
    box = Box(low=size[:, 0], high=size[:, 1])
    both = System(box, Particles.from_arrays(pos, ptype=ptype), [lj, bond])
    v, f, w = both.potential_and_force()

    lj_only = System(box, Particles.from_arrays(pos, ptype=ptype), [lj])
    v_lj, f_lj, w_lj = lj_only.potential_and_force()
    length = size[:, 1] - size[:, 0]
    # the oracle's DoubleWellPair on the ACTIVE subset equals the full evaluation (SURVEY.md 8c iv)
    active = np.array([first, near])
    par = to.dwpair_params(well["rzero"], well["width"], well["height"])
    f_dw, w_dw = to.dwpair_force(pos[active], ptype[active], length, 1.0 / length, [True] * 3, (1, 1), par)
    v_dw = to.dwpair_potential(pos[active], ptype[active], length, 1.0 / length, [True] * 3, (1, 1), par)
    expect = f_lj.copy()
    expect[active] += f_dw
    assert rel_norm(f, expect) < 1e-13
    assert abs(v - (v_lj + v_dw)) <= 1e-12 * abs(v_lj)
    assert rel_norm(w, w_lj + w_dw) < 1e-12
    # and the LJ part itself against the oracle on sampled rows, the two dimer atoms included
    rows = np.concatenate([active, np.random.default_rng(13).choice(n, 48, replace=False)])
    fr, _ = _sampled_rows_check(lj_only, pos, size, ptype, params, rows)
    assert rel_norm(f_lj[rows], fr) < TOL
    assert lj.nlist_stats()["calls"] >= 1  # the list path ran (not the all-pairs fallback)


def test_dimer_on_the_list_path_follows_the_reference_trajectory():
    """traj_vv_dimer_864.npz: the reference's own VelocityVerlet + LennardJonesCut + DoubleWellPair run."""
    from turtlemd_b200.integrators import VelocityVerlet
    from turtlemd_b200.potentials import DoubleWellPair, LennardJonesCut
    from turtlemd_b200.system import Box, Particles, System

    g = load_golden("traj_vv_dimer_864.npz")
    lj = LennardJonesCut(dim=3, shift=True, method="nlist")
    lj.set_parameters({0: dict(ONE), 1: dict(ONE)})
    bond = DoubleWellPair(types=(1, 1))
    bond.set_parameters({"rzero": 2.0 ** (1.0 / 6.0), "height": 6.0, "width": 0.25})
    part = Particles.from_arrays(g["pos0"], vel=g["vel0"], mass=g["mass"], ptype=g["ptype"])
    system = System(Box(low=g["low"], high=g["high"]), part, [lj, bond])
    system.potential_and_force()
    integ = VelocityVerlet(float(g["dt"]))
    frames = [int(k) for k in g["frames"]]
    done = 0
    for slot, frame in enumerate(frames):
        if frame > done:
            integ.run(system, frame - done)
            done = frame
        p = system.particles
        assert rel_norm(p.pos, g["pos"][slot]) < TRAJ_TOL, frame
        assert rel_norm(p.vel, g["vel"][slot]) < TRAJ_TOL, frame
        assert rel_norm(p.force, g["force"][slot]) < TRAJ_TOL, frame
        assert abs(p.v_pot - g["v_pot"][slot]) <= TRAJ_TOL * abs(g["v_pot"][slot])
        assert rel_norm(p.virial, g["virial"][slot]) < TRAJ_TOL
    assert lj.nlist_stats()["calls"] >= 1


# ---- (iv) C-ABI entry points without a Python caller -----------------------------------------------------


def test_host_vv_lj_step_follows_the_reference_trajectory():
    """tmd_host_vv_lj_step = what a maintainer binds behind VelocityVerlet.integration_step:
    host arrays in and out, ten calls against the reference-generated traj_vv_lj_108.npz."""
    from oracle import turtle_oracle as to
    from turtlemd_b200 import _lib

    g = load_golden("traj_vv_lj_108.npz")
    pos, vel = g["pos0"].copy(), g["vel0"].copy()
    force = g["force"][0].copy()
    n = len(pos)
    imass = np.ascontiguousarray(1.0 / g["mass"][:, 0])
    table, _ = to.lj_tables({0: dict(ONE)})
    table = np.ascontiguousarray(table, dtype=np.float64)
    tidx = np.zeros(n, dtype=np.int64)
    length = g["high"] - g["low"]
    box = _lib.Box()
    box.dim = 3
    for k in range(3):
        box.periodic[k], box.length[k], box.ilength[k] = 1, length[k], 1.0 / length[k]
    vw = np.zeros(10)
    for frame in range(1, 11):
        _lib.call("tmd_host_vv_lj_step", _lib.hptr(pos), _lib.hptr(vel), _lib.hptr(force), _lib.hptr(imass),
                  _lib.hptr(tidx), n, C.byref(box), _lib.hptr(table), 1, 0, float(g["dt"]), _lib.hptr(vw))
        assert rel_norm(pos, g["pos"][frame]) < TRAJ_TOL, frame
        assert rel_norm(vel, g["vel"][frame]) < TRAJ_TOL, frame
        assert rel_norm(force, g["force"][frame]) < TRAJ_TOL, frame
        assert abs(vw[0] - g["v_pot"][frame]) <= TRAJ_TOL * abs(g["v_pot"][frame])
        assert rel_norm(vw[1:].reshape(3, 3), g["virial"][frame]) < TRAJ_TOL


_IPC_CHILD = textwrap.dedent("""
    import ctypes as C, sys
    sys.path.insert(0, sys.argv[1])
    from turtlemd_b200 import _lib
    handle = (C.c_ubyte * 64).from_buffer_copy(bytes.fromhex(sys.argv[2]))
    _lib.call("tmd_set_device", 0)
    ptr = C.c_void_p()
    _lib.call("tmd_ipc_open_handle", handle, C.byref(ptr))
    import numpy as np
    got = np.zeros(256)
    _lib.call("tmd_memcpy_d2h_async", _lib.hptr(got), ptr, got.nbytes, None)
    _lib.call("tmd_stream_synchronize", None)
    assert np.array_equal(got, np.arange(256.0)), got[:4]
    back = np.arange(256.0) * 2.0 + 1.0
    _lib.call("tmd_memcpy_h2d_async", ptr, _lib.hptr(back), back.nbytes, None)
    _lib.call("tmd_stream_synchronize", None)
    _lib.call("tmd_ipc_close_handle", ptr)
    print("IPC_CHILD ok")
""")


def test_ipc_handle_round_trip_between_two_processes():
    """tmd_ipc_get_handle / open / close: another process maps this process's device buffer, reads what is
    there and writes an answer (the plumbing of the peer-store halo exchange; one GPU is enough)."""
    from turtlemd_b200 import _lib

    ptr = C.c_void_p()
    _lib.call("tmd_malloc", C.byref(ptr), 256 * 8)
    try:
        data = np.arange(256.0)
        _lib.call("tmd_memcpy_h2d_async", ptr, _lib.hptr(data), data.nbytes, None)
        _lib.call("tmd_stream_synchronize", None)
        handle = (C.c_ubyte * 64)()
        _lib.call("tmd_ipc_get_handle", ptr, handle)
        env = dict(os.environ, CUDA_VISIBLE_DEVICES=os.environ.get("CUDA_VISIBLE_DEVICES", "0").split(",")[0])
        run = subprocess.run([sys.executable, "-c", _IPC_CHILD, str(REPO), bytes(handle).hex()],
                             capture_output=True, text=True, timeout=180, env=env)
        assert run.returncode == 0 and "IPC_CHILD ok" in run.stdout, run.stdout + run.stderr
        got = np.zeros(256)
        _lib.call("tmd_memcpy_d2h_async", _lib.hptr(got), ptr, got.nbytes, None)
        _lib.call("tmd_stream_synchronize", None)
        assert np.array_equal(got, data * 2.0 + 1.0)
    finally:
        _lib.call("tmd_free", ptr)
