"""Pin the CPU oracle: reference-generated fixtures + the reference's own known answers.

No GPU, no /root/reference: only ``oracle/`` and ``tests/golden/*.npz``.
Bit-exact (``array_equal``) wherever the restatement keeps the reference's
operation order, which is everywhere except summation over Python-level pair
loops that the oracle visits in the same order anyway.
"""
from __future__ import annotations

import glob

import numpy as np
import pytest

from conftest import GOLDEN, load_golden, rel_norm
from oracle import philox_ref as pr
from oracle import turtle_oracle as to

LJ_CASES = sorted(p.split("/")[-1] for p in glob.glob(str(GOLDEN / "lj_*.npz")))


def _box(g):
    length, ilength = to.box_lengths(g["low"], g["high"], g["periodic"])
    return length, ilength, list(g["periodic"])


def _tidx(g):
    lut = {int(t): n for n, t in enumerate(g["types"])}
    return np.array([lut[int(t)] for t in g["ptype"]])


@pytest.mark.parametrize("name", LJ_CASES)
def test_lj_matches_reference_bitwise(name):
    g = load_golden(name)
    box = _box(g)
    tidx = _tidx(g)
    v, f, w = to.lj_potential_and_force(g["pos"], tidx, *box, g["table"])
    assert v == g["v_pf"]
    assert np.array_equal(f, g["f_pf"])
    assert np.array_equal(w, g["w_pf"])
    if len(g["pos"]) <= 1000:
        f2, w2 = to.lj_force(g["pos"], tidx, *box, g["table"])
        assert np.array_equal(f2, g["f_f"]) and np.array_equal(w2, g["w_f"])
        assert to.lj_potential(g["pos"], tidx, *box, g["table"]) == g["v_p"]


TRI_PARAMS = {0: {"sigma": 1.0, "epsilon": 1.0, "rcut": 2.5}, 1: {"sigma": 1.1, "epsilon": 0.8, "rcut": 2.2}}


@pytest.mark.parametrize("tag", ["3d_150", "3d_90_longcut", "2d_80"])
def test_lj_in_a_triclinic_cell_matches_reference_bitwise(tag):
    """TriclinicBox.pbc_dist (box.py:271-296) under the LJ row loops; tests/golden/make_golden_triclinic.py."""
    g = load_golden(f"triclinic_lj_{tag}.npz")
    alpha, beta, gamma = (None if np.isnan(a) else float(a) for a in g["angles"])
    cell = to.TriclinicCell(g["high"], alpha, beta, gamma)
    assert np.array_equal(cell.matrix, g["box_matrix"]) and cell.half_width == float(g["half_width"])
    params = {k: dict(v, rcut=v["rcut"] * float(g["rscale"])) for k, v in TRI_PARAMS.items()}
    table, types = to.lj_tables(params)
    assert types == [0, 1]
    tidx = np.asarray(g["ptype"])
    v, f, w = to.lj_potential_and_force(g["pos"], tidx, cell, None, None, table)
    assert v == g["v"] and np.array_equal(f, g["f"]) and np.array_equal(w, g["w"])
    assert to.lj_potential(g["pos"], tidx, cell, None, None, table) == g["v_only"]
    f2, w2 = to.lj_force(g["pos"], tidx, cell, None, None, table)
    assert np.array_equal(f2, g["f"]) and np.array_equal(w2, g["w_only"])


def test_lj_tables_match_reference():
    single = {0: {"sigma": 1.0, "epsilon": 1.0, "rcut": 2.5}}
    g = load_golden("lj_fcc_108.npz")
    table, types = to.lj_tables(single)
    assert types == [0] and np.array_equal(table, g["table"])
    two = {0: {"sigma": 1.0, "epsilon": 1.0, "rcut": 2.5}, 1: {"sigma": 1.2, "epsilon": 0.7, "rcut": 3.0}}
    g = load_golden("lj_two_types_geometric.npz")
    assert np.array_equal(to.lj_tables(two)[0], g["table"])
    g = load_golden("lj_two_types_arithmetic_noshift.npz")
    assert np.array_equal(to.lj_tables(two, mixing="arithmetic", shift=False)[0], g["table"])
    three = dict(two)
    three[2] = {"sigma": 0.9, "epsilon": 1.3, "rcut": 2.2}
    three[(0, 2)] = {"sigma": 1.1, "epsilon": 0.5, "rcut": 2.0}
    g = load_golden("lj_three_types_sixthpower_explicit.npz")
    assert np.array_equal(to.lj_tables(three, mixing="sixthpower")[0], g["table"])
    with pytest.raises(ValueError):
        to.mix(1, 1, 1, 1, 1, 1, "nope")


def test_lj_reference_known_answer():
    """tests/potentials/test_lennardjones.py:16-43,194-220: 3 atoms, sigma=eps=1, rc=2.5, shifted."""
    k = load_golden("reference_kats.npz")
    pos = np.array([[1.0, 1.0, 1.0], [1.5, 1.0, 1.0], [1.5, 1.5, 1.0]])  # test_lennardjones.py:29-43
    length, ilength = to.box_lengths([0, 0, 0], [10, 10, 10], [True] * 3)
    table, _ = to.lj_tables({0: {"sigma": 1.0, "epsilon": 1.0, "rcut": 2.5}})
    tidx = np.zeros(3, dtype=int)
    v, f, w = to.lj_potential_and_force(pos, tidx, length, ilength, [True] * 3, table)
    assert v == pytest.approx(float(k["lj_vpot"]))
    assert f == pytest.approx(k["lj_force"])
    assert w == pytest.approx(k["lj_virial"])
    # rcut = 0 must not crash (test_lennardjones.py:186-191)
    t0, _ = to.lj_tables({0: {"sigma": 1.0, "epsilon": 1.0, "rcut": 0.0}})
    assert t0[4, 0, 0] == 0.0


def test_lj_rows_agree_with_full_evaluation():
    g = load_golden("lj_two_types_geometric.npz")
    box, tidx = _box(g), _tidx(g)
    rows = np.array([0, 3, 100, 255])
    frc, vrow = to.lj_rows(g["pos"], tidx, *box, g["table"], rows)
    assert rel_norm(frc, g["f_pf"][rows]) < 1e-13


def test_box_minimum_image():
    g = load_golden("box_pbc.npz")
    length, ilength = to.box_lengths(g["low"], g["high"], g["periodic"])
    per = list(g["periodic"])
    assert np.array_equal(to.pbc_dist_matrix(g["dist"], length, ilength, per), g["dist_matrix"])
    each = np.array([to.pbc_dist(d, length, ilength, per) for d in g["dist"]])
    assert np.array_equal(each, g["dist_each"])
    # defaults: open directions get an infinite length and a zero reciprocal (box.py:142-158)
    length, ilength = to.box_lengths(None, None, [False])
    assert np.isinf(length[0]) and ilength[0] == 0.0


@pytest.mark.parametrize("name", ["doublewell_1d_11.npz", "doublewell_3d_257.npz"])
def test_doublewell(name):
    g = load_golden(name)
    a, b, c = g["abc"]
    assert to.doublewell_potential(g["pos"], a, b, c) == g["v"]
    f, w = to.doublewell_force(g["pos"], a, b, c)
    assert np.array_equal(f, g["f"]) and np.array_equal(w, g["w"])


@pytest.mark.parametrize("types", ["11", "12", "02"])
def test_doublewell_pair(types):
    g = load_golden(f"dwpair_108_types{types}.npz")
    par = to.dwpair_params(float(g["rzero"]), float(g["width"]), float(g["height"]))
    box = _box(g)
    tt = tuple(int(t) for t in g["types"])
    assert to.dwpair_potential(g["pos"], g["ptype"], *box, tt, par) == g["v"]
    f, w = to.dwpair_force(g["pos"], g["ptype"], *box, tt, par)
    assert np.array_equal(f, g["f"]) and np.array_equal(w, g["w"])


@pytest.mark.parametrize("name", ["3d_90_longcut_types11", "3d_90_longcut_types01", "2d_80_types11", "2d_80_types01"])
def test_doublewell_pair_in_a_triclinic_cell(name):
    """well.py:230-286 with TriclinicBox.pbc_dist (box.py:271-296); tests/golden/make_golden_triclinic.py."""
    g = load_golden(f"triclinic_dwpair_{name}.npz")
    alpha, beta, gamma = (None if np.isnan(a) else float(a) for a in g["angles"])
    cell = to.TriclinicCell(g["high"], alpha, beta, gamma)
    par = to.dwpair_params(float(g["rzero"]), float(g["width"]), float(g["height"]))
    tt = tuple(int(t) for t in g["types"])
    assert to.dwpair_potential(g["pos"], g["ptype"], cell, None, None, tt, par) == g["v"]
    f, w = to.dwpair_force(g["pos"], g["ptype"], cell, None, None, tt, par)
    assert np.array_equal(f, g["f"]) and np.array_equal(w, g["w"])


def test_doublewell_pair_reference_known_answer():
    """tests/potentials/test_wellwca.py:12-14,35-45,87-110."""
    k = load_golden("reference_kats.npz")
    # POS2 of test_wellwca.py:22-27; types (0,0,1,1); potential active for (0,0)
    pos = np.array([[1.0, 1.0, 1.0], [1.2, 1.2, 1.2], [1.0, 2.0, 1.0], [1.0, 3.25, 1.0]])
    ptype = np.array([0, 0, 1, 1])
    par = to.dwpair_params(1.0, 0.25, 6.0)
    assert par["width2"] == 0.0625 and par["rwidth"] == 1.25 and par["height4"] == 24.0
    assert to.dwpair_function(1.25, par) == pytest.approx(6.0)
    pos1 = np.array([[1.0, 1.0, 1.0], [2.25, 1.0, 1.0], [1.0, 2.0, 1.0], [1.0, 3.25, 1.0]])  # POS1, :16-21
    l10, il10 = to.box_lengths([0, 0, 0], [10, 10, 10], [True] * 3)
    assert to.dwpair_potential(pos1, np.array([0, 0, 1, 1]), l10, il10, [True] * 3, (0, 0), par) == pytest.approx(6.0)
    length, ilength = to.box_lengths([0, 0, 0], [10, 10, 10], [True] * 3)
    f, w = to.dwpair_force(pos, ptype, length, ilength, [True] * 3, (0, 0), par)
    v = to.dwpair_potential(pos, ptype, length, ilength, [True] * 3, (0, 0), par)
    assert f[0] == pytest.approx(k["dwp_force"]) and f[1] == pytest.approx(-k["dwp_force"])
    assert np.all(f[2:] == 0.0)
    assert w == pytest.approx(k["dwp_virial"])
    assert v == pytest.approx(float(k["dwp_vpot"]))
    assert to.dwpair_activate(0, 1, (0, 1)) and to.dwpair_activate(1, 0, (0, 1))
    assert not to.dwpair_activate(0, 0, (0, 1))


def _ff_lj(g, rc=2.5):
    length, ilength = to.box_lengths(g["low"], g["high"], [True] * 3)
    table, _ = to.lj_tables({0: {"sigma": 1.0, "epsilon": 1.0, "rcut": rc}})
    tidx = np.zeros(len(g["pos0"]), dtype=int)
    return to.ForceField(length, ilength, [True] * 3, [("lj", (tidx, table))])


def _check_traj(st, g, frame, bitwise=True):
    cmp = np.array_equal if bitwise else (lambda a, b: rel_norm(a, b) < 1e-13)
    assert cmp(st.pos, g["pos"][frame]), frame
    assert cmp(st.vel, g["vel"][frame]), frame


def test_velocity_verlet_lj_trajectories():
    for name in ("traj_vv_lj_108.npz", "traj_vv_lj_256_masses.npz"):
        g = load_golden(name)
        ff = _ff_lj(g)
        st = to.State(g["pos0"], g["vel0"], g["mass"])
        to._eval_all(ff, st)  # MDSimulation's step 0 (mdsimulation.py:73-77)
        for frame in range(11):
            _check_traj(st, g, frame)
            assert np.array_equal(st.force, g["force"][frame])
            assert st.v_pot == g["v_pot"][frame]
            assert np.array_equal(st.virial, g["virial"][frame])
            to.velocity_verlet_step(st, ff, float(g["dt"]))


def test_verlet_lj_trajectory():
    g = load_golden("traj_verlet_lj_108.npz")
    ff = _ff_lj(g)
    st = to.State(g["pos0"], g["vel0"], g["mass"])
    to._eval_all(ff, st)
    stepper = to.VerletStepper(float(g["dt"]))
    for frame in range(11):
        _check_traj(st, g, frame)
        stepper.step(st, ff)


def test_dimer_trajectory():
    g = load_golden("traj_vv_dimer_108.npz")
    rc = float(g["rc"])
    length, ilength = to.box_lengths(g["low"], g["high"], [True] * 3)
    table, types = to.lj_tables({0: {"sigma": 1.0, "epsilon": 1.0, "rcut": rc},
                                 1: {"sigma": 1.0, "epsilon": 1.0, "rcut": rc}})
    tidx = np.array([types.index(int(t)) for t in g["ptype"]])
    par = to.dwpair_params(rc, 0.25, 6.0)
    ff = to.ForceField(length, ilength, [True] * 3,
                       [("lj", (tidx, table)), ("dwpair", (g["ptype"], (1, 1), par))])
    st = to.State(g["pos0"], g["vel0"], g["mass"])
    to._eval_all(ff, st)
    for frame in range(11):
        _check_traj(st, g, frame)
        assert st.v_pot == g["v_pot"][frame]
        to.velocity_verlet_step(st, ff, float(g["dt"]))


@pytest.mark.parametrize("dim", [1, 2])
@pytest.mark.parametrize("kind", ["inertia", "overdamped"])
def test_langevin_doublewell_trajectories(kind, dim):
    g = load_golden(f"traj_langevin_{kind}_{dim}d_64.npz")
    a, b, c = g["abc"]
    inf = np.full(dim, np.inf)
    ff = to.ForceField(inf, np.zeros(dim), [False] * dim, [("dw", (a, b, c))])
    st = to.State(g["pos0"], g["vel0"], g["mass"])
    to._eval_all(ff, st)
    rgen = pr.OraclePhiloxNormal(key=int(g["key"]))
    dt, gamma, beta = float(g["dt"]), float(g["gamma"]), float(g["beta"])
    stepper = (to.InertiaStepper(dt, gamma, beta, rgen) if kind == "inertia"
               else to.OverdampedStepper(dt, gamma, rgen, beta))
    for frame in range(11):
        _check_traj(st, g, frame)
        assert np.array_equal(st.force, g["force"][frame])
        stepper.step(st, ff)


def test_langevin_inertia_lj_trajectory():
    g = load_golden("traj_langevin_inertia_lj_108.npz")
    ff = _ff_lj(g)
    st = to.State(g["pos0"], g["vel0"], g["mass"])
    to._eval_all(ff, st)
    stepper = to.InertiaStepper(float(g["dt"]), float(g["gamma"]), float(g["beta"]),
                                pr.OraclePhiloxNormal(key=int(g["key"])))
    for frame in range(11):
        _check_traj(st, g, frame)
        stepper.step(st, ff)


def test_langevin_inertia_lj_trajectory_under_the_reference_default_generator():
    """The reference's own default, default_rng(seed) (integrators.py:216), fed to the oracle's stepper: the fixture the
    reference wrote with a stock LangevinInertia(seed=3) script (tests/golden/make_golden_r2.py)."""
    g = load_golden("traj_langevin_inertia_lj_108_default_rng.npz")
    ff = _ff_lj(g)
    st = to.State(g["pos0"], g["vel0"], g["mass"])
    to._eval_all(ff, st)
    stepper = to.InertiaStepper(float(g["dt"]), float(g["gamma"]), float(g["beta"]), np.random.default_rng(int(g["seed"])))
    for frame in range(11):
        _check_traj(st, g, frame)
        stepper.step(st, ff)


def test_langevin_constants():
    g = load_golden("langevin_constants.npz")
    for tag in "abc":
        dt, gamma, beta = g[f"{tag}_in"]
        par = to.langevin_constants(dt, gamma, beta, g["imass"])
        assert par["c0"] == g[f"{tag}_c0"] and par["a1"] == g[f"{tag}_a1"]
        for key in ("a2", "b1", "b2"):
            assert np.array_equal(par[key], g[f"{tag}_{key}"])
        assert np.array_equal(np.array(par["cov"]), g[f"{tag}_cov"])
        assert np.array_equal(np.array(par["cho"]), g[f"{tag}_cho"])
    # the reference's own parameter KATs (tests/test_integrators.py:305-353, 504-521)
    for tag, (dt, gamma) in (("1", (0.002, 0.3)), ("2", (0.0005, 0.5))):
        k = load_golden(f"reference_langevin_parameter{tag}.npz")
        par = to.langevin_constants(dt, gamma, 1.0, np.ones((1, 1)))
        assert par["c0"] == pytest.approx(float(k["c0"])) and par["a1"] == pytest.approx(float(k["a1"]))
        for key in ("a2", "b1", "b2"):
            assert par[key] == pytest.approx(k[key])
        assert np.array(par["cov"]) == pytest.approx(k["cov"])
        assert np.array(par["cho"]) == pytest.approx(k["cho"])


class _FakeRandom:
    """Cycles the 20-number table of tests/test_integrators.py:356-407 and ignores ``scale``."""

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


def _one_particle():
    ff = to.ForceField(np.array([np.inf]), np.zeros(1), [False], [("dw", (1.0, 2.0, 0.0))])
    st = to.State([[-1.0]], [[0.78008020]], [[1.0]])
    return ff, st


def test_reference_integrator_series():
    """tests/test_integrators.py:453-501: one particle in DoubleWell(1,2,0), dt=0.002."""
    k = load_golden("reference_kats.npz")
    ff, st = _one_particle()
    for i in range(21):
        assert st.pos[0, 0] == pytest.approx(k["vv_pos"][i]) and st.vel[0, 0] == pytest.approx(k["vv_vel"][i])
        to.velocity_verlet_step(st, ff, 0.002)
    ff, st = _one_particle()
    stepper = to.VerletStepper(0.002)
    for i in range(5):
        stepper.step(st, ff)
        assert st.pos[0, 0] == pytest.approx(k["verlet_pos"][i]) and st.vel[0, 0] == pytest.approx(k["verlet_vel"][i])
    ff, st = _one_particle()
    stepper = to.OverdampedStepper(0.002, 0.3, _FakeRandom(k["fake_table"], 1), 1.0)
    for i in range(51):
        assert st.pos[0, 0] == pytest.approx(k["over_pos"][i]) and st.vel[0, 0] == pytest.approx(k["over_vel"][i])
        stepper.step(st, ff)


def test_reference_langevin_inertia_series():
    """tests/test_integrators.py:410-417,542-562: draws are 0.01*(fake numbers), no Cholesky."""
    k = load_golden("reference_kats.npz")
    ff, st = _one_particle()
    stepper = to.InertiaStepper(0.002, 0.3, 1.0, _FakeRandom(k["fake_table"], 1))

    def fake_draw(state):
        norm = stepper.rgen.normal(size=2).reshape(1, 2)
        xv = 0.01 * norm
        return xv[:, :1].copy(), xv[:, 1:].copy()

    stepper.draw = fake_draw
    for i in range(51):
        assert st.pos[0, 0] == pytest.approx(k["inertia_pos"][i]) and st.vel[0, 0] == pytest.approx(k["inertia_vel"][i])
        stepper.step(st, ff)


def test_reference_md_trajectory():
    """tests/simulation/test_mdsimulation.py:98-141 with md-traj.npy and velocities.txt.gz."""
    k = load_golden("reference_kats.npz")
    # generate_lattice("fcc", [3,3,3], density=0.9) (tools/tools.py:37-78)
    cell = np.array([[0.0, 0.0, 0.0], [0.5, 0.5, 0.0], [0.0, 0.5, 0.5], [0.5, 0.0, 0.5]])
    lcon = (4 / 0.9) ** (1.0 / 3.0)
    pos = np.array([lcon * (np.array(i) + c) for i in np.ndindex(3, 3, 3) for c in cell])
    assert pos == pytest.approx(k["md_traj"][0][:, :3])
    high = np.array([3 * lcon] * 3)
    length, ilength = to.box_lengths(np.zeros(3), high, [True] * 3)
    table, _ = to.lj_tables({0: {"sigma": 1.0, "epsilon": 1.0, "rcut": 2.5}})
    ff = to.ForceField(length, ilength, [True] * 3, [("lj", (np.zeros(108, dtype=int), table))])
    st = to.State(pos, k["md_vel"], np.ones(108))
    to._eval_all(ff, st)
    vol = to.box_volume(length)
    for i in range(11):
        th = to.thermo(st, vol, np.ones(3))
        assert th["temperature"] == pytest.approx(k["md_temperature"][i])
        assert th["pressure"] == pytest.approx(k["md_pressure"][i])
        assert th["ekin"] == pytest.approx(k["md_ekin"][i])
        assert th["vpot"] == pytest.approx(k["md_vpot"][i])
        assert st.pos == pytest.approx(k["md_traj"][i][:, :3])
        assert st.vel == pytest.approx(k["md_traj"][i][:, 3:])
        to.velocity_verlet_step(st, ff, 0.002)


# ---- RNG ------------------------------------------------------------------


def test_philox_matches_numpy_bit_generator():
    for key in (0, 1, 12345, 2**63 + 5):
        raw = np.random.Philox(key=key).random_raw(64)
        assert np.array_equal(pr.raw_words(key, 0, 64), raw)
        assert np.array_equal(pr.raw_words(key, 13, 7), raw[13:20])
        assert [int(x) for x in raw[:4]] == pr.philox4x64_10_scalar([1, 0, 0, 0], [key, 0])
    bitgen = np.random.Philox(key=99)
    bitgen.advance(1000)  # advance() adds to the block counter
    assert np.array_equal(pr.raw_words(99, 4000, 8), bitgen.random_raw(8))
    uni = np.random.Generator(np.random.Philox(key=5)).random(16)
    mine = (pr.raw_words(5, 0, 16) >> np.uint64(11)).astype(np.float64) * 2.0**-53
    assert np.array_equal(uni, mine)


def test_fixed_arithmetic_normals():
    rng = np.random.default_rng(0)
    u = (rng.integers(0, 2**53, 20000, dtype=np.uint64) + np.uint64(1)).astype(np.float64) * 2.0**-53
    u = np.concatenate([u, [2.0**-53, 1.0, 0.5, np.sqrt(0.5)]])
    lg = pr.fixed_log(u)
    ref = np.log(u)
    assert np.max(np.abs(lg - ref) / np.maximum(np.abs(ref), 1e-300)) < 4e-16 and lg[-3] == 0.0
    v = np.concatenate([rng.random(20000), np.arange(8) / 8.0, [1 - 2.0**-53]])
    s, c = pr.fixed_sincos_2pi(v)
    assert np.max(np.abs(s - np.sin(2 * np.pi * v))) < 1e-15
    assert np.max(np.abs(c - np.cos(2 * np.pi * v))) < 1e-15
    z = pr.normals(42, 0, 400000)
    assert abs(z.mean()) < 6e-3 and abs(z.std() - 1) < 5e-3
    assert abs((z**3).mean()) < 3e-2 and abs((z**4).mean() - 3) < 6e-2
    assert np.array_equal(pr.normals(42, 1001, 33), z[1001:1034])  # random access
    gen = pr.OraclePhiloxNormal(key=42)
    a = gen.normal(0.0, 1.0, size=(5, 2))
    b = gen.normal(0.0, np.full((3, 1), 2.0), size=(3, 3))
    assert np.array_equal(a.reshape(-1), z[:10]) and np.array_equal(b.reshape(-1), 2.0 * z[10:19])
