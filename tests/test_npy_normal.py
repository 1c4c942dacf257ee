"""NumPy's normal stream (PCG64 + ziggurat) restated for host and device: csrc/npy_normal.h, csrc/npy_normal.cu.

The checker is NumPy itself -- the third-party dependency the reference draws its noise from
(turtlemd/integrators.py:156-160, :216, :312).  CPU tests: the host twin of the library, the log1p restatement against
libm (in both of glibc's builds), and the DEVICE passes -- the kernels of npy_normal.cu compiled for the host with
tests/emul/cuda_emul.h and run as threads.  GPU tests: the same stream on the device, and the Langevin integrators fed by
it against their host-draw path.
"""
from __future__ import annotations

import ctypes as C
import os
import pathlib
import shutil
import subprocess
import sys

import numpy as np
import pytest

REPO = pathlib.Path(__file__).resolve().parent.parent
MASK = (1 << 64) - 1


def state_words(rgen) -> np.ndarray:
    st = rgen.bit_generator.state["state"]
    return np.array([st["state"] >> 64, st["state"] & MASK, st["inc"] >> 64, st["inc"] & MASK], dtype=np.uint64)


def same_bits(a, b) -> bool:
    return np.array_equal(np.asarray(a).view(np.uint64), np.asarray(b).view(np.uint64))


# ---------------------------------------------------------------------------------------------- host twin (library)
def _lib():
    from turtlemd_b200 import _lib as lib

    lib.load()
    return lib


@pytest.mark.parametrize("seed", [0, 3, 12345])
def test_host_twin_is_numpy_bit_for_bit(seed):
    lib = _lib()
    rgen = np.random.default_rng(seed)
    words = state_words(rgen)
    n = 20_000_000
    out = np.empty(n)
    used = C.c_uint64(0)
    lib.call("tmd_host_npstream_normals", lib.hptr(words), lib.hptr(out), n, C.byref(used))
    ref = rgen.standard_normal(n)
    assert same_bits(out, ref)
    assert np.array_equal(words, state_words(rgen))  # the generator state after the draws
    assert 1.0215 < used.value / n < 1.0226  # 64-bit words per normal: what the device window is sized on


def test_host_normals_helper_moves_the_generator():
    from turtlemd_b200 import npstream

    a, b = np.random.default_rng(9), np.random.default_rng(9)
    first = npstream.host_normals(a, 1000)
    assert same_bits(first, b.standard_normal(1000))
    assert same_bits(a.standard_normal(7), b.standard_normal(7))  # both generators stand at the same place
    # normal(loc, scale, size) is loc + scale * standard_normal: what the integrators rely on
    c, d = np.random.default_rng(4), np.random.default_rng(4)
    sigma = np.array([[0.3], [1.7]])
    assert same_bits(c.normal(loc=0.0, scale=sigma, size=(2, 3)), sigma * d.standard_normal((2, 3)))
    # one (n, dim, 2) draw is the reference's n calls of normal(size=2 dim), in its order (integrators.py:284-297)
    e, f = np.random.default_rng(5), np.random.default_rng(5)
    per_particle = np.array([e.normal(loc=0.0, scale=1.0, size=6).reshape(3, 2) for _ in range(50)])
    assert same_bits(per_particle, f.standard_normal((50, 3, 2)))


def test_usable_only_for_pcg64_generators():
    from turtlemd_b200 import npstream

    assert not npstream.usable(object())
    assert not npstream.usable(np.random.Generator(np.random.MT19937(1)))
    assert not npstream.usable(np.random.Generator(np.random.PCG64DXSM(1)))
    assert not npstream.usable(np.random.RandomState(1))


# ------------------------------------------------------------------------------- device passes run as host threads
@pytest.fixture(scope="module")
def emul(tmp_path_factory):
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    out = tmp_path_factory.mktemp("npy_emul") / "npy_emul.so"
    cmd = ["g++", "-std=c++20", "-O2", "-ffp-contract=off", "-DTMD_NPY_EMULATE", f"-I{REPO / 'tests' / 'emul'}",
           f"-I{REPO / 'turtlemd_b200' / 'csrc'}", "-x", "c++", str(REPO / "turtlemd_b200" / "csrc" / "npy_normal.cu"),
           "-shared", "-fPIC", "-o", str(out), "-lpthread"]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    assert proc.returncode == 0, proc.stderr[-3000:]
    lib = C.CDLL(str(out))
    lib.emul_last_error.restype = C.c_char_p
    return lib


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def test_log1p_restatement_is_libm(emul):
    variant = emul.tmd_npstream_libm_variant()
    assert variant in (0, 1), "libm's log1p is neither glibc build restated in npy_normal.h"
    r = np.random.default_rng(1)
    x = np.concatenate([
        -r.random(400_000), -r.random(100_000) * 1e-6, -r.random(100_000) ** 8, -r.random(100_000) * 2.0 ** -28,
        -r.random(100_000) * 2.0 ** -19, -1 + r.random(100_000) * 1e-6, -(1 - 2.0 ** -np.arange(1, 54)),
        -(2.0 ** -np.arange(1, 60)), -(0.5 + 2.0 ** -np.arange(2, 54)), -(0.5 - 2.0 ** -np.arange(2, 54)),
        [-0.0, -2.0 ** -53, -1 + 2.0 ** -53, -0.25, -0.75, -0.5]])
    y = np.empty_like(x)
    emul.emul_log1p(_ptr(x), _ptr(y), C.c_int64(len(x)), C.c_int32(variant))
    import math

    ref = np.array([math.log1p(v) for v in x])
    assert same_bits(y, ref)
    other = np.empty_like(x)
    emul.emul_log1p(_ptr(x), _ptr(other), C.c_int64(len(x)), C.c_int32(1 - variant))
    assert not same_bits(other, ref)  # the two builds do differ: the probe in libm_variant() has something to see


def test_pcg_jump_ahead_is_numpy_advance(emul):
    for seed, delta in ((1, 1), (2, 31), (3, 2 ** 40 + 12345), (4, 2 ** 63 + 7)):
        bg = np.random.PCG64(seed)
        words = state_words(np.random.Generator(bg))
        emul.emul_pcg_advance(_ptr(words), C.c_uint64(delta))
        bg.advance(delta)
        assert np.array_equal(words[:2], state_words(np.random.Generator(bg))[:2])


def test_device_passes_reproduce_numpy(emul):
    handle = C.c_void_p()
    assert emul.tmd_npstream_create(C.byref(handle)) == 0
    rewalks_seen = off_chain_seen = 0
    for seed, counts in ((0, [6, 1, 0, 7, 100, 4097, 50_000, 3]), (3, [400_000, 13]), (2 ** 63 + 5, [33_000, 33_000])):
        rgen = np.random.default_rng(seed)
        words = state_words(rgen)
        assert emul.tmd_npstream_seed(handle, _ptr(words), None) == 0
        total_words = 0
        for n in counts:
            out = np.full(n + 8, -77.0)
            assert emul.tmd_npstream_normals(handle, _ptr(out), C.c_int64(n), None) == 0, emul.emul_last_error()
            ref = rgen.standard_normal(n)
            assert same_bits(out[:n], ref), (seed, n)
            assert (out[n:] == -77.0).all()  # nothing written behind the normals asked for
            state = np.zeros(2, dtype=np.uint64)
            used, counters = C.c_uint64(0), np.zeros(2, dtype=np.uint64)
            rc = emul.tmd_npstream_state(handle, _ptr(state), C.byref(used), _ptr(counters), None)
            assert rc == 0, emul.emul_last_error()
            assert np.array_equal(state, state_words(rgen)[:2]), (seed, n)  # the stream continues where NumPy stands
            assert used.value >= total_words + n
            total_words = used.value
        rewalks_seen += int(counters[0])
        off_chain_seen += int(counters[1])
    assert rewalks_seen > 0  # chains off the canonical one were met and walked again ...
    assert off_chain_seen > 0  # ... also by the pass that places the normals (a true entry off the canonical chain)
    assert emul.tmd_npstream_destroy(handle) == 0


def test_device_passes_many_short_calls(emul):
    """The stream picked up again and again at arbitrary places (every call ends inside some chunk, the next one starts
    there): outputs and generator state after each of 30 calls of random length."""
    handle = C.c_void_p()
    assert emul.tmd_npstream_create(C.byref(handle)) == 0
    rgen = np.random.default_rng(99)
    assert emul.tmd_npstream_seed(handle, _ptr(state_words(rgen)), None) == 0
    for n in np.random.default_rng(1).integers(1, 3000, size=30):
        n = int(n)
        out = np.full(n + 4, -77.0)
        assert emul.tmd_npstream_normals(handle, _ptr(out), C.c_int64(n), None) == 0, emul.emul_last_error()
        assert same_bits(out[:n], rgen.standard_normal(n)) and (out[n:] == -77.0).all()
        state = np.zeros(2, dtype=np.uint64)
        assert emul.tmd_npstream_state(handle, _ptr(state), None, None, None) == 0, emul.emul_last_error()
        assert np.array_equal(state, state_words(rgen)[:2])
    assert emul.tmd_npstream_destroy(handle) == 0


def test_device_stream_refuses_unseeded_and_bad_increment(emul):
    handle = C.c_void_p()
    assert emul.tmd_npstream_create(C.byref(handle)) == 0
    out = np.zeros(4)
    assert emul.tmd_npstream_normals(handle, _ptr(out), C.c_int64(4), None) != 0  # never seeded
    even = np.array([1, 2, 3, 4], dtype=np.uint64)
    assert emul.tmd_npstream_seed(handle, _ptr(even), None) != 0  # a PCG increment is odd
    emul.tmd_npstream_destroy(handle)


_OTHER_BUILD = r"""
import ctypes as C, math, sys
import numpy as np
sys.path.insert(0, sys.argv[1])
from turtlemd_b200 import _lib
lib = _lib.load()
variant = lib.tmd_npstream_libm_variant()
rgen = np.random.default_rng(21)
st = rgen.bit_generator.state["state"]
M = (1 << 64) - 1
words = np.array([st["state"] >> 64, st["state"] & M, st["inc"] >> 64, st["inc"] & M], dtype=np.uint64)
out = np.empty(8_000_000)
_lib.call("tmd_host_npstream_normals", _lib.hptr(words), _lib.hptr(out), out.size, None)
ok = np.array_equal(out.view(np.uint64), rgen.standard_normal(out.size).view(np.uint64))
print(variant, int(ok))
"""


def test_both_glibc_log1p_builds_are_followed():
    """glibc >= 2.39 picks a log1p compiled with FMA on CPUs that have it; GLIBC_TUNABLES turns that off.  The
    library must recognise whichever build its process runs and stay bit-exact with NumPy in it."""
    seen = {}
    for label, tunables in (("default", None), ("baseline", "glibc.cpu.hwcaps=-FMA,-AVX2")):
        env = dict(os.environ)
        if tunables is None:
            env.pop("GLIBC_TUNABLES", None)
        else:
            env["GLIBC_TUNABLES"] = tunables
        proc = subprocess.run([sys.executable, "-c", _OTHER_BUILD, str(REPO)], capture_output=True, text=True, env=env,
                              timeout=600)
        assert proc.returncode == 0, proc.stderr[-2000:]
        variant, ok = (int(v) for v in proc.stdout.split())
        assert variant in (0, 1) and ok == 1, (label, proc.stdout)
        seen[label] = variant
    assert seen["baseline"] == 0


# ------------------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
def test_gpu_stream_is_numpy_bit_for_bit():
    from turtlemd_b200 import npstream

    rgen = np.random.default_rng(2024)
    twin = np.random.default_rng(2024)
    assert npstream.usable(rgen)
    stream = npstream.NumpyStream(rgen)
    for n in (6, 0, 1, 4097, 6_000_000, 100_003, 17_000_000):
        stream.begin()
        got = stream.normals(n).cpu().numpy()
        stream.end()
        assert same_bits(got, twin.standard_normal(n)), n
        assert rgen.bit_generator.state == twin.bit_generator.state, n
    # a draw on the host between two device calls is seen (begin() re-seeds from the host generator)
    assert rgen.random() == twin.random()
    stream.begin()
    a = stream.normals(1000).cpu().numpy()
    b = stream.normals(50).cpu().numpy()  # two calls inside one begin / end
    stream.end()
    assert same_bits(a, twin.standard_normal(1000)) and same_bits(b, twin.standard_normal(50))
    assert rgen.bit_generator.state == twin.bit_generator.state
    words, rewalks, off_chain = stream.stats()
    assert words > 1050 and rewalks > 0


def _replicas(n, dim, seed):
    from turtlemd_b200.potentials import DoubleWell
    from turtlemd_b200.system import Box, Particles, System

    r = np.random.default_rng(seed)
    pos = -1.0 + 0.1 * r.standard_normal((n, dim))
    vel = 0.2 * r.standard_normal((n, dim))
    mass = np.where(np.arange(n) % 3 == 0, 2.5, 1.0)
    system = System(Box(periodic=[False] * dim), Particles.from_arrays(pos, vel=vel, mass=mass),
                    [DoubleWell(a=1.0, b=2.0, c=0.0)])
    system.potential_and_force()
    return system


@pytest.mark.gpu
@pytest.mark.parametrize("dim", [1, 3])
def test_gpu_langevin_inertia_under_default_rng_matches_its_host_draw_path(dim):
    """LangevinInertia(seed=s) -- the reference's default_rng(s) -- with the stream continued on the device against the
    same integrator drawing on the host like the reference does; run() against single steps; generator positions."""
    from turtlemd_b200 import npstream
    from turtlemd_b200.integrators import LangevinInertia

    n, steps = 4001, 12
    systems = {}
    gens = {}
    for label in ("device_run", "device_steps", "host"):
        system = _replicas(n, dim, 8)
        integ = LangevinInertia(timestep=0.002, gamma=0.3, beta=1.0 / 0.07, seed=31)
        npstream.ENABLED = label != "host"
        try:
            if label == "device_run":
                integ.run(system, steps)
                assert integ._np_stream is not None  # the device stream was the path taken
            else:
                for _ in range(steps):
                    integ.integration_step(system)
        finally:
            npstream.ENABLED = True
        systems[label] = system
        gens[label] = integ.rgen
    ref = systems["host"].particles
    for label in ("device_run", "device_steps"):
        part = systems[label].particles
        assert np.max(np.abs(part.pos - ref.pos)) < 1e-12, label
        assert np.max(np.abs(part.vel - ref.vel)) < 1e-12, label
        assert np.max(np.abs(part.force - ref.force)) < 1e-10, label
        assert abs(part.v_pot - ref.v_pot) < 1e-9 * max(1.0, abs(ref.v_pot)), label
        assert gens[label].bit_generator.state == gens["host"].bit_generator.state, label
    assert np.array_equal(systems["device_run"].particles.pos, systems["device_steps"].particles.pos)
    assert np.array_equal(systems["device_run"].particles.vel, systems["device_steps"].particles.vel)


@pytest.mark.gpu
def test_gpu_langevin_overdamped_under_default_rng_matches_its_host_draw_path():
    from turtlemd_b200 import npstream
    from turtlemd_b200.integrators import LangevinOverdamped

    n, dim, steps = 3000, 2, 10
    out = {}
    for label in ("device", "host"):
        system = _replicas(n, dim, 5)
        rgen = np.random.default_rng(77)
        integ = LangevinOverdamped(timestep=0.002, gamma=0.3, rgen=rgen, beta=1.0 / 0.07)
        npstream.ENABLED = label == "device"
        try:
            integ.run(system, steps // 2)
            for _ in range(steps - steps // 2):
                integ.integration_step(system)
        finally:
            npstream.ENABLED = True
        out[label] = (system.particles.pos.copy(), system.particles.vel.copy(), rgen.bit_generator.state)
    assert np.max(np.abs(out["device"][0] - out["host"][0])) < 1e-12
    assert np.max(np.abs(out["device"][1] - out["host"][1])) < 1e-12
    assert out["device"][2] == out["host"][2]


@pytest.mark.gpu
def test_gpu_lj_langevin_under_default_rng_matches_its_host_draw_path():
    """The neighbour-list fluid (N = 4000) under LangevinInertia(seed=...): device stream against host draws."""
    from turtlemd_b200 import npstream
    from turtlemd_b200.integrators import LangevinInertia
    from turtlemd_b200.potentials import LennardJonesCut
    from turtlemd_b200.system import Box, Particles, System
    from turtlemd_b200.tools import generate_lattice

    pos0, size = generate_lattice("fcc", [10, 10, 10], density=0.8)
    r = np.random.default_rng(2)
    pos0 = pos0 + 0.05 * (2.0 * r.random(pos0.shape) - 1.0)
    vel0 = 0.5 * r.standard_normal(pos0.shape)
    out = {}
    for label in ("device", "host"):
        pot = LennardJonesCut(dim=3, shift=True)
        pot.set_parameters({0: {"sigma": 1.0, "epsilon": 1.0, "rcut": 2.5}})
        system = System(Box(low=size[:, 0], high=size[:, 1]), Particles.from_arrays(pos0, vel=vel0), [pot])
        system.potential_and_force()
        integ = LangevinInertia(timestep=0.002, gamma=0.5, beta=1.25, seed=6)
        npstream.ENABLED = label == "device"
        try:
            integ.run(system, 8)
        finally:
            npstream.ENABLED = True
        out[label] = (system.particles.pos.copy(), system.particles.vel.copy(), integ.rgen.bit_generator.state)
    assert np.max(np.abs(out["device"][0] - out["host"][0])) < 1e-10
    assert np.max(np.abs(out["device"][1] - out["host"][1])) < 1e-9
    assert out["device"][2] == out["host"][2]
