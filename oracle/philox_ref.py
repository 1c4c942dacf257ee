"""Oracle for the counter-based RNG of the Langevin kernels (test infrastructure).

The reference draws its noise from ``numpy.random.Generator.normal``
(``turtlemd/integrators.py:156-160`` and ``:312``); the generator is a
caller-owned, duck-typed object.  The CUDA path needs a *counter-based*
stream, so the product defines one ("tmd-normal-v1"):

* raw words: Philox4x64-10 (Salmon et al., SC'11 "Parallel random numbers:
  as easy as 1, 2, 3"), laid out exactly like ``numpy.random.Philox(key=k)``:
  raw word ``m`` is lane ``m % 4`` of the block with counter
  ``(m // 4 + 1, 0, 0, 0)`` and key ``(k, 0)``.  NumPy (2.3.5 here) is the
  third-party dependency that pins this; `test_oracle_golden.py` compares
  against ``Philox(key).random_raw``.
* uniforms: ``u = (raw >> 11) * 2**-53`` (NumPy's ``random()``).
* normals: Box-Muller on consecutive raw words ``(2P, 2P+1)`` with
  ``u1 = ((r0 >> 11) + 1) * 2**-53`` in (0, 1] and ``u2 = (r1 >> 11) * 2**-53``;
  ``z[2P] = sqrt(-2 log u1) cos(2 pi u2)``, ``z[2P+1] = ... sin(2 pi u2)``.
  ``log`` and ``sincos`` are evaluated with a fixed sequence of individually
  rounded IEEE-754 ``+ - * / sqrt`` (no FMA), so this NumPy restatement, the
  host C code and the CUDA kernel produce bit-identical doubles.

Everything here is written with plain NumPy element-wise operations, each of
which rounds once, which is what makes it a bit-exact twin.
"""
from __future__ import annotations

import numpy as np

MASK64 = (1 << 64) - 1
PHILOX_M0 = 0xD2E7470EE14C6C93
PHILOX_M1 = 0xCA5A826395121157
PHILOX_W0 = 0x9E3779B97F4A7C15
PHILOX_W1 = 0xBB67AE8584CAA73B


def philox4x64_10_scalar(counter, key):
    """One Philox4x64-10 block with Python integers (exact, slow)."""
    c = [int(x) & MASK64 for x in counter]
    k = [int(x) & MASK64 for x in key]
    for _ in range(10):
        p0 = PHILOX_M0 * c[0]
        p1 = PHILOX_M1 * c[2]
        c = [
            (p1 >> 64) ^ c[1] ^ k[0],
            p1 & MASK64,
            (p0 >> 64) ^ c[3] ^ k[1],
            p0 & MASK64,
        ]
        k = [(k[0] + PHILOX_W0) & MASK64, (k[1] + PHILOX_W1) & MASK64]
    return c


def _mulhilo64(a: int, b: np.ndarray):
    """(hi, lo) words of a * b for a Python int ``a`` and uint64 array ``b``."""
    a_lo = np.uint64(a & 0xFFFFFFFF)
    a_hi = np.uint64(a >> 32)
    m32 = np.uint64(0xFFFFFFFF)
    s32 = np.uint64(32)
    b_lo = b & m32
    b_hi = b >> s32
    ll = a_lo * b_lo
    lh = a_lo * b_hi
    hl = a_hi * b_lo
    hh = a_hi * b_hi
    mid = (ll >> s32) + (lh & m32) + (hl & m32)
    lo = (ll & m32) | ((mid & m32) << s32)
    hi = hh + (lh >> s32) + (hl >> s32) + (mid >> s32)
    return hi, lo


def philox4x64_10(counter0: np.ndarray, key0: int, key1: int = 0) -> np.ndarray:
    """Vectorised Philox4x64-10 for counters ``(counter0[i], 0, 0, 0)``.

    Returns a ``(len(counter0), 4)`` uint64 array.
    """
    with np.errstate(over="ignore"):
        c0 = np.asarray(counter0, dtype=np.uint64).copy()
        c1 = np.zeros_like(c0)
        c2 = np.zeros_like(c0)
        c3 = np.zeros_like(c0)
        k0 = int(key0) & MASK64
        k1 = int(key1) & MASK64
        for _ in range(10):
            hi0, lo0 = _mulhilo64(PHILOX_M0, c0)
            hi1, lo1 = _mulhilo64(PHILOX_M1, c2)
            c0, c1, c2, c3 = (
                hi1 ^ c1 ^ np.uint64(k0),
                lo1,
                hi0 ^ c3 ^ np.uint64(k1),
                lo0,
            )
            k0 = (k0 + PHILOX_W0) & MASK64
            k1 = (k1 + PHILOX_W1) & MASK64
    return np.stack([c0, c1, c2, c3], axis=1)


def raw_words(key: int, first: int, count: int) -> np.ndarray:
    """Raw uint64 words ``first .. first+count-1`` of the NumPy-compatible stream."""
    if count == 0:
        return np.zeros(0, dtype=np.uint64)
    b0 = first // 4
    b1 = (first + count - 1) // 4
    ctr = np.arange(b0 + 1, b1 + 2, dtype=np.uint64)
    words = philox4x64_10(ctr, key).reshape(-1)
    off = first - 4 * b0
    return words[off : off + count]


# ---- the fixed-arithmetic normal transform ("tmd-normal-v1") -------------

TWO_M53 = 2.0**-53
LN2_HI = 6.93147180369123816490e-01
LN2_LO = 1.90821492927058770002e-10
LG1 = 6.666666666666735130e-01
LG2 = 3.999999999940941908e-01
LG3 = 2.857142874366239149e-01
LG4 = 2.222219843214978396e-01
LG5 = 1.818357216161805012e-01
LG6 = 1.531383769920937332e-01
LG7 = 1.479819860511658591e-01
SQRT2 = 1.41421356237309514547e00
PI_4 = 7.85398163397448278999e-01
# Taylor coefficients (-1)^k/(2k+1)! and (-1)^k/(2k)!, correctly rounded.
SIN_C = [
    -1.66666666666666657415e-01,
    8.33333333333333321769e-03,
    -1.98412698412698412526e-04,
    2.75573192239858925110e-06,
    -2.50521083854417202239e-08,
    1.60590438368216133196e-10,
    -7.64716373181981641310e-13,
    2.81145725434552059581e-15,
]
COS_C = [
    -5.00000000000000000000e-01,
    4.16666666666666643537e-02,
    -1.38888888888888894189e-03,
    2.48015873015873015658e-05,
    -2.75573192239858882758e-07,
    2.08767569878681001532e-09,
    -1.14707455977297245072e-11,
    4.77947733238738524982e-14,
    -1.56192069685862252482e-16,
]


def fixed_log(x: np.ndarray) -> np.ndarray:
    """log(x) for normal doubles in (0, 1] by the classic k*ln2 + log1p(f) split."""
    x = np.asarray(x, dtype=np.float64)
    bits = x.view(np.uint64)
    k = ((bits >> np.uint64(52)) & np.uint64(0x7FF)).astype(np.int64) - 1023
    mant = (bits & np.uint64(0x000FFFFFFFFFFFFF)) | np.uint64(0x3FF0000000000000)
    m = mant.view(np.float64)
    big = m > SQRT2
    m = np.where(big, m * 0.5, m)
    k = np.where(big, k + 1, k)
    dk = k.astype(np.float64)
    f = m - 1.0
    s = f / (2.0 + f)
    z = s * s
    w = z * z
    t1 = w * (LG2 + w * (LG4 + w * LG6))
    t2 = z * (LG1 + w * (LG3 + w * (LG5 + w * LG7)))
    r = t2 + t1
    hfsq = (0.5 * f) * f
    return dk * LN2_HI - ((hfsq - (s * (hfsq + r) + dk * LN2_LO)) - f)


def fixed_sincos_2pi(u: np.ndarray):
    """(sin, cos) of 2*pi*u for u in [0, 1) by exact octant reduction."""
    u = np.asarray(u, dtype=np.float64)
    y = 8.0 * u
    q = np.floor(y)
    r = y - q
    qi = q.astype(np.int64)
    odd = (qi & 1) == 1
    r = np.where(odd, 1.0 - r, r)
    t = PI_4 * r
    z = t * t
    ps = SIN_C[7]
    for c in SIN_C[6::-1]:
        ps = c + z * ps
    st = t + t * (z * ps)
    pc = COS_C[8]
    for c in COS_C[7::-1]:
        pc = c + z * pc
    ct = 1.0 + z * pc
    # octant 0: (s,c)  1: (c,s)  2: (c,-s)  3: (s,-c)  4: (-s,-c)  5: (-c,-s)  6: (-c,s)  7: (-s,c)
    swap = ((qi + 1) & 2) == 2
    sin_mag = np.where(swap, ct, st)
    cos_mag = np.where(swap, st, ct)
    sin_neg = qi >= 4
    cos_neg = ((qi + 2) & 4) == 4
    return np.where(sin_neg, -sin_mag, sin_mag), np.where(cos_neg, -cos_mag, cos_mag)


def normals(key: int, first: int, count: int) -> np.ndarray:
    """Standard normals ``first .. first+count-1`` of the stream with ``key``."""
    if count == 0:
        return np.zeros(0)
    p0 = first // 2
    p1 = (first + count - 1) // 2
    raw = raw_words(key, 2 * p0, 2 * (p1 - p0 + 1)).reshape(-1, 2)
    u1 = ((raw[:, 0] >> np.uint64(11)) + np.uint64(1)).astype(np.float64) * TWO_M53
    u2 = (raw[:, 1] >> np.uint64(11)).astype(np.float64) * TWO_M53
    rad = np.sqrt(-2.0 * fixed_log(u1))
    sn, cs = fixed_sincos_2pi(u2)
    z = np.stack([rad * cs, rad * sn], axis=1).reshape(-1)
    off = first - 2 * p0
    return z[off : off + count]


class OraclePhiloxNormal:
    """Duck-typed ``rgen`` (only ``normal(loc, scale, size)``) over the stream above.

    Handing this to the *reference* integrators makes them consume exactly the
    draws the CUDA kernels generate in-kernel: the reference calls
    ``rgen.normal`` particle-major (``turtlemd/integrators.py:284-297``), i.e.
    draw index ``i*2d + 2k + c``, and for the overdamped integrator in one
    C-order call (``:156-160``), i.e. draw index ``i*d + k``.
    """

    def __init__(self, key: int = 0, position: int = 0):
        self.key = int(key)
        self.position = int(position)

    def normal(self, loc=0.0, scale=1.0, size=None):
        shape = () if size is None else (tuple(size) if np.iterable(size) else (int(size),))
        count = int(np.prod(shape)) if shape else 1
        z = normals(self.key, self.position, count)
        self.position += count
        out = loc + scale * z.reshape(shape) if shape else loc + scale * z[0]
        return out
