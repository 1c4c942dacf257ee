// "numpy-normal": numpy.random.Generator(PCG64).standard_normal restated for host and device.
//
// The reference draws all of its noise from a caller-owned numpy Generator -- rgen.normal(...) at
// turtlemd/integrators.py:156-160 (LangevinOverdamped), :312 (LangevinInertia through multivariate_normal) and
// turtlemd/system/particles.py:260 (Maxwell velocities) -- and LangevinInertia's default is numpy.random.default_rng(seed)
// (integrators.py:216).  NumPy is a third-party dependency of the reference (2.3.5 in this image; the stream has been
// frozen since 1.17 by NumPy's compatibility policy); what is restated here is its published algorithm:
//   * PCG64 = O'Neill's pcg_setseq_128 LCG (multiplier 0x2360ED051FC65DA44385DF649FCCF645) with the XSL-RR 128/64
//     output function, stepped BEFORE the output is taken; next_double = (next_uint64 >> 11) * 2^-53;
//   * standard_normal = the 256-layer ziggurat of numpy/random/src/distributions/distributions.c
//     (random_standard_normal): one 64-bit word gives layer (8 bits), sign (1 bit) and a 52-bit abscissa; 98.8 % of the
//     words return on the spot; the wedges cost one more word, the tail two per attempt;
//   * the tail uses log1p, whose VALUE is returned: glibc's log1p (sysdeps/ieee754/dbl-64/s_log1p.c, the fdlibm
//     algorithm) is restated operation by operation so that host, device and NumPy agree to the last bit; the wedge test
//     only COMPARES against exp(-x^2/2), any exp correct to an ulp decides identically (a disagreement needs the two
//     sides to be equal to 2^-52 relative: ~1e-16 per wedge test).
// Pinned by tests/test_npy_normal.py against NumPy itself, bit for bit: the host twin over 6e7 draws, log1p against libm
// in both builds, and the device passes (the kernels of npy_normal.cu run as host threads) -- all without a GPU; on the
// GPU the stream is compared with NumPy again and the Langevin integrators with their host-draw path.
//
// The layer tables are data of the published algorithm (npy_ziggurat_tables.h, written by make_npy_tables.py).
#pragma once

#include <math.h>
#include <stdint.h>
#include <string.h>

#include "npy_ziggurat_tables.h"
#include "philox_normal.h"  // TMD_HD, TMD_ADD / TMD_MUL ... : individually rounded IEEE operations on both sides

namespace tmd {
namespace npy {

typedef unsigned __int128 u128;

struct Pcg64 {
    u128 state, inc;
};

TMD_HD u128 make_u128(uint64_t hi, uint64_t lo) { return ((u128)hi << 64) | (u128)lo; }
TMD_HD u128 pcg_multiplier() { return make_u128(2549297995355413924ull, 4865540595714422341ull); }

TMD_HD uint64_t pcg_output(u128 s) {
    const uint64_t hi = (uint64_t)(s >> 64), lo = (uint64_t)s;
    const uint64_t x = hi ^ lo;
    const unsigned rot = (unsigned)(hi >> 58);
    return (x >> rot) | (x << ((64u - rot) & 63u));
}

TMD_HD uint64_t pcg_next(Pcg64 &g) {
    g.state = g.state * pcg_multiplier() + g.inc;
    return pcg_output(g.state);
}

// the state `delta` steps ahead (Brown, "Random number generation with arbitrary strides"): O(log delta)
TMD_HD u128 pcg_advance(u128 state, u128 inc, uint64_t delta) {
    u128 acc_mult = 1, acc_plus = 0, cur_mult = pcg_multiplier(), cur_plus = inc;
    while (delta > 0) {
        if (delta & 1ull) {
            acc_mult *= cur_mult;
            acc_plus = acc_plus * cur_mult + cur_plus;
        }
        cur_plus = (cur_mult + 1) * cur_plus;
        cur_mult *= cur_mult;
        delta >>= 1;
    }
    return acc_mult * state + acc_plus;
}

TMD_HD double u64_as_double(uint64_t b) {
#if defined(__CUDA_ARCH__)
    return __longlong_as_double((long long)b);
#else
    double d;
    memcpy(&d, &b, sizeof d);
    return d;
#endif
}
TMD_HD uint64_t double_as_u64(double d) {
#if defined(__CUDA_ARCH__)
    return (uint64_t)__double_as_longlong(d);
#else
    uint64_t b;
    memcpy(&b, &d, sizeof b);
    return b;
#endif
}
TMD_HD int32_t high_word(double d) { return (int32_t)(double_as_u64(d) >> 32); }
TMD_HD double with_high_word(double d, int32_t hi) {
    return u64_as_double((double_as_u64(d) & 0xffffffffull) | ((uint64_t)(uint32_t)hi << 32));
}

#if defined(__CUDA_ARCH__)
#define TMD_FMA_EXACT(a, b, c) __fma_rn((a), (b), (c))
#else
#define TMD_FMA_EXACT(a, b, c) __builtin_fma((a), (b), (c))
#endif

// a*b + c the way the libm variant in use evaluates it: glibc >= 2.39 on x86-64 selects (IFUNC) a log1p compiled with
// -mfma -mavx2 on CPUs that have both, in which the compiler contracted a fixed set of multiply-adds; the baseline
// variant rounds the product and the sum separately.  `fused` says which variant NumPy's process uses (every call
// site below is one the FMA build contracts -- read off the disassembly of libm 2.39's two variants).
TMD_HD double mul_add(bool fused, double a, double b, double c) {
    return fused ? TMD_FMA_EXACT(a, b, c) : TMD_ADD(TMD_MUL(a, b), c);
}

// glibc's log1p (sysdeps/ieee754/dbl-64/s_log1p.c, the fdlibm algorithm) for -1 < x < 0.41 -- the ziggurat tail calls it
// with x = -U, U in [0, 1) -- operation by operation in glibc's order.  x = -0.0 (U = 0) returns x like glibc.
TMD_HD double log1p_glibc(double x, bool fused) {
    const double ln2_hi = 6.93147180369123816490e-01, ln2_lo = 1.90821492927058770002e-10;
    const double Lp1 = 6.666666666666735130e-01, Lp2 = 3.999999999940941908e-01, Lp3 = 2.857142874366239149e-01,
                 Lp4 = 2.222219843214978396e-01, Lp5 = 1.818357216161805012e-01, Lp6 = 1.531383769920937332e-01,
                 Lp7 = 1.479819860511658591e-01;
    const int32_t hx = high_word(x);
    const int32_t ax = hx & 0x7fffffff;
    int32_t k = 1, hu = 0;
    double f = 0.0, c = 0.0;
    if (hx < 0x3FDA827A) {           // x < 0.41422
        if (ax >= 0x3ff00000) {      // x <= -1: not reachable from the ziggurat (U < 1)
            return (x == -1.0) ? -HUGE_VAL : u64_as_double(0x7ff8000000000000ull);
        }
        if (ax < 0x3e200000) {       // |x| < 2^-29
            if (ax < 0x3c900000) return x;
            return mul_add(fused, -TMD_MUL(x, x), 0.5, x);  // x - x*x*0.5
        }
        if (hx > 0 || hx <= (int32_t)0xbfd2bec3) {  // -0.2929 < x < 0.41422
            k = 0;
            f = x;
            hu = 1;
        }
    }
    if (k != 0) {
        double u = TMD_ADD(1.0, x);
        hu = high_word(u);
        k = (hu >> 20) - 1023;
        c = (k > 0) ? TMD_SUB(1.0, TMD_SUB(u, x)) : TMD_SUB(x, TMD_SUB(u, 1.0));  // correction term
        c = TMD_DIV(c, u);
        hu &= 0x000fffff;
        if (hu < 0x6a09e) {
            u = with_high_word(u, hu | 0x3ff00000);  // normalise u
        } else {
            k += 1;
            u = with_high_word(u, hu | 0x3fe00000);  // normalise u / 2
            hu = (0x00100000 - hu) >> 2;
        }
        f = TMD_SUB(u, 1.0);
    }
    const double hfsq = TMD_MUL(TMD_MUL(0.5, f), f);
    const double dk = (double)k;
    if (hu == 0) {  // |f| < 2^-20
        if (f == 0.0) {
            if (k == 0) return 0.0;
            return mul_add(fused, dk, ln2_hi, mul_add(fused, dk, ln2_lo, c));
        }
        const double R0 = TMD_MUL(hfsq, mul_add(fused, -f, 0.66666666666666666, 1.0));
        if (k == 0) return TMD_SUB(f, R0);
        return mul_add(fused, dk, ln2_hi, -TMD_SUB(TMD_SUB(R0, mul_add(fused, dk, ln2_lo, c)), f));
    }
    const double s = TMD_DIV(f, TMD_ADD(2.0, f));
    const double z = TMD_MUL(s, s);
    const double z2 = TMD_MUL(z, z), z4 = TMD_MUL(z2, z2), z6 = TMD_MUL(z4, z2);
    const double R2 = mul_add(fused, z, Lp3, Lp2);
    const double R3 = mul_add(fused, z, Lp5, Lp4);
    const double R4 = mul_add(fused, z, Lp7, Lp6);
    // R = R1 + z2*R2 + z4*R3 + z6*R4 with R1 = z*Lp1, summed left to right
    const double R = mul_add(fused, z6, R4, mul_add(fused, z4, R3, mul_add(fused, z, Lp1, TMD_MUL(z2, R2))));
    const double sr = TMD_MUL(s, TMD_ADD(hfsq, R));
    if (k == 0) return TMD_SUB(f, TMD_SUB(hfsq, sr));
    return mul_add(fused, dk, ln2_hi, -TMD_SUB(TMD_SUB(hfsq, TMD_ADD(sr, mul_add(fused, dk, ln2_lo, c))), f));
}

struct ZigTables {
    const uint64_t *ki;
    const double *wi, *fi;
    bool fused_log1p;  // which libm log1p variant NumPy's process runs (see mul_add)
};

TMD_HD double next_double(Pcg64 &g, uint32_t &used) {
    ++used;
    return TMD_MUL((double)(pcg_next(g) >> 11), 1.0 / 9007199254740992.0);
}

// random_standard_normal of distributions.c in two pieces, so that a kernel can run the common case of many lanes in
// lock step and gather the rare remainder.
// Piece 1: the candidate x of word r, and whether it is accepted on the spot (98.8 % of the words).
TMD_HD bool zig_fast(uint64_t r, const ZigTables &t, double &x) {
    const int idx = (int)(r & 0xff);
    const uint64_t rabs = (r >> 9) & 0x000fffffffffffffull;
    x = TMD_MUL((double)rabs, t.wi[idx]);
    if ((r >> 8) & 0x1) x = -x;
    return rabs < t.ki[idx];
}

// Piece 2: word r was not accepted on the spot; the wedge test (one more word) or the tail (two words per attempt), and
// after a rejected wedge the whole procedure again from a fresh word.  `used` counts the words drawn here.
TMD_HD double zig_slow(Pcg64 &g, const ZigTables &t, uint64_t r, double x, uint32_t &used) {
    const double nor_r = 3.6541528853610087963519472518, nor_inv_r = 0.27366123732975827203338247596;
    for (;;) {
        const int idx = (int)(r & 0xff);
        if (idx == 0) {
            const uint64_t rabs = (r >> 9) & 0x000fffffffffffffull;
            for (;;) {
                const double xx = TMD_MUL(-nor_inv_r, log1p_glibc(-next_double(g, used), t.fused_log1p));
                const double yy = -log1p_glibc(-next_double(g, used), t.fused_log1p);
                if (TMD_ADD(yy, yy) > TMD_MUL(xx, xx))
                    return ((rabs >> 8) & 0x1) ? -TMD_ADD(nor_r, xx) : TMD_ADD(nor_r, xx);
            }
        }
        const double u = next_double(g, used);
        const double lhs = TMD_ADD(TMD_MUL(TMD_SUB(t.fi[idx - 1], t.fi[idx]), u), t.fi[idx]);
        if (lhs < exp(TMD_MUL(TMD_MUL(-0.5, x), x))) return x;
        r = pcg_next(g);
        ++used;
        if (zig_fast(r, t, x)) return x;
    }
}

// One standard normal; `used` counts the 64-bit words it consumed.
TMD_HD double standard_normal(Pcg64 &g, const ZigTables &t, uint32_t &used) {
    const uint64_t r = pcg_next(g);
    ++used;
    double x;
    if (zig_fast(r, t, x)) return x;
    return zig_slow(g, t, r, x, used);
}

}  // namespace npy
}  // namespace tmd
