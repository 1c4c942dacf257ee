// "tmd-normal-v1": the counter-based normal stream of the Langevin kernels.
//
// One source for host and device.  Raw words are Philox4x64-10 (Salmon, Moraes, Dror, Shaw,
// SC'11) laid out like numpy.random.Philox(key=k): raw word m = lane m%4 of the block with
// counter (m/4 + 1, 0, 0, 0), key (k, 0).  Normals come in Box-Muller pairs from raw words
// (2P, 2P+1).  log and sincos are fixed sequences of individually rounded IEEE add/mul/div/sqrt
// (explicit *_rn intrinsics on the device, -ffp-contract=off on the host), so the host twin, the
// NumPy oracle (oracle/philox_ref.py) and the kernels agree to the last bit.
//
// Replaces: rgen.normal(...) at turtlemd/integrators.py:156-160 and :312 (the reference's
// generator is a caller-owned duck-typed object; see DESIGN.md "RNG contract").
#pragma once

#include <stdint.h>

#if defined(__CUDACC__)
#define TMD_HD __host__ __device__ __forceinline__
#else
#define TMD_HD static inline
#endif

#if defined(__CUDA_ARCH__)
#define TMD_ADD(a, b) __dadd_rn((a), (b))
#define TMD_SUB(a, b) __dsub_rn((a), (b))
#define TMD_MUL(a, b) __dmul_rn((a), (b))
#define TMD_DIV(a, b) __ddiv_rn((a), (b))
#define TMD_SQRT(a) __dsqrt_rn((a))
#else
#include <math.h>
#define TMD_ADD(a, b) ((a) + (b))
#define TMD_SUB(a, b) ((a) - (b))
#define TMD_MUL(a, b) ((a) * (b))
#define TMD_DIV(a, b) ((a) / (b))
#define TMD_SQRT(a) sqrt((a))
#endif

namespace tmd {

struct Philox4 {
    uint64_t w[4];
};

TMD_HD void mulhilo64(uint64_t a, uint64_t b, uint64_t &hi, uint64_t &lo) {
#if defined(__CUDA_ARCH__)
    hi = __umul64hi(a, b);
    lo = a * b;
#else
    unsigned __int128 p = (unsigned __int128)a * b;
    hi = (uint64_t)(p >> 64);
    lo = (uint64_t)p;
#endif
}

// Block `block` (0-based) of the NumPy-compatible stream: counter = block + 1.
TMD_HD Philox4 philox4x64_10(uint64_t block, uint64_t key0) {
    uint64_t c0 = block + 1, c1 = 0, c2 = 0, c3 = 0;
    uint64_t k0 = key0, k1 = 0;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint64_t hi0, lo0, hi1, lo1;
        mulhilo64(0xD2E7470EE14C6C93ull, c0, hi0, lo0);
        mulhilo64(0xCA5A826395121157ull, c2, hi1, lo1);
        uint64_t n0 = hi1 ^ c1 ^ k0;
        uint64_t n2 = hi0 ^ c3 ^ k1;
        c0 = n0;
        c1 = lo1;
        c2 = n2;
        c3 = lo0;
        k0 += 0x9E3779B97F4A7C15ull;
        k1 += 0xBB67AE8584CAA73Bull;
    }
    Philox4 out;
    out.w[0] = c0;
    out.w[1] = c1;
    out.w[2] = c2;
    out.w[3] = c3;
    return out;
}

TMD_HD double bits_to_double(uint64_t b) {
#if defined(__CUDA_ARCH__)
    return __longlong_as_double((long long)b);
#else
    union {
        uint64_t u;
        double d;
    } cv;
    cv.u = b;
    return cv.d;
#endif
}

TMD_HD uint64_t double_to_bits(double d) {
#if defined(__CUDA_ARCH__)
    return (uint64_t)__double_as_longlong(d);
#else
    union {
        uint64_t u;
        double d;
    } cv;
    cv.d = d;
    return cv.u;
#endif
}

// log(x) for normal doubles in (0, 1]: x = 2^k m, m in (sqrt(1/2), sqrt(2)], f = m - 1,
// s = f/(2+f), log(1+f) = f - f^2/2 + s (f^2/2 + R(s^2)); R = classic degree-14 minimax.
TMD_HD double fixed_log(double x) {
    const double ln2_hi = 6.93147180369123816490e-01, ln2_lo = 1.90821492927058770002e-10;
    const double Lg1 = 6.666666666666735130e-01, Lg2 = 3.999999999940941908e-01,
                 Lg3 = 2.857142874366239149e-01, Lg4 = 2.222219843214978396e-01,
                 Lg5 = 1.818357216161805012e-01, Lg6 = 1.531383769920937332e-01,
                 Lg7 = 1.479819860511658591e-01;
    uint64_t bits = double_to_bits(x);
    int k = (int)((bits >> 52) & 0x7FF) - 1023;
    double m = bits_to_double((bits & 0x000FFFFFFFFFFFFFull) | 0x3FF0000000000000ull);
    if (m > 1.41421356237309514547e+00) {
        m = TMD_MUL(m, 0.5);
        k += 1;
    }
    const double dk = (double)k;
    const double f = TMD_SUB(m, 1.0);
    const double s = TMD_DIV(f, TMD_ADD(2.0, f));
    const double z = TMD_MUL(s, s);
    const double w = TMD_MUL(z, z);
    const double t1 = TMD_MUL(w, TMD_ADD(Lg2, TMD_MUL(w, TMD_ADD(Lg4, TMD_MUL(w, Lg6)))));
    const double t2 =
        TMD_MUL(z, TMD_ADD(Lg1, TMD_MUL(w, TMD_ADD(Lg3, TMD_MUL(w, TMD_ADD(Lg5, TMD_MUL(w, Lg7)))))));
    const double r = TMD_ADD(t2, t1);
    const double hfsq = TMD_MUL(TMD_MUL(0.5, f), f);
    // dk*ln2_hi - ((hfsq - (s*(hfsq+r) + dk*ln2_lo)) - f)
    const double inner = TMD_ADD(TMD_MUL(s, TMD_ADD(hfsq, r)), TMD_MUL(dk, ln2_lo));
    return TMD_SUB(TMD_MUL(dk, ln2_hi), TMD_SUB(TMD_SUB(hfsq, inner), f));
}

// (sin, cos)(2 pi u), u in [0, 1): exact octant split, Taylor on [0, pi/4].
TMD_HD void fixed_sincos_2pi(double u, double &sn, double &cs) {
    const double y = TMD_MUL(8.0, u);  // exact
    const int q = (int)y;              // floor, y >= 0
    double r = TMD_SUB(y, (double)q);  // exact
    if (q & 1) r = TMD_SUB(1.0, r);    // exact
    const double t = TMD_MUL(7.85398163397448278999e-01, r);
    const double z = TMD_MUL(t, t);
    double ps = 2.81145725434552059581e-15;
    ps = TMD_ADD(-7.64716373181981641310e-13, TMD_MUL(z, ps));
    ps = TMD_ADD(1.60590438368216133196e-10, TMD_MUL(z, ps));
    ps = TMD_ADD(-2.50521083854417202239e-08, TMD_MUL(z, ps));
    ps = TMD_ADD(2.75573192239858925110e-06, TMD_MUL(z, ps));
    ps = TMD_ADD(-1.98412698412698412526e-04, TMD_MUL(z, ps));
    ps = TMD_ADD(8.33333333333333321769e-03, TMD_MUL(z, ps));
    ps = TMD_ADD(-1.66666666666666657415e-01, TMD_MUL(z, ps));
    const double st = TMD_ADD(t, TMD_MUL(t, TMD_MUL(z, ps)));
    double pc = -1.56192069685862252482e-16;
    pc = TMD_ADD(4.77947733238738524982e-14, TMD_MUL(z, pc));
    pc = TMD_ADD(-1.14707455977297245072e-11, TMD_MUL(z, pc));
    pc = TMD_ADD(2.08767569878681001532e-09, TMD_MUL(z, pc));
    pc = TMD_ADD(-2.75573192239858882758e-07, TMD_MUL(z, pc));
    pc = TMD_ADD(2.48015873015873015658e-05, TMD_MUL(z, pc));
    pc = TMD_ADD(-1.38888888888888894189e-03, TMD_MUL(z, pc));
    pc = TMD_ADD(4.16666666666666643537e-02, TMD_MUL(z, pc));
    pc = TMD_ADD(-5.00000000000000000000e-01, TMD_MUL(z, pc));
    const double ct = TMD_ADD(1.0, TMD_MUL(z, pc));
    const bool swap = ((q + 1) & 2) != 0;
    const double smag = swap ? ct : st;
    const double cmag = swap ? st : ct;
    sn = (q >= 4) ? -smag : smag;
    cs = ((q + 2) & 4) ? -cmag : cmag;
}

// Box-Muller pair from two raw words: z0 = normal #2P, z1 = normal #2P+1.
TMD_HD void normal_pair_from_raw(uint64_t r0, uint64_t r1, double &z0, double &z1) {
    const double two_m53 = 1.1102230246251565404e-16;  // 2^-53
    const double u1 = TMD_MUL((double)((r0 >> 11) + 1ull), two_m53);
    const double u2 = TMD_MUL((double)(r1 >> 11), two_m53);
    const double rad = TMD_SQRT(TMD_MUL(-2.0, fixed_log(u1)));
    double sn, cs;
    fixed_sincos_2pi(u2, sn, cs);
    z0 = TMD_MUL(rad, cs);
    z1 = TMD_MUL(rad, sn);
}

// Normal pair P of the stream with key `key`.
TMD_HD void normal_pair(uint64_t key, uint64_t pair, double &z0, double &z1) {
    Philox4 b = philox4x64_10(pair >> 1, key);
    const int h = (int)(pair & 1ull) * 2;
    normal_pair_from_raw(b.w[h], b.w[h + 1], z0, z1);
}

// Normal number q of the stream.
TMD_HD double normal_at(uint64_t key, uint64_t q) {
    double z0, z1;
    normal_pair(key, q >> 1, z0, z1);
    return (q & 1ull) ? z1 : z0;
}

}  // namespace tmd
