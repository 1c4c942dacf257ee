// Per-coordinate arithmetic of the integrators (turtlemd/integrators.py), shared by the element-wise
// kernels (integrators.cu) and the slot-ordered slab kernels (slab.cu).  Explicit *_rn intrinsics in
// the order of the reference's NumPy expressions.
#pragma once

#include "common.cuh"
#include "philox_normal.h"

namespace tmd {

// v += (half*F)*imass   (integrators.py:88 and :94)
__device__ __forceinline__ double kick(double v, double f, double im, double half) {
    return __dadd_rn(v, __dmul_rn(__dmul_rn(half, f), im));
}

// normals q0 and q0+1 of the stream
__device__ __forceinline__ void two_normals(uint64_t key, uint64_t q0, bool two, double &z0, double &z1) {
    if ((q0 & 1ull) == 0) {
        normal_pair(key, q0 >> 1, z0, z1);
    } else {
        z0 = normal_at(key, q0);
        z1 = two ? normal_at(key, q0 + 1) : 0.0;
    }
}

// normals q0 .. q0 + 3 of the stream, q0 a multiple of four: the four words of ONE Philox block
__device__ __forceinline__ void four_normals(uint64_t key, uint64_t q0, double (&z)[4]) {
    const Philox4 b = philox4x64_10(q0 >> 2, key);
    normal_pair_from_raw(b.w[0], b.w[1], z[0], z[1]);
    normal_pair_from_raw(b.w[2], b.w[3], z[2], z[3]);
}

// normals q0 .. q0 + 5 of the stream, q0 even (the six draws of one particle in three dimensions): they lie in TWO
// consecutive Philox blocks -- words (0,1), (2,3) of the first and (0,1) of the second when q0 is a multiple of four,
// words (2,3) of the first and (0,1), (2,3) of the second otherwise -- instead of the three blocks three pair draws cost
__device__ __forceinline__ void six_normals(uint64_t key, uint64_t q0, double (&z)[6]) {
    const Philox4 a = philox4x64_10(q0 >> 2, key), b = philox4x64_10((q0 >> 2) + 1, key);
    const bool low = (q0 & 2ull) == 0;
    normal_pair_from_raw(low ? a.w[0] : a.w[2], low ? a.w[1] : a.w[3], z[0], z[1]);
    normal_pair_from_raw(low ? a.w[2] : b.w[0], low ? a.w[3] : b.w[1], z[2], z[3]);
    normal_pair_from_raw(low ? b.w[0] : b.w[2], low ? b.w[1] : b.w[3], z[4], z[5]);
}

struct InertiaPerParticle {
    double a2, b1, b2, l00, l10, l11;
};

__device__ __forceinline__ InertiaPerParticle inertia_constants(const tmd_langevin_consts &c, double im) {
    InertiaPerParticle p;
    p.a2 = __dmul_rn(c.a2f, im);  // integrators.py:239-241
    p.b1 = __dmul_rn(c.b1f, im);
    p.b2 = __dmul_rn(c.b2f, im);
    // integrators.py:249-256 and the 2x2 Cholesky of :261
    const double sr2 = __dmul_rn(__ddiv_rn(__dmul_rn(c.dt, im), c.beta_gamma), c.kfac);
    const double sv2 = __ddiv_rn(__dmul_rn(c.one_minus_e2, im), c.beta);
    const double crv = __dmul_rn(__ddiv_rn(im, c.beta_gamma), c.one_minus_e_sq);
    p.l00 = __dsqrt_rn(sr2);
    p.l10 = __ddiv_rn(crv, p.l00);
    p.l11 = __dsqrt_rn(__dsub_rn(sv2, __dmul_rn(p.l10, p.l10)));
    return p;
}

__device__ __forceinline__ void inertia_pre_update(double &x, double &v, double f, double xr, double vr,
                                                   const tmd_langevin_consts &c, const InertiaPerParticle &p) {
    // pos += a1*vel + a2*force + pos_rand ; vel2 = c0*vel + b1*force + vel_rand (integrators.py:270-279)
    const double dx = __dadd_rn(__dadd_rn(__dmul_rn(c.a1, v), __dmul_rn(p.a2, f)), xr);
    const double v2 = __dadd_rn(__dadd_rn(__dmul_rn(c.c0, v), __dmul_rn(p.b1, f)), vr);
    x = __dadd_rn(x, dx);
    v = v2;
}

}  // namespace tmd
