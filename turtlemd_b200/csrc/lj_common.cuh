// Pieces shared by the two Lennard-Jones kernels (all-pairs and cell list): the type table, the
// exact per-pair arithmetic of the reference and the finishing pass.
#pragma once

#include "common.cuh"

namespace tmd {

constexpr int kMaxT2 = TMD_MAX_TYPES * TMD_MAX_TYPES;

// lj1, lj2, lj3, lj4, offset, rcut2 per ordered type pair (lennardjones.py:94-115)
struct LjTable {
    double c[6][kMaxT2];
};

struct LjBox {
    double len[3], ilen[3], half[3];
};

static inline void make_lj_box(const tmd_box *box, LjBox &b, int &pmask) {
    pmask = 0;
    for (int k = 0; k < 3; ++k) {
        const bool on = k < box->dim && box->periodic[k];
        if (on) pmask |= 1 << k;
        b.len[k] = on ? box->length[k] : 0.0;
        b.ilen[k] = on ? box->ilength[k] : 0.0;
        b.half[k] = on ? 0.5 * box->length[k] : 0.0;
    }
}

// Validates the table and copies it into the kernel-parameter struct; returns max rcut2.
int load_lj_table(const double *table, int ntypes, LjTable &out, double &rc2max);

// x^3 carrying the rounding error of x*x: NumPy's r2inv**3 / rsq**3 (lennardjones.py:172,255) is a
// (nearly) correctly rounded pow, a plain (x*x)*x is off by up to ~1.5 ulp.
__device__ __forceinline__ double cube_cr(double x) {
    const double x2 = __dmul_rn(x, x);
    const double e2 = __fma_rn(x, x, -x2);
    const double h3 = __dmul_rn(x2, x);
    const double e3 = __fma_rn(x2, x, -h3);
    return __dadd_rn(h3, __fma_rn(e2, x, e3));
}

// Box.pbc_dist_matrix for one component (box.py:398-402): '>=' half box, rint, stored reciprocal.
__device__ __forceinline__ double min_image(double d, double len, double ilen, double half) {
    if (fabs(d) >= half) d = __dsub_rn(d, __dmul_rn(rint(__dmul_rn(d, ilen)), len));
    return d;
}

struct PairAcc {
    double f[3];
    double v;
    double w[9];
};

// The body of the reference's row loop for ONE pair whose exact rsq passed the cut-off test
// (lennardjones.py:253-272).  d = r_i - r_j after minimum image.
// FAST = false: the all-pairs path; r6inv through a compensated cube (tracks NumPy's pow), the full
//               dim x dim virial accumulated entry by entry like np.einsum("ij,ik->jk") does.
// FAST = true : the cell / neighbour-list paths, where this body IS the hot loop: plain products,
//               and only the upper triangle of the (mathematically symmetric) virial -- the caller
//               mirrors it.  Differences to FAST = false are at the 1e-16 level.
template <int DIM, bool FAST = false>
__device__ __forceinline__ void lj_pair(const double (&d)[3], double rsq, double lj1, double lj2, double lj3,
                                        double lj4, double offset, uint32_t flags, PairAcc &acc) {
    double r2inv, r6inv;
    if (flags & TMD_V_STANDALONE) {
        r2inv = 0.0;
        r6inv = __ddiv_rn(1.0, FAST ? (rsq * rsq) * rsq : cube_cr(rsq));  // potential(): 1.0 / rsq**3  (:172)
    } else {
        r2inv = __ddiv_rn(1.0, rsq);                            // :254
        r6inv = FAST ? (r2inv * r2inv) * r2inv : cube_cr(r2inv);  // :255
    }
    if (flags & TMD_WANT_V)                    // r6inv*(lj3*r6inv - lj4) - offset  (:276-282)
        acc.v += __dsub_rn(__dmul_rn(r6inv, __dsub_rn(__dmul_rn(lj3, r6inv), lj4)), offset);
    if (flags & (TMD_WANT_F | TMD_WANT_W)) {
        // r2inv*r6inv*(lj1*r6inv - lj2)  (:285-288)
        const double flj = __dmul_rn(__dmul_rn(r2inv, r6inv), __dsub_rn(__dmul_rn(lj1, r6inv), lj2));
#pragma unroll
        for (int a = 0; a < DIM; ++a) {
            const double fa = __dmul_rn(flj, d[a]);  // :269
            acc.f[a] += fa;                          // :270
            if (flags & TMD_WANT_W) {
#pragma unroll
                for (int b = (FAST ? a : 0); b < DIM; ++b) acc.w[a * 3 + b] = fma(fa, d[b], acc.w[a * 3 + b]);  // :272
            }
        }
    }
}

// FAST accumulates the upper triangle only: copy it down before reducing.
__device__ __forceinline__ void mirror_virial(PairAcc &acc) {
    acc.w[3] = acc.w[1];
    acc.w[6] = acc.w[2];
    acc.w[7] = acc.w[5];
}

// force[first+r] (=|+=) sum over splits of partial[s][r], fixed order.
int launch_lj_finish(const double *partial_f, int n_split, int64_t count, int dim, int64_t first,
                     uint32_t flags, double *force, cudaStream_t stream);

}  // namespace tmd
