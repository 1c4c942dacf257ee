// K1: tiled all-pairs FP64 Lennard-Jones with minimum-image periodic boundaries.
//
// Replaces LennardJonesCut.potential_and_force / force / potential
// (turtlemd/potentials/lennardjones.py:145-273) and the Box.pbc_dist_matrix call inside their row
// loop (turtlemd/system/box.py:384-403).  Every pair of the reference's N(N-1)/2 is distance-checked.
//
// Structure (FP64-pipe bound; positions of a 32k system are 768 KB and live in L2/shared memory):
//   * grid = (tiles of 128 rows) x (splits of the column range); one row atom per thread
//   * columns are staged through shared memory 256 at a time, read back as broadcasts
//   * FILTER, per pair, on a copy of the coordinates wrapped into the box: per periodic direction
//     |d'| = L/2 - | |d| - L/2 |   (three FP64 adds instead of mul + rint + fma) and
//     rsq' = sum d'^2; the comparison with the cut-off is done on the integer bit patterns so it
//     runs on the ALU pipe.  The filter is conservative (cut-off widened by 1e-9 relative).
//   * pairs that pass are only recorded in a per-thread bit mask; after each tile the warp drains
//     the masks together, so the expensive EXACT path (the reference's own arithmetic on the
//     unwrapped coordinates: '>=' half-box test, rint with the stored reciprocal, strict
//     rsq < rcut2, division, force, virial) runs with many lanes active instead of 1-in-600
//   * per-thread force / energy / virial accumulators; per-split partial forces are summed in a
//     fixed order by the finishing pass -> no atomics, bit-reproducible runs.
#include <stdlib.h>
#include <string.h>

#include "lj_common.cuh"

namespace tmd {

constexpr int kRows = 128;   // threads per block = row atoms per block
constexpr int kTile = 256;   // column atoms staged per iteration
constexpr int kWords = kTile / 32;

struct AllPairsArgs {
    const double *pos;
    const int32_t *tidx;
    const double4 *wrapped;
    const float4 *wrapped_f;  // single-precision copy for the FP32 filter (fully periodic boxes)
    float half_f[3];
    float rc2_f;
    int64_t n, first, count;
    int n_itile, n_split;
    int64_t split_len;
    LjBox box;
    int ntypes;
    uint32_t flags;
    unsigned long long rc2_filter_bits;
    double *partial_f;
    double *partial_c;       // symmetric kernel: column (reaction) forces, [n_itile][n][DIM]
    double *partial_vw;
    LjTable tab;
};

template <int DIM>
__global__ void __launch_bounds__(256) lj_wrap_kernel(const double *__restrict__ pos, int64_t n, LjBox box,
                                                      int pmask, double4 *__restrict__ wrapped,
                                                      float4 *__restrict__ wrapped_f) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double x[3] = {0.0, 0.0, 0.0};
#pragma unroll
    for (int k = 0; k < DIM; ++k) {
        double v = pos[i * DIM + k];
        if (pmask & (1 << k)) v = fma(-rint(v * box.ilen[k]), box.len[k], v);  // into [-L/2, L/2]
        x[k] = v;
    }
    wrapped[i] = make_double4(x[0], x[1], x[2], 0.0);
    if (wrapped_f) wrapped_f[i] = make_float4((float)x[0], (float)x[1], (float)x[2], 0.f);
}

template <int DIM, int PMASK>
__device__ __forceinline__ double filter_rsq(const double (&xw)[3], const double4 q, const double (&h)[3]) {
    double d = xw[0] - q.x;
    if (PMASK & 1) d = h[0] - fabs(fabs(d) - h[0]);
    double r = d * d;
    if (DIM > 1) {
        d = xw[1] - q.y;
        if (PMASK & 2) d = h[1] - fabs(fabs(d) - h[1]);
        r = fma(d, d, r);
    }
    if (DIM > 2) {
        d = xw[2] - q.z;
        if (PMASK & 4) d = h[2] - fabs(fabs(d) - h[2]);
        r = fma(d, d, r);
    }
    return r;
}

// The same filter in single precision (fully periodic boxes, coordinates wrapped into [-L/2, L/2]):
// twelve FP32 instructions per pair on the FMA pipe, which issues twice as many lanes per clock as
// the FP64 pipe and leaves the latter to the exact path.  Conservative: the threshold is widened by
// the single-precision rounding of the wrapped coordinates.
template <int DIM>
__device__ __forceinline__ float filter_rsq_f(const float (&xw)[3], const float4 q, const float (&h)[3]) {
    float d = xw[0] - q.x;
    d = h[0] - fabsf(fabsf(d) - h[0]);
    float r = d * d;
    if (DIM > 1) {
        d = xw[1] - q.y;
        d = h[1] - fabsf(fabsf(d) - h[1]);
        r = fmaf(d, d, r);
    }
    if (DIM > 2) {
        d = xw[2] - q.z;
        d = h[2] - fabsf(fabsf(d) - h[2]);
        r = fmaf(d, d, r);
    }
    return r;
}

template <int DIM, int PMASK, bool F32>
__global__ void __launch_bounds__(kRows, 4) lj_allpairs_kernel(const AllPairsArgs a) {
    __shared__ double4 s_w[F32 ? 1 : kTile];
    __shared__ float4 s_wf[F32 ? kTile : 1];
    __shared__ double s_o[kTile * DIM];
    __shared__ int s_t[kTile];
    __shared__ unsigned s_mask[kWords][kRows];
    __shared__ double s_tab[6 * kMaxT2];
    __shared__ double s_red[10 * (kRows / 32)];

    const int tid = threadIdx.x;
    const int itile = blockIdx.x % a.n_itile;
    const int split = blockIdx.x / a.n_itile;
    const int64_t row = (int64_t)itile * kRows + tid;
    const bool valid = row < a.count;
    const int64_t i = a.first + (valid ? row : 0);
    const int ntypes = a.ntypes;

    if (ntypes > 1) {
        for (int k = tid; k < 6 * kMaxT2; k += kRows) s_tab[k] = a.tab.c[k / kMaxT2][k % kMaxT2];
    }

    double xw[3], xo[3] = {0.0, 0.0, 0.0};
    {
        const double4 w = a.wrapped[i];
        xw[0] = w.x;
        xw[1] = w.y;
        xw[2] = w.z;
#pragma unroll
        for (int k = 0; k < DIM; ++k) xo[k] = a.pos[i * DIM + k];
    }
    const int ti = (ntypes > 1) ? a.tidx[i] : 0;
    const double h[3] = {a.box.half[0], a.box.half[1], a.box.half[2]};
    const long long rc2_bits = valid ? (long long)a.rc2_filter_bits : -1;  // invalid rows never hit
    float xf[3] = {0.f, 0.f, 0.f};
    const float hf[3] = {a.half_f[0], a.half_f[1], a.half_f[2]};
    const float rc2_f = valid ? a.rc2_f : -1.f;
    if (F32) {
        const float4 w = a.wrapped_f[i];
        xf[0] = w.x;
        xf[1] = w.y;
        xf[2] = w.z;
    }

    PairAcc acc;
#pragma unroll
    for (int k = 0; k < 3; ++k) acc.f[k] = 0.0;
    acc.v = 0.0;
#pragma unroll
    for (int k = 0; k < 9; ++k) acc.w[k] = 0.0;

    const int64_t j_begin = (int64_t)split * a.split_len;
    int64_t j_end = j_begin + a.split_len;
    if (j_end > a.n) j_end = a.n;

    for (int64_t jt = j_begin; jt < j_end; jt += kTile) {
        __syncthreads();
#pragma unroll
        for (int s = tid; s < kTile; s += kRows) {
            const int64_t j = jt + s;
            if (j < j_end) {
                if (F32) s_wf[s] = a.wrapped_f[j];
                else s_w[s] = a.wrapped[j];
#pragma unroll
                for (int k = 0; k < DIM; ++k) s_o[s * DIM + k] = a.pos[j * DIM + k];
                s_t[s] = (ntypes > 1) ? a.tidx[j] : 0;
            } else if (F32) {
                s_wf[s] = make_float4(1e18f, 1e18f, 1e18f, 0.f);  // rsq' ~ h^2 at most, but see the mask below
            } else {
                s_w[s] = make_double4(1e300, 1e300, 1e300, 0.0);  // rsq' = inf: never passes
            }
        }
        __syncthreads();

        // ---- filter: 12 instructions per pair in 3-D periodic (FP64, or FP32 for fully periodic boxes) ----
        const int64_t in_tile = j_end - jt;  // columns of this tile that exist
#pragma unroll 1
        for (int word = 0; word < kWords; ++word) {
            unsigned m = 0;
#pragma unroll
            for (int b = 0; b < 32; ++b) {
                if (F32) {
                    if (filter_rsq_f<DIM>(xf, s_wf[word * 32 + b], hf) < rc2_f) m |= 1u << b;
                } else {
                    const double r = filter_rsq<DIM, PMASK>(xw, s_w[word * 32 + b], h);
                    if (__double_as_longlong(r) < rc2_bits) m |= 1u << b;
                }
            }
            // the folding h - ||d| - h| maps any far-away pad value back into the box: mask the columns
            // of a partial last tile that do not exist
            if (F32 && in_tile < (int64_t)(word + 1) * 32)
                m &= in_tile <= (int64_t)word * 32 ? 0u : (0xffffffffu >> (32 - (int)(in_tile - (int64_t)word * 32)));
            s_mask[word][tid] = m;
        }

        // ---- drain: exact reference arithmetic for the recorded pairs ----
#pragma unroll 1
        for (int word = 0; word < kWords; ++word) {
            unsigned m = s_mask[word][tid];
            while (__any_sync(0xffffffffu, m != 0)) {
                if (m != 0) {
                    const int b = __ffs(m) - 1;
                    m &= m - 1;
                    const int s = word * 32 + b;
                    if (jt + s != i) {
                        double d[3] = {0.0, 0.0, 0.0};
                        double rsq = 0.0;
#pragma unroll
                        for (int k = 0; k < DIM; ++k) {
                            double dk = __dsub_rn(xo[k], s_o[s * DIM + k]);  // lennardjones.py:244
                            if (PMASK & (1 << k)) dk = min_image(dk, a.box.len[k], a.box.ilen[k], a.box.half[k]);
                            d[k] = dk;
                            rsq = (k == 0) ? __dmul_rn(dk, dk) : __dadd_rn(rsq, __dmul_rn(dk, dk));  // :246
                        }
                        double lj1, lj2, lj3, lj4, off, rc2;
                        if (ntypes > 1) {
                            const int p = ti * ntypes + s_t[s];
                            lj1 = s_tab[0 * kMaxT2 + p];
                            lj2 = s_tab[1 * kMaxT2 + p];
                            lj3 = s_tab[2 * kMaxT2 + p];
                            lj4 = s_tab[3 * kMaxT2 + p];
                            off = s_tab[4 * kMaxT2 + p];
                            rc2 = s_tab[5 * kMaxT2 + p];
                        } else {
                            lj1 = a.tab.c[0][0];
                            lj2 = a.tab.c[1][0];
                            lj3 = a.tab.c[2][0];
                            lj4 = a.tab.c[3][0];
                            off = a.tab.c[4][0];
                            rc2 = a.tab.c[5][0];
                        }
                        if (rsq < rc2)  // strict '<'  (:291-294)
                            lj_pair<DIM>(d, rsq, lj1, lj2, lj3, lj4, off, a.flags, acc);
                    }
                }
            }
        }
    }

    if (valid && (a.flags & TMD_WANT_F)) {
        double *dst = a.partial_f + ((size_t)split * a.count + row) * DIM;
#pragma unroll
        for (int k = 0; k < DIM; ++k) dst[k] = acc.f[k];
    }
    double red[10];
    red[0] = acc.v;
#pragma unroll
    for (int k = 0; k < 9; ++k) red[1 + k] = acc.w[k];
    block_sum<10, kRows>(red, s_red);
    if (tid == 0) {
#pragma unroll
        for (int k = 0; k < 10; ++k) a.partial_vw[(size_t)blockIdx.x * 10 + k] = red[k];
    }
}

// Symmetric variant (whole system only): every unordered pair is FILTERED ONCE, by the row with the
// smaller index -- exactly the reference's loop structure (lennardjones.py:242-244).  The hit bit
// matrix of a tile is then read twice: by rows (thread = row atom i: energy, virial, F_i += f) and,
// after a ballot transposition of the few non-empty columns, by columns (thread = column atom j:
// F_j -= f, lennardjones.py:271).  Hits are ~0.2 % of the pairs, so evaluating them twice costs
// nothing next to a filter that now does half the work.  Column forces of each (row tile, column)
// go to their own slot of partial_c and are summed in a fixed order by lj_finish_sym_kernel: no
// atomics, bit-reproducible.
template <int DIM, int PMASK, bool F32>
__global__ void __launch_bounds__(kRows, 4) lj_allpairs_sym_kernel(const AllPairsArgs a) {
    __shared__ double4 s_w[F32 ? 1 : kTile];
    __shared__ float4 s_wf[F32 ? kTile : 1];
    __shared__ double s_o[kTile * DIM];
    __shared__ int s_t[kTile];
    __shared__ unsigned s_mask[kWords][kRows];
    __shared__ unsigned s_colmask[kRows / 32][kTile];
    __shared__ double s_row[kRows * DIM];
    __shared__ int s_rowt[kRows];
    __shared__ double s_tab[6 * kMaxT2];
    __shared__ double s_red[10 * (kRows / 32)];

    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int itile = blockIdx.x % a.n_itile;
    const int split = blockIdx.x / a.n_itile;
    const int64_t row0 = (int64_t)itile * kRows;
    const int64_t row = row0 + tid;
    const bool valid = row < a.n;
    const int64_t i = valid ? row : 0;
    const int ntypes = a.ntypes;

    if (ntypes > 1) {
        for (int k = tid; k < 6 * kMaxT2; k += kRows) s_tab[k] = a.tab.c[k / kMaxT2][k % kMaxT2];
    }

    double xw[3], xo[3] = {0.0, 0.0, 0.0};
    {
        const double4 w = a.wrapped[i];
        xw[0] = w.x;
        xw[1] = w.y;
        xw[2] = w.z;
#pragma unroll
        for (int k = 0; k < DIM; ++k) {
            xo[k] = a.pos[i * DIM + k];
            s_row[tid * DIM + k] = xo[k];
        }
    }
    const int ti = (ntypes > 1) ? a.tidx[i] : 0;
    s_rowt[tid] = ti;
    const double h[3] = {a.box.half[0], a.box.half[1], a.box.half[2]};
    const long long rc2_bits = valid ? (long long)a.rc2_filter_bits : -1;  // invalid rows never hit
    float xf[3] = {0.f, 0.f, 0.f};
    const float hf[3] = {a.half_f[0], a.half_f[1], a.half_f[2]};
    const float rc2_f = valid ? a.rc2_f : -1.f;
    if (F32) {
        const float4 w = a.wrapped_f[i];
        xf[0] = w.x;
        xf[1] = w.y;
        xf[2] = w.z;
    }

    PairAcc acc;
#pragma unroll
    for (int k = 0; k < 3; ++k) acc.f[k] = 0.0;
    acc.v = 0.0;
#pragma unroll
    for (int k = 0; k < 9; ++k) acc.w[k] = 0.0;

    const int64_t j_begin = (int64_t)split * a.split_len;
    int64_t j_end = j_begin + a.split_len;
    if (j_end > a.n) j_end = a.n;

    for (int64_t jt = j_begin; jt < j_end; jt += kTile) {
        if (jt + kTile <= row0) continue;  // every column of this tile has a smaller index than every row: not ours
        __syncthreads();
#pragma unroll
        for (int s = tid; s < kTile; s += kRows) {
            const int64_t j = jt + s;
            if (j < j_end) {
                if (F32) s_wf[s] = a.wrapped_f[j];
                else s_w[s] = a.wrapped[j];
#pragma unroll
                for (int k = 0; k < DIM; ++k) s_o[s * DIM + k] = a.pos[j * DIM + k];
                s_t[s] = (ntypes > 1) ? a.tidx[j] : 0;
            } else if (F32) {
                s_wf[s] = make_float4(1e18f, 1e18f, 1e18f, 0.f);
            } else {
                s_w[s] = make_double4(1e300, 1e300, 1e300, 0.0);
            }
#pragma unroll
            for (int w = 0; w < kRows / 32; ++w) s_colmask[w][s] = 0u;
        }
        __syncthreads();

        // ---- filter (pairs i < j only) ----
        const int64_t in_tile = j_end - jt;  // columns of this tile that exist
        const int64_t above = i - jt;        // columns 0 .. above of this tile have index <= i
#pragma unroll 1
        for (int word = 0; word < kWords; ++word) {
            unsigned m = 0;
#pragma unroll
            for (int b = 0; b < 32; ++b) {
                if (F32) {
                    if (filter_rsq_f<DIM>(xf, s_wf[word * 32 + b], hf) < rc2_f) m |= 1u << b;
                } else {
                    const double r = filter_rsq<DIM, PMASK>(xw, s_w[word * 32 + b], h);
                    if (__double_as_longlong(r) < rc2_bits) m |= 1u << b;
                }
            }
            if (in_tile < (int64_t)(word + 1) * 32)  // partial last tile: drop the columns that do not exist
                m &= in_tile <= (int64_t)word * 32 ? 0u : (0xffffffffu >> (32 - (int)(in_tile - (int64_t)word * 32)));
            if (above >= (int64_t)word * 32)         // diagonal tile: keep j > i only
                m &= above >= (int64_t)word * 32 + 31 ? 0u : (0xffffffffu << ((int)(above - (int64_t)word * 32) + 1));
            s_mask[word][tid] = m;
            // transpose the non-empty columns of this word for the column pass
            unsigned any = __reduce_or_sync(0xffffffffu, m);
            while (any) {
                const int b = __ffs(any) - 1;
                any &= any - 1;
                const unsigned rows = __ballot_sync(0xffffffffu, (m >> b) & 1u);
                if (lane == 0) s_colmask[warp][word * 32 + b] = rows;
            }
        }

        // ---- row pass: exact reference arithmetic for the recorded pairs (V, W, F_i) ----
#pragma unroll 1
        for (int word = 0; word < kWords; ++word) {
            unsigned m = s_mask[word][tid];
            while (__any_sync(0xffffffffu, m != 0)) {
                if (m != 0) {
                    const int b = __ffs(m) - 1;
                    m &= m - 1;
                    const int s = word * 32 + b;
                    double d[3] = {0.0, 0.0, 0.0};
                    double rsq = 0.0;
#pragma unroll
                    for (int k = 0; k < DIM; ++k) {
                        double dk = __dsub_rn(xo[k], s_o[s * DIM + k]);  // lennardjones.py:244
                        if (PMASK & (1 << k)) dk = min_image(dk, a.box.len[k], a.box.ilen[k], a.box.half[k]);
                        d[k] = dk;
                        rsq = (k == 0) ? __dmul_rn(dk, dk) : __dadd_rn(rsq, __dmul_rn(dk, dk));  // :246
                    }
                    const int p = (ntypes > 1) ? ti * ntypes + s_t[s] : 0;
                    const double *tab = (ntypes > 1) ? s_tab : nullptr;
                    const double lj1 = tab ? tab[0 * kMaxT2 + p] : a.tab.c[0][0], lj2 = tab ? tab[1 * kMaxT2 + p] : a.tab.c[1][0],
                                 lj3 = tab ? tab[2 * kMaxT2 + p] : a.tab.c[2][0], lj4 = tab ? tab[3 * kMaxT2 + p] : a.tab.c[3][0],
                                 off = tab ? tab[4 * kMaxT2 + p] : a.tab.c[4][0], rc2 = tab ? tab[5 * kMaxT2 + p] : a.tab.c[5][0];
                    if (rsq < rc2)  // strict '<'  (:291-294)
                        lj_pair<DIM>(d, rsq, lj1, lj2, lj3, lj4, off, a.flags, acc);
                }
            }
        }
        __syncthreads();  // all column masks are in place

        // ---- column pass: F_j -= f for the same pairs, thread = column atom ----
        if (a.flags & TMD_WANT_F) {
#pragma unroll 1
            for (int c = tid; c < kTile; c += kRows) {
                double fc[3] = {0.0, 0.0, 0.0};
                const int tj = (ntypes > 1) ? s_t[c] : 0;
#pragma unroll 1
                for (int w = 0; w < kRows / 32; ++w) {
                    unsigned rows = s_colmask[w][c];
                    while (rows) {
                        const int r = w * 32 + __ffs(rows) - 1;
                        rows &= rows - 1;
                        double d[3] = {0.0, 0.0, 0.0};
                        double rsq = 0.0;
#pragma unroll
                        for (int k = 0; k < DIM; ++k) {
                            double dk = __dsub_rn(s_row[r * DIM + k], s_o[c * DIM + k]);  // the same x_i - x_j
                            if (PMASK & (1 << k)) dk = min_image(dk, a.box.len[k], a.box.ilen[k], a.box.half[k]);
                            d[k] = dk;
                            rsq = (k == 0) ? __dmul_rn(dk, dk) : __dadd_rn(rsq, __dmul_rn(dk, dk));
                        }
                        const int p = (ntypes > 1) ? s_rowt[r] * ntypes + tj : 0;
                        const double *tab = (ntypes > 1) ? s_tab : nullptr;
                        const double lj1 = tab ? tab[0 * kMaxT2 + p] : a.tab.c[0][0], lj2 = tab ? tab[1 * kMaxT2 + p] : a.tab.c[1][0],
                                     rc2 = tab ? tab[5 * kMaxT2 + p] : a.tab.c[5][0];
                        if (rsq < rc2) {
                            const double r2inv = __ddiv_rn(1.0, rsq);
                            const double r6inv = cube_cr(r2inv);
                            const double flj = __dmul_rn(__dmul_rn(r2inv, r6inv), __dsub_rn(__dmul_rn(lj1, r6inv), lj2));
#pragma unroll
                            for (int k = 0; k < DIM; ++k) fc[k] -= __dmul_rn(flj, d[k]);  // :271
                        }
                    }
                }
                const int64_t j = jt + c;
                if (j < j_end) {
                    double *dst = a.partial_c + ((size_t)itile * a.n + j) * DIM;
#pragma unroll
                    for (int k = 0; k < DIM; ++k) dst[k] = fc[k];
                }
            }
        }
    }

    if (valid && (a.flags & TMD_WANT_F)) {
        double *dst = a.partial_f + ((size_t)split * a.n + row) * DIM;
#pragma unroll
        for (int k = 0; k < DIM; ++k) dst[k] = acc.f[k];
    }
    double red[10];
    red[0] = acc.v;
#pragma unroll
    for (int k = 0; k < 9; ++k) red[1 + k] = acc.w[k];
    block_sum<10, kRows>(red, s_red);
    if (tid == 0) {
#pragma unroll
        for (int k = 0; k < 10; ++k) a.partial_vw[(size_t)blockIdx.x * 10 + k] = red[k];
    }
}

// force[j] (=|+=) sum over splits of partial_f[s][j] + sum over the row tiles I that processed
// column j (all tiles whose first row is not beyond the end of j's column tile) of partial_c[I][j].
__global__ void __launch_bounds__(256) lj_finish_sym_kernel(const double *__restrict__ partial_f,
                                                            const double *__restrict__ partial_c, int n_split,
                                                            int n_itile, int64_t n, int dim, uint32_t flags,
                                                            double *__restrict__ force) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t total = n * dim;
    if (e >= total) return;
    const int64_t j = e / dim;
    double s = partial_f[e];
    for (int p = 1; p < n_split; ++p) s += partial_f[(size_t)p * total + e];
    int64_t last = ((j / kTile) * kTile + kTile - 1) / kRows;  // last row tile that starts inside or before j's column tile
    if (last > n_itile - 1) last = n_itile - 1;
    for (int64_t it = 0; it <= last; ++it) s += partial_c[(size_t)it * total + e];
    force[e] = (flags & TMD_ACCUMULATE) ? force[e] + s : s;
}

__global__ void __launch_bounds__(256) lj_finish_kernel(const double *__restrict__ partial_f, int n_split,
                                                        int64_t total, int64_t offset, uint32_t flags,
                                                        double *__restrict__ force) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= total) return;
    double s = partial_f[e];
    for (int p = 1; p < n_split; ++p) s += partial_f[(size_t)p * total + e];
    double *dst = force + offset + e;
    *dst = (flags & TMD_ACCUMULATE) ? *dst + s : s;
}

int launch_lj_finish(const double *partial_f, int n_split, int64_t count, int dim, int64_t first,
                     uint32_t flags, double *force, cudaStream_t stream) {
    const int64_t total = count * dim;
    if (total == 0) return TMD_OK;
    const int blocks = (int)((total + 255) / 256);
    ::tmd::count_launch(), lj_finish_kernel<<<blocks, 256, 0, stream>>>(partial_f, n_split, total, first * dim, flags, force);
    TMD_CUDA(cudaGetLastError());
    return TMD_OK;
}

int load_lj_table(const double *table, int ntypes, LjTable &out, double &rc2max) {
    if (!table || ntypes < 1) {
        set_error("LJ table is NULL or ntypes < 1");
        return TMD_ERR_INVALID;
    }
    if (ntypes > TMD_MAX_TYPES) {
        set_error("%d particle types exceed TMD_MAX_TYPES=%d of the device path", ntypes, TMD_MAX_TYPES);
        return TMD_ERR_UNSUPPORTED;
    }
    rc2max = 0.0;
    for (int c = 0; c < 6; ++c)
        for (int p = 0; p < kMaxT2; ++p) out.c[c][p] = 0.0;
    for (int c = 0; c < 6; ++c)
        for (int p = 0; p < ntypes * ntypes; ++p) out.c[c][p] = table[(size_t)c * ntypes * ntypes + p];
    for (int p = 0; p < ntypes * ntypes; ++p) {
        const double rc2 = out.c[5][p];
        if (rc2 == rc2 && rc2 > rc2max) rc2max = rc2;  // NaN marks pairs the reference has no entry for
    }
    return TMD_OK;
}

template <int DIM>
static int run_allpairs(AllPairsArgs &a, int pmask, bool f32, int blocks, cudaStream_t s) {
    if (f32) {
        ::tmd::count_launch(), lj_allpairs_kernel<DIM, (1 << DIM) - 1, true><<<blocks, kRows, 0, s>>>(a);
        TMD_CUDA(cudaGetLastError());
        return TMD_OK;
    }
#define TMD_AP(P)                                                                        \
    case P:                                                                              \
        if constexpr ((P) < (1 << DIM)) ::tmd::count_launch(), lj_allpairs_kernel<DIM, P, false><<<blocks, kRows, 0, s>>>(a); \
        break;
    switch (pmask & ((1 << DIM) - 1)) {
        TMD_AP(0)
        TMD_AP(1)
        TMD_AP(2)
        TMD_AP(3)
        TMD_AP(4)
        TMD_AP(5)
        TMD_AP(6)
        TMD_AP(7)
        default: break;
    }
#undef TMD_AP
    TMD_CUDA(cudaGetLastError());
    return TMD_OK;
}

template <int DIM>
static int run_allpairs_sym(AllPairsArgs &a, int pmask, bool f32, int blocks, cudaStream_t s) {
    if (f32) {
        ::tmd::count_launch(), lj_allpairs_sym_kernel<DIM, (1 << DIM) - 1, true><<<blocks, kRows, 0, s>>>(a);
        TMD_CUDA(cudaGetLastError());
        return TMD_OK;
    }
#define TMD_AP(P)                                                                        \
    case P:                                                                              \
        if constexpr ((P) < (1 << DIM)) ::tmd::count_launch(), lj_allpairs_sym_kernel<DIM, P, false><<<blocks, kRows, 0, s>>>(a); \
        break;
    switch (pmask & ((1 << DIM) - 1)) {
        TMD_AP(0)
        TMD_AP(1)
        TMD_AP(2)
        TMD_AP(3)
        TMD_AP(4)
        TMD_AP(5)
        TMD_AP(6)
        TMD_AP(7)
        default: break;
    }
#undef TMD_AP
    TMD_CUDA(cudaGetLastError());
    return TMD_OK;
}

}  // namespace tmd

using namespace tmd;

extern "C" int tmd_lj_allpairs(const double *pos, const int32_t *tidx, int64_t n, const tmd_box *box,
                               const double *table, int32_t ntypes, uint32_t flags, int64_t first,
                               int64_t count, double *force, double *vw, tmd_stream_t stream) {
    TMD_REQUIRE(box && box->dim >= 1 && box->dim <= 3, "box->dim must be 1..3");
    TMD_REQUIRE(n >= 0 && first >= 0 && count >= 0 && first + count <= n, "rows [first, first+count) must lie in [0, n)");
    TMD_REQUIRE(pos || n == 0, "pos is NULL");
    TMD_REQUIRE(ntypes == 1 || tidx, "tidx is required when ntypes > 1");
    TMD_REQUIRE(!(flags & TMD_WANT_F) || force, "force is NULL but TMD_WANT_F is set");
    TMD_REQUIRE(vw != nullptr, "vw is NULL");
    TMD_CHECK(require_device());
    cudaStream_t s = as_stream(stream);
    const int dim = box->dim;

    AllPairsArgs a;
    double rc2max = 0.0;
    TMD_CHECK(load_lj_table(table, ntypes, a.tab, rc2max));
    int pmask = 0;
    make_lj_box(box, a.box, pmask);

    if (count == 0 || n == 0)  // nothing to compute: the wanted outputs become 0 (+ old value)
        return launch_reduce_vw(nullptr, 0, 0.5, flags, vw, s);

    a.pos = pos;
    a.tidx = tidx;
    a.n = n;
    a.first = first;
    a.count = count;
    a.ntypes = ntypes;
    a.flags = flags;
    a.n_itile = (int)((count + kRows - 1) / kRows);
    // enough blocks for ~6 waves of 4 resident blocks per SM, column ranges in whole tiles
    const int64_t n_jtile = (n + kTile - 1) / kTile;
    // (the symmetric kernel leaves the blocks below the diagonal idle: twice as many, so that the busy
    // half still makes ~6 waves; TMD_AP_WAVES overrides)
    static const int waves_env = [] {
        const char *env = getenv("TMD_AP_WAVES");
        return env ? atoi(env) : 0;
    }();
    const bool whole_system = first == 0 && count == n;
    const int waves = waves_env > 0 ? waves_env : (whole_system ? 12 : 6);
    int64_t want_split = (int64_t)num_sms() * 4 * waves / a.n_itile;
    if (want_split < 1) want_split = 1;
    if (want_split > n_jtile) want_split = n_jtile;
    const int64_t tiles_per_split = (n_jtile + want_split - 1) / want_split;
    a.split_len = tiles_per_split * kTile;
    a.n_split = (int)((n_jtile + tiles_per_split - 1) / tiles_per_split);
    // conservative filter cut-off; compared as integers, both sides are non-negative doubles
    const double rc2f = rc2max * (1.0 + 1e-9);
    memcpy(&a.rc2_filter_bits, &rc2f, sizeof(double));

    const int blocks = a.n_itile * a.n_split;
    double4 *wrapped = nullptr;
    TMD_CHECK(arena_get(SLOT_WRAPPED, (size_t)n * sizeof(double4), (void **)&wrapped));
    TMD_CHECK(arena_get(SLOT_PARTIAL_F, (size_t)a.n_split * count * dim * sizeof(double), (void **)&a.partial_f));
    TMD_CHECK(arena_get(SLOT_PARTIAL_VW, (size_t)blocks * 10 * sizeof(double), (void **)&a.partial_vw));
    a.wrapped = wrapped;

    // FP32 filter when every direction is periodic (TMD_AP_FILTER=fp64 keeps the FP64 filter): the
    // wrapped coordinates are bounded by L/2, so their single-precision images are off by at most
    // 2^-24 L/2 each; with the roundings of the twelve operations the distance is off by < 2^-20 Lmax.
    bool f32 = false;
    float4 *wrapped_f = nullptr;
    a.wrapped_f = nullptr;
    a.rc2_f = 0.f;
    a.half_f[0] = a.half_f[1] = a.half_f[2] = 0.f;
    static const bool fp64_only = [] {
        const char *env = getenv("TMD_AP_FILTER");
        return env && strcmp(env, "fp64") == 0;
    }();
    if (!fp64_only && pmask == (1 << dim) - 1 && rc2max > 0.0) {
        double lmax = 0.0;
        for (int k = 0; k < dim; ++k) lmax = fmax(lmax, box->length[k]);
        const double reach = sqrt(rc2max) * (1.0 + 1e-6) + lmax * (1.0 / 1048576.0);
        a.rc2_f = (float)(reach * reach * (1.0 + 1e-6));
        for (int k = 0; k < dim; ++k) a.half_f[k] = (float)a.box.half[k];
        TMD_CHECK(arena_get(SLOT_CELL_KEYS, (size_t)n * sizeof(float4), (void **)&wrapped_f));
        a.wrapped_f = wrapped_f;
        f32 = true;
    }

    const int wblocks = (int)((n + 255) / 256);
    if (dim == 1) ::tmd::count_launch(), lj_wrap_kernel<1><<<wblocks, 256, 0, s>>>(pos, n, a.box, pmask, wrapped, wrapped_f);
    else if (dim == 2) ::tmd::count_launch(), lj_wrap_kernel<2><<<wblocks, 256, 0, s>>>(pos, n, a.box, pmask, wrapped, wrapped_f);
    else ::tmd::count_launch(), lj_wrap_kernel<3><<<wblocks, 256, 0, s>>>(pos, n, a.box, pmask, wrapped, wrapped_f);
    TMD_CUDA(cudaGetLastError());

    // Whole system: the symmetric kernel (each pair filtered once).  A block of rows (atom
    // decomposition) needs every column for its own rows and keeps the full-row kernel.
    static const bool no_sym = [] {
        const char *env = getenv("TMD_AP_SYMMETRIC");
        return env && strcmp(env, "0") == 0;
    }();
    const size_t c_bytes = (size_t)a.n_itile * n * dim * sizeof(double);
    if (!no_sym && first == 0 && count == n && c_bytes <= ((size_t)1 << 30)) {
        TMD_CHECK(arena_get(SLOT_PARTIAL_C, c_bytes, (void **)&a.partial_c));
        if (dim == 1) TMD_CHECK(run_allpairs_sym<1>(a, pmask, f32, blocks, s));
        else if (dim == 2) TMD_CHECK(run_allpairs_sym<2>(a, pmask, f32, blocks, s));
        else TMD_CHECK(run_allpairs_sym<3>(a, pmask, f32, blocks, s));
        if (flags & TMD_WANT_F) {
            const int64_t total = n * dim;
            ::tmd::count_launch(), lj_finish_sym_kernel<<<(int)((total + 255) / 256), 256, 0, s>>>(
                a.partial_f, a.partial_c, a.n_split, a.n_itile, n, dim, flags, force);
            TMD_CUDA(cudaGetLastError());
        }
        return launch_reduce_vw(a.partial_vw, blocks, 1.0, flags, vw, s);  // every pair was visited once
    }
    a.partial_c = nullptr;

    if (dim == 1) TMD_CHECK(run_allpairs<1>(a, pmask, f32, blocks, s));
    else if (dim == 2) TMD_CHECK(run_allpairs<2>(a, pmask, f32, blocks, s));
    else TMD_CHECK(run_allpairs<3>(a, pmask, f32, blocks, s));

    if (flags & TMD_WANT_F) TMD_CHECK(launch_lj_finish(a.partial_f, a.n_split, count, dim, first, flags, force, s));
    // every pair was visited from both of its rows
    return launch_reduce_vw(a.partial_vw, blocks, 0.5, flags, vw, s);
}
