// K4: DoubleWell (one-body, element-wise, HBM-bound) and K5: DoubleWellPair (all pairs of the few
// atoms whose type the potential is active for).  Reference: turtlemd/potentials/well.py.
#include "common.cuh"
#include "doublewell.cuh"
#include "triclinic.cuh"

namespace tmd {

// ---- K4 -------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) doublewell_kernel(const double *__restrict__ pos, int64_t total, double a,
                                                         double b, double c, uint32_t flags,
                                                         double *__restrict__ force,
                                                         double *__restrict__ partials) {
    __shared__ double smem[8];
    double vsum[1] = {0.0};
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
        const double x = pos[e];
        if (flags & TMD_WANT_F) {
            const double f = dw_force(x, a, b, c);
            force[e] = (flags & TMD_ACCUMULATE) ? force[e] + f : f;
        }
        if (flags & TMD_WANT_V) vsum[0] += dw_energy(x, a, b, c);
    }
    block_sum<1, 256>(vsum, smem);
    if (threadIdx.x == 0) {
        double *row = partials + (size_t)blockIdx.x * 10;
        row[0] = vsum[0];
#pragma unroll
        for (int k = 1; k < 10; ++k) row[k] = 0.0;
    }
}

// ---- K5 -------------------------------------------------------------------------------------------
struct BoxDev {
    double len[3], ilen[3], half[3];
    int per[3];
};

static inline BoxDev make_box(const tmd_box *box) {
    BoxDev b;
    for (int k = 0; k < 3; ++k) {
        const bool on = k < box->dim && box->periodic[k];
        b.per[k] = on ? 1 : 0;
        b.len[k] = on ? box->length[k] : 0.0;
        b.ilen[k] = on ? box->ilength[k] : 0.0;
        b.half[k] = on ? 0.5 * box->length[k] : 0.0;
    }
    return b;
}

// One thread per active atom; it walks all its partners in index order and sums its own force,
// half of the pair energies and half of the pair virials: no atomics, fixed order.
// TRI: the box is a TriclinicBox and the minimum image is TriclinicBox.pbc_dist (box.py:271-296).
template <int DIM, bool TRI>
__global__ void __launch_bounds__(128) dwpair_kernel(const double *__restrict__ pos, BoxDev box, TriCell cell,
                                                     const int32_t *__restrict__ idx0, int n0,
                                                     const int32_t *__restrict__ idx1, int n1, int same_type,
                                                     double rwidth, double width2, double height,
                                                     uint32_t flags, double *__restrict__ force,
                                                     double *__restrict__ partials) {
    __shared__ double smem[10 * 4];
    double acc[10];
#pragma unroll
    for (int k = 0; k < 10; ++k) acc[k] = 0.0;
    const int nown = same_type ? n0 : n0 + n1;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < nown) {
        const bool in0 = t < n0;
        const int me = in0 ? idx0[t] : idx1[t - n0];
        const int32_t *other = same_type ? idx0 : (in0 ? idx1 : idx0);
        const int nother = same_type ? n0 : (in0 ? n1 : n0);
        double xi[3] = {0.0, 0.0, 0.0}, fi[3] = {0.0, 0.0, 0.0};
#pragma unroll
        for (int k = 0; k < DIM; ++k) xi[k] = pos[(int64_t)me * DIM + k];
        const double height4 = __dmul_rn(4.0, height);  // well.py:212
        for (int p = 0; p < nother; ++p) {
            const int you = other[p];
            if (you == me) continue;
            const bool first = me < you;  // the reference visits the pair as (i, j) with i < j (particles.py:129-132)
            double d[3] = {0.0, 0.0, 0.0};
            double rsq = 0.0;
#pragma unroll
            for (int k = 0; k < DIM; ++k) {
                const double xj = pos[(int64_t)you * DIM + k];
                double dk = first ? __dsub_rn(xi[k], xj) : __dsub_rn(xj, xi[k]);
                // Box.pbc_dist: strict '>' half box, stored reciprocal (box.py:376-378)
                if (!TRI && box.per[k] && fabs(dk) > box.half[k])
                    dk = __dsub_rn(dk, __dmul_rn(rint(__dmul_rn(dk, box.ilen[k])), box.len[k]));
                d[k] = dk;
            }
            if (TRI) tri_min_image<DIM>(cell, d);
#pragma unroll
            for (int k = 0; k < DIM; ++k) rsq = (k == 0) ? __dmul_rn(d[k], d[k]) : __dadd_rn(rsq, __dmul_rn(d[k], d[k]));
            const double delr = __dsqrt_rn(rsq);
            const double diff = __dsub_rn(delr, rwidth);
            const double q = __ddiv_rn(__dmul_rn(diff, diff), width2);
            if (flags & TMD_WANT_V) {  // height*(1 - diff^2/width2)^2  (well.py:250-260)
                const double om = __dsub_rn(1.0, q);
                acc[0] += __dmul_rn(height, __dmul_rn(om, om));
            }
            if (flags & (TMD_WANT_F | TMD_WANT_W)) {
                // height4*(1 - diff^2/width2)*(diff/width2) * delta/delr  (well.py:279-282)
                const double mag = __dmul_rn(__dmul_rn(height4, __dsub_rn(1.0, q)), __ddiv_rn(diff, width2));
#pragma unroll
                for (int k = 0; k < DIM; ++k) {
                    const double fk = __ddiv_rn(__dmul_rn(mag, d[k]), delr);
                    fi[k] += first ? fk : -fk;
#pragma unroll
                    for (int m = 0; m < DIM; ++m) acc[1 + k * 3 + m] += __dmul_rn(fk, d[m]);  // np.outer (well.py:285)
                }
            }
        }
        if (flags & TMD_WANT_F) {
#pragma unroll
            for (int k = 0; k < DIM; ++k) {
                double *dst = force + (int64_t)me * DIM + k;
                *dst = (flags & TMD_ACCUMULATE) ? *dst + fi[k] : fi[k];
            }
        }
    }
    block_sum<10, 128>(acc, smem);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < 10; ++k) partials[(size_t)blockIdx.x * 10 + k] = acc[k];
    }
}

}  // namespace tmd

using namespace tmd;

extern "C" {

int tmd_doublewell(const double *pos, int64_t n, int32_t dim, double a, double b, double c, uint32_t flags,
                   double *force, double *vw, tmd_stream_t stream) {
    TMD_REQUIRE(dim >= 1 && dim <= 3 && n >= 0, "dim must be 1..3 and n >= 0");
    TMD_REQUIRE(!(flags & TMD_WANT_F) || force, "force is NULL but TMD_WANT_F is set");
    TMD_REQUIRE(vw != nullptr, "vw is NULL");
    TMD_CHECK(require_device());
    cudaStream_t s = as_stream(stream);
    const int64_t total = n * dim;
    int blocks = (int)((total + 255) / 256);
    const int cap = num_sms() * 16;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    double *partials = nullptr;
    TMD_CHECK(arena_get(SLOT_PARTIAL_VW, (size_t)blocks * 10 * sizeof(double), (void **)&partials));
    ::tmd::count_launch(), doublewell_kernel<<<blocks, 256, 0, s>>>(pos, total, a, b, c, flags, force, partials);
    TMD_CUDA(cudaGetLastError());
    // virial = zeros (well.py:80-81): the W partials are all zero, so TMD_WANT_W defines W = 0 (+ old W)
    return launch_reduce_vw(partials, blocks, 1.0, flags, vw, s);
}

static int launch_dwpair(const double *pos, int64_t n, int dim, const tmd_box *box, const tmd_cell *tri,
                         const int32_t *idx0, int32_t n0, const int32_t *idx1, int32_t n1, int32_t same_type,
                         double rwidth, double width2, double height, uint32_t flags, double *force, double *vw,
                         cudaStream_t s) {
    TMD_REQUIRE(n0 >= 0 && n1 >= 0, "negative index count");
    TMD_REQUIRE(!(flags & TMD_WANT_F) || force, "force is NULL but TMD_WANT_F is set");
    TMD_REQUIRE(vw != nullptr, "vw is NULL");
    TMD_CHECK(require_device());
    if ((flags & TMD_WANT_F) && !(flags & TMD_ACCUMULATE))
        TMD_CUDA(cudaMemsetAsync(force, 0, (size_t)n * dim * sizeof(double), s));
    const int nown = same_type ? n0 : n0 + n1;
    int blocks = (nown + 127) / 128;
    if (blocks < 1) blocks = 1;
    double *partials = nullptr;
    TMD_CHECK(arena_get(SLOT_PARTIAL_VW, (size_t)blocks * 10 * sizeof(double), (void **)&partials));
    BoxDev bd = {};
    TriCell cell = {};
    if (tri) TMD_CHECK(load_tri_cell(tri, cell));
    else bd = make_box(box);
    // the kernel adds into force[] (zeroed above when overwriting)
    const uint32_t kflags = flags | TMD_ACCUMULATE;
#define TMD_DWP(D, T) \
    ::tmd::count_launch(), dwpair_kernel<D, T><<<blocks, 128, 0, s>>>(pos, bd, cell, idx0, n0, idx1, n1, same_type, rwidth, width2, height, kflags, force, partials)
    if (tri) { if (dim == 2) TMD_DWP(2, true); else TMD_DWP(3, true); }
    else if (dim == 1) TMD_DWP(1, false);
    else if (dim == 2) TMD_DWP(2, false);
    else TMD_DWP(3, false);
#undef TMD_DWP
    TMD_CUDA(cudaGetLastError());
    // every pair was visited from both ends
    return launch_reduce_vw(partials, blocks, 0.5, flags, vw, s);
}

int tmd_dwpair(const double *pos, int64_t n, const tmd_box *box, const int32_t *idx0, int32_t n0,
               const int32_t *idx1, int32_t n1, int32_t same_type, double rwidth, double width2, double height,
               uint32_t flags, double *force, double *vw, tmd_stream_t stream) {
    TMD_REQUIRE(box && box->dim >= 1 && box->dim <= 3 && n >= 0, "bad box or n");
    return launch_dwpair(pos, n, box->dim, box, nullptr, idx0, n0, idx1, n1, same_type, rwidth, width2, height, flags,
                         force, vw, as_stream(stream));
}

int tmd_dwpair_triclinic(const double *pos, int64_t n, const tmd_cell *cell, const int32_t *idx0, int32_t n0,
                         const int32_t *idx1, int32_t n1, int32_t same_type, double rwidth, double width2,
                         double height, uint32_t flags, double *force, double *vw, tmd_stream_t stream) {
    TMD_REQUIRE(cell && cell->dim >= 2 && cell->dim <= 3 && n >= 0, "bad cell or n");
    return launch_dwpair(pos, n, cell->dim, nullptr, cell, idx0, n0, idx1, n1, same_type, rwidth, width2, height, flags,
                         force, vw, as_stream(stream));
}

}  // extern "C"
