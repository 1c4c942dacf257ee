// Lennard-Jones in a TriclinicBox: the reference's pair loop
// (potentials/lennardjones.py:145-273) with the nearest image of TriclinicBox.pbc_dist
// (system/box.py:271-296, after Mezei 1992): reduce with the cell matrix,
//     dist = d - M rint(M^-1 d),
// accept it when |dist| < half the shortest diagonal entry of M, otherwise scan the 3^dim
// neighbouring images `dist + T_k` and take the first shortest one if it is shorter.  The
// reference applies this in every direction whatever `periodic` says; so does the kernel.
//
// Tiled all-pairs kernel, every pair visited from both of its rows (halved at the end), column
// ranges split over blocks, partial forces summed in a fixed order: deterministic.  A first,
// correct kernel for the skewed cells TurtleMD users set up by hand (N up to a few 10^4): the exact
// image scan costs ~170 FP64 instructions for every pair the early exit does not catch, and there
// is no FP32 pre-filter yet.
#include "lj_common.cuh"
#include "triclinic.cuh"

namespace tmd {

constexpr int kTriRows = 128;
constexpr int kTriTile = 128;

struct TriArgs {
    const double *pos;
    const int32_t *tidx;
    int64_t n, first, count;
    int ntypes;
    uint32_t flags;
    int n_itile, n_split;
    int64_t cols_per_split;
    TriCell cell;
    LjTable tab;
    double *partial_f;   // [n_split][count*dim]
    double *partial_vw;  // [blocks][10]
};

template <int DIM>
__global__ void __launch_bounds__(kTriRows) lj_triclinic_kernel(const TriArgs a) {
    __shared__ double s_x[kTriTile * 3];
    __shared__ int s_t[kTriTile];
    __shared__ double s_red[10 * (kTriRows / 32)];
    const int tid = threadIdx.x;
    const int itile = blockIdx.x % a.n_itile, split = blockIdx.x / a.n_itile;
    const int64_t r = (int64_t)itile * kTriRows + tid;  // row inside the block of rows
    const bool valid = r < a.count;
    const int64_t i = a.first + (valid ? r : 0);
    double xi[3] = {0.0, 0.0, 0.0};
#pragma unroll
    for (int k = 0; k < DIM; ++k) xi[k] = a.pos[i * DIM + k];
    const int ti = a.ntypes > 1 ? a.tidx[i] : 0;
    PairAcc acc;
    acc.v = 0.0;
#pragma unroll
    for (int k = 0; k < 3; ++k) acc.f[k] = 0.0;
#pragma unroll
    for (int k = 0; k < 9; ++k) acc.w[k] = 0.0;

    const int64_t jbegin = (int64_t)split * a.cols_per_split;
    const int64_t jend = min(a.n, jbegin + a.cols_per_split);
    for (int64_t j0 = jbegin; j0 < jend; j0 += kTriTile) {
        __syncthreads();
        {
            const int64_t j = j0 + tid;
            if (j < jend) {
#pragma unroll
                for (int k = 0; k < DIM; ++k) s_x[tid * 3 + k] = a.pos[j * DIM + k];
                s_t[tid] = a.ntypes > 1 ? a.tidx[j] : 0;
            }
        }
        __syncthreads();
        const int cnt = (int)min((int64_t)kTriTile, jend - j0);
        if (!valid) continue;
        for (int c = 0; c < cnt; ++c) {
            if (j0 + c == i) continue;
            double d[3] = {0.0, 0.0, 0.0};
#pragma unroll
            for (int k = 0; k < DIM; ++k) d[k] = __dsub_rn(xi[k], s_x[c * 3 + k]);  // pos[i] - pos[j]  (:158)
            tri_min_image<DIM>(a.cell, d);
            double rsq = __dmul_rn(d[0], d[0]);
            if (DIM > 1) rsq = __dadd_rn(rsq, __dmul_rn(d[1], d[1]));
            if (DIM > 2) rsq = __dadd_rn(rsq, __dmul_rn(d[2], d[2]));
            const int p = ti * a.ntypes + s_t[c];
            if (rsq < a.tab.c[5][p])  // strict '<' (:291-294)
                lj_pair<DIM, false>(d, rsq, a.tab.c[0][p], a.tab.c[1][p], a.tab.c[2][p], a.tab.c[3][p], a.tab.c[4][p],
                                    a.flags, acc);
        }
    }
    if (valid && (a.flags & TMD_WANT_F)) {
        double *dst = a.partial_f + ((size_t)split * a.count + r) * DIM;
#pragma unroll
        for (int k = 0; k < DIM; ++k) dst[k] = acc.f[k];
    }
    double red[10];
    red[0] = acc.v;
#pragma unroll
    for (int k = 0; k < 9; ++k) red[1 + k] = (a.flags & TMD_WANT_W) ? acc.w[k] : 0.0;
    block_sum<10, kTriRows>(red, s_red);
    if (tid == 0) {
#pragma unroll
        for (int k = 0; k < 10; ++k) a.partial_vw[(size_t)blockIdx.x * 10 + k] = red[k];
    }
}

int load_tri_cell(const tmd_cell *cell, TriCell &out) {
    if (!cell || cell->dim < 2 || cell->dim > 3) {
        set_error("tmd_cell: dim must be 2 or 3 (a TriclinicBox has dim > 1, box.py:222-230)");
        return TMD_ERR_INVALID;
    }
    const int dim = cell->dim;
    out.nimg = dim == 2 ? 9 : 27;
    out.half = cell->half_width;
    for (int k = 0; k < 9; ++k) {
        out.m[k] = cell->matrix[k];
        out.inv[k] = cell->inverse[k];
    }
    for (int k = 0; k < 27 * 3; ++k) out.t[k / 3][k % 3] = 0.0;
    for (int g = 0; g < out.nimg; ++g)
        for (int k = 0; k < dim; ++k) out.t[g][k] = cell->translations[g * 3 + k];
    return TMD_OK;
}

}  // namespace tmd

using namespace tmd;

extern "C" {

int tmd_lj_triclinic(const double *pos, const int32_t *tidx, int64_t n, const tmd_cell *cell, const double *table,
                     int32_t ntypes, uint32_t flags, int64_t first, int64_t count, double *force, double *vw,
                     tmd_stream_t stream) {
    TMD_REQUIRE(n >= 0 && first >= 0 && count >= 0 && first + count <= n, "rows [first, first+count) must lie in [0, n)");
    TMD_REQUIRE(pos || n == 0, "pos is NULL");
    TMD_REQUIRE(ntypes == 1 || tidx, "tidx is required when ntypes > 1");
    TMD_REQUIRE(!(flags & TMD_WANT_F) || force, "force is NULL but TMD_WANT_F is set");
    TMD_REQUIRE(vw != nullptr, "vw is NULL");
    TMD_CHECK(require_device());
    cudaStream_t s = as_stream(stream);
    TriArgs a;
    TMD_CHECK(load_tri_cell(cell, a.cell));
    double rc2max = 0.0;
    TMD_CHECK(load_lj_table(table, ntypes, a.tab, rc2max));
    if (count == 0 || n == 0) return launch_reduce_vw(nullptr, 0, 0.5, flags, vw, s);
    const int dim = cell->dim;
    a.pos = pos;
    a.tidx = tidx;
    a.n = n;
    a.first = first;
    a.count = count;
    a.ntypes = ntypes;
    a.flags = flags;
    a.n_itile = (int)((count + kTriRows - 1) / kTriRows);
    // ~4 waves of resident blocks, column ranges in whole tiles
    const int64_t n_jtile = (n + kTriTile - 1) / kTriTile;
    int64_t want_split = (int64_t)num_sms() * 16 / a.n_itile;
    if (want_split < 1) want_split = 1;
    if (want_split > n_jtile) want_split = n_jtile;
    const int64_t tiles_per_split = (n_jtile + want_split - 1) / want_split;
    a.cols_per_split = tiles_per_split * kTriTile;
    a.n_split = (int)((n_jtile + tiles_per_split - 1) / tiles_per_split);
    const int blocks = a.n_itile * a.n_split;
    TMD_CHECK(arena_get(SLOT_PARTIAL_F, (size_t)a.n_split * count * dim * sizeof(double) + 8, (void **)&a.partial_f));
    TMD_CHECK(arena_get(SLOT_PARTIAL_VW, (size_t)blocks * 10 * sizeof(double), (void **)&a.partial_vw));
    if (dim == 2) ::tmd::count_launch(), lj_triclinic_kernel<2><<<blocks, kTriRows, 0, s>>>(a);
    else ::tmd::count_launch(), lj_triclinic_kernel<3><<<blocks, kTriRows, 0, s>>>(a);
    TMD_CUDA(cudaGetLastError());
    if (flags & TMD_WANT_F) TMD_CHECK(launch_lj_finish(a.partial_f, a.n_split, count, dim, first, flags, force, s));
    return launch_reduce_vw(a.partial_vw, blocks, 0.5, flags, vw, s);  // every pair was visited from both rows
}

}  // extern "C"
