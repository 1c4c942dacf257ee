// K2 + K3: cell-list Lennard-Jones for large N (the reference has no neighbour structure at all:
// "This could be some kind of fancy neighbour list; here, we ignore this",
// turtlemd/potentials/lennardjones.py:158-160 -- it is O(N^2) and cannot run N = 1M).
//
// Same contract as tmd_lj_allpairs (lennardjones.py:145-273 with box.py:384-403 fused), for a
// fully periodic orthorhombic box with >= 3 cells of edge >= max rcut per direction, where the
// single-nearest-image rule of the reference and the 3^dim stencil select the same pairs.
//
// K2 build, every call (positions move every step; nothing persists between calls):
//   assign   wrap a COPY of the coordinates into the box (the API-visible pos stays unwrapped, as
//            the reference integrators leave it, integrators.py:90), cell index, atomic rank
//   scan     exclusive prefix sum of the cell populations (one block)
//   scatter  slot = start[cell] + rank
//   sort     one thread per cell orders its atoms by index (so the result does not depend on the
//            order the atomics happened to resolve: bit-reproducible) and writes the sorted,
//            padded (x, y, z, type) records
// K3 traversal: one thread per atom in cell order.  Phase 1 walks the 3^dim neighbour cells and only
// FILTERS (7 FP64 instructions per candidate: the periodic image is a per-cell constant folded
// into the row atom's coordinates) and appends passing candidates to a per-thread list in shared
// memory; phase 2 drains the lists warp-synchronously, so the expensive pair arithmetic runs with
// ~3/4 of the lanes active instead of ~1/7.  One thread owns one atom: no atomics on forces.
#include <string.h>

#include "cell_build.cuh"

namespace tmd {

constexpr int kCellThreads = 128;
constexpr int kHitCap = 96;  // per-thread candidate list (mean ~52 at rho=0.8, rc=2.5)

template <int DIM>
__global__ void __launch_bounds__(kCellThreads, 4) lj_cells_kernel(const CellArgs a) {
    extern __shared__ unsigned s_hits[];  // [kHitCap][kCellThreads]
    __shared__ double s_tab[6 * kMaxT2];
    __shared__ double s_red[10 * (kCellThreads / 32)];

    const int tid = threadIdx.x;
    const int64_t row = (int64_t)blockIdx.x * kCellThreads + tid;
    const bool valid = row < a.count;
    const bool whole = a.first == 0 && a.count == a.n;
    // whole system: thread <-> slot (best locality); a slab of rows: thread <-> atom first + row
    const int slot = valid ? (whole ? (int)row : a.slot_of[a.first + row]) : 0;
    const int atom = valid ? (whole ? a.order[slot] : (int)(a.first + row)) : 0;
    const int ntypes = a.ntypes;
    if (ntypes > 1) {
        for (int k = tid; k < 6 * kMaxT2; k += kCellThreads) s_tab[k] = a.tab.c[k / kMaxT2][k % kMaxT2];
        __syncthreads();
    }

    const double4 me = a.sorted[slot];
    const int ti = (int)__double_as_longlong(me.w);
    int c[3];
    {
        const int cell = a.cell_of[atom];
        c[0] = cell % a.grid.nc[0];
        c[1] = (cell / a.grid.nc[0]) % a.grid.nc[1];
        c[2] = cell / (a.grid.nc[0] * a.grid.nc[1]);
    }
    const long long rc2_bits = valid ? __double_as_longlong(a.rc2max) : -1;

    PairAcc acc;
#pragma unroll
    for (int k = 0; k < 3; ++k) acc.f[k] = 0.0;
    acc.v = 0.0;
#pragma unroll
    for (int k = 0; k < 9; ++k) acc.w[k] = 0.0;

    int cnt = 0;

    auto drain = [&]() {
        int most = cnt;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) most = max(most, __shfl_xor_sync(0xffffffffu, most, off));
        for (int k = 0; k < most; ++k) {
            if (k < cnt) {
                const unsigned e = s_hits[k * kCellThreads + tid];
                const int j = (int)(e & 0x07FFFFFFu);
                const int code = (int)(e >> 27);
                const double4 q = a.sorted[j];
                const int s0 = code % 3 - 1, s1 = (code / 3) % 3 - 1, s2 = code / 9 - 1;
                double d[3] = {0.0, 0.0, 0.0};
                d[0] = (me.x - s0 * a.grid.len[0]) - q.x;
                if (DIM > 1) d[1] = (me.y - s1 * a.grid.len[1]) - q.y;
                if (DIM > 2) d[2] = (me.z - s2 * a.grid.len[2]) - q.z;
                double rsq = d[0] * d[0];
                if (DIM > 1) rsq = fma(d[1], d[1], rsq);
                if (DIM > 2) rsq = fma(d[2], d[2], rsq);
                double lj1, lj2, lj3, lj4, off, rc2;
                if (ntypes > 1) {
                    const int p = ti * ntypes + (int)__double_as_longlong(q.w);
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
                if (rsq < rc2) lj_pair<DIM>(d, rsq, lj1, lj2, lj3, lj4, off, a.flags, acc);
            }
        }
        cnt = 0;
    };

    // ---- phase 1: filter the 3^DIM neighbour cells ----
    for (int dz = (DIM > 2 ? -1 : 0); dz <= (DIM > 2 ? 1 : 0); ++dz) {
        for (int dy = (DIM > 1 ? -1 : 0); dy <= (DIM > 1 ? 1 : 0); ++dy) {
            for (int dx = -1; dx <= 1; ++dx) {
                int nb[3] = {c[0] + dx, c[1] + dy, c[2] + dz};
                int sh[3] = {0, 0, 0};
#pragma unroll
                for (int k = 0; k < DIM; ++k) {
                    if (nb[k] < 0) { nb[k] += a.grid.nc[k]; sh[k] = -1; }
                    else if (nb[k] >= a.grid.nc[k]) { nb[k] -= a.grid.nc[k]; sh[k] = 1; }
                }
                // neighbour image sits at x_j + sh*L; fold the shift into the row atom instead
                const double xs = me.x - sh[0] * a.grid.len[0];
                const double ys = me.y - sh[1] * a.grid.len[1];
                const double zs = me.z - sh[2] * a.grid.len[2];
                const unsigned code = (unsigned)((sh[0] + 1) + 3 * (sh[1] + 1) + 9 * (sh[2] + 1)) << 27;
                const int ncell = (nb[2] * a.grid.nc[1] + nb[1]) * a.grid.nc[0] + nb[0];
                int j = a.starts[ncell];
                const int jend = a.starts[ncell + 1];
                // chunks of <= 32 candidates; drain first if any lane could overflow its list
                int more = valid ? jend - j : 0;
                while (__any_sync(0xffffffffu, more > 0)) {
                    if (__any_sync(0xffffffffu, cnt + 32 > kHitCap)) drain();
                    const int stop = min(jend, j + 32);
                    for (; j < stop && valid; ++j) {
                        const double4 q = a.sorted[j];
                        double d = xs - q.x;
                        double r = d * d;
                        if (DIM > 1) { d = ys - q.y; r = fma(d, d, r); }
                        if (DIM > 2) { d = zs - q.z; r = fma(d, d, r); }
                        if (__double_as_longlong(r) < rc2_bits && j != slot) {
                            s_hits[cnt * kCellThreads + tid] = code | (unsigned)j;
                            ++cnt;
                        }
                    }
                    more = valid ? jend - j : 0;
                }
            }
        }
    }
    // ---- phase 2 ----
    drain();

    if (valid && (a.flags & TMD_WANT_F)) {
        double *dst = a.force + (int64_t)atom * DIM;
#pragma unroll
        for (int k = 0; k < DIM; ++k) dst[k] = (a.flags & TMD_ACCUMULATE) ? dst[k] + acc.f[k] : acc.f[k];
    }
    double red[10];
    red[0] = acc.v;
#pragma unroll
    for (int k = 0; k < 9; ++k) red[1 + k] = acc.w[k];
    block_sum<10, kCellThreads>(red, s_red);
    if (tid == 0) {
#pragma unroll
        for (int k = 0; k < 10; ++k) a.partial_vw[(size_t)blockIdx.x * 10 + k] = red[k];
    }
}

}  // namespace tmd

using namespace tmd;

extern "C" int tmd_lj_cells_usable(int64_t n, const tmd_box *box, const double *table, int32_t ntypes,
                                   int *usable) {
    TMD_REQUIRE(usable != nullptr, "usable is NULL");
    TMD_REQUIRE(box && box->dim >= 1 && box->dim <= 3, "box->dim must be 1..3");
    *usable = 0;
    LjTable tab;
    double rc2max = 0.0;
    TMD_CHECK(load_lj_table(table, ntypes, tab, rc2max));
    CellGrid g;
    if (n >= 2 && n < (1ll << 27) && make_grid(box, sqrt(rc2max), g)) *usable = 1;
    return TMD_OK;
}

extern "C" int tmd_lj_cells(const double *pos, const int32_t *tidx, int64_t n, const tmd_box *box,
                            const double *table, int32_t ntypes, uint32_t flags, int64_t first, int64_t count,
                            double *force, double *vw, tmd_stream_t stream) {
    TMD_REQUIRE(box && box->dim >= 1 && box->dim <= 3, "box->dim must be 1..3");
    TMD_REQUIRE(n >= 0 && first >= 0 && count >= 0 && first + count <= n, "rows [first, first+count) must lie in [0, n)");
    TMD_REQUIRE(pos || n == 0, "pos is NULL");
    TMD_REQUIRE(ntypes == 1 || tidx, "tidx is required when ntypes > 1");
    TMD_REQUIRE(!(flags & TMD_WANT_F) || force, "force is NULL but TMD_WANT_F is set");
    TMD_REQUIRE(vw != nullptr, "vw is NULL");
    TMD_CHECK(require_device());
    cudaStream_t s = as_stream(stream);
    const int dim = box->dim;

    CellArgs a;
    TMD_CHECK(load_lj_table(table, ntypes, a.tab, a.rc2max));
    if (n < 2 || n >= (1ll << 27) || !make_grid(box, sqrt(a.rc2max), a.grid)) {
        set_error("tmd_lj_cells: needs a fully periodic box with >= 3 cells of edge rcut per direction "
                  "and 2 <= n < 2^27; use tmd_lj_allpairs");
        return TMD_ERR_UNSUPPORTED;
    }
    if (count == 0) return launch_reduce_vw(nullptr, 0, 0.5, flags, vw, s);

    a.pos = pos;
    a.tidx = tidx;
    a.n = n;
    a.first = first;
    a.count = count;
    a.ntypes = ntypes;
    a.flags = flags;
    a.force = force;
    const int ncells = a.grid.ncells;
    const int blocks = (int)((count + kCellThreads - 1) / kCellThreads);
    int *ints = nullptr;
    TMD_CHECK(arena_get(SLOT_CELL_COUNTS, (size_t)(2 * (ncells + 1)) * sizeof(int), (void **)&ints));
    a.counts = ints;
    a.starts = ints + (ncells + 1);
    TMD_CHECK(arena_get(SLOT_CELL_KEYS, (size_t)n * 4 * sizeof(int), (void **)&ints));
    a.cell_of = ints;
    a.rank_of = ints + n;
    a.order = ints + 2 * n;
    a.slot_of = ints + 3 * n;
    TMD_CHECK(arena_get(SLOT_WRAPPED, (size_t)n * sizeof(double4), (void **)&a.wrapped));
    TMD_CHECK(arena_get(SLOT_CELL_SORTED, (size_t)n * sizeof(double4), (void **)&a.sorted));
    TMD_CHECK(arena_get(SLOT_PARTIAL_VW, (size_t)blocks * 10 * sizeof(double), (void **)&a.partial_vw));

    // ---- K2 ----
    a.gate = nullptr;
    a.image = nullptr;
    a.sortedf = nullptr;
    TMD_CHECK(launch_cell_build(a, dim, s));

    // ---- K3 ----
    const size_t smem = (size_t)kHitCap * kCellThreads * sizeof(unsigned);
    static bool attr_set[3] = {false, false, false};
    if (!attr_set[dim - 1]) {
        if (dim == 1) TMD_CUDA(cudaFuncSetAttribute(lj_cells_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        else if (dim == 2) TMD_CUDA(cudaFuncSetAttribute(lj_cells_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        else TMD_CUDA(cudaFuncSetAttribute(lj_cells_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set[dim - 1] = true;
    }
    if (dim == 1) ::tmd::count_launch(), lj_cells_kernel<1><<<blocks, kCellThreads, smem, s>>>(a);
    else if (dim == 2) ::tmd::count_launch(), lj_cells_kernel<2><<<blocks, kCellThreads, smem, s>>>(a);
    else ::tmd::count_launch(), lj_cells_kernel<3><<<blocks, kCellThreads, smem, s>>>(a);
    TMD_CUDA(cudaGetLastError());
    // every pair was visited from both of its atoms
    return launch_reduce_vw(a.partial_vw, blocks, 0.5, flags, vw, s);
}
