// K2: cell binning shared by the stateless cell-list kernel (lj_cells.cu) and the neighbour-list
// path (lj_nlist.cu).  Counting sort by cell; atoms of a cell end up ordered by index, so nothing
// downstream depends on the order in which the ranking atomics resolved.
//
// `gate`: optional device flag.  When non-NULL the kernels return at once unless *gate != 0, which
// lets a stream-ordered sequence "check displacement -> rebuild if needed -> traverse" run without
// a host round trip.
#pragma once

#include <math.h>

#include "lj_common.cuh"

namespace tmd {

struct CellGrid {
    int nc[3];
    int ncells;
    double len[3], ilen[3];
    double scale[3];  // nc / L
};

struct CellArgs {
    const double *pos;
    const int32_t *tidx;
    int64_t n, first, count;
    int ntypes;
    uint32_t flags;
    CellGrid grid;
    int *counts;      // [ncells + 1]
    int *starts;      // [ncells + 1]
    int *cell_of;     // [n]
    int *rank_of;     // [n]
    int *order;       // [n] slot -> atom
    int *slot_of;     // [n] atom -> slot
    double4 *wrapped; // [n] by atom
    double4 *sorted;  // [n] by slot; .w carries the type index
    float4 *sortedf;  // [n] by slot, single-precision copy for conservative filtering (or NULL)
    double rc2max;
    double *force;
    double *partial_vw;
    const int *gate;  // see above; NULL = always run
    double *image;    // [n*3] by atom: the lattice vector floor(x/L)*L removed when wrapping (nlist path) or NULL
    LjTable tab;
};

#define TMD_GATE(g) \
    if ((g) != nullptr && *(g) == 0) return;

static __global__ void __launch_bounds__(256) cell_clear_kernel(int *counts, int n, const int *gate) {
    TMD_GATE(gate)
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) counts[i] = 0;
}

template <int DIM>
static __global__ void __launch_bounds__(256) cell_assign_kernel(CellArgs a) {
    TMD_GATE(a.gate)
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    double x[3] = {0.0, 0.0, 0.0};
    int c[3] = {0, 0, 0};
#pragma unroll
    for (int k = 0; k < DIM; ++k) {
        const double v = a.pos[i * DIM + k];
        const double w = fma(-floor(v * a.grid.ilen[k]), a.grid.len[k], v);  // into [0, L) up to rounding
        int ck = (int)(w * a.grid.scale[k]);
        ck = ck < 0 ? 0 : (ck >= a.grid.nc[k] ? a.grid.nc[k] - 1 : ck);
        x[k] = w;
        c[k] = ck;
        if (a.image) a.image[i * 3 + k] = floor(v * a.grid.ilen[k]) * a.grid.len[k];
    }
    const int cell = (c[2] * a.grid.nc[1] + c[1]) * a.grid.nc[0] + c[0];
    const int t = (a.ntypes > 1) ? a.tidx[i] : 0;
    a.wrapped[i] = make_double4(x[0], x[1], x[2], __longlong_as_double((long long)t));
    a.cell_of[i] = cell;
    a.rank_of[i] = atomicAdd(&a.counts[cell], 1);
}

// exclusive scan of counts[0..ncells) into starts[0..ncells]; one block walks the array in chunks of
// 4096 (coalesced int4 loads, warp-shuffle scan, running carry).  counts[] is left ZEROED, so a
// persistent owner (the neighbour list) never needs the clear kernel again.
static __global__ void __launch_bounds__(1024) cell_scan_kernel(int *__restrict__ counts, int *__restrict__ starts,
                                                                int ncells, const int *gate) {
    TMD_GATE(gate)
    __shared__ int s_warp[32];
    __shared__ int s_carry;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < ncells; base += 4096) {
        const int e = base + tid * 4;
        int v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            v[k] = (e + k < ncells) ? counts[e + k] : 0;
            if (e + k < ncells) counts[e + k] = 0;
        }
        const int mine = v[0] + v[1] + v[2] + v[3];
        int incl = mine;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, off);
            if (lane >= off) incl += t;
        }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            const int w = s_warp[lane];
            int wi = w;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, wi, off);
                if (lane >= off) wi += t;
            }
            s_warp[lane] = wi - w;  // exclusive prefix of the warp totals
        }
        __syncthreads();
        int run = s_carry + s_warp[warp] + incl - mine;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (e + k < ncells) starts[e + k] = run;
            run += v[k];
        }
        __syncthreads();
        if (tid == 1023) s_carry = run;
        __syncthreads();
    }
    if (tid == 0) starts[ncells] = s_carry;
}

static __global__ void __launch_bounds__(256) cell_scatter_kernel(CellArgs a) {
    TMD_GATE(a.gate)
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    a.order[a.starts[a.cell_of[i]] + a.rank_of[i]] = (int)i;
}

static __global__ void __launch_bounds__(128) cell_sort_kernel(CellArgs a) {
    TMD_GATE(a.gate)
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= a.grid.ncells) return;
    const int lo = a.starts[c], hi = a.starts[c + 1];
    for (int p = lo + 1; p < hi; ++p) {  // insertion sort, ~13 atoms per cell
        const int key = a.order[p];
        int q = p - 1;
        while (q >= lo && a.order[q] > key) {
            a.order[q + 1] = a.order[q];
            --q;
        }
        a.order[q + 1] = key;
    }
    for (int p = lo; p < hi; ++p) {
        const int i = a.order[p];
        const double4 w = a.wrapped[i];
        a.sorted[p] = w;
        if (a.sortedf) a.sortedf[p] = make_float4((float)w.x, (float)w.y, (float)w.z, 0.f);
        a.slot_of[i] = p;
    }
}

static inline bool make_grid(const tmd_box *box, double reach, CellGrid &g) {
    if (!(reach > 0.0)) return false;
    const double edge = reach * (1.0 + 1e-6);  // slack: an atom on a face may land in either cell
    long long ncells = 1;
    for (int k = 0; k < 3; ++k) {
        g.nc[k] = 1;
        g.len[k] = g.ilen[k] = g.scale[k] = 0.0;
    }
    for (int k = 0; k < box->dim; ++k) {
        if (!box->periodic[k]) return false;
        const double len = box->length[k];
        if (!(len > 0.0) || len != len || len > 1e300) return false;
        const double q = floor(len / edge);
        if (q < 3.0) return false;  // fewer than 3 cells: the stencil would see an atom twice
        g.nc[k] = q > 2048.0 ? 2048 : (int)q;
        g.len[k] = len;
        g.ilen[k] = box->ilength[k];
        g.scale[k] = g.nc[k] / len;
        ncells *= g.nc[k];
    }
    if (ncells > (1ll << 30)) return false;
    g.ncells = (int)ncells;
    return true;
}


// Enqueue assign -> scan -> scatter -> sort for `a` (arrays already set).
// clear_first = false: the caller guarantees counts[] is zero (cudaMemset at allocation; the scan
// re-zeroes it every build).
static inline int launch_cell_build(CellArgs &a, int dim, cudaStream_t s, bool clear_first = true) {
    const int ncells = a.grid.ncells;
    const int nb = (int)((a.n + 255) / 256);
    if (clear_first) ::tmd::count_launch(), cell_clear_kernel<<<(ncells + 1 + 255) / 256, 256, 0, s>>>(a.counts, ncells + 1, a.gate);
    if (dim == 1) ::tmd::count_launch(), cell_assign_kernel<1><<<nb, 256, 0, s>>>(a);
    else if (dim == 2) ::tmd::count_launch(), cell_assign_kernel<2><<<nb, 256, 0, s>>>(a);
    else ::tmd::count_launch(), cell_assign_kernel<3><<<nb, 256, 0, s>>>(a);
    ::tmd::count_launch(), cell_scan_kernel<<<1, 1024, 0, s>>>(a.counts, a.starts, ncells, a.gate);
    ::tmd::count_launch(), cell_scatter_kernel<<<nb, 256, 0, s>>>(a);
    ::tmd::count_launch(), cell_sort_kernel<<<(ncells + 127) / 128, 128, 0, s>>>(a);
    TMD_CUDA(cudaGetLastError());
    return TMD_OK;
}

}  // namespace tmd
