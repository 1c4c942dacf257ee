// K2 + K3 with a persistent Verlet neighbour list: the large-N Lennard-Jones path.
//
// Same contract as tmd_lj_allpairs / tmd_lj_cells (lennardjones.py:145-273 + box.py:384-403), but
// the per-atom candidate lists are kept between calls in a tmd_nlist handle and reused while no
// atom has moved more than skin/2 since they were built (then every pair closer than rcut is still
// listed, so forces are those of the full evaluation -- this is an acceleration structure, not an
// approximation).  Everything is stream-ordered, there is no host round trip.  Per call:
//
//   gather    xs[slot] = pos[atom] - image[atom]  (32-byte records in cell order; the lattice
//   + check   vector `image` is frozen at build time, so coordinates stay continuous between
//             builds) and, in the same pass, |xs[slot] - build position|^2 > (skin/2)^2 -> flag
//   build     two kernels that return at once unless the flag is set.  (1) one cooperative kernel,
//             phases separated by grid syncs: bin a wrapped copy into cells (edge >=
//             (rcut+skin)/H, H = 1), exclusive scan, scatter, per-cell ordering by atom index (so
//             nothing depends on how the ranking atomics resolved).  (2) one thread per row atom walks
//             the (2H+1)^dim cells around it with a conservative SINGLE-precision filter on the packed
//             FP32 pipe and appends to its row-major list, logically ascending by (stencil row, slot),
//             stored in transposed blocks of 16 entries (two 32-byte stores per block).
//   traverse  four lanes per row atom, persistent blocks; 6 FP64 instructions per candidate to the exact
//             cut-off test (strict rsq < rcut2, lennardjones.py:291-294), ~70 % pass.  Per-lane force
//             accumulators, shuffle reduction in a fixed order, no atomics -> bit-reproducible.  The
//             block that finishes last reduces the per-block (V, W) partials.
//
// Virial.  The reference sums f (x) delta over pairs (lennardjones.py:272).  With delta_ij =
// x_i - x_j - S_ij (S = the lattice vector of the image) and f_ij = -f_ji this equals
//     W = sum_i F_i (x) x_i  -  1/2 sum_i sum_j f_ij (x) S_ij ,
// i.e. six FMAs per ATOM plus a correction only for pairs that cross a box face, instead of six
// FMAs per pair.  x_i is taken relative to the box centre.  Agreement with the pair sum: ~1e-14.
//
// A row whose candidate count exceeds the list capacity is not truncated: its count is stored
// as-is and the traversal walks the build-time cells for that row instead (slow, exact).
#include <cooperative_groups.h>
#include <stdlib.h>
#include <string.h>

#include <new>
#include <vector>

#include "lj_common.cuh"
#include "integrator_math.cuh"

namespace cg = cooperative_groups;

namespace tmd {

constexpr int kNlThreads = 128;
constexpr unsigned kSlotMask = 0x07FFFFFFu;
constexpr int kHomeCode = 13;  // (0,0,0) shift
constexpr int kNlFlagWords = 16;  // nlflags[]: see NlArgs; [8], [9] = the two gate words of a resident run
constexpr int kNeverHi = 0x7FF80BAD;  // high word of a NaN payload no arithmetic produces (see nl_force_kernel)
constexpr int kBinThreads = 256;           // cooperative binning kernel
constexpr int kScanChunk = kBinThreads * 8;

struct NlGrid {
    int nc[3];
    int ncells;
    int h;            // stencil half width in cells (y, z; x when the grid is not refined)
    int hx;           // stencil half width along x in (possibly finer) x cells: cells are cut finer along x, the
                      // direction in which the cells of a stencil row are consecutive in slot order (nl_make_grid)
    double len[3], ilen[3];
    double scale[3];  // nc / L
};

// Step barrier of a multi-GPU slab run (slabrun.cuh), executed by the traversal block that finishes last: device-resident
// description of where the ranks' mailboxes are and which words of this rank the barrier maintains.
constexpr int kNlMaxWorld = 16;
struct NlStepBarrier {
    int world, rank;
    unsigned long long *mail[kNlMaxWorld];  // every rank's mailbox [2][world] (peer memory)
    unsigned long long *epoch;              // this rank's barrier count
    int *local_flag;                        // raised by this rank's epilogues: an atom moved more than skin / 2
    int *gates;                             // [2]: the rebuild gates of even / odd steps
    int *counters;                          // [5]: append counters of a rebuild, zeroed here
    int *err;
};

struct NlArgs {
    const double *pos;
    const int32_t *tidx;
    int64_t n, first, count;
    int row_mode;        // 0: rows are atoms first .. first+count-1, forces indexed by atom;
                         // 1: rows are cell-order SLOTS first .. first+count-1, forces indexed by slot (slab path)
    int ntypes;
    uint32_t flags;
    NlGrid grid;
    int *nlflags;        // [0] rebuild, [1] rows on the slow path, [2] builds, [3] calls,
                         // [4], [5] slots below / above the row range that the lists reference (slab path)
    int *counts;         // [ncells]     zero between builds
    int *starts;         // [ncells + 1]
    int *chunk_sums;     // [ceil(ncells / kScanChunk)]
    int *cell_of;        // [n] by atom
    int *rank_of;        // [n] by atom
    int *order;          // [n] slot -> atom
    int *order_tmp;      // [n] provisional order of a rebuild
    int *slot_of;        // [n] atom -> slot
    double4 *wrapped;    // [n] by atom: wrapped position at build time, .w = type
    double4 *xb;         // [n] by slot: the same, in cell order (build positions)
    float4 *xbf;         // [n/2][2] by slot pair: single-precision build positions {x0 x1 y0 y1} {z0 z1 - -}
    double4 *xs;         // [n] by slot: current positions (image removed), .w = type
    double *image;       // [n*3] by atom: lattice vector removed when wrapping
    unsigned *entries;   // [rows][cap]
    int *nneigh;         // [rows]
    int cap;
    int uniform_table;   // every type pair has the same six coefficients: the one-type kernel applies
    float reach2f;       // conservative single-precision filter radius^2
    double half_skin2;
    double centre[3];
    double *force;
    double *partial_vw;
    double *vw;          // (V, W) of the call, written by the traversal block that finishes last
    const int *gate;     // the build kernels return at once unless *gate != 0 (nlflags[0], or a resident run's flag word)
    // ---- resident run (tmd_nlist_run_vv): slot-ordered state advanced by the traversal's epilogue ----
    double *xu;          // [n*dim] by slot: unwrapped positions (the reference's particles.pos, never wrapped)
    double *vs;          // [n*dim] by slot: velocities
    double *fs;          // [n*dim] by slot: forces of the last evaluation
    const double *fx;    // [n*dim] by slot or NULL: forces of the other potentials of the system, added to fs
    const double *imass_s;  // [n] by slot
    const double *image_s;  // [n*3] by slot: lattice vector removed at build time (image[] in slot order)
    double4 *xs_next;    // [n] by slot: records of the NEXT step (double buffer; xs is read by every block)
    int *flag_next;      // set when an atom has moved more than skin/2: gates the build of the next step
    int *flag_clear;     // the word that gated this step's build: cleared for its next use
    double dt, half_dt;
    int epi_last;        // 1: last step of the run, closing half kick only
    int md_rebuild;      // the build kernels move the resident state: binning reads xu / vs through the OLD slot_of and
                         // leaves pos_w / vel_w (atom order) behind, the list kernel fills the NEW slots from them
    double *pos_w, *vel_w;
    const double *imass; // [n] by atom
    // ---- slab runs (slabrun.cuh): sizes decided on the device, in-cell order by global index, halo push ----
    const int *dyn;      // NULL, or [0] = atoms to bin (own atoms + ghost copies of the neighbours' boundary layers)
    int row_cell_lo, row_cell_hi;  // row_cell_hi > 0: rows = slots of cells [lo, hi): first = starts[lo], count = starts[hi] - first
    const int *gid;      // NULL, or global atom index by local index: slots inside a cell ascend by it
    const int *halo;     // EPI = 2: [0] lo_begin [1] lo_end [2] dest_lo [3] hi_begin [4] hi_end [5] dest_hi (slots)
    double4 *peer_lo_next, *peer_hi_next;  // EPI = 2: the record buffers of the next step on the two neighbouring ranks
    int dynamic_rows;          // EPI >= 2: hand rows out dynamically on steps that need no (V, W)
    int *work;                 // EPI = 2: next unclaimed chunk of rows (steps that need no (V, W) hand rows out dynamically)
    const NlStepBarrier *bar;  // EPI = 2: non-NULL = the block that finishes last runs the step barrier (real ranks)
    int bar_gate_out;          //          which gate word receives the OR of the ranks' flags
    int bar_has_cond;          //          inside a step graph: also decide the IF node around the next step's rebuild
    cudaGraphConditionalHandle bar_cond;
    LjTable tab;
};

}  // namespace tmd

struct tmd_nlist {
    double skin;
    int device;
    // signature of what the lists were built for (a change forces a rebuild)
    int64_t n, first, count;
    int dim, ntypes;
    double len[3];
    double reach;
    const int32_t *tidx;  // the type array the records were built from (types are baked into them): a rebound array rebuilds
    // device memory, owned
    int *flags;
    unsigned *entries;
    int *nneigh;
    int *ints;           // counts, starts, chunk_sums, cell_of, rank_of, order, slot_of, order_tmp
    double4 *wrapped, *xb, *xs;
    float4 *xbf;
    double *image;
    int64_t cap_n, cap_rows, cap_cells;
    int cap;
    int build_blocks[3];  // list-kernel grid per dim (0 = not yet queried)
    int bin_blocks[3];    // cooperative binning grid per dim
    // slab path (tmd_nlist_build_slots / tmd_nlist_force_slots)
    int row_mode;         // signature: 0 = atom rows, 1 = slot rows
    double4 *xs_ext;      // caller-owned current-position records used instead of xs (exchange buffer)
    tmd::NlArgs *saved;   // launch arguments of the last slot build, reused by the traversal
    // resident runs (tmd_nlist_run_vv): slot-ordered state, owned
    double4 *xs2;         // second record buffer (the epilogue of a step writes the records of the next)
    double *md;           // xu, vs, fs, fx, image_s (3 doubles per slot each), imass_s (1)
    int64_t cap_md;
    void *dw_scratch;     // resident runs with a DoubleWellPair term: prev_slot[128] ints, then partials[10] doubles
    // measurement (tmd_nlist_set_timing): CUDA events around every traversal launch of a resident run
    int timing;
    double traversal_ms;
    int64_t traversals;
};

namespace tmd {

// 1/x to <= 1 ulp: MUFU.RCP64H seed (2^-23) + two Newton steps.  x is a squared distance inside
// the cut-off: positive, normal.
__device__ __forceinline__ double rcp_fast(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    e = fma(-x, r, 1.0);
    return fma(r, e, r);
}

// 32-byte record in one 256-bit load (LDG.E.256 on sm_100a); xs is not written while a traversal runs
__device__ __forceinline__ double4 ld_record(const double4 *p) {
    double4 v;
    asm("ld.global.nc.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(v.x), "=d"(v.y), "=d"(v.z), "=d"(v.w) : "l"(p));
    return v;
}

// 8 bytes global -> shared without a register in between (LDGSTS); completion: cp_async_wait_all()
__device__ __forceinline__ void cp_async8(double *smem_dst, const double *gsrc) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

__device__ __forceinline__ void st_entries8(unsigned *p, const unsigned (&e)[8]) {
    asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(e[0]), "r"(e[1]), "r"(e[2]),
                 "r"(e[3]), "r"(e[4]), "r"(e[5]), "r"(e[6]), "r"(e[7])
                 : "memory");
}

// Single-precision build positions (NlArgs::xbf): float index of coordinate c (0..2) of slot t in the
// 32-byte pair records {x0, x1, y0, y1, z0, z1, -, -} of two consecutive slots (see nl_list_kernel)
__device__ __forceinline__ size_t pair_index(int t, int c) { return (size_t)(t >> 1) * 8 + 2 * c + (t & 1); }

template <int DIM>
__global__ void __launch_bounds__(256) nl_gather_check_kernel(const double *__restrict__ pos,
                                                              const double *__restrict__ image,
                                                              const int *__restrict__ order,
                                                              const double4 *__restrict__ xb, int64_t n,
                                                              double half_skin2, double4 *__restrict__ xs,
                                                              int *flags) {
    const int64_t slot = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= n) return;
    const int atom = order[slot];
    const double4 b = xb[slot];  // position at build time, .w = type
    double x[3] = {0.0, 0.0, 0.0};
#pragma unroll
    for (int k = 0; k < DIM; ++k) x[k] = pos[(int64_t)atom * DIM + k] - image[(int64_t)atom * 3 + k];
    double r = (x[0] - b.x) * (x[0] - b.x);
    if (DIM > 1) r = fma(x[1] - b.y, x[1] - b.y, r);
    if (DIM > 2) r = fma(x[2] - b.z, x[2] - b.z, r);
    xs[slot] = make_double4(x[0], x[1], x[2], b.w);
    if (!(r <= half_skin2)) flags[0] = 1;  // also catches NaN
}

// ---- stencil walk shared by the list build and the traversal's slow path -------------------------
// Calls body(jbegin, jend, code) for every run of consecutive slots of the (2H+1)^DIM cells around
// cell (c0, c1, c2); code = image shift of that run as (sx+1) + 3 (sy+1) + 9 (sz+1).  Cells of one
// x-row are consecutive in slot order, so a row is one run (up to three where it wraps).
// xwin(dy, dz, lo, hi) -> false to skip a stencil row, else the x-cell range [lo, hi] to walk (indices may leave
// [0, nc): the wrap is handled here); full_x_window walks the whole stencil width.
struct full_x_window {
    int c0, hx;
    __device__ __forceinline__ bool operator()(int, int, int &lo, int &hi) const {
        lo = c0 - hx;
        hi = c0 + hx;
        return true;
    }
};

template <int DIM, typename XWin, typename Body>
__device__ __forceinline__ void for_each_run(const NlGrid &g, const int *__restrict__ starts, int c1, int c2, XWin xwin,
                                             Body body) {
    const int h = g.h;
    const int hz = DIM > 2 ? h : 0, hy = DIM > 1 ? h : 0;
    for (int dz = -hz; dz <= hz; ++dz) {
        int z = c2 + dz, sz = 0;
        if (z < 0) { z += g.nc[2]; sz = -1; }
        else if (z >= g.nc[2]) { z -= g.nc[2]; sz = 1; }
        for (int dy = -hy; dy <= hy; ++dy) {
            int y = c1 + dy, sy = 0;
            if (y < 0) { y += g.nc[1]; sy = -1; }
            else if (y >= g.nc[1]) { y -= g.nc[1]; sy = 1; }
            int lo, hi;
            if (!xwin(dy, dz, lo, hi)) continue;
            const int rowbase = (z * g.nc[1] + y) * g.nc[0];
            const int yz = 3 * (sy + 1) + 9 * (sz + 1);
            if (lo < 0) body(starts[rowbase + lo + g.nc[0]], starts[rowbase + g.nc[0]], 0 + yz);
            body(starts[rowbase + (lo < 0 ? 0 : lo)], starts[rowbase + (hi >= g.nc[0] ? g.nc[0] : hi + 1)], 1 + yz);
            if (hi >= g.nc[0]) body(starts[rowbase], starts[rowbase + hi - g.nc[0] + 1], 2 + yz);
        }
    }
}

// kBinThreads threads; returns the exclusive prefix of v over the block, total = block sum
__device__ __forceinline__ int block_exclusive_scan(int v, int *s_warp, int &total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int incl = v;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, off);
        if (lane >= off) incl += t;
    }
    __syncthreads();  // s_warp may still be read from a previous call
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    int before = 0, sum = 0;
#pragma unroll
    for (int w = 0; w < kBinThreads / 32; ++w) {
        const int t = s_warp[w];
        if (w < warp) before += t;
        sum += t;
    }
    total = sum;
    return before + incl - v;
}

// Binning (phases 1-4 of a rebuild) in one cooperative launch; gated on the rebuild flag.
template <int DIM>
__global__ void __launch_bounds__(kBinThreads) nl_bin_kernel(const NlArgs a) {
    if (*a.gate == 0) return;  // same answer in every block: nobody writes the flag while this kernel runs
    cg::grid_group grid = cg::this_grid();
    __shared__ int s_warp[kBinThreads / 32];
    const NlGrid &g = a.grid;
    const int tid = threadIdx.x;
    const int64_t gtid = (int64_t)blockIdx.x * kBinThreads + tid;
    const int64_t gthreads = (int64_t)gridDim.x * kBinThreads;
    const int ncells = g.ncells;
    const int64_t n = a.dyn ? (int64_t)a.dyn[0] : a.n;  // slab runs: the population of a rank changes from build to build

    // ---- 1. wrap a copy into the box, cell index, rank inside the cell (atomics) ----
    for (int64_t i = gtid; i < n; i += gthreads) {
        double x[3] = {0.0, 0.0, 0.0};
        int c[3] = {0, 0, 0};
#pragma unroll
        for (int k = 0; k < DIM; ++k) {
            double v;
            if (a.md_rebuild) {  // resident run: the state lives in (old) slot order; atom order is the staging area
                const int64_t e = (int64_t)a.slot_of[i] * DIM + k;
                v = a.xu[e];
                a.pos_w[i * DIM + k] = v;
                a.vel_w[i * DIM + k] = a.vs[e];
            } else {
                v = a.pos[i * DIM + k];
            }
            const double im = floor(v * g.ilen[k]) * g.len[k];
            const double w = v - im;  // into [0, L) up to rounding
            int ck = (int)(w * g.scale[k]);
            ck = ck < 0 ? 0 : (ck >= g.nc[k] ? g.nc[k] - 1 : ck);
            x[k] = w;
            c[k] = ck;
            a.image[i * 3 + k] = im;
        }
        const int cell = (c[2] * g.nc[1] + c[1]) * g.nc[0] + c[0];
        const int t = (a.ntypes > 1) ? a.tidx[i] : 0;
        a.wrapped[i] = make_double4(x[0], x[1], x[2], __longlong_as_double((long long)t));
        a.cell_of[i] = cell;
        a.rank_of[i] = atomicAdd(&a.counts[cell], 1);
    }
    grid.sync();

    // ---- 2. exclusive scan of the cell populations (chunk sums, then chunk-local scans) ----
    const int nchunks = (ncells + kScanChunk - 1) / kScanChunk;
    for (int ch = blockIdx.x; ch < nchunks; ch += gridDim.x) {
        const int e = ch * kScanChunk + tid * 8;
        int mine = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) mine += (e + k < ncells) ? a.counts[e + k] : 0;
        int total;
        block_exclusive_scan(mine, s_warp, total);
        if (tid == 0) a.chunk_sums[ch] = total;
    }
    grid.sync();
    for (int ch = blockIdx.x; ch < nchunks; ch += gridDim.x) {
        int part = 0;
        for (int q = tid; q < ch; q += kBinThreads) part += a.chunk_sums[q];
        int offset;
        block_exclusive_scan(part, s_warp, offset);  // offset = sum of all chunks before this one
        const int e = ch * kScanChunk + tid * 8;
        int v[8], mine = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            v[k] = (e + k < ncells) ? a.counts[e + k] : 0;
            if (e + k < ncells) a.counts[e + k] = 0;  // ready for the next build
            mine += v[k];
        }
        int total;
        int run = offset + block_exclusive_scan(mine, s_warp, total);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            if (e + k < ncells) a.starts[e + k] = run;
            run += v[k];
        }
    }
    if (gtid == 0) a.starts[ncells] = (int)n;
    grid.sync();

    // ---- 3. scatter atoms to a provisional slot inside their cell ----
    for (int64_t i = gtid; i < n; i += gthreads) a.order_tmp[a.starts[a.cell_of[i]] + a.rank_of[i]] = (int)i;
    grid.sync();

    // ---- 4. final slot = rank by atom index inside the cell (nothing depends on how the atomics
    //         resolved); write the records ----
    for (int64_t p = gtid; p < n; p += gthreads) {
        const int i = a.order_tmp[p];
        const int cell = a.cell_of[i];
        const int lo = a.starts[cell], hi = a.starts[cell + 1];
        int rank = 0;
        if (a.gid) {  // local indices are arbitrary on a slab: order by the global atom index
            const int mine = a.gid[i];
            for (int q = lo; q < hi; ++q) rank += a.gid[a.order_tmp[q]] < mine;
        } else {
            for (int q = lo; q < hi; ++q) rank += a.order_tmp[q] < i;
        }
        const int t = lo + rank;
        const double4 w = a.wrapped[i];
        a.order[t] = i;
        a.slot_of[i] = t;
        a.xb[t] = w;
        a.xs[t] = w;
        {
            float *xf = reinterpret_cast<float *>(a.xbf);
            xf[pair_index(t, 0)] = (float)w.x;
            xf[pair_index(t, 1)] = (float)w.y;
            xf[pair_index(t, 2)] = (float)w.z;
        }
    }
}

// Candidate lists (phase 5 of a rebuild): one thread per row atom walks the stencil runs with the
// single-precision filter.  A hit costs one shared-memory store into the thread's 32-deep ring;
// between filter rounds the ring is drained to the row-major list one 16-entry block (two STG.256)
// at a time.  Gated on the rebuild flag; persistent grid.
//
// List layout.  Logically a row is ascending by (stencil row, slot).  In memory every block of
// kBlockEntries = 16 logical entries is stored TRANSPOSED as a 4 x 4 matrix: logical entry q of a
// block sits at position (q % 4) * 4 + q / 4, so that lane l of the traversal's four-lane group
// finds the entries of its next four rounds (q = l, l + 4, l + 8, l + 12) in ONE aligned 16-byte
// load, while the four lanes of a round still work on four consecutive logical entries (nearly
// consecutive slots, whose 32-byte records share cache lines).
//
// Filter arithmetic.  The single-precision build positions are stored as 32-byte PAIR records
// {x0, x1, y0, y1, z0, z1, -, -} of two consecutive slots, and the distance test runs on the packed
// FP32 pipe (FADD2 / FMUL2 / FFMA2 on sm_100a): six arithmetic instructions per two candidates, the
// same individually rounded operations as the scalar form (d = xs - x; r = d d; r = fma(dy, dy, r); ...).
//
// Ring drains.  A lane's ring holds 64 entries.  Lanes fill at different rates; if each drained whenever
// it had a block waiting, a warp would run the drain code after nearly every filter round with one or two
// lanes active.  Instead the lanes that are converged vote: when any of them could overflow in the next
// round, ALL write their complete blocks.
constexpr int kRing = 64;
constexpr int kBlockEntries = 16;
constexpr int kChunk = 16;  // candidates per filter round (eight pair records)
constexpr int kDrainAt = kRing - kBlockEntries - kChunk;  // vote to drain when a lane holds more than this
static_assert(kDrainAt >= kBlockEntries - 1 && kDrainAt + kChunk <= kRing, "a filter round must always fit into the ring");

__device__ __forceinline__ int block_position(int q) { return (q & 3) * 4 + (q >> 2); }  // its own inverse

template <int DIM>
__global__ void __launch_bounds__(kNlThreads) nl_list_kernel(const NlArgs a) {
    if (*a.gate == 0) return;
    __shared__ unsigned s_ring[kRing][kNlThreads];
    const NlGrid &g = a.grid;
    const int tid = threadIdx.x;
    const int64_t gthreads = (int64_t)gridDim.x * kNlThreads;
    int64_t row_first = a.first, row_count = a.count;
    if (a.row_cell_hi > 0) {  // slab runs: the rows are the slots of a range of cells, known only after binning
        row_first = a.starts[a.row_cell_lo];
        row_count = a.starts[a.row_cell_hi] - row_first;
    }
    const bool whole = a.row_cell_hi == 0 && a.first == 0 && a.count == a.n;
    const float lx = (float)g.len[0], ly = (float)g.len[1], lz = (float)g.len[2];
    const float reach2 = a.reach2f;
    const int cap = a.cap;
    for (int64_t row = (int64_t)blockIdx.x * kNlThreads + tid; row < row_count; row += gthreads) {
        const int slot = whole ? (int)row : (a.row_mode ? (int)(row_first + row) : a.slot_of[row_first + row]);
        const int atom = (whole || a.row_mode) ? a.order[slot] : (int)(row_first + row);
        const float *xf = reinterpret_cast<const float *>(a.xbf);
        const float mex = xf[pair_index(slot, 0)], mey = xf[pair_index(slot, 1)], mez = xf[pair_index(slot, 2)];
        const int cell = a.cell_of[atom];
        const int c0 = cell % g.nc[0], c1 = (cell / g.nc[0]) % g.nc[1], c2 = cell / (g.nc[0] * g.nc[1]);
        unsigned *list = a.entries + (size_t)row * cap;  // cap is a multiple of 16: rows are 64-byte aligned
        int cnt = 0, flushed = 0;                        // flushed is a multiple of 16
        auto drain = [&]() {                             // write whole blocks while at least one is waiting
            while (cnt - flushed >= kBlockEntries) {
                if (flushed + kBlockEntries <= cap) {
#pragma unroll
                    for (int half = 0; half < 2; ++half) {
                        unsigned e[8];
#pragma unroll
                        for (int k = 0; k < 8; ++k)
                            e[k] = s_ring[(flushed + block_position(8 * half + k)) & (kRing - 1)][tid];
                        st_entries8(list + flushed + 8 * half, e);
                    }
                }
                flushed += kBlockEntries;
            }
        };
        // x window of a stencil row: its atoms are at least dmin away in the (y, z) plane (distance from this atom to
        // the faces of its own cell), so only |dx| <= sqrt(reach^2 - dmin^2) can be within reach.  Single precision
        // with a margin far above its rounding: the window is conservative, the distance filter below is what decides.
        const float ey = DIM > 1 ? (float)(g.len[1] / g.nc[1]) : 0.f, ez = DIM > 2 ? (float)(g.len[2] / g.nc[2]) : 0.f;
        const float fy = DIM > 1 ? fminf(fmaxf(mey - c1 * ey, 0.f), ey) : 0.f, fz = DIM > 2 ? fminf(fmaxf(mez - c2 * ez, 0.f), ez) : 0.f;
        const float sx = (float)g.scale[0];
        const float margin = 1e-3f * sqrtf(reach2) + 4e-6f * lx;
        auto xwin = [&](int dy, int dz, int &lo, int &hi) {
            const float gy = dy == 0 ? 0.f : ((dy < 0 ? fy : ey - fy) + (abs(dy) - 1) * ey);
            const float gz = dz == 0 ? 0.f : ((dz < 0 ? fz : ez - fz) + (abs(dz) - 1) * ez);
            const float w2 = reach2 - gy * gy - gz * gz;
            if (w2 <= 0.f) return false;
            const float w = sqrtf(w2) + margin;
            lo = max((int)floorf((mex - w) * sx), c0 - g.hx);
            hi = min((int)floorf((mex + w) * sx), c0 + g.hx);
            return true;
        };
        for_each_run<DIM>(g, a.starts, c1, c2, xwin, [&](int jbegin, int jend, int code) {
            const float xs = mex - (code % 3 - 1) * lx;
            const float ys = mey - ((code / 3) % 3 - 1) * ly;
            const float zs = mez - (code / 9 - 1) * lz;
            const unsigned long long xs2 = pack2(xs, xs), ys2 = pack2(ys, ys), zs2 = pack2(zs, zs);
            const unsigned tag = (unsigned)code << 27;
            // kChunk candidates = eight pair records at a time, starting at the pair that holds jbegin (loads
            // not predicated: xbf is padded, and the bits of candidates outside the run are masked afterwards);
            // hits recorded as bits, then the few set bits are turned into entries.
            for (int p0 = jbegin >> 1; 2 * p0 < jend; p0 += kChunk / 2) {
                const float4 *rec = a.xbf + 2 * (size_t)p0;
                unsigned m = 0;
#pragma unroll
                for (int u = 0; u < kChunk / 2; ++u) {
                    const float4 qa = __ldg(rec + 2 * u);  // x0 x1 y0 y1
                    unsigned long long d = sub2(xs2, pack2(qa.x, qa.y));
                    unsigned long long r = mul2(d, d);
                    if (DIM > 1) { d = sub2(ys2, pack2(qa.z, qa.w)); r = fma2(d, d, r); }
                    if (DIM > 2) {
                        const float4 qb = __ldg(rec + 2 * u + 1);  // z0 z1 - -
                        d = sub2(zs2, pack2(qb.x, qb.y));
                        r = fma2(d, d, r);
                    }
                    float r0, r1;
                    unpack2(r, r0, r1);
                    if (r0 < reach2) m |= 1u << (2 * u);
                    if (r1 < reach2) m |= 2u << (2 * u);
                }
                const int j0 = 2 * p0;
                const int left = jend - j0;
                if (left < kChunk) m &= (1u << left) - 1u;
                if (j0 < jbegin) m &= ~1u;  // the first pair record may start one slot before the run
                const unsigned self = (unsigned)(slot - j0);
                if (self < (unsigned)kChunk) m &= ~(1u << self);
                while (m) {
                    const int b = __ffs(m) - 1;
                    m &= m - 1;
                    s_ring[cnt & (kRing - 1)][tid] = tag | (unsigned)(j0 + b);
                    ++cnt;
                }
                if (__any_sync(__activemask(), cnt - flushed > kDrainAt)) drain();
            }
        });
        drain();
        if (cnt <= cap) {  // the 0..15 entries of the last, partial block (same transposed positions)
            for (int k = flushed; k < cnt; ++k) list[flushed + block_position(k - flushed)] = s_ring[k & (kRing - 1)][tid];
        } else {
            atomicAdd(&a.nlflags[1], 1);  // statistics only: the traversal handles such a row exactly
        }
        a.nneigh[row] = cnt;
        if (a.md_rebuild) {  // whole system: row == slot
#pragma unroll
            for (int k = 0; k < DIM; ++k) {
                a.xu[(int64_t)slot * DIM + k] = a.pos_w[(int64_t)atom * DIM + k];
                a.vs[(int64_t)slot * DIM + k] = a.vel_w[(int64_t)atom * DIM + k];
            }
#pragma unroll
            for (int k = 0; k < 3; ++k) const_cast<double *>(a.image_s)[(int64_t)slot * 3 + k] = a.image[(int64_t)atom * 3 + k];
            const_cast<double *>(a.imass_s)[slot] = a.imass[atom];
        }
    }
}

// Slab path: how far outside the row range [first, first+count) the lists of those rows reach, in
// slots.  A row in layer z (slowest direction) references slots of layers z-h .. z+h only, and a
// layer is a contiguous slot range, so the extents follow from the cell table; they tell the halo
// exchange how many records of the neighbouring ranks must be current.  nlflags[4] = below, [5] = above,
// [6] = atoms in the fullest layer.
__global__ void nl_reach_kernel(const NlArgs a, int dim) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const NlGrid &g = a.grid;
    const int nz = g.nc[dim - 1];
    const int layer = g.ncells / nz;
    const int64_t n = a.n, lo = a.first, hi = a.first + a.count;
    const int z_lo = a.cell_of[a.order[lo]] / layer, z_hi = a.cell_of[a.order[hi - 1]] / layer;
    int below, above;
    if (z_hi - z_lo + 1 + 2 * g.h >= nz) {
        below = (int)lo;  // the rows and their stencil span every layer: everything is needed
        above = (int)(n - hi);
    } else {
        const int zb = ((z_lo - g.h) % nz + nz) % nz, za = (z_hi + g.h) % nz;
        const int64_t need_lo = a.starts[zb * layer], need_hi = a.starts[(za + 1) * layer];
        below = (int)(need_lo <= lo ? lo - need_lo : lo + n - need_lo);
        above = (int)(need_hi >= hi ? need_hi - hi : need_hi + n - hi);
    }
    a.nlflags[4] = below;
    a.nlflags[5] = above;
    int most = 0;  // fullest layer: with the stencil width it bounds what any slab can ever need
    for (int z = 0; z < nz; ++z) most = max(most, a.starts[(z + 1) * layer] - a.starts[z * layer]);
    a.nlflags[6] = most;
}

// The step barrier (one warp, lane r talks to rank r): fence, write (epoch << 1 | flag) into every rank's mailbox, wait for
// all ranks of this epoch (bounded: ~2 s, then the run is marked failed), OR of the flags -> the gate of the next step's
// rebuild.  Mailboxes are double-buffered by epoch parity: a rank that is one barrier ahead writes the other half.
__device__ __forceinline__ void nl_step_barrier(const NlStepBarrier &b, int gate_out, int has_cond,
                                                cudaGraphConditionalHandle cond, int lane) {
    const unsigned long long e = *b.epoch + 1;
    const int mine = *b.local_flag != 0;
    __syncwarp();
    __threadfence_system();
    if (lane < b.world) {
        volatile unsigned long long *dst = b.mail[lane] + (e & 1ull) * b.world + b.rank;
        *dst = (e << 1) | (unsigned long long)mine;
    }
    int flag = 0, failed = 0;
    if (lane < b.world) {
        volatile unsigned long long *src = b.mail[b.rank] + (e & 1ull) * b.world + lane;
        unsigned long long t0, now;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        unsigned long long v = *src;
        while ((v >> 1) != e) {
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            if (now - t0 > 2000000000ull) { failed = 1; break; }
            __nanosleep(32);
            v = *src;
        }
        flag = (int)(v & 1ull);
    }
    failed = __any_sync(0xffffffffu, failed);
    flag = __any_sync(0xffffffffu, flag);
    __threadfence_system();
    if (lane == 0) {
        if (failed) atomicOr(b.err, 1);
        if (has_cond) cudaGraphSetConditional(cond, flag ? 1u : 0u);
        b.gates[gate_out] = flag;
        *b.local_flag = 0;
#pragma unroll
        for (int k = 0; k < 5; ++k) b.counters[k] = 0;
        *b.epoch = e;
    }
}

template <int G>
__device__ __forceinline__ double group_sum(double v) {
#pragma unroll
    for (int off = G / 2; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    return v;
}

// Traversal.  G = 4 lanes share one row atom: in round r of a 16-entry block lane l takes logical entry
// 4 r + l, so the lanes of a group gather records of nearly consecutive slots; the transposed block
// layout (see nl_list_kernel) hands a lane the entries of four rounds in one 16-byte load.  Blocks are
// persistent (grid-stride over groups of kNlThreads/G rows): the energy / virial partials stay in
// registers for the whole kernel and are reduced once.  SINGLE: one particle type (coefficients in
// registers, the shift `offset` applied once per lane from its hit count).
//
// EPI = 1 (resident run, tmd_nlist_run_vv): rows are cell-order slots and the state lives in slot order.
// When a row's force is complete, lanes 0 .. DIM-1 of its group each take one coordinate of the
// VelocityVerlet update (integrators.py:84-94): the closing half kick of this step and -- unless this is
// the last step of the run -- the opening half kick and the drift of the next one, the new 32-byte
// record for the next traversal (double buffer) and the Verlet-list displacement test.  Same roundings,
// in the same order, as tmd_vv_kick_kick_drift + nl_gather_check_kernel: bit-identical trajectories.
//
// EPI = 2 (slab run, VelocityVerlet): as 1, rows = the slots of this rank's cell layers, boundary records also stored into
// the neighbours' buffers, optional in-kernel step barrier.  EPI = 3 (slab run, other integrators): the slab's rows, forces
// by slot, NO integration (a separate pass integrates and pushes the halo).
template <int DIM, bool SINGLE, int G, bool WV, bool WF, int MINB, int EPI = 0>
__global__ void __launch_bounds__(kNlThreads, MINB) nl_force_kernel(const NlArgs a) {
    constexpr bool kInt = EPI == 1 || EPI == 2;  // the epilogue integrates
    constexpr bool kSlab = EPI >= 2;             // rows decided on the device (slab runs)
    static_assert(G * 4 == kBlockEntries, "the block layout of the lists is for four lanes per row atom");
    __shared__ double s_tab[SINGLE ? 1 : 6 * kMaxT2];
    __shared__ double s_red[10 * (kNlThreads / 32)];
    // per-thread virial partials (xx xy xz yy yz zz).  They are touched once per row (lane 0) and once
    // per pair that crosses a box face, so they live in shared memory: the twelve registers they would
    // hold go to a third record buffer (deeper gather pipeline) instead.
    __shared__ double s_w[6][kNlThreads];
    // EPI: what the epilogue of the current row needs (v, x, lattice shift, build position, 1/m, other forces) is
    // requested when the row starts and lands here while its pairs are evaluated
    __shared__ double s_epi[kInt ? 6 : 1][kInt ? kNlThreads : 1];
    constexpr int kRowsPerBlock = kNlThreads / G;
    const int tid = threadIdx.x;
#pragma unroll
    for (int w = 0; w < 6; ++w) s_w[w][tid] = 0.0;
    const int l = tid % G, grp = tid / G;
    // EPI = 2 (slab runs): the rows are the slots of a range of cells, known only after binning (see nl_list_kernel);
    // otherwise the two are kernel parameters (constant bank, no registers: this loop is at its register limit)
    const int64_t row_first = kSlab ? (int64_t)a.starts[a.row_cell_lo] : a.first;
    const int64_t row_count = kSlab ? (int64_t)(a.starts[a.row_cell_hi] - a.starts[a.row_cell_lo]) : a.count;
    // EPI = 2: which of this rank's slots are copied to which slots of the two neighbouring ranks
    int halo_lo_begin = 0, halo_lo_end = 0, halo_dest_lo = 0, halo_hi_begin = 0, halo_dest_hi = 0;
    if (EPI == 2) {
        halo_lo_begin = a.halo[0]; halo_lo_end = a.halo[1]; halo_dest_lo = a.halo[2];
        halo_hi_begin = a.halo[3]; halo_dest_hi = a.halo[5];
    }
    const bool whole = EPI ? false : (a.first == 0 && a.count == a.n);
    const int ntypes = a.ntypes;
    if (!SINGLE) {
        for (int k = tid; k < 6 * kMaxT2; k += kNlThreads) s_tab[k] = a.tab.c[k / kMaxT2][k % kMaxT2];
        __syncthreads();
    }
    if (blockIdx.x == 0 && tid == 0) {  // the gated build of this call is behind us in the stream
        int *gate_word = EPI ? a.flag_clear : a.nlflags;  // (EPI: this step's epilogues raise a.flag_next, another word)
        if (*gate_word) {
            *gate_word = 0;
            a.nlflags[2] += 1;
        }
        a.nlflags[3] += 1;
    }
    const uint32_t flags = a.flags;
    const double lx = a.grid.len[0], ly = a.grid.len[1], lz = a.grid.len[2];
    const double lj1 = a.tab.c[0][0], lj2 = a.tab.c[1][0], lj3 = a.tab.c[2][0], lj4 = a.tab.c[3][0],
                 off = a.tab.c[4][0], rc2 = a.tab.c[5][0];
    // per-lane totals over all rows of this block: energy, hit count (the shift is applied at the end),
    // virial (xx xy xz yy yz zz) = 2 F (x) x of the rows this lane leads - sum f (x) S of crossing pairs
    double vtot = 0.0;
    int hits = 0;
    // per-atom virial term F_i (x) x_i: lane p < DIM of a group accumulates row p of the upper triangle,
    // (p, p), (p, p+1), ... in registers; only the rare crossing corrections go through s_w
    double wreg[3] = {0.0, 0.0, 0.0};

    // Everything a row needs before its first pair can be evaluated; fetched one row ahead so that
    // the dependent chain (row -> list head -> record) of the NEXT row overlaps this row's arithmetic.
    struct Row {
        bool valid;
        int slot, nn;
        double4 me;
        uint4 e;  // this lane's entries of the first block
    };
    auto fetch = [&](int64_t base) {
        Row r;
        const int64_t row = base + grp;
        r.valid = row < row_count;
        r.slot = !r.valid ? 0 : (whole ? (int)row : ((EPI || a.row_mode) ? (int)(row_first + row) : a.slot_of[row_first + row]));
        r.me = a.xs[r.slot];
        r.nn = r.valid ? a.nneigh[row] : 0;
        // first entries of the row, whatever its length (cap >= 16: in bounds; unused words are never dereferenced)
        r.e = make_uint4(0u, 0u, 0u, 0u);
        if (r.valid) r.e = __ldg(reinterpret_cast<const uint4 *>(a.entries + (size_t)row * a.cap) + l);
        return r;
    };

    // Warp-uniform trip count: the group shuffles below need every lane of the warp in the loop.  Rows are dealt out
    // statically (block b takes row groups b, b + gridDim, ...: per-block (V, W) partials are then reproducible), except
    // on the steps of a slab run that need no (V, W): there every warp claims its next chunk of 32 / G rows from a
    // counter, so that no SM idles while another still has a whole iteration to go (the rows per rank are few).
    const int64_t stride = (int64_t)gridDim.x * kRowsPerBlock;
    const bool dynamic = kSlab && !a.epi_last && a.dynamic_rows;
    constexpr int kRowsPerWarp = 32 / G;
    const int64_t limit = kSlab ? row_count - (int64_t)(tid >> 5) * kRowsPerWarp : row_count;
    auto claim = [&]() -> int64_t {  // base of this warp's next chunk, such that row = base + grp
        int c = 0;
        if ((tid & 31) == 0) c = atomicAdd(a.work, 1);
        c = __shfl_sync(0xffffffffu, c, 0);
        return ((int64_t)c - (tid >> 5)) * kRowsPerWarp;
    };
    int64_t base = dynamic ? claim() : (int64_t)blockIdx.x * kRowsPerBlock, base_next = 0;
    Row next = fetch(base);
    for (; base < limit; base = base_next) {
        const Row cur = next;
        base_next = dynamic ? claim() : base + stride;
        if (base_next < limit) next = fetch(base_next);
        const int64_t row = base + grp;
        const bool valid = cur.valid;
        const int slot = cur.slot, nn = cur.nn;
        const double4 me = cur.me;
        const int ti = SINGLE ? 0 : (int)__double_as_longlong(me.w);
        double facc[3] = {0.0, 0.0, 0.0};
        if (kInt) {
            if (valid && l < DIM) {
                const int64_t e = (int64_t)slot * DIM + l;
                cp_async8(&s_epi[0][tid], a.vs + e);
                cp_async8(&s_epi[4][tid], a.imass_s + slot);
                if (a.fx) cp_async8(&s_epi[5][tid], a.fx + e);
                if (!a.epi_last) {
                    cp_async8(&s_epi[1][tid], a.xu + e);
                    cp_async8(&s_epi[2][tid], a.image_s + (int64_t)slot * 3 + l);
                    cp_async8(&s_epi[3][tid], reinterpret_cast<const double *>(a.xb + slot) + l);
                }
            }
            cp_async_commit();
        }

        // A candidate is handled in two halves.  consume(): the three subtractions that read the gathered
        // record (the scoreboard wait of its load happens here).  compute(): everything else.
        struct Delta {
            double d[3];
            int t;  // type of the neighbour (multi-type systems)
        };
        auto consume = [&](const double4 q) {
            Delta r;
            r.d[0] = me.x - q.x;
            r.d[1] = DIM > 1 ? me.y - q.y : 0.0;
            r.d[2] = DIM > 2 ? me.z - q.z : 0.0;
            r.t = SINGLE ? 0 : (int)__double_as_longlong(q.w);
            return r;
        };
        auto compute = [&](const Delta &dl, int code) {
            double d[3] = {dl.d[0], dl.d[1], dl.d[2]};
            double l1 = lj1, l2 = lj2, l3 = lj3, l4 = lj4, of = off, r2c = rc2;
            if (!SINGLE) {
                const int p = ti * ntypes + dl.t;
                l1 = s_tab[0 * kMaxT2 + p];
                l2 = s_tab[1 * kMaxT2 + p];
                l3 = s_tab[2 * kMaxT2 + p];
                l4 = s_tab[3 * kMaxT2 + p];
                of = s_tab[4 * kMaxT2 + p];
                r2c = s_tab[5 * kMaxT2 + p];
            }
            double S[3] = {0.0, 0.0, 0.0};
            const bool crossing = code != kHomeCode;  // the neighbour is a periodic image: atoms near a face only
            if (crossing) {
                S[0] = (code % 3 - 1) * lx;
                if (DIM > 1) S[1] = ((code / 3) % 3 - 1) * ly;
                if (DIM > 2) S[2] = (code / 9 - 1) * lz;
#pragma unroll
                for (int k = 0; k < DIM; ++k) d[k] -= S[k];
            }
            double rsq = d[0] * d[0];
            if (DIM > 1) rsq = fma(d[1], d[1], rsq);
            if (DIM > 2) rsq = fma(d[2], d[2], rsq);
            if (rsq < r2c) {                                                     // strict '<' (lennardjones.py:291-294)
                const double r2inv = rcp_fast(rsq);                              // :254
                const double r6inv = (r2inv * r2inv) * r2inv;                    // :255
                if (WV) {                                                        // :276-282
                    vtot = fma(r6inv, fma(l3, r6inv, -l4), vtot);
                    if (SINGLE) hits += 1;
                    else vtot -= of;
                }
                if (WF) {
                    const double flj = (r2inv * r6inv) * fma(l1, r6inv, -l2);    // :285-288
                    if (!crossing) {
#pragma unroll
                        for (int k = 0; k < DIM; ++k) facc[k] = fma(flj, d[k], facc[k]);  // :269-270
                    } else {
                        int w = 0;
#pragma unroll
                        for (int k = 0; k < 3; ++k) {
                            const double fk = k < DIM ? flj * d[k] : 0.0;
                            if (k < DIM) facc[k] += fk;
#pragma unroll
                            for (int b = k; b < 3; ++b, ++w)
                                if (b < DIM && S[b] != 0.0) s_w[w][tid] = fma(-fk, S[b], s_w[w][tid]);
                        }
                    }
                }
            }
        };

        if (nn <= a.cap) {
            // Software pipeline over half blocks (two rounds).  ptxas tracks every gather of this loop with
            // ONE scoreboard, so the wait in front of a record's first use waits for every load in flight: a
            // gather issued just before that wait would be waited for at once (ncu, r01p: 40 % of the stall
            // samples sat there).  Hence the order: consume the two records that are due (one wait), THEN
            // issue the two gathers of the next half block (and, once per block, the next entries), then
            // do the arithmetic of the two candidates while those loads fly.  `go` makes the gathers depend
            // on a consumed value so that the assembler cannot hoist them above the wait; it is false only
            // for one particular NaN payload.  k = logical index of this lane's candidate in round 0.
            const uint4 *blk = reinterpret_cast<const uint4 *>(a.entries + (size_t)row * a.cap) + l;
            int k = l;
            uint4 e = cur.e;
            double4 qa = me, qb = me;
            if (k < nn) qa = ld_record(a.xs + (e.x & kSlotMask));
            if (k + G < nn) qb = ld_record(a.xs + (e.y & kSlotMask));
            while (k < nn) {
                Delta da = consume(qa), db = consume(qb);
                bool go = __double2hiint(da.d[0]) != kNeverHi;
                blk += kBlockEntries / 4;
                uint4 en = e;
                if (go && k + 4 * G < nn) en = __ldg(blk);
                if (go && k + 2 * G < nn) qa = ld_record(a.xs + (e.z & kSlotMask));
                if (go && k + 3 * G < nn) qb = ld_record(a.xs + (e.w & kSlotMask));
                compute(da, (int)(e.x >> 27));
                if (k + G < nn) compute(db, (int)(e.y >> 27));
                if (k + 2 * G >= nn) break;
                da = consume(qa);
                db = consume(qb);
                go = __double2hiint(da.d[0]) != kNeverHi;
                if (go && k + 4 * G < nn) qa = ld_record(a.xs + (en.x & kSlotMask));
                if (go && k + 5 * G < nn) qb = ld_record(a.xs + (en.y & kSlotMask));
                compute(da, (int)(e.z >> 27));
                if (k + 3 * G < nn) compute(db, (int)(e.w >> 27));
                k += 4 * G;
                e = en;
            }
        } else {
            // list capacity exceeded at build time: walk the build-time cells (every pair now inside
            // rcut was inside rcut + skin then, hence in this stencil)
            const int cell = a.cell_of[(whole || a.row_mode || EPI) ? a.order[slot] : (int)(row_first + row)];
            const int c0 = cell % a.grid.nc[0], c1 = (cell / a.grid.nc[0]) % a.grid.nc[1],
                      c2 = cell / (a.grid.nc[0] * a.grid.nc[1]);
            for_each_run<DIM>(a.grid, a.starts, c1, c2, full_x_window{c0, a.grid.hx}, [&](int jbegin, int jend, int code) {
                for (int j = jbegin + l; j < jend; j += G)
                    if (j != slot) compute(consume(a.xs[j]), code);
            });
        }

        // ---- finish this row: F over the G lanes, then the per-atom virial term on lane 0 ----
        double f[3] = {0.0, 0.0, 0.0};
#pragma unroll
        for (int k = 0; k < DIM; ++k) f[k] = group_sum<G>(facc[k]);
        if (kInt) {
            static_assert(!kInt || G >= 3, "one lane per coordinate");
            // lane l < DIM owns coordinate l of this row
            const bool mine = valid && l < DIM;
            const int64_t e = (int64_t)slot * DIM + (l < DIM ? l : 0);
            const double fl = l == 0 ? f[0] : (l == 1 ? f[1] : f[2]);
            double fk = fl, v = 0.0, x = 0.0, bk = 0.0;
            cp_async_wait_all();  // this lane's own requests of the row start
            if (mine) {
                if (a.fx) fk = s_epi[5][tid] + fl;  // the system's other potentials were evaluated into fx before this launch
                v = s_epi[0][tid];
                const double im = s_epi[4][tid];
                a.fs[e] = fk;
                v = kick(v, fk, im, a.half_dt);  // integrators.py:94 of this step
                if (!a.epi_last) {
                    v = kick(v, fk, im, a.half_dt);                      // :88 of the next step
                    x = __dadd_rn(s_epi[1][tid], __dmul_rn(a.dt, v));    // :90
                    a.xu[e] = x;
                    x = __dsub_rn(x, s_epi[2][tid]);  // the record: what nl_gather_check_kernel forms, pos - image
                    bk = s_epi[3][tid];
                }
                a.vs[e] = v;
            }
            if (!a.epi_last) {  // block-uniform
                const int base = (tid & 31) & ~(G - 1);
                const double ry = __shfl_sync(0xffffffffu, x, base + (DIM > 1 ? 1 : 0));
                const double rz = __shfl_sync(0xffffffffu, x, base + (DIM > 2 ? 2 : 0));
                const double by = __shfl_sync(0xffffffffu, bk, base + (DIM > 1 ? 1 : 0));
                const double bz = __shfl_sync(0xffffffffu, bk, base + (DIM > 2 ? 2 : 0));
                if (l == 0 && valid) {
                    const double4 rec = make_double4(x, DIM > 1 ? ry : 0.0, DIM > 2 ? rz : 0.0, me.w);
                    a.xs_next[slot] = rec;
                    if (EPI == 2) {  // boundary layers: the neighbours keep copies (peer stores over NVLink)
                        if (slot < halo_lo_end) a.peer_lo_next[halo_dest_lo + (slot - halo_lo_begin)] = rec;
                        if (slot >= halo_hi_begin) a.peer_hi_next[halo_dest_hi + (slot - halo_hi_begin)] = rec;
                    }
                    double r = (x - bk) * (x - bk);
                    if (DIM > 1) r = fma(ry - by, ry - by, r);
                    if (DIM > 2) r = fma(rz - bz, rz - bz, r);
                    if (!(r <= a.half_skin2)) *a.flag_next = 1;  // also catches NaN
                }
            }
        }
        if (l == 0 && valid) {
            if (!kInt && (flags & TMD_WANT_F)) {
                const int atom = a.row_mode ? slot : (whole ? a.order[slot] : (int)(row_first + row));
                double *dst = a.force + (int64_t)atom * DIM;
#pragma unroll
                for (int k = 0; k < DIM; ++k) dst[k] = (flags & TMD_ACCUMULATE) ? dst[k] + f[k] : f[k];
            }
        }
        if (WF && (flags & TMD_WANT_W)) {
            const double xc[3] = {me.x - a.centre[0], me.y - a.centre[1], me.z - a.centre[2]};
            const double f2 = 2.0 * (l == 0 ? f[0] : (l == 1 ? f[1] : (l == 2 ? f[2] : 0.0)));
            const double x0 = l == 0 ? xc[0] : (l == 1 ? xc[1] : xc[2]);
            const double x1 = l == 0 ? xc[1] : (l == 1 ? xc[2] : 0.0);
            const double x2 = l == 0 ? xc[2] : 0.0;
            wreg[0] = fma(f2, x0, wreg[0]);
            if (DIM > 1) wreg[1] = fma(f2, x1, wreg[1]);
            if (DIM > 2) wreg[2] = fma(f2, x2, wreg[2]);
        }
    }

    // per-block share of (V, W) in the "every pair seen from both atoms, halve at the end" convention
    double red[10];
    red[0] = SINGLE ? vtot - hits * off : vtot;
    double wtot[6];
#pragma unroll
    for (int w = 0; w < 6; ++w) wtot[w] = s_w[w][tid];
    if (l == 0) { wtot[0] += wreg[0]; wtot[1] += wreg[1]; wtot[2] += wreg[2]; }
    if (l == 1 && DIM > 1) { wtot[3] += wreg[0]; wtot[4] += wreg[1]; }
    if (l == 2 && DIM > 2) wtot[5] += wreg[0];
    red[1] = wtot[0]; red[2] = wtot[1]; red[3] = wtot[2];
    red[4] = wtot[1]; red[5] = wtot[3]; red[6] = wtot[4];
    red[7] = wtot[2]; red[8] = wtot[4]; red[9] = wtot[5];
    if (!(flags & TMD_WANT_W)) {
#pragma unroll
        for (int k = 1; k < 10; ++k) red[k] = 0.0;
    }
    if (EPI == 2) __threadfence_system();  // this thread's stores into the neighbours' buffers are ahead of the barrier below
    block_sum<10, kNlThreads>(red, s_red);
    // The block that finishes last adds up the partials of all blocks (no second launch).  The order of
    // the sum is fixed by block index and thread index, not by who arrives when: deterministic.
    // Resident runs (EPI): nobody can look at (V, W) between the steps of a run, so only the closing step reduces
    // them -- the reduction is a serial tail of ~10 us during which all other SMs idle.
    const bool reduce_now = !(EPI && !a.epi_last);
    const bool barrier_here = EPI == 2 && a.bar != nullptr;
    if (!reduce_now && !barrier_here) {
        if (tid == 0) {
#pragma unroll
            for (int k = 0; k < 10; ++k) a.partial_vw[(size_t)blockIdx.x * 10 + k] = red[k];
        }
        return;  // (emulated ranks of a slab run: the separate barrier kernel that follows resets the row counter)
    }
    __shared__ int s_last;
    if (tid == 0) {
#pragma unroll
        for (int k = 0; k < 10; ++k) a.partial_vw[(size_t)blockIdx.x * 10 + k] = red[k];
        __threadfence();
        s_last = atomicInc(reinterpret_cast<unsigned *>(&a.nlflags[7]), gridDim.x - 1) == gridDim.x - 1;  // wraps to 0
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    if (kSlab) {
        if (tid == 0) *a.work = 0;  // every block is past its loop: the counter is ready for the next launch
        if (barrier_here && tid < 32) nl_step_barrier(*a.bar, a.bar_gate_out, a.bar_has_cond, a.bar_cond, tid);
        if (!reduce_now) return;
        __syncthreads();
    }
    // thread t sums blocks t, t + 128, ... for all ten components (ten independent chains of coalesced loads),
    // then the block sums its 128 threads in a fixed order
    double fin[10];
#pragma unroll
    for (int k = 0; k < 10; ++k) fin[k] = 0.0;
    for (int b = tid; b < (int)gridDim.x; b += kNlThreads) {
        const double *src = a.partial_vw + (size_t)b * 10;
#pragma unroll
        for (int k = 0; k < 10; ++k) fin[k] += __ldcg(src + k);
    }
    block_sum<10, kNlThreads>(fin, s_red);
    if (tid == 0) {
#pragma unroll
        for (int k = 0; k < 10; ++k) {
            // every pair was seen from both atoms: halve; quantities that were not asked for stay untouched
            if (k == 0 ? (flags & TMD_WANT_V) : (flags & TMD_WANT_W))
                a.vw[k] = ((flags & TMD_ACCUMULATE) ? a.vw[k] : 0.0) + 0.5 * fin[k];
        }
    }
}

// ---- resident runs: atom order <-> slot order ---------------------------------------------------------
// gate == NULL: always; otherwise only when *gate != 0 (the rebuild of a step, decided on the device)
template <int DIM>
__global__ void __launch_bounds__(256) md_import_kernel(const int *__restrict__ order, int64_t n,
                                                        const double *__restrict__ pos, const double *__restrict__ vel,
                                                        const double *__restrict__ force,
                                                        const double *__restrict__ imass,
                                                        const double *__restrict__ image, double *__restrict__ xu,
                                                        double *__restrict__ vs, double *__restrict__ fs,
                                                        double *__restrict__ imass_s, double *__restrict__ image_s,
                                                        const int *gate) {
    if (gate && *gate == 0) return;
    const int64_t slot = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= n) return;
    const int64_t atom = order[slot];
#pragma unroll
    for (int k = 0; k < DIM; ++k) {
        xu[slot * DIM + k] = pos[atom * DIM + k];
        vs[slot * DIM + k] = vel[atom * DIM + k];
        if (force) fs[slot * DIM + k] = force[atom * DIM + k];
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) image_s[slot * 3 + k] = image[atom * 3 + k];
    imass_s[slot] = imass[atom];
}

template <int DIM>
__global__ void __launch_bounds__(256) md_export_kernel(const int *__restrict__ order, int64_t n,
                                                        const double *__restrict__ xu, const double *__restrict__ vs,
                                                        const double *__restrict__ fs, double *__restrict__ pos,
                                                        double *__restrict__ vel, double *__restrict__ force,
                                                        const int *gate) {
    if (gate && *gate == 0) return;
    const int64_t slot = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= n) return;
    const int64_t atom = order[slot];
#pragma unroll
    for (int k = 0; k < DIM; ++k) {
        pos[atom * DIM + k] = xu[slot * DIM + k];
        vel[atom * DIM + k] = vs[slot * DIM + k];
        if (force) force[atom * DIM + k] = fs[slot * DIM + k];
    }
}

// opening half kick + drift of the first step of a run (integrators.py:88-90) on slot-ordered state, the
// records the first traversal reads and the displacement test against the build positions
template <int DIM>
__global__ void __launch_bounds__(256) md_pre_kernel(int64_t n, double *__restrict__ xu, double *__restrict__ vs,
                                                     const double *__restrict__ fs,
                                                     const double *__restrict__ imass_s,
                                                     const double *__restrict__ image_s,
                                                     const double4 *__restrict__ xb, double4 *__restrict__ xs,
                                                     double dt, double half, double half_skin2, int *flag) {
    const int64_t slot = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= n) return;
    const double im = imass_s[slot];
    const double4 b = xb[slot];
    double rec[3] = {0.0, 0.0, 0.0};
#pragma unroll
    for (int k = 0; k < DIM; ++k) {
        const double v = kick(vs[slot * DIM + k], fs[slot * DIM + k], im, half);
        const double x = __dadd_rn(xu[slot * DIM + k], __dmul_rn(dt, v));
        vs[slot * DIM + k] = v;
        xu[slot * DIM + k] = x;
        rec[k] = __dsub_rn(x, image_s[slot * 3 + k]);
    }
    xs[slot] = make_double4(rec[0], rec[1], rec[2], b.w);
    double r = (rec[0] - b.x) * (rec[0] - b.x);
    if (DIM > 1) r = fma(rec[1] - b.y, rec[1] - b.y, r);
    if (DIM > 2) r = fma(rec[2] - b.z, rec[2] - b.z, r);
    if (!(r <= half_skin2)) *flag = 1;
}

// LangevinInertia on slot-ordered state (resident run, tmd_nlist_run_langevin; integrators.py:265-282).  mode bit 0: the
// closing update v = v' + b2 F of the step whose forces were just computed (:281); bit 1: the opening update of the next
// step (:270-279) with the noise of a coordinate addressed by its ATOM index (draws q_base + 2 (atom dim + k) and + 1, as the
// reference consumes them and as inertia_pre_kernel does), the records of the next traversal and the displacement test.
template <int DIM>
__global__ void __launch_bounds__(256) md_langevin_kernel(int64_t n, const int *__restrict__ order, double *__restrict__ xu,
                                                          double *__restrict__ vs, const double *__restrict__ fs,
                                                          const double *__restrict__ imass_s,
                                                          const double *__restrict__ image_s,
                                                          const double4 *__restrict__ xb, double4 *__restrict__ xs,
                                                          tmd_langevin_consts c, uint64_t key, uint64_t q_base, int mode,
                                                          double half_skin2, int *flag) {
    const int64_t slot = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= n) return;
    const double im = imass_s[slot];
    const InertiaPerParticle p = inertia_constants(c, im);
    const int64_t atom = order[slot];
    const double4 b = xb[slot];
    double rec[3] = {0.0, 0.0, 0.0};
    // three dimensions, even stream position: the six normals of this atom from two Philox blocks
    double z6[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    const uint64_t q0 = q_base + 2ull * (uint64_t)(atom * DIM);
    const bool six = DIM == 3 && (mode & 2) && (q0 & 1ull) == 0;
    if (six) six_normals(key, q0, z6);
#pragma unroll
    for (int k = 0; k < DIM; ++k) {
        const double f = fs[slot * DIM + k];
        double v = vs[slot * DIM + k];
        double x = xu[slot * DIM + k];
        if (mode & 1) v = __dadd_rn(v, __dmul_rn(__dmul_rn(c.b2f, im), f));  // exactly inertia_post_kernel
        if (mode & 2) {
            double n0, n1;
            if (six) { n0 = z6[2 * k]; n1 = z6[2 * k + 1]; }
            else two_normals(key, q_base + 2ull * (uint64_t)(atom * DIM + k), true, n0, n1);
            const double xr = __dmul_rn(p.l00, n0);  // norm @ cho.T  (integrators.py:315)
            const double vr = __dadd_rn(__dmul_rn(p.l10, n0), __dmul_rn(p.l11, n1));
            inertia_pre_update(x, v, f, xr, vr, c, p);
            xu[slot * DIM + k] = x;
        }
        vs[slot * DIM + k] = v;
        rec[k] = __dsub_rn(x, image_s[slot * 3 + k]);
    }
    if (!(mode & 2)) return;
    xs[slot] = make_double4(rec[0], rec[1], rec[2], b.w);
    double r = (rec[0] - b.x) * (rec[0] - b.x);
    if (DIM > 1) r = fma(rec[1] - b.y, rec[1] - b.y, r);
    if (DIM > 2) r = fma(rec[2] - b.z, rec[2] - b.z, r);
    if (!(r <= half_skin2)) *flag = 1;
}

// after an ungated build outside a traversal: what the traversal's block 0 does with the rebuild flag
__global__ void md_built_kernel(int *nlflags) {
    if (nlflags[0]) {
        nlflags[0] = 0;
        nlflags[2] += 1;
    }
}

// DoubleWellPair (well.py:230-286) as the SECOND potential of a resident run: the few atoms of the active types, one
// thread each in ONE block, read their positions from the slot-ordered state, and leave their forces in fx (slot order),
// which the traversal's epilogue adds to the Lennard-Jones force -- System.potential_and_force's sum (system.py:58-76).
// The arithmetic is dwpair_kernel's (wells.cu), term for term: same roundings, same order of partners -> the same bits.
// prev_slot remembers where this thread wrote last (slots change at every rebuild): those entries are cleared first.
struct MdDwPair {
    const int32_t *idx0, *idx1;  // atoms of types[0] / types[1] (ascending)
    int n0, n1, same_type;
    double rwidth, width2, height;
    double len[3], ilen[3], half[3];
    int per[3];
    int *prev_slot;              // [n0 (+ n1)], -1 = nothing written yet
    double *partials;            // [10]: V and W of the pairs, every pair seen from both ends (reduced with scale 0.5)
};

template <int DIM>
__global__ void __launch_bounds__(128) md_dwpair_kernel(const MdDwPair w, const int *__restrict__ slot_of,
                                                        const double *__restrict__ xu, double *__restrict__ fx,
                                                        uint32_t flags) {
    __shared__ double smem[10 * 4];
    double acc[10];
#pragma unroll
    for (int k = 0; k < 10; ++k) acc[k] = 0.0;
    const int nown = w.same_type ? w.n0 : w.n0 + w.n1;
    const int t = threadIdx.x;
    if (t < nown) {
        const int ps = w.prev_slot[t];
        if (ps >= 0) {
#pragma unroll
            for (int k = 0; k < DIM; ++k) fx[(int64_t)ps * DIM + k] = 0.0;
        }
    }
    __syncthreads();
    if (t < nown) {
        const bool in0 = t < w.n0;
        const int me = in0 ? w.idx0[t] : w.idx1[t - w.n0];
        const int32_t *other = w.same_type ? w.idx0 : (in0 ? w.idx1 : w.idx0);
        const int nother = w.same_type ? w.n0 : (in0 ? w.n1 : w.n0);
        const int myslot = slot_of[me];
        double xi[3] = {0.0, 0.0, 0.0}, fi[3] = {0.0, 0.0, 0.0};
#pragma unroll
        for (int k = 0; k < DIM; ++k) xi[k] = xu[(int64_t)myslot * DIM + k];
        const double height4 = __dmul_rn(4.0, w.height);  // well.py:212
        for (int p = 0; p < nother; ++p) {
            const int you = other[p];
            if (you == me) continue;
            const bool first = me < you;  // the reference visits the pair as (i, j) with i < j (particles.py:129-132)
            const int yourslot = slot_of[you];
            double d[3] = {0.0, 0.0, 0.0};
            double rsq = 0.0;
#pragma unroll
            for (int k = 0; k < DIM; ++k) {
                const double xj = xu[(int64_t)yourslot * DIM + k];
                double dk = first ? __dsub_rn(xi[k], xj) : __dsub_rn(xj, xi[k]);
                // Box.pbc_dist: strict '>' half box, stored reciprocal (box.py:376-378)
                if (w.per[k] && fabs(dk) > w.half[k])
                    dk = __dsub_rn(dk, __dmul_rn(rint(__dmul_rn(dk, w.ilen[k])), w.len[k]));
                d[k] = dk;
            }
#pragma unroll
            for (int k = 0; k < DIM; ++k) rsq = (k == 0) ? __dmul_rn(d[k], d[k]) : __dadd_rn(rsq, __dmul_rn(d[k], d[k]));
            const double delr = __dsqrt_rn(rsq);
            const double diff = __dsub_rn(delr, w.rwidth);
            const double q = __ddiv_rn(__dmul_rn(diff, diff), w.width2);
            if (flags & TMD_WANT_V) {  // height*(1 - diff^2/width2)^2  (well.py:250-260)
                const double om = __dsub_rn(1.0, q);
                acc[0] += __dmul_rn(w.height, __dmul_rn(om, om));
            }
            // height4*(1 - diff^2/width2)*(diff/width2) * delta/delr  (well.py:279-282)
            const double mag = __dmul_rn(__dmul_rn(height4, __dsub_rn(1.0, q)), __ddiv_rn(diff, w.width2));
#pragma unroll
            for (int k = 0; k < DIM; ++k) {
                const double fk = __ddiv_rn(__dmul_rn(mag, d[k]), delr);
                fi[k] += first ? fk : -fk;
#pragma unroll
                for (int m = 0; m < DIM; ++m) acc[1 + k * 3 + m] += __dmul_rn(fk, d[m]);  // np.outer (well.py:285)
            }
        }
#pragma unroll
        for (int k = 0; k < DIM; ++k) fx[(int64_t)myslot * DIM + k] = fi[k];  // (dwpair_kernel: 0 + fi[k])
        w.prev_slot[t] = myslot;
    }
    block_sum<10, 128>(acc, smem);  // (as dwpair_kernel's only block when there are at most 128 active atoms)
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < 10; ++k) w.partials[k] = acc[k];
    }
}

// ---- host side -----------------------------------------------------------------------------------

static void nl_release(tmd_nlist *nl) {
    delete nl->saved;
    nl->saved = nullptr;
    cudaFree(nl->flags);
    cudaFree(nl->entries);
    cudaFree(nl->nneigh);
    cudaFree(nl->ints);
    cudaFree(nl->wrapped);
    cudaFree(nl->xb);
    cudaFree(nl->xbf);
    cudaFree(nl->xs);
    cudaFree(nl->image);
    cudaFree(nl->xs2);
    cudaFree(nl->md);
    cudaFree(nl->dw_scratch);
    nl->dw_scratch = nullptr;
    nl->xs2 = nullptr;
    nl->md = nullptr;
    nl->cap_md = 0;
    nl->xs = nullptr;
    nl->flags = nullptr;
    nl->entries = nullptr;
    nl->nneigh = nullptr;
    nl->ints = nullptr;
    nl->wrapped = nl->xb = nl->xs = nullptr;
    nl->xbf = nullptr;
    nl->image = nullptr;
    nl->cap_n = nl->cap_rows = nl->cap_cells = 0;
    nl->cap = 0;
}

static int nl_env_int(const char *name, int fallback) {
    const char *env = getenv(name);
    return env ? atoi(env) : fallback;
}

// stencil half width: cells of edge >= (rcut + skin) / H (TMD_NL_H = 1 | 2 | 3; default 1).  Finer
// cells halve the candidates a rebuild tests but cut them into many short runs; measured on B200
// (N = 1M, rho = 0.8) H = 1 rebuilds fastest and the traversal does not care.
static int nl_half_width() {
    static int h = 0;
    if (h == 0) {
        const int v = nl_env_int("TMD_NL_H", 1);
        h = (v >= 1 && v <= 3) ? v : 1;
    }
    return h;
}

static int nl_x_refinement() {
    static int f = 0;
    if (f == 0) {
        const int v = nl_env_int("TMD_NL_FX", 4);
        f = (v == 1 || v == 2 || v == 4 || v == 8) ? v : 4;
    }
    return f;
}

// Fine grid for `reach` = rcut + skin: false when the box is not fully periodic or a direction holds
// fewer than 2H+1 cells (the stencil would then see an atom twice; use tmd_lj_allpairs).
static bool nl_make_grid(const tmd_box *box, double reach, int64_t n, NlGrid &g) {
    if (!(reach > 0.0) || n < 2 || n >= (1ll << 27)) return false;
    const int h = nl_half_width();
    const double edge = reach / h * (1.0 + 1e-6);  // slack: an atom on a face may land in either cell
    long long ncells = 1;
    g.h = h;
    for (int k = 0; k < 3; ++k) {
        g.nc[k] = 1;
        g.len[k] = g.ilen[k] = g.scale[k] = 0.0;
    }
    for (int k = 0; k < box->dim; ++k) {
        if (!box->periodic[k]) return false;
        const double len = box->length[k];
        if (!(len > 0.0) || len != len || len > 1e300) return false;
        const double q = floor(len / edge);
        if (q < 2.0 * h + 1.0) return false;
        g.nc[k] = q > 1024.0 ? 1024 : (int)q;
        g.len[k] = len;
        g.ilen[k] = box->ilength[k];
        g.scale[k] = g.nc[k] / len;
        ncells *= g.nc[k];
    }
    // Cells are cut finer along x (TMD_NL_FX = 1 | 2 | 4 | 8, default 4): the cells of a stencil row are consecutive
    // in slot order along x, so a finer cut keeps every row ONE run of slots while the list build can clip the run to
    // the x window that can hold neighbours at all (nl_list_kernel): ~300 instead of ~490 candidates per atom.
    const int fx = nl_x_refinement();
    g.hx = h * fx;
    if (fx > 1) {
        const long long fine = (long long)g.nc[0] * fx;
        if (fine <= 4096) {
            ncells = ncells / g.nc[0] * fine;
            g.nc[0] = (int)fine;
            g.scale[0] = g.nc[0] / g.len[0];
        } else {
            g.hx = h;
        }
    }
    if (ncells > (1ll << 28)) return false;
    g.ncells = (int)ncells;
    return true;
}

// resident traversal blocks per SM (TMD_NL_MINB = 5 .. 8): the register budget of a thread is 65536 /
// (128 MINB); see profiles/r01q_sweep_minb.txt
static int nl_min_blocks() {
    static int m = 0;
    if (m == 0) {
        const int v = nl_env_int("TMD_NL_MINB", 5);
        m = (v >= 5 && v <= 8) ? v : 5;
    }
    return m;
}

template <int DIM, int MINB>
static void nl_launch_force_m(const NlArgs &a, int blocks, cudaStream_t s) {
    const bool wv = a.flags & TMD_WANT_V, wf = a.flags & (TMD_WANT_F | TMD_WANT_W);
#define TMD_NL_LAUNCH(SINGLE, WV, WF) \
    ::tmd::count_launch(), nl_force_kernel<DIM, SINGLE, 4, WV, WF, MINB><<<blocks, kNlThreads, 0, s>>>(a)
    if (a.ntypes == 1 || a.uniform_table) {
        if (wv && wf) TMD_NL_LAUNCH(true, true, true);
        else if (wf) TMD_NL_LAUNCH(true, false, true);
        else TMD_NL_LAUNCH(true, true, false);
    } else {
        if (wv && wf) TMD_NL_LAUNCH(false, true, true);
        else if (wf) TMD_NL_LAUNCH(false, false, true);
        else TMD_NL_LAUNCH(false, true, false);
    }
#undef TMD_NL_LAUNCH
}

// resident run: V, F and W every step (what VelocityVerlet's potential_and_force call computes)
template <int DIM>
static void nl_launch_force_epi(const NlArgs &a, int blocks, cudaStream_t s) {
    ::tmd::count_launch();
    if (a.ntypes == 1 || a.uniform_table) nl_force_kernel<DIM, true, 4, true, true, 5, 1><<<blocks, kNlThreads, 0, s>>>(a);
    else nl_force_kernel<DIM, false, 4, true, true, 5, 1><<<blocks, kNlThreads, 0, s>>>(a);
}

template <int DIM>
static void nl_launch_force(const NlArgs &a, int blocks, cudaStream_t s) {
    if (DIM == 3) {  // the occupancy variants exist for the 3-D kernel only
        switch (nl_min_blocks()) {
            case 6: nl_launch_force_m<DIM, 6>(a, blocks, s); return;
            case 7: nl_launch_force_m<DIM, 7>(a, blocks, s); return;
            case 8: nl_launch_force_m<DIM, 8>(a, blocks, s); return;
            default: break;
        }
    }
    nl_launch_force_m<DIM, 5>(a, blocks, s);
}

// persistent traversal grid: resident blocks per SM x SMs, never more blocks than row groups
static int nl_force_blocks(int64_t count, int dim) {
    const int g = 4;
    const int64_t groups = (count + (kNlThreads / g) - 1) / (kNlThreads / g);
    const int64_t want = (int64_t)num_sms() * (dim == 3 ? nl_min_blocks() : 5);
    return (int)(groups < want ? groups : want);
}

template <int DIM>
static int nl_launch_build(tmd_nlist *nl, NlArgs &a, cudaStream_t s) {
    if (nl->bin_blocks[DIM - 1] == 0) {
        int per_sm = 0;
        TMD_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, nl_bin_kernel<DIM>, kBinThreads, 0));
        if (per_sm < 1) per_sm = 1;
        nl->bin_blocks[DIM - 1] = per_sm * num_sms();
    }
    if (nl->build_blocks[DIM - 1] == 0) {
        int per_sm = 0;
        TMD_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, nl_list_kernel<DIM>, kNlThreads, 0));
        if (per_sm < 1) per_sm = 1;
        nl->build_blocks[DIM - 1] = per_sm * num_sms();
    }
    void *args[] = {(void *)&a};
    ::tmd::count_launch();
    TMD_CUDA(cudaLaunchCooperativeKernel((const void *)nl_bin_kernel<DIM>, dim3(nl->bin_blocks[DIM - 1]),
                                         dim3(kBinThreads), args, 0, s));
    ::tmd::count_launch(), nl_list_kernel<DIM><<<nl->build_blocks[DIM - 1], kNlThreads, 0, s>>>(a);
    return TMD_OK;
}

}  // namespace tmd

using namespace tmd;

extern "C" {

int tmd_nlist_create(tmd_nlist **out, double skin) {
    TMD_REQUIRE(out != nullptr, "out is NULL");
    TMD_REQUIRE(skin > 0.0 && skin < 1e6, "skin must be positive");
    TMD_CHECK(require_device());
    tmd_nlist *nl = new (std::nothrow) tmd_nlist();
    if (!nl) return TMD_ERR_NOMEM;
    memset(nl, 0, sizeof(*nl));
    nl->skin = skin;
    nl->n = -1;
    TMD_CUDA(cudaGetDevice(&nl->device));
    *out = nl;
    return TMD_OK;
}

int tmd_nlist_destroy(tmd_nlist *nl) {
    if (!nl) return TMD_OK;
    nl_release(nl);
    delete nl;
    return TMD_OK;
}

int tmd_nlist_set_timing(tmd_nlist *nl, int32_t on) {
    TMD_REQUIRE(nl != nullptr, "nl is NULL");
    nl->timing = on ? 1 : 0;
    nl->traversal_ms = 0.0;
    nl->traversals = 0;
    return TMD_OK;
}

int tmd_nlist_timing(tmd_nlist *nl, double *traversal_ms, int64_t *traversals) {
    TMD_REQUIRE(nl != nullptr && traversal_ms && traversals, "NULL argument");
    *traversal_ms = nl->traversal_ms;
    *traversals = nl->traversals;
    return TMD_OK;
}

int tmd_nlist_stats(tmd_nlist *nl, int64_t *builds, int64_t *calls, int32_t *overflow) {
    TMD_REQUIRE(nl != nullptr, "nl is NULL");
    int host[4] = {0, 0, 0, 0};
    if (nl->flags) TMD_CUDA(cudaMemcpy(host, nl->flags, sizeof(host), cudaMemcpyDeviceToHost));
    if (builds) *builds = host[2];
    if (calls) *calls = host[3];
    if (overflow) *overflow = host[1];
    return TMD_OK;
}

int tmd_lj_nlist_usable(int64_t n, const tmd_box *box, const double *table, int32_t ntypes, double skin,
                        int *usable) {
    TMD_REQUIRE(usable != nullptr, "usable is NULL");
    TMD_REQUIRE(box && box->dim >= 1 && box->dim <= 3, "box->dim must be 1..3");
    TMD_REQUIRE(skin > 0.0, "skin must be positive");
    *usable = 0;
    LjTable tab;
    double rc2max = 0.0;
    TMD_CHECK(load_lj_table(table, ntypes, tab, rc2max));
    NlGrid g;
    if (nl_make_grid(box, sqrt(rc2max) + skin, n, g)) *usable = 1;
    return TMD_OK;
}

// Everything a call needs apart from pos / tidx / flags / force: coefficient table, grid, (re)allocation,
// signature check (a change forces a rebuild), pointers.  `same` = the lists of the previous call
// are still for this system and row range.
static int nl_prepare(tmd_nlist *nl, int64_t n, const tmd_box *box, const double *table, int32_t ntypes,
                      int64_t first, int64_t count, int row_mode, cudaStream_t s, NlArgs &a, bool &same,
                      const int32_t *tidx_now = nullptr) {
    const int dim = box->dim;
    double rc2max = 0.0;
    TMD_CHECK(load_lj_table(table, ntypes, a.tab, rc2max));
    // e.g. a solute that differs from the solvent only in its bonded terms (examples/movie-doublewell-pair.py
    // at scale): identical Lennard-Jones coefficients for every type pair need no table look-ups
    a.uniform_table = 1;
    for (int c = 0; c < 6; ++c)
        for (int p = 1; p < ntypes * ntypes; ++p)
            if (memcmp(&a.tab.c[c][p], &a.tab.c[c][0], sizeof(double)) != 0) a.uniform_table = 0;
    const double reach = sqrt(rc2max) + nl->skin;
    if (!nl_make_grid(box, reach, n, a.grid)) {
        set_error("tmd_lj_nlist: needs a fully periodic box with >= %d cells of edge (rcut+skin)/%d per direction "
                  "and 2 <= n < 2^27; use tmd_lj_allpairs", 2 * nl_half_width() + 1, nl_half_width());
        return TMD_ERR_UNSUPPORTED;
    }

    // ---- (re)allocate and decide whether the signature changed ----
    bool fresh = false;
    const int ncells = a.grid.ncells;
    if (n > nl->cap_n || count > nl->cap_rows || ncells > nl->cap_cells) {
        TMD_CUDA(cudaStreamSynchronize(s));
        nl_release(nl);
        // expected neighbours within reach at this density, +35 % and a floor, rounded to whole blocks
        double volume = 1.0;
        for (int k = 0; k < dim; ++k) volume *= box->length[k];
        const double sphere = dim == 3 ? 4.18879020478639 * reach * reach * reach
                                       : (dim == 2 ? 3.14159265358979 * reach * reach : 2.0 * reach);
        int cap = (int)(1.35 * (double)n / volume * sphere) + 24;
        cap = (cap + kBlockEntries - 1) & ~(kBlockEntries - 1);
        if (cap > 1024) cap = 1024;
        nl->cap = cap;
        nl->cap_n = n;
        nl->cap_rows = count;
        nl->cap_cells = ncells;
        const size_t n_chunks = ((size_t)ncells + kScanChunk - 1) / kScanChunk;
        const size_t n_ints = (size_t)2 * (ncells + 1) + n_chunks + (size_t)5 * n;
        // all or nothing: a failed allocation must not leave capacities behind that a retried call would trust
        const size_t xbf_bytes = ((size_t)(n + 1) / 2 + kChunk / 2) * 2 * sizeof(float4);
        bool ok = cudaMalloc(&nl->flags, kNlFlagWords * sizeof(int)) == cudaSuccess &&
                  cudaMalloc(&nl->entries, (size_t)nl->cap * count * sizeof(unsigned)) == cudaSuccess &&
                  cudaMalloc(&nl->nneigh, (size_t)count * sizeof(int)) == cudaSuccess &&
                  cudaMalloc(&nl->ints, n_ints * sizeof(int)) == cudaSuccess &&
                  cudaMalloc(&nl->wrapped, (size_t)n * sizeof(double4)) == cudaSuccess &&
                  cudaMalloc(&nl->xb, (size_t)n * sizeof(double4)) == cudaSuccess &&
                  cudaMalloc(&nl->xbf, xbf_bytes) == cudaSuccess &&  // pair records; the list filter reads eight at a time: padded
                  (nl->xs_ext || cudaMalloc(&nl->xs, (size_t)n * sizeof(double4)) == cudaSuccess) &&
                  cudaMalloc(&nl->image, (size_t)n * 3 * sizeof(double)) == cudaSuccess;
        if (!ok) {
            cudaGetLastError();
            nl_release(nl);
            nl->n = -1;
            set_error("tmd_lj_nlist: out of device memory for the neighbour list of %lld atoms", (long long)n);
            return TMD_ERR_NOMEM;
        }
        TMD_CUDA(cudaMemsetAsync(nl->flags, 0, kNlFlagWords * sizeof(int), s));
        TMD_CUDA(cudaMemsetAsync(nl->ints, 0, n_ints * sizeof(int), s));  // counts[] must start at zero
        TMD_CUDA(cudaMemsetAsync(nl->xbf, 0, xbf_bytes, s));
        fresh = true;
    }
    same = !fresh && nl->n == n && nl->first == first && nl->count == count && nl->dim == dim &&
           nl->ntypes == ntypes && nl->reach == reach && nl->row_mode == row_mode && nl->tidx == tidx_now;
    for (int k = 0; k < dim && same; ++k) same = nl->len[k] == box->length[k];
    if (!same) {
        nl->n = n;
        nl->first = first;
        nl->count = count;
        nl->dim = dim;
        nl->ntypes = ntypes;
        nl->reach = reach;
        nl->row_mode = row_mode;
        nl->tidx = tidx_now;
        for (int k = 0; k < 3; ++k) nl->len[k] = k < dim ? box->length[k] : 0.0;
        // force a build (counts[] is zero: the scan of every build re-zeroes it)
        TMD_CUDA(cudaMemsetAsync(nl->flags, 1, sizeof(int), s));
    }

    a.pos = nullptr;
    a.tidx = nullptr;
    a.n = n;
    a.first = first;
    a.count = count;
    a.row_mode = row_mode;
    a.ntypes = ntypes;
    a.flags = 0;
    a.force = nullptr;
    a.partial_vw = nullptr;
    a.nlflags = nl->flags;
    a.gate = nl->flags;
    a.xu = a.vs = a.fs = nullptr;
    a.fx = a.imass_s = a.image_s = nullptr;
    a.xs_next = nullptr;
    a.flag_next = a.flag_clear = nullptr;
    a.dt = a.half_dt = 0.0;
    a.epi_last = 0;
    a.md_rebuild = 0;
    a.pos_w = a.vel_w = nullptr;
    a.imass = nullptr;
    a.dyn = nullptr;
    a.row_cell_lo = a.row_cell_hi = 0;
    a.gid = nullptr;
    a.halo = nullptr;
    a.peer_lo_next = a.peer_hi_next = nullptr;
    a.work = nullptr;
    a.dynamic_rows = 0;
    a.bar = nullptr;
    a.bar_gate_out = a.bar_has_cond = 0;
    {
        const size_t cells1 = (size_t)nl->cap_cells + 1;
        const size_t chunks = ((size_t)nl->cap_cells + kScanChunk - 1) / kScanChunk;
        a.counts = nl->ints;
        a.starts = nl->ints + cells1;
        a.chunk_sums = nl->ints + 2 * cells1;
        a.cell_of = nl->ints + 2 * cells1 + chunks;
        a.rank_of = a.cell_of + nl->cap_n;
        a.order = a.cell_of + 2 * nl->cap_n;
        a.slot_of = a.cell_of + 3 * nl->cap_n;
        a.order_tmp = a.cell_of + 4 * nl->cap_n;
    }
    a.wrapped = nl->wrapped;
    a.xb = nl->xb;
    a.xbf = nl->xbf;
    a.xs = nl->xs_ext ? nl->xs_ext : nl->xs;
    a.image = nl->image;
    a.entries = nl->entries;
    a.nneigh = nl->nneigh;
    a.cap = nl->cap;
    {
        // single-precision filter: coordinates in [0, L) (+- L after folding the shift) carry an
        // absolute error <= 5 * 2^-24 * L per component, i.e. < 2^-20 * Lmax on the distance
        double lmax = 0.0;
        for (int k = 0; k < dim; ++k) lmax = fmax(lmax, box->length[k]);
        const double rf = reach + lmax * (1.0 / 1048576.0);
        a.reach2f = (float)(rf * rf * (1.0 + 1e-6));
    }
    a.half_skin2 = 0.25 * nl->skin * nl->skin;
    for (int k = 0; k < 3; ++k) a.centre[k] = k < dim ? 0.5 * box->length[k] : 0.0;
    return TMD_OK;
}

static int nl_traverse(NlArgs &a, int dim, uint32_t flags, double *force, double *vw, cudaStream_t s) {
    a.flags = flags;
    a.force = force;
    a.vw = vw;
    const int fblocks = nl_force_blocks(a.count, dim);
    TMD_CHECK(arena_get(SLOT_PARTIAL_VW, (size_t)fblocks * 10 * sizeof(double), (void **)&a.partial_vw));
    if (dim == 1) nl_launch_force<1>(a, fblocks, s);
    else if (dim == 2) nl_launch_force<2>(a, fblocks, s);
    else nl_launch_force<3>(a, fblocks, s);
    TMD_CUDA(cudaGetLastError());
    return TMD_OK;  // the traversal reduces (V, W) itself
}

int tmd_lj_nlist(tmd_nlist *nl, const double *pos, const int32_t *tidx, int64_t n, const tmd_box *box,
                 const double *table, int32_t ntypes, uint32_t flags, int64_t first, int64_t count,
                 double *force, double *vw, tmd_stream_t stream) {
    TMD_REQUIRE(nl != nullptr, "nl is NULL");
    TMD_REQUIRE(box && box->dim >= 1 && box->dim <= 3, "box->dim must be 1..3");
    TMD_REQUIRE(n >= 0 && first >= 0 && count >= 0 && first + count <= n, "rows [first, first+count) must lie in [0, n)");
    TMD_REQUIRE(pos || n == 0, "pos is NULL");
    TMD_REQUIRE(ntypes == 1 || tidx, "tidx is required when ntypes > 1");
    TMD_REQUIRE(!(flags & TMD_WANT_F) || force, "force is NULL but TMD_WANT_F is set");
    TMD_REQUIRE(vw != nullptr, "vw is NULL");
    TMD_REQUIRE(nl->xs_ext == nullptr, "this handle is bound to a slab (tmd_nlist_bind_records)");
    TMD_CHECK(require_device());
    cudaStream_t s = as_stream(stream);
    const int dim = box->dim;
    if (count == 0) return launch_reduce_vw(nullptr, 0, 0.5, flags, vw, s);

    NlArgs a;
    bool same = false;
    TMD_CHECK(nl_prepare(nl, n, box, table, ntypes, first, count, 0, s, a, same, ntypes > 1 ? tidx : nullptr));
    a.pos = pos;
    a.tidx = tidx;

    const int nb = (int)((n + 255) / 256);
    if (same) {  // otherwise the flag is already set and order / image / xb are not valid yet
        if (dim == 1) ::tmd::count_launch(), nl_gather_check_kernel<1><<<nb, 256, 0, s>>>(pos, nl->image, a.order, nl->xb, n, a.half_skin2, nl->xs, nl->flags);
        else if (dim == 2) ::tmd::count_launch(), nl_gather_check_kernel<2><<<nb, 256, 0, s>>>(pos, nl->image, a.order, nl->xb, n, a.half_skin2, nl->xs, nl->flags);
        else ::tmd::count_launch(), nl_gather_check_kernel<3><<<nb, 256, 0, s>>>(pos, nl->image, a.order, nl->xb, n, a.half_skin2, nl->xs, nl->flags);
    }
    // ---- gated build ----
    if (dim == 1) TMD_CHECK(nl_launch_build<1>(nl, a, s));
    else if (dim == 2) TMD_CHECK(nl_launch_build<2>(nl, a, s));
    else TMD_CHECK(nl_launch_build<3>(nl, a, s));
    // ---- traversal ----
    return nl_traverse(a, dim, flags, force, vw, s);
}

// ---- resident VelocityVerlet run -------------------------------------------------------------------
// K steps of VelocityVerlet (integrators.py:84-94) over a LennardJonesCut force field without leaving the
// device and without a host decision: the state is moved into cell order once, every step is [gated
// rebuild] + ONE traversal whose epilogue integrates (nl_force_kernel<EPI = 1>), and the state is moved
// back at the end.  The lists, their build positions and the rebuild rule are those of tmd_lj_nlist on the
// same handle, so the trajectory is bit-identical to K x (tmd_vv_kick_drift, tmd_lj_nlist, tmd_vv_kick).
}  // extern "C"

// slot-ordered state of a resident run (second record buffer; xu, vs, fs, fx, image_s: 3 doubles per slot, imass_s: 1)
static int md_reserve_state(tmd_nlist *nl, cudaStream_t s) {
    if (nl->cap_md >= nl->cap_n) return TMD_OK;
    TMD_CUDA(cudaStreamSynchronize(s));
    cudaFree(nl->xs2);
    cudaFree(nl->md);
    nl->xs2 = nullptr;
    nl->md = nullptr;
    nl->cap_md = 0;
    TMD_CUDA(cudaMalloc(&nl->xs2, (size_t)nl->cap_n * sizeof(double4)));
    TMD_CUDA(cudaMalloc(&nl->md, (size_t)nl->cap_n * 16 * sizeof(double)));
    nl->cap_md = nl->cap_n;
    return TMD_OK;
}

template <int DIM>
static int md_run_vv(tmd_nlist *nl, NlArgs &a, bool same, double *pos, double *vel, double *force,
                     const double *imass, int64_t n, int steps, cudaStream_t s, const MdDwPair *bond = nullptr) {
    double4 *rec[2] = {nl->xs, nl->xs2};
    int *gates = nl->flags + 8;
    const int nb = (int)((n + 255) / 256);
    const int fblocks = nl_force_blocks(n, DIM);
    TMD_CHECK(arena_get(SLOT_PARTIAL_VW, (size_t)fblocks * 10 * sizeof(double), (void **)&a.partial_vw));
    double *xu = a.xu, *vs = a.vs, *fs = a.fs;
    double *imass_s = const_cast<double *>(a.imass_s), *image_s = const_cast<double *>(a.image_s);

    if (!same) {  // no lists for this system yet (nl_prepare raised nlflags[0]): build them from pos as it is
        a.gate = nl->flags;
        a.xs = rec[0];
        TMD_CHECK(nl_launch_build<DIM>(nl, a, s));
        ::tmd::count_launch(), md_built_kernel<<<1, 1, 0, s>>>(nl->flags);
    }
    TMD_CUDA(cudaMemsetAsync(gates, 0, 2 * sizeof(int), s));
    ::tmd::count_launch(), md_import_kernel<DIM><<<nb, 256, 0, s>>>(a.order, n, pos, vel, force, imass, a.image, xu, vs, fs,
                                                                   imass_s, image_s, nullptr);
    if (bond) {  // forces of the second potential: zero except at the slots of the active atoms
        TMD_CUDA(cudaMemsetAsync(const_cast<double *>(a.fx), 0, (size_t)n * DIM * sizeof(double), s));
        TMD_CUDA(cudaMemsetAsync(bond->prev_slot, 0xFF, 128 * sizeof(int), s));  // -1: nothing written yet
    }
    int p = 0;
    ::tmd::count_launch(), md_pre_kernel<DIM><<<nb, 256, 0, s>>>(n, xu, vs, fs, imass_s, image_s, a.xb, rec[p], a.dt,
                                                                a.half_dt, a.half_skin2, gates + 0);
    std::vector<cudaEvent_t> events;
    if (nl->timing) {
        events.resize(2 * (size_t)steps);
        for (auto &e : events) TMD_CUDA(cudaEventCreate(&e));
    }
    for (int step = 0; step < steps; ++step) {
        const int g = step & 1;
        const bool last = step == steps - 1;
        // ---- rebuild, if the positions written by the previous pass say so (decided on the device) ----
        a.gate = gates + g;
        a.xs = rec[p];
        a.md_rebuild = 1;  // the two build kernels also carry the state from the old slots to the new ones
        a.pos_w = pos;
        a.vel_w = vel;
        a.imass = imass;
        TMD_CHECK(nl_launch_build<DIM>(nl, a, s));
        a.md_rebuild = 0;
        if (bond)  // the DoubleWellPair forces at the positions this traversal evaluates (after the rebuild: current slots)
            ::tmd::count_launch(), md_dwpair_kernel<DIM><<<1, 128, 0, s>>>(*bond, a.slot_of, xu, const_cast<double *>(a.fx),
                                                                          TMD_WANT_F | (last ? (TMD_WANT_V | TMD_WANT_W) : 0u));
        // ---- forces + closing kick (+ opening kick and drift of the next step) ----
        a.xs_next = rec[p ^ 1];
        a.flag_clear = gates + g;
        a.flag_next = gates + (g ^ 1);
        a.epi_last = last ? 1 : 0;
        if (nl->timing) TMD_CUDA(cudaEventRecord(events[2 * step], s));
        nl_launch_force_epi<DIM>(a, fblocks, s);
        if (nl->timing) TMD_CUDA(cudaEventRecord(events[2 * step + 1], s));
        if (!last) p ^= 1;
    }
    if (bond)  // System sums the potentials in order (system.py:58-76): Lennard-Jones wrote (V, W), the bond adds its share
        TMD_CHECK(launch_reduce_vw(bond->partials, 1, 0.5, TMD_WANT_V | TMD_WANT_W | TMD_ACCUMULATE, a.vw, s));
    ::tmd::count_launch(), md_export_kernel<DIM><<<nb, 256, 0, s>>>(a.order, n, xu, vs, fs, pos, vel, force, nullptr);
    TMD_CUDA(cudaGetLastError());
    if (nl->timing) {  // measurement runs only: waits for the run
        TMD_CUDA(cudaStreamSynchronize(s));
        for (int step = 0; step < steps; ++step) {
            float ms = 0.f;
            TMD_CUDA(cudaEventElapsedTime(&ms, events[2 * step], events[2 * step + 1]));
            nl->traversal_ms += ms;
            nl->traversals += 1;
        }
        for (auto &e : events) cudaEventDestroy(e);
    }
    return TMD_OK;
}

extern "C" {

static int run_vv_impl(tmd_nlist *nl, double *pos, double *vel, double *force, const double *imass,
                       const int32_t *tidx, int64_t n, const tmd_box *box, const double *table, int32_t ntypes,
                       double dt, int32_t steps, const tmd_dwpair_term *term, double *vw, tmd_stream_t stream) {
    TMD_REQUIRE(nl != nullptr, "nl is NULL");
    TMD_REQUIRE(box && box->dim >= 1 && box->dim <= 3, "box->dim must be 1..3");
    TMD_REQUIRE(n >= 2 && steps >= 1, "needs n >= 2 and steps >= 1");
    TMD_REQUIRE(pos && vel && force && imass && vw, "NULL argument");
    TMD_REQUIRE(ntypes == 1 || tidx, "tidx is required when ntypes > 1");
    TMD_REQUIRE(nl->xs_ext == nullptr, "this handle is bound to a slab (tmd_nlist_bind_records)");
    TMD_CHECK(require_device());
    cudaStream_t s = as_stream(stream);
    const int dim = box->dim;

    NlArgs a;
    bool same = false;
    TMD_CHECK(nl_prepare(nl, n, box, table, ntypes, 0, n, 0, s, a, same, ntypes > 1 ? tidx : nullptr));
    a.pos = pos;
    a.tidx = tidx;
    TMD_CHECK(md_reserve_state(nl, s));
    const size_t stride = (size_t)nl->cap_md * 3;
    a.xu = nl->md;
    a.vs = nl->md + stride;
    a.fs = nl->md + 2 * stride;
    a.fx = nullptr;
    MdDwPair bond, *bp = nullptr;
    if (term) {  // a DoubleWellPair on top of the Lennard-Jones potential: its forces arrive through fx
        const int nown = term->same_type ? term->n0 : term->n0 + term->n1;
        TMD_REQUIRE(term->n0 >= 0 && term->n1 >= 0 && nown <= 128, "the resident run takes at most 128 atoms of the bonded types");
        TMD_REQUIRE(nown == 0 || term->idx0 || term->n0 == 0, "idx0 is NULL");
        if (!nl->dw_scratch) TMD_CUDA(cudaMalloc(&nl->dw_scratch, 128 * sizeof(int) + 16 * sizeof(double)));
        bond.idx0 = term->idx0;
        bond.idx1 = term->idx1;
        bond.n0 = term->n0;
        bond.n1 = term->n1;
        bond.same_type = term->same_type;
        bond.rwidth = term->rwidth;
        bond.width2 = term->width2;
        bond.height = term->height;
        for (int k = 0; k < 3; ++k) {
            const bool on = k < dim && box->periodic[k];
            bond.per[k] = on ? 1 : 0;
            bond.len[k] = on ? box->length[k] : 0.0;
            bond.ilen[k] = on ? box->ilength[k] : 0.0;
            bond.half[k] = on ? 0.5 * box->length[k] : 0.0;
        }
        bond.prev_slot = reinterpret_cast<int *>(nl->dw_scratch);
        bond.partials = reinterpret_cast<double *>(reinterpret_cast<char *>(nl->dw_scratch) + 128 * sizeof(int));
        bp = &bond;
        a.fx = nl->md + 3 * stride;
    }
    a.image_s = nl->md + 4 * stride;
    a.imass_s = nl->md + 5 * stride;
    a.dt = dt;
    a.half_dt = dt * 0.5;  // VelocityVerlet.half_timestep (integrators.py:82)
    a.flags = TMD_WANT_V | TMD_WANT_F | TMD_WANT_W;
    a.force = nullptr;
    a.vw = vw;
    if (dim == 1) return md_run_vv<1>(nl, a, same, pos, vel, force, imass, n, steps, s, bp);
    if (dim == 2) return md_run_vv<2>(nl, a, same, pos, vel, force, imass, n, steps, s, bp);
    return md_run_vv<3>(nl, a, same, pos, vel, force, imass, n, steps, s, bp);
}

int tmd_nlist_run_vv(tmd_nlist *nl, double *pos, double *vel, double *force, const double *imass,
                     const int32_t *tidx, int64_t n, const tmd_box *box, const double *table, int32_t ntypes,
                     double dt, int32_t steps, double *vw, tmd_stream_t stream) {
    return run_vv_impl(nl, pos, vel, force, imass, tidx, n, box, table, ntypes, dt, steps, nullptr, vw, stream);
}

int tmd_nlist_run_vv_dwpair(tmd_nlist *nl, double *pos, double *vel, double *force, const double *imass,
                            const int32_t *tidx, int64_t n, const tmd_box *box, const double *table, int32_t ntypes,
                            double dt, int32_t steps, const tmd_dwpair_term *term, double *vw, tmd_stream_t stream) {
    TMD_REQUIRE(term != nullptr, "term is NULL");
    return run_vv_impl(nl, pos, vel, force, imass, tidx, n, box, table, ntypes, dt, steps, term, vw, stream);
}

}  // extern "C"

// K steps of LangevinInertia (integrators.py:265-282) over a LennardJonesCut force field, resident like md_run_vv: slot-ordered
// state, per step [gated rebuild] + a traversal that leaves the forces by slot (EPI = 3: no integration inside) + ONE pass over
// the state (closing update of the step + opening update of the next).  The list, its build positions and the rebuild rule are
// those of tmd_lj_nlist; the noise is addressed by atom index: bit-identical to the step-by-step path.
template <int DIM>
static int md_run_langevin(tmd_nlist *nl, NlArgs &a, bool same, double *pos, double *vel, double *force, const double *imass,
                           int64_t n, int steps, const tmd_langevin_consts &c, uint64_t key, uint64_t offset, cudaStream_t s) {
    double4 *rec[2] = {nl->xs, nl->xs2};
    int *gates = nl->flags + 8;
    const int nb = (int)((n + 255) / 256);
    const int fblocks = nl_force_blocks(n, DIM);
    const uint64_t stride = 2ull * (uint64_t)n * DIM;  // normals one step consumes
    TMD_CHECK(arena_get(SLOT_PARTIAL_VW, (size_t)fblocks * 10 * sizeof(double), (void **)&a.partial_vw));
    double *xu = a.xu, *vs = a.vs, *fs = a.fs;
    double *imass_s = const_cast<double *>(a.imass_s), *image_s = const_cast<double *>(a.image_s);
    if (!same) {
        a.gate = nl->flags;
        a.xs = rec[0];
        TMD_CHECK(nl_launch_build<DIM>(nl, a, s));
        ::tmd::count_launch(), md_built_kernel<<<1, 1, 0, s>>>(nl->flags);
    }
    TMD_CUDA(cudaMemsetAsync(gates, 0, 2 * sizeof(int), s));
    ::tmd::count_launch(), md_import_kernel<DIM><<<nb, 256, 0, s>>>(a.order, n, pos, vel, force, imass, a.image, xu, vs, fs,
                                                                   imass_s, image_s, nullptr);
    int p = 0;
    ::tmd::count_launch(), md_langevin_kernel<DIM><<<nb, 256, 0, s>>>(n, a.order, xu, vs, fs, imass_s, image_s, a.xb, rec[p], c, key,
                                                                     offset, 2, a.half_skin2, gates + 0);
    for (int step = 0; step < steps; ++step) {
        const int g = step & 1;
        const bool last = step == steps - 1;
        a.gate = gates + g;
        a.xs = rec[p];
        a.md_rebuild = 1;  // the two build kernels also carry the state from the old slots to the new ones
        a.pos_w = pos;
        a.vel_w = vel;
        a.imass = imass;
        TMD_CHECK(nl_launch_build<DIM>(nl, a, s));
        a.md_rebuild = 0;
        a.flag_clear = gates + g;
        a.epi_last = last ? 1 : 0;
        a.force = fs;
        ::tmd::count_launch();
        if (a.ntypes == 1 || a.uniform_table) nl_force_kernel<DIM, true, 4, true, true, 5, 3><<<fblocks, kNlThreads, 0, s>>>(a);
        else nl_force_kernel<DIM, false, 4, true, true, 5, 3><<<fblocks, kNlThreads, 0, s>>>(a);
        ::tmd::count_launch(), md_langevin_kernel<DIM><<<nb, 256, 0, s>>>(n, a.order, xu, vs, fs, imass_s, image_s, a.xb, rec[p ^ 1], c,
                                                                         key, offset + (uint64_t)(step + 1) * stride,
                                                                         last ? 1 : 3, a.half_skin2, gates + (g ^ 1));
        if (!last) p ^= 1;
    }
    ::tmd::count_launch(), md_export_kernel<DIM><<<nb, 256, 0, s>>>(a.order, n, xu, vs, fs, pos, vel, force, nullptr);
    TMD_CUDA(cudaGetLastError());
    return TMD_OK;
}

extern "C" {

int tmd_nlist_run_langevin(tmd_nlist *nl, double *pos, double *vel, double *force, const double *imass,
                           const int32_t *tidx, int64_t n, const tmd_box *box, const double *table, int32_t ntypes,
                           const tmd_langevin_consts *consts, const tmd_rng *rng, int32_t steps, double *vw,
                           tmd_stream_t stream) {
    TMD_REQUIRE(nl != nullptr, "nl is NULL");
    TMD_REQUIRE(box && box->dim >= 1 && box->dim <= 3, "box->dim must be 1..3");
    TMD_REQUIRE(n >= 2 && steps >= 1, "needs n >= 2 and steps >= 1");
    TMD_REQUIRE(pos && vel && force && imass && vw && consts && rng, "NULL argument");
    TMD_REQUIRE(rng->step_counter == nullptr, "the resident run takes the stream position in rng->offset");
    TMD_REQUIRE(ntypes == 1 || tidx, "tidx is required when ntypes > 1");
    TMD_REQUIRE(nl->xs_ext == nullptr, "this handle is bound to a slab (tmd_nlist_bind_records)");
    TMD_CHECK(require_device());
    cudaStream_t s = as_stream(stream);
    const int dim = box->dim;
    NlArgs a;
    bool same = false;
    TMD_CHECK(nl_prepare(nl, n, box, table, ntypes, 0, n, 0, s, a, same, ntypes > 1 ? tidx : nullptr));
    a.pos = pos;
    a.tidx = tidx;
    TMD_CHECK(md_reserve_state(nl, s));
    const size_t stride = (size_t)nl->cap_md * 3;
    a.xu = nl->md;
    a.vs = nl->md + stride;
    a.fs = nl->md + 2 * stride;
    a.fx = nullptr;
    a.image_s = nl->md + 4 * stride;
    a.imass_s = nl->md + 5 * stride;
    a.flags = TMD_WANT_V | TMD_WANT_F | TMD_WANT_W;  // force() and potential() at the same positions: one traversal
    a.vw = vw;
    a.row_mode = 1;  // forces by slot
    a.row_cell_lo = 0;
    a.row_cell_hi = a.grid.ncells;  // EPI = 3 takes its rows from the cell table: all slots
    a.work = nl->flags + 10;
    a.dynamic_rows = 0;
    if (dim == 1) return md_run_langevin<1>(nl, a, same, pos, vel, force, imass, n, steps, *consts, rng->key, rng->offset, s);
    if (dim == 2) return md_run_langevin<2>(nl, a, same, pos, vel, force, imass, n, steps, *consts, rng->key, rng->offset, s);
    return md_run_langevin<3>(nl, a, same, pos, vel, force, imass, n, steps, *consts, rng->key, rng->offset, s);
}

// ---- slab path: rows are cell-order slots, state lives in slot order ------------------------------

int tmd_nlist_bind_records(tmd_nlist *nl, double *records) {
    TMD_REQUIRE(nl != nullptr, "nl is NULL");
    TMD_REQUIRE(nl->cap_n == 0, "bind the record buffer before the first build");
    TMD_REQUIRE((reinterpret_cast<uintptr_t>(records) & 31u) == 0, "records must be 32-byte aligned");
    nl->xs_ext = reinterpret_cast<double4 *>(records);
    return TMD_OK;
}

int tmd_nlist_build_slots(tmd_nlist *nl, const double *pos, const int32_t *tidx, int64_t n, const tmd_box *box,
                          const double *table, int32_t ntypes, int64_t slot_first, int64_t slot_count,
                          int32_t gated, tmd_stream_t stream) {
    TMD_REQUIRE(nl != nullptr, "nl is NULL");
    TMD_REQUIRE(box && box->dim >= 1 && box->dim <= 3, "box->dim must be 1..3");
    TMD_REQUIRE(n >= 2 && slot_first >= 0 && slot_count > 0 && slot_first + slot_count <= n,
                "slots [slot_first, slot_first+slot_count) must lie in [0, n)");
    TMD_REQUIRE(pos != nullptr, "pos is NULL");
    TMD_REQUIRE(ntypes == 1 || tidx, "tidx is required when ntypes > 1");
    TMD_CHECK(require_device());
    cudaStream_t s = as_stream(stream);
    const int dim = box->dim;
    NlArgs a;
    bool same = false;
    TMD_CHECK(nl_prepare(nl, n, box, table, ntypes, slot_first, slot_count, 1, s, a, same, ntypes > 1 ? tidx : nullptr));
    a.pos = pos;
    a.tidx = tidx;
    if (!gated) TMD_CUDA(cudaMemsetAsync(nl->flags, 1, sizeof(int), s));  // any non-zero value raises the flag
    if (dim == 1) TMD_CHECK(nl_launch_build<1>(nl, a, s));
    else if (dim == 2) TMD_CHECK(nl_launch_build<2>(nl, a, s));
    else TMD_CHECK(nl_launch_build<3>(nl, a, s));
    ::tmd::count_launch(), nl_reach_kernel<<<1, 32, 0, s>>>(a, dim);
    TMD_CUDA(cudaGetLastError());
    if (!nl->saved) {
        nl->saved = new (std::nothrow) NlArgs();
        if (!nl->saved) return TMD_ERR_NOMEM;
    }
    *nl->saved = a;
    return TMD_OK;
}

int tmd_nlist_force_slots(tmd_nlist *nl, uint32_t flags, double *force_s, double *vw, tmd_stream_t stream) {
    TMD_REQUIRE(nl != nullptr && nl->saved != nullptr, "no slot build on this handle");
    TMD_REQUIRE(!(flags & TMD_WANT_F) || force_s, "force_s is NULL but TMD_WANT_F is set");
    TMD_REQUIRE(vw != nullptr, "vw is NULL");
    NlArgs a = *nl->saved;
    return nl_traverse(a, nl->dim, flags, force_s, vw, as_stream(stream));
}

int tmd_nlist_view(tmd_nlist *nl, tmd_nlist_buffers *out) {
    TMD_REQUIRE(nl != nullptr && nl->saved != nullptr && out != nullptr, "no slot build on this handle");
    const NlArgs &a = *nl->saved;
    out->records = reinterpret_cast<double *>(a.xs);
    out->build_records = reinterpret_cast<double *>(a.xb);
    out->order = a.order;
    out->slot_of = a.slot_of;
    out->image = a.image;
    out->flags = a.nlflags;
    for (int k = 0; k < 3; ++k) out->cells[k] = a.grid.nc[k];
    out->half_width = a.grid.h;
    out->capacity = a.cap;
    return TMD_OK;
}

}  // extern "C"
#include "slabrun.cuh"
