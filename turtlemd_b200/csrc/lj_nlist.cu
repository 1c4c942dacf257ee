// K2 + K3 with a persistent Verlet neighbour list: the large-N Lennard-Jones path.
//
// Same contract as tmd_lj_allpairs / tmd_lj_cells (lennardjones.py:145-273 + box.py:384-403), but
// the per-atom candidate lists are kept between calls in a tmd_nlist handle and reused while no
// atom has moved more than skin/2 since they were built (then every pair closer than rcut is still
// listed, so forces are those of the full evaluation -- this is an acceleration structure, not an
// approximation).  Everything is stream-ordered, there is no host round trip:
//
//   gather    xs[slot] = pos[atom] - image[atom]  (32-byte records in cell order; the lattice
//   + check   vector `image` is frozen at build time, so coordinates stay continuous between
//             builds) and, in the same pass, |xs[slot] - build position|^2 > (skin/2)^2 -> flag
//   build     every kernel returns at once unless the flag is set:
//             bin a wrapped copy into cells of edge >= rcut+skin (cell_build.cuh), then one thread
//             per row atom walks its 3^dim cells with a conservative SINGLE-precision distance
//             filter and appends (neighbour slot | image code << 27) to its list; lists are stored
//             transposed (entry k of all rows contiguous) so the traversal reads them coalesced
//   traverse  one thread per row atom: ~75 candidates, 6 FP64 instructions each to the exact
//             cut-off test (strict rsq < rcut2, lennardjones.py:291-294), ~70 % pass.  Per-thread
//             force accumulators, no atomics, fixed order -> bit-reproducible.
//
// Virial.  The reference sums f (x) delta over pairs (lennardjones.py:272).  With delta_ij =
// x_i - x_j - S_ij (S = the lattice vector of the image) and f_ij = -f_ji this equals
//     W = sum_i F_i (x) x_i  -  1/2 sum_i sum_j f_ij (x) S_ij ,
// i.e. six FMAs per ATOM plus a correction only for pairs that cross a box face, instead of six
// FMAs per pair.  x_i is taken relative to the box centre.  Agreement with the pair sum: ~1e-14.
//
// A row whose candidate count exceeds the list capacity is not truncated: its count is stored
// as-is and the traversal walks the build-time cells for that row instead (slow, exact).
//
// FP64-pipe bound; HBM/L2 traffic per atom-step: list indices (~4 B x 75) + records (32 B) + force.
#include <string.h>

#include <new>

#include "cell_build.cuh"

struct tmd_nlist {
    double skin;
    int device;
    // signature of what the lists were built for (a change forces a rebuild)
    int64_t n, first, count;
    int dim, ntypes;
    double len[3];
    double reach;
    // device memory, owned
    int *flags;          // [0] rebuild, [1] rows on the slow path, [2] builds, [3] calls
    unsigned *entries;   // [cap][stride]
    int *nneigh;         // [stride]
    int *ints;           // counts, starts, cell_of, rank_of, order, slot_of
    double4 *wrapped, *sorted, *xs;
    float4 *sortedf;
    double *image;       // [n*3]
    int64_t cap_n, cap_rows, cap_cells;
    int cap, stride;
};

namespace tmd {

constexpr int kNlThreads = 128;
constexpr unsigned kSlotMask = 0x07FFFFFFu;
constexpr int kHomeCode = 13;  // (0,0,0) shift

struct NlArgs {
    CellArgs c;
    const int *gate;
    int *flags;
    unsigned *entries;
    int *nneigh;
    int cap, stride;
    float reach2f;       // conservative single-precision filter radius^2
    double4 *xs;
    double half_skin2;
    double centre[3];
};

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

template <int DIM>
__global__ void __launch_bounds__(256) nl_gather_check_kernel(const double *__restrict__ pos,
                                                              const double *__restrict__ image,
                                                              const int *__restrict__ order,
                                                              const double4 *__restrict__ sorted, int64_t n,
                                                              double half_skin2, double4 *__restrict__ xs,
                                                              int *flags) {
    const int64_t slot = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= n) return;
    const int atom = order[slot];
    const double4 b = sorted[slot];  // position at build time, .w = type
    double x[3] = {0.0, 0.0, 0.0};
#pragma unroll
    for (int k = 0; k < DIM; ++k) x[k] = pos[(int64_t)atom * DIM + k] - image[(int64_t)atom * 3 + k];
    double r = (x[0] - b.x) * (x[0] - b.x);
    if (DIM > 1) r = fma(x[1] - b.y, x[1] - b.y, r);
    if (DIM > 2) r = fma(x[2] - b.z, x[2] - b.z, r);
    xs[slot] = make_double4(x[0], x[1], x[2], b.w);
    if (!(r <= half_skin2)) flags[0] = 1;  // also catches NaN
}

// after a (gated) build the records ARE the build positions
static __global__ void __launch_bounds__(256) nl_copy_sorted_kernel(const double4 *__restrict__ sorted,
                                                                    double4 *__restrict__ xs, int64_t n,
                                                                    const int *gate) {
    TMD_GATE(gate)
    const int64_t slot = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (slot < n) xs[slot] = sorted[slot];
}

struct Stencil {
    int ncell;
    unsigned code;
    float sx, sy, sz;   // shift folded into the row atom (single precision filter)
};

template <int DIM>
__device__ __forceinline__ void neighbour_cell(const CellGrid &g, int c0, int c1, int c2, int dx, int dy, int dz,
                                               int &ncell, int (&sh)[3]) {
    int nb[3] = {c0 + dx, c1 + dy, c2 + dz};
    sh[0] = sh[1] = sh[2] = 0;
#pragma unroll
    for (int k = 0; k < DIM; ++k) {
        if (nb[k] < 0) { nb[k] += g.nc[k]; sh[k] = -1; }
        else if (nb[k] >= g.nc[k]) { nb[k] -= g.nc[k]; sh[k] = 1; }
    }
    ncell = (nb[2] * g.nc[1] + nb[1]) * g.nc[0] + nb[0];
}

template <int DIM>
__global__ void __launch_bounds__(kNlThreads) nl_build_kernel(const NlArgs a) {
    TMD_GATE(a.gate)
    const CellArgs &c = a.c;
    const int64_t row = (int64_t)blockIdx.x * kNlThreads + threadIdx.x;
    if (row >= c.count) return;
    const bool whole = c.first == 0 && c.count == c.n;
    const int slot = whole ? (int)row : c.slot_of[c.first + row];
    const int atom = whole ? c.order[slot] : (int)(c.first + row);
    const float4 me = c.sortedf[slot];
    const int cell = c.cell_of[atom];
    const int c0 = cell % c.grid.nc[0], c1 = (cell / c.grid.nc[0]) % c.grid.nc[1],
              c2 = cell / (c.grid.nc[0] * c.grid.nc[1]);
    const float lx = (float)c.grid.len[0], ly = (float)c.grid.len[1], lz = (float)c.grid.len[2];
    const float reach2 = a.reach2f;
    int cnt = 0;
    for (int dz = (DIM > 2 ? -1 : 0); dz <= (DIM > 2 ? 1 : 0); ++dz)
        for (int dy = (DIM > 1 ? -1 : 0); dy <= (DIM > 1 ? 1 : 0); ++dy)
            for (int dx = -1; dx <= 1; ++dx) {
                int ncell, sh[3];
                neighbour_cell<DIM>(c.grid, c0, c1, c2, dx, dy, dz, ncell, sh);
                const float xs = me.x - sh[0] * lx;
                const float ys = me.y - sh[1] * ly;
                const float zs = me.z - sh[2] * lz;
                const unsigned code = (unsigned)((sh[0] + 1) + 3 * (sh[1] + 1) + 9 * (sh[2] + 1)) << 27;
                const int jend = c.starts[ncell + 1];
                for (int j = c.starts[ncell]; j < jend; ++j) {
                    const float4 q = c.sortedf[j];
                    float d = xs - q.x;
                    float r = d * d;
                    if (DIM > 1) { d = ys - q.y; r = fmaf(d, d, r); }
                    if (DIM > 2) { d = zs - q.z; r = fmaf(d, d, r); }
                    if (r < reach2 && j != slot) {
                        if (cnt < a.cap) a.entries[(size_t)cnt * a.stride + row] = code | (unsigned)j;
                        ++cnt;
                    }
                }
            }
    if (cnt > a.cap) atomicAdd(&a.flags[1], 1);  // statistics only: the traversal handles it exactly
    a.nneigh[row] = cnt;
}

struct NlAcc {
    double f[3];
    double v;
    double wc[9];   // crossing correction: sum f (x) S, upper triangle used
    int hits;
};

// One candidate pair; d = x_i - x_j - S already formed.  SINGLE: one particle type (coefficients in
// registers, the shift `offset` is applied once per thread from the hit count).
template <int DIM, bool SINGLE>
__device__ __forceinline__ void nl_pair(const double (&d)[3], double rsq, double lj1, double lj2, double lj3,
                                        double lj4, double off, uint32_t flags, NlAcc &acc) {
    const double r2inv = rcp_fast(rsq);                                  // lennardjones.py:254
    const double r6inv = (r2inv * r2inv) * r2inv;                        // :255
    if (flags & TMD_WANT_V) {                                            // :276-282
        acc.v = fma(r6inv, fma(lj3, r6inv, -lj4), acc.v);
        if (SINGLE) acc.hits += 1;
        else acc.v -= off;
    }
    if (flags & (TMD_WANT_F | TMD_WANT_W)) {
        const double flj = (r2inv * r6inv) * fma(lj1, r6inv, -lj2);      // :285-288
#pragma unroll
        for (int k = 0; k < DIM; ++k) acc.f[k] = fma(flj, d[k], acc.f[k]);  // :269-270
    }
}

// the rare pair whose neighbour is a periodic image: also feeds the virial correction
template <int DIM, bool SINGLE>
__device__ __forceinline__ void nl_pair_crossing(const double (&d)[3], const double (&S)[3], double rsq, double lj1,
                                                 double lj2, double lj3, double lj4, double off, uint32_t flags,
                                                 NlAcc &acc) {
    const double r2inv = rcp_fast(rsq);
    const double r6inv = (r2inv * r2inv) * r2inv;
    if (flags & TMD_WANT_V) {
        acc.v = fma(r6inv, fma(lj3, r6inv, -lj4), acc.v);
        if (SINGLE) acc.hits += 1;
        else acc.v -= off;
    }
    if (flags & (TMD_WANT_F | TMD_WANT_W)) {
        const double flj = (r2inv * r6inv) * fma(lj1, r6inv, -lj2);
#pragma unroll
        for (int k = 0; k < DIM; ++k) {
            const double fk = flj * d[k];
            acc.f[k] += fk;
#pragma unroll
            for (int b = k; b < DIM; ++b) acc.wc[k * 3 + b] = fma(fk, S[b], acc.wc[k * 3 + b]);
        }
    }
}

template <int DIM, bool SINGLE>
__global__ void __launch_bounds__(kNlThreads, 4) nl_force_kernel(const NlArgs a) {
    __shared__ double s_tab[SINGLE ? 1 : 6 * kMaxT2];
    __shared__ double s_red[10 * (kNlThreads / 32)];
    const CellArgs &c = a.c;
    const int tid = threadIdx.x;
    const int64_t row = (int64_t)blockIdx.x * kNlThreads + tid;
    const bool valid = row < c.count;
    const bool whole = c.first == 0 && c.count == c.n;
    const int ntypes = c.ntypes;
    if (!SINGLE) {
        for (int k = tid; k < 6 * kMaxT2; k += kNlThreads) s_tab[k] = c.tab.c[k / kMaxT2][k % kMaxT2];
        __syncthreads();
    }
    if (blockIdx.x == 0 && tid == 0) {  // every gated kernel of this call is behind us in the stream
        if (a.flags[0]) {
            a.flags[0] = 0;
            a.flags[2] += 1;
        }
        a.flags[3] += 1;
    }
    NlAcc acc;
#pragma unroll
    for (int k = 0; k < 3; ++k) acc.f[k] = 0.0;
    acc.v = 0.0;
#pragma unroll
    for (int k = 0; k < 9; ++k) acc.wc[k] = 0.0;
    acc.hits = 0;

    int atom = 0;
    double xme[3] = {0.0, 0.0, 0.0};
    const uint32_t flags = c.flags;
    const double lx = c.grid.len[0], ly = c.grid.len[1], lz = c.grid.len[2];
    const double lj1 = c.tab.c[0][0], lj2 = c.tab.c[1][0], lj3 = c.tab.c[2][0], lj4 = c.tab.c[3][0],
                 off = c.tab.c[4][0], rc2 = c.tab.c[5][0];
    if (valid) {
        const int slot = whole ? (int)row : c.slot_of[c.first + row];
        atom = whole ? c.order[slot] : (int)(c.first + row);
        const double4 me = a.xs[slot];
        xme[0] = me.x;
        xme[1] = me.y;
        xme[2] = me.z;
        const int ti = SINGLE ? 0 : (int)__double_as_longlong(me.w);
        const int nn = a.nneigh[row];

        auto candidate = [&](int j, int code) {
            const double4 q = a.xs[j];
            double d[3] = {0.0, 0.0, 0.0};
            d[0] = me.x - q.x;
            if (DIM > 1) d[1] = me.y - q.y;
            if (DIM > 2) d[2] = me.z - q.z;
            double l1 = lj1, l2 = lj2, l3 = lj3, l4 = lj4, of = off, r2c = rc2;
            if (!SINGLE) {
                const int p = ti * ntypes + (int)__double_as_longlong(q.w);
                l1 = s_tab[0 * kMaxT2 + p];
                l2 = s_tab[1 * kMaxT2 + p];
                l3 = s_tab[2 * kMaxT2 + p];
                l4 = s_tab[3 * kMaxT2 + p];
                of = s_tab[4 * kMaxT2 + p];
                r2c = s_tab[5 * kMaxT2 + p];
            }
            if (code == kHomeCode) {
                double rsq = d[0] * d[0];
                if (DIM > 1) rsq = fma(d[1], d[1], rsq);
                if (DIM > 2) rsq = fma(d[2], d[2], rsq);
                if (rsq < r2c) nl_pair<DIM, SINGLE>(d, rsq, l1, l2, l3, l4, of, flags, acc);
            } else {  // the neighbour is a periodic image: only for atoms near a box face
                double S[3] = {0.0, 0.0, 0.0};
                S[0] = (code % 3 - 1) * lx;
                if (DIM > 1) S[1] = ((code / 3) % 3 - 1) * ly;
                if (DIM > 2) S[2] = (code / 9 - 1) * lz;
#pragma unroll
                for (int k = 0; k < DIM; ++k) d[k] -= S[k];
                double rsq = d[0] * d[0];
                if (DIM > 1) rsq = fma(d[1], d[1], rsq);
                if (DIM > 2) rsq = fma(d[2], d[2], rsq);
                if (rsq < r2c) nl_pair_crossing<DIM, SINGLE>(d, S, rsq, l1, l2, l3, l4, of, flags, acc);
            }
        };

        if (nn <= a.cap) {
            const unsigned *list = a.entries + row;
            int k = 0;
            for (; k + 4 <= nn; k += 4) {  // four independent gathers in flight
                unsigned e[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) e[u] = __ldg(list + (size_t)(k + u) * a.stride);
#pragma unroll
                for (int u = 0; u < 4; ++u) candidate((int)(e[u] & kSlotMask), (int)(e[u] >> 27));
            }
            for (; k < nn; ++k) {
                const unsigned e = __ldg(list + (size_t)k * a.stride);
                candidate((int)(e & kSlotMask), (int)(e >> 27));
            }
        } else {
            // list capacity exceeded at build time: walk the build-time cells (every pair now inside
            // rcut was inside rcut + skin then, hence in this stencil)
            const int cell = c.cell_of[atom];
            const int c0 = cell % c.grid.nc[0], c1 = (cell / c.grid.nc[0]) % c.grid.nc[1],
                      c2 = cell / (c.grid.nc[0] * c.grid.nc[1]);
            for (int dz = (DIM > 2 ? -1 : 0); dz <= (DIM > 2 ? 1 : 0); ++dz)
                for (int dy = (DIM > 1 ? -1 : 0); dy <= (DIM > 1 ? 1 : 0); ++dy)
                    for (int dx = -1; dx <= 1; ++dx) {
                        int ncell, sh[3];
                        neighbour_cell<DIM>(c.grid, c0, c1, c2, dx, dy, dz, ncell, sh);
                        const int code = (sh[0] + 1) + 3 * (sh[1] + 1) + 9 * (sh[2] + 1);
                        const int jend = c.starts[ncell + 1];
                        for (int j = c.starts[ncell]; j < jend; ++j)
                            if (j != slot) candidate(j, code);
                    }
        }
    }

    if (valid && (flags & TMD_WANT_F)) {
        double *dst = c.force + (int64_t)atom * DIM;
#pragma unroll
        for (int k = 0; k < DIM; ++k) dst[k] = (flags & TMD_ACCUMULATE) ? dst[k] + acc.f[k] : acc.f[k];
    }
    // per-thread share of (V, W) in the "every pair seen from both atoms, halve at the end" convention
    double red[10];
    red[0] = SINGLE ? acc.v - acc.hits * off : acc.v;
#pragma unroll
    for (int k = 0; k < 9; ++k) red[1 + k] = 0.0;
    if (flags & TMD_WANT_W) {
#pragma unroll
        for (int p = 0; p < DIM; ++p)
#pragma unroll
            for (int q = p; q < DIM; ++q)
                red[1 + p * 3 + q] = fma(2.0 * acc.f[p], xme[q] - a.centre[q], -acc.wc[p * 3 + q]);
        red[1 + 3] = red[1 + 1];
        red[1 + 6] = red[1 + 2];
        red[1 + 7] = red[1 + 5];
    }
    block_sum<10, kNlThreads>(red, s_red);
    if (tid == 0) {
#pragma unroll
        for (int k = 0; k < 10; ++k) c.partial_vw[(size_t)blockIdx.x * 10 + k] = red[k];
    }
}

static void nl_release(tmd_nlist *nl) {
    cudaFree(nl->flags);
    cudaFree(nl->entries);
    cudaFree(nl->nneigh);
    cudaFree(nl->ints);
    cudaFree(nl->wrapped);
    cudaFree(nl->sorted);
    cudaFree(nl->sortedf);
    cudaFree(nl->xs);
    cudaFree(nl->image);
    nl->flags = nullptr;
    nl->entries = nullptr;
    nl->nneigh = nullptr;
    nl->ints = nullptr;
    nl->wrapped = nl->sorted = nl->xs = nullptr;
    nl->sortedf = nullptr;
    nl->image = nullptr;
    nl->cap_n = nl->cap_rows = nl->cap_cells = 0;
    nl->cap = nl->stride = 0;
}

static bool nl_grid(const tmd_box *box, double rc2max, double skin, int64_t n, CellGrid &g, double &reach) {
    reach = sqrt(rc2max) + skin;
    return n >= 2 && n < (1ll << 27) && make_grid(box, reach, g);
}

template <int DIM>
static void nl_launch_force(const NlArgs &a, int blocks, cudaStream_t s) {
    if (a.c.ntypes == 1) ::tmd::count_launch(), nl_force_kernel<DIM, true><<<blocks, kNlThreads, 0, s>>>(a);
    else ::tmd::count_launch(), nl_force_kernel<DIM, false><<<blocks, kNlThreads, 0, s>>>(a);
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
    double rc2max = 0.0, reach = 0.0;
    TMD_CHECK(load_lj_table(table, ntypes, tab, rc2max));
    CellGrid g;
    if (nl_grid(box, rc2max, skin, n, g, reach)) *usable = 1;
    return TMD_OK;
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
    TMD_CHECK(require_device());
    cudaStream_t s = as_stream(stream);
    const int dim = box->dim;

    NlArgs a;
    CellArgs &c = a.c;
    TMD_CHECK(load_lj_table(table, ntypes, c.tab, c.rc2max));
    double reach = 0.0;
    if (!nl_grid(box, c.rc2max, nl->skin, n, c.grid, reach)) {
        set_error("tmd_lj_nlist: needs a fully periodic box with >= 3 cells of edge rcut+skin per direction "
                  "and 2 <= n < 2^27; use tmd_lj_allpairs");
        return TMD_ERR_UNSUPPORTED;
    }
    if (count == 0) return launch_reduce_vw(nullptr, 0, 0.5, flags, vw, s);

    // ---- (re)allocate and decide whether the signature changed ----
    bool fresh = false;
    const int ncells = c.grid.ncells;
    if (n > nl->cap_n || count > nl->cap_rows || ncells > nl->cap_cells) {
        TMD_CUDA(cudaStreamSynchronize(s));
        nl_release(nl);
        // expected neighbours within reach at this density, +35 % and a floor, rounded to 8
        double volume = 1.0, lmax = 0.0;
        for (int k = 0; k < dim; ++k) {
            volume *= box->length[k];
            lmax = fmax(lmax, box->length[k]);
        }
        const double sphere = dim == 3 ? 4.18879020478639 * reach * reach * reach
                                       : (dim == 2 ? 3.14159265358979 * reach * reach : 2.0 * reach);
        int cap = (int)(1.35 * (double)n / volume * sphere) + 24;
        cap = (cap + 7) & ~7;
        if (cap > 1024) cap = 1024;
        nl->cap = cap;
        nl->stride = (int)((count + 31) & ~31ll);
        nl->cap_n = n;
        nl->cap_rows = count;
        nl->cap_cells = ncells;
        const size_t n_ints = (size_t)2 * (ncells + 1) + (size_t)4 * n;
        TMD_CUDA(cudaMalloc(&nl->flags, 4 * sizeof(int)));
        TMD_CUDA(cudaMemsetAsync(nl->flags, 0, 4 * sizeof(int), s));
        TMD_CUDA(cudaMalloc(&nl->entries, (size_t)nl->cap * nl->stride * sizeof(unsigned)));
        TMD_CUDA(cudaMalloc(&nl->nneigh, (size_t)nl->stride * sizeof(int)));
        TMD_CUDA(cudaMalloc(&nl->ints, n_ints * sizeof(int)));
        TMD_CUDA(cudaMemsetAsync(nl->ints, 0, n_ints * sizeof(int), s));  // counts[] must start at zero
        TMD_CUDA(cudaMalloc(&nl->wrapped, (size_t)n * sizeof(double4)));
        TMD_CUDA(cudaMalloc(&nl->sorted, (size_t)n * sizeof(double4)));
        TMD_CUDA(cudaMalloc(&nl->sortedf, (size_t)n * sizeof(float4)));
        TMD_CUDA(cudaMalloc(&nl->xs, (size_t)n * sizeof(double4)));
        TMD_CUDA(cudaMalloc(&nl->image, (size_t)n * 3 * sizeof(double)));
        fresh = true;
    }
    bool same = !fresh && nl->n == n && nl->first == first && nl->count == count && nl->dim == dim &&
                nl->ntypes == ntypes && nl->reach == reach;
    for (int k = 0; k < dim && same; ++k) same = nl->len[k] == box->length[k];
    if (!same) {
        nl->n = n;
        nl->first = first;
        nl->count = count;
        nl->dim = dim;
        nl->ntypes = ntypes;
        nl->reach = reach;
        nl->stride = (int)((count + 31) & ~31ll);
        for (int k = 0; k < 3; ++k) nl->len[k] = k < dim ? box->length[k] : 0.0;
        // force a build; counts[] may hold a stale grid of another size: clear the cell integers
        static const int one = 1;
        TMD_CUDA(cudaMemsetAsync(nl->ints, 0, (size_t)2 * (nl->cap_cells + 1) * sizeof(int), s));
        TMD_CUDA(cudaMemcpyAsync(nl->flags, &one, sizeof(int), cudaMemcpyHostToDevice, s));
    }

    c.pos = pos;
    c.tidx = tidx;
    c.n = n;
    c.first = first;
    c.count = count;
    c.ntypes = ntypes;
    c.flags = flags;
    c.force = force;
    c.counts = nl->ints;
    c.starts = nl->ints + (ncells + 1);
    c.cell_of = nl->ints + 2 * ((size_t)nl->cap_cells + 1);
    c.rank_of = c.cell_of + nl->cap_n;
    c.order = c.cell_of + 2 * nl->cap_n;
    c.slot_of = c.cell_of + 3 * nl->cap_n;
    c.wrapped = nl->wrapped;
    c.sorted = nl->sorted;
    c.sortedf = nl->sortedf;
    c.image = nl->image;
    c.gate = nl->flags;
    const int blocks = (int)((count + kNlThreads - 1) / kNlThreads);
    TMD_CHECK(arena_get(SLOT_PARTIAL_VW, (size_t)blocks * 10 * sizeof(double), (void **)&c.partial_vw));
    a.gate = nl->flags;
    a.flags = nl->flags;
    a.entries = nl->entries;
    a.nneigh = nl->nneigh;
    a.cap = nl->cap;
    a.stride = nl->stride;
    {
        // single-precision filter: coordinates in [0, L) (+- L after folding the shift) carry an
        // absolute error <= 5 * 2^-24 * L per component, i.e. < 2^-20 * Lmax on the distance
        double lmax = 0.0;
        for (int k = 0; k < dim; ++k) lmax = fmax(lmax, box->length[k]);
        const double rf = reach + lmax * (1.0 / 1048576.0);
        a.reach2f = (float)(rf * rf * (1.0 + 1e-6));
    }
    a.xs = nl->xs;
    a.half_skin2 = 0.25 * nl->skin * nl->skin;
    for (int k = 0; k < 3; ++k) a.centre[k] = k < dim ? 0.5 * box->length[k] : 0.0;

    const int nb = (int)((n + 255) / 256);
    if (same) {  // otherwise the flag is already set and order / image / sorted are not valid yet
        if (dim == 1) ::tmd::count_launch(), nl_gather_check_kernel<1><<<nb, 256, 0, s>>>(pos, nl->image, c.order, nl->sorted, n, a.half_skin2, nl->xs, nl->flags);
        else if (dim == 2) ::tmd::count_launch(), nl_gather_check_kernel<2><<<nb, 256, 0, s>>>(pos, nl->image, c.order, nl->sorted, n, a.half_skin2, nl->xs, nl->flags);
        else ::tmd::count_launch(), nl_gather_check_kernel<3><<<nb, 256, 0, s>>>(pos, nl->image, c.order, nl->sorted, n, a.half_skin2, nl->xs, nl->flags);
    }
    // ---- gated build ----
    TMD_CHECK(launch_cell_build(c, dim, s, /*clear_first=*/false));
    ::tmd::count_launch(), nl_copy_sorted_kernel<<<nb, 256, 0, s>>>(nl->sorted, nl->xs, n, nl->flags);
    if (dim == 1) ::tmd::count_launch(), nl_build_kernel<1><<<blocks, kNlThreads, 0, s>>>(a);
    else if (dim == 2) ::tmd::count_launch(), nl_build_kernel<2><<<blocks, kNlThreads, 0, s>>>(a);
    else ::tmd::count_launch(), nl_build_kernel<3><<<blocks, kNlThreads, 0, s>>>(a);
    // ---- every call ----
    if (dim == 1) nl_launch_force<1>(a, blocks, s);
    else if (dim == 2) nl_launch_force<2>(a, blocks, s);
    else nl_launch_force<3>(a, blocks, s);
    TMD_CUDA(cudaGetLastError());
    return launch_reduce_vw(c.partial_vw, blocks, 0.5, flags, vw, s);
}

}  // extern "C"
