// Micro-benchmark (lever 1c): can TMA `cp.async.bulk.tensor.2d ... tile::gather4` feed the
// neighbour-list traversal's record gathers cheaper than LDG.256 through the L1 data pipe?
//
// A warp-level "gather" = 32 records of 32 bytes (x, y, z, type as four doubles).
//   mode L : every lane loads its record with one LDG.E.256 (what nl_force_kernel does today)
//   mode T : eight lanes issue one gather4 each (four 32-byte rows of a [rows][4] FP64 tensor) into a 1 KB
//            staging buffer of the warp, one mbarrier per stage, STAGES gathers in flight per warp; after the
//            barrier flips every lane reads its record back from shared memory (2 x LDS.128)
// Index patterns: consecutive, random in a 2048-record window, and REAL: the lists of a periodic LJ fluid
// (rho = 0.8, reach 2.8, cells of edge 2.8, x-fastest cell order) walked the way the traversal walks them
// (8 row atoms per warp, 4 lanes per atom, lane l takes entries 4 r + l).
// Reports SM cycles per warp-level gather.  Every barrier wait is bounded: a wrong transaction count prints
// "TIMEOUT" instead of hanging the box.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tma_gather4 tma_gather4.cu
#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x)                                                                                   \
    do {                                                                                        \
        cudaError_t e_ = (x);                                                                   \
        if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 1; } \
    } while (0)

__device__ __forceinline__ double4 ld256(const double4 *p) {
    double4 v;
    asm volatile("ld.global.nc.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(v.x), "=d"(v.y), "=d"(v.z), "=d"(v.w) : "l"(p));
    return v;
}

__global__ void __launch_bounds__(128) ldg_kernel(const double4 *__restrict__ recs, const int *__restrict__ idx, int iters,
                                                  int reps, double *out, long long *cycles) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double acc = 0.0;
    const long long t0 = clock64();
    for (int rep = 0; rep < reps; ++rep) {
#pragma unroll 4
        for (int k = 0; k < iters; ++k) {
            const int j = idx[((k + warp * 7 + blockIdx.x * 13) % iters) * 32 + lane];
            const double4 q = ld256(recs + j);
            acc += q.x + q.y + q.z;
        }
    }
    const long long t1 = clock64();
    if (acc == 1.2345) out[0] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ bool mbar_wait(unsigned bar, unsigned parity) {
    const long long t0 = clock64();
    unsigned done = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if (!done && clock64() - t0 > 200000000ll) return false;  // ~0.1 s: give up, never hang
    }
    return true;
}

template <int STAGES>
__global__ void __launch_bounds__(128) tma_kernel(const __grid_constant__ CUtensorMap tmap, const int *__restrict__ idx,
                                                  int iters, int reps, double *out, long long *cycles, int *timeouts) {
    extern __shared__ __align__(1024) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned char *buf = smem + (size_t)warp * STAGES * 1024;
    __shared__ __align__(8) unsigned long long bars[4][STAGES];
    if (lane == 0) {
        for (int s = 0; s < STAGES; ++s)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[warp][s])));
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    const int total = iters * reps;
    auto issue = [&](int it) {
        const int s = it % STAGES;
        const unsigned bar = smem_u32(&bars[warp][s]);
        const int k = (it + warp * 7 + blockIdx.x * 13) % iters;
        if (lane == 0)
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], 1024;" ::"r"(bar) : "memory");
        __syncwarp();
        if (lane < 8) {
            const int4 rows = *reinterpret_cast<const int4 *>(idx + k * 32 + lane * 4);
            const unsigned dst = smem_u32(buf + s * 1024 + lane * 128);
            asm volatile(
                "cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, "
                "%4, %5, %6}], [%7];" ::"r"(dst),
                "l"(&tmap), "r"(0), "r"(rows.x), "r"(rows.y), "r"(rows.z), "r"(rows.w), "r"(bar)
                : "memory");
        }
    };
    double acc = 0.0;
    bool ok = true;
    const long long t0 = clock64();
    for (int it = 0; it < STAGES && it < total; ++it) issue(it);
    for (int it = 0; it < total && ok; ++it) {
        const int s = it % STAGES;
        ok = mbar_wait(smem_u32(&bars[warp][s]), (it / STAGES) & 1);
        const double2 *p = reinterpret_cast<const double2 *>(buf + s * 1024 + lane * 32);
        const double2 a = p[0], b = p[1];
        acc += a.x + a.y + b.x;
        __syncwarp();
        if (it + STAGES < total) issue(it + STAGES);
    }
    const long long t1 = clock64();
    if (!ok && lane == 0) atomicAdd(timeouts, 1);
    if (acc == 1.2345) out[0] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    // the correctness probe: block 0 / warp 0 leaves the sum of what it read in the last iteration
    if (blockIdx.x == 0 && warp == 0) {
        const double2 *p = reinterpret_cast<const double2 *>(buf + ((total - 1) % STAGES) * 1024 + lane * 32);
        out[1 + lane] = p[0].x;
    }
}

// ---- REAL pattern: lists of a periodic fluid, walked like nl_force_kernel walks them ------------------------
static void real_pattern(std::vector<int> &idx, int &iters, int &nrec, bool bucketed) {
    const int nc = 8;
    const double edge = 2.8, L = nc * edge, reach2 = 2.8 * 2.8;
    const int n = (int)(L * L * L * 0.8);
    srand(7);
    std::vector<double> x(n), y(n), z(n);
    for (int i = 0; i < n; ++i) {
        x[i] = L * (rand() / (RAND_MAX + 1.0));
        y[i] = L * (rand() / (RAND_MAX + 1.0));
        z[i] = L * (rand() / (RAND_MAX + 1.0));
    }
    auto cell = [&](int i) { return ((int)(z[i] / edge) * nc + (int)(y[i] / edge)) * nc + (int)(x[i] / edge); };
    std::vector<int> order(n);
    for (int i = 0; i < n; ++i) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return cell(a) < cell(b); });
    std::vector<int> start(nc * nc * nc + 1, 0);
    for (int s = 0; s < n; ++s) start[cell(order[s]) + 1]++;
    for (int c = 0; c < nc * nc * nc; ++c) start[c + 1] += start[c];
    std::vector<std::vector<int>> lists(n);
    for (int s = 0; s < n; ++s) {
        const int i = order[s], c = cell(i), cx = c % nc, cy = (c / nc) % nc, cz = c / (nc * nc);
        for (int dz = -1; dz <= 1; ++dz)
            for (int dy = -1; dy <= 1; ++dy)
                for (int dx = -1; dx <= 1; ++dx) {
                    const int c2 = (((cz + dz + nc) % nc) * nc + (cy + dy + nc) % nc) * nc + (cx + dx + nc) % nc;
                    for (int t = start[c2]; t < start[c2 + 1]; ++t) {
                        if (t == s) continue;
                        const int j = order[t];
                        double ddx = x[i] - x[j], ddy = y[i] - y[j], ddz = z[i] - z[j];
                        ddx -= L * std::nearbyint(ddx / L);
                        ddy -= L * std::nearbyint(ddy / L);
                        ddz -= L * std::nearbyint(ddz / L);
                        if (ddx * ddx + ddy * ddy + ddz * ddz < reach2) lists[s].push_back(t);
                    }
                }
        // (stencil row, slot) order: rows of three cells are contiguous in slot order except across the wrap
    }
    if (bucketed) {
        // lane l of a row's group takes the neighbours whose slot is = l (mod 4): the four lanes of a group then
        // always read four DIFFERENT sector positions of their 128-byte lines.  Streams are evened out at the end
        // (entries moved from the longest to the shortest), as a list build would do it.
        for (auto &li : lists) {
            std::vector<int> st[4];
            for (int t : li) st[t & 3].push_back(t);
            const size_t nn = li.size();
            for (;;) {
                int hi = 0, lo = 0;
                for (int r = 1; r < 4; ++r) {
                    if ((long)st[r].size() - (long)((nn - r + 3) / 4) > (long)st[hi].size() - (long)((nn - hi + 3) / 4)) hi = r;
                    if ((long)st[r].size() - (long)((nn - r + 3) / 4) < (long)st[lo].size() - (long)((nn - lo + 3) / 4)) lo = r;
                }
                if (st[hi].size() <= (nn - hi + 3) / 4) break;
                st[lo].push_back(st[hi].back());
                st[hi].pop_back();
            }
            for (size_t k = 0; k < nn; ++k) li[k] = st[k & 3][k >> 2];
        }
    }
    // warp = 8 consecutive rows; round r: lane 4 g + l gathers entry 4 r + l of row g (idle lanes repeat entry 0)
    idx.clear();
    for (int base = 0; base + 8 <= n && (int)idx.size() < 32 * 4096; base += 8 * 37) {  // scattered row groups
        size_t rounds = 0;
        for (int g = 0; g < 8; ++g) rounds = std::max(rounds, (lists[base + g].size() + 3) / 4);
        for (size_t r = 0; r < rounds; ++r)
            for (int lane = 0; lane < 32; ++lane) {
                const auto &li = lists[base + lane / 4];
                const size_t e = 4 * r + lane % 4;
                idx.push_back(li[e < li.size() ? e : 0]);
            }
    }
    iters = (int)idx.size() / 32;
    nrec = n;
    double mean = 0;
    for (auto &l : lists) mean += l.size();
    printf("# real pattern: %d atoms, %.1f neighbours per atom, %d warp-gathers\n", n, mean / n, iters);
}

typedef CUresult (*EncodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    const int blocks_per_sm = 5, blocks = 148 * blocks_per_sm, threads = 128, reps = 4;
    EncodeTiled encode = nullptr;
    {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
        if (!fn || q != cudaDriverEntryPointSuccess) { printf("cuTensorMapEncodeTiled not available\n"); return 1; }
        encode = (EncodeTiled)fn;
    }
    double *out; long long *cyc; int *timeouts;
    CK(cudaMalloc(&out, 64 * sizeof(double)));
    CK(cudaMalloc(&cyc, sizeof(long long) * blocks));
    CK(cudaMalloc(&timeouts, sizeof(int)));
    std::vector<long long> hc(blocks);

    for (int pat = 0; pat < 4; ++pat) {
        std::vector<int> h;
        int iters = 256, nrec = 2048 * 64;
        const char *name = pat == 0 ? "consecutive" : (pat == 1 ? "random in a 2048-record window"
                           : (pat == 2 ? "REAL lists (rho 0.8, reach 2.8)" : "REAL lists, lane l <- slot%4 == l"));
        if (pat >= 2) real_pattern(h, iters, nrec, pat == 3);
        else {
            srand(1234);
            h.resize((size_t)iters * 32);
            for (int k = 0; k < iters; ++k) {
                const int b0 = rand() % (2048 - 256);
                for (int l = 0; l < 32; ++l) h[k * 32 + l] = pat == 0 ? b0 + l : rand() % 2048;
            }
        }
        double4 *recs; int *idx;
        CK(cudaMalloc(&recs, sizeof(double4) * (size_t)nrec));
        std::vector<double4> hr(nrec);
        for (int i = 0; i < nrec; ++i) hr[i] = make_double4(i, 2.0 * i, 3.0 * i, 0.0);
        CK(cudaMemcpy(recs, hr.data(), sizeof(double4) * (size_t)nrec, cudaMemcpyHostToDevice));
        CK(cudaMalloc(&idx, sizeof(int) * h.size()));
        CK(cudaMemcpy(idx, h.data(), sizeof(int) * h.size(), cudaMemcpyHostToDevice));

        CUtensorMap tmap;
        const cuuint64_t gdim[2] = {4, (cuuint64_t)nrec};
        const cuuint64_t gstride[1] = {32};
        const cuuint32_t box[2] = {4, 1}, estr[2] = {1, 1};
        const CUresult r = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, recs, gdim, gstride, box, estr,
                                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed: %d\n", (int)r); return 1; }

        auto report = [&](const char *mode) -> int {
            CK(cudaDeviceSynchronize());
            CK(cudaMemcpy(hc.data(), cyc, sizeof(long long) * blocks, cudaMemcpyDeviceToHost));
            double mean = 0;
            for (int b = 0; b < blocks; ++b) mean += hc[b];
            mean /= blocks;
            // blocks_per_sm blocks x 4 warps each issue iters * reps gathers on an SM: SM cycles per warp-gather
            printf("%-34s | %-22s %7.2f cyc / warp-gather  (%5.2f cyc / record)\n", name, mode,
                   mean / ((double)iters * reps * 4 * blocks_per_sm), mean / ((double)iters * reps * 4 * blocks_per_sm) / 32.0);
            return 0;
        };
        for (int w = 0; w < 2; ++w) ldg_kernel<<<blocks, threads>>>(recs, idx, iters, reps, out, cyc);
        if (report("LDG.256 per lane")) return 1;
#define RUN_TMA(S)                                                                                                  \
    do {                                                                                                            \
        CK(cudaMemset(timeouts, 0, sizeof(int)));                                                                   \
        CK(cudaFuncSetAttribute(tma_kernel<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * S * 1024));          \
        for (int w = 0; w < 2; ++w) tma_kernel<S><<<blocks, threads, 4 * S * 1024>>>(tmap, idx, iters, reps, out, cyc, timeouts); \
        char label[64];                                                                                             \
        snprintf(label, sizeof(label), "TMA gather4, %d stages", S);                                                \
        if (report(label)) return 1;                                                                                \
        int ht = 0;                                                                                                 \
        double ho[33];                                                                                              \
        CK(cudaMemcpy(&ht, timeouts, sizeof(int), cudaMemcpyDeviceToHost));                                         \
        CK(cudaMemcpy(ho, out, sizeof(ho), cudaMemcpyDeviceToHost));                                                \
        const int klast = ((iters * reps - 1) + 0 * 7 + 0 * 13) % iters;                                            \
        int wrong = 0;                                                                                              \
        for (int l = 0; l < 32; ++l) wrong += ho[1 + l] != (double)h[klast * 32 + l];                               \
        if (ht || wrong) printf("    !! %d warps TIMED OUT, %d of 32 probe records wrong\n", ht, wrong);             \
    } while (0)
        RUN_TMA(2);
        RUN_TMA(4);
        RUN_TMA(8);
        cudaFree(recs);
        cudaFree(idx);
    }
    return 0;
}
