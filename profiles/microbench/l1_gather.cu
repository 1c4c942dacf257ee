// Micro-benchmark: cost of gathering 32-byte records (the neighbour-list traversal's access
// pattern) through L1 with LDG.256 / 2x LDG.128 and through shared memory, for several lane->record
// patterns.  Reports SM cycles per warp-level gather.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ double4 ld256(const double4 *p) {
    double4 v;
    asm volatile("ld.global.nc.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(v.x), "=d"(v.y), "=d"(v.z), "=d"(v.w) : "l"(p));
    return v;
}

// idx table: [iters][32] record indices inside a window of `window` records (per block offset added)
template <int MODE>  // 0 = LDG.256, 1 = 2x LDG.128, 2 = smem AoS 2x LDS.128, 3 = smem SoA 3x LDS.64, 4 = 2x tex1Dfetch<int4>, 5 = LDG.256 and tex alternating
__global__ void __launch_bounds__(256) gather_kernel(const double4 *__restrict__ recs, const int *__restrict__ idx,
                                                     int iters, int window, double *out, long long *cycles,
                                                     cudaTextureObject_t tex) {
    extern __shared__ double smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const double4 *base = recs + (size_t)(blockIdx.x % 64) * window;
    if (MODE == 2 || MODE == 3) {
        for (int i = threadIdx.x; i < window; i += blockDim.x) {
            const double4 r = base[i];
            if (MODE == 2) { smem[4 * i] = r.x; smem[4 * i + 1] = r.y; smem[4 * i + 2] = r.z; smem[4 * i + 3] = r.w; }
            else { smem[i] = r.x; smem[window + i] = r.y; smem[2 * window + i] = r.z; }
        }
        __syncthreads();
    }
    double acc = 0.0;
    const long long t0 = clock64();
    for (int rep = 0; rep < 8; ++rep) {
#pragma unroll 4
        for (int k = 0; k < iters; ++k) {
            const int j = idx[((k + warp * 7) % iters) * 32 + lane];
            if (MODE == 0) { const double4 q = ld256(base + j); acc += q.x + q.y + q.z; }
            else if (MODE == 4 || (MODE == 5 && (k & 1))) {
                const int t = ((blockIdx.x % 64) * window + j) * 2;
                const int4 a = tex1Dfetch<int4>(tex, t), b = tex1Dfetch<int4>(tex, t + 1);
                acc += __hiloint2double(a.y, a.x) + __hiloint2double(a.w, a.z) + __hiloint2double(b.y, b.x);
            }
            else if (MODE == 5) { const double4 q = ld256(base + j); acc += q.x + q.y + q.z; }
            else if (MODE == 1) {
                const double2 *p = reinterpret_cast<const double2 *>(base + j);
                const double2 a = __ldg(p), b = __ldg(p + 1);
                acc += a.x + a.y + b.x;
            } else if (MODE == 2) {
                const double2 *p = reinterpret_cast<const double2 *>(smem + 4 * j);
                const double2 a = p[0], b = p[1];
                acc += a.x + a.y + b.x;
            } else { acc += smem[j] + smem[window + j] + smem[2 * window + j]; }
        }
    }
    const long long t1 = clock64();
    if (acc == 1.2345) out[0] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

int main() {
    const int window = 2048, iters = 256, blocks = 148 * 2, threads = 256;
    double4 *recs; int *idx; double *out; long long *cyc;
    cudaMalloc(&recs, sizeof(double4) * window * 64);
    cudaMemset(recs, 0, sizeof(double4) * window * 64);
    cudaMalloc(&idx, sizeof(int) * iters * 32);
    cudaMalloc(&out, 8); cudaMalloc(&cyc, sizeof(long long) * blocks);
    cudaTextureObject_t tex = 0;
    {
        cudaResourceDesc rd = {};
        rd.resType = cudaResourceTypeLinear;
        rd.res.linear.devPtr = recs;
        rd.res.linear.desc = cudaCreateChannelDesc<int4>();
        rd.res.linear.sizeInBytes = sizeof(double4) * window * 64;
        cudaTextureDesc td = {};
        td.readMode = cudaReadModeElementType;
        if (cudaCreateTextureObject(&tex, &rd, &td, nullptr) != cudaSuccess) { printf("texture object failed\n"); return 1; }
    }
    int *h = (int *)malloc(sizeof(int) * iters * 32);
    long long *hc = (long long *)malloc(sizeof(long long) * blocks);
    const char *names[] = {"consecutive (8 lines)", "stride2 (16 lines)", "stride4 (32 lines)", "random in 64-record window",
                           "broadcast (1 record)", "8 groups x 4 consecutive, groups random in 128", "pairs share a record (16 consecutive)",
                           "random in 2048 window", "4 groups x 8 consecutive, groups random in 128"};
    for (int pat = 0; pat < 9; ++pat) {
        srand(1234);
        for (int k = 0; k < iters; ++k) {
            const int b0 = rand() % (window - 256);
            int g[8]; for (int q = 0; q < 8; ++q) g[q] = b0 + rand() % 124;
            for (int l = 0; l < 32; ++l) {
                int j;
                switch (pat) {
                    case 0: j = b0 + l; break;
                    case 1: j = b0 + 2 * l; break;
                    case 2: j = b0 + 4 * l; break;
                    case 3: j = b0 + rand() % 64; break;
                    case 4: j = b0; break;
                    case 5: j = g[l / 4] + l % 4; break;
                    case 6: j = b0 + l / 2; break;
                    case 7: j = rand() % window; break;
                    default: j = g[l / 8] + l % 8; break;
                }
                h[k * 32 + l] = j;
            }
        }
        cudaMemcpy(idx, h, sizeof(int) * iters * 32, cudaMemcpyHostToDevice);
        printf("%-48s", names[pat]);
        for (int mode = 0; mode < 6; ++mode) {
            const size_t smem = (mode == 2 || mode == 3) ? sizeof(double) * 4 * window : 0;
            for (int r = 0; r < 2; ++r) {
                switch (mode) {
                    case 0: gather_kernel<0><<<blocks, threads, smem>>>(recs, idx, iters, window, out, cyc, tex); break;
                    case 1: gather_kernel<1><<<blocks, threads, smem>>>(recs, idx, iters, window, out, cyc, tex); break;
                    case 2: cudaFuncSetAttribute(gather_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                            gather_kernel<2><<<blocks, threads, smem>>>(recs, idx, iters, window, out, cyc, tex); break;
                    case 3: cudaFuncSetAttribute(gather_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                             gather_kernel<3><<<blocks, threads, smem>>>(recs, idx, iters, window, out, cyc, tex); break;
                    case 4: gather_kernel<4><<<blocks, threads, 0>>>(recs, idx, iters, window, out, cyc, tex); break;
                    default: gather_kernel<5><<<blocks, threads, 0>>>(recs, idx, iters, window, out, cyc, tex); break;
                }
                cudaDeviceSynchronize();
            }
            cudaMemcpy(hc, cyc, sizeof(long long) * blocks, cudaMemcpyDeviceToHost);
            double mean = 0; for (int b = 0; b < blocks; ++b) mean += hc[b]; mean /= blocks;
            // 2 blocks/SM x 8 warps each issue iters*8 gathers: SM cycles per warp-gather
            const double per = mean / (8.0 * iters * 8 * 2);
            printf(" | mode%d %6.2f cyc/gather", mode, per);
        }
        printf("\n");
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
    }
    return 0;
}
