// TEST INFRASTRUCTURE (tests/test_npy_normal.py): just enough of the CUDA execution model to run the passes of
// turtlemd_b200/csrc/npy_normal.cu -- the SAME source the GPU library is built from -- as host threads, so that the chunk /
// super-chunk bookkeeping of the device stream is checked against NumPy on a machine without a GPU.  One block at a time,
// one std::thread per CUDA thread, __syncthreads as a barrier, __shared__ as a function-local static (blocks do not
// overlap).  Nothing of the product includes this file.
#pragma once

#include <math.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <barrier>
#include <memory>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __shared__ static
#define __forceinline__ inline
#define __noinline__
#define __restrict__
#define __launch_bounds__(...)
#define __CUDACC_EMULATED__ 1

struct EmulDim3 {
    unsigned x = 0, y = 0, z = 0;
};
static thread_local EmulDim3 threadIdx;
static EmulDim3 blockIdx, blockDim, gridDim;

struct uint2 {
    uint32_t x, y;
};
static inline uint2 make_uint2(uint32_t x, uint32_t y) { return uint2{x, y}; }
static inline int __popc(uint32_t v) { return __builtin_popcount(v); }
static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v) {
    return __atomic_fetch_add(p, v, __ATOMIC_RELAXED);
}
using std::max;
using std::min;

namespace emul {
static std::unique_ptr<std::barrier<>> g_barrier;
static std::vector<std::unique_ptr<std::barrier<>>> g_warp_barriers;  // __syncwarp: the 32 threads of a warp

template <class F>
static void run(unsigned grid, unsigned block, F body) {
    gridDim.x = grid;
    blockDim.x = block;
    for (unsigned b = 0; b < grid; ++b) {
        blockIdx.x = b;
        g_barrier = std::make_unique<std::barrier<>>((std::ptrdiff_t)block);
        g_warp_barriers.clear();
        for (unsigned w = 0; w < block; w += 32)
            g_warp_barriers.push_back(std::make_unique<std::barrier<>>((std::ptrdiff_t)std::min(32u, block - w)));
        std::vector<std::thread> threads;
        threads.reserve(block);
        for (unsigned t = 0; t < block; ++t)
            threads.emplace_back([t, &body] {
                threadIdx.x = t;
                body();
                g_barrier->arrive_and_drop();  // a thread that has returned no longer takes part in later barriers
                g_warp_barriers[t / 32]->arrive_and_drop();
            });
        for (auto &th : threads) th.join();
    }
}

template <class K>
struct Bound {
    K kernel;
    unsigned grid, block;
    template <class... A>
    void operator()(A... args) {
        run(grid, block, [&] { kernel(args...); });
    }
};
template <class K>
static Bound<K> bind(K kernel, unsigned grid, unsigned block) {
    return Bound<K>{kernel, grid, block};
}
}  // namespace emul

static inline void __syncthreads() { emul::g_barrier->arrive_and_wait(); }
static inline void __syncwarp() { emul::g_warp_barriers[threadIdx.x / 32]->arrive_and_wait(); }
#define TMD_LAUNCH(kernel, grid, block, stream) ::emul::bind(kernel, (grid), (block))

// ---- the slice of the runtime API / library plumbing the file uses ---------------------------------------------
typedef int cudaError_t;
typedef void *cudaStream_t;
typedef void *tmd_stream_t;
enum { cudaSuccess = 0, cudaMemcpyDeviceToHost = 2 };
enum { TMD_OK = 0, TMD_ERR_INVALID = 1, TMD_ERR_NO_DEVICE = 2, TMD_ERR_CUDA = 3, TMD_ERR_NOMEM = 4 };

template <class T>
static inline cudaError_t cudaMalloc(T **p, size_t bytes) {
    *p = static_cast<T *>(calloc(1, bytes ? bytes : 1));
    return *p ? cudaSuccess : 2;
}
static inline cudaError_t cudaFree(void *p) {
    free(p);
    return cudaSuccess;
}
static inline cudaError_t cudaMemcpyAsync(void *dst, const void *src, size_t n, int, cudaStream_t) {
    memcpy(dst, src, n);
    return cudaSuccess;
}
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaGetDevice(int *d) {
    *d = 0;
    return cudaSuccess;
}

namespace tmd {
static char g_emul_error[512];
static inline void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_emul_error, sizeof g_emul_error, fmt, ap);
    va_end(ap);
}
static inline int cuda_fail(cudaError_t, const char *what, const char *, int) {
    set_error("emulated CUDA failure in %s", what);
    return TMD_ERR_CUDA;
}
static inline int require_device() { return TMD_OK; }
static inline void count_launch() {}
static inline cudaStream_t as_stream(tmd_stream_t s) { return s; }
}  // namespace tmd

#define TMD_CUDA(call)                                                                                   \
    do {                                                                                                 \
        if ((call) != cudaSuccess) return ::tmd::cuda_fail(1, #call, __FILE__, __LINE__);                \
    } while (0)
#define TMD_REQUIRE(cond, msg)                                                                           \
    do {                                                                                                 \
        if (!(cond)) {                                                                                   \
            ::tmd::set_error("%s: %s", __func__, msg);                                                   \
            return TMD_ERR_INVALID;                                                                      \
        }                                                                                                \
    } while (0)
#define TMD_CHECK(expr)                                                                                  \
    do {                                                                                                 \
        int tmd_rc__ = (expr);                                                                           \
        if (tmd_rc__ != TMD_OK) return tmd_rc__;                                                         \
    } while (0)

extern "C" const char *emul_last_error(void) { return tmd::g_emul_error; }
