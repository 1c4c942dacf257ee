// Shared helpers for the turtlemd_b200 CUDA library (sm_100a).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/turtlemd_b200.h"

namespace tmd {

// ---- error plumbing -----------------------------------------------------------------------
void set_error(const char *fmt, ...);
int cuda_fail(cudaError_t err, const char *what, const char *file, int line);
int require_device();  // TMD_OK or TMD_ERR_NO_DEVICE

#define TMD_CUDA(call)                                                          \
    do {                                                                        \
        cudaError_t tmd_err__ = (call);                                         \
        if (tmd_err__ != cudaSuccess)                                           \
            return ::tmd::cuda_fail(tmd_err__, #call, __FILE__, __LINE__);      \
    } while (0)

#define TMD_REQUIRE(cond, msg)                                                  \
    do {                                                                        \
        if (!(cond)) {                                                          \
            ::tmd::set_error("%s: %s", __func__, msg);                          \
            return TMD_ERR_INVALID;                                             \
        }                                                                       \
    } while (0)

#define TMD_CHECK(expr)                                                         \
    do {                                                                        \
        int tmd_rc__ = (expr);                                                  \
        if (tmd_rc__ != TMD_OK) return tmd_rc__;                                \
    } while (0)

// ---- per-device scratch arena -------------------------------------------------------------
// Grows only; the hot path never allocates once the first step has run.  One arena per device,
// several named slots so that independent kernels of one step do not alias.  A slot that has to grow
// gets a new block and the old one is RETIRED, never freed while the process lives (captured CUDA
// graphs and launches still in flight keep their pointers valid).  The slots are shared by every
// stream of a device: callers that run the library on several streams of one device at once must
// order those streams themselves (the Python layer uses one stream per process).
enum ArenaSlot {
    SLOT_PARTIAL_F = 0,   // per-split force partials of the pair kernels
    SLOT_PARTIAL_VW,      // per-block (V, W) partials
    SLOT_PARTIAL_C,       // symmetric all-pairs kernel: column (reaction) force partials
    SLOT_WRAPPED,         // wrapped / padded copy of the positions (double4 per atom)
    SLOT_TABLE,           // LJ type table on the device
    SLOT_CELL_KEYS,       // cell index per atom
    SLOT_CELL_COUNTS,     // atoms per cell, then exclusive prefix
    SLOT_CELL_ORDER,      // sorted -> original index
    SLOT_CELL_SORTED,     // sorted, wrapped double4 positions (+type)
    SLOT_CELL_SCRATCH,    // scan scratch
    SLOT_HOST_A,          // staging of the tmd_host_* entry points
    SLOT_HOST_B,
    SLOT_HOST_C,
    SLOT_HOST_D,
    SLOT_HOST_E,
    SLOT_HOST_F,
    SLOT_MISC,
    SLOT_COUNT
};
int arena_get(ArenaSlot slot, size_t bytes, void **ptr);
int num_sms();
// every kernel launch of the library goes through this (tmd_launch_count: bench.py's gpu_launches)
void count_launch();
long long launch_count_value();
void add_launches_value(long long delta);  // recording a CUDA graph is not launching: take the count back

static inline cudaStream_t as_stream(tmd_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

// ---- packed single precision (FADD2 / FMUL2 / FFMA2 on sm_100a): two floats per 64-bit register, each
// operation individually rounded like its scalar form -------------------------------------------
__device__ __forceinline__ unsigned long long pack2(float lo, float hi) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(unsigned long long v, float &lo, float &hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ unsigned long long sub2(unsigned long long a, unsigned long long b) {
    unsigned long long r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ unsigned long long mul2(unsigned long long a, unsigned long long b) {
    unsigned long long r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}

__device__ __forceinline__ unsigned long long add2(unsigned long long a, unsigned long long b) {
    unsigned long long r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}

// ---- device-side reductions ---------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    return v;
}

// Sum NV values per thread over a block of NT threads (NT multiple of 32); result valid in
// thread 0.  Fixed order: lanes by butterfly, then warps in index order -> deterministic.
template <int NV, int NT>
__device__ __forceinline__ void block_sum(double (&v)[NV], double *smem /* NV * NT/32 */) {
    constexpr int NW = NT / 32;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < NV; ++k) v[k] = warp_sum(v[k]);
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < NV; ++k) smem[k * NW + warp] = v[k];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            double s = smem[k * NW];
            for (int w = 1; w < NW; ++w) s += smem[k * NW + w];
            v[k] = s;
        }
    }
    __syncthreads();
}

// Deterministic final reduction of per-block (V, W[9]) partials into vw[10].
int launch_reduce_vw(const double *partials, int nblocks, double scale, uint32_t flags,
                     double *vw, cudaStream_t stream);

}  // namespace tmd
