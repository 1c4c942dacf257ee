// Library plumbing: errors, device selection, scratch arena, memory helpers, the common
// (V, W) reduction and the two roofline micro-benchmarks (DFMA peak, copy bandwidth).
#include <stdarg.h>
#include <string.h>

#include <atomic>
#include <mutex>
#include <vector>

#include "common.cuh"

namespace tmd {

static thread_local char g_error[512] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t err, const char *what, const char *file, int line) {
    set_error("CUDA error %d (%s) in %s at %s:%d", (int)err, cudaGetErrorString(err), what, file, line);
    if (err == cudaErrorNoDevice || err == cudaErrorInsufficientDriver) return TMD_ERR_NO_DEVICE;
    if (err == cudaErrorMemoryAllocation) return TMD_ERR_NOMEM;
    return TMD_ERR_CUDA;
}

int require_device() {
    int count = 0;
    cudaError_t err = cudaGetDeviceCount(&count);
    if (err != cudaSuccess || count == 0) {
        cudaGetLastError();
        set_error("no CUDA device available (%s); turtlemd_b200 has no CPU path",
                  err == cudaSuccess ? "device count is 0" : cudaGetErrorString(err));
        return TMD_ERR_NO_DEVICE;
    }
    return TMD_OK;
}

static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
long long launch_count_value() { return g_launches.load(std::memory_order_relaxed); }
void add_launches_value(long long delta) { g_launches.fetch_add(delta, std::memory_order_relaxed); }

// ---- arena ---------------------------------------------------------------------------------
struct Arena {
    void *ptr[SLOT_COUNT];
    size_t cap[SLOT_COUNT];
    std::vector<void *> retired;  // outgrown blocks: see arena_get
};
static constexpr int kMaxDevices = 32;
static Arena g_arena[kMaxDevices];
static int g_sms[kMaxDevices];
static std::mutex g_arena_mutex;

int arena_get(ArenaSlot slot, size_t bytes, void **ptr) {
    int dev = 0;
    TMD_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= kMaxDevices) {
        set_error("device index %d out of range", dev);
        return TMD_ERR_INVALID;
    }
    std::lock_guard<std::mutex> lock(g_arena_mutex);
    Arena &a = g_arena[dev];
    if (a.cap[slot] < bytes) {
        // A slot that grows gets a NEW block; the old one is retired, not freed: CUDA graphs captured by the
        // decompositions (parallel.py) bake scratch pointers, and kernels of another stream may still be
        // using the old block.  Growth is geometric, so the retired blocks of a slot sum to less than its
        // final size; they are released with the process.
        if (a.ptr[slot]) a.retired.push_back(a.ptr[slot]);
        const size_t want = bytes + bytes / 2 + 4096;
        void *fresh = nullptr;
        TMD_CUDA(cudaMalloc(&fresh, want));
        a.ptr[slot] = fresh;
        a.cap[slot] = want;
    }
    *ptr = a.ptr[slot];
    return TMD_OK;
}

int num_sms() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return 148;
    if (g_sms[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        g_sms[dev] = n;
    }
    return g_sms[dev];
}

// ---- (V, W) reduction ------------------------------------------------------------------------
// partials: [nblocks][10].  960 threads read the flat array coalesced; 960 is a multiple of 10, so a
// thread always meets the same component.  Fixed order everywhere -> deterministic.
constexpr int kReduceThreads = 960;

__global__ void __launch_bounds__(kReduceThreads) reduce_vw_kernel(const double *__restrict__ partials, int nblocks,
                                                                   double scale, uint32_t flags,
                                                                   double *__restrict__ vw) {
    __shared__ double smem[kReduceThreads];
    const int total = nblocks * 10;
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    int e = threadIdx.x;
    for (; e + 3 * kReduceThreads < total; e += 4 * kReduceThreads) {  // four loads in flight
#pragma unroll
        for (int u = 0; u < 4; ++u) acc[u] += partials[e + u * kReduceThreads];
    }
    for (int u = 0; e < total; e += kReduceThreads, ++u) acc[u & 3] += partials[e];
    smem[threadIdx.x] = (acc[0] + acc[1]) + (acc[2] + acc[3]);
    __syncthreads();
    if (threadIdx.x < 10) {
        double s = 0.0;
        for (int q = threadIdx.x; q < kReduceThreads; q += 10) s += smem[q];
        // quantities that were not asked for are left untouched (force() must not clobber v_pot)
        const bool accumulate = flags & TMD_ACCUMULATE;
        const int k = threadIdx.x;
        if (k == 0 ? (flags & TMD_WANT_V) : (flags & TMD_WANT_W)) vw[k] = (accumulate ? vw[k] : 0.0) + scale * s;
    }
}

int launch_reduce_vw(const double *partials, int nblocks, double scale, uint32_t flags, double *vw,
                     cudaStream_t stream) {
    ::tmd::count_launch(), reduce_vw_kernel<<<1, kReduceThreads, 0, stream>>>(partials, nblocks, scale, flags, vw);
    TMD_CUDA(cudaGetLastError());
    return TMD_OK;
}

// ---- micro-benchmarks ------------------------------------------------------------------------
// 16 independent DFMA chains per thread; 8 warps/SMSP resident -> the FP64 pipe is the only limit.
__global__ void __launch_bounds__(256) dfma_peak_kernel(double *out, int iters, double seed) {
    double a[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) a[k] = seed + k * 1e-3 + threadIdx.x * 1e-6;
    const double m = 1.0000001, c = 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 16; ++k) a[k] = fma(a[k], m, c);
    }
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < 16; ++k) s += a[k];
    if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;  // never true; keeps the chain alive
}

__global__ void __launch_bounds__(256) copy_kernel(const double2 *__restrict__ src, double2 *__restrict__ dst, size_t n2) {
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += stride) dst[i] = src[i];
}

}  // namespace tmd

using namespace tmd;

extern "C" {

int tmd_abi_version(void) { return TMD_ABI_VERSION; }

const char *tmd_last_error(void) { return g_error; }

const char *tmd_status_string(int status) {
    switch (status) {
        case TMD_OK: return "ok";
        case TMD_ERR_INVALID: return "invalid argument";
        case TMD_ERR_CUDA: return "CUDA runtime error";
        case TMD_ERR_NO_DEVICE: return "no CUDA device (no CPU fallback exists)";
        case TMD_ERR_UNSUPPORTED: return "outside the supported hot path";
        case TMD_ERR_NOMEM: return "out of device memory";
        default: return "unknown status";
    }
}

int tmd_launch_count(int64_t *launches) {
    TMD_REQUIRE(launches != nullptr, "launches is NULL");
    *launches = g_launches.load(std::memory_order_relaxed);
    return TMD_OK;
}

int tmd_device_count(int *count) {
    TMD_REQUIRE(count != nullptr, "count is NULL");
    cudaError_t err = cudaGetDeviceCount(count);
    if (err != cudaSuccess) {
        cudaGetLastError();
        *count = 0;
    }
    return TMD_OK;
}

int tmd_set_device(int device) {
    TMD_CHECK(require_device());
    TMD_CUDA(cudaSetDevice(device));
    return TMD_OK;
}

int tmd_device_info(int device, char *name, int *sm, size_t *total_mem, int *sms) {
    TMD_CHECK(require_device());
    cudaDeviceProp prop;
    TMD_CUDA(cudaGetDeviceProperties(&prop, device));
    if (name) {
        strncpy(name, prop.name, 255);
        name[255] = 0;
    }
    if (sm) *sm = prop.major * 10 + prop.minor;
    if (total_mem) *total_mem = prop.totalGlobalMem;
    if (sms) *sms = prop.multiProcessorCount;
    return TMD_OK;
}

int tmd_stream_synchronize(tmd_stream_t stream) {
    TMD_CHECK(require_device());
    TMD_CUDA(cudaStreamSynchronize(as_stream(stream)));
    return TMD_OK;
}

int tmd_malloc(void **dptr, size_t bytes) {
    TMD_REQUIRE(dptr != nullptr, "dptr is NULL");
    TMD_CHECK(require_device());
    TMD_CUDA(cudaMalloc(dptr, bytes ? bytes : 1));
    return TMD_OK;
}

int tmd_free(void *dptr) {
    if (!dptr) return TMD_OK;
    TMD_CUDA(cudaFree(dptr));
    return TMD_OK;
}

int tmd_memset_async(void *dptr, int value, size_t bytes, tmd_stream_t stream) {
    TMD_CUDA(cudaMemsetAsync(dptr, value, bytes, as_stream(stream)));
    return TMD_OK;
}

int tmd_memcpy_h2d_async(void *dptr, const void *hptr, size_t bytes, tmd_stream_t stream) {
    TMD_CUDA(cudaMemcpyAsync(dptr, hptr, bytes, cudaMemcpyHostToDevice, as_stream(stream)));
    return TMD_OK;
}

int tmd_memcpy_d2h_async(void *hptr, const void *dptr, size_t bytes, tmd_stream_t stream) {
    TMD_CUDA(cudaMemcpyAsync(hptr, dptr, bytes, cudaMemcpyDeviceToHost, as_stream(stream)));
    return TMD_OK;
}

int tmd_memcpy_d2d_async(void *dst, const void *src, size_t bytes, tmd_stream_t stream) {
    TMD_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, as_stream(stream)));
    return TMD_OK;
}

int tmd_host_register(void *hptr, size_t bytes) {
    TMD_CHECK(require_device());
    TMD_CUDA(cudaHostRegister(hptr, bytes, cudaHostRegisterDefault));
    return TMD_OK;
}

int tmd_host_unregister(void *hptr) {
    TMD_CUDA(cudaHostUnregister(hptr));
    return TMD_OK;
}

int tmd_ipc_get_handle(void *dptr, unsigned char handle[64]) {
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    cudaIpcMemHandle_t h;
    TMD_CUDA(cudaIpcGetMemHandle(&h, dptr));
    memcpy(handle, &h, 64);
    return TMD_OK;
}

int tmd_ipc_open_handle(const unsigned char handle[64], void **dptr) {
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    TMD_CUDA(cudaIpcOpenMemHandle(dptr, h, cudaIpcMemLazyEnablePeerAccess));
    return TMD_OK;
}

int tmd_ipc_close_handle(void *dptr) {
    TMD_CUDA(cudaIpcCloseMemHandle(dptr));
    return TMD_OK;
}

int tmd_bench_dfma(int32_t repeats, double *tflops, double *milliseconds) {
    TMD_REQUIRE(tflops && milliseconds, "output pointer is NULL");
    TMD_CHECK(require_device());
    const int iters = 4096, threads = 256;
    const int blocks = num_sms() * 8;
    double *out = nullptr;
    TMD_CHECK(arena_get(SLOT_MISC, (size_t)blocks * threads * sizeof(double), (void **)&out));
    cudaEvent_t e0, e1;
    TMD_CUDA(cudaEventCreate(&e0));
    TMD_CUDA(cudaEventCreate(&e1));
    double best = 1e30;
    for (int r = 0; r < repeats + 2; ++r) {
        TMD_CUDA(cudaEventRecord(e0, 0));
        ::tmd::count_launch(), dfma_peak_kernel<<<blocks, threads>>>(out, iters, 1.0 + r);
        TMD_CUDA(cudaEventRecord(e1, 0));
        TMD_CUDA(cudaEventSynchronize(e1));
        float ms = 0.f;
        TMD_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        if (r >= 2 && ms < best) best = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    double flop = 2.0 * 16.0 * iters * (double)blocks * threads;
    *milliseconds = best;
    *tflops = flop / (best * 1e-3) / 1e12;
    return TMD_OK;
}

int tmd_bench_copy(size_t bytes, int32_t repeats, double *gbs, double *milliseconds) {
    TMD_REQUIRE(gbs && milliseconds, "output pointer is NULL");
    TMD_CHECK(require_device());
    size_t n2 = bytes / sizeof(double2);
    double2 *src = nullptr, *dst = nullptr;
    TMD_CUDA(cudaMalloc(&src, n2 * sizeof(double2)));
    if (cudaMalloc(&dst, n2 * sizeof(double2)) != cudaSuccess) {
        cudaFree(src);
        return cuda_fail(cudaErrorMemoryAllocation, "cudaMalloc(dst)", __FILE__, __LINE__);
    }
    cudaMemset(src, 1, n2 * sizeof(double2));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    double best = 1e30;
    const int blocks = num_sms() * 16;
    for (int r = 0; r < repeats + 2; ++r) {
        cudaEventRecord(e0, 0);
        ::tmd::count_launch(), copy_kernel<<<blocks, 256>>>(src, dst, n2);
        cudaEventRecord(e1, 0);
        cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        if (r >= 2 && ms < best) best = ms;
    }
    cudaError_t err = cudaGetLastError();
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(src);
    cudaFree(dst);
    if (err != cudaSuccess) return cuda_fail(err, "copy_kernel", __FILE__, __LINE__);
    *milliseconds = best;
    *gbs = 2.0 * n2 * sizeof(double2) / (best * 1e-3) / 1e9;
    return TMD_OK;
}

}  // extern "C"
