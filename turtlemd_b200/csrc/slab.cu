// Slab path of the integrators: the state of a rank lives in CELL ORDER (the slot order of its
// neighbour list, lj_nlist.cu) instead of atom order.
//
//   records[slot] = (x, y, z, type) with the lattice vector of the last build removed -- the very
//                   array the pair kernel gathers from, so the position update of the integrator
//                   (integrators.py:90 / :270-274) IS the refresh of the pair kernel's input: no
//                   gather pass, and on several GPUs the records of a rank's boundary slots are the
//                   halo message
//   vel_s, force_s  (slot, dim) ; imass_s (slot)
//
// A rank owns the contiguous slots [slot_first, slot_first + slot_count).  The drift kernels also do
// the Verlet-list displacement test against the build-time records (|x - x_build| > skin/2 -> flag).
// tmd_slab_unsort / tmd_slab_sort move the state between atom order and slot order around a
// rebuild; both can be gated on a device flag so that a single-GPU step needs no host decision.
#include "integrator_math.cuh"

namespace tmd {

template <int DIM>
__device__ __forceinline__ void displacement_flag(const double (&x)[3], const double4 b, double half_skin2, int *flag) {
    double r = (x[0] - b.x) * (x[0] - b.x);
    if (DIM > 1) r = fma(x[1] - b.y, x[1] - b.y, r);
    if (DIM > 2) r = fma(x[2] - b.z, x[2] - b.z, r);
    if (!(r <= half_skin2)) *flag = 1;  // also catches NaN
}

// VelocityVerlet, first half (integrators.py:88-90) on slot-ordered state
template <int DIM>
__global__ void __launch_bounds__(256) slab_vv_kick_drift_kernel(double4 *__restrict__ records,
                                                                 double *__restrict__ vel_s,
                                                                 const double *__restrict__ force_s,
                                                                 const double *__restrict__ imass_s,
                                                                 const double4 *__restrict__ build_records,
                                                                 int64_t first, int64_t count, double dt, double half,
                                                                 double half_skin2, int *flag) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= count) return;
    const int64_t slot = first + t;
    const double im = imass_s[slot];
    double4 rec = records[slot];
    double x[3] = {rec.x, rec.y, rec.z};
#pragma unroll
    for (int k = 0; k < DIM; ++k) {
        const double v = kick(vel_s[slot * DIM + k], force_s[slot * DIM + k], im, half);
        vel_s[slot * DIM + k] = v;
        x[k] = __dadd_rn(x[k], __dmul_rn(dt, v));
    }
    rec.x = x[0];
    rec.y = x[1];
    rec.z = x[2];
    records[slot] = rec;
    displacement_flag<DIM>(x, build_records[slot], half_skin2, flag);
}

// LangevinInertia, part 1 (integrators.py:265-279); the noise of a coordinate is addressed by its
// ATOM index (draws 2*(atom*dim + k) and + 1 of the step), so the trajectory does not depend on the
// cell order or on the number of ranks.
template <int DIM>
__global__ void __launch_bounds__(256) slab_inertia_pre_kernel(double4 *__restrict__ records,
                                                               double *__restrict__ vel_s,
                                                               const double *__restrict__ force_s,
                                                               const double *__restrict__ imass_s,
                                                               const double4 *__restrict__ build_records,
                                                               const int *__restrict__ order, int64_t first,
                                                               int64_t count, tmd_langevin_consts c, uint64_t key,
                                                               uint64_t q_base, double half_skin2, int *flag) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= count) return;
    const int64_t slot = first + t;
    const int64_t atom = order[slot];
    const InertiaPerParticle p = inertia_constants(c, imass_s[slot]);
    double4 rec = records[slot];
    double x[3] = {rec.x, rec.y, rec.z};
#pragma unroll
    for (int k = 0; k < DIM; ++k) {
        double n0, n1;
        two_normals(key, q_base + 2ull * (uint64_t)(atom * DIM + k), true, n0, n1);
        const double xr = __dmul_rn(p.l00, n0);  // norm @ cho.T  (integrators.py:315)
        const double vr = __dadd_rn(__dmul_rn(p.l10, n0), __dmul_rn(p.l11, n1));
        double v = vel_s[slot * DIM + k];
        inertia_pre_update(x[k], v, force_s[slot * DIM + k], xr, vr, c, p);
        vel_s[slot * DIM + k] = v;
    }
    rec.x = x[0];
    rec.y = x[1];
    rec.z = x[2];
    records[slot] = rec;
    displacement_flag<DIM>(x, build_records[slot], half_skin2, flag);
}

// slot order -> atom order: pos = record + image, vel, force (each optional)
template <int DIM>
__global__ void __launch_bounds__(256) slab_unsort_kernel(const double4 *__restrict__ records,
                                                          const double *__restrict__ vel_s,
                                                          const double *__restrict__ force_s,
                                                          const int *__restrict__ order,
                                                          const double *__restrict__ image, int64_t n,
                                                          double *__restrict__ pos, double *__restrict__ vel,
                                                          double *__restrict__ force, const int *gate) {
    if (gate && *gate == 0) return;
    const int64_t slot = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= n) return;
    const int64_t atom = order[slot];
    if (pos) {
        const double4 rec = records[slot];
        const double x[3] = {rec.x, rec.y, rec.z};
#pragma unroll
        for (int k = 0; k < DIM; ++k) pos[atom * DIM + k] = __dadd_rn(x[k], image[atom * 3 + k]);
    }
    if (vel) {
#pragma unroll
        for (int k = 0; k < DIM; ++k) vel[atom * DIM + k] = vel_s[slot * DIM + k];
    }
    if (force) {
#pragma unroll
        for (int k = 0; k < DIM; ++k) force[atom * DIM + k] = force_s[slot * DIM + k];
    }
}

// atom order -> slot order (after a build): vel, imass, force (each optional)
template <int DIM>
__global__ void __launch_bounds__(256) slab_sort_kernel(const int *__restrict__ order, int64_t n,
                                                        const double *__restrict__ vel,
                                                        const double *__restrict__ imass,
                                                        const double *__restrict__ force,
                                                        double *__restrict__ vel_s, double *__restrict__ imass_s,
                                                        double *__restrict__ force_s, const int *gate) {
    if (gate && *gate == 0) return;
    const int64_t slot = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= n) return;
    const int64_t atom = order[slot];
    if (vel_s) {
#pragma unroll
        for (int k = 0; k < DIM; ++k) vel_s[slot * DIM + k] = vel[atom * DIM + k];
    }
    if (imass_s) imass_s[slot] = imass[atom];
    if (force_s) {
#pragma unroll
        for (int k = 0; k < DIM; ++k) force_s[slot * DIM + k] = force[atom * DIM + k];
    }
}

}  // namespace tmd

using namespace tmd;

#define TMD_SLAB_DIM(dim, CALL) \
    do {                        \
        if ((dim) == 1) { CALL(1); } else if ((dim) == 2) { CALL(2); } else { CALL(3); } \
    } while (0)

extern "C" {

int tmd_slab_vv_kick_drift(double *records, double *vel_s, const double *force_s, const double *imass_s,
                           const double *build_records, int64_t slot_first, int64_t slot_count, int32_t dim,
                           double dt, double half_skin2, int32_t *flag, tmd_stream_t stream) {
    TMD_REQUIRE(dim >= 1 && dim <= 3 && slot_first >= 0 && slot_count >= 0, "dim must be 1..3, slots >= 0");
    TMD_REQUIRE(records && vel_s && force_s && imass_s && build_records && flag, "NULL argument");
    TMD_CHECK(require_device());
    if (slot_count == 0) return TMD_OK;
    cudaStream_t s = as_stream(stream);
    const int blocks = (int)((slot_count + 255) / 256);
#define TMD_CALL(D)                                                                                               \
    ::tmd::count_launch(), slab_vv_kick_drift_kernel<D><<<blocks, 256, 0, s>>>(                                    \
        reinterpret_cast<double4 *>(records), vel_s, force_s, imass_s, reinterpret_cast<const double4 *>(build_records), \
        slot_first, slot_count, dt, dt * 0.5, half_skin2, flag)
    TMD_SLAB_DIM(dim, TMD_CALL);
#undef TMD_CALL
    TMD_CUDA(cudaGetLastError());
    return TMD_OK;
}

int tmd_slab_langevin_inertia_pre(double *records, double *vel_s, const double *force_s, const double *imass_s,
                                  const double *build_records, const int32_t *order, int64_t slot_first,
                                  int64_t slot_count, int32_t dim, const tmd_langevin_consts *consts,
                                  const tmd_rng *rng, double half_skin2, int32_t *flag, tmd_stream_t stream) {
    TMD_REQUIRE(dim >= 1 && dim <= 3 && slot_first >= 0 && slot_count >= 0, "dim must be 1..3, slots >= 0");
    TMD_REQUIRE(records && vel_s && force_s && imass_s && build_records && order && flag && consts && rng, "NULL argument");
    TMD_CHECK(require_device());
    if (slot_count == 0) return TMD_OK;
    cudaStream_t s = as_stream(stream);
    const int blocks = (int)((slot_count + 255) / 256);
#define TMD_CALL(D)                                                                                               \
    ::tmd::count_launch(), slab_inertia_pre_kernel<D><<<blocks, 256, 0, s>>>(                                      \
        reinterpret_cast<double4 *>(records), vel_s, force_s, imass_s, reinterpret_cast<const double4 *>(build_records), \
        order, slot_first, slot_count, *consts, rng->key, rng->offset, half_skin2, flag)
    TMD_SLAB_DIM(dim, TMD_CALL);
#undef TMD_CALL
    TMD_CUDA(cudaGetLastError());
    return TMD_OK;
}

int tmd_slab_unsort(const double *records, const double *vel_s, const double *force_s, const int32_t *order,
                    const double *image, int64_t n, int32_t dim, double *pos, double *vel, double *force,
                    const int32_t *gate, tmd_stream_t stream) {
    TMD_REQUIRE(dim >= 1 && dim <= 3 && n >= 0, "dim must be 1..3, n >= 0");
    TMD_REQUIRE(order != nullptr, "order is NULL");
    TMD_REQUIRE(!pos || (records && image), "pos needs records and image");
    TMD_REQUIRE(!vel || vel_s, "vel needs vel_s");
    TMD_REQUIRE(!force || force_s, "force needs force_s");
    TMD_CHECK(require_device());
    if (n == 0) return TMD_OK;
    cudaStream_t s = as_stream(stream);
    const int blocks = (int)((n + 255) / 256);
#define TMD_CALL(D)                                                                                     \
    ::tmd::count_launch(), slab_unsort_kernel<D><<<blocks, 256, 0, s>>>(                                 \
        reinterpret_cast<const double4 *>(records), vel_s, force_s, order, image, n, pos, vel, force, gate)
    TMD_SLAB_DIM(dim, TMD_CALL);
#undef TMD_CALL
    TMD_CUDA(cudaGetLastError());
    return TMD_OK;
}

int tmd_slab_sort(const int32_t *order, int64_t n, int32_t dim, const double *vel, const double *imass,
                  const double *force, double *vel_s, double *imass_s, double *force_s, const int32_t *gate,
                  tmd_stream_t stream) {
    TMD_REQUIRE(dim >= 1 && dim <= 3 && n >= 0, "dim must be 1..3, n >= 0");
    TMD_REQUIRE(order != nullptr, "order is NULL");
    TMD_REQUIRE(!vel_s || vel, "vel_s needs vel");
    TMD_REQUIRE(!imass_s || imass, "imass_s needs imass");
    TMD_REQUIRE(!force_s || force, "force_s needs force");
    TMD_CHECK(require_device());
    if (n == 0) return TMD_OK;
    cudaStream_t s = as_stream(stream);
    const int blocks = (int)((n + 255) / 256);
#define TMD_CALL(D) \
    ::tmd::count_launch(), slab_sort_kernel<D><<<blocks, 256, 0, s>>>(order, n, vel, imass, force, vel_s, imass_s, force_s, gate)
    TMD_SLAB_DIM(dim, TMD_CALL);
#undef TMD_CALL
    TMD_CUDA(cudaGetLastError());
    return TMD_OK;
}

}  // extern "C"
