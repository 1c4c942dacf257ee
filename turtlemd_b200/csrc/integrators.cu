// K6/K7: per-particle integrator updates (turtlemd/integrators.py), element-wise over the flat
// (n*dim) arrays.  HBM-bound: two coordinates per thread through 16-byte loads/stores when the
// arrays allow it.  Arithmetic uses explicit *_rn intrinsics in the order of the reference's NumPy
// expressions, so given the same forces the updates are bit-identical to the reference's.
#include "common.cuh"
#include "philox_normal.h"
#include "doublewell.cuh"
#include "integrator_math.cuh"

namespace tmd {

// ---- element pair access ---------------------------------------------------------------------
struct D2 {
    double a, b;
};

template <bool VEC>
__device__ __forceinline__ D2 ld2(const double *p, int64_t e, bool two) {
    D2 r;
    if (VEC && two) {
        double2 v = *reinterpret_cast<const double2 *>(p + e);
        r.a = v.x;
        r.b = v.y;
    } else {
        r.a = p[e];
        r.b = two ? p[e + 1] : 0.0;
    }
    return r;
}

template <bool VEC>
__device__ __forceinline__ void st2(double *p, int64_t e, bool two, D2 v) {
    if (VEC && two) {
        *reinterpret_cast<double2 *>(p + e) = make_double2(v.a, v.b);
    } else {
        p[e] = v.a;
        if (two) p[e + 1] = v.b;
    }
}

template <int DIM>
__device__ __forceinline__ D2 ld_imass(const double *imass, int64_t e, bool two) {
    D2 r;
    const int64_t i0 = e / DIM;
    r.a = imass[i0];
    if (DIM == 1) {
        r.b = two ? imass[i0 + 1] : 0.0;
    } else {
        const int64_t i1 = (e + 1) / DIM;
        r.b = (i1 == i0 || !two) ? r.a : imass[i1];
    }
    return r;
}

static inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

struct Launch {
    int blocks, threads;
};
static inline Launch ew_launch(int64_t total) {
    Launch l;
    l.threads = 256;
    int64_t pairs = (total + 1) / 2;
    l.blocks = (int)((pairs + l.threads - 1) / l.threads);
    if (l.blocks < 1) l.blocks = 1;
    return l;
}

#define TMD_EW_PROLOGUE                                                   \
    const int64_t e = 2 * ((int64_t)blockIdx.x * blockDim.x + threadIdx.x); \
    if (e >= total) return;                                               \
    const bool two = (e + 1) < total;

__global__ void counter_add_kernel(uint64_t *counter, uint64_t increment) { counter[0] += increment; }

// ---- Velocity Verlet -----------------------------------------------------------------------------
template <int DIM, bool VEC, int KICKS, bool DRIFT>
__global__ void __launch_bounds__(256) vv_kernel(double *__restrict__ pos, double *__restrict__ vel,
                                                 const double *__restrict__ force,
                                                 const double *__restrict__ imass, int64_t total,
                                                 double dt, double half) {
    TMD_EW_PROLOGUE
    D2 v = ld2<VEC>(vel, e, two);
    const D2 f = ld2<VEC>(force, e, two);
    const D2 im = ld_imass<DIM>(imass, e, two);
#pragma unroll
    for (int k = 0; k < KICKS; ++k) {
        v.a = kick(v.a, f.a, im.a, half);
        v.b = kick(v.b, f.b, im.b, half);
    }
    st2<VEC>(vel, e, two, v);
    if (DRIFT) {
        D2 x = ld2<VEC>(pos, e, two);
        x.a = __dadd_rn(x.a, __dmul_rn(dt, v.a));  // integrators.py:90
        x.b = __dadd_rn(x.b, __dmul_rn(dt, v.b));
        st2<VEC>(pos, e, two, x);
    }
}

template <int KICKS, bool DRIFT>
static int launch_vv(double *pos, double *vel, const double *force, const double *imass, int64_t n,
                     int32_t dim, double dt, cudaStream_t stream) {
    const int64_t total = n * dim;
    if (total == 0) return TMD_OK;
    const bool vec = aligned16(vel) && aligned16(force) && (!DRIFT || aligned16(pos));
    const Launch l = ew_launch(total);
    const double half = dt * 0.5;  // VelocityVerlet.half_timestep (integrators.py:82)
#define TMD_VV(D, V)                                                                         \
    ::tmd::count_launch(), vv_kernel<D, V, KICKS, DRIFT><<<l.blocks, l.threads, 0, stream>>>(pos, vel, force, imass, \
                                                                      total, dt, half)
    if (dim == 1) { if (vec) TMD_VV(1, true); else TMD_VV(1, false); }
    else if (dim == 2) { if (vec) TMD_VV(2, true); else TMD_VV(2, false); }
    else { if (vec) TMD_VV(3, true); else TMD_VV(3, false); }
#undef TMD_VV
    TMD_CUDA(cudaGetLastError());
    return TMD_OK;
}

// ---- Verlet -----------------------------------------------------------------------------------------
template <int DIM, bool VEC>
__global__ void __launch_bounds__(256) verlet_kernel(double *__restrict__ pos, double *__restrict__ vel,
                                                     double *__restrict__ prev,
                                                     const double *__restrict__ force,
                                                     const double *__restrict__ imass, int64_t total,
                                                     double dt, double dt2, double half_idt, int first_call) {
    TMD_EW_PROLOGUE
    const D2 x = ld2<VEC>(pos, e, two);
    const D2 f = ld2<VEC>(force, e, two);
    const D2 im = ld_imass<DIM>(imass, e, two);
    D2 xp;
    if (first_call) {  // previous_pos = pos - vel*dt  (integrators.py:56-57)
        const D2 v = ld2<VEC>(vel, e, two);
        xp.a = __dsub_rn(x.a, __dmul_rn(v.a, dt));
        xp.b = __dsub_rn(x.b, __dmul_rn(v.b, dt));
    } else {
        xp = ld2<VEC>(prev, e, two);
    }
    D2 xn, vn;
    // pos = 2*pos - previous + acc*dt^2 ; vel = (pos - previous) * (0.5/dt)   (integrators.py:58-64)
    xn.a = __dadd_rn(__dsub_rn(__dmul_rn(2.0, x.a), xp.a), __dmul_rn(__dmul_rn(f.a, im.a), dt2));
    xn.b = __dadd_rn(__dsub_rn(__dmul_rn(2.0, x.b), xp.b), __dmul_rn(__dmul_rn(f.b, im.b), dt2));
    vn.a = __dmul_rn(__dsub_rn(xn.a, xp.a), half_idt);
    vn.b = __dmul_rn(__dsub_rn(xn.b, xp.b), half_idt);
    st2<VEC>(vel, e, two, vn);
    st2<VEC>(prev, e, two, x);
    st2<VEC>(pos, e, two, xn);
}

// ---- noise ------------------------------------------------------------------------------------------
// ---- Langevin, overdamped ------------------------------------------------------------------------
// NOISE: 0 = the in-kernel Philox stream, 1 = noise drawn by the caller (already scaled), 2 = STANDARD normals of the
// caller (tmd_npstream_normals: NumPy's own stream generated on the device), scaled here like the Philox ones
template <int DIM, bool VEC, int NOISE>
__global__ void __launch_bounds__(256) overdamped_kernel(double *__restrict__ pos, double *__restrict__ vel,
                                                         const double *__restrict__ force,
                                                         const double *__restrict__ imass, int64_t total,
                                                         double dt, double gamma, double beta, uint64_t key,
                                                         uint64_t q_base, const double *__restrict__ noise) {
    TMD_EW_PROLOGUE
    const D2 f = ld2<VEC>(force, e, two);
    const D2 im = ld_imass<DIM>(imass, e, two);
    D2 x = ld2<VEC>(pos, e, two);
    D2 xi;
    if (NOISE == 1) {
        xi = ld2<VEC>(noise, e, two);
    } else {
        double z0, z1;
        if (NOISE == 2) {  // rgen.normal(loc=0, scale=sigma, size=vel.shape): element e takes normal e
            const D2 z = ld2<VEC>(noise, e, two);
            z0 = z.a;
            z1 = z.b;
        } else {
            two_normals(key, q_base + (uint64_t)e, two, z0, z1);
        }
        // sigma = sqrt(2*dt*imass/(beta*gamma))  (integrators.py:145-147)
        const double bg = __dmul_rn(beta, gamma), tdt = __dmul_rn(2.0, dt);
        xi.a = __dmul_rn(__dsqrt_rn(__ddiv_rn(__dmul_rn(tdt, im.a), bg)), z0);
        xi.b = __dmul_rn(__dsqrt_rn(__ddiv_rn(__dmul_rn(tdt, im.b), bg)), z1);
    }
    // pos += bddt*force + rands ; vel = rands ; bddt = dt*imass/gamma  (integrators.py:148,161-162)
    x.a = __dadd_rn(x.a, __dadd_rn(__dmul_rn(__ddiv_rn(__dmul_rn(dt, im.a), gamma), f.a), xi.a));
    x.b = __dadd_rn(x.b, __dadd_rn(__dmul_rn(__ddiv_rn(__dmul_rn(dt, im.b), gamma), f.b), xi.b));
    st2<VEC>(pos, e, two, x);
    st2<VEC>(vel, e, two, xi);
}

// ---- Langevin with inertia ----------------------------------------------------------------------
// POST: the closing velocity update of the previous step, v = v' + b2 F (integrators.py:281), is applied
// first -- the same F as this step's opening update uses -- so a run of steps needs one pass per step.
// NOISE as above; with NOISE == 2, noise_pos holds the standard normals (element e takes normals 2e and 2e + 1, the
// reshape of multivariate_normal, integrators.py:312-313) and noise_vel the reference's own Cholesky factors, (l00, l10,
// l11) per particle (np.linalg.cholesky, integrators.py:261).  norm @ cho.T is evaluated the way NumPy's matmul rounds it
// (dgemm accumulates the second product with a fused multiply-add): x-noise = l00 z0, v-noise = fma(l11, z1, l10 z0).
template <int DIM, bool VEC, int NOISE, bool POST = false>
__global__ void __launch_bounds__(256) inertia_pre_kernel(double *__restrict__ pos, double *__restrict__ vel,
                                                          const double *__restrict__ force,
                                                          const double *__restrict__ imass, int64_t total,
                                                          tmd_langevin_consts c, uint64_t key, uint64_t q_base,
                                                          const double *__restrict__ noise_pos,
                                                          const double *__restrict__ noise_vel,
                                                          const uint64_t *__restrict__ step_counter,
                                                          uint64_t step_stride) {
    TMD_EW_PROLOGUE
    if (step_counter) q_base += step_counter[0] * step_stride;  // graph replays: the position lives on the device
    const D2 f = ld2<VEC>(force, e, two);
    const D2 im = ld_imass<DIM>(imass, e, two);
    D2 x = ld2<VEC>(pos, e, two);
    D2 v = ld2<VEC>(vel, e, two);
    if (POST) {  // exactly inertia_post_kernel's arithmetic
        v.a = __dadd_rn(v.a, __dmul_rn(__dmul_rn(c.b2f, im.a), f.a));
        v.b = __dadd_rn(v.b, __dmul_rn(__dmul_rn(c.b2f, im.b), f.b));
    }
    const InertiaPerParticle pa = inertia_constants(c, im.a);
    const InertiaPerParticle pb = (im.b == im.a) ? pa : inertia_constants(c, im.b);
    D2 xr, vr;
    if (NOISE == 1) {
        xr = ld2<VEC>(noise_pos, e, two);
        vr = ld2<VEC>(noise_vel, e, two);
    } else if (NOISE == 2) {
        const D2 za = ld2<VEC>(noise_pos, 2 * e, true);
        const double *__restrict__ ca = noise_vel + 3 * (e / DIM);
        xr.a = __dmul_rn(ca[0], za.a);
        vr.a = __fma_rn(ca[2], za.b, __dmul_rn(ca[1], za.a));
        if (two) {
            const D2 zb = ld2<VEC>(noise_pos, 2 * e + 2, true);
            const double *__restrict__ cb = noise_vel + 3 * ((e + 1) / DIM);
            xr.b = __dmul_rn(cb[0], zb.a);
            vr.b = __fma_rn(cb[2], zb.b, __dmul_rn(cb[1], zb.a));
        } else {
            xr.b = vr.b = 0.0;
        }
    } else {
        // element e consumes normals q_base + 2e (position) and + 2e + 1 (velocity):
        // multivariate_normal reshapes its 2*dim draws to (dim, 2)  (integrators.py:312-313)
        double n0, n1;
        const uint64_t q = q_base + 2ull * (uint64_t)e;
        if (two && (q & 3ull) == 0) {  // both pairs of this thread are the halves of one Philox block
            double z[4];
            four_normals(key, q, z);
            xr.a = __dmul_rn(pa.l00, z[0]);
            vr.a = __dadd_rn(__dmul_rn(pa.l10, z[0]), __dmul_rn(pa.l11, z[1]));
            xr.b = __dmul_rn(pb.l00, z[2]);
            vr.b = __dadd_rn(__dmul_rn(pb.l10, z[2]), __dmul_rn(pb.l11, z[3]));
        } else {
        two_normals(key, q, true, n0, n1);
        xr.a = __dmul_rn(pa.l00, n0);  // norm @ cho.T  (integrators.py:315)
        vr.a = __dadd_rn(__dmul_rn(pa.l10, n0), __dmul_rn(pa.l11, n1));
        if (two) {
            two_normals(key, q_base + 2ull * (uint64_t)(e + 1), true, n0, n1);
            xr.b = __dmul_rn(pb.l00, n0);
            vr.b = __dadd_rn(__dmul_rn(pb.l10, n0), __dmul_rn(pb.l11, n1));
        } else {
            xr.b = vr.b = 0.0;
        }
        }
    }
    inertia_pre_update(x.a, v.a, f.a, xr.a, vr.a, c, pa);
    inertia_pre_update(x.b, v.b, f.b, xr.b, vr.b, c, pb);
    st2<VEC>(pos, e, two, x);
    st2<VEC>(vel, e, two, v);
}

template <int DIM, bool VEC>
__global__ void __launch_bounds__(256) inertia_post_kernel(double *__restrict__ vel,
                                                           const double *__restrict__ force,
                                                           const double *__restrict__ imass, int64_t total,
                                                           double b2f) {
    TMD_EW_PROLOGUE
    const D2 f = ld2<VEC>(force, e, two);
    const D2 im = ld_imass<DIM>(imass, e, two);
    D2 v = ld2<VEC>(vel, e, two);
    // vel = vel2 + b2*force  (integrators.py:281)
    v.a = __dadd_rn(v.a, __dmul_rn(__dmul_rn(b2f, im.a), f.a));
    v.b = __dadd_rn(v.b, __dmul_rn(__dmul_rn(b2f, im.b), f.b));
    st2<VEC>(vel, e, two, v);
}

// ---- fused replica kernel: nsteps of LangevinInertia over a DoubleWell-only system ---------------
template <int DIM>
__global__ void __launch_bounds__(256) inertia_doublewell_kernel(
    double *__restrict__ pos, double *__restrict__ vel, double *__restrict__ force,
    const double *__restrict__ imass, int64_t total, double a, double b, double cpar,
    tmd_langevin_consts c, uint64_t key, uint64_t q_base, uint64_t q_step, int nsteps,
    double *__restrict__ partial_v) {
    __shared__ double smem[8];
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    double vsum[1] = {0.0};
    if (e < total) {
        const double im = imass[e / DIM];
        const InertiaPerParticle p = inertia_constants(c, im);
        double x = pos[e], v = vel[e], f = force[e];
        uint64_t q = q_base + 2ull * (uint64_t)e;
        for (int s = 0; s < nsteps; ++s, q += q_step) {
            double n0, n1;
            two_normals(key, q, true, n0, n1);
            const double xr = __dmul_rn(p.l00, n0);
            const double vr = __dadd_rn(__dmul_rn(p.l10, n0), __dmul_rn(p.l11, n1));
            inertia_pre_update(x, v, f, xr, vr, c, p);
            f = dw_force(x, a, b, cpar);                 // system.force()      (integrators.py:280)
            v = __dadd_rn(v, __dmul_rn(p.b2, f));        // vel = vel2 + b2*F   (integrators.py:281)
        }
        pos[e] = x;
        vel[e] = v;
        force[e] = f;
        vsum[0] = dw_energy(x, a, b, cpar);              // system.potential()  (integrators.py:282)
    }
    block_sum<1, 256>(vsum, smem);
    if (threadIdx.x == 0) {
        double *row = partial_v + (size_t)blockIdx.x * 10;
        row[0] = vsum[0];
#pragma unroll
        for (int k = 1; k < 10; ++k) row[k] = 0.0;
    }
}

// Two coordinates per thread: when the stream position of every even coordinate is a multiple of four,
// the two Box-Muller pairs a thread needs per step are the two halves of ONE Philox block (the kernel is
// bound by the 64-bit multiplies of Philox: this halves them).  Same numbers as inertia_doublewell_kernel.
template <int DIM, bool VEC>
__global__ void __launch_bounds__(256) inertia_doublewell_pair_kernel(
    double *__restrict__ pos, double *__restrict__ vel, double *__restrict__ force,
    const double *__restrict__ imass, int64_t total, double a, double b, double cpar,
    tmd_langevin_consts c, uint64_t key, uint64_t q_base, uint64_t q_step, int nsteps,
    double *__restrict__ partial_v) {
    __shared__ double smem[8];
    const int64_t e = 2 * ((int64_t)blockIdx.x * blockDim.x + threadIdx.x);
    double vsum[1] = {0.0};
    if (e < total) {
        const bool two = (e + 1) < total;
        const D2 im = ld_imass<DIM>(imass, e, two);
        const InertiaPerParticle pa = inertia_constants(c, im.a);
        const InertiaPerParticle pb = (im.b == im.a) ? pa : inertia_constants(c, im.b);
        D2 x = ld2<VEC>(pos, e, two), v = ld2<VEC>(vel, e, two), f = ld2<VEC>(force, e, two);
        uint64_t q = q_base + 2ull * (uint64_t)e;  // a multiple of four (checked by the launcher)
        for (int s = 0; s < nsteps; ++s, q += q_step) {
            double z[4];
            four_normals(key, q, z);
            const double xra = __dmul_rn(pa.l00, z[0]);
            const double vra = __dadd_rn(__dmul_rn(pa.l10, z[0]), __dmul_rn(pa.l11, z[1]));
            const double xrb = __dmul_rn(pb.l00, z[2]);
            const double vrb = __dadd_rn(__dmul_rn(pb.l10, z[2]), __dmul_rn(pb.l11, z[3]));
            inertia_pre_update(x.a, v.a, f.a, xra, vra, c, pa);
            inertia_pre_update(x.b, v.b, f.b, xrb, vrb, c, pb);
            f.a = dw_force(x.a, a, b, cpar);                   // system.force()      (integrators.py:280)
            f.b = dw_force(x.b, a, b, cpar);
            v.a = __dadd_rn(v.a, __dmul_rn(pa.b2, f.a));       // vel = vel2 + b2*F   (integrators.py:281)
            v.b = __dadd_rn(v.b, __dmul_rn(pb.b2, f.b));
        }
        st2<VEC>(pos, e, two, x);
        st2<VEC>(vel, e, two, v);
        st2<VEC>(force, e, two, f);
        vsum[0] = dw_energy(x.a, a, b, cpar);                  // system.potential()  (integrators.py:282)
        if (two) vsum[0] += dw_energy(x.b, a, b, cpar);
    }
    block_sum<1, 256>(vsum, smem);
    if (threadIdx.x == 0) {
        double *row = partial_v + (size_t)blockIdx.x * 10;
        row[0] = vsum[0];
#pragma unroll
        for (int k = 1; k < 10; ++k) row[k] = 0.0;
    }
}

__global__ void __launch_bounds__(256) normal_fill_kernel(double *__restrict__ out, int64_t count, uint64_t key,
                                                          uint64_t offset) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < count) out[t] = normal_at(key, offset + (uint64_t)t);
}

// ---- kinetic tensor (system/particles.py:156-168) ------------------------------------------------
template <int DIM>
__global__ void __launch_bounds__(256) kinetic_kernel(const double *__restrict__ vel,
                                                      const double *__restrict__ mass, int64_t n,
                                                      double *__restrict__ partials) {
    __shared__ double smem[10 * 8];
    double acc[10];
#pragma unroll
    for (int k = 0; k < 10; ++k) acc[k] = 0.0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const double m = mass[i];
        double v[3] = {0.0, 0.0, 0.0};
#pragma unroll
        for (int k = 0; k < DIM; ++k) v[k] = vel[i * DIM + k];
#pragma unroll
        for (int p = 0; p < DIM; ++p)
#pragma unroll
            for (int q = 0; q < DIM; ++q) acc[1 + p * 3 + q] += (v[p] * m) * v[q];
    }
    block_sum<10, 256>(acc, smem);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < 10; ++k) partials[(size_t)blockIdx.x * 10 + k] = acc[k];
    }
}

// ---- velocity set-up (system/particles.py:140-153, 227-269) ---------------------------------------
template <int DIM>
__global__ void __launch_bounds__(256) maxwell_draw_kernel(double *__restrict__ vel, const double *__restrict__ imass,
                                                           int64_t total, uint64_t key, uint64_t q_base) {
    TMD_EW_PROLOGUE
    const D2 im = ld_imass<DIM>(imass, e, two);
    double z0, z1;
    two_normals(key, q_base + (uint64_t)e, two, z0, z1);
    vel[e] = __dmul_rn(__dsqrt_rn(im.a), z0);  // np.sqrt(imass) * rgen.normal(...)
    if (two) vel[e + 1] = __dmul_rn(__dsqrt_rn(im.b), z1);
}

template <int DIM>
__global__ void __launch_bounds__(256) momentum_kernel(const double *__restrict__ vel, const double *__restrict__ mass,
                                                       int64_t n, double *__restrict__ partials) {
    __shared__ double smem[10 * 8];
    double acc[10];
#pragma unroll
    for (int k = 0; k < 10; ++k) acc[k] = 0.0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const double m = mass[i];
        acc[0] += m;
#pragma unroll
        for (int k = 0; k < DIM; ++k) acc[1 + k] += __dmul_rn(vel[i * DIM + k], m);  // particles.vel * particles.mass
    }
    block_sum<10, 256>(acc, smem);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < 10; ++k) partials[(size_t)blockIdx.x * 10 + k] = acc[k];
    }
}

struct Shift3 {
    double s[3];
};

template <int DIM>
__global__ void __launch_bounds__(256) vel_shift_scale_kernel(double *__restrict__ vel, int64_t total, Shift3 shift,
                                                              int do_shift, int do_scale, double scale) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= total) return;
    double v = vel[e];
    if (do_shift) v = __dsub_rn(v, shift.s[e % DIM]);
    if (do_scale) v = __dmul_rn(v, scale);
    vel[e] = v;
}

}  // namespace tmd

using namespace tmd;

extern "C" {

int tmd_vv_kick_drift(double *pos, double *vel, const double *force, const double *imass, int64_t n,
                      int32_t dim, double dt, tmd_stream_t stream) {
    TMD_REQUIRE(dim >= 1 && dim <= 3 && n >= 0, "dim must be 1..3 and n >= 0");
    TMD_CHECK(require_device());
    return launch_vv<1, true>(pos, vel, force, imass, n, dim, dt, as_stream(stream));
}

int tmd_vv_kick(double *vel, const double *force, const double *imass, int64_t n, int32_t dim, double dt,
                tmd_stream_t stream) {
    TMD_REQUIRE(dim >= 1 && dim <= 3 && n >= 0, "dim must be 1..3 and n >= 0");
    TMD_CHECK(require_device());
    return launch_vv<1, false>(nullptr, vel, force, imass, n, dim, dt, as_stream(stream));
}

int tmd_vv_kick_kick_drift(double *pos, double *vel, const double *force, const double *imass, int64_t n,
                           int32_t dim, double dt, tmd_stream_t stream) {
    TMD_REQUIRE(dim >= 1 && dim <= 3 && n >= 0, "dim must be 1..3 and n >= 0");
    TMD_CHECK(require_device());
    return launch_vv<2, true>(pos, vel, force, imass, n, dim, dt, as_stream(stream));
}

int tmd_verlet_step(double *pos, double *vel, double *prev, const double *force, const double *imass,
                    int64_t n, int32_t dim, double dt, int32_t first_call, tmd_stream_t stream) {
    TMD_REQUIRE(dim >= 1 && dim <= 3 && n >= 0, "dim must be 1..3 and n >= 0");
    TMD_CHECK(require_device());
    const int64_t total = n * dim;
    if (total == 0) return TMD_OK;
    const bool vec = aligned16(pos) && aligned16(vel) && aligned16(prev) && aligned16(force);
    const Launch l = ew_launch(total);
    const double dt2 = dt * dt, half_idt = 0.5 / dt;  // integrators.py:49-50 (timestep**2 == dt*dt in IEEE)
    cudaStream_t s = as_stream(stream);
#define TMD_VERLET(D, V) \
    ::tmd::count_launch(), verlet_kernel<D, V><<<l.blocks, l.threads, 0, s>>>(pos, vel, prev, force, imass, total, dt, dt2, half_idt, first_call)
    if (dim == 1) { if (vec) TMD_VERLET(1, true); else TMD_VERLET(1, false); }
    else if (dim == 2) { if (vec) TMD_VERLET(2, true); else TMD_VERLET(2, false); }
    else { if (vec) TMD_VERLET(3, true); else TMD_VERLET(3, false); }
#undef TMD_VERLET
    TMD_CUDA(cudaGetLastError());
    return TMD_OK;
}

int tmd_langevin_overdamped(double *pos, double *vel, const double *force, const double *imass, int64_t n,
                            int32_t dim, double dt, double gamma, double beta, const tmd_rng *rng,
                            int64_t first, const double *noise, tmd_stream_t stream) {
    TMD_REQUIRE(dim >= 1 && dim <= 3 && n >= 0 && first >= 0, "dim must be 1..3, n >= 0, first >= 0");
    TMD_REQUIRE(noise != nullptr || rng != nullptr, "need either host-drawn noise or an rng");
    TMD_CHECK(require_device());
    const int64_t total = n * dim;
    if (total == 0) return TMD_OK;
    const bool vec = aligned16(pos) && aligned16(vel) && aligned16(force) && (!noise || aligned16(noise));
    const Launch l = ew_launch(total);
    const uint64_t key = rng ? rng->key : 0;
    const uint64_t q_base = rng ? rng->offset + (uint64_t)first * (uint64_t)dim : 0;
    cudaStream_t s = as_stream(stream);
#define TMD_OD(D, V, H) \
    ::tmd::count_launch(), overdamped_kernel<D, V, H><<<l.blocks, l.threads, 0, s>>>(pos, vel, force, imass, total, dt, gamma, beta, key, q_base, noise)
#define TMD_OD_D(D)                                                        \
    do {                                                                   \
        if (noise) { if (vec) TMD_OD(D, true, true); else TMD_OD(D, false, true); } \
        else { if (vec) TMD_OD(D, true, false); else TMD_OD(D, false, false); }     \
    } while (0)
    if (dim == 1) TMD_OD_D(1); else if (dim == 2) TMD_OD_D(2); else TMD_OD_D(3);
#undef TMD_OD_D
#undef TMD_OD
    TMD_CUDA(cudaGetLastError());
    return TMD_OK;
}

int tmd_langevin_inertia_pre(double *pos, double *vel, const double *force, const double *imass, int64_t n,
                             int32_t dim, const tmd_langevin_consts *consts, const tmd_rng *rng, int64_t first,
                             const double *noise_pos, const double *noise_vel, tmd_stream_t stream) {
    TMD_REQUIRE(dim >= 1 && dim <= 3 && n >= 0 && first >= 0, "dim must be 1..3, n >= 0, first >= 0");
    TMD_REQUIRE(consts != nullptr, "consts is NULL");
    TMD_REQUIRE((noise_pos == nullptr) == (noise_vel == nullptr), "give both noise arrays or neither");
    TMD_REQUIRE(noise_pos != nullptr || rng != nullptr, "need either host-drawn noise or an rng");
    TMD_CHECK(require_device());
    const int64_t total = n * dim;
    if (total == 0) return TMD_OK;
    const bool vec = aligned16(pos) && aligned16(vel) && aligned16(force) &&
                     (!noise_pos || (aligned16(noise_pos) && aligned16(noise_vel)));
    const Launch l = ew_launch(total);
    const uint64_t key = rng ? rng->key : 0;
    const uint64_t q_base = rng ? rng->offset + 2ull * (uint64_t)first * (uint64_t)dim : 0;
    cudaStream_t s = as_stream(stream);
#define TMD_IP(D, V, H) \
    ::tmd::count_launch(), inertia_pre_kernel<D, V, H><<<l.blocks, l.threads, 0, s>>>(pos, vel, force, imass, total, *consts, key, q_base, noise_pos, noise_vel, rng ? rng->step_counter : nullptr, rng ? rng->step_stride : 0)
#define TMD_IP_D(D)                                                            \
    do {                                                                       \
        if (noise_pos) { if (vec) TMD_IP(D, true, true); else TMD_IP(D, false, true); } \
        else { if (vec) TMD_IP(D, true, false); else TMD_IP(D, false, false); }         \
    } while (0)
    if (dim == 1) TMD_IP_D(1); else if (dim == 2) TMD_IP_D(2); else TMD_IP_D(3);
#undef TMD_IP_D
#undef TMD_IP
    TMD_CUDA(cudaGetLastError());
    return TMD_OK;
}

int tmd_langevin_inertia_post_pre(double *pos, double *vel, const double *force, const double *imass, int64_t n,
                                  int32_t dim, const tmd_langevin_consts *consts, const tmd_rng *rng, int64_t first,
                                  tmd_stream_t stream) {
    TMD_REQUIRE(dim >= 1 && dim <= 3 && n >= 0 && first >= 0, "dim must be 1..3, n >= 0, first >= 0");
    TMD_REQUIRE(consts != nullptr && rng != nullptr, "consts and rng are required");
    TMD_CHECK(require_device());
    const int64_t total = n * dim;
    if (total == 0) return TMD_OK;
    const bool vec = aligned16(pos) && aligned16(vel) && aligned16(force);
    const Launch l = ew_launch(total);
    const uint64_t q_base = rng->offset + 2ull * (uint64_t)first * (uint64_t)dim;
    cudaStream_t s = as_stream(stream);
#define TMD_IPP(D, V) \
    ::tmd::count_launch(), inertia_pre_kernel<D, V, false, true><<<l.blocks, l.threads, 0, s>>>(pos, vel, force, imass, total, *consts, rng->key, q_base, nullptr, nullptr, rng->step_counter, rng->step_stride)
    if (dim == 1) { if (vec) TMD_IPP(1, true); else TMD_IPP(1, false); }
    else if (dim == 2) { if (vec) TMD_IPP(2, true); else TMD_IPP(2, false); }
    else { if (vec) TMD_IPP(3, true); else TMD_IPP(3, false); }
#undef TMD_IPP
    TMD_CUDA(cudaGetLastError());
    return TMD_OK;
}

int tmd_langevin_overdamped_normals(double *pos, double *vel, const double *force, const double *imass, int64_t n,
                                    int32_t dim, double dt, double gamma, double beta, const double *normals,
                                    tmd_stream_t stream) {
    TMD_REQUIRE(dim >= 1 && dim <= 3 && n >= 0, "dim must be 1..3 and n >= 0");
    TMD_REQUIRE(normals != nullptr, "normals is NULL");
    TMD_CHECK(require_device());
    const int64_t total = n * dim;
    if (total == 0) return TMD_OK;
    const bool vec = aligned16(pos) && aligned16(vel) && aligned16(force) && aligned16(normals);
    const Launch l = ew_launch(total);
    cudaStream_t s = as_stream(stream);
#define TMD_ODN(D, V) \
    ::tmd::count_launch(), overdamped_kernel<D, V, 2><<<l.blocks, l.threads, 0, s>>>(pos, vel, force, imass, total, dt, gamma, beta, 0, 0, normals)
    if (dim == 1) { if (vec) TMD_ODN(1, true); else TMD_ODN(1, false); }
    else if (dim == 2) { if (vec) TMD_ODN(2, true); else TMD_ODN(2, false); }
    else { if (vec) TMD_ODN(3, true); else TMD_ODN(3, false); }
#undef TMD_ODN
    TMD_CUDA(cudaGetLastError());
    return TMD_OK;
}

int tmd_langevin_inertia_pre_normals(double *pos, double *vel, const double *force, const double *imass, int64_t n,
                                     int32_t dim, const tmd_langevin_consts *consts, const double *normals,
                                     const double *chol, int32_t post, tmd_stream_t stream) {
    TMD_REQUIRE(dim >= 1 && dim <= 3 && n >= 0, "dim must be 1..3 and n >= 0");
    TMD_REQUIRE(consts != nullptr && normals != nullptr && chol != nullptr, "consts / normals / chol is NULL");
    TMD_CHECK(require_device());
    const int64_t total = n * dim;
    if (total == 0) return TMD_OK;
    const bool vec = aligned16(pos) && aligned16(vel) && aligned16(force) && aligned16(normals);
    const Launch l = ew_launch(total);
    cudaStream_t s = as_stream(stream);
#define TMD_IPN(D, V, P) \
    ::tmd::count_launch(), inertia_pre_kernel<D, V, 2, P><<<l.blocks, l.threads, 0, s>>>(pos, vel, force, imass, total, *consts, 0, 0, normals, chol, nullptr, 0)
#define TMD_IPN_D(D)                                                            \
    do {                                                                        \
        if (post) { if (vec) TMD_IPN(D, true, true); else TMD_IPN(D, false, true); } \
        else { if (vec) TMD_IPN(D, true, false); else TMD_IPN(D, false, false); }    \
    } while (0)
    if (dim == 1) TMD_IPN_D(1); else if (dim == 2) TMD_IPN_D(2); else TMD_IPN_D(3);
#undef TMD_IPN_D
#undef TMD_IPN
    TMD_CUDA(cudaGetLastError());
    return TMD_OK;
}

int tmd_langevin_inertia_post(double *vel, const double *force, const double *imass, int64_t n, int32_t dim,
                              const tmd_langevin_consts *consts, tmd_stream_t stream) {
    TMD_REQUIRE(dim >= 1 && dim <= 3 && n >= 0, "dim must be 1..3 and n >= 0");
    TMD_REQUIRE(consts != nullptr, "consts is NULL");
    TMD_CHECK(require_device());
    const int64_t total = n * dim;
    if (total == 0) return TMD_OK;
    const bool vec = aligned16(vel) && aligned16(force);
    const Launch l = ew_launch(total);
    cudaStream_t s = as_stream(stream);
#define TMD_IPO(D, V) ::tmd::count_launch(), inertia_post_kernel<D, V><<<l.blocks, l.threads, 0, s>>>(vel, force, imass, total, consts->b2f)
    if (dim == 1) { if (vec) TMD_IPO(1, true); else TMD_IPO(1, false); }
    else if (dim == 2) { if (vec) TMD_IPO(2, true); else TMD_IPO(2, false); }
    else { if (vec) TMD_IPO(3, true); else TMD_IPO(3, false); }
#undef TMD_IPO
    TMD_CUDA(cudaGetLastError());
    return TMD_OK;
}

int tmd_langevin_inertia_doublewell(double *pos, double *vel, double *force, const double *imass, int64_t n,
                                    int32_t dim, double a, double b, double c,
                                    const tmd_langevin_consts *consts, const tmd_rng *rng, int64_t first,
                                    int64_t n_global, int32_t nsteps, double *vw, tmd_stream_t stream) {
    TMD_REQUIRE(dim >= 1 && dim <= 3 && n >= 0 && first >= 0 && nsteps >= 0, "bad sizes");
    TMD_REQUIRE(consts && rng && vw, "consts, rng and vw are required");
    TMD_REQUIRE(n_global >= first + n, "n_global must cover first + n");
    TMD_CHECK(require_device());
    const int64_t total = n * dim;
    cudaStream_t s = as_stream(stream);
    int blocks = (int)((total + 255) / 256);
    if (blocks == 0) {
        TMD_CUDA(cudaMemsetAsync(vw, 0, 10 * sizeof(double), s));
        return TMD_OK;
    }
    double *partials = nullptr;
    TMD_CHECK(arena_get(SLOT_PARTIAL_VW, (size_t)blocks * 10 * sizeof(double), (void **)&partials));
    const uint64_t q_base = rng->offset + 2ull * (uint64_t)first * (uint64_t)dim;
    const uint64_t q_step = 2ull * (uint64_t)n_global * (uint64_t)dim;
    if ((q_base & 3ull) == 0 && (q_step & 3ull) == 0) {
        // coordinate pairs share a Philox block on every step: two coordinates per thread
        const bool vec = aligned16(pos) && aligned16(vel) && aligned16(force);
        blocks = ew_launch(total).blocks;
#define TMD_IDW2(D, V) \
    ::tmd::count_launch(), inertia_doublewell_pair_kernel<D, V><<<blocks, 256, 0, s>>>(pos, vel, force, imass, total, a, b, c, *consts, rng->key, q_base, q_step, nsteps, partials)
        if (dim == 1) { if (vec) TMD_IDW2(1, true); else TMD_IDW2(1, false); }
        else if (dim == 2) { if (vec) TMD_IDW2(2, true); else TMD_IDW2(2, false); }
        else { if (vec) TMD_IDW2(3, true); else TMD_IDW2(3, false); }
#undef TMD_IDW2
        TMD_CUDA(cudaGetLastError());
        return launch_reduce_vw(partials, blocks, 1.0, TMD_WANT_V, vw, s);
    }
#define TMD_IDW(D) \
    ::tmd::count_launch(), inertia_doublewell_kernel<D><<<blocks, 256, 0, s>>>(pos, vel, force, imass, total, a, b, c, *consts, rng->key, q_base, q_step, nsteps, partials)
    if (dim == 1) TMD_IDW(1); else if (dim == 2) TMD_IDW(2); else TMD_IDW(3);
#undef TMD_IDW
    TMD_CUDA(cudaGetLastError());
    return launch_reduce_vw(partials, blocks, 1.0, TMD_WANT_V, vw, s);
}

int tmd_langevin_inertia_constants(double dt, double gamma, double beta, tmd_langevin_consts *out) {
    TMD_REQUIRE(out != nullptr, "out is NULL");
    // integrators.py:229-241, 249-256 (scalar parts)
    const double gdt = gamma * dt;
    const double e = exp(-gdt);
    const double c1 = (1.0 - e) / gdt;
    const double c2 = (1.0 - c1) / gdt;
    out->c0 = e;
    out->a1 = c1 * dt;
    out->a2f = c2 * (dt * dt);
    out->b1f = (c1 - c2) * dt;
    out->b2f = c2 * dt;
    out->dt = dt;
    out->beta = beta;
    out->beta_gamma = beta * gamma;
    out->kfac = 2.0 - (3.0 - 4.0 * e + e * e) / gdt;
    out->one_minus_e2 = 1.0 - e * e;
    out->one_minus_e_sq = (1.0 - e) * (1.0 - e);
    return TMD_OK;
}

int tmd_counter_add(uint64_t *counter, uint64_t increment, tmd_stream_t stream) {
    TMD_REQUIRE(counter != nullptr, "counter is NULL");
    TMD_CHECK(require_device());
    ::tmd::count_launch(), counter_add_kernel<<<1, 1, 0, as_stream(stream)>>>(counter, increment);
    TMD_CUDA(cudaGetLastError());
    return TMD_OK;
}

int tmd_philox_normal_fill(double *out, int64_t count, const tmd_rng *rng, tmd_stream_t stream) {
    TMD_REQUIRE(rng != nullptr && count >= 0, "rng is NULL or count < 0");
    TMD_CHECK(require_device());
    if (count == 0) return TMD_OK;
    const int blocks = (int)((count + 255) / 256);
    ::tmd::count_launch(), normal_fill_kernel<<<blocks, 256, 0, as_stream(stream)>>>(out, count, rng->key, rng->offset);
    TMD_CUDA(cudaGetLastError());
    return TMD_OK;
}

int tmd_host_philox_normal_fill(double *out, int64_t count, uint64_t key, uint64_t offset) {
    TMD_REQUIRE(out != nullptr || count == 0, "out is NULL");
    for (int64_t t = 0; t < count; ++t) out[t] = normal_at(key, offset + (uint64_t)t);
    return TMD_OK;
}

int tmd_host_philox_raw_fill(uint64_t *out, int64_t count, uint64_t key, uint64_t offset) {
    TMD_REQUIRE(out != nullptr || count == 0, "out is NULL");
    for (int64_t t = 0; t < count; ++t) {
        const uint64_t m = offset + (uint64_t)t;
        const Philox4 b = philox4x64_10(m >> 2, key);
        out[t] = b.w[m & 3ull];
    }
    return TMD_OK;
}

int tmd_maxwell_draw(double *vel, const double *imass, int64_t n, int32_t dim, const tmd_rng *rng, int64_t first,
                     tmd_stream_t stream) {
    TMD_REQUIRE(dim >= 1 && dim <= 3 && n >= 0 && first >= 0, "bad sizes");
    TMD_REQUIRE(rng != nullptr && (n == 0 || (vel && imass)), "vel, imass and rng are required");
    TMD_CHECK(require_device());
    const int64_t total = n * dim;
    if (total == 0) return TMD_OK;
    const Launch l = ew_launch(total);
    cudaStream_t s = as_stream(stream);
    const uint64_t q_base = rng->offset + (uint64_t)first * (uint64_t)dim;
#define TMD_MXW(D) ::tmd::count_launch(), maxwell_draw_kernel<D><<<l.blocks, l.threads, 0, s>>>(vel, imass, total, rng->key, q_base)
    if (dim == 1) TMD_MXW(1); else if (dim == 2) TMD_MXW(2); else TMD_MXW(3);
#undef TMD_MXW
    TMD_CUDA(cudaGetLastError());
    return TMD_OK;
}

int tmd_momentum(const double *vel, const double *mass, int64_t n, int32_t dim, double *out, tmd_stream_t stream) {
    TMD_REQUIRE(dim >= 1 && dim <= 3 && n >= 0 && out, "bad arguments");
    TMD_REQUIRE(n == 0 || (vel && mass), "vel and mass are required");
    TMD_CHECK(require_device());
    cudaStream_t s = as_stream(stream);
    int blocks = (int)((n + 255) / 256);
    const int cap = num_sms() * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    double *partials = nullptr, *vw = nullptr;
    TMD_CHECK(arena_get(SLOT_PARTIAL_VW, (size_t)blocks * 10 * sizeof(double), (void **)&partials));
    TMD_CHECK(arena_get(SLOT_MISC, 10 * sizeof(double), (void **)&vw));
    if (dim == 1) ::tmd::count_launch(), momentum_kernel<1><<<blocks, 256, 0, s>>>(vel, mass, n, partials);
    else if (dim == 2) ::tmd::count_launch(), momentum_kernel<2><<<blocks, 256, 0, s>>>(vel, mass, n, partials);
    else ::tmd::count_launch(), momentum_kernel<3><<<blocks, 256, 0, s>>>(vel, mass, n, partials);
    TMD_CUDA(cudaGetLastError());
    TMD_CHECK(launch_reduce_vw(partials, blocks, 1.0, TMD_WANT_V | TMD_WANT_W, vw, s));
    TMD_CUDA(cudaMemcpyAsync(out, vw, 4 * sizeof(double), cudaMemcpyDeviceToDevice, s));
    return TMD_OK;
}

int tmd_vel_shift_scale(double *vel, int64_t n, int32_t dim, const double *shift, int32_t do_scale, double scale,
                        tmd_stream_t stream) {
    TMD_REQUIRE(dim >= 1 && dim <= 3 && n >= 0, "bad sizes");
    TMD_REQUIRE(n == 0 || vel, "vel is NULL");
    TMD_CHECK(require_device());
    const int64_t total = n * dim;
    if (total == 0 || (!shift && !do_scale)) return TMD_OK;
    Shift3 sh = {{0.0, 0.0, 0.0}};
    for (int k = 0; k < dim && shift; ++k) sh.s[k] = shift[k];
    const int blocks = (int)((total + 255) / 256);
    cudaStream_t s = as_stream(stream);
#define TMD_VSS(D) ::tmd::count_launch(), vel_shift_scale_kernel<D><<<blocks, 256, 0, s>>>(vel, total, sh, shift != nullptr, do_scale, scale)
    if (dim == 1) TMD_VSS(1); else if (dim == 2) TMD_VSS(2); else TMD_VSS(3);
#undef TMD_VSS
    TMD_CUDA(cudaGetLastError());
    return TMD_OK;
}

int tmd_kinetic_tensor(const double *vel, const double *mass, int64_t n, int32_t dim, double *kin,
                       tmd_stream_t stream) {
    TMD_REQUIRE(dim >= 1 && dim <= 3 && n >= 0 && kin, "bad arguments");
    TMD_CHECK(require_device());
    cudaStream_t s = as_stream(stream);
    int blocks = (int)((n + 255) / 256);
    const int cap = num_sms() * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    double *partials = nullptr, *vw = nullptr;
    TMD_CHECK(arena_get(SLOT_PARTIAL_VW, (size_t)blocks * 10 * sizeof(double), (void **)&partials));
    TMD_CHECK(arena_get(SLOT_MISC, 10 * sizeof(double), (void **)&vw));
    if (dim == 1) ::tmd::count_launch(), kinetic_kernel<1><<<blocks, 256, 0, s>>>(vel, mass, n, partials);
    else if (dim == 2) ::tmd::count_launch(), kinetic_kernel<2><<<blocks, 256, 0, s>>>(vel, mass, n, partials);
    else ::tmd::count_launch(), kinetic_kernel<3><<<blocks, 256, 0, s>>>(vel, mass, n, partials);
    TMD_CUDA(cudaGetLastError());
    TMD_CHECK(launch_reduce_vw(partials, blocks, 0.5, TMD_WANT_W, vw, s));
    TMD_CUDA(cudaMemcpyAsync(kin, vw + 1, 9 * sizeof(double), cudaMemcpyDeviceToDevice, s));
    return TMD_OK;
}

}  // extern "C"
