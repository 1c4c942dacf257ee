// tmd_host_*: the plug-in boundary with HOST buffers laid out like the reference's NumPy arrays.
// Copy in, launch the same kernels the resident path uses, copy out, synchronise.  Staging buffers
// come from the per-device arena, so repeated calls do not allocate.
#include <vector>

#include "common.cuh"

namespace tmd {

static int upload(ArenaSlot slot, const void *host, size_t bytes, void **dev, cudaStream_t s) {
    TMD_CHECK(arena_get(slot, bytes ? bytes : 1, dev));
    if (bytes) TMD_CUDA(cudaMemcpyAsync(*dev, host, bytes, cudaMemcpyHostToDevice, s));
    return TMD_OK;
}

// int64 (NumPy ptype, already dense) -> int32 on the device
static int upload_types(const int64_t *tidx, int64_t n, int ntypes, int32_t **dev, cudaStream_t s) {
    *dev = nullptr;
    if (ntypes <= 1 || tidx == nullptr) return TMD_OK;
    std::vector<int32_t> tmp((size_t)n);
    for (int64_t i = 0; i < n; ++i) {
        if (tidx[i] < 0 || tidx[i] >= ntypes) {
            set_error("type index %lld of atom %lld is outside 0..%d", (long long)tidx[i], (long long)i, ntypes - 1);
            return TMD_ERR_INVALID;
        }
        tmp[(size_t)i] = (int32_t)tidx[i];
    }
    TMD_CHECK(arena_get(SLOT_HOST_B, (size_t)n * sizeof(int32_t) + 4, (void **)dev));
    // pageable source: the copy is staged before the call returns, tmp may go out of scope
    TMD_CUDA(cudaMemcpyAsync(*dev, tmp.data(), (size_t)n * sizeof(int32_t), cudaMemcpyHostToDevice, s));
    TMD_CUDA(cudaStreamSynchronize(s));
    return TMD_OK;
}

static int pick_lj(const double *pos, const int32_t *tidx, int64_t n, const tmd_box *box, const double *table,
                   int ntypes, uint32_t flags, int method, double *force, double *vw, cudaStream_t s) {
    int usable = 0;
    if (method != 1) TMD_CHECK(tmd_lj_cells_usable(n, box, table, ntypes, &usable));
    if (method == 2 && !usable) {
        set_error("cell list requested but the box holds fewer than 3 cells in a periodic direction");
        return TMD_ERR_UNSUPPORTED;
    }
    const bool cells = method == 2 || (method == 0 && usable && n >= 4096);
    if (cells) return tmd_lj_cells(pos, tidx, n, box, table, ntypes, flags, 0, n, force, vw, s);
    return tmd_lj_allpairs(pos, tidx, n, box, table, ntypes, flags, 0, n, force, vw, s);
}

static int download_results(uint32_t flags, size_t fbytes, const double *d_force, const double *d_vw,
                            double *force, double *vw, cudaStream_t s) {
    if ((flags & TMD_WANT_F) && force && fbytes)
        TMD_CUDA(cudaMemcpyAsync(force, d_force, fbytes, cudaMemcpyDeviceToHost, s));
    TMD_CUDA(cudaMemcpyAsync(vw, d_vw, 10 * sizeof(double), cudaMemcpyDeviceToHost, s));
    TMD_CUDA(cudaStreamSynchronize(s));
    return TMD_OK;
}

}  // namespace tmd

using namespace tmd;

extern "C" {

int tmd_host_lj(const double *pos, const int64_t *tidx, int64_t n, const tmd_box *box, const double *table,
                int32_t ntypes, uint32_t flags, int32_t method, double *force, double *vw) {
    TMD_REQUIRE(box && box->dim >= 1 && box->dim <= 3 && n >= 0, "bad box or n");
    TMD_REQUIRE(vw != nullptr, "vw is NULL");
    TMD_REQUIRE(!(flags & TMD_ACCUMULATE), "TMD_ACCUMULATE is meaningless with host buffers");
    TMD_CHECK(require_device());
    cudaStream_t s = 0;
    const size_t fbytes = (size_t)n * box->dim * sizeof(double);
    double *d_pos = nullptr, *d_force = nullptr, *d_vw = nullptr;
    int32_t *d_t = nullptr;
    TMD_CHECK(upload(SLOT_HOST_A, pos, fbytes, (void **)&d_pos, s));
    TMD_CHECK(upload_types(tidx, n, ntypes, &d_t, s));
    TMD_CHECK(arena_get(SLOT_HOST_C, fbytes + 8, (void **)&d_force));
    TMD_CHECK(arena_get(SLOT_HOST_D, 10 * sizeof(double), (void **)&d_vw));
    TMD_CHECK(pick_lj(d_pos, d_t, n, box, table, ntypes, flags, method, d_force, d_vw, s));
    return download_results(flags, fbytes, d_force, d_vw, force, vw, s);
}

int tmd_host_lj_triclinic(const double *pos, const int64_t *tidx, int64_t n, const tmd_cell *cell, const double *table,
                          int32_t ntypes, uint32_t flags, double *force, double *vw) {
    TMD_REQUIRE(cell && cell->dim >= 2 && cell->dim <= 3 && n >= 0, "bad cell or n");
    TMD_REQUIRE(vw != nullptr, "vw is NULL");
    TMD_REQUIRE(!(flags & TMD_ACCUMULATE), "TMD_ACCUMULATE is meaningless with host buffers");
    TMD_CHECK(require_device());
    cudaStream_t s = 0;
    const size_t fbytes = (size_t)n * cell->dim * sizeof(double);
    double *d_pos = nullptr, *d_force = nullptr, *d_vw = nullptr;
    int32_t *d_t = nullptr;
    TMD_CHECK(upload(SLOT_HOST_A, pos, fbytes, (void **)&d_pos, s));
    TMD_CHECK(upload_types(tidx, n, ntypes, &d_t, s));
    TMD_CHECK(arena_get(SLOT_HOST_C, fbytes + 8, (void **)&d_force));
    TMD_CHECK(arena_get(SLOT_HOST_D, 10 * sizeof(double), (void **)&d_vw));
    TMD_CHECK(tmd_lj_triclinic(d_pos, d_t, n, cell, table, ntypes, flags, 0, n, d_force, d_vw, s));
    return download_results(flags, fbytes, d_force, d_vw, force, vw, s);
}

int tmd_host_doublewell(const double *pos, int64_t n, int32_t dim, double a, double b, double c, uint32_t flags,
                        double *force, double *vw) {
    TMD_REQUIRE(dim >= 1 && dim <= 3 && n >= 0, "bad dim or n");
    TMD_REQUIRE(vw != nullptr, "vw is NULL");
    TMD_REQUIRE(!(flags & TMD_ACCUMULATE), "TMD_ACCUMULATE is meaningless with host buffers");
    TMD_CHECK(require_device());
    cudaStream_t s = 0;
    const size_t fbytes = (size_t)n * dim * sizeof(double);
    double *d_pos = nullptr, *d_force = nullptr, *d_vw = nullptr;
    TMD_CHECK(upload(SLOT_HOST_A, pos, fbytes, (void **)&d_pos, s));
    TMD_CHECK(arena_get(SLOT_HOST_C, fbytes + 8, (void **)&d_force));
    TMD_CHECK(arena_get(SLOT_HOST_D, 10 * sizeof(double), (void **)&d_vw));
    TMD_CHECK(tmd_doublewell(d_pos, n, dim, a, b, c, flags, d_force, d_vw, s));
    return download_results(flags, fbytes, d_force, d_vw, force, vw, s);
}

static int host_dwpair(const double *pos, const int64_t *ptype, int64_t n, int dim, const tmd_box *box,
                       const tmd_cell *cell, int64_t type0, int64_t type1, double rwidth, double width2,
                       double height, uint32_t flags, double *force, double *vw);

int tmd_host_dwpair(const double *pos, const int64_t *ptype, int64_t n, const tmd_box *box, int64_t type0,
                    int64_t type1, double rwidth, double width2, double height, uint32_t flags, double *force,
                    double *vw) {
    TMD_REQUIRE(box && box->dim >= 1 && box->dim <= 3 && n >= 0, "bad box or n");
    return host_dwpair(pos, ptype, n, box->dim, box, nullptr, type0, type1, rwidth, width2, height, flags, force, vw);
}

int tmd_host_dwpair_triclinic(const double *pos, const int64_t *ptype, int64_t n, const tmd_cell *cell,
                              int64_t type0, int64_t type1, double rwidth, double width2, double height,
                              uint32_t flags, double *force, double *vw) {
    TMD_REQUIRE(cell && cell->dim >= 2 && cell->dim <= 3 && n >= 0, "bad cell or n");
    return host_dwpair(pos, ptype, n, cell->dim, nullptr, cell, type0, type1, rwidth, width2, height, flags, force, vw);
}

static int host_dwpair(const double *pos, const int64_t *ptype, int64_t n, int dim, const tmd_box *box,
                       const tmd_cell *cell, int64_t type0, int64_t type1, double rwidth, double width2,
                       double height, uint32_t flags, double *force, double *vw) {
    TMD_REQUIRE(ptype != nullptr || n == 0, "ptype is NULL");
    TMD_REQUIRE(vw != nullptr, "vw is NULL");
    TMD_REQUIRE(!(flags & TMD_ACCUMULATE), "TMD_ACCUMULATE is meaningless with host buffers");
    TMD_CHECK(require_device());
    cudaStream_t s = 0;
    std::vector<int32_t> idx0, idx1;
    for (int64_t i = 0; i < n; ++i) {
        if (ptype[i] == type0) idx0.push_back((int32_t)i);
        else if (ptype[i] == type1) idx1.push_back((int32_t)i);
    }
    const int same = type0 == type1;
    const size_t fbytes = (size_t)n * dim * sizeof(double);
    double *d_pos = nullptr, *d_force = nullptr, *d_vw = nullptr;
    int32_t *d_i0 = nullptr, *d_i1 = nullptr;
    TMD_CHECK(upload(SLOT_HOST_A, pos, fbytes, (void **)&d_pos, s));
    TMD_CHECK(upload(SLOT_HOST_B, idx0.data(), idx0.size() * sizeof(int32_t), (void **)&d_i0, s));
    TMD_CHECK(upload(SLOT_HOST_E, idx1.data(), idx1.size() * sizeof(int32_t), (void **)&d_i1, s));
    TMD_CUDA(cudaStreamSynchronize(s));  // idx vectors are pageable temporaries
    TMD_CHECK(arena_get(SLOT_HOST_C, fbytes + 8, (void **)&d_force));
    TMD_CHECK(arena_get(SLOT_HOST_D, 10 * sizeof(double), (void **)&d_vw));
    if (cell)
        TMD_CHECK(tmd_dwpair_triclinic(d_pos, n, cell, d_i0, (int32_t)idx0.size(), d_i1, (int32_t)idx1.size(), same,
                                       rwidth, width2, height, flags, d_force, d_vw, s));
    else
        TMD_CHECK(tmd_dwpair(d_pos, n, box, d_i0, (int32_t)idx0.size(), d_i1, (int32_t)idx1.size(), same, rwidth,
                             width2, height, flags, d_force, d_vw, s));
    return download_results(flags, fbytes, d_force, d_vw, force, vw, s);
}

int tmd_host_vv_lj_step(double *pos, double *vel, double *force, const double *imass, const int64_t *tidx,
                        int64_t n, const tmd_box *box, const double *table, int32_t ntypes, int32_t method,
                        double dt, double *vw) {
    TMD_REQUIRE(box && box->dim >= 1 && box->dim <= 3 && n >= 0, "bad box or n");
    TMD_REQUIRE(pos && vel && force && imass && vw, "NULL array");
    TMD_CHECK(require_device());
    cudaStream_t s = 0;
    const int dim = box->dim;
    const size_t fbytes = (size_t)n * dim * sizeof(double);
    double *d_pos = nullptr, *d_vel = nullptr, *d_force = nullptr, *d_im = nullptr, *d_vw = nullptr;
    int32_t *d_t = nullptr;
    TMD_CHECK(upload(SLOT_HOST_A, pos, fbytes, (void **)&d_pos, s));
    TMD_CHECK(upload(SLOT_HOST_E, vel, fbytes, (void **)&d_vel, s));
    TMD_CHECK(upload(SLOT_HOST_C, force, fbytes, (void **)&d_force, s));
    TMD_CHECK(upload(SLOT_HOST_F, imass, (size_t)n * sizeof(double), (void **)&d_im, s));
    TMD_CHECK(upload_types(tidx, n, ntypes, &d_t, s));
    TMD_CHECK(arena_get(SLOT_HOST_D, 10 * sizeof(double), (void **)&d_vw));
    // VelocityVerlet.integration_step (integrators.py:84-94)
    TMD_CHECK(tmd_vv_kick_drift(d_pos, d_vel, d_force, d_im, n, dim, dt, s));
    TMD_CHECK(pick_lj(d_pos, d_t, n, box, table, ntypes, TMD_WANT_V | TMD_WANT_F | TMD_WANT_W, method, d_force,
                      d_vw, s));
    TMD_CHECK(tmd_vv_kick(d_vel, d_force, d_im, n, dim, dt, s));
    if (fbytes) {
        TMD_CUDA(cudaMemcpyAsync(pos, d_pos, fbytes, cudaMemcpyDeviceToHost, s));
        TMD_CUDA(cudaMemcpyAsync(vel, d_vel, fbytes, cudaMemcpyDeviceToHost, s));
    }
    return download_results(TMD_WANT_F, fbytes, d_force, d_vw, force, vw, s);
}

}  // extern "C"
