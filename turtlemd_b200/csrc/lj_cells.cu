// K2/K3 placeholder (filled in below in this round): cell-list Lennard-Jones.
#include "lj_common.cuh"

using namespace tmd;

extern "C" int tmd_lj_cells_usable(int64_t n, const tmd_box *box, const double *table, int32_t ntypes, int *usable) {
    TMD_REQUIRE(usable != nullptr, "usable is NULL");
    (void)n; (void)box; (void)table; (void)ntypes;
    *usable = 0;
    return TMD_OK;
}

extern "C" int tmd_lj_cells(const double *, const int32_t *, int64_t, const tmd_box *, const double *, int32_t,
                            uint32_t, int64_t, int64_t, double *, double *, tmd_stream_t) {
    set_error("tmd_lj_cells: not built yet");
    return TMD_ERR_UNSUPPORTED;
}
