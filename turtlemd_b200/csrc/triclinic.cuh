// TriclinicBox.pbc_dist (turtlemd/system/box.py:271-296) for one distance vector, shared by the pair kernels.
#pragma once

#include "common.cuh"

namespace tmd {

struct TriCell {
    double m[9], inv[9];   // box_matrix and box_matrix_inv, row-major 3x3 (zero padded for dim = 2)
    double half;           // shortest_width_half (box.py:257-258)
    double t[27][3];       // translations: box_matrix @ index for index in product([-1, 0, 1], repeat=dim) (:251-256)
    int nimg;              // 3^dim
};

// In place: d <- nearest image of d.  Individually rounded products and left-to-right sums, the order of
// NumPy's small matrix-vector products; `rint` = np.rint (half to even).
template <int DIM>
__device__ __forceinline__ void tri_min_image(const TriCell &c, double (&d)[3]) {
    double s[3] = {0.0, 0.0, 0.0};
#pragma unroll
    for (int r = 0; r < DIM; ++r) {  // np.rint(box_matrix_inv @ distance)  (:286-288)
        double f = __dmul_rn(c.inv[r * 3], d[0]);
#pragma unroll
        for (int k = 1; k < DIM; ++k) f = __dadd_rn(f, __dmul_rn(c.inv[r * 3 + k], d[k]));
        s[r] = rint(f);
    }
    double dist[3] = {0.0, 0.0, 0.0};
#pragma unroll
    for (int r = 0; r < DIM; ++r) {  // distance - box_matrix @ (...)
        double f = __dmul_rn(c.m[r * 3], s[0]);
#pragma unroll
        for (int k = 1; k < DIM; ++k) f = __dadd_rn(f, __dmul_rn(c.m[r * 3 + k], s[k]));
        dist[r] = __dsub_rn(d[r], f);
    }
    double best = __dmul_rn(dist[0], dist[0]);  // np.dot(dist, dist)  (:289)
#pragma unroll
    for (int k = 1; k < DIM; ++k) best = __dadd_rn(best, __dmul_rn(dist[k], dist[k]));
    if (!(__dsqrt_rn(best) < c.half)) {  // :290-291
        int idx = 0;
        double shortest = 0.0;
        for (int g = 0; g < c.nimg; ++g) {  // images = dist + translations; lengths; argmin = first minimum  (:292-294)
            double im = __dadd_rn(dist[0], c.t[g][0]);
            double len = __dmul_rn(im, im);
#pragma unroll
            for (int k = 1; k < DIM; ++k) {
                im = __dadd_rn(dist[k], c.t[g][k]);
                len = __dadd_rn(len, __dmul_rn(im, im));
            }
            if (g == 0 || len < shortest) {
                shortest = len;
                idx = g;
            }
        }
        if (shortest < best) {  // :295-296
#pragma unroll
            for (int k = 0; k < DIM; ++k) dist[k] = __dadd_rn(dist[k], c.t[idx][k]);
        }
    }
#pragma unroll
    for (int k = 0; k < DIM; ++k) d[k] = dist[k];
}

int load_tri_cell(const tmd_cell *cell, TriCell &out);

}  // namespace tmd
