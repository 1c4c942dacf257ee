// DoubleWell element math (turtlemd/potentials/well.py:65-81), shared by wells.cu and the fused
// replica kernel in integrators.cu.
#pragma once

namespace tmd {

// x^3 and x^4 with the rounding error of x*x carried along: within 1 ulp of a correctly rounded
// power, which is what NumPy's pos**3 / pos**4 (well.py:68,77) deliver.
__device__ __forceinline__ void cube_quart(double x, double &x3, double &x4) {
    const double x2 = __dmul_rn(x, x);
    const double e2 = __fma_rn(x, x, -x2);
    const double h3 = __dmul_rn(x2, x);
    const double e3 = __fma_rn(x2, x, -h3);
    x3 = __dadd_rn(h3, __fma_rn(e2, x, e3));
    const double h4 = __dmul_rn(x2, x2);
    const double e4 = __fma_rn(x2, x2, -h4);
    x4 = __dadd_rn(h4, __fma_rn(__dmul_rn(2.0, x2), e2, e4));
}

__device__ __forceinline__ double dw_force(double x, double a, double b, double c) {
    double x3, x4;
    cube_quart(x, x3, x4);
    // -4.0*(a*pos**3) + 2.0*(b*(pos - c))  (well.py:77-79)
    return __dadd_rn(__dmul_rn(-4.0, __dmul_rn(a, x3)), __dmul_rn(2.0, __dmul_rn(b, __dsub_rn(x, c))));
}

__device__ __forceinline__ double dw_energy(double x, double a, double b, double c) {
    double x3, x4;
    cube_quart(x, x3, x4);
    const double xc = __dsub_rn(x, c);
    // a*pos**4 - b*(pos - c)**2  (well.py:68-71)
    return __dsub_rn(__dmul_rn(a, x4), __dmul_rn(b, __dmul_rn(xc, xc)));
}

}  // namespace tmd
