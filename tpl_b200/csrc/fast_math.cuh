// Branch-free fp64 elementary functions for the model code.
//
// CUDA's libdevice sin/cos/tan, `/`, sqrt each contain a rarely-taken slow path behind a
// branch (huge arguments, denormals).  The branch ends the basic block, so ptxas cannot
// interleave two independent calls — in the rollout chain every transcendental then runs
// strictly after the previous one.  The versions below are the same algorithms' fast paths
// only (3-constant Cody-Waite reduction + fdlibm kernels, MUFU seed + Newton steps) as
// straight-line code, accurate to <= 2 ulp on their stated domains:
//
//   sincos_bf, sin_bf, cos_bf, tan_bf :  |x| <= 1e5   (heading / steering angles)
//   inv_bf, rsqrt_bf, sqrt_bf         :  normal, finite arguments; 0, inf and NaN give
//                                        inf/NaN results that the line search rejects the
//                                        same way the reference rejects them (optim.c:842)
//
// tests/test_gpu_parity.py::test_fast_math_accuracy checks them against libdevice.
#pragma once

namespace tplb {

__device__ __forceinline__ double rcp_seed(double a) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(a));
    return r;
}

__device__ __forceinline__ double rsqrt_seed(double a) {
    double r;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(a));
    return r;
}

// Newton steps after the hardware seed (rcp / rsqrt.approx.ftz.f64: relative error about 2^-23).
// Two quadratic steps reach 2^-92, i.e. the result is as good as the FMA roundings allow: measured
// on B200 over 2^20 arguments, 1/x is correctly rounded and rsqrt within 2 ulp with two steps
// exactly as with three (scripts/ulp_check.py), and the headline is 2.7 % faster.
#ifndef TPLB_NEWTON_STEPS
#define TPLB_NEWTON_STEPS 2
#endif

// 1/a
__device__ __forceinline__ double inv_bf(double a) {
    double r = rcp_seed(a);
#pragma unroll
    for (int i = 0; i < TPLB_NEWTON_STEPS; ++i) {
        const double e = fma(-a, r, 1.0);
        r = fma(r, e, r);
    }
    return r;
}

// 1/sqrt(a)
__device__ __forceinline__ double rsqrt_bf(double a) {
    double y = rsqrt_seed(a);
    const double h = 0.5 * a;
#pragma unroll
    for (int i = 0; i < TPLB_NEWTON_STEPS; ++i) {
        const double e = fma(-h * y, y, 0.5);       // 0.5 - 0.5 a y^2
        y = fma(y, e, y);
    }
    return y;
}

__device__ __forceinline__ double sqrt_bf(double a) {
    const double y = rsqrt_bf(a);
    double s = a * y;
    const double d = fma(-s, s, a);       // residual
    return fma(0.5 * y, d, s);
}

// a/b with one residual correction (not always correctly rounded; <= 1 ulp)
__device__ __forceinline__ double div_bf(double a, double b) {
    const double r = inv_bf(b);
    const double q = a * r;
    return fma(fma(-b, q, a), r, q);
}

// fdlibm __kernel_sin / __kernel_cos on |r| <= pi/4 (with the tail of the reduction)
__device__ __forceinline__ double ksin(double r, double z) {
    double p = fma(z, 1.58969099521155010221e-10, -2.50507602534068634195e-08);
    p = fma(p, z, 2.75573137070700676789e-06);
    p = fma(p, z, -1.98412698298579493134e-04);
    p = fma(p, z, 8.33333333332248946124e-03);
    p = fma(p, z, -1.66666666666666324348e-01);
    return fma(z * r, p, r);
}

__device__ __forceinline__ double kcos(double z) {
    double p = fma(z, -1.13596475577881948265e-11, 2.08757232129817482790e-09);
    p = fma(p, z, -2.75573143513906633035e-07);
    p = fma(p, z, 2.48015872894767294178e-05);
    p = fma(p, z, -1.38888888888741095749e-03);
    p = fma(p, z, 4.16666666666666019037e-02);
    return fma(z, fma(z, p, -0.5), 1.0);
}

// x = j*pi/2 + r, |r| <= pi/4; returns r and the quadrant
__device__ __forceinline__ double reduce_pio2(double x, int& quadrant) {
    const double magic = 6755399441055744.0;                 // 1.5 * 2^52: round to nearest integer
    const double t = fma(x, 6.36619772367581382433e-01, magic);
    quadrant = __double2loint(t);
    const double j = t - magic;
    double r = fma(-j, 1.57079632679489655800e+00, x);       // pi/2 split in three parts
    r = fma(-j, 6.12323399573676603587e-17, r);
    r = fma(-j, -1.49738490485916983479e-33, r);
    return r;
}

__device__ __forceinline__ void sincos_bf(double x, double* s, double* c) {
    int q;
    const double r = reduce_pio2(x, q);
    const double z = r * r;
    const double sr = ksin(r, z), cr = kcos(z);
    double ss = (q & 1) ? cr : sr;
    double cc = (q & 1) ? sr : cr;
    ss = (q & 2) ? -ss : ss;
    cc = ((q + 1) & 2) ? -cc : cc;
    *s = ss;
    *c = cc;
}

__device__ __forceinline__ double sin_bf(double x) {
    double s, c;
    sincos_bf(x, &s, &c);
    return s;
}

__device__ __forceinline__ double cos_bf(double x) {
    double s, c;
    sincos_bf(x, &s, &c);
    return c;
}

__device__ __forceinline__ double tan_bf(double x) {
    int q;
    const double r = reduce_pio2(x, q);
    const double z = r * r;
    const double sr = ksin(r, z), cr = kcos(z);
    // odd quadrants: tan(x) = -cos(r)/sin(r)
    const double num = (q & 1) ? -cr : sr;
    const double den = (q & 1) ? sr : cr;
    return div_bf(num, den);
}

}  // namespace tplb
