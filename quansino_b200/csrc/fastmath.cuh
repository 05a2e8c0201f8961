// Branch-free FP64 elementary functions for the hot kernels (sm_100a).
//
// The CUDA math library versions carry a special-case branch (BSSY / BSYNC / slow-path call) per call and
// re-materialise their coefficients; in the sweep kernel that overhead was 4x the FP64 work itself
// (profiles/r1_a_baseline.md).  These versions are straight-line DFMA chains with coefficients in the constant
// bank, valid on the argument ranges the kernels guarantee (stated per function), accurate to <= 1-2 ulp --
// far inside the 1e-10 parity budget.  tests/test_gpu_math.py checks each against numpy float64.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace qmc {

// exp(r) = 1 + r + r^2 q(r) on |r| <= ln(2)/2: q = degree-9 Chebyshev fit of (exp(r) - 1 - r) / r^2 (mpmath, 50 digits),
// maximum relative error of the whole expression 1.6e-17 with the coefficients rounded to double
__device__ __constant__ double FM_EXP[10] = {0.5000000000000001,    0.16666666666666669,   0.041666666666624164,  0.008333333333330065,
                                             0.0013888888917196736, 0.00019841269863040559, 2.480152132234228e-05, 2.7557268480289716e-06,
                                             2.7620075892548164e-07, 2.51003758422237e-08};
__device__ __constant__ double FM_LOG[7] = {6.666666666666735130e-01, 3.999999999940941908e-01, 2.857142874366239149e-01,
                                            2.222219843214978396e-01, 1.818357216161805012e-01, 1.531383769920937332e-01,
                                            1.479819860511658591e-01};
__device__ __constant__ double FM_SIN[6] = {-1.66666666666666324348e-01, 8.33333333332248946124e-03, -1.98412698298579493134e-04,
                                            2.75573137070700676789e-06,  -2.50507602534068634195e-08, 1.58969099521155010221e-10};
__device__ __constant__ double FM_COS[6] = {4.16666666666666019037e-02,  -1.38888888888741095749e-03, 2.48015872894767294178e-05,
                                            -2.75573143513906633035e-07, 2.08757232129817482790e-09,  -1.13596475577881948265e-11};

// log2(e), -ln2 (high part), -ln2 (low part): in the constant bank next to the coefficients, so that one 16-byte uniform load
// brings two of them (as 64-bit immediates each costs two instructions every time the compiler re-materialises it)
__device__ __constant__ double FM_RED[4] = {1.4426950408889634, -6.93147180369123816490e-01, -1.90821492927058770002e-10, 0.0};

constexpr double FM_MAGIC = 6755399441055744.0;  // 1.5 * 2^52
constexpr double FM_LN2_HI = 6.93147180369123816490e-01, FM_LN2_LO = 1.90821492927058770002e-10;

// 1 / d for normal d (|d| in [1e-300, 1e300]): MUFU.RCP64H seed (2^-20) + two Newton steps
__device__ __forceinline__ double rcp_fast(double d) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
    double e = fma(-d, y, 1.0);
    y = fma(y, e, y);
    e = fma(-d, y, 1.0);
    return fma(y, e, y);
}

// sqrt(x) for normal positive x: MUFU.RSQ64H seed + Goldschmidt, with a final residual correction
__device__ __forceinline__ double sqrt_fast(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double g = x * y, h = 0.5 * y;  // g ~ sqrt(x), h ~ 1 / (2 sqrt(x)), relative error ~2^-20
    const double r = fma(-g, h, 0.5);
    g = fma(g, r, g);  // ~2^-39
    h = fma(h, r, h);
    const double d = fma(-g, g, x);  // exact residual: the correction is one more Newton step (~2^-77 before rounding)
    return fma(d, h, g);
}

// exp(x) for x in [-700, 700]
__device__ __forceinline__ double exp_fast(double x) {
    const double t = fma(x, FM_RED[0], FM_MAGIC);
    const double k = t - FM_MAGIC;
    double r = fma(k, FM_RED[1], x);
    r = fma(k, FM_RED[2], r);
    double q = FM_EXP[9];
#pragma unroll
    for (int i = 8; i >= 0; --i) q = fma(q, r, FM_EXP[i]);
    const double p = fma(r, fma(r, q, 1.0), 1.0);
    const int n = __double2loint(t);
    return __hiloint2double(__double2hiint(p) + (n << 20), __double2loint(p));
}

// two exponentials in lockstep
__device__ __forceinline__ void exp2_fast(double x0, double x1, double &e0, double &e1) {
    const double l2e = FM_RED[0], nl2h = FM_RED[1], nl2l = FM_RED[2];
    const double t0 = fma(x0, l2e, FM_MAGIC), t1 = fma(x1, l2e, FM_MAGIC);
    const double k0 = t0 - FM_MAGIC, k1 = t1 - FM_MAGIC;
    double r0 = fma(k0, nl2h, x0), r1 = fma(k1, nl2h, x1);
    r0 = fma(k0, nl2l, r0);
    r1 = fma(k1, nl2l, r1);
    double q0 = FM_EXP[9], q1 = FM_EXP[9];
#pragma unroll
    for (int i = 8; i >= 0; --i) {
        const double c = FM_EXP[i];
        q0 = fma(q0, r0, c);
        q1 = fma(q1, r1, c);
    }
    const double p0 = fma(r0, fma(r0, q0, 1.0), 1.0), p1 = fma(r1, fma(r1, q1, 1.0), 1.0);
    e0 = __hiloint2double(__double2hiint(p0) + (__double2loint(t0) << 20), __double2loint(p0));
    e1 = __hiloint2double(__double2hiint(p1) + (__double2loint(t1) << 20), __double2loint(p1));
}

// three exponentials in lockstep: one coefficient fetch per Horner step serves all three chains
__device__ __forceinline__ void exp3_fast(double x0, double x1, double x2, double &e0, double &e1, double &e2) {
    const double l2e = FM_RED[0], nl2h = FM_RED[1], nl2l = FM_RED[2];
    const double t0 = fma(x0, l2e, FM_MAGIC), t1 = fma(x1, l2e, FM_MAGIC), t2 = fma(x2, l2e, FM_MAGIC);
    const double k0 = t0 - FM_MAGIC, k1 = t1 - FM_MAGIC, k2 = t2 - FM_MAGIC;
    double r0 = fma(k0, nl2h, x0), r1 = fma(k1, nl2h, x1), r2 = fma(k2, nl2h, x2);
    r0 = fma(k0, nl2l, r0);
    r1 = fma(k1, nl2l, r1);
    r2 = fma(k2, nl2l, r2);
    double q0 = FM_EXP[9], q1 = FM_EXP[9], q2 = FM_EXP[9];
#pragma unroll
    for (int i = 8; i >= 0; --i) {
        const double c = FM_EXP[i];
        q0 = fma(q0, r0, c);
        q1 = fma(q1, r1, c);
        q2 = fma(q2, r2, c);
    }
    const double p0 = fma(r0, fma(r0, q0, 1.0), 1.0), p1 = fma(r1, fma(r1, q1, 1.0), 1.0), p2 = fma(r2, fma(r2, q2, 1.0), 1.0);
    e0 = __hiloint2double(__double2hiint(p0) + (__double2loint(t0) << 20), __double2loint(p0));
    e1 = __hiloint2double(__double2hiint(p1) + (__double2loint(t1) << 20), __double2loint(p1));
    e2 = __hiloint2double(__double2hiint(p2) + (__double2loint(t2) << 20), __double2loint(p2));
}

// log(x) for normal positive x (fdlibm e_log.c scheme)
__device__ __forceinline__ double log_fast(double x) {
    int hx = __double2hiint(x);
    const int lx = __double2loint(x);
    int k = (hx >> 20) - 1023;
    hx &= 0x000fffff;
    const int i = (hx + 0x95f64) & 0x100000;
    k += i >> 20;
    const double m = __hiloint2double(hx | (i ^ 0x3ff00000), lx);  // in [sqrt(2)/2, sqrt(2))
    const double f = m - 1.0;
    const double s = f * rcp_fast(2.0 + f);
    const double z = s * s;
    double R = FM_LOG[6];
#pragma unroll
    for (int j = 5; j >= 0; --j) R = fma(R, z, FM_LOG[j]);
    R *= z;
    const double hfsq = 0.5 * f * f;
    const double dk = (double)k;
    // dk * ln2_hi - ((hfsq - (s * (hfsq + R) + dk * ln2_lo)) - f), with the two constants taken (negated, exactly) from FM_RED
    return -(dk * FM_RED[1]) - ((hfsq - (s * (hfsq + R) - dk * FM_RED[2])) - f);
}

// sin and cos of phi in [0, 2 pi]  (fdlibm k_sin / k_cos kernels after a two-term pi/2 reduction)
__device__ __forceinline__ void sincos_2pi(double phi, double &sn, double &cs) {
    const double t = fma(phi, 6.36619772367581382433e-01, FM_MAGIC);
    const double q = t - FM_MAGIC;
    double r = fma(q, -1.57079632673412561417e+00, phi);
    r = fma(q, -6.07710050650619224932e-11, r);
    const double z = r * r;
    double ps = FM_SIN[5], pc = FM_COS[5];
#pragma unroll
    for (int j = 4; j >= 0; --j) {
        ps = fma(ps, z, FM_SIN[j]);
        pc = fma(pc, z, FM_COS[j]);
    }
    const double s = fma(r * z, ps, r);
    const double c = fma(z * z, pc, fma(z, -0.5, 1.0));
    const int n = __double2loint(t) & 3;
    const double a = (n & 1) ? c : s, b = (n & 1) ? s : c;
    sn = (n & 2) ? -a : a;
    cs = ((n + 1) & 2) ? -b : b;
}

}  // namespace qmc
