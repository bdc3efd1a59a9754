// fd_eval_common.cuh -- device helpers shared by the evaluation kernels (fd_eval.cu, fd_eval64.cu).
#pragma once

#include "fd_internal.h"

// ---- FP64 kernel functions from few DFMAs ---------------------------------------------------------------------------
// The FP64 pipe is the bound of the FP64 evaluation (64 lanes/clk/SM), so the kernel functions avoid libdevice's IEEE
// sqrt / log / exp (~20 / ~40 / ~35 FP64 instructions): 2^-40 relative accuracy is ample for a result that is rounded
// to FP32 -- the point of FP64 here is the cancellation in sum_j w_j phi_j, not the last bits of phi.

// sqrt: MUFU.RSQ64H seed (rsqrt.approx.f64, ~2^-20) + one Newton step in FP64 -> ~2^-40 relative, 4 FP64 instructions
// and no FP32 <-> FP64 conversions (F2F runs at a quarter of the FP64 rate).  x > 0.
__device__ __forceinline__ double fd_fast_sqrt64(double x)
{
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double t = x * y;
    const double e = fma(-t, y, 1.0);
    return fma(0.5 * t, e, t);
}

// 0.5 * ln|x| by a 128-entry table of (1 / c_i, 0.5 ln c_i), c_i = 1 + (i + 0.5) / 128, and a degree-4 series in
// d = m / c_i - 1, |d| <= 2^-8 (truncation d^5 / 5 < 2^-42); x = 0 gives a finite value (the caller multiplies by x).
__device__ __forceinline__ double fd_half_log64(double x, const double2* __restrict__ s_tab)
{
    const int hi = __double2hiint(x);
    const int e = ((hi >> 20) & 0x7ff) - 1023;
    const double m = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, __double2loint(x)); // [1, 2)
    const double2 t = s_tab[(hi >> 13) & 127];
    const double d = fma(m, t.x, -1.0);
    double p = fma(d, -0.125, 1.0 / 6.0);
    p = fma(d, p, -0.25);
    p = fma(d, p, 0.5);
    // (double)e without an I2F conversion: 2^52 + 2^31 + e as raw bits, minus the magic constant (exact)
    const double ed = __hiloint2double(0x43300000, e ^ 0x80000000) - 4503601774854144.0;
    return fma(ed, 0.34657359027997264, fma(d, p, t.y)); // e * ln2 / 2 + 0.5 ln c + 0.5 ln(1 + d)
}
__device__ __forceinline__ void fd_half_log64_table(double2* s_tab, int tid) // 128 threads fill the table
{
    const double c = 1.0 + ((double)tid + 0.5) / 128.0;
    s_tab[tid] = make_double2(1.0 / c, 0.5 * log(c));
}

// 2^t for t <= ~0 (the Gaussian's argument in base 2): t = k / 64 + r with |r| <= 1 / 128,
// 2^t = 2^(k >> 6) * T[k & 63] * (1 + r ln2 + ... degree 5), truncation (ln2 / 128)^6 / 720 < 2^-54.
// 11 FP64 instructions + 2 integer; s_tab[i] = 2^(i / 64).  Arguments below -1000 return ~2^-1000 (0 for every use).
__device__ __forceinline__ double fd_exp2_64(double t, const double* __restrict__ s_tab)
{
    t = fmax(t, -1000.0);
    const double magic = 6755399441055744.0;                // 1.5 * 2^52: the low word of (x + magic) is rint(x)
    const double kf = fma(t, 64.0, magic);
    const int k = __double2loint(kf);
    const double r = fma(kf - magic, -0.015625, t);          // t - k / 64, exact
    const double x = r * 0.69314718055994530942;
    double p = fma(x, 1.0 / 120.0, 1.0 / 24.0);
    p = fma(x, p, 1.0 / 6.0);
    p = fma(x, p, 0.5);
    p = fma(x, p, 1.0);
    p = fma(x, p, 1.0);
    const double s = s_tab[k & 63] * p;                      // in [1, 2 sqrt2): the exponent add below cannot carry wrongly
    return __hiloint2double(__double2hiint(s) + ((k >> 6) << 20), __double2loint(s));
}
__device__ __forceinline__ void fd_exp2_64_table(double* s_tab, int tid) // 64 threads fill the table
{
    s_tab[tid] = exp2((double)tid / 64.0);
}

__device__ __forceinline__ void fd_normalize3(float a[3])
{
    const float len = sqrtf(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]);
    if (len > 0.0f) {
        const float inv = 1.0f / len;
        a[0] *= inv;
        a[1] *= inv;
        a[2] *= inv;
    }
}

// SOP_FaceDeform.hpp:28-41, FP32, row-vector convention (see the oracle for the derivation)
__device__ __forceinline__ void fd_project_to_tangents(const float u[3], const float v[3], const float n[3], float d[3])
{
    float B[3][3];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) B[i][j] = u[i] * u[j] + v[i] * v[j] + n[i] * n[j];
    float a1[3], a2[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        a1[j] = u[0] * B[0][j] + u[1] * B[1][j] + u[2] * B[2][j];
        a2[j] = v[0] * B[0][j] + v[1] * B[1][j] + v[2] * B[2][j];
    }
    fd_normalize3(a1);
    fd_normalize3(a2);
    const float da1 = d[0] * a1[0] + d[1] * a1[1] + d[2] * a1[2];
    const float da2 = d[0] * a2[0] + d[1] * a2[1] + d[2] * a2[2];
#pragma unroll
    for (int k = 0; k < 3; ++k) d[k] = a1[k] * da1 + a2[k] * da2;
}

// one FP64 tensor-pipe instruction: C[8x8] += A[8x4] * B[4x8]; lane l holds A[l / 4][l % 4], B[l % 4][l / 4] and
// C[l / 4][2 (l % 4) + {0, 1}]
__device__ __forceinline__ void fd_dmma884(double& c0, double& c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}
