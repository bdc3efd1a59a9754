// fd_factor.cu -- K2: blocked right-looking LU with partial pivoting, FP64, column-major, in place.
//
// Replaces the solver inside alglib::rbfbuildmodel (reference SOP_FaceDeform.cpp:363) by the dense direct
// factorisation north_star names.  The saddle-point system is symmetric indefinite (and non-symmetric with
// per-centre QNN radii), so the general path is pivoted LU:
//   for each block column k0 (width NB = 32), two launches:
//     k_lu_panel_cluster : unblocked LU of the m x 32 panel held in the registers of a thread-block cluster
//     k_lu_update        : interchanges on every column, U12 = L11^-1 A12 and A22 -= L21 * U12, fused per 16 columns
//   k_lu_perm            : folds the interchanges into one permutation vector for the right-hand-side gather
#include <cooperative_groups.h>
#include <stdlib.h>

#include "fd_internal.h"

// The kernels live in fd_factor_impl.inl and are instantiated twice: FP64 (the default) and FP32 (factor for the
// FP32 + iterative-refinement mode, fd_params.factor_precision = FD_FACTOR_FP32_IR).

__device__ __forceinline__ unsigned fd_mag_key(double v) { return (unsigned)__double2hiint(v) & 0x7fffffffu; }
__device__ __forceinline__ unsigned fd_mag_key(float v) { return __float_as_uint(v) & 0x7fffffffu; }
__device__ __forceinline__ double fd_fast_rcp(double d)
{
    // MUFU.RCP64H seed + two Newton steps (full double accuracy up to the last ulps; LU does not need a correctly
    // rounded quotient), 5 dependent instructions instead of the ~20 of an IEEE division
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
    double e = fma(-d, r, 1.0);
    r = fma(r, e, r);
    e = fma(-d, r, 1.0);
    return fma(r, e, r);
}
__device__ __forceinline__ float fd_fast_rcp(float d) { return __frcp_rn(d); }
__device__ __forceinline__ double2 fd_make2(double a, double b) { return make_double2(a, b); }
__device__ __forceinline__ float2 fd_make2(float a, float b) { return make_float2(a, b); }

#define REAL double
#define REAL2 double2
#define FD_LU_NS lu_f64
#define FD_LU_DMMA 1 // the tile product of the fused LU on the FP64 tensor pipe (mma.sync.m8n8k4.f64)
#include "fd_factor_impl.inl"
#undef FD_LU_DMMA
#undef REAL
#undef REAL2
#undef FD_LU_NS

#define REAL float
#define REAL2 float2
#define FD_LU_NS lu_f32
#include "fd_factor_impl.inl"
#undef REAL
#undef REAL2
#undef FD_LU_NS

cudaError_t fd_factor_setup(fd_ctx* ctx)
{
    (void)ctx; // the current device is the ctx's (fd_ctx_create)
    cudaError_t e = lu_f64::setup_attributes();
    if (e == cudaSuccess) e = lu_f32::setup_attributes();
    return e;
}

cudaError_t fd_launch_lu(fd_ctx* ctx, double* d_A, int lda, int n, int* d_ipiv, int* d_perm, int* d_flags,
                         double* d_pivstat, int* d_win)
{
    return lu_f64::launch_lu(ctx, d_A, lda, n, d_ipiv, d_perm, d_flags, d_pivstat, d_win);
}
cudaError_t fd_launch_lu_nopivot(fd_ctx* ctx, double* d_A, int lda, int n, int* d_ipiv, int* d_perm, int* d_flags,
                                 double* d_pivstat)
{
    return lu_f64::launch_lu_nopivot(ctx, d_A, lda, n, d_ipiv, d_perm, d_flags, d_pivstat);
}
// sym != 0: the matrix is symmetric (K + lambda I with its polynomial border, or the reduced null-space block)
cudaError_t fd_launch_lu_nopivot_fused(fd_ctx* ctx, double* d_A, int lda, int n, int* d_ipiv, int* d_perm, int* d_flags,
                                       double* d_pivstat, double* d_Tinv, int sym)
{
    return lu_f64::launch_lu_nopivot_fused(ctx, d_A, lda, n, d_ipiv, d_perm, d_flags, d_pivstat, d_Tinv, sym);
}
cudaError_t fd_launch_lu_f32(fd_ctx* ctx, float* d_A, int lda, int n, int* d_ipiv, int* d_perm, int* d_flags,
                             double* d_pivstat, int* d_win)
{
    return lu_f32::launch_lu(ctx, d_A, lda, n, d_ipiv, d_perm, d_flags, d_pivstat, d_win);
}
cudaError_t fd_launch_lu_nopivot_f32(fd_ctx* ctx, float* d_A, int lda, int n, int* d_ipiv, int* d_perm, int* d_flags,
                                     double* d_pivstat)
{
    return lu_f32::launch_lu_nopivot(ctx, d_A, lda, n, d_ipiv, d_perm, d_flags, d_pivstat);
}
