// fd_factor.cu -- K2: blocked right-looking LU with partial pivoting, FP64, column-major, in place.
//
// Replaces the solver inside alglib::rbfbuildmodel (reference SOP_FaceDeform.cpp:363) by the dense direct
// factorisation north_star names.  The saddle-point system is symmetric indefinite (and non-symmetric with
// per-centre QNN radii), so the general path is pivoted LU:
//   for each block column k0 (width NB):
//     k_lu_panel   : unblocked LU of the m x NB panel by one CTA (panel staged in shared memory when it fits)
//     k_lu_swap_trsm: row interchanges on every other column, and U12 = L11^-1 A12 on the columns to the right
//     k_lu_gemm    : A22 -= L21 * U12  (register-tiled FP64 FMA, 64x64 tile per CTA)
//   k_lu_perm     : folds the interchanges into one permutation vector for the right-hand-side gather
#include "fd_internal.h"

namespace {

constexpr int NB = 32;            // block-column width
constexpr int PANEL_THREADS = 1024;
constexpr int PANEL_SMEM_MAX = 200 * 1024;

struct ArgMax {
    double v;
    int i;
};

__device__ __forceinline__ ArgMax argmax_combine(ArgMax a, ArgMax b)
{
    // larger magnitude wins; ties -> lower row index (deterministic, matches a serial first-max scan)
    if (b.v > a.v || (b.v == a.v && b.i < a.i)) return b;
    return a;
}

// Unblocked LU with partial pivoting of the panel A[k0:n, k0:k0+nb].
// P points at the panel storage (global memory or the shared-memory copy) with leading dimension ldp.
template <bool IN_SMEM>
__global__ void __launch_bounds__(PANEL_THREADS) k_lu_panel(double* __restrict__ A, int lda, int n, int k0, int nb,
                                                            int* __restrict__ ipiv, int* __restrict__ flags,
                                                            double* __restrict__ pivstat)
{
    extern __shared__ double s_panel[];
    __shared__ ArgMax s_red[PANEL_THREADS / 32];
    __shared__ double s_pivrow[NB];
    __shared__ int s_piv;
    const int m = n - k0;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    double* G = A + (size_t)k0 * lda + k0; // panel origin in global memory
    const int ldp = IN_SMEM ? (m | 1) : lda;
    double* P = IN_SMEM ? s_panel : G;
    if (IN_SMEM) {
        for (int c = 0; c < nb; ++c)
            for (int r = tid; r < m; r += blockDim.x) P[(size_t)c * ldp + r] = G[(size_t)c * lda + r];
        __syncthreads();
    }
    double pmin = pivstat[0], pmax = pivstat[1];
    for (int j = 0; j < nb; ++j) {
        // (1) pivot search in column j, rows j..m-1
        ArgMax best = {-1.0, 0x7fffffff};
        const double* col = P + (size_t)j * ldp;
        for (int r = j + tid; r < m; r += blockDim.x) {
            const double v = fabs(col[r]);
            if (v > best.v) best = {v, r}; // rows visited in increasing order per thread
        }
        for (int o = 16; o > 0; o >>= 1) {
            ArgMax other = {__shfl_xor_sync(0xffffffffu, best.v, o), __shfl_xor_sync(0xffffffffu, best.i, o)};
            best = argmax_combine(best, other);
        }
        if (lane == 0) s_red[warp] = best;
        __syncthreads();
        if (warp == 0) {
            best = lane < nwarps ? s_red[lane] : ArgMax{-1.0, 0x7fffffff};
            for (int o = 16; o > 0; o >>= 1) {
                ArgMax other = {__shfl_xor_sync(0xffffffffu, best.v, o), __shfl_xor_sync(0xffffffffu, best.i, o)};
                best = argmax_combine(best, other);
            }
            if (lane == 0) {
                if (best.i >= m) best.i = j; // an all-NaN column: keep the diagonal, flagged singular below
                s_piv = best.i;
                ipiv[k0 + j] = k0 + best.i;
                if (!(best.v > 0.0) && flags[FD_FLAG_SINGULAR] == 0) flags[FD_FLAG_SINGULAR] = k0 + j + 1;
                pmin = fmin(pmin, best.v);
                pmax = fmax(pmax, best.v);
            }
        }
        __syncthreads();
        const int p = s_piv;
        // (2) swap rows j and p inside the panel; keep the pivot row in shared memory
        if (tid < nb) {
            const double a = P[(size_t)tid * ldp + j], b = P[(size_t)tid * ldp + p];
            P[(size_t)tid * ldp + j] = b;
            P[(size_t)tid * ldp + p] = a;
            s_pivrow[tid] = b;
        }
        __syncthreads();
        // (3) scale the column and rank-1 update the columns to its right
        const double piv = s_pivrow[j];
        if (piv != 0.0) {
            const double inv = 1.0 / piv;
            for (int r = j + 1 + tid; r < m; r += blockDim.x) {
                const double l = P[(size_t)j * ldp + r] * inv;
                P[(size_t)j * ldp + r] = l;
#pragma unroll 4
                for (int c = j + 1; c < nb; ++c) P[(size_t)c * ldp + r] -= l * s_pivrow[c];
            }
        }
        __syncthreads();
    }
    if (IN_SMEM) {
        for (int c = 0; c < nb; ++c)
            for (int r = tid; r < m; r += blockDim.x) G[(size_t)c * lda + r] = P[(size_t)c * ldp + r];
    }
    if (tid == 0) {
        pivstat[0] = pmin;
        pivstat[1] = pmax;
    }
}

// One thread per column outside the panel: apply the nb interchanges; columns right of the panel also get
// the unit-lower triangular solve with L11 (staged in shared memory).
__global__ void __launch_bounds__(128) k_lu_swap_trsm(double* __restrict__ A, int lda, int n, int k0, int nb,
                                                      const int* __restrict__ ipiv)
{
    __shared__ double s_L[NB][NB + 1];
    __shared__ int s_ip[NB];
    for (int t = threadIdx.x; t < NB * NB; t += blockDim.x) {
        const int r = t % NB, c = t / NB;
        s_L[r][c] = (r < nb && c < nb && r > c) ? A[(size_t)(k0 + c) * lda + k0 + r] : 0.0;
    }
    if (threadIdx.x < NB) s_ip[threadIdx.x] = threadIdx.x < nb ? ipiv[k0 + threadIdx.x] : 0;
    __syncthreads();
    int c = blockIdx.x * blockDim.x + threadIdx.x; // index over the n - nb columns outside the panel
    if (c >= n - nb) return;
    if (c >= k0) c += nb;
    double* col = A + (size_t)c * lda;
    // interchanges, in order (row p >= k0 + j)
    for (int j = 0; j < nb; ++j) {
        const int p = s_ip[j];
        if (p != k0 + j) {
            const double t = col[k0 + j];
            col[k0 + j] = col[p];
            col[p] = t;
        }
    }
    if (c < k0) return;
    double x[NB];
#pragma unroll
    for (int j = 0; j < NB; ++j) x[j] = j < nb ? col[k0 + j] : 0.0;
    { // right of the panel: forward substitution with the unit-lower L11
#pragma unroll
        for (int j = 0; j < NB; ++j) {
            const double xj = x[j];
#pragma unroll
            for (int r = j + 1; r < NB; ++r) x[r] -= s_L[r][j] * xj;
        }
    }
#pragma unroll
    for (int j = 0; j < NB; ++j)
        if (j < nb) col[k0 + j] = x[j];
}

// C[m2 x m2] -= L21[m2 x nb] * U12[nb x m2]; 64x64 tile per CTA, 256 threads, 4x4 outputs per thread.
constexpr int GT = 64;
__global__ void __launch_bounds__(256) k_lu_gemm(double* __restrict__ A, int lda, int n, int k0, int nb)
{
    __shared__ double s_a[NB][GT + 2]; // L21 tile, [k][row]
    __shared__ double s_b[NB][GT + 2]; // U12 tile, [k][col]
    const int r0 = k0 + nb + blockIdx.x * GT;
    const int c0 = k0 + nb + blockIdx.y * GT;
    const int tid = threadIdx.x;
    for (int t = tid; t < NB * GT; t += 256) {
        const int rr = t % GT, k = t / GT;
        s_a[k][rr] = (k < nb && r0 + rr < n) ? A[(size_t)(k0 + k) * lda + r0 + rr] : 0.0;
    }
    for (int t = tid; t < NB * GT; t += 256) {
        const int k = t % NB, cc = t / NB;
        s_b[k][cc] = (k < nb && c0 + cc < n) ? A[(size_t)(c0 + cc) * lda + k0 + k] : 0.0;
    }
    __syncthreads();
    const int tr = (tid % 16) * 4, tc = (tid / 16) * 4;
    double acc[4][4] = {};
#pragma unroll 8
    for (int k = 0; k < NB; ++k) {
        double a[4], b[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            a[i] = s_a[k][tr + i];
            b[i] = s_b[k][tc + i];
        }
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] += a[i] * b[j];
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int c = c0 + tc + j;
        if (c >= n) continue;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int r = r0 + tr + i;
            if (r < n) A[(size_t)c * lda + r] -= acc[i][j];
        }
    }
}

// perm[i] = original row that ends up in row i after all interchanges (single CTA, shared-memory resident)
__global__ void __launch_bounds__(256) k_lu_perm(const int* __restrict__ ipiv, int n, int* __restrict__ perm)
{
    extern __shared__ int s_perm[];
    for (int i = threadIdx.x; i < n; i += blockDim.x) s_perm[i] = i;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 0; k < n; ++k) {
            const int p = ipiv[k];
            if (p != k) {
                const int t = s_perm[k];
                s_perm[k] = s_perm[p];
                s_perm[p] = t;
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) perm[i] = s_perm[i];
}

__global__ void k_lu_init(int* flags, double* pivstat)
{
    if (threadIdx.x == 0) {
        flags[FD_FLAG_SINGULAR] = 0;
        flags[FD_FLAG_NONFINITE] = 0;
        pivstat[0] = INFINITY;
        pivstat[1] = 0.0;
    }
}

} // namespace

cudaError_t fd_launch_lu(fd_ctx* ctx, double* d_A, int lda, int n, int* d_ipiv, int* d_perm, int* d_flags,
                         double* d_pivstat)
{
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(k_lu_panel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, PANEL_SMEM_MAX);
        cudaFuncSetAttribute(k_lu_perm, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
        attr_set = true;
    }
    cudaStream_t s = ctx->stream;
    k_lu_init<<<1, 32, 0, s>>>(d_flags, d_pivstat);
    ctx->launches += 1;
    for (int k0 = 0; k0 < n; k0 += NB) {
        const int nb = min(NB, n - k0);
        const int m = n - k0;
        const size_t smem = (size_t)(m | 1) * nb * sizeof(double);
        const int threads = min(PANEL_THREADS, fd_round_up(m, 32));
        if (smem <= (size_t)PANEL_SMEM_MAX)
            k_lu_panel<true><<<1, threads, smem, s>>>(d_A, lda, n, k0, nb, d_ipiv, d_flags, d_pivstat);
        else
            k_lu_panel<false><<<1, threads, 0, s>>>(d_A, lda, n, k0, nb, d_ipiv, d_flags, d_pivstat);
        ctx->launches += 1;
        if (n - nb > 0) {
            k_lu_swap_trsm<<<(n - nb + 127) / 128, 128, 0, s>>>(d_A, lda, n, k0, nb, d_ipiv);
            ctx->launches += 1;
        }
        const int m2 = n - k0 - nb;
        if (m2 > 0) {
            dim3 grid((m2 + GT - 1) / GT, (m2 + GT - 1) / GT);
            k_lu_gemm<<<grid, 256, 0, s>>>(d_A, lda, n, k0, nb);
            ctx->launches += 1;
        }
    }
    if ((size_t)n * sizeof(int) > 64 * 1024) return cudaErrorInvalidValue;
    k_lu_perm<<<1, 256, (size_t)n * sizeof(int), s>>>(d_ipiv, n, d_perm);
    ctx->launches += 1;
    return cudaGetLastError();
}
