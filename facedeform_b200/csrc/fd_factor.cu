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
#include <stdlib.h>

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
    __shared__ ArgMax s_red[2][PANEL_THREADS / 32];
    __shared__ double s_pivrow[2][NB];
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
    // Two block-wide barriers per column: every thread owns a fixed set of rows (r = tid mod blockDim), so the pivot
    // search of column j+1 only reads elements the same thread updated in column j.
    for (int j = 0; j < nb; ++j) {
        // (1) pivot search in column j, rows j..m-1
        ArgMax best = {-1.0, 0x7fffffff};
        const double* col = P + (size_t)j * ldp;
        for (int r = tid; r < m; r += blockDim.x) {
            if (r < j) continue;
            const double v = fabs(col[r]);
            if (v > best.v) best = {v, r}; // rows visited in increasing order per thread
        }
        for (int o = 16; o > 0; o >>= 1) {
            ArgMax other = {__shfl_xor_sync(0xffffffffu, best.v, o), __shfl_xor_sync(0xffffffffu, best.i, o)};
            best = argmax_combine(best, other);
        }
        if (lane == 0) s_red[j & 1][warp] = best;
        __syncthreads();
        best = lane < nwarps ? s_red[j & 1][lane] : ArgMax{-1.0, 0x7fffffff}; // every warp reduces the partials itself
        for (int o = 16; o > 0; o >>= 1) {
            ArgMax other = {__shfl_xor_sync(0xffffffffu, best.v, o), __shfl_xor_sync(0xffffffffu, best.i, o)};
            best = argmax_combine(best, other);
        }
        if (best.i >= m) best.i = j; // an all-NaN column: keep the diagonal, flagged singular below
        const int p = best.i;
        if (tid == 0) {
            ipiv[k0 + j] = k0 + p;
            if (!(best.v > 0.0) && flags[FD_FLAG_SINGULAR] == 0) flags[FD_FLAG_SINGULAR] = k0 + j + 1;
            pmin = fmin(pmin, best.v);
            pmax = fmax(pmax, best.v);
        }
        // (2) swap rows j and p inside the panel; keep the pivot row in shared memory
        if (tid < nb) {
            const double x = P[(size_t)tid * ldp + j], y = P[(size_t)tid * ldp + p];
            P[(size_t)tid * ldp + j] = y;
            P[(size_t)tid * ldp + p] = x;
            s_pivrow[j & 1][tid] = y;
        }
        __syncthreads();
        // (3) scale the column and rank-1 update the columns to its right (own rows only)
        const double piv = s_pivrow[j & 1][j];
        if (piv != 0.0) {
            const double inv = 1.0 / piv;
            for (int r = tid; r < m; r += blockDim.x) {
                if (r <= j) continue;
                const double l = P[(size_t)j * ldp + r] * inv;
                P[(size_t)j * ldp + r] = l;
#pragma unroll 4
                for (int c = j + 1; c < nb; ++c) P[(size_t)c * ldp + r] -= l * s_pivrow[j & 1][c];
            }
        }
    }
    __syncthreads();
    if (IN_SMEM) {
        for (int c = 0; c < nb; ++c)
            for (int r = tid; r < m; r += blockDim.x) G[(size_t)c * lda + r] = P[(size_t)c * ldp + r];
    }
    if (tid == 0) {
        pivstat[0] = pmin;
        pivstat[1] = pmax;
    }
}

// Register-resident panel for short panels (m <= RPT * 256 rows): thread t keeps the not-yet-eliminated part of rows
// t, t+256, ... in registers.  The row registers are shifted left by one column per step (fused into the rank-1
// update: a'[c-1] = a[c] - l * u[c]), so the loop body has static register indices and is NOT unrolled over the
// columns -- a fully unrolled panel is ~200 KB of straight-line code executed once and ran at instruction-fetch
// speed.  Finished L and U entries are collected in a shared-memory copy of the panel and written out at the end.
// Two block barriers per column.
template <int RPT>
__global__ void __launch_bounds__(256) k_lu_panel_reg(double* __restrict__ A, int lda, int n, int k0, int nb,
                                                      int* __restrict__ ipiv, int* __restrict__ flags,
                                                      double* __restrict__ pivstat)
{
    extern __shared__ double s_out[]; // [NB][ldo] finished panel, column-major
    __shared__ ArgMax s_red[2][8];
    __shared__ double s_rowj[2][NB]; // the row that sat at position j before the swap
    __shared__ double s_rowp[2][NB]; // the pivot row
    const int m = n - k0;
    const int ldo = m | 1;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double* G = A + (size_t)k0 * lda + k0;
    double a[RPT][NB];
#pragma unroll
    for (int q = 0; q < RPT; ++q) {
        const int r = tid + q * 256;
#pragma unroll
        for (int c = 0; c < NB; ++c) a[q][c] = (r < m && c < nb) ? G[(size_t)c * lda + r] : 0.0;
    }
    double pmin = pivstat[0], pmax = pivstat[1];
#pragma unroll 1
    for (int j = 0; j < nb; ++j) {
        ArgMax best = {-1.0, 0x7fffffff};
#pragma unroll
        for (int q = 0; q < RPT; ++q) {
            const int r = tid + q * 256;
            const double v = fabs(a[q][0]);
            if (r >= j && r < m && v > best.v) best = {v, r};
        }
        for (int o = 16; o > 0; o >>= 1) {
            ArgMax other = {__shfl_xor_sync(0xffffffffu, best.v, o), __shfl_xor_sync(0xffffffffu, best.i, o)};
            best = argmax_combine(best, other);
        }
        if (lane == 0) s_red[j & 1][warp] = best;
        __syncthreads();
        best = lane < 8 ? s_red[j & 1][lane] : ArgMax{-1.0, 0x7fffffff};
        for (int o = 4; o > 0; o >>= 1) {
            ArgMax other = {__shfl_xor_sync(0xffffffffu, best.v, o), __shfl_xor_sync(0xffffffffu, best.i, o)};
            best = argmax_combine(best, other);
        }
        best.v = __shfl_sync(0xffffffffu, best.v, 0);
        best.i = __shfl_sync(0xffffffffu, best.i, 0);
        if (best.i >= m) best.i = j;
        const int p = best.i;
        if (tid == 0) {
            ipiv[k0 + j] = k0 + p;
            if (!(best.v > 0.0) && flags[FD_FLAG_SINGULAR] == 0) flags[FD_FLAG_SINGULAR] = k0 + j + 1;
            pmin = fmin(pmin, best.v);
            pmax = fmax(pmax, best.v);
        }
        // publish rows j and p (by their owners)
#pragma unroll
        for (int q = 0; q < RPT; ++q) {
            const int r = tid + q * 256;
            if (r == j) {
#pragma unroll
                for (int c = 0; c < NB; ++c) s_rowj[j & 1][c] = a[q][c];
            }
            if (r == p) {
#pragma unroll
                for (int c = 0; c < NB; ++c) s_rowp[j & 1][c] = a[q][c];
            }
        }
        __syncthreads();
        const double* u = s_rowp[j & 1]; // u[c] = U(j, j + c)
        const double piv = u[0];
        const double inv = piv != 0.0 ? 1.0 / piv : 0.0;
        if (tid < nb - j) s_out[(size_t)(j + tid) * ldo + j] = u[tid];          // row j of U is final
        if (tid < j && p != j) {                                                 // interchange in the finished L columns
            const double x = s_out[(size_t)tid * ldo + j];
            s_out[(size_t)tid * ldo + j] = s_out[(size_t)tid * ldo + p];
            s_out[(size_t)tid * ldo + p] = x;
        }
#pragma unroll
        for (int q = 0; q < RPT; ++q) {
            const int r = tid + q * 256;
            if (r == p && p != j) { // the row that was at position j moves to position p
#pragma unroll
                for (int c = 0; c < NB; ++c) a[q][c] = s_rowj[j & 1][c];
            }
            if (r > j && r < m) {
                const double l = a[q][0] * inv;
                s_out[(size_t)j * ldo + r] = l;
#pragma unroll
                for (int c = 1; c < NB; ++c) a[q][c - 1] = a[q][c] - l * u[c]; // rank-1 update fused with the shift
                a[q][NB - 1] = 0.0;
            }
        }
    }
    __syncthreads();
    for (int c = 0; c < nb; ++c)
        for (int r = tid; r < m; r += 256) G[(size_t)c * lda + r] = s_out[(size_t)c * ldo + r];
    if (tid == 0) {
        pivstat[0] = pmin;
        pivstat[1] = pmax;
    }
}

// One thread per column outside the panel: apply the nb interchanges; columns right of the panel also get the
// unit-lower triangular solve with L11.  The interchanges touch at most 2 nb distinct rows (the nb top rows and the
// pivot rows); warp 0 lists them once, every thread then gathers its column's values of those rows with independent
// loads, replays the swaps in shared memory and scatters the rows back -- no chain of dependent global accesses.
constexpr int SW_THREADS = 128;
__global__ void __launch_bounds__(SW_THREADS) k_lu_swap_trsm(double* __restrict__ A, int lda, int n, int k0, int nb,
                                                             const int* __restrict__ ipiv)
{
    extern __shared__ double s_vals[]; // [2 * NB][SW_THREADS]
    __shared__ double s_L[NB][NB + 1];
    __shared__ int s_rows[2 * NB];
    __shared__ int s_ib[NB];
    __shared__ int s_cnt;
    const int tid = threadIdx.x, lane = tid & 31;
    for (int t = tid; t < NB * NB; t += SW_THREADS) {
        const int r = t % NB, c = t / NB;
        s_L[r][c] = (r < nb && c < nb && r > c) ? A[(size_t)(k0 + c) * lda + k0 + r] : 0.0;
    }
    if (tid < 32) {
        if (lane < nb) s_rows[lane] = k0 + lane;
        int cnt = nb;
        __syncwarp();
        for (int j = 0; j < nb; ++j) {
            const int p = ipiv[k0 + j];
            int ib;
            if (p < k0 + nb) {
                ib = p - k0;
            } else {
                const bool hit = (nb + lane < cnt) && s_rows[nb + lane] == p;
                const unsigned m = __ballot_sync(0xffffffffu, hit);
                if (m) {
                    ib = nb + __ffs(m) - 1;
                } else {
                    if (lane == 0) s_rows[cnt] = p;
                    ib = cnt++;
                    __syncwarp();
                }
            }
            if (lane == 0) s_ib[j] = ib;
        }
        if (lane == 0) s_cnt = cnt;
    }
    __syncthreads();
    int c = blockIdx.x * SW_THREADS + tid; // index over the n - nb columns outside the panel
    if (c >= n - nb) return;
    if (c >= k0) c += nb;
    double* col = A + (size_t)c * lda;
    const int cnt = s_cnt;
    for (int e = 0; e < cnt; ++e) s_vals[e * SW_THREADS + tid] = col[s_rows[e]];
    for (int j = 0; j < nb; ++j) {
        const int ib = s_ib[j];
        if (ib != j) {
            const double t = s_vals[j * SW_THREADS + tid];
            s_vals[j * SW_THREADS + tid] = s_vals[ib * SW_THREADS + tid];
            s_vals[ib * SW_THREADS + tid] = t;
        }
    }
    if (c > k0) { // right of the panel: forward substitution with the unit-lower L11
        double x[NB];
#pragma unroll
        for (int j = 0; j < NB; ++j) x[j] = j < nb ? s_vals[j * SW_THREADS + tid] : 0.0;
#pragma unroll
        for (int j = 0; j < NB; ++j) {
            const double xj = x[j];
#pragma unroll
            for (int r = j + 1; r < NB; ++r) x[r] -= s_L[r][j] * xj;
        }
#pragma unroll
        for (int j = 0; j < NB; ++j)
            if (j < nb) s_vals[j * SW_THREADS + tid] = x[j];
    }
    for (int e = 0; e < cnt; ++e) col[s_rows[e]] = s_vals[e * SW_THREADS + tid];
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
        cudaFuncSetAttribute(k_lu_panel_reg<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, PANEL_SMEM_MAX);
        cudaFuncSetAttribute(k_lu_panel_reg<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, PANEL_SMEM_MAX);
        cudaFuncSetAttribute(k_lu_swap_trsm, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * NB * SW_THREADS * (int)sizeof(double));
        attr_set = true;
    }
    cudaStream_t s = ctx->stream;
    k_lu_init<<<1, 32, 0, s>>>(d_flags, d_pivstat);
    ctx->launches += 1;
    static const bool dbg = getenv("FD_LU_DEBUG") != nullptr; // development aid: in-stream time per kernel kind
    static cudaEvent_t ev[4];
    static bool ev_init = false;
    if (dbg && !ev_init) { for (auto& e : ev) cudaEventCreate(&e); ev_init = true; }
    float t_panel = 0, t_swap = 0, t_gemm = 0;
    for (int k0 = 0; k0 < n; k0 += NB) {
        if (dbg) cudaEventRecord(ev[0], s);
        const int nb = min(NB, n - k0);
        const int m = n - k0;
        const size_t smem = (size_t)(m | 1) * nb * sizeof(double);
        // few rows: fewer warps make the two barriers per column cheaper
        const int threads = m <= 1024 ? 256 : (m <= 4096 ? 512 : PANEL_THREADS);
        if (m <= 256)
            k_lu_panel_reg<1><<<1, 256, smem, s>>>(d_A, lda, n, k0, nb, d_ipiv, d_flags, d_pivstat);
        else if (m <= 512)
            k_lu_panel_reg<2><<<1, 256, smem, s>>>(d_A, lda, n, k0, nb, d_ipiv, d_flags, d_pivstat);
        else if (smem <= (size_t)PANEL_SMEM_MAX)
            k_lu_panel<true><<<1, threads, smem, s>>>(d_A, lda, n, k0, nb, d_ipiv, d_flags, d_pivstat);
        else
            k_lu_panel<false><<<1, threads, 0, s>>>(d_A, lda, n, k0, nb, d_ipiv, d_flags, d_pivstat);
        ctx->launches += 1;
        if (dbg) cudaEventRecord(ev[1], s);
        if (n - nb > 0) {
            k_lu_swap_trsm<<<(n - nb + SW_THREADS - 1) / SW_THREADS, SW_THREADS, 2 * NB * SW_THREADS * sizeof(double), s>>>(
                d_A, lda, n, k0, nb, d_ipiv);
            ctx->launches += 1;
        }
        if (dbg) cudaEventRecord(ev[2], s);
        const int m2 = n - k0 - nb;
        if (m2 > 0) {
            dim3 grid((m2 + GT - 1) / GT, (m2 + GT - 1) / GT);
            k_lu_gemm<<<grid, 256, 0, s>>>(d_A, lda, n, k0, nb);
            ctx->launches += 1;
        }
        if (dbg) {
            cudaEventRecord(ev[3], s);
            cudaEventSynchronize(ev[3]);
            float a, b, c;
            cudaEventElapsedTime(&a, ev[0], ev[1]);
            cudaEventElapsedTime(&b, ev[1], ev[2]);
            cudaEventElapsedTime(&c, ev[2], ev[3]);
            t_panel += a; t_swap += b; t_gemm += c;
        }
    }
    if (dbg) fprintf(stderr, "[fd_lu] n=%d: panel %.3f ms, swap+trsm %.3f ms, gemm %.3f ms\n", n, t_panel, t_swap, t_gemm);
    if ((size_t)n * sizeof(int) > 64 * 1024) return cudaErrorInvalidValue;
    k_lu_perm<<<1, 256, (size_t)n * sizeof(int), s>>>(d_ipiv, n, d_perm);
    ctx->launches += 1;
    return cudaGetLastError();
}
