// fd_refine.cu -- FP32 factorisation + FP64 iterative refinement (fd_params.factor_precision = FD_FACTOR_FP32_IR).
//
// BASELINE.json config 4 ("8192 control points ... FP64 vs FP32+refinement tolerance study") and north_star (2):
// "tiled right-looking LU ... FP64 DMMA, or FP32 with iterative refinement".  The system replaces the solver inside
// alglib::rbfbuildmodel (reference SOP_FaceDeform.cpp:363); the success criterion stays report.terminationtype == 1
// (:365-368), iterationscount carries the refinement sweeps (:370-373 prints it).
//
//   fit    A (FP64, kept) -> A32 = fl32(A);  P A32 = L U  in FP32 (the same blocked LU kernels, float instantiation)
//   solve  X = 0, R = B;  repeat:  D = U^-1 L^-1 P fl32(R)  (FP32) ;  X += D (FP64) ;  R = B - A X (FP64)
//          until |R|_max <= tol * |B|_max (converged), the residual stops shrinking (stagnated) or FD_IR_MAX_SWEEPS.
// Each sweep gains about -log10(cond(A) * 2^-24) digits, so the mode works while cond(A) is comfortably below 2^24 ~
// 1.7e7 and stagnates / diverges beyond (the study in profiles/ walks the Gaussian radius across that boundary).
#include "fd_internal.h"

namespace {

constexpr int SB = 32;

__global__ void __launch_bounds__(256) k_to_f32(const double* __restrict__ A, float* __restrict__ A32, size_t count)
{
    const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (i < count) A32[i] = (float)A[i];
}

// B[i][3f+k] = (double)(deform[f][i][k] - rest[i][k])  (FP32 subtract, SOP_FaceDeform.cpp:276-284), polynomial rows 0
__global__ void __launch_bounds__(256) k_ir_rhs(const float* __restrict__ rest, const float* __restrict__ deform, int N,
                                                int F, double* __restrict__ B, double* __restrict__ R,
                                                double* __restrict__ X, int ldw, unsigned long long* __restrict__ norm_b)
{
    const int c = blockIdx.x * 256 + threadIdx.x;
    const int i = blockIdx.y;
    double v = 0.0;
    if (c < ldw) {
        if (c < 3 * F && i < N) {
            const int f = c / 3, k = c - 3 * f;
            v = (double)(deform[((size_t)f * N + i) * 3 + k] - rest[3 * i + k]);
        }
        B[(size_t)i * ldw + c] = v;
        R[(size_t)i * ldw + c] = v;
        X[(size_t)i * ldw + c] = 0.0;
    }
    // |B|_max: non-negative doubles order like their bit patterns
    double mx = fabs(v);
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0 && mx > 0.0) atomicMax(norm_b, (unsigned long long)__double_as_longlong(mx));
}

// D32[i][c] = fl32(R[perm[i]][c])
__global__ void __launch_bounds__(256) k_ir_gather(const double* __restrict__ R, const int* __restrict__ perm, int ldw,
                                                   float* __restrict__ D32)
{
    const int c = blockIdx.x * 256 + threadIdx.x;
    const int i = blockIdx.y;
    if (c < ldw) D32[(size_t)i * ldw + c] = (float)R[(size_t)perm[i] * ldw + c];
}

// X += D32 (FP64 accumulate)
__global__ void __launch_bounds__(256) k_ir_axpy(const float* __restrict__ D32, double* __restrict__ X, size_t count)
{
    const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (i < count) X[i] += (double)D32[i];
}

// FP32 triangular sweeps on the FP32 factors (same blocking as the FP64 fallback in fd_solve.cu)
template <bool LOWER>
__global__ void __launch_bounds__(128) k_ir_trsm_diag(const float* __restrict__ A, int lda, int k0, int nb,
                                                      float* __restrict__ B, int ldw, int nrhs)
{
    __shared__ float s_T[SB][SB + 1];
    for (int t = threadIdx.x; t < SB * SB; t += blockDim.x) {
        const int r = t % SB, c = t / SB;
        s_T[r][c] = (r < nb && c < nb) ? A[(size_t)(k0 + c) * lda + k0 + r] : (r == c ? 1.0f : 0.0f);
    }
    __syncthreads();
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nrhs) return;
    float x[SB];
#pragma unroll
    for (int j = 0; j < SB; ++j) x[j] = j < nb ? B[(size_t)(k0 + j) * ldw + c] : 0.0f;
    if (LOWER) {
#pragma unroll
        for (int j = 0; j < SB; ++j) {
            const float xj = x[j];
#pragma unroll
            for (int r = j + 1; r < SB; ++r) x[r] -= s_T[r][j] * xj;
        }
    } else {
#pragma unroll
        for (int j = SB - 1; j >= 0; --j) {
            const float xj = x[j] / s_T[j][j];
            x[j] = xj;
#pragma unroll
            for (int r = 0; r < j; ++r) x[r] -= s_T[r][j] * xj;
        }
    }
#pragma unroll
    for (int j = 0; j < SB; ++j)
        if (j < nb) B[(size_t)(k0 + j) * ldw + c] = x[j];
}

template <bool LOWER>
__global__ void __launch_bounds__(256) k_ir_trsm_update(const float* __restrict__ A, int lda, int n, int k0, int nb,
                                                        float* __restrict__ B, int ldw, int nrhs)
{
    __shared__ float s_t[SB][64 + 1]; // [k][row]
    __shared__ float s_x[SB][32 + 1]; // [k][col]
    const int row_begin = LOWER ? k0 + nb : 0;
    const int row_end = LOWER ? n : k0;
    const int r0 = row_begin + blockIdx.y * 64;
    const int c0 = blockIdx.x * 32;
    for (int t = threadIdx.x; t < SB * 64; t += 256) {
        const int rr = t % 64, k = t / 64;
        s_t[k][rr] = (k < nb && r0 + rr < row_end) ? A[(size_t)(k0 + k) * lda + r0 + rr] : 0.0f;
    }
    for (int t = threadIdx.x; t < SB * 32; t += 256) {
        const int cc = t % 32, k = t / 32;
        s_x[k][cc] = (k < nb && c0 + cc < nrhs) ? B[(size_t)(k0 + k) * ldw + c0 + cc] : 0.0f;
    }
    __syncthreads();
    const int cc = threadIdx.x % 32, rg = threadIdx.x / 32;
    float acc[8] = {};
#pragma unroll 8
    for (int k = 0; k < SB; ++k) {
        const float xv = s_x[k][cc];
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] += s_t[k][rg * 8 + i] * xv;
    }
    if (c0 + cc < nrhs) {
        float bv[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int r = r0 + rg * 8 + i;
            bv[i] = r < row_end ? B[(size_t)r * ldw + c0 + cc] : 0.0f;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int r = r0 + rg * 8 + i;
            if (r < row_end) B[(size_t)r * ldw + c0 + cc] = bv[i] - acc[i];
        }
    }
}

// R = B - A X in FP64 (A column-major n x n, X / B / R row-major n x ldw); CTA tile 64 rows x 32 columns, k in steps
// of 32 through shared memory; |R|_max via atomicMax on the bit pattern.  A is read exactly once per column tile.
__global__ void __launch_bounds__(256) k_ir_residual(const double* __restrict__ A, int lda, int n,
                                                     const double* __restrict__ X, const double* __restrict__ B,
                                                     double* __restrict__ R, int ldw, int nrhs,
                                                     unsigned long long* __restrict__ norm_r)
{
    __shared__ double s_a[SB][64 + 1]; // [k][row]
    __shared__ double s_x[SB][32 + 1]; // [k][col]
    const int r0 = blockIdx.y * 64, c0 = blockIdx.x * 32;
    const int cc = threadIdx.x % 32, rg = threadIdx.x / 32;
    double acc[8] = {};
    for (int k0 = 0; k0 < n; k0 += SB) {
        __syncthreads();
        for (int t = threadIdx.x; t < SB * 64; t += 256) {
            const int rr = t % 64, k = t / 64;
            s_a[k][rr] = (k0 + k < n && r0 + rr < n) ? A[(size_t)(k0 + k) * lda + r0 + rr] : 0.0;
        }
        for (int t = threadIdx.x; t < SB * 32; t += 256) {
            const int c = t % 32, k = t / 32;
            s_x[k][c] = (k0 + k < n && c0 + c < nrhs) ? X[(size_t)(k0 + k) * ldw + c0 + c] : 0.0;
        }
        __syncthreads();
#pragma unroll 8
        for (int k = 0; k < SB; ++k) {
            const double xv = s_x[k][cc];
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] = fma(s_a[k][rg * 8 + i], xv, acc[i]);
        }
    }
    double mx = 0.0;
    if (c0 + cc < nrhs) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int r = r0 + rg * 8 + i;
            if (r < n) {
                const double v = B[(size_t)r * ldw + c0 + cc] - acc[i];
                R[(size_t)r * ldw + c0 + cc] = v;
                mx = fmax(mx, fabs(v));
            }
        }
    }
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0 && mx > 0.0) atomicMax(norm_r, (unsigned long long)__double_as_longlong(mx));
}

} // namespace

cudaError_t fd_launch_to_f32(fd_ctx* ctx, const double* d_A, float* d_A32, size_t count)
{
    k_to_f32<<<(unsigned)((count + 255) / 256), 256, 0, ctx->stream>>>(d_A, d_A32, count);
    ctx->launches += 1;
    return cudaGetLastError();
}

// Solves for all right-hand sides by refinement; synchronises the stream once per sweep to read the residual norm
// (this mode is the tolerance study of config 4, not the per-frame hot path).  On return m->d_W holds X and
// m->ir_* the sweep count / final relative residual / outcome.
cudaError_t fd_refine_solve(fd_ctx* ctx, fd_model* m, const float* d_deform, int F)
{
    cudaStream_t s = ctx->stream;
    const int n = m->n, nrhs = 3 * F, ldw = m->ldw;
    const size_t count = (size_t)n * ldw;
    unsigned long long* d_norm = reinterpret_cast<unsigned long long*>(m->d_ir_norm); // [0] |B|, [1] |R|
    cudaError_t e = cudaMemsetAsync(d_norm, 0, 2 * sizeof(unsigned long long), s);
    if (e != cudaSuccess) return e;
    dim3 grid_rows((ldw + 255) / 256, n);
    k_ir_rhs<<<grid_rows, 256, 0, s>>>(m->d_rest, d_deform, m->N, F, m->d_B, m->d_R, m->d_W, ldw, d_norm);
    ctx->launches += 1;
    double h_norm[2] = {0.0, 0.0};
    e = cudaMemcpyAsync(h_norm, d_norm, sizeof(h_norm), cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) return e;
    const double norm_b = h_norm[0];
    m->ir_sweeps = 0;
    m->ir_residual = 0.0;
    m->ir_converged = true;
    if (!(norm_b > 0.0)) return cudaSuccess; // zero deltas: X = 0
    const int cblocks = (nrhs + 127) / 128;
    double prev = 1.0, rel = 1.0;
    int stalled = 0;
    m->ir_converged = false;
    for (int sweep = 1; sweep <= FD_IR_MAX_SWEEPS; ++sweep) {
        k_ir_gather<<<grid_rows, 256, 0, s>>>(m->d_R, m->d_perm, ldw, m->d_D32);
        ctx->launches += 1;
        for (int k0 = 0; k0 < n; k0 += SB) { // L y = P r
            const int nb = min(SB, n - k0);
            k_ir_trsm_diag<true><<<cblocks, 128, 0, s>>>(m->d_A32, m->lda, k0, nb, m->d_D32, ldw, nrhs);
            const int rows = n - k0 - nb;
            if (rows > 0) {
                dim3 grid((nrhs + 31) / 32, (rows + 63) / 64);
                k_ir_trsm_update<true><<<grid, 256, 0, s>>>(m->d_A32, m->lda, n, k0, nb, m->d_D32, ldw, nrhs);
                ctx->launches += 1;
            }
            ctx->launches += 1;
        }
        for (int k0 = (n - 1) / SB * SB; k0 >= 0; k0 -= SB) { // U d = y
            const int nb = min(SB, n - k0);
            k_ir_trsm_diag<false><<<cblocks, 128, 0, s>>>(m->d_A32, m->lda, k0, nb, m->d_D32, ldw, nrhs);
            ctx->launches += 1;
            if (k0 > 0) {
                dim3 grid((nrhs + 31) / 32, (k0 + 63) / 64);
                k_ir_trsm_update<false><<<grid, 256, 0, s>>>(m->d_A32, m->lda, n, k0, nb, m->d_D32, ldw, nrhs);
                ctx->launches += 1;
            }
        }
        k_ir_axpy<<<(unsigned)((count + 255) / 256), 256, 0, s>>>(m->d_D32, m->d_W, count);
        cudaMemsetAsync(d_norm + 1, 0, sizeof(unsigned long long), s);
        dim3 grid_res((nrhs + 31) / 32, (n + 63) / 64);
        k_ir_residual<<<grid_res, 256, 0, s>>>(m->d_A, m->lda, n, m->d_W, m->d_B, m->d_R, ldw, nrhs, d_norm + 1);
        ctx->launches += 2;
        e = cudaMemcpyAsync(h_norm, d_norm, sizeof(h_norm), cudaMemcpyDeviceToHost, s);
        if (e == cudaSuccess) e = cudaStreamSynchronize(s);
        if (e != cudaSuccess) return e;
        rel = h_norm[1] / norm_b;
        m->ir_sweeps = sweep;
        m->ir_residual = rel;
        if (!(rel == rel) || rel > 1e30) break;               // NaN / blown up: diverged
        if (rel <= FD_IR_TOLERANCE) { m->ir_converged = true; break; }
        stalled = rel > 0.5 * prev ? stalled + 1 : 0;         // less than one bit gained
        if (stalled >= 2) {                                   // stagnated: at the FP64 residual floor, or not converging
            m->ir_converged = rel <= FD_IR_FLOOR_OK;
            break;
        }
        prev = rel;
    }
    return cudaGetLastError();
}
