// fd_solve.cu -- K2c: multi-RHS triangular solves for all 3F per-frame control displacements at once, and the
// packing of the solved weights into the evaluation tables.
//
// Replaces the delta half of the pack loop (reference SOP_FaceDeform.cpp:268-287: FP32 subtract, FP64 store)
// and the solve inside alglib::rbfbuildmodel (:363).  Right-hand sides / weights live row-major as
// (N + npoly) x ldw doubles, column 3f + k = frame f, axis k, so one row is one control point's weights for
// every frame -- the layout the evaluation kernels stage.
#include <cuda_fp16.h>
#include <stdlib.h>

#include "fd_internal.h"

namespace {

constexpr int SB = 32; // block size of the triangular sweeps

// B[i][3f+k] = (double)(deform[f][perm[i]][k] - rest[perm[i]][k])   (FP32 subtract), 0 for polynomial rows
__global__ void __launch_bounds__(256) k_build_rhs(const float* __restrict__ rest, const float* __restrict__ deform,
                                                   const int* __restrict__ perm, int N, int n, int F,
                                                   double* __restrict__ B, int ldw)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y;
    if (c >= ldw) return;
    double v = 0.0;
    const int src = perm ? perm[i] : i;
    if (c < 3 * F && src < N) {
        const int f = c / 3, k = c - 3 * f;
        const float d = deform[((size_t)f * N + src) * 3 + k] - rest[3 * src + k];
        v = (double)d;
    }
    B[(size_t)i * ldw + c] = v;
}

// ---- explicit inverse for per-cook solves -----------------------------------------------------------------------------
// X[i][c] = (row i of the permuted identity): perm == NULL gives I, else W[i][c] = (perm[i] == c)
__global__ void __launch_bounds__(256) k_identity_rhs(double* __restrict__ X, int n, int ld, const int* __restrict__ perm)
{
    const int c = blockIdx.x * 256 + threadIdx.x;
    const int i = blockIdx.y;
    if (c >= ld) return;
    const int src = perm ? perm[i] : i;
    X[(size_t)i * ld + c] = (c == src && c < n) ? 1.0 : 0.0;
}

// B8[c][q] = W[c][q] for q < nrhs (0 beyond): the right-hand sides out of the in-place weight block
__global__ void __launch_bounds__(256) k_gather_rhs8(const double* __restrict__ W, int ldw, int n, int nrhs, double* __restrict__ B8)
{
    const int t = blockIdx.x * 256 + threadIdx.x;
    if (t >= n * 8) return;
    const int c = t >> 3, q = t & 7;
    B8[t] = q < nrhs ? W[(size_t)c * ldw + q] : 0.0;
}

// out[i][q] = sum_c inv[i][c] * B8[c][q], q < NQ <= 8.  A CTA owns 16 rows (two per warp) and walks the columns in chunks
// of 512: the chunk of right-hand sides is staged transposed in shared memory ([q][c]: conflict-free reads), the rows of
// the inverse stream once, coalesced, straight from L2 / HBM -- the kernel is bound by those n^2 x 8 bytes.
constexpr int IA_ROWS = 16, IA_CH = 512;
template <int NQ>
__global__ void __launch_bounds__(256) k_inv_apply(const double* __restrict__ inv, int ld, int n, const double* __restrict__ B8,
                                                   double* __restrict__ out, int ldo)
{
    __shared__ double s_b[NQ][IA_CH];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int r0 = blockIdx.x * IA_ROWS + 2 * warp;
    const double* row0 = inv + (size_t)min(r0, n - 1) * ld;
    const double* row1 = inv + (size_t)min(r0 + 1, n - 1) * ld;
    double acc0[NQ], acc1[NQ];
#pragma unroll
    for (int q = 0; q < NQ; ++q) acc0[q] = acc1[q] = 0.0;
    for (int c0 = 0; c0 < n; c0 += IA_CH) {
        __syncthreads();
        for (int t = threadIdx.x; t < IA_CH * NQ; t += 256) {
            const int c = t / NQ, q = t - c * NQ;
            s_b[q][c] = c0 + c < n ? B8[(size_t)(c0 + c) * 8 + q] : 0.0;
        }
        __syncthreads();
        const int cend = min(IA_CH, n - c0);
#pragma unroll 4
        for (int c = lane; c < cend; c += 32) {
            const double a0 = __ldcs(row0 + c0 + c), a1 = __ldcs(row1 + c0 + c); // streamed: read once per solve
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
                const double bq = s_b[q][c];
                acc0[q] = fma(a0, bq, acc0[q]);
                acc1[q] = fma(a1, bq, acc1[q]);
            }
        }
    }
#pragma unroll
    for (int q = 0; q < NQ; ++q)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            acc0[q] += __shfl_xor_sync(0xffffffffu, acc0[q], o);
            acc1[q] += __shfl_xor_sync(0xffffffffu, acc1[q], o);
        }
    if (lane < ldo && lane < 8) { // padding columns of the weight block are zero, like after the sweeps
        double v0 = 0.0, v1 = 0.0;
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            v0 = lane == q ? acc0[q] : v0;
            v1 = lane == q ? acc1[q] : v1;
        }
        if (r0 < n) out[(size_t)r0 * ldo + lane] = v0;
        if (r0 + 1 < n) out[(size_t)(r0 + 1) * ldo + lane] = v1;
    }
}

// ---- a handful of right-hand sides (one frame per cook, the reference's own usage): one launch per 32-row block step.
// Every CTA recomputes x_k = T_kk^-1 b_k from the pre-inverted diagonal block (32 x 32 x nrhs FMAs, redundantly: cheaper
// than a second launch), then updates its own 64 rows outside the block: upd[rows] -= T[rows, k-block] x_k.  Sources and
// targets are separate buffers (forward: reads B, writes y_k to Y, updates B below; backward: reads Y, writes x_k to X,
// updates Y above), so no CTA reads rows another one writes in the same launch.
constexpr int FEW_MAX = 8;
template <bool LOWER>
__global__ void __launch_bounds__(256) k_few_step(const double* __restrict__ A, int lda, int n, int k0, int nb,
                                                  const double* __restrict__ Tinv_blk, const double* src, int lds,
                                                  double* xout, int ldx, double* upd, int ldu, int nrhs)
{
    __shared__ __align__(16) double s_I[SB * SB]; // s_I[j * 32 + r] = inverse[r][j]
    __shared__ double s_b[SB][FEW_MAX];
    __shared__ double s_x[SB][FEW_MAX];
    const int tid = threadIdx.x;
    for (int t = tid; t < SB * SB; t += 256) s_I[t] = Tinv_blk[t];
    {
        const int j = tid & 31, c = tid >> 5;
        s_b[j][c] = (j < nb && c < nrhs) ? src[(size_t)(k0 + j) * lds + c] : 0.0;
    }
    __syncthreads();
    {
        const int r = tid & 31, c = tid >> 5;
        double x = 0.0;
#pragma unroll 8
        for (int j = 0; j < SB; ++j) x = fma(s_I[j * SB + r], s_b[j][c], x);
        s_x[r][c] = x;
        if (blockIdx.x == 0 && r < nb && c < nrhs) xout[(size_t)(k0 + r) * ldx + c] = x;
    }
    __syncthreads();
    const int row_begin = LOWER ? k0 + nb : 0, row_end = LOWER ? n : k0;
    const int r = row_begin + blockIdx.x * 64 + (tid & 63);
    if (r >= row_end) return;
    const int cg = tid >> 6; // 4 column groups of 2
    double a0 = 0.0, a1 = 0.0;
#pragma unroll 8
    for (int k = 0; k < SB; ++k) {
        const double t = k < nb ? A[(size_t)(k0 + k) * lda + r] : 0.0;
        a0 = fma(t, s_x[k][2 * cg], a0);
        a1 = fma(t, s_x[k][2 * cg + 1], a1);
    }
    if (2 * cg < nrhs) upd[(size_t)r * ldu + 2 * cg] -= a0;
    if (2 * cg + 1 < nrhs) upd[(size_t)r * ldu + 2 * cg + 1] -= a1;
}

// Panel update of the blocked sweeps: C[r][c] -= sum_k T[r][kb + k] * X[kb + k][c] for r in [r_lo, r_hi), k < K.
// T column-major (the LU factors, or any column-major operand), X and C row-major.  CTA tile 128 x 128 on the FP64 tensor
// pipe (mma.sync.m8n8k4.f64): 8 warps as 4 x 2, a warp owns 32 x 64 outputs = 32 accumulator tiles and loads 12 fragments
// per 32 DMMAs (8192 FMAs); K in chunks of 16, double-buffered with cp.async (zero-filled beyond K, r_hi, nrhs).  Stage
// strides = 4 mod 16 doubles: the fragment loads of a half warp (4 rows x 4 k) fall into 16 different 8-byte banks.
constexpr int PG_TM = 128, PG_TN = 128, PG_KC = 16;
constexpr int PG_LDA = PG_TM + 4, PG_LDX = PG_TN + 4;
constexpr int PG_STAGE_DOUBLES = PG_KC * (PG_LDA + PG_LDX);
constexpr int PG_SMEM_BYTES = 2 * PG_STAGE_DOUBLES * 8;
__device__ __forceinline__ void pg_cp16(void* smem_dst, const void* gsrc, bool valid)
{
    const int bytes = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void pg_tile_dmma(const double* __restrict__ T, int lda, int r_lo, int r_hi, int K,
                                             const double* __restrict__ X, int ldx, double* __restrict__ C, int ldc, int nrhs)
{
    extern __shared__ __align__(16) double pg_smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp & 3, wn = warp >> 2, fr = lane >> 2, fk = lane & 3;
    const int r0 = r_lo + blockIdx.y * PG_TM, c0 = blockIdx.x * PG_TN;
    double acc[4][8][2];
#pragma unroll
    for (int mi = 0; mi < 4; ++mi)
#pragma unroll
        for (int ni = 0; ni < 8; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
    auto issue = [&](int k0, int st) {
        double* sa = pg_smem + st * PG_STAGE_DOUBLES;
        double* sx = sa + PG_KC * PG_LDA;
        for (int t = tid; t < PG_KC * (PG_TM / 2); t += 256) {
            const int k = t / (PG_TM / 2), q = t - k * (PG_TM / 2);
            const bool ok = k0 + k < K && r0 + 2 * q < r_hi; // r_hi even or the pair's second row is a valid, masked address
            pg_cp16(sa + k * PG_LDA + 2 * q, T + (size_t)(ok ? k0 + k : 0) * lda + (ok ? r0 + 2 * q : 0), ok);
        }
        for (int t = tid; t < PG_KC * (PG_TN / 2); t += 256) {
            const int k = t / (PG_TN / 2), q = t - k * (PG_TN / 2);
            const bool ok = k0 + k < K && c0 + 2 * q < nrhs;
            pg_cp16(sx + k * PG_LDX + 2 * q, X + (size_t)(ok ? k0 + k : 0) * ldx + (ok ? c0 + 2 * q : 0), ok);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    const int nch = (K + PG_KC - 1) / PG_KC;
    issue(0, 0);
    for (int ch = 0; ch < nch; ++ch) {
        if (ch + 1 < nch) {
            issue((ch + 1) * PG_KC, (ch + 1) & 1);
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        __syncthreads();
        const double* sa = pg_smem + (ch & 1) * PG_STAGE_DOUBLES + fk * PG_LDA + wm * 32 + fr;
        const double* sx = pg_smem + (ch & 1) * PG_STAGE_DOUBLES + PG_KC * PG_LDA + fk * PG_LDX + wn * 64 + fr;
#pragma unroll
        for (int k4 = 0; k4 < PG_KC / 4; ++k4) {
            double af[4], bf[8];
#pragma unroll
            for (int mi = 0; mi < 4; ++mi) af[mi] = sa[k4 * 4 * PG_LDA + mi * 8];
#pragma unroll
            for (int ni = 0; ni < 8; ++ni) bf[ni] = sx[k4 * 4 * PG_LDX + ni * 8];
#pragma unroll
            for (int mi = 0; mi < 4; ++mi)
#pragma unroll
                for (int ni = 0; ni < 8; ++ni)
                    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                                 : "+d"(acc[mi][ni][0]), "+d"(acc[mi][ni][1])
                                 : "d"(af[mi]), "d"(bf[ni]));
        }
        __syncthreads();
    }
#pragma unroll
    for (int mi = 0; mi < 4; ++mi) {
        const int r = r0 + wm * 32 + mi * 8 + fr;
        if (r >= r_hi) continue;
#pragma unroll
        for (int ni = 0; ni < 8; ++ni) {
            const int c = c0 + wn * 64 + ni * 8 + 2 * fk;
            double* dst = C + (size_t)r * ldc + c;
            if (c + 1 < nrhs) { // c is even and the row strides are multiples of 4 doubles: 16-byte aligned
                double2 v = *reinterpret_cast<double2*>(dst);
                v.x -= acc[mi][ni][0];
                v.y -= acc[mi][ni][1];
                *reinterpret_cast<double2*>(dst) = v;
            } else if (c < nrhs) {
                dst[0] -= acc[mi][ni][0];
            }
        }
    }
}

__global__ void __launch_bounds__(256, 1) k_panel_gemm(const double* __restrict__ A, int lda, int r_lo, int r_hi, int kb, int K,
                                                    double* __restrict__ B, int ldw, int nrhs)
{
    pg_tile_dmma(A + (size_t)kb * lda, lda, r_lo, r_hi, K, B + (size_t)kb * ldw, ldw, B, ldw, nrhs);
}

// C[r][c] -= sum_k A[r][k] X[k][c], r < rows, k < K: A column-major (lda), X and C row-major with the same row stride.
// (the residual update of the layered fit, fd_api.cu)
__global__ void __launch_bounds__(256, 1) k_gemm_sub(const double* __restrict__ A, int lda, int rows, int K,
                                                     const double* __restrict__ X, double* __restrict__ C, int ldw, int nrhs)
{
    pg_tile_dmma(A, lda, 0, rows, K, X, ldw, C, ldw, nrhs);
}

// ---- slab solve: one CTA owns RC right-hand-side columns for the whole forward/backward substitution --------------
// The slab (n x RC doubles) lives in shared memory, L and U stream once from L2; no inter-CTA dependency, one launch.
// Warp w owns slab columns {2w, 2w+1}; inside a 32-row diagonal block lane r holds row r, so the triangular solve
// of the block needs only warp shuffles.  The rows outside the block are updated by all 256 threads.
constexpr int RC = 16;
constexpr int SLAB_THREADS = 256;

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// stage the (transposed) inverse of a diagonal block: 8 KB = two 16-byte cp.async per thread
__device__ __forceinline__ void slab_prefetch_tinv(double* s_T, const double* __restrict__ Tinv)
{
    for (int t = threadIdx.x; t < SB * SB / 2; t += SLAB_THREADS) cp_async16(s_T + 2 * t, Tinv + 2 * t);
    cp_async_commit();
}

// One block step of a sweep.  s_T holds TinvT for this step (s_T[k * SB + r] = Tinv[r][k]); the inverse of the next
// step's block is prefetched into s_Tnext with cp.async, and the first pass of L/U rows is loaded into registers
// before the block product, so the L2 round trips overlap the arithmetic instead of adding to it.
template <bool LOWER>
__device__ __forceinline__ void slab_sweep_block(const double* __restrict__ A, int lda, int n, int k0, int nb,
                                                 const double* __restrict__ Tinv_next, double* __restrict__ s_B,
                                                 const double* __restrict__ s_T, double* __restrict__ s_Tnext,
                                                 double (*s_X)[RC])
{
    const int tid = threadIdx.x;
    if (Tinv_next) slab_prefetch_tinv(s_Tnext, Tinv_next);
    const int row_begin = LOWER ? k0 + nb : 0;
    const int row_end = LOWER ? n : k0;
    const int half = tid & 1; // 8 of the 16 columns
    // first pass of the outside rows: loads issued now (clamped, unconditional), consumed after the block product
    const int i0 = row_begin + (tid >> 1);
    double l[SB];
    {
        const double* Ti = A + (size_t)k0 * lda + min(i0, n - 1);
#pragma unroll
        for (int k = 0; k < SB; ++k) l[k] = Ti[(size_t)min(k, nb - 1) * lda];
    }
    {
        const int r = tid & 31, cpair = (tid >> 5) * 2;
        double x0 = 0.0, x1 = 0.0, y0 = 0.0, y1 = 0.0; // two partial sums per output: shorter dependent chains
#pragma unroll 8
        for (int k = 0; k < SB; k += 2) {
            const double t0 = s_T[k * SB + r], t1 = s_T[(k + 1) * SB + r];
            const double* b0 = s_B + (size_t)(k0 + min(k, nb - 1)) * RC + cpair;
            const double* b1 = s_B + (size_t)(k0 + min(k + 1, nb - 1)) * RC + cpair;
            const double m0 = k < nb ? 1.0 : 0.0, m1 = k + 1 < nb ? 1.0 : 0.0;
            x0 += t0 * (m0 * b0[0]);
            x1 += t0 * (m0 * b0[1]);
            y0 += t1 * (m1 * b1[0]);
            y1 += t1 * (m1 * b1[1]);
        }
        __syncthreads(); // every thread has read its B_k rows
        s_X[r][cpair] = r < nb ? x0 + y0 : 0.0;
        s_X[r][cpair + 1] = r < nb ? x1 + y1 : 0.0;
        if (r < nb) {
            s_B[(size_t)(k0 + r) * RC + cpair] = x0 + y0;
            s_B[(size_t)(k0 + r) * RC + cpair + 1] = x1 + y1;
        }
    }
    __syncthreads();
    // rows outside the block: B[i][:] -= T[i][k0:k0+nb] * X
    for (int i = i0; i < row_end; i += SLAB_THREADS / 2) {
        if (i != i0) {
            const double* Ti = A + (size_t)k0 * lda + i;
#pragma unroll
            for (int k = 0; k < SB; ++k) l[k] = Ti[(size_t)min(k, nb - 1) * lda];
        }
        double acc[8] = {};
#pragma unroll
        for (int k = 0; k < SB; ++k) {
#pragma unroll
            for (int c = 0; c < 8; ++c) acc[c] += l[k] * s_X[k][half * 8 + c]; // rows k >= nb of s_X are zero
        }
        double* bi = s_B + (size_t)i * RC + half * 8;
#pragma unroll
        for (int c = 0; c < 8; ++c) bi[c] -= acc[c];
    }
    cp_async_wait_all();
    __syncthreads();
}

// inverse of every 32x32 diagonal block of L (unit lower) and U, once per factorisation; thread c computes column c
__global__ void __launch_bounds__(64) k_lu_invdiag(const double* __restrict__ A, int lda, int n, double* __restrict__ Tinv)
{
    __shared__ double s_T[SB][SB + 1];
    const int blk = blockIdx.x, k0 = blk * SB, nb = min(SB, n - k0);
    for (int t = threadIdx.x; t < SB * SB; t += 64) {
        const int r = t % SB, c = t / SB;
        s_T[r][c] = (r < nb && c < nb) ? A[(size_t)(k0 + c) * lda + k0 + r] : (r == c ? 1.0 : 0.0);
    }
    __syncthreads();
    const int c = threadIdx.x & 31;
    const bool upper = threadIdx.x >= 32;
    double x[SB];
#pragma unroll
    for (int r = 0; r < SB; ++r) x[r] = r == c ? 1.0 : 0.0;
    if (!upper) {
#pragma unroll
        for (int j = 0; j < SB; ++j) {
            const double xj = x[j];
#pragma unroll
            for (int r = j + 1; r < SB; ++r) x[r] -= s_T[r][j] * xj;
        }
    } else {
#pragma unroll
        for (int j = SB - 1; j >= 0; --j) {
            const double xj = x[j] / s_T[j][j];
            x[j] = xj;
#pragma unroll
            for (int r = 0; r < j; ++r) x[r] -= s_T[r][j] * xj;
        }
    }
    double* out = Tinv + ((size_t)blk * 2 + (upper ? 1 : 0)) * SB * SB; // transposed: out[c * SB + r] = inverse[r][c]
#pragma unroll
    for (int r = 0; r < SB; ++r) out[c * SB + r] = x[r];
}

__global__ void __launch_bounds__(SLAB_THREADS) k_solve_slab(const double* __restrict__ A, int lda, int n, int N,
                                                             const int* __restrict__ perm, const float* __restrict__ rest,
                                                             const float* __restrict__ deform, int F,
                                                             const double* __restrict__ Tinv, double* __restrict__ W, int ldw)
{
    extern __shared__ double s_B[]; // n x RC
    __shared__ __align__(16) double s_T[2][SB * SB];
    __shared__ double s_X[SB][RC];
    const int c0 = blockIdx.x * RC;
    const int nrhs = 3 * F;
    const int nblk = (n + SB - 1) / SB;
    slab_prefetch_tinv(s_T[0], Tinv); // block 0 of L, overlapped with the right-hand-side build
    // right-hand sides, permuted: delta subtracted in FP32 then widened (SOP_FaceDeform.cpp:276-284)
    for (int t = threadIdx.x; t < n * RC; t += SLAB_THREADS) {
        const int i = t / RC, c = c0 + (t % RC);
        double v = 0.0;
        const int src = perm[i];
        if (c < nrhs && src < N) {
            const int f = c / 3, k = c - 3 * f;
            v = (double)(deform[((size_t)f * N + src) * 3 + k] - rest[3 * src + k]);
        }
        s_B[t] = v;
    }
    cp_async_wait_all();
    __syncthreads();
    int buf = 0;
    for (int blk = 0; blk < nblk; ++blk, buf ^= 1) { // L y = P b; the last step prefetches the first block of the U sweep
        const double* next = blk + 1 < nblk ? Tinv + (size_t)(blk + 1) * 2 * SB * SB : Tinv + ((size_t)(nblk - 1) * 2 + 1) * SB * SB;
        slab_sweep_block<true>(A, lda, n, blk * SB, min(SB, n - blk * SB), next, s_B, s_T[buf], s_T[buf ^ 1], s_X);
    }
    for (int blk = nblk - 1; blk >= 0; --blk, buf ^= 1) { // U x = y
        const double* next = blk > 0 ? Tinv + ((size_t)(blk - 1) * 2 + 1) * SB * SB : nullptr;
        slab_sweep_block<false>(A, lda, n, blk * SB, min(SB, n - blk * SB), next, s_B, s_T[buf], s_T[buf ^ 1], s_X);
    }
    for (int t = threadIdx.x; t < n * RC; t += SLAB_THREADS) {
        const int i = t / RC, c = c0 + (t % RC);
        if (c < ldw) W[(size_t)i * ldw + c] = s_B[t];
    }
}

// ---- slab solve on the FP64 tensor pipe -----------------------------------------------------------------------------
// One CTA owns 8 right-hand-side columns (one n8 tile of mma.sync.m8n8k4.f64) for the whole forward / backward sweep.
// The slab [n_pad][8] lives in shared memory.  Per 32-row block step: the diagonal block is applied through its
// pre-computed inverse (4 warps, one m8 tile each), then every row outside the block takes  B_i -= T_i,k * X_k  with
// the L (or U) panel streamed from L2 in 128-row chunks, double buffered with cp.async so that the round trips hide
// behind the previous chunk's DMMAs.  A DMMA needs 2 operand loads per 256 FMA (the X fragments stay in registers for
// the whole step), which takes the shared-memory pipe off the critical path that bounded the DFMA version.
constexpr int S8_RC = 8;
constexpr int S8_THREADS = 256;
constexpr int S8_CH = 128;           // panel rows per chunk
constexpr int S8_LDP = S8_CH + 8;    // chunk row stride [k][row]: = 8 mod 16 keeps the A-fragment loads conflict free
constexpr int S8_LDT = SB + 8;       // same for the inverted diagonal block
constexpr int S8_CHUNK_DOUBLES = SB * S8_LDP;
constexpr int S8_TINV_DOUBLES = SB * S8_LDT;

__device__ __forceinline__ void dmma8(double& c0, double& c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}
template <int N_> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N_) : "memory"); }

// chunk of the panel of block column k0: rows [r0, r0 + 128) x columns [k0, k0 + 32) -> s_P[k][row]
__device__ __forceinline__ void s8_prefetch_chunk(double* s_P, const double* __restrict__ A, int lda, int k0, int nb, int r0,
                                                  int n_pad)
{
    for (int t = threadIdx.x; t < SB * (S8_CH / 2); t += S8_THREADS) {
        const int kk = t / (S8_CH / 2), q = t % (S8_CH / 2);
        const int r = min(r0 + 2 * q, n_pad - 2); // rows past the (sub-)matrix: a valid address of the column, results masked
        cp_async16(s_P + kk * S8_LDP + 2 * q, A + (size_t)(k0 + min(kk, nb - 1)) * lda + r);
    }
}
// transposed inverse of a diagonal block ([k][row], 32 x 32 doubles) -> padded rows
__device__ __forceinline__ void s8_prefetch_tinv(double* s_T, const double* __restrict__ Tinv)
{
    for (int t = threadIdx.x; t < SB * SB / 2; t += S8_THREADS) {
        const int kk = t / (SB / 2), q = t % (SB / 2);
        cp_async16(s_T + kk * S8_LDT + 2 * q, Tinv + kk * SB + 2 * q);
    }
}

__global__ void __launch_bounds__(S8_THREADS) k_solve_slab8(const double* __restrict__ A, int lda, int n, int N,
                                                            const int* __restrict__ perm, const float* __restrict__ rest,
                                                            const float* __restrict__ deform, int F,
                                                            const double* __restrict__ Tinv, double* __restrict__ W, int ldw,
                                                            const fd_tc_pack_args pk, int mode)
{
    // mode 0: L sweep then U sweep (a whole system); 1: L sweep only; 2: U sweep only -- the blocked solve of systems too
    // large for one slab runs this kernel on 1024-row diagonal panels (A, Tinv, W offset to the panel; perm == NULL)
    extern __shared__ __align__(16) double s8_smem[];
    __shared__ double s_red[32][S8_RC + 1];
    __shared__ float s_scale[S8_RC];
    const int nblk = (n + SB - 1) / SB, n_pad = nblk * SB;
    double* s_B = s8_smem;                           // [n_pad][8]
    double* s_P = s_B + (size_t)n_pad * S8_RC;       // [2][32][S8_LDP]
    double* s_T = s_P + 2 * S8_CHUNK_DOUBLES;        // [2][32][S8_LDT]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int fr = lane >> 2, fk = lane & 3;         // fragment coordinates: row (or column) group, k within the k4 step
    const int c0 = blockIdx.x * S8_RC;
    const int nrhs = 3 * F;
    if (c0 >= ldw) { // only with the fused pack: padded columns of the tensor path's tables (zero weights, unit scale)
        for (int cc = tid; cc < S8_RC; cc += S8_THREADS)
            if (c0 + cc < pk.ncol_pad) {
                pk.scale[c0 + cc] = 1.0f;
                pk.unscale[c0 + cc] = (float)ldexp(1.0, -pk.phi_shift);
            }
        for (int t = tid; t < S8_RC * pk.Kpad; t += S8_THREADS) {
            const int c = c0 + t / pk.Kpad, k = t % pk.Kpad;
            if (c < pk.ncol_pad) {
                ((__half*)pk.wt_hi)[(size_t)c * pk.Kpad + k] = __float2half_rn(0.f);
                ((__half*)pk.wt_lo)[(size_t)c * pk.Kpad + k] = __float2half_rn(0.f);
            }
        }
        return;
    }
    const int s_begin = mode == 2 ? nblk : 0, s_end = mode == 1 ? nblk : 2 * nblk;
    { // the inverse of the first step's diagonal block
        const int blk0 = s_begin < nblk ? s_begin : 2 * nblk - 1 - s_begin;
        s8_prefetch_tinv(s_T, Tinv + ((size_t)blk0 * 2 + (s_begin < nblk ? 0 : 1)) * SB * SB);
    }
    cp_async_commit();
    // right-hand sides, permuted: delta subtracted in FP32 then widened (SOP_FaceDeform.cpp:276-284); pad rows are zero
    for (int t = tid; t < n_pad * S8_RC; t += S8_THREADS) {
        const int i = t / S8_RC, c = c0 + (t % S8_RC);
        double v = 0.0;
        if (i < n) {
            const int src = perm ? perm[i] : i;
            if (!deform) { // right-hand sides prebuilt in W (the null-space path, fd_nullspace.cu)
                if (c < ldw) v = W[(size_t)src * ldw + c];
            } else if (c < nrhs && src < N) {
                const int f = c / 3, k = c - 3 * f;
                v = (double)(deform[((size_t)f * N + src) * 3 + k] - rest[3 * src + k]);
            }
        }
        s_B[t] = v;
    }
    int tbuf = 0, pbuf = 0;
    // 2 * nblk block steps: L sweep down, then U sweep up (or one of the two: mode)
    for (int step = s_begin; step < s_end; ++step) {
        const bool lower = step < nblk;
        const int blk = lower ? step : 2 * nblk - 1 - step;
        const int k0 = blk * SB, nb = min(SB, n - k0);
        const int row_lo = lower ? k0 + SB : 0, row_hi = lower ? n : k0; // rows outside the block that depend on it
        const int nchunks = row_hi > row_lo ? (row_hi - row_lo + S8_CH - 1) / S8_CH : 0;
        // prefetch: first panel chunk of this step and the inverse of the next step's block (one cp.async group)
        if (nchunks > 0) s8_prefetch_chunk(s_P + pbuf * S8_CHUNK_DOUBLES, A, lda, k0, nb, row_lo, n_pad);
        if (step + 1 < s_end) {
            const int nblk_next = step + 1 < nblk ? step + 1 : 2 * nblk - 2 - step;
            s8_prefetch_tinv(s_T + (tbuf ^ 1) * S8_TINV_DOUBLES, Tinv + ((size_t)nblk_next * 2 + (step + 1 < nblk ? 0 : 1)) * SB * SB);
        }
        cp_async_commit();
        cp_async_wait<1>(); // everything but the group just issued: this step's inverse has landed
        __syncthreads();    // ... for every thread; the slab updates of the previous step are visible
        // ---- X_k = T_kk^-1 B_k: warps 0..3, one m8 tile each
        double x0 = 0.0, x1 = 0.0;
        if (warp < 4) {
            const double* T = s_T + tbuf * S8_TINV_DOUBLES + 8 * warp + fr;
#pragma unroll
            for (int s4 = 0; s4 < SB / 4; ++s4) {
                const double av = T[(4 * s4 + fk) * S8_LDT];
                const double bv = s_B[(size_t)(k0 + 4 * s4 + fk) * S8_RC + fr];
                dmma8(x0, x1, av, bv);
            }
            asm volatile("bar.sync 1, 128;" ::: "memory"); // the four warps have read B_k
            *reinterpret_cast<double2*>(s_B + (size_t)(k0 + 8 * warp + fr) * S8_RC + 2 * fk) = make_double2(x0, x1);
        }
        __syncthreads();
        if (nchunks > 0) {
            // X fragments of the whole block: 8 doubles per lane, reused by every row tile of the step
            double xf[SB / 4];
#pragma unroll
            for (int s4 = 0; s4 < SB / 4; ++s4) xf[s4] = s_B[(size_t)(k0 + 4 * s4 + fk) * S8_RC + fr];
            for (int ch = 0; ch < nchunks; ++ch) {
                const int r0 = row_lo + ch * S8_CH;
                if (ch + 1 < nchunks) {
                    s8_prefetch_chunk(s_P + (pbuf ^ 1) * S8_CHUNK_DOUBLES, A, lda, k0, nb, r0 + S8_CH, n_pad);
                    cp_async_commit();
                    cp_async_wait<1>();
                } else {
                    cp_async_wait<0>();
                }
                __syncthreads();
                const double* P = s_P + pbuf * S8_CHUNK_DOUBLES;
#pragma unroll
                for (int tt = 0; tt < S8_CH / 8 / (S8_THREADS / 32); ++tt) {
                    const int tile = warp + (S8_THREADS / 32) * tt;
                    const int i0 = r0 + 8 * tile;
                    if (i0 < row_hi) {
                        double a0 = 0.0, a1 = 0.0;
                        const double* Pt = P + 8 * tile + fr;
#pragma unroll
                        for (int s4 = 0; s4 < SB / 4; ++s4) dmma8(a0, a1, Pt[(4 * s4 + fk) * S8_LDP], xf[s4]);
                        if (i0 + fr < row_hi) {
                            double2* dst = reinterpret_cast<double2*>(s_B + (size_t)(i0 + fr) * S8_RC + 2 * fk);
                            double2 v = *dst;
                            v.x -= a0;
                            v.y -= a1;
                            *dst = v;
                        }
                    }
                }
                __syncthreads(); // chunk buffer free for the prefetch after next
                pbuf ^= 1;
            }
        }
        tbuf ^= 1;
    }
    cp_async_wait<0>();
    __syncthreads();
    for (int t = tid; t < n * S8_RC; t += S8_THREADS) {
        const int i = t / S8_RC, c = c0 + (t % S8_RC);
        if (c < ldw) W[(size_t)i * ldw + c] = s_B[t];
    }
    if (!pk.enabled) return;
    // ---- fused pack for the tensor-core evaluation (the arithmetic of tc::k_tc_colscale + tc::k_tc_pack, fd_eval_tc.cu):
    // this CTA holds its 8 columns for every row, so the per-column power-of-two scale and the transposed FP16 hi/lo
    // tiles W^T[c][k] come straight out of shared memory -- two launches and a pass over W less per solve.
    const int N_ = pk.N, np_ = pk.np;
    const float n0 = pk.norm[0], n1 = pk.norm[1], n2 = pk.norm[2], n3 = pk.norm[3];
    auto weff = [&](int k, int cc) -> double { // effective weight of row k (affine rows in normalised coordinates)
        if (k < N_) return s_B[(size_t)k * S8_RC + cc];
        if (np_ == 0) return 0.0;
        if (k == N_) {
            double v = s_B[(size_t)N_ * S8_RC + cc];
            if (np_ == 4) {
                v += s_B[(size_t)(N_ + 1) * S8_RC + cc] * (double)n0;
                v += s_B[(size_t)(N_ + 2) * S8_RC + cc] * (double)n1;
                v += s_B[(size_t)(N_ + 3) * S8_RC + cc] * (double)n2;
            }
            return v;
        }
        if (np_ == 4 && k <= N_ + 3) return s_B[(size_t)k * S8_RC + cc] / (double)n3;
        return 0.0;
    };
    {
        const int cc = tid % S8_RC, kg = tid / S8_RC; // 32 row groups x 8 columns
        double mx = 0.0, chk = 0.0;
        if (c0 + cc < pk.ncol)
            for (int k = kg; k < N_ + 4; k += S8_THREADS / S8_RC) {
                const double w = weff(k, cc);
                mx = fmax(mx, fabs(w));
                chk = fma(w, 0.0, chk);
            }
        if (chk != 0.0) atomicExch(&pk.flags[FD_FLAG_NONFINITE], 1); // NaN / Inf weights -> terminationtype -3
        s_red[kg][cc] = mx;
    }
    __syncthreads();
    if (tid < S8_RC) {
        double mx = 0.0;
        for (int g = 0; g < S8_THREADS / S8_RC; ++g) mx = fmax(mx, s_red[g][tid]);
        int e = 0;
        if (mx > 0.0 && isfinite(mx)) {
            frexp(mx, &e);
            e = max(-60, min(60, 14 - e)); // mx * 2^e in [8192, 16384)
        }
        s_scale[tid] = (float)ldexp(1.0, e);
        if (c0 + tid < pk.ncol_pad) {
            pk.scale[c0 + tid] = s_scale[tid];
            pk.unscale[c0 + tid] = (float)ldexp(1.0, -e - pk.phi_shift);
        }
    }
    __syncthreads();
    for (int cc = 0; cc < S8_RC; ++cc) {
        const int c = c0 + cc;
        if (c >= pk.ncol_pad) break;
        __half* hi = (__half*)pk.wt_hi + (size_t)c * pk.Kpad;
        __half* lo = (__half*)pk.wt_lo + (size_t)c * pk.Kpad;
        const double sc = (double)s_scale[cc];
        for (int k = tid; k < pk.Kpad; k += S8_THREADS) {
            float v = 0.f;
            if (c < pk.ncol) v = (float)(weff(k, cc) * sc);
            const __half h = __float2half_rn(v);
            hi[k] = h;
            lo[k] = __float2half_rn(v - __half2float(h));
        }
    }
}

// weights -> evaluation tables.  FP32: centre table (cx, cy, cz, kernel parameter) and weights n x ldw32;
// FP64 centre table when the evaluation runs in double.  Flags non-finite weights.
__global__ void __launch_bounds__(256) k_pack_tables(const float* __restrict__ rest, const double* __restrict__ radii,
                                                     int N, int Npad, int kernel, float4* __restrict__ ctab32,
                                                     float* __restrict__ ctab_pair, double4* __restrict__ ctab64,
                                                     double4* __restrict__ ctabx, double* __restrict__ cscx)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= Npad) return;
    // the tensor path reads pairs of centres interleaved as (x0 x1 y0 y1 z0 z1 prm0 prm1) for the packed FP32 pipe
    float* pr = ctab_pair + (j >> 1) * 8 + (j & 1);
    if (j >= N) { // padding of the centre table (multiquadric parameter 1 keeps the padded basis finite)
        const float pad = kernel == FD_KERNEL_MULTIQUADRIC ? 1.f : 0.f;
        ctab32[j] = make_float4(0.f, 0.f, 0.f, pad);
        pr[0] = 0.f, pr[2] = 0.f, pr[4] = 0.f, pr[6] = pad;
        if (ctab64) ctab64[j] = make_double4(0.0, 0.0, 0.0, 0.0);
        if (ctabx) {
            ctabx[j] = make_double4(0.0, 0.0, 0.0, 0.0);
            cscx[j] = 0.0;
        }
        return;
    }
    const double R = radii[j];
    double prm;
    if (kernel == FD_KERNEL_GAUSSIAN) prm = -1.0 / (R * R);
    else if (kernel == FD_KERNEL_MULTIQUADRIC) prm = R * R;
    else prm = 0.0;
    const float x = rest[3 * j], y = rest[3 * j + 1], z = rest[3 * j + 2];
    // FP32 Gaussian evaluates ex2(r2 * (-log2(e) / R^2))
    const double prm32 = kernel == FD_KERNEL_GAUSSIAN ? prm * 1.4426950408889634074 : prm;
    ctab32[j] = make_float4(x, y, z, (float)prm32);
    pr[0] = x, pr[2] = y, pr[4] = z, pr[6] = (float)prm32;
    if (ctabx) { // the exact-digit tensor-core kernel (Gaussian): t = log2 phi = sc |p - c|^2 expanded around centre 0
        const double sc = prm * 1.4426950408889634074;
        const double cx = (double)x - (double)rest[0], cy = (double)y - (double)rest[1], cz = (double)z - (double)rest[2];
        ctabx[j] = make_double4(-2.0 * sc * cx, -2.0 * sc * cy, -2.0 * sc * cz, sc * (cx * cx + cy * cy + cz * cz));
        cscx[j] = sc;
    }
    if (ctab64) {
        if (kernel == FD_KERNEL_GAUSSIAN) {
            ctab64[j] = make_double4((double)x, (double)y, (double)z, prm);
        } else { // expanded-distance form of k_eval_f64: -2 (c - o), |c - o|^2 + kernel parameter, o = centre 0
            const double cx = (double)x - (double)rest[0], cy = (double)y - (double)rest[1], cz = (double)z - (double)rest[2];
            ctab64[j] = make_double4(-2.0 * cx, -2.0 * cy, -2.0 * cz, cx * cx + cy * cy + cz * cz + prm);
        }
    }
}

__global__ void __launch_bounds__(256) k_pack_weights(const double* __restrict__ W, int n, int ldw, int nrhs,
                                                      float* __restrict__ W32, int ldw32, int* __restrict__ flags)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y;
    if (c >= ldw32) return;
    double v = 0.0;
    if (c < nrhs) {
        v = W[(size_t)i * ldw + c];
        if (!isfinite(v)) atomicExch(&flags[FD_FLAG_NONFINITE], 1);
    }
    W32[(size_t)i * ldw32 + c] = (float)v;
}

} // namespace

// per-device function attributes (fd_ctx_create)
cudaError_t fd_solve_setup(fd_ctx* ctx)
{
    (void)ctx;
    cudaError_t e = cudaFuncSetAttribute(k_solve_slab8, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_solve_slab, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_panel_gemm, cudaFuncAttributeMaxDynamicSharedMemorySize, PG_SMEM_BYTES);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_gemm_sub, cudaFuncAttributeMaxDynamicSharedMemorySize, PG_SMEM_BYTES);
    return e;
}

cudaError_t fd_launch_solve(fd_ctx* ctx, fd_model* m, const float* d_deform, int F)
{
    m->tc_packed_by_solve = false;

    cudaStream_t s = ctx->stream;
    const int n = m->n, nrhs = 3 * F, ldw = m->ldw;
    {
        cudaError_t e_inv;
        if (fd_try_inverse_solve(ctx, m, m->d_A, m->lda, n, m->d_perm, m->d_Tinv, d_deform, F, m->d_W, ldw, false, &e_inv))
            return e_inv;
    }
    {
        // FP64 tensor-pipe slab solve: 8 right-hand sides per CTA
        const int n_pad = fd_round_up(n, SB);
        const size_t bytes8 = ((size_t)n_pad * S8_RC + 2 * S8_CHUNK_DOUBLES + 2 * S8_TINV_DOUBLES) * sizeof(double);
        if (bytes8 <= 220 * 1024 && !ctx->dbg.solve_dfma) {
            fd_tc_pack_args pk;
            pk.enabled = 0;
            if (m->use_tc && !ctx->dbg.no_fused_pack) fd_tc_pack_args_fill(m, &pk);
            const int cols = pk.enabled ? max(ldw, pk.ncol_pad) : ldw;
            k_solve_slab8<<<(cols + S8_RC - 1) / S8_RC, S8_THREADS, bytes8, s>>>(m->d_A, m->lda, n, m->N, m->d_perm, m->d_rest,
                                                                                d_deform, F, m->d_Tinv, m->d_W, ldw, pk, 0);
            ctx->launches += 1;
            m->tc_packed_by_solve = pk.enabled != 0;
            return cudaGetLastError();
        }
    }
    const size_t slab_bytes = (size_t)n * RC * sizeof(double);
    if (slab_bytes <= 200 * 1024 && (nrhs >= 2 * RC || n <= 1024)) {
        // one launch: every CTA solves its 16 right-hand sides start to finish out of shared memory
        k_solve_slab<<<(ldw + RC - 1) / RC, SLAB_THREADS, slab_bytes, s>>>(m->d_A, m->lda, n, m->N, m->d_perm, m->d_rest,
                                                                          d_deform, F, m->d_Tinv, m->d_W, ldw);
        ctx->launches += 1;
        return cudaGetLastError();
    }
    {
        dim3 grid((ldw + 255) / 256, n);
        k_build_rhs<<<grid, 256, 0, s>>>(m->d_rest, d_deform, m->d_perm, m->N, n, F, m->d_W, ldw);
        ctx->launches += 1;
    }
    return fd_launch_solve_prebuilt(ctx, m, nrhs);
}

// forward / backward sweeps on right-hand sides that already sit (row-permuted) in m->d_W
cudaError_t fd_launch_solve_prebuilt(fd_ctx* ctx, fd_model* m, int nrhs)
{
    return fd_launch_solve_sub(ctx, m->d_A, m->lda, m->n, nullptr, m->d_Tinv, m->d_W, m->ldw, nrhs);
}

// the same on an explicit system: LU factors A (n x n, column stride lda), inverted diagonal blocks Tinv, right-hand
// sides W (row stride ldw) already in the LU's row order.  perm != NULL (an identity for the no-pivot LU) allows the
// one-launch slab solve when the slab fits.
cudaError_t fd_launch_solve_sub(fd_ctx* ctx, const double* d_A, int lda, int n, const int* d_perm, const double* d_Tinv,
                                double* d_W, int ldw, int nrhs)
{
    cudaStream_t s = ctx->stream;
    if (d_perm) {
        const int n_pad = fd_round_up(n, SB);
        const size_t bytes8 = ((size_t)n_pad * S8_RC + 2 * S8_CHUNK_DOUBLES + 2 * S8_TINV_DOUBLES) * sizeof(double);
        if (bytes8 <= 220 * 1024) {
            fd_tc_pack_args pk;
            pk.enabled = 0;
            k_solve_slab8<<<(ldw + S8_RC - 1) / S8_RC, S8_THREADS, bytes8, s>>>(d_A, lda, n, n, d_perm, nullptr, nullptr,
                                                                              nrhs / 3, d_Tinv, d_W, ldw, pk, 0);
            ctx->launches += 1;
            return cudaGetLastError();
        }
    }
    // beyond the slab: a handful of right-hand sides take one launch per block step (8.0 ms against 9.0 ms for the
    // two-launch rank-32 sweeps at n = 8196; where the slab fits, its single long-running CTA is still faster)
    if (nrhs <= FEW_MAX && n >= 512 && !ctx->dbg.no_few_rhs) {
        double* d_Y = nullptr; // forward-substituted right-hand sides, n x 8
        cudaError_t e = cudaMallocAsync((void**)&d_Y, (size_t)n * FEW_MAX * sizeof(double), s);
        if (e != cudaSuccess) return e;
        if (ctx->dbg.poison) cudaMemsetAsync(d_Y, 0xFF, (size_t)n * FEW_MAX * sizeof(double), s);
        for (int k0 = 0; k0 < n; k0 += SB) { // L y = b
            const int nb = min(SB, n - k0), rows = n - k0 - nb;
            k_few_step<true><<<max(1, (rows + 63) / 64), 256, 0, s>>>(d_A, lda, n, k0, nb, d_Tinv + (size_t)(k0 / SB) * 2 * SB * SB,
                                                                      d_W, ldw, d_Y, FEW_MAX, d_W, ldw, nrhs);
        }
        for (int k0 = (n - 1) / SB * SB; k0 >= 0; k0 -= SB) { // U x = y
            const int nb = min(SB, n - k0);
            k_few_step<false><<<max(1, (k0 + 63) / 64), 256, 0, s>>>(d_A, lda, n, k0, nb,
                                                                     d_Tinv + ((size_t)(k0 / SB) * 2 + 1) * SB * SB, d_Y, FEW_MAX,
                                                                     d_W, ldw, d_Y, FEW_MAX, nrhs);
        }
        ctx->launches += 2 * ((n + SB - 1) / SB);
        e = cudaGetLastError();
        cudaFreeAsync(d_Y, s);
        return e;
    }
    // Blocked sweeps for systems whose slab does not fit in shared memory: diagonal panels of 1024 rows.  Inside a panel
    // the slab kernel runs its L (or U) sweep for every right-hand side (one launch, 8 columns per CTA, the panel's
    // triangle streamed from L2, DMMA block products); the rows outside the panel take one rank-1024 update on the FP64
    // tensor pipe (k_panel_gemm).  n = 4100: 5 + 5 slab launches and 4 + 4 GEMMs instead of ~540 block-step launches.
    constexpr int PB = 1024;
    const size_t slab_bytes = ((size_t)PB * S8_RC + 2 * S8_CHUNK_DOUBLES + 2 * S8_TINV_DOUBLES) * sizeof(double);
    fd_tc_pack_args pk0;
    pk0.enabled = 0;
    const int slab_grid = (ldw + S8_RC - 1) / S8_RC;
    for (int p0 = 0; p0 < n; p0 += PB) { // L y = P b
        const int pe = min(p0 + PB, n);
        k_solve_slab8<<<slab_grid, S8_THREADS, slab_bytes, s>>>(d_A + (size_t)p0 * lda + p0, lda, pe - p0, pe - p0, nullptr, nullptr,
                                                              nullptr, nrhs / 3, d_Tinv + (size_t)(p0 / SB) * 2 * SB * SB,
                                                              d_W + (size_t)p0 * ldw, ldw, pk0, 1);
        ctx->launches += 1;
        if (pe < n) {
            dim3 grid((nrhs + PG_TN - 1) / PG_TN, (n - pe + PG_TM - 1) / PG_TM);
            k_panel_gemm<<<grid, 256, PG_SMEM_BYTES, s>>>(d_A, lda, pe, n, p0, pe - p0, d_W, ldw, nrhs);
            ctx->launches += 1;
        }
    }
    for (int p0 = (n - 1) / PB * PB; p0 >= 0; p0 -= PB) { // U x = y
        const int pe = min(p0 + PB, n);
        k_solve_slab8<<<slab_grid, S8_THREADS, slab_bytes, s>>>(d_A + (size_t)p0 * lda + p0, lda, pe - p0, pe - p0, nullptr, nullptr,
                                                              nullptr, nrhs / 3, d_Tinv + (size_t)(p0 / SB) * 2 * SB * SB,
                                                              d_W + (size_t)p0 * ldw, ldw, pk0, 2);
        ctx->launches += 1;
        if (p0 > 0) {
            dim3 grid((nrhs + PG_TN - 1) / PG_TN, (p0 + PG_TM - 1) / PG_TM);
            k_panel_gemm<<<grid, 256, PG_SMEM_BYTES, s>>>(d_A, lda, 0, p0, p0, pe - p0, d_W, ldw, nrhs);
            ctx->launches += 1;
        }
    }
    return cudaGetLastError();
}

// Per-cook fast path (see fd_model::d_inv).  The inverse is n_f solves against the identity through the ordinary sweeps,
// so X = A^-1 includes the row interchanges and applies to right-hand sides in their original order.
bool fd_try_inverse_solve(fd_ctx* ctx, fd_model* m, const double* d_A, int lda, int n_f, const int* d_perm, const double* d_Tinv,
                          const float* d_deform, int F, double* d_W, int ldw, bool rhs_in_W, cudaError_t* err)
{
    *err = cudaSuccess;
    const int nrhs = 3 * F;
    if (nrhs > 8 || ctx->dbg.no_inverse || n_f < 256) return false; // small systems: the slab solve is already a few microseconds
    cudaStream_t s = ctx->stream;
    m->small_solves += 1;
    if (!m->d_inv) {
        if (m->small_solves < 2) return false; // a single solve per fit (fit + solve + eval per step) never pays for the inverse
        const int ld = fd_round_up(n_f, 4);
        cudaError_t e = cudaMallocAsync((void**)&m->d_inv, (size_t)n_f * ld * sizeof(double), s);
        if (e == cudaSuccess) e = cudaMallocAsync((void**)&m->d_inv_rhs, (size_t)n_f * 8 * sizeof(double), s);
        if (e != cudaSuccess) { // no memory for the inverse: keep solving through the sweeps
            cudaGetLastError();
            if (m->d_inv) { cudaFreeAsync(m->d_inv, s); m->d_inv = nullptr; }
            m->small_solves = -1000000;
            return false;
        }
        m->ld_inv = ld;
        if (ctx->dbg.poison) {
            cudaMemsetAsync(m->d_inv, 0xFF, (size_t)n_f * ld * sizeof(double), s);
            cudaMemsetAsync(m->d_inv_rhs, 0xFF, (size_t)n_f * 8 * sizeof(double), s);
        }
        const int n_pad = fd_round_up(n_f, SB);
        const bool slab = ((size_t)n_pad * S8_RC + 2 * S8_CHUNK_DOUBLES + 2 * S8_TINV_DOUBLES) * sizeof(double) <= 220 * 1024;
        dim3 grid((ld + 255) / 256, n_f);
        // the slab kernel gathers its rows through perm itself; the blocked sweeps expect the rows already interchanged
        k_identity_rhs<<<grid, 256, 0, s>>>(m->d_inv, n_f, ld, slab ? nullptr : d_perm);
        ctx->launches += 1;
        e = fd_launch_solve_sub(ctx, d_A, lda, n_f, slab ? d_perm : nullptr, d_Tinv, m->d_inv, ld, n_f);
        if (e != cudaSuccess) { *err = e; return true; }
    }
    if (rhs_in_W) {
        k_gather_rhs8<<<(n_f * 8 + 255) / 256, 256, 0, s>>>(d_W, ldw, n_f, nrhs, m->d_inv_rhs);
    } else {
        dim3 grid(1, n_f);
        k_build_rhs<<<grid, 8, 0, s>>>(m->d_rest, d_deform, nullptr, m->N, n_f, F, m->d_inv_rhs, 8);
    }
    const int ia_grid = (n_f + IA_ROWS - 1) / IA_ROWS;
    if (nrhs <= 3) k_inv_apply<3><<<ia_grid, 256, 0, s>>>(m->d_inv, m->ld_inv, n_f, m->d_inv_rhs, d_W, ldw);
    else if (nrhs <= 6) k_inv_apply<6><<<ia_grid, 256, 0, s>>>(m->d_inv, m->ld_inv, n_f, m->d_inv_rhs, d_W, ldw);
    else k_inv_apply<8><<<ia_grid, 256, 0, s>>>(m->d_inv, m->ld_inv, n_f, m->d_inv_rhs, d_W, ldw);
    ctx->launches += 2;
    *err = cudaGetLastError();
    return true;
}

// the tables that depend on the centres and radii only (built once per fit; receivers build them when the radii
// have arrived): centre tables of every evaluation kernel + the bounding-box normalisation of the tensor path
cudaError_t fd_launch_pack_tables(fd_ctx* ctx, fd_model* m)
{
    cudaStream_t s = ctx->stream;
    const int npad = fd_tc_kpad(m->N);
    k_pack_tables<<<(npad + 255) / 256, 256, 0, s>>>(m->d_rest, m->d_radii, m->N, npad, m->prm.kernel, m->d_ctab32,
                                                    reinterpret_cast<float*>(m->d_ctab_pair), m->d_ctab64, m->d_ctab_tcx, m->d_csc_tcx);
    ctx->launches += 1;
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = fd_launch_tc_norm(ctx, m);
    m->tables_packed = e == cudaSuccess;
    return e;
}

// per solve: the weight tables of the evaluation kernel that will run (FP32 rows for the FMA/SFU kernel, or the
// column-scaled FP16 hi/lo tiles of the tensor path; both flag non-finite weights)
cudaError_t fd_launch_pack(fd_ctx* ctx, fd_model* m)
{
    cudaStream_t s = ctx->stream;
    cudaError_t e = cudaSuccess;
    if (!m->tables_packed || m->receiver) e = fd_launch_pack_tables(ctx, m);
    if (e != cudaSuccess) return e;
    if (m->use_tcx) { // the exact-digit tensor-core kernel while its (small) error bound holds, else FP64: settled on the device
        e = fd_launch_pack_tcx(ctx, m);
        if (e == cudaSuccess) e = fd_launch_cancel_select(ctx, m, 0, 0, 1, 0);
        return e;
    }
    // FD_EVAL_AUTO (Gaussian) keeps every FP32 candidate ready: the choice is made on the device after this pack
    const bool want_simt = !m->use_tc || (m->auto_sel && m->prm.eval_path != FD_PATH_TENSOR);
    if (m->use_tc) e = fd_launch_pack_tc(ctx, m);
    if (e == cudaSuccess && want_simt) {
        dim3 grid((m->ldw32 + 255) / 256, m->n);
        k_pack_weights<<<grid, 256, 0, s>>>(fd_w_src(m), m->n, m->ldw, 3 * m->F, m->d_W32, m->ldw32, m->d_flags);
        ctx->launches += 1;
        e = cudaGetLastError();
    }
    // the cancellation of these weights, and with it the evaluation kernel of FD_EVAL_AUTO (Gaussian; fd_eval64.cu)
    if (e == cudaSuccess && !m->eval64 && m->prm.kernel == FD_KERNEL_GAUSSIAN)
        e = fd_launch_cancel_select(ctx, m, m->use_tc ? 1 : 0, want_simt ? 1 : 0, 0,
                                    m->auto_sel ? 0 : (m->use_tc ? FD_SEL_TENSOR : FD_SEL_SIMT));
    return e;
}

// d_W <- d_W_src (peer memory), only when the FP64 evaluation will run: statically (eval64) or by the device-side choice
__global__ void __launch_bounds__(256) k_pull_weights(const double2* __restrict__ src, double2* __restrict__ dst, size_t count2,
                                                      const int* __restrict__ sel)
{
    if (sel && *sel != FD_SEL_FP64) return;
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < count2; i += (size_t)gridDim.x * 256) dst[i] = src[i];
}

cudaError_t fd_launch_pull_weights(fd_ctx* ctx, fd_model* m)
{
    if (!m->d_W_src || m->d_W_src == m->d_W) return cudaSuccess;
    if (!m->eval64 && !m->auto_sel) return cudaSuccess; // FP32 evaluation only: the local tables are all it reads
    const size_t count2 = (size_t)m->n * m->ldw / 2;    // ldw is a multiple of 4
    const int grid = (int)((count2 + 255) / 256 < (size_t)ctx->sm_count * 4 ? (count2 + 255) / 256 : (size_t)ctx->sm_count * 4);
    k_pull_weights<<<grid > 0 ? grid : 1, 256, 0, ctx->stream>>>(reinterpret_cast<const double2*>(m->d_W_src),
                                                               reinterpret_cast<double2*>(m->d_W), count2,
                                                               m->eval64 ? nullptr : m->d_sel);
    ctx->launches += 1;
    return cudaGetLastError();
}

// after the LU: inverted diagonal blocks for the slab solve
cudaError_t fd_launch_invdiag(fd_ctx* ctx, fd_model* m)
{
    const int nblk = (m->n + SB - 1) / SB;
    k_lu_invdiag<<<nblk, 64, 0, ctx->stream>>>(m->d_A, m->lda, m->n, m->d_Tinv);
    ctx->launches += 1;
    return cudaGetLastError();
}

// ---- "ALGLIB v1 like" layered fit (fd_params.fidelity = FD_FIDELITY_ALGLIB_V1; orchestration in fd_api.cu) --------------
// one thread per right-hand-side column: R = delta - P v with v the least-squares polynomial of the deltas on
// [1 x y z] (np = 4), their mean (np = 1) or nothing -- rbfsetlinterm / constterm / zeroterm fitted FIRST
// (SOP_FaceDeform.cpp:351-361 [recollection of ALGLIB v1: two-stage])
__global__ void __launch_bounds__(128) k_v1_rhs_poly(const float* __restrict__ rest, const float* __restrict__ deform, int N,
                                                     int F, int np, double* __restrict__ R, double* __restrict__ V, int ldw,
                                                     int* __restrict__ flags)
{
    const int c = blockIdx.x * 128 + threadIdx.x;
    if (c >= ldw) return;
    const bool live = c < 3 * F;
    const int f = c / 3, k = c - 3 * f;
    auto delta = [&](int i) -> double {
        return live ? (double)(deform[((size_t)f * N + i) * 3 + k] - rest[3 * i + k]) : 0.0; // FP32 subtract, :276-284
    };
    double y[4] = {0, 0, 0, 0}, G[4][4] = {};
    for (int i = 0; i < N && np > 0; ++i) {
        const double p[4] = {1.0, (double)rest[3 * i], (double)rest[3 * i + 1], (double)rest[3 * i + 2]};
        const double d = delta(i);
        for (int a = 0; a < np; ++a) {
            y[a] += p[a] * d;
            for (int b = 0; b < np; ++b) G[a][b] += p[a] * p[b];
        }
    }
    bool bad = false;
    for (int q = 0; q < np; ++q) { // elimination without pivoting on the Gram matrix (the oracle does the same)
        const double dg = G[q][q];
        if (!(dg > 0.0)) { bad = true; break; }
        for (int r = q + 1; r < np; ++r) {
            const double l = G[r][q] / dg;
            for (int t = q; t < np; ++t) G[r][t] -= l * G[q][t];
            y[r] -= l * y[q];
        }
    }
    if (bad) {
        atomicExch(&flags[FD_FLAG_SINGULAR], 1);
        for (int a = 0; a < np; ++a) y[a] = 0.0;
    } else {
        for (int q = np - 1; q >= 0; --q) {
            double sacc = y[q];
            for (int t = q + 1; t < np; ++t) sacc -= G[q][t] * y[t];
            y[q] = sacc / G[q][q];
        }
    }
    for (int a = 0; a < np; ++a) V[(size_t)a * ldw + c] = y[a];
    for (int i = 0; i < N; ++i) {
        double t = np > 0 ? y[0] : 0.0;
        if (np == 4) t += y[1] * (double)rest[3 * i] + y[2] * (double)rest[3 * i + 1] + y[3] * (double)rest[3 * i + 2];
        R[(size_t)i * ldw + c] = delta(i) - t;
    }
}

// W[i][c] = R[perm[i]][c]: right-hand sides in the row order of the layer's LU
__global__ void __launch_bounds__(256) k_v1_gather(const double* __restrict__ R, const int* __restrict__ perm, int ldw,
                                                   double* __restrict__ W)
{
    const int c = blockIdx.x * 256 + threadIdx.x;
    const int i = blockIdx.y;
    if (c < ldw) W[(size_t)i * ldw + c] = R[(size_t)perm[i] * ldw + c];
}

cudaError_t fd_launch_v1_rhs_poly(fd_ctx* ctx, const float* d_rest, const float* d_deform, int N, int F, int np, double* d_R,
                                  double* d_V, int ldw, int* d_flags)
{
    k_v1_rhs_poly<<<(ldw + 127) / 128, 128, 0, ctx->stream>>>(d_rest, d_deform, N, F, np, d_R, d_V, ldw, d_flags);
    ctx->launches += 1;
    return cudaGetLastError();
}

cudaError_t fd_launch_v1_gather(fd_ctx* ctx, const double* d_R, const int* d_perm, int N, int ldw, double* d_W)
{
    dim3 grid((ldw + 255) / 256, N);
    k_v1_gather<<<grid, 256, 0, ctx->stream>>>(d_R, d_perm, ldw, d_W);
    ctx->launches += 1;
    return cudaGetLastError();
}

cudaError_t fd_launch_gemm_sub(fd_ctx* ctx, const double* d_A, int lda, int rows, int K, const double* d_X, double* d_C, int ldw,
                               int nrhs)
{
    dim3 grid((nrhs + PG_TN - 1) / PG_TN, (rows + PG_TM - 1) / PG_TM);
    k_gemm_sub<<<grid, 256, PG_SMEM_BYTES, ctx->stream>>>(d_A, lda, rows, K, d_X, d_C, ldw, nrhs);
    ctx->launches += 1;
    return cudaGetLastError();
}
