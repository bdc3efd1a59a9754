// fd_nullspace.cu -- multiquadric / thin-plate systems without pivoting: the null-space transform.
//
// The saddle-point system [[K + lambda I, P], [P^T, 0]] [w; a] = [D; 0] of the conditionally definite kernels
// (reference call site: alglib::rbfbuildmodel, SOP_FaceDeform.cpp:363, in the dense formulation of north_star) is
// symmetric INDEFINITE, so the general path is the pivoted LU -- 2.5-4x slower than the fused no-pivot kernel.
// With P = Q [R; 0] (Householder QR of the N x 4 polynomial block [1 x y z], Q = H_0 H_1 H_2 H_3) every admissible
// w (P^T w = 0) is w = Q [0; z], and the equations split:
//      S z = (Q^T D)[4:],   S = (Q^T K Q)[4:, 4:]            (N - 4) x (N - 4), DEFINITE on the null space:
//                                                            positive for r^2 log r, negative for sqrt(r^2 + R^2)
//      R a = (Q^T D)[:4] - (Q^T K Q)[:4, 4:] z
// so S takes the fused no-pivot LU (fd_factor.cu) and its slab solve.  The transform costs 8 passes over K
// (per reflector: p = tau K v, q = p - (tau/2)(v.p) v, K -= v q^T + q v^T), HBM bound and small beside the LU.
// Conditions (fd_api.cu): uniform radius (symmetric K), linear term.  Smoothing keeps S definite: + lambda I for the
// thin plate, - lambda I for the multiquadric (k_assemble).
#include "fd_internal.h"

namespace {

constexpr int NP = 4;

__device__ __forceinline__ double cta_sum(double v, double* s_red)
{
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = blockDim.x >> 5;
    __syncthreads();
    if (l == 0) s_red[w] = v;
    __syncthreads();
    double t = 0.0;
    for (int i = 0; i < nw; ++i) t += s_red[i]; // every thread sums the same values in the same order
    return t;
}

// Householder QR of P = [1 x y z] (N x 4): V (N x 4 column-major, v_j[j] = 1, zeros above), tau[4], R (4 x 4 row-major
// upper triangle).  One CTA; LAPACK conventions (beta = -sign(alpha) |x|).
__global__ void __launch_bounds__(1024) k_ns_house(const float* __restrict__ rest, int N, double* __restrict__ V,
                                                   double* __restrict__ tau, double* __restrict__ R)
{
    __shared__ double s_red[32];
    double* Pw = V;              // the working copy lives in V (overwritten column by column with the reflectors)
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        Pw[i] = 1.0;
        Pw[N + i] = (double)rest[3 * i];
        Pw[2 * N + i] = (double)rest[3 * i + 1];
        Pw[3 * N + i] = (double)rest[3 * i + 2];
    }
    if (threadIdx.x < NP * NP) R[threadIdx.x] = 0.0;
    __syncthreads();
    for (int j = 0; j < NP; ++j) {
        double* x = Pw + (size_t)j * N;
        double ss = 0.0;
        for (int i = threadIdx.x; i < N; i += blockDim.x)
            if (i > j) ss = fma(x[i], x[i], ss);
        ss = cta_sum(ss, s_red);
        const double alpha = x[j];
        double t = 0.0, scale = 1.0, beta = alpha;
        if (ss != 0.0) {
            beta = sqrt(alpha * alpha + ss);
            if (alpha >= 0.0) beta = -beta;
            t = (beta - alpha) / beta;
            scale = 1.0 / (alpha - beta);
        }
        __syncthreads(); // everyone has read alpha
        for (int i = threadIdx.x; i < N; i += blockDim.x) x[i] = i < j ? 0.0 : (i == j ? 1.0 : x[i] * scale);
        if (threadIdx.x == 0) {
            tau[j] = t;
            R[j * NP + j] = beta;
        }
        __syncthreads();
        for (int c = j + 1; c < NP; ++c) { // apply H_j to the remaining columns
            double* y = Pw + (size_t)c * N;
            double d = 0.0;
            for (int i = threadIdx.x; i < N; i += blockDim.x)
                if (i >= j) d = fma(x[i], y[i], d);
            d = cta_sum(d, s_red) * t;
            for (int i = threadIdx.x; i < N; i += blockDim.x)
                if (i >= j) y[i] = fma(-d, x[i], y[i]);
            __syncthreads();
            if (threadIdx.x == 0) R[j * NP + c] = y[j];
            __syncthreads();
        }
    }
}

// p[c] = tau * sum_r K[r][c] v[r]  (K symmetric: K v read column-wise, coalesced); one warp per column
__global__ void __launch_bounds__(256) k_ns_gemv(const double* __restrict__ K, int lda, int N, const double* __restrict__ v,
                                                 const double* __restrict__ tau_j, double* __restrict__ p)
{
    const int c = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (c >= N) return;
    const double* col = K + (size_t)c * lda;
    double acc = 0.0;
    for (int r = lane; r < N; r += 32) acc = fma(col[r], v[r], acc);
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) p[c] = acc * tau_j[0];
}

// q = p - (tau / 2) (v . p) v   (one CTA)
__global__ void __launch_bounds__(1024) k_ns_q(const double* __restrict__ v, const double* __restrict__ tau_j, int N,
                                               double* __restrict__ p)
{
    __shared__ double s_red[32];
    double d = 0.0;
    for (int i = threadIdx.x; i < N; i += blockDim.x) d = fma(v[i], p[i], d);
    d = cta_sum(d, s_red) * 0.5 * tau_j[0];
    for (int i = threadIdx.x; i < N; i += blockDim.x) p[i] = fma(-d, v[i], p[i]);
}

// K[r][c] -= v[r] q[c] + q[r] v[c]
__global__ void __launch_bounds__(256) k_ns_rank2(double* __restrict__ K, int lda, int N, const double* __restrict__ v,
                                                  const double* __restrict__ q)
{
    const int r = blockIdx.x * 32 + (threadIdx.x & 31);
    const int c0 = blockIdx.y * 32;
    if (r >= N) return;
    const double vr = v[r], qr = q[r];
    for (int cc = threadIdx.x >> 5; cc < 32; cc += 8) {
        const int c = c0 + cc;
        if (c < N) K[(size_t)c * lda + r] -= vr * q[c] + qr * v[c];
    }
}

// D[i][c] = (double)(deform[f][i][k] - rest[i][k]) for i < N (FP32 subtract, SOP_FaceDeform.cpp:276-284), rows N.. zero
__global__ void __launch_bounds__(256) k_ns_rhs(const float* __restrict__ rest, const float* __restrict__ deform, int N, int n,
                                                int F, double* __restrict__ W, int ldw)
{
    const int c = blockIdx.x * 256 + threadIdx.x;
    const int i = blockIdx.y;
    if (c >= ldw) return;
    double v = 0.0;
    if (c < 3 * F && i < N) {
        const int f = c / 3, k = c - 3 * f;
        v = (double)(deform[((size_t)f * N + i) * 3 + k] - rest[3 * i + k]);
    }
    (void)n;
    W[(size_t)i * ldw + c] = v;
}

// W <- H_j W on the rows [j, N) of 8 columns per CTA: s = tau v^T W, W -= v s^T  (threads: 8 columns x 32 row slots)
__global__ void __launch_bounds__(256) k_ns_reflect(const double* __restrict__ v, const double* __restrict__ tau_j, int j, int N,
                                                    double* __restrict__ W, int ldw, int nrhs)
{
    __shared__ double s_part[32][9];
    const int tx = threadIdx.x & 7, ty = threadIdx.x >> 3;
    const int c = blockIdx.x * 8 + tx;
    double acc = 0.0;
    if (c < nrhs)
        for (int i = j + ty; i < N; i += 32) acc = fma(v[i], W[(size_t)i * ldw + c], acc);
    s_part[ty][tx] = acc;
    __syncthreads();
    double s = 0.0;
    for (int g = 0; g < 32; ++g) s += s_part[g][tx];
    s *= tau_j[0];
    if (c < nrhs)
        for (int i = j + ty; i < N; i += 32) W[(size_t)i * ldw + c] = fma(-s, v[i], W[(size_t)i * ldw + c]);
}

// a = R^-1 (D'[:4] - K'[:4, 4:] z): rows 0..3 of W hold D'[:4], rows 4..N-1 hold z; a goes to rows N..N+3, then rows 0..3
// are cleared (w = Q [0; z]).  8 columns per CTA like k_ns_reflect.
__global__ void __launch_bounds__(256) k_ns_poly(const double* __restrict__ K, int lda, int N, const double* __restrict__ R,
                                                 double* __restrict__ W, int ldw, int nrhs)
{
    __shared__ double s_part[32][8][NP + 1];
    const int tx = threadIdx.x & 7, ty = threadIdx.x >> 3;
    const int c = blockIdx.x * 8 + tx;
    double acc[NP] = {0.0, 0.0, 0.0, 0.0};
    if (c < nrhs)
        for (int k = NP + ty; k < N; k += 32) {
            const double z = W[(size_t)k * ldw + c];
            const double* col = K + (size_t)k * lda; // K'[0..3][k]
#pragma unroll
            for (int i = 0; i < NP; ++i) acc[i] = fma(col[i], z, acc[i]);
        }
#pragma unroll
    for (int i = 0; i < NP; ++i) s_part[ty][tx][i] = acc[i];
    __syncthreads();
    if (ty != 0 || c >= nrhs) return;
    double t[NP];
#pragma unroll
    for (int i = 0; i < NP; ++i) {
        double s = 0.0;
        for (int g = 0; g < 32; ++g) s += s_part[g][tx][i];
        t[i] = W[(size_t)i * ldw + c] - s;
    }
    for (int i = NP - 1; i >= 0; --i) { // back substitution with the upper triangle of R
        double s = t[i];
        for (int q = i + 1; q < NP; ++q) s -= R[i * NP + q] * t[q];
        t[i] = s / R[i * NP + i];
    }
#pragma unroll
    for (int i = 0; i < NP; ++i) {
        W[(size_t)(N + i) * ldw + c] = t[i];
        W[(size_t)i * ldw + c] = 0.0;
    }
}

} // namespace

// K (N x N, in m->d_A) -> Q^T K Q; the reflectors, tau and R stay in m->d_ns (N * 4 + 4 + 16 doubles, then 2 N scratch)
cudaError_t fd_launch_ns_transform(fd_ctx* ctx, fd_model* m)
{
    cudaStream_t s = ctx->stream;
    const int N = m->N;
    double* V = m->d_ns;
    double* tau = V + (size_t)N * NP;
    double* R = tau + NP;
    double* p = R + NP * NP;
    k_ns_house<<<1, 1024, 0, s>>>(m->d_rest, N, V, tau, R);
    ctx->launches += 1;
    for (int j = 0; j < NP; ++j) {
        const double* v = V + (size_t)j * N;
        k_ns_gemv<<<(N + 7) / 8, 256, 0, s>>>(m->d_A, m->lda, N, v, tau + j, p);
        k_ns_q<<<1, 1024, 0, s>>>(v, tau + j, N, p);
        dim3 grid((N + 31) / 32, (N + 31) / 32);
        k_ns_rank2<<<grid, 256, 0, s>>>(m->d_A, m->lda, N, v, p);
        ctx->launches += 3;
    }
    return cudaGetLastError();
}

// right-hand sides into m->d_W and D' = Q^T D = H_3 H_2 H_1 H_0 D
cudaError_t fd_launch_ns_rhs(fd_ctx* ctx, fd_model* m, const float* d_deform, int F)
{
    cudaStream_t s = ctx->stream;
    const int N = m->N, nrhs = 3 * F;
    dim3 grid((m->ldw + 255) / 256, m->n);
    k_ns_rhs<<<grid, 256, 0, s>>>(m->d_rest, d_deform, N, m->n, F, m->d_W, m->ldw);
    ctx->launches += 1;
    const double* V = m->d_ns;
    const double* tau = V + (size_t)N * NP;
    for (int j = 0; j < NP; ++j) {
        k_ns_reflect<<<(nrhs + 7) / 8, 256, 0, s>>>(V + (size_t)j * N, tau + j, j, N, m->d_W, m->ldw, nrhs);
        ctx->launches += 1;
    }
    return cudaGetLastError();
}

// after z = S^-1 D'[4:] sits in rows 4..N-1: the polynomial coefficients, then w = Q [0; z] = H_0 H_1 H_2 H_3 [0; z]
cudaError_t fd_launch_ns_finish(fd_ctx* ctx, fd_model* m, int F)
{
    cudaStream_t s = ctx->stream;
    const int N = m->N, nrhs = 3 * F;
    const double* V = m->d_ns;
    const double* tau = V + (size_t)N * NP;
    const double* R = tau + NP;
    k_ns_poly<<<(nrhs + 7) / 8, 256, 0, s>>>(m->d_A, m->lda, N, R, m->d_W, m->ldw, nrhs);
    ctx->launches += 1;
    for (int j = NP - 1; j >= 0; --j) {
        k_ns_reflect<<<(nrhs + 7) / 8, 256, 0, s>>>(V + (size_t)j * N, tau + j, j, N, m->d_W, m->ldw, nrhs);
        ctx->launches += 1;
    }
    return cudaGetLastError();
}
