// fd_dbse.cu -- DirectBSEdit, the SOP's "morph space" post-pass, on the GPU.
//
// Replaces reference src/dbse.cpp behind the same entry points (dbse.hpp:16-22):
//   fd_dbse_init             DirectBSEdit::init            dbse.cpp:9-35   shapes matrix M (3P x S) + Householder QR
//   fd_dbse_compute_weights  DirectBSEdit::computeWeights  dbse.cpp:37-58  w_s = sum_i (pos - rest)_i * QR(i, s)
//   fd_dbse_displace         DirectBSEdit::displaceVector for every point + the SOP's write
//                                                          dbse.cpp:60-75, SOP_FaceDeform.cpp:460-472
//   fd_dbse_get_weights      DirectBSEdit::getWeights      dbse.cpp:77-87
// Eigen's HouseholderQR (dbse.cpp:31) is restated as the unblocked Householder QR with LAPACK's conventions (the
// oracle pins them against LAPACK): packed storage = R above, essential reflector parts below the diagonal -- the
// reference multiplies by this packed matrix (matrixQR(), :53), not by Q, and so does this file.
//
// Everything is HBM-bound streaming over the tall-skinny matrix (3P x S doubles): per column one norm pass, one dot
// pass and one update pass over the trailing columns; reductions go through per-chunk partials summed in a fixed
// order, so results do not depend on the launch.  The per-cook work is two passes: weights (read QR once) and
// displace (read the FP32 copy of M once; FP32, un-fused, columns in order like the reference's loop).
#include <math.h>
#include <new>
#include <string.h>

#include "fd_internal.h"

struct fd_dbse {
    fd_ctx* ctx;
    int64_t P, m, lda; // points, rows = 3P, column stride of QR
    int32_t S;
    int nchunk;        // row chunks of the reductions
    float* d_M32;      // m x S column-major: (shape - rest) in FP32 (exact copy of the reference's double matrix)
    double* d_QR;      // lda x S column-major, packed Householder QR
    double* d_tau;     // S
    double* d_w;       // S weights
    float* d_cw;       // S: clamp((float)(w * 3))
    double* d_part;    // nchunk x max(S, 1) partial sums
    double* d_hh;      // [tau, scale] of the current column
    float* d_a;        // staging: pos
    float* d_b;        // staging: rest
    float* d_o;        // staging: output
    bool computed;
};

namespace {

constexpr int CHUNK = 4096; // rows per reduction chunk
constexpr int CT = 8;       // columns per CTA in the dot / update passes

__device__ __forceinline__ double block_sum(double v, double* s_red)
{
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) s_red[w] = v;
    __syncthreads();
    double t = 0.0;
    if (threadIdx.x == 0)
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += s_red[i];
    return t; // valid in thread 0
}

// M32[s][i] = shapes[s][i] - rest[i] (FP32, dbse.cpp:24), QR = (double) of it (:25-27)
__global__ void __launch_bounds__(256) k_dbse_build(const float* __restrict__ rest, const float* __restrict__ shapes,
                                                    int64_t m, int64_t lda, float* __restrict__ M32, double* __restrict__ QR)
{
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    const int s = blockIdx.y;
    if (i >= m) return;
    const float d = __fsub_rn(shapes[(int64_t)s * m + i], rest[i]);
    M32[(int64_t)s * m + i] = d;
    QR[(int64_t)s * lda + i] = (double)d;
}

// partial sums of x_i^2 over the rows below the diagonal of column j
__global__ void __launch_bounds__(256) k_qr_colnorm(const double* __restrict__ x, int64_t m, int j, double* __restrict__ part)
{
    __shared__ double s_red[8];
    const int64_t r0 = (int64_t)blockIdx.x * CHUNK;
    double acc = 0.0;
    for (int t = threadIdx.x; t < CHUNK; t += 256) {
        const int64_t i = r0 + t;
        if (i > j && i < m) acc = fma(x[i], x[i], acc);
    }
    const double tot = block_sum(acc, s_red);
    if (threadIdx.x == 0) part[blockIdx.x] = tot;
}

// the reflector of column j: beta = -sign(alpha) |x|, tau = (beta - alpha) / beta, scale = 1 / (alpha - beta)
__global__ void __launch_bounds__(256) k_qr_house(double* __restrict__ x, int j, const double* __restrict__ part, int nchunk,
                                                  double* __restrict__ tau, double* __restrict__ hh)
{
    __shared__ double s_red[8];
    double acc = 0.0;
    for (int b = threadIdx.x; b < nchunk; b += 256) acc += part[b];
    const double ss = block_sum(acc, s_red);
    if (threadIdx.x != 0) return;
    const double alpha = x[j];
    if (ss == 0.0) { // zero tail: H = I
        tau[j] = 0.0;
        hh[0] = 0.0;
        hh[1] = 1.0;
        return;
    }
    double beta = sqrt(alpha * alpha + ss);
    if (alpha >= 0.0) beta = -beta;
    tau[j] = (beta - alpha) / beta;
    hh[0] = tau[j];
    hh[1] = 1.0 / (alpha - beta);
    x[j] = beta;
}

__global__ void __launch_bounds__(256) k_qr_scale(double* __restrict__ x, int64_t m, int j, const double* __restrict__ hh)
{
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i > j && i < m) x[i] *= hh[1];
}

// partial dots v . y_c over one row chunk for CT trailing columns (v_j = 1 implied)
__global__ void __launch_bounds__(256) k_qr_dots(const double* __restrict__ QR, int64_t m, int64_t lda, int j, int S,
                                                 double* __restrict__ part, int nchunk)
{
    __shared__ double s_red[8];
    const int64_t r0 = (int64_t)blockIdx.x * CHUNK;
    const double* __restrict__ v = QR + (int64_t)j * lda;
    double vr[CHUNK / 256];
#pragma unroll
    for (int t = 0; t < CHUNK / 256; ++t) {
        const int64_t i = r0 + threadIdx.x + 256 * t;
        vr[t] = (i < j || i >= m) ? 0.0 : (i == j ? 1.0 : v[i]);
    }
    for (int cc = 0; cc < CT; ++cc) {
        const int c = j + 1 + blockIdx.y * CT + cc;
        if (c >= S) break;
        const double* __restrict__ y = QR + (int64_t)c * lda;
        double acc = 0.0;
#pragma unroll
        for (int t = 0; t < CHUNK / 256; ++t) {
            const int64_t i = r0 + threadIdx.x + 256 * t;
            if (i >= j && i < m) acc = fma(vr[t], y[i], acc);
        }
        const double tot = block_sum(acc, s_red);
        if (threadIdx.x == 0) part[(int64_t)c * nchunk + blockIdx.x] = tot;
    }
}

// w_c = tau * sum of the partials (fixed order), kept in part[c * nchunk] for the update pass
__global__ void __launch_bounds__(256) k_qr_dotsum(double* __restrict__ part, int nchunk, int j, const double* __restrict__ hh)
{
    __shared__ double s_red[8];
    const int c = j + 1 + blockIdx.x;
    double acc = 0.0;
    for (int b = threadIdx.x; b < nchunk; b += 256) acc += part[(int64_t)c * nchunk + b];
    const double tot = block_sum(acc, s_red);
    if (threadIdx.x == 0) part[(int64_t)c * nchunk] = tot * hh[0];
}

// y_c -= w_c v on the rows >= j of the trailing columns
__global__ void __launch_bounds__(256) k_qr_apply(double* __restrict__ QR, int64_t m, int64_t lda, int j, int S,
                                                  const double* __restrict__ part, int nchunk)
{
    const int64_t r0 = (int64_t)blockIdx.x * CHUNK;
    const double* __restrict__ v = QR + (int64_t)j * lda;
    double vr[CHUNK / 256];
#pragma unroll
    for (int t = 0; t < CHUNK / 256; ++t) {
        const int64_t i = r0 + threadIdx.x + 256 * t;
        vr[t] = (i < j || i >= m) ? 0.0 : (i == j ? 1.0 : v[i]);
    }
    for (int cc = 0; cc < CT; ++cc) {
        const int c = j + 1 + blockIdx.y * CT + cc;
        if (c >= S) break;
        const double w = part[(int64_t)c * nchunk];
        if (w == 0.0) continue;
        double* __restrict__ y = QR + (int64_t)c * lda;
#pragma unroll
        for (int t = 0; t < CHUNK / 256; ++t) {
            const int64_t i = r0 + threadIdx.x + 256 * t;
            if (i >= j && i < m) y[i] = fma(-w, vr[t], y[i]);
        }
    }
}

// partial sums of (pos - rest)_i * QR(i, s): dbse.cpp:46-48 (FP32 subtract, widened), :53-54
__global__ void __launch_bounds__(256) k_dbse_wpart(const double* __restrict__ QR, int64_t m, int64_t lda,
                                                    const float* __restrict__ pos, const float* __restrict__ rest,
                                                    double* __restrict__ part, int nchunk)
{
    __shared__ double s_red[8];
    const int64_t r0 = (int64_t)blockIdx.x * CHUNK;
    const int s = blockIdx.y;
    const double* __restrict__ q = QR + (int64_t)s * lda;
    double acc = 0.0;
    for (int t = threadIdx.x; t < CHUNK; t += 256) {
        const int64_t i = r0 + t;
        if (i < m) acc = fma((double)__fsub_rn(pos[i], rest[i]), q[i], acc);
    }
    const double tot = block_sum(acc, s_red);
    if (threadIdx.x == 0) part[(int64_t)s * nchunk + blockIdx.x] = tot;
}

// w_s and the clamped FP32 factor of displaceVector: (float)(w * 3), SYSclamp (dbse.cpp:69-71)
__global__ void __launch_bounds__(256) k_dbse_wsum(const double* __restrict__ part, int nchunk, double* __restrict__ w,
                                                   float* __restrict__ cw, int doclamp, float lo, float hi)
{
    __shared__ double s_red[8];
    const int s = blockIdx.x;
    if (part) {
        double acc = 0.0;
        for (int b = threadIdx.x; b < nchunk; b += 256) acc += part[(int64_t)s * nchunk + b];
        const double tot = block_sum(acc, s_red);
        if (threadIdx.x == 0) w[s] = tot;
    }
    if (threadIdx.x == 0) {
        float f = (float)(w[s] * 3);
        if (doclamp) f = f < lo ? lo : (f > hi ? hi : f);
        cw[s] = f;
    }
}

// one thread per coordinate: disp = sum_s M[i][s] * cw_s in column order, FP32, never fused (the reference's
// UT_Vector3 arithmetic, dbse.cpp:65-72); then SOP_FaceDeform.cpp:467-471
__global__ void __launch_bounds__(256) k_dbse_displace(const float* __restrict__ M32, int64_t m, int S,
                                                       const float* __restrict__ cw, const float* __restrict__ pos,
                                                       const float* __restrict__ rest, int dofalloff, float falloffradius,
                                                       float* __restrict__ out)
{
    extern __shared__ float s_cw[];
    for (int s = threadIdx.x; s < S; s += 256) s_cw[s] = cw[s];
    __syncthreads();
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= m) return;
    float disp = 0.f;
    for (int s = 0; s < S; ++s) disp = __fadd_rn(disp, __fmul_rn(M32[(int64_t)s * m + i], s_cw[s]));
    const float r = rest[i];
    if (dofalloff && falloffradius != 0.f) disp = __fadd_rn(disp, __fmul_rn(__fsub_rn(pos[i], r), falloffradius));
    out[i] = __fadd_rn(r, disp);
}

#define DB_ERR(ctx, ...) snprintf((ctx)->err, sizeof((ctx)->err), __VA_ARGS__)

template <typename T> int db_alloc(fd_ctx* ctx, T** p, size_t count)
{
    cudaError_t e = cudaMalloc((void**)p, (count ? count : 1) * sizeof(T));
    if (e != cudaSuccess) {
        DB_ERR(ctx, "cudaMalloc(%zu bytes) failed: %s", count * sizeof(T), cudaGetErrorString(e));
        *p = nullptr;
        return e == cudaErrorMemoryAllocation ? FD_E_NOMEM : FD_E_CUDA;
    }
    return FD_OK;
}

struct DevGuard {
    int prev = -1;
    explicit DevGuard(int dev) { cudaGetDevice(&prev); if (prev != dev) cudaSetDevice(dev); else prev = -1; }
    ~DevGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

} // namespace

extern "C" {

void fd_dbse_destroy(fd_dbse* h)
{
    if (!h) return;
    fd_ctx* owner = h->ctx;
    {
        DevGuard g(owner->device);
        cudaStreamSynchronize(owner->stream);
        void* blocks[] = {h->d_M32, h->d_QR, h->d_tau, h->d_w, h->d_cw, h->d_part, h->d_hh, h->d_a, h->d_b, h->d_o};
        for (void* b : blocks)
            if (b) cudaFree(b);
        delete h;
    }
    fd_ctx_release(owner);
}

int fd_dbse_init(fd_ctx* ctx, const float* rest_P, int64_t n_pts, const float* shapes, int32_t n_shapes, fd_dbse** out)
{
    if (!ctx || !out) return FD_E_INVALID;
    *out = nullptr;
    if (!rest_P || !shapes || n_pts < 1 || n_shapes < 1) { DB_ERR(ctx, "dbse: no points or no blendshapes"); return FD_E_INVALID; }
    DevGuard g(ctx->device);
    fd_dbse* h = new (std::nothrow) fd_dbse();
    if (!h) return FD_E_NOMEM;
    memset(h, 0, sizeof(*h));
    h->ctx = ctx;
    fd_ctx_retain(ctx);
    h->P = n_pts;
    h->m = 3 * n_pts;
    h->lda = (h->m + 3) / 4 * 4;
    h->S = n_shapes;
    h->nchunk = (int)((h->m + CHUNK - 1) / CHUNK);
    const size_t m = (size_t)h->m, S = (size_t)n_shapes;
    int st = db_alloc(ctx, &h->d_M32, m * S);
    if (st == FD_OK) st = db_alloc(ctx, &h->d_QR, (size_t)h->lda * S);
    if (st == FD_OK) st = db_alloc(ctx, &h->d_tau, S);
    if (st == FD_OK) st = db_alloc(ctx, &h->d_w, S);
    if (st == FD_OK) st = db_alloc(ctx, &h->d_cw, S);
    if (st == FD_OK) st = db_alloc(ctx, &h->d_part, (size_t)h->nchunk * S);
    if (st == FD_OK) st = db_alloc(ctx, &h->d_hh, 2);
    if (st == FD_OK) st = db_alloc(ctx, &h->d_a, m);
    if (st == FD_OK) st = db_alloc(ctx, &h->d_b, m);
    if (st == FD_OK) st = db_alloc(ctx, &h->d_o, m);
    // the S shapes are staged through d_QR's tail-free region: copy them to a scratch the build kernel reads
    float* d_shapes = nullptr;
    if (st == FD_OK) st = db_alloc(ctx, &d_shapes, m * S);
    if (st != FD_OK) { if (d_shapes) cudaFree(d_shapes); fd_dbse_destroy(h); return st; }
    cudaStream_t s = ctx->stream;
    cudaError_t e = cudaMemcpyAsync(h->d_b, rest_P, m * sizeof(float), cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_shapes, shapes, m * S * sizeof(float), cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess) e = cudaMemsetAsync(h->d_tau, 0, S * sizeof(double), s);
    if (e == cudaSuccess) {
        dim3 grid((unsigned)((m + 255) / 256), (unsigned)S);
        k_dbse_build<<<grid, 256, 0, s>>>(h->d_b, d_shapes, h->m, h->lda, h->d_M32, h->d_QR);
        ctx->launches += 1;
        const int ncol = (int)(S < m ? S : m);
        cudaEventRecord(ctx->ev_begin[FD_PH_FACTOR], s); // fd_ctx_phase_ms(ctx, 1): the QR alone, without the H2D copies
        for (int j = 0; j < ncol; ++j) {
            double* x = h->d_QR + (size_t)j * h->lda;
            k_qr_colnorm<<<h->nchunk, 256, 0, s>>>(x, h->m, j, h->d_part);
            k_qr_house<<<1, 256, 0, s>>>(x, j, h->d_part, h->nchunk, h->d_tau, h->d_hh);
            k_qr_scale<<<(unsigned)((m + 255) / 256), 256, 0, s>>>(x, h->m, j, h->d_hh);
            ctx->launches += 3;
            const int trailing = (int)S - j - 1;
            if (trailing > 0) {
                dim3 g2((unsigned)h->nchunk, (unsigned)((trailing + CT - 1) / CT));
                k_qr_dots<<<g2, 256, 0, s>>>(h->d_QR, h->m, h->lda, j, (int)S, h->d_part, h->nchunk);
                k_qr_dotsum<<<trailing, 256, 0, s>>>(h->d_part, h->nchunk, j, h->d_hh);
                k_qr_apply<<<g2, 256, 0, s>>>(h->d_QR, h->m, h->lda, j, (int)S, h->d_part, h->nchunk);
                ctx->launches += 3;
            }
        }
        cudaEventRecord(ctx->ev_end[FD_PH_FACTOR], s);
        ctx->phase_valid[FD_PH_FACTOR] = true;
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    cudaFree(d_shapes);
    if (e != cudaSuccess) { DB_ERR(ctx, "dbse init: %s", cudaGetErrorString(e)); fd_dbse_destroy(h); return FD_E_CUDA; }
    *out = h;
    return FD_OK;
}

int fd_dbse_compute_weights(fd_dbse* h, const float* pos, const float* rest, double* weights_out)
{
    if (!h || !pos || !rest) return FD_E_INVALID;
    fd_ctx* ctx = h->ctx;
    DevGuard g(ctx->device);
    cudaStream_t s = ctx->stream;
    const size_t m = (size_t)h->m;
    FD_CUDA_OK(ctx, cudaMemcpyAsync(h->d_a, pos, m * sizeof(float), cudaMemcpyHostToDevice, s));
    FD_CUDA_OK(ctx, cudaMemcpyAsync(h->d_b, rest, m * sizeof(float), cudaMemcpyHostToDevice, s));
    dim3 grid((unsigned)h->nchunk, (unsigned)h->S);
    cudaEventRecord(ctx->ev_begin[FD_PH_SOLVE], s); // fd_ctx_phase_ms(ctx, 2): the two weight kernels
    k_dbse_wpart<<<grid, 256, 0, s>>>(h->d_QR, h->m, h->lda, h->d_a, h->d_b, h->d_part, h->nchunk);
    k_dbse_wsum<<<h->S, 256, 0, s>>>(h->d_part, h->nchunk, h->d_w, h->d_cw, 0, 0.f, 0.f);
    cudaEventRecord(ctx->ev_end[FD_PH_SOLVE], s);
    ctx->phase_valid[FD_PH_SOLVE] = true;
    ctx->launches += 2;
    FD_CUDA_OK(ctx, cudaGetLastError());
    if (weights_out) FD_CUDA_OK(ctx, cudaMemcpyAsync(weights_out, h->d_w, (size_t)h->S * sizeof(double), cudaMemcpyDeviceToHost, s));
    FD_CUDA_OK(ctx, cudaStreamSynchronize(s));
    h->computed = true;
    return FD_OK;
}

int fd_dbse_displace(fd_dbse* h, const float* pos, const float* rest, int32_t doclamp, const float* weightrange,
                     int32_t dofalloff, float falloffradius, float* P_out)
{
    if (!h || !pos || !rest || !P_out || (doclamp && !weightrange)) return FD_E_INVALID;
    fd_ctx* ctx = h->ctx;
    DevGuard g(ctx->device);
    if (!h->computed) { DB_ERR(ctx, "dbse: computeWeights first"); return FD_E_STATE; }
    cudaStream_t s = ctx->stream;
    const size_t m = (size_t)h->m;
    FD_CUDA_OK(ctx, cudaMemcpyAsync(h->d_a, pos, m * sizeof(float), cudaMemcpyHostToDevice, s));
    FD_CUDA_OK(ctx, cudaMemcpyAsync(h->d_b, rest, m * sizeof(float), cudaMemcpyHostToDevice, s));
    cudaEventRecord(ctx->ev_begin[FD_PH_EVAL], s); // fd_ctx_phase_ms(ctx, 3): the displacement kernels
    k_dbse_wsum<<<h->S, 256, 0, s>>>(nullptr, h->nchunk, h->d_w, h->d_cw, doclamp ? 1 : 0, doclamp ? weightrange[0] : 0.f,
                                     doclamp ? weightrange[1] : 0.f);
    k_dbse_displace<<<(unsigned)((m + 255) / 256), 256, (size_t)h->S * sizeof(float), s>>>(
        h->d_M32, h->m, h->S, h->d_cw, h->d_a, h->d_b, dofalloff, falloffradius, h->d_o);
    cudaEventRecord(ctx->ev_end[FD_PH_EVAL], s);
    ctx->phase_valid[FD_PH_EVAL] = true;
    ctx->launches += 2;
    FD_CUDA_OK(ctx, cudaGetLastError());
    FD_CUDA_OK(ctx, cudaMemcpyAsync(P_out, h->d_o, m * sizeof(float), cudaMemcpyDeviceToHost, s));
    FD_CUDA_OK(ctx, cudaStreamSynchronize(s));
    return FD_OK;
}

int fd_dbse_get_weights(fd_dbse* h, double* weights)
{
    if (!h || !weights) return FD_E_INVALID;
    fd_ctx* ctx = h->ctx;
    DevGuard g(ctx->device);
    if (!h->computed) { DB_ERR(ctx, "dbse: weights not computed"); return FD_E_STATE; } // getWeights returns false, :79-81
    FD_CUDA_OK(ctx, cudaMemcpyAsync(weights, h->d_w, (size_t)h->S * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    FD_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
    return FD_OK;
}

int fd_dbse_get_qr(fd_dbse* h, double* qr, double* tau)
{
    if (!h) return FD_E_INVALID;
    fd_ctx* ctx = h->ctx;
    DevGuard g(ctx->device);
    if (qr)
        FD_CUDA_OK(ctx, cudaMemcpy2DAsync(qr, (size_t)h->m * sizeof(double), h->d_QR, (size_t)h->lda * sizeof(double),
                                          (size_t)h->m * sizeof(double), h->S, cudaMemcpyDeviceToHost, ctx->stream));
    if (tau) FD_CUDA_OK(ctx, cudaMemcpyAsync(tau, h->d_tau, (size_t)h->S * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    FD_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
    return FD_OK;
}

int fd_dbse_info(const fd_dbse* h, int64_t* n_pts, int32_t* n_shapes, int32_t* computed)
{
    if (!h) return FD_E_INVALID;
    if (n_pts) *n_pts = h->P;
    if (n_shapes) *n_shapes = h->S;
    if (computed) *computed = h->computed ? 1 : 0;
    return FD_OK;
}

} // extern "C"
