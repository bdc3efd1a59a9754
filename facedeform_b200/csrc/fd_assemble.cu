// fd_assemble.cu -- K1: per-centre radii and assembly of the (N+p)^2 RBF system in FP64.
//
// Replaces the design-matrix build inside alglib::rbfbuildmodel (reference SOP_FaceDeform.cpp:363) and the
// radius rules selected by rbfsetalgoqnn / rbfsetalgomultilayer (:342-349), in the dense saddle-point form
//   [[K + lambda I, P], [P^T, 0]],  K_ij = phi_j(|c_i - c_j|),  P_i = [1, x, y, z] / [1] / nothing.
// Layout: column-major, column stride lda (multiple of 32 doubles), so the LU panel kernels read columns coalesced.
#include "fd_internal.h"

namespace {

constexpr int kTile = 32;

// distance to the nearest other centre (FP64 from FP32 coordinates), R_i = qcoef * d_nn(i)
__global__ void __launch_bounds__(256) k_nn_radius(const float* __restrict__ rest, int N, double qcoef,
                                                   double* __restrict__ radii)
{
    __shared__ float s_c[256 * 3];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    double xi = 0, yi = 0, zi = 0;
    if (i < N) {
        xi = rest[3 * i];
        yi = rest[3 * i + 1];
        zi = rest[3 * i + 2];
    }
    double best = INFINITY;
    for (int j0 = 0; j0 < N; j0 += 256) {
        const int cnt = min(256, N - j0);
        __syncthreads();
        for (int t = threadIdx.x; t < cnt * 3; t += blockDim.x) s_c[t] = rest[3 * j0 + t];
        __syncthreads();
        for (int j = 0; j < cnt; ++j) {
            const double dx = xi - (double)s_c[3 * j], dy = yi - (double)s_c[3 * j + 1], dz = zi - (double)s_c[3 * j + 2];
            const double d2 = dx * dx + dy * dy + dz * dz;
            if (j0 + j != i && d2 < best) best = d2;
        }
    }
    if (i < N) radii[i] = qcoef * sqrt(best);
}

// R_i = min(R_i, zcoef * median(R)), median = the element of rank N/2 (counting rank, ties by index)
__global__ void __launch_bounds__(256) k_radius_median(const double* __restrict__ radii, int N, double* __restrict__ median)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const double ri = radii[i];
    int rank = 0;
    for (int k = 0; k < N; ++k) {
        const double rk = radii[k];
        rank += (rk < ri) || (rk == ri && k < i);
    }
    if (rank == N / 2) *median = ri;
}

__global__ void __launch_bounds__(256) k_radius_cap(double* __restrict__ radii, int N, double zcoef,
                                                    const double* __restrict__ median, int* __restrict__ flags)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    double r = radii[i];
    const double cap = zcoef * (*median);
    if (r > cap) r = cap;
    radii[i] = r;
    if (!(r > 0.0)) atomicExch(&flags[FD_FLAG_ZERO_RADIUS], 1);
}

__global__ void __launch_bounds__(256) k_radius_fill(double* __restrict__ radii, int N, double r)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < N) radii[i] = r;
}

template <int KERNEL>
__device__ __forceinline__ double phi64(double r2, double R)
{
    if (KERNEL == FD_KERNEL_GAUSSIAN) return exp(-r2 / (R * R));
    if (KERNEL == FD_KERNEL_MULTIQUADRIC) return sqrt(r2 + R * R);
    return r2 > 0.0 ? 0.5 * r2 * log(r2) : 0.0;
}

// one 32x32 tile of the system per CTA; threadIdx.x walks rows so column-major stores are coalesced
template <int KERNEL>
__global__ void __launch_bounds__(kTile* 8) k_assemble(const float* __restrict__ rest, const double* __restrict__ radii,
                                                        int N, int np, double lambda, double* __restrict__ A, int lda)
{
    __shared__ float s_ci[kTile][3];
    __shared__ float s_cj[kTile][3];
    __shared__ double s_rj[kTile];
    const int n = N + np;
    const int i0 = blockIdx.x * kTile, j0 = blockIdx.y * kTile;
    const int t = threadIdx.y * kTile + threadIdx.x;
    if (t < kTile * 3) {
        const int r = t / 3, k = t % 3;
        s_ci[r][k] = (i0 + r < N) ? rest[3 * (i0 + r) + k] : 0.f;
        s_cj[r][k] = (j0 + r < N) ? rest[3 * (j0 + r) + k] : 0.f;
    }
    if (t < kTile) s_rj[t] = (j0 + t < N) ? radii[j0 + t] : 1.0;
    __syncthreads();
    const int i = i0 + threadIdx.x;
    if (i >= n) return;
    for (int jj = threadIdx.y; jj < kTile; jj += 8) {
        const int j = j0 + jj;
        if (j >= n) break;
        double v;
        if (i < N && j < N) {
            const double dx = (double)s_ci[threadIdx.x][0] - (double)s_cj[jj][0];
            const double dy = (double)s_ci[threadIdx.x][1] - (double)s_cj[jj][1];
            const double dz = (double)s_ci[threadIdx.x][2] - (double)s_cj[jj][2];
            v = phi64<KERNEL>(dx * dx + dy * dy + dz * dz, s_rj[jj]);
            // smoothing; the multiquadric is conditionally NEGATIVE definite, its shift carries the opposite sign
            if (i == j) v += KERNEL == FD_KERNEL_MULTIQUADRIC ? -lambda : lambda;
        } else if (i < N) { // polynomial column j - N of row i
            const int k = j - N;
            v = (k == 0) ? 1.0 : (double)s_ci[threadIdx.x][k - 1];
        } else if (j < N) { // polynomial row
            const int k = i - N;
            v = (k == 0) ? 1.0 : (double)s_cj[jj][k - 1];
        } else {
            v = 0.0;
        }
        A[(size_t)j * lda + i] = v;
    }
}

} // namespace

cudaError_t fd_launch_radii(fd_ctx* ctx, const fd_params& prm, const float* d_rest, int N, double* d_radii, int* d_flags)
{
    const int blocks = (N + 255) / 256;
    if (prm.model != FD_MODEL_QNN || N == 1) {
        k_radius_fill<<<blocks, 256, 0, ctx->stream>>>(d_radii, N, (double)prm.radius);
        ctx->launches += 1;
        return cudaGetLastError();
    }
    double* d_median = (double*)ctx->stage_dev[FD_STAGE_MISC]; // reserved by fd_rbf_fit_dev
    k_nn_radius<<<blocks, 256, 0, ctx->stream>>>(d_rest, N, (double)prm.qcoef, d_radii);
    k_radius_median<<<blocks, 256, 0, ctx->stream>>>(d_radii, N, d_median);
    k_radius_cap<<<blocks, 256, 0, ctx->stream>>>(d_radii, N, (double)prm.zcoef, d_median, d_flags);
    ctx->launches += 3;
    return cudaGetLastError();
}

cudaError_t fd_launch_assemble(fd_ctx* ctx, const fd_params& prm, const float* d_rest, const double* d_radii, int N,
                               int np, double* d_A, int lda)
{
    const int n = N + np;
    dim3 grid((n + kTile - 1) / kTile, (n + kTile - 1) / kTile), block(kTile, 8);
    switch (prm.kernel) {
    case FD_KERNEL_GAUSSIAN:
        k_assemble<FD_KERNEL_GAUSSIAN><<<grid, block, 0, ctx->stream>>>(d_rest, d_radii, N, np, (double)prm.lambda, d_A, lda);
        break;
    case FD_KERNEL_MULTIQUADRIC:
        k_assemble<FD_KERNEL_MULTIQUADRIC><<<grid, block, 0, ctx->stream>>>(d_rest, d_radii, N, np, (double)prm.lambda, d_A, lda);
        break;
    default:
        k_assemble<FD_KERNEL_THINPLATE><<<grid, block, 0, ctx->stream>>>(d_rest, d_radii, N, np, (double)prm.lambda, d_A, lda);
        break;
    }
    ctx->launches += 1;
    return cudaGetLastError();
}
