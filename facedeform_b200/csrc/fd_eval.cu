// fd_eval.cu -- K3 (FMA/SFU path): fused vertex evaluation
//   P_out[f][v] = P[v] + falloff(v) * tangent_project( sum_j w_j^f phi_j(|P[v] - c_j|) + a0^f + A^f P[v] )
//
// Replaces the serial loop of the reference, SOP_FaceDeform.cpp:404-439 (gate :405-410, rbfcalc :414, tangent
// projection :416-422 + SOP_FaceDeform.hpp:28-41, falloff :423-425, position write :437-438).  Phi is generated
// on the fly and never touches HBM: a CTA stages a tile of centres (float4: c, kernel parameter) and the
// matching weight rows for a chunk of FC frames in shared memory, every thread owns VPT vertices and keeps
// their 3*FC accumulators in registers.  Shared-memory reads are warp-wide broadcasts (LDS.128), so per
// (vertex, centre) pair the issue slots are 6 (distance) + 1-2 (kernel) + 1 MUFU + 3*FC FMA.
// The same template instantiates the FP64 variant used for multiquadric / thin-plate accuracy.
#include "fd_internal.h"

namespace {

constexpr int EVAL_THREADS = 256;
constexpr int TJ = 256; // centres per shared-memory stage

template <typename T> struct Vec4;
template <> struct Vec4<float> { using type = float4; };
template <> struct Vec4<double> { using type = double4; };

__device__ __forceinline__ float ex2_approx(float x)
{
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float lg2_approx(float x)
{
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float sqrt_approx(float x)
{
    float y;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

template <int KERNEL> __device__ __forceinline__ float phi(float r2, float prm)
{
    if (KERNEL == FD_KERNEL_GAUSSIAN) return ex2_approx(r2 * prm);          // prm = -log2(e) / R^2
    if (KERNEL == FD_KERNEL_MULTIQUADRIC) return sqrt_approx(r2 + prm);     // prm = R^2
    return (0.34657359027997264f * r2) * lg2_approx(fmaxf(r2, 1e-37f));     // 0.5 ln2 r^2 log2 r^2
}
template <int KERNEL> __device__ __forceinline__ double phi(double r2, double prm)
{
    if (KERNEL == FD_KERNEL_GAUSSIAN) return exp(r2 * prm);                 // prm = -1 / R^2
    if (KERNEL == FD_KERNEL_MULTIQUADRIC) return sqrt(r2 + prm);
    return r2 > 0.0 ? 0.5 * r2 * log(r2) : 0.0;
}

__device__ __forceinline__ void normalize3(float a[3])
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
__device__ __forceinline__ void project_to_tangents(const float u[3], const float v[3], const float n[3], float d[3])
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
    normalize3(a1);
    normalize3(a2);
    const float da1 = d[0] * a1[0] + d[1] * a1[1] + d[2] * a1[2];
    const float da2 = d[0] * a2[0] + d[1] * a2[1] + d[2] * a2[2];
#pragma unroll
    for (int k = 0; k < 3; ++k) d[k] = a1[k] * da1 + a2[k] * da2;
}

struct EvalArgs {
    const void* ctab;  // float4 / double4 [N]
    const void* W;     // float / double, n x ldw
    int ldw;
    int N, np, F;
    const float* P;
    int64_t V;
    const float* dist2;
    const float* tu;
    const float* tv;
    const float* nrm;
    float* P_out;
    float* falloff_out;
    float radius2;     // radius * radius in FP32 (SOP_FaceDeform.cpp:402)
    float falloffrate;
    int do_tangent;
};

template <typename T, int KERNEL, int FC, int VPT>
__global__ void __launch_bounds__(EVAL_THREADS) k_eval_simt(const EvalArgs a)
{
    using V4 = typename Vec4<T>::type;
    constexpr int WPAD = (3 * FC + 3) / 4 * 4; // weights per centre in shared memory, padded for 128-bit reads
    __shared__ V4 s_c[TJ];
    __shared__ __align__(16) T s_w[TJ * WPAD];

    const int f0 = blockIdx.y * FC;
    const int64_t vbase = (int64_t)blockIdx.x * (EVAL_THREADS * VPT) + threadIdx.x;
    T px[VPT], py[VPT], pz[VPT];
    float pos[VPT][3];
#pragma unroll
    for (int u = 0; u < VPT; ++u) {
        const int64_t v = vbase + (int64_t)u * EVAL_THREADS;
        if (v < a.V) {
            pos[u][0] = a.P[3 * v];
            pos[u][1] = a.P[3 * v + 1];
            pos[u][2] = a.P[3 * v + 2];
        } else {
            pos[u][0] = pos[u][1] = pos[u][2] = 0.f;
        }
        px[u] = (T)pos[u][0];
        py[u] = (T)pos[u][1];
        pz[u] = (T)pos[u][2];
    }
    T acc[VPT][3 * FC];
#pragma unroll
    for (int u = 0; u < VPT; ++u)
#pragma unroll
        for (int c = 0; c < 3 * FC; ++c) acc[u][c] = (T)0;

    const V4* __restrict__ ctab = (const V4*)a.ctab;
    const T* __restrict__ W = (const T*)a.W;
    const int ncol = min(3 * FC, 3 * (a.F - f0)); // valid columns of this frame chunk

    for (int j0 = 0; j0 < a.N; j0 += TJ) {
        const int cnt = min(TJ, a.N - j0);
        __syncthreads();
        for (int t = threadIdx.x; t < TJ; t += EVAL_THREADS) {
            V4 c;
            if (t < cnt) c = ctab[j0 + t];
            else { c.x = c.y = c.z = (T)0; c.w = (T)(KERNEL == FD_KERNEL_MULTIQUADRIC ? 1 : 0); }
            s_c[t] = c;
        }
        for (int t = threadIdx.x; t < TJ * WPAD; t += EVAL_THREADS) {
            const int j = t / WPAD, c = t - j * WPAD;
            s_w[t] = (j < cnt && c < ncol) ? W[(size_t)(j0 + j) * a.ldw + 3 * f0 + c] : (T)0;
        }
        __syncthreads();
        const int jn = (cnt + 3) & ~3; // padded centres carry zero weights
#pragma unroll 4
        for (int j = 0; j < jn; ++j) {
            const V4 c = s_c[j];
            T w[WPAD];
#pragma unroll
            for (int q = 0; q < WPAD; q += 4) {
                const V4 t4 = *reinterpret_cast<const V4*>(&s_w[j * WPAD + q]);
                w[q] = t4.x; w[q + 1] = t4.y; w[q + 2] = t4.z; w[q + 3] = t4.w;
            }
#pragma unroll
            for (int u = 0; u < VPT; ++u) {
                const T dx = px[u] - c.x, dy = py[u] - c.y, dz = pz[u] - c.z;
                const T r2 = dx * dx + dy * dy + dz * dz;
                const T ph = phi<KERNEL>(r2, c.w);
#pragma unroll
                for (int q = 0; q < 3 * FC; ++q) acc[u][q] += w[q] * ph;
            }
        }
    }

    // polynomial block: rows N .. N+np-1 of W hold a0 and the three columns of A (oracle: fdo_calc)
    if (a.np >= 1) {
#pragma unroll
        for (int q = 0; q < 3 * FC; ++q) {
            if (q < ncol) {
                const T a0 = W[(size_t)a.N * a.ldw + 3 * f0 + q];
                T ax = (T)0, ay = (T)0, az = (T)0;
                if (a.np == 4) {
                    ax = W[(size_t)(a.N + 1) * a.ldw + 3 * f0 + q];
                    ay = W[(size_t)(a.N + 2) * a.ldw + 3 * f0 + q];
                    az = W[(size_t)(a.N + 3) * a.ldw + 3 * f0 + q];
                }
#pragma unroll
                for (int u = 0; u < VPT; ++u) acc[u][q] += a0 + ax * px[u] + ay * py[u] + az * pz[u];
            }
        }
    }

    // epilogue: gate, tangent projection, falloff, position write
#pragma unroll
    for (int u = 0; u < VPT; ++u) {
        const int64_t v = vbase + (int64_t)u * EVAL_THREADS;
        if (v >= a.V) continue;
        const float d2 = a.dist2 ? a.dist2[v] : 0.f;
        const bool skip = d2 > a.radius2;                       // SOP_FaceDeform.cpp:408-410
        float fo = fminf(d2 / a.radius2, 1.0f);                 // :423
        fo = powf(1.0f - fo, a.falloffrate);                    // :424
        if (skip) fo = 0.f;
        if (a.falloff_out && blockIdx.y == 0) a.falloff_out[v] = fo;
        float tu[3], tv[3], tn[3];
        if (a.do_tangent) {
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                tu[k] = a.tu[3 * v + k];
                tv[k] = a.tv[3 * v + k];
                tn[k] = a.nrm[3 * v + k];
            }
            normalize3(tu);
            normalize3(tv);
            normalize3(tn);
        }
#pragma unroll
        for (int f = 0; f < FC; ++f) {
            if (f0 + f >= a.F) break;
            float d[3] = {(float)acc[u][3 * f], (float)acc[u][3 * f + 1], (float)acc[u][3 * f + 2]};
            if (a.do_tangent) project_to_tangents(tu, tv, tn, d);
            float* o = a.P_out + ((size_t)(f0 + f) * (size_t)a.V + (size_t)v) * 3;
#pragma unroll
            for (int k = 0; k < 3; ++k) o[k] = skip ? pos[u][k] : pos[u][k] + d[k] * fo;
        }
    }
}

template <typename T, int KERNEL, int FC, int VPT>
cudaError_t launch_one(fd_ctx* ctx, const EvalArgs& a)
{
    dim3 grid((unsigned)((a.V + EVAL_THREADS * VPT - 1) / (EVAL_THREADS * VPT)), (unsigned)((a.F + FC - 1) / FC));
    k_eval_simt<T, KERNEL, FC, VPT><<<grid, EVAL_THREADS, 0, ctx->stream>>>(a);
    ctx->launches += 1;
    return cudaGetLastError();
}

template <typename T, int KERNEL>
cudaError_t launch_fc(fd_ctx* ctx, const EvalArgs& a)
{
    if (sizeof(T) == 8) {
        if (a.F >= 2) return launch_one<T, KERNEL, 2, 2>(ctx, a);
        return launch_one<T, KERNEL, 1, 2>(ctx, a);
    }
    if (a.F >= 4) return launch_one<T, KERNEL, 4, 2>(ctx, a);
    if (a.F >= 2) return launch_one<T, KERNEL, 2, 2>(ctx, a);
    return launch_one<T, KERNEL, 1, 2>(ctx, a);
}

template <typename T>
cudaError_t launch_kernel(fd_ctx* ctx, int kernel, const EvalArgs& a)
{
    switch (kernel) {
    case FD_KERNEL_GAUSSIAN: return launch_fc<T, FD_KERNEL_GAUSSIAN>(ctx, a);
    case FD_KERNEL_MULTIQUADRIC: return launch_fc<T, FD_KERNEL_MULTIQUADRIC>(ctx, a);
    default: return launch_fc<T, FD_KERNEL_THINPLATE>(ctx, a);
    }
}

} // namespace

cudaError_t fd_launch_eval(fd_ctx* ctx, const fd_model* m, const float* P, int64_t V, const float* dist2,
                           const float* tu, const float* tv, const float* nrm, float* P_out, float* falloff_out)
{
    if (V <= 0) return cudaSuccess;
    if (m->use_tc) return fd_launch_eval_tc(ctx, m, P, V, dist2, tu, tv, nrm, P_out, falloff_out);
    EvalArgs a;
    a.N = m->N;
    a.np = m->np;
    a.F = m->F;
    a.P = P;
    a.V = V;
    a.dist2 = dist2;
    a.tu = tu;
    a.tv = tv;
    a.nrm = nrm;
    a.P_out = P_out;
    a.falloff_out = falloff_out;
    a.radius2 = m->prm.radius * m->prm.radius;
    a.falloffrate = m->prm.falloffrate;
    a.do_tangent = (m->prm.tangent && tu && tv && nrm) ? 1 : 0; // SOP_FaceDeform.cpp:293-294
    if (m->eval64) {
        a.ctab = m->d_ctab64;
        a.W = m->d_W;
        a.ldw = m->ldw;
        return launch_kernel<double>(ctx, m->prm.kernel, a);
    }
    a.ctab = m->d_ctab32;
    a.W = m->d_W32;
    a.ldw = m->ldw32;
    return launch_kernel<float>(ctx, m->prm.kernel, a);
}
