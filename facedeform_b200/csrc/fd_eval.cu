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
#include <stdlib.h>

#include "fd_eval_common.cuh"

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

// packed FP32 pairs (FADD2 / FMUL2 / FFMA2 on sm_100a): the same arithmetic for two vertices in one issue slot
__device__ __forceinline__ uint64_t pack2(float lo, float hi)
{
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(uint64_t v, float& lo, float& hi)
{
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t sub2(uint64_t a, uint64_t b)
{
    uint64_t r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b)
{
    uint64_t r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b)
{
    uint64_t r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c)
{
    uint64_t r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
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

// (the FP64 kernel functions fd_fast_sqrt64 / fd_half_log64 and the tangent projection live in fd_eval_common.cuh)

struct EvalArgs {
    const void* ctab;  // float4 / double4 [N]
    const float* origin; // FP64 multiquadric / thin plate: coordinates are taken relative to this point (centre 0)
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
    const int* sel; // device word holding the evaluation kernel FD_EVAL_AUTO chose, or NULL (no choice to make);
    int sel_id;     // this launch's id: the kernel returns at once when *sel differs
};

// polynomial block + SOP epilogue (gate, tangent projection, falloff, position write), shared by the evaluation kernels
template <typename T, int FC, int VPT, int NT = EVAL_THREADS>
__device__ __forceinline__ void finish_vertices(const EvalArgs& a, const T* __restrict__ W, int f0, int ncol, int64_t vbase,
                                                const T (&px)[VPT], const T (&py)[VPT], const T (&pz)[VPT],
                                                const float (&pos)[VPT][3], T (&acc)[VPT][3 * FC])
{
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
        const int64_t v = vbase + (int64_t)u * NT;
        if (v >= a.V) continue;
        const float d2 = a.dist2 ? a.dist2[v] : 0.f;
        const bool skip = d2 > a.radius2;                       // SOP_FaceDeform.cpp:408-410
        float fo = fminf(d2 / a.radius2, 1.0f);                 // :423
        fo = powf(1.0f - fo, a.falloffrate);                    // :424
        if (skip) fo = 0.f;
        if (a.falloff_out && f0 == 0) a.falloff_out[v] = fo;
        float tu[3], tv[3], tn[3];
        if (a.do_tangent) {
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                tu[k] = a.tu[3 * v + k];
                tv[k] = a.tv[3 * v + k];
                tn[k] = a.nrm[3 * v + k];
            }
            fd_normalize3(tu);
            fd_normalize3(tv);
            fd_normalize3(tn);
        }
#pragma unroll
        for (int f = 0; f < FC; ++f) {
            if (f0 + f >= a.F) break;
            float d[3] = {(float)acc[u][3 * f], (float)acc[u][3 * f + 1], (float)acc[u][3 * f + 2]};
            if (a.do_tangent) fd_project_to_tangents(tu, tv, tn, d);
            float* o = a.P_out + ((size_t)(f0 + f) * (size_t)a.V + (size_t)v) * 3;
#pragma unroll
            for (int k = 0; k < 3; ++k) o[k] = skip ? pos[u][k] : pos[u][k] + d[k] * fo;
        }
    }
}

template <typename T, int KERNEL, int FC, int VPT>
__global__ void __launch_bounds__(EVAL_THREADS) k_eval_simt(const EvalArgs a)
{
    using V4 = typename Vec4<T>::type;
    constexpr int WPAD = (3 * FC + 3) / 4 * 4; // weights per centre in shared memory, padded for 128-bit reads
    __shared__ V4 s_c[TJ];
    __shared__ __align__(16) T s_w[TJ * WPAD];

    if (a.sel && *a.sel != a.sel_id) return;
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

    finish_vertices<T, FC, VPT>(a, W, f0, ncol, vbase, px, py, pz, pos, acc);
}

// FP32 FMA/SFU path with packed pairs: a thread owns two vertices and carries them as the two halves of FP32x2
// registers, so the distance, the kernel argument and the 3 FC accumulations of both take one issue slot each
// (per vertex-centre pair 3.5 + 1.5 FC packed instructions + 1 MUFU instead of 8 + 3 FC + 1: the scalar kernel was
// issue-bound at 64 % of the FMA pipe).  Lane-wise identical arithmetic to k_eval_simt<float> (same rounding per op).
template <int KERNEL, int FC, int VP>
__global__ void __launch_bounds__(EVAL_THREADS) k_eval_f32x2(const EvalArgs a)
{
    constexpr int VPT = 2 * VP; // VP packed pairs of vertices per thread
    constexpr int WPAD = (3 * FC + 3) / 4 * 4;
    __shared__ float4 s_c[TJ];
    __shared__ __align__(16) float s_w[TJ * WPAD];

    if (a.sel && *a.sel != a.sel_id) return;
    // persistent walk over (vertex tile, frame chunk) pairs, vertex tiles fastest: a grid of a few CTAs per SM costs
    // nothing when FD_EVAL_AUTO settled on another kernel, and consecutive tiles reuse the chunk's weight rows in L2
    const int64_t nbx = (a.V + EVAL_THREADS * VPT - 1) / (EVAL_THREADS * VPT);
    const int64_t ntiles = nbx * ((a.F + FC - 1) / FC);
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int f0 = (int)(tile / nbx) * FC;
    const int64_t vbase = (tile % nbx) * (EVAL_THREADS * VPT) + threadIdx.x;
    float px[VPT], py[VPT], pz[VPT];
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
        px[u] = pos[u][0];
        py[u] = pos[u][1];
        pz[u] = pos[u][2];
    }
    uint64_t px2[VP], py2[VP], pz2[VP];
#pragma unroll
    for (int g = 0; g < VP; ++g) {
        px2[g] = pack2(px[2 * g], px[2 * g + 1]);
        py2[g] = pack2(py[2 * g], py[2 * g + 1]);
        pz2[g] = pack2(pz[2 * g], pz[2 * g + 1]);
    }
    uint64_t acc2[VP][3 * FC];
#pragma unroll
    for (int g = 0; g < VP; ++g)
#pragma unroll
        for (int c = 0; c < 3 * FC; ++c) acc2[g][c] = pack2(0.f, 0.f);

    const float4* __restrict__ ctab = (const float4*)a.ctab;
    const float* __restrict__ W = (const float*)a.W;
    const int ncol = min(3 * FC, 3 * (a.F - f0));

    for (int j0 = 0; j0 < a.N; j0 += TJ) {
        const int cnt = min(TJ, a.N - j0);
        __syncthreads();
        for (int t = threadIdx.x; t < TJ; t += EVAL_THREADS)
            s_c[t] = t < cnt ? ctab[j0 + t] : make_float4(0.f, 0.f, 0.f, KERNEL == FD_KERNEL_MULTIQUADRIC ? 1.f : 0.f);
        for (int t = threadIdx.x; t < TJ * WPAD; t += EVAL_THREADS) {
            const int j = t / WPAD, c = t - j * WPAD;
            s_w[t] = (j < cnt && c < ncol) ? W[(size_t)(j0 + j) * a.ldw + 3 * f0 + c] : 0.f;
        }
        __syncthreads();
        const int jn = (cnt + 3) & ~3; // padded centres carry zero weights
#pragma unroll 4
        for (int j = 0; j < jn; ++j) {
            const float4 c = s_c[j];
            float w[WPAD];
#pragma unroll
            for (int q = 0; q < WPAD; q += 4) {
                const float4 t4 = *reinterpret_cast<const float4*>(&s_w[j * WPAD + q]);
                w[q] = t4.x; w[q + 1] = t4.y; w[q + 2] = t4.z; w[q + 3] = t4.w;
            }
            const uint64_t cx = pack2(c.x, c.x), cy = pack2(c.y, c.y), cz = pack2(c.z, c.z), cw = pack2(c.w, c.w);
#pragma unroll
            for (int g = 0; g < VP; ++g) {
                const uint64_t dx = sub2(px2[g], cx), dy = sub2(py2[g], cy), dz = sub2(pz2[g], cz);
                const uint64_t r2 = fma2(dz, dz, fma2(dy, dy, mul2(dx, dx)));
                float t0, t1;
                uint64_t ph2;
                if (KERNEL == FD_KERNEL_GAUSSIAN) {
                    unpack2(mul2(r2, cw), t0, t1);
                    ph2 = pack2(ex2_approx(t0), ex2_approx(t1));
                } else if (KERNEL == FD_KERNEL_MULTIQUADRIC) {
                    unpack2(add2(r2, cw), t0, t1);
                    ph2 = pack2(sqrt_approx(t0), sqrt_approx(t1));
                } else {
                    unpack2(r2, t0, t1);
                    ph2 = pack2(phi<KERNEL>(t0, 0.f), phi<KERNEL>(t1, 0.f));
                }
#pragma unroll
                for (int q = 0; q < 3 * FC; ++q) acc2[g][q] = fma2(pack2(w[q], w[q]), ph2, acc2[g][q]);
            }
        }
    }
    float acc[VPT][3 * FC];
#pragma unroll
    for (int g = 0; g < VP; ++g)
#pragma unroll
        for (int q = 0; q < 3 * FC; ++q) unpack2(acc2[g][q], acc[2 * g][q], acc[2 * g + 1][q]);
    finish_vertices<float, FC, VPT>(a, W, f0, ncol, vbase, px, py, pz, pos, acc);
    } // tile
}

// FP32 FMA/SFU path for frame chunks of FC >= 4 (an even number of columns): the packed FMAs run along the COLUMNS.
// A packed register holds the weights of two adjacent columns exactly as they lie in shared memory, so an LDS.128
// delivers two ready operands; the other operand is (phi, phi), one MOV per basis value.  Per (vertex pair, centre):
// 7 packed instructions for the two distances, 2 MUFU, 2 MOV, 3 FC FFMA2 and 2 + 3 FC / 4 shared-memory reads -- the
// kernel sits on the FP32 pipe (an FFMA2 occupies it for two cycles), not on issue slots or shared-memory bandwidth.
// (Packing along the vertices needs every weight duplicated: one MOV per FFMA2 when done in registers -- issue bound --
// or twice the LDS.128 broadcasts when done in shared memory -- LSU bound: 0.98 ms and 1.40 ms at BASELINE configs[1].)
// Lane-wise the same arithmetic per output as k_eval_simt<float> (same rounding per operation, same order over centres).
template <int KERNEL, int FC, int VP, int UNR = 2, int NT = EVAL_THREADS>
__global__ void __launch_bounds__(NT) k_eval_f32c(const EvalArgs a)
{
    constexpr int VPT = 2 * VP;      // vertices per thread; the distances of a pair share packed instructions
    constexpr int NC = 3 * FC;       // columns of the chunk (even)
    constexpr int TJ = FC <= 8 ? 256 : 128; // centres per shared-memory stage (static shared memory: 48 KB)
    static_assert(NC % 4 == 0, "column count must be a multiple of 4 (128-bit weight reads)");
    __shared__ __align__(16) uint64_t s_c2[TJ * 4];  // (x, x), (y, y), (z, z), (parameter, parameter)
    __shared__ __align__(16) float s_w[TJ * NC];

    if (a.sel && *a.sel != a.sel_id) return;
    const int64_t nbx = (a.V + NT * VPT - 1) / (NT * VPT);
    const int64_t ntiles = nbx * ((a.F + FC - 1) / FC);
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int f0 = (int)(tile / nbx) * FC;
        const int64_t vbase = (tile % nbx) * (NT * VPT) + threadIdx.x;
        float px[VPT], py[VPT], pz[VPT];
        float pos[VPT][3];
#pragma unroll
        for (int u = 0; u < VPT; ++u) {
            const int64_t v = vbase + (int64_t)u * NT;
            if (v < a.V) {
                pos[u][0] = a.P[3 * v];
                pos[u][1] = a.P[3 * v + 1];
                pos[u][2] = a.P[3 * v + 2];
            } else {
                pos[u][0] = pos[u][1] = pos[u][2] = 0.f;
            }
            px[u] = pos[u][0];
            py[u] = pos[u][1];
            pz[u] = pos[u][2];
        }
        uint64_t px2[VP], py2[VP], pz2[VP];
#pragma unroll
        for (int g = 0; g < VP; ++g) {
            px2[g] = pack2(px[2 * g], px[2 * g + 1]);
            py2[g] = pack2(py[2 * g], py[2 * g + 1]);
            pz2[g] = pack2(pz[2 * g], pz[2 * g + 1]);
        }
        uint64_t acc2[VPT][NC / 2]; // [vertex][column pair]
#pragma unroll
        for (int u = 0; u < VPT; ++u)
#pragma unroll
            for (int c = 0; c < NC / 2; ++c) acc2[u][c] = pack2(0.f, 0.f);

        const float4* __restrict__ ctab = (const float4*)a.ctab;
        const float* __restrict__ W = (const float*)a.W;
        const int ncol = min(NC, 3 * (a.F - f0));

        for (int j0 = 0; j0 < a.N; j0 += TJ) {
            const int cnt = min(TJ, a.N - j0);
            __syncthreads();
            for (int t = threadIdx.x; t < TJ; t += NT) {
                const float4 c = t < cnt ? ctab[j0 + t] : make_float4(0.f, 0.f, 0.f, KERNEL == FD_KERNEL_MULTIQUADRIC ? 1.f : 0.f);
                s_c2[4 * t] = pack2(c.x, c.x);
                s_c2[4 * t + 1] = pack2(c.y, c.y);
                s_c2[4 * t + 2] = pack2(c.z, c.z);
                s_c2[4 * t + 3] = pack2(c.w, c.w);
            }
            for (int t = threadIdx.x; t < TJ * NC; t += NT) {
                const int j = t / NC, c = t - j * NC;
                s_w[t] = (j < cnt && c < ncol) ? W[(size_t)(j0 + j) * a.ldw + 3 * f0 + c] : 0.f;
            }
            __syncthreads();
            const int jn = (cnt + 3) & ~3; // padded centres carry zero weights
#pragma unroll UNR
            for (int j = 0; j < jn; ++j) {
                const ulonglong2 cxy = *reinterpret_cast<const ulonglong2*>(&s_c2[4 * j]);
                const ulonglong2 czw = *reinterpret_cast<const ulonglong2*>(&s_c2[4 * j + 2]);
                uint64_t w2[NC / 2];
#pragma unroll
                for (int q = 0; q < NC / 2; q += 2) {
                    const ulonglong2 t2 = *reinterpret_cast<const ulonglong2*>(&s_w[j * NC + 2 * q]);
                    w2[q] = t2.x;     // (w[2q], w[2q+1])
                    w2[q + 1] = t2.y; // (w[2q+2], w[2q+3])
                }
#pragma unroll
                for (int g = 0; g < VP; ++g) {
                    const uint64_t dx = sub2(px2[g], cxy.x), dy = sub2(py2[g], cxy.y), dz = sub2(pz2[g], czw.x);
                    const uint64_t r2 = fma2(dz, dz, fma2(dy, dy, mul2(dx, dx)));
                    float t0, t1, e0, e1;
                    if (KERNEL == FD_KERNEL_GAUSSIAN) {
                        unpack2(mul2(r2, czw.y), t0, t1);
                        e0 = ex2_approx(t0), e1 = ex2_approx(t1);
                    } else if (KERNEL == FD_KERNEL_MULTIQUADRIC) {
                        unpack2(add2(r2, czw.y), t0, t1);
                        e0 = sqrt_approx(t0), e1 = sqrt_approx(t1);
                    } else {
                        unpack2(r2, t0, t1);
                        e0 = phi<KERNEL>(t0, 0.f), e1 = phi<KERNEL>(t1, 0.f);
                    }
                    const uint64_t pa = pack2(e0, e0), pb = pack2(e1, e1);
#pragma unroll
                    for (int q = 0; q < NC / 2; ++q) {
                        acc2[2 * g][q] = fma2(w2[q], pa, acc2[2 * g][q]);
                        acc2[2 * g + 1][q] = fma2(w2[q], pb, acc2[2 * g + 1][q]);
                    }
                }
            }
        }
        float acc[VPT][NC];
#pragma unroll
        for (int u = 0; u < VPT; ++u)
#pragma unroll
            for (int q = 0; q < NC / 2; ++q) unpack2(acc2[u][q], acc[u][2 * q], acc[u][2 * q + 1]);
        finish_vertices<float, FC, VPT, NT>(a, W, f0, ncol, vbase, px, py, pz, pos, acc);
    } // tile
}

template <int KERNEL, int FC, int VP, int UNR = 2, int NT = EVAL_THREADS>
cudaError_t launch_f32c(fd_ctx* ctx, const EvalArgs& a)
{
    const int64_t ntiles = ((a.V + NT * 2 * VP - 1) / (NT * 2 * VP)) * ((a.F + FC - 1) / FC);
    const int64_t cap = (int64_t)ctx->sm_count * 8;
    k_eval_f32c<KERNEL, FC, VP, UNR, NT><<<(unsigned)(ntiles < cap ? ntiles : cap), NT, 0, ctx->stream>>>(a);
    ctx->launches += 1;
    return cudaGetLastError();
}

template <int KERNEL, int FC, int VP>
cudaError_t launch_f32x2(fd_ctx* ctx, const EvalArgs& a)
{
    const int64_t ntiles = ((a.V + EVAL_THREADS * 2 * VP - 1) / (EVAL_THREADS * 2 * VP)) * ((a.F + FC - 1) / FC);
    const int64_t cap = (int64_t)ctx->sm_count * 8; // up to 8 resident CTAs of 256 threads per SM
    k_eval_f32x2<KERNEL, FC, VP><<<(unsigned)(ntiles < cap ? ntiles : cap), EVAL_THREADS, 0, ctx->stream>>>(a);
    ctx->launches += 1;
    return cudaGetLastError();
}

template <int KERNEL>
cudaError_t launch_f32x2_fc(fd_ctx* ctx, const EvalArgs& a)
{
    const int vp_env = ctx->dbg.eval_vp;
    // frame chunks of 8 / 4: packed along the columns (k_eval_f32c); 4 vertices per thread when the mesh fills the GPU that way
    const int fc = a.F >= 8 ? 8 : 4;
    // 12 frames per chunk when the mesh still fills the GPU: 36 accumulations per basis value instead of 24 (the distance and
    // the kernel function are 7 of 31 FP32-pipe cycles per pair at 8 frames, 7 of 43 at 12; BASELINE configs[1]: 1.03 -> 0.93 ms;
    // 212 registers, one CTA per SM like the 8-frame kernel's 172 -- 16 frames with two vertices per thread lost to the
    // shared-memory reads, 384-thread CTAs and deeper unrolling moved nothing: profiles/r2_eval_f32c_variants.txt)
    if (a.F >= 12 && ((a.V + EVAL_THREADS * 4 - 1) / (EVAL_THREADS * 4)) * ((a.F + 11) / 12) >= 2 * (int64_t)ctx->sm_count)
        return launch_f32c<KERNEL, 12, 2>(ctx, a);
    const bool many = ((a.V + EVAL_THREADS * 4 - 1) / (EVAL_THREADS * 4)) * ((a.F + fc - 1) / fc) >= 2 * (int64_t)ctx->sm_count;
    if (a.F >= 8) return many ? launch_f32c<KERNEL, 8, 2>(ctx, a) : launch_f32c<KERNEL, 8, 1>(ctx, a);
    if (a.F >= 4) return many ? launch_f32c<KERNEL, 4, 2>(ctx, a) : launch_f32c<KERNEL, 4, 1>(ctx, a);
    if (a.F >= 2) return launch_f32x2<KERNEL, 2, 1>(ctx, a); // one or two frames: packed along the vertices (k_eval_f32x2)
    // one frame: two packed pairs per thread when there are enough vertices to fill the GPU that way
    const bool wide = vp_env ? vp_env == 2 : a.V >= (int64_t)ctx->sm_count * EVAL_THREADS * 4 * 4;
    return wide ? launch_f32x2<KERNEL, 1, 2>(ctx, a) : launch_f32x2<KERNEL, 1, 1>(ctx, a);
}

// FP64 multiquadric / thin plate: distance in the expanded form (the centre table holds -2 (c - o) and |c - o|^2 +
// kernel parameter, o = centre 0; cancellation is harmless at 53 bits), kernel functions from fast_sqrt64 / half_log64:
// per (vertex, centre) pair 4 + 4 (+ 8 thin plate) + 3 FC FP64 instructions instead of ~26 / ~60 with libdevice.
template <int KERNEL, int FC, int VPT>
__global__ void __launch_bounds__(EVAL_THREADS) k_eval_f64(const EvalArgs a)
{
    constexpr int WPAD = (3 * FC + 1) / 2 * 2; // weights per centre in shared memory, padded for 128-bit reads
    __shared__ double4 s_c[TJ];
    __shared__ __align__(16) double s_w[TJ * WPAD];
    __shared__ double2 s_tab[128];

    if (a.sel && *a.sel != a.sel_id) return; // FD_EVAL_AUTO settled on another kernel (fd_eval64.cu: k_cancel_select)
    if (KERNEL == FD_KERNEL_THINPLATE && threadIdx.x < 128) fd_half_log64_table(s_tab, threadIdx.x);
    const int f0 = blockIdx.y * FC;
    const int64_t vbase = (int64_t)blockIdx.x * (EVAL_THREADS * VPT) + threadIdx.x;
    const double ox = (double)a.origin[0], oy = (double)a.origin[1], oz = (double)a.origin[2];
    double px[VPT], py[VPT], pz[VPT], pp[VPT], qx[VPT], qy[VPT], qz[VPT];
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
        px[u] = (double)pos[u][0];
        py[u] = (double)pos[u][1];
        pz[u] = (double)pos[u][2];
        qx[u] = px[u] - ox;
        qy[u] = py[u] - oy;
        qz[u] = pz[u] - oz;
        pp[u] = qx[u] * qx[u] + qy[u] * qy[u] + qz[u] * qz[u];
    }
    double acc[VPT][3 * FC];
#pragma unroll
    for (int u = 0; u < VPT; ++u)
#pragma unroll
        for (int c = 0; c < 3 * FC; ++c) acc[u][c] = 0.0;

    const double4* __restrict__ ctab = (const double4*)a.ctab;
    const double* __restrict__ W = (const double*)a.W;
    const int ncol = min(3 * FC, 3 * (a.F - f0));

    for (int j0 = 0; j0 < a.N; j0 += TJ) {
        const int cnt = min(TJ, a.N - j0);
        __syncthreads();
        for (int t = threadIdx.x; t < TJ; t += EVAL_THREADS)
            s_c[t] = t < cnt ? ctab[j0 + t] : make_double4(0.0, 0.0, 0.0, 1.0); // padded centres carry zero weights
        for (int t = threadIdx.x; t < TJ * WPAD; t += EVAL_THREADS) {
            const int j = t / WPAD, c = t - j * WPAD;
            s_w[t] = (j < cnt && c < ncol) ? W[(size_t)(j0 + j) * a.ldw + 3 * f0 + c] : 0.0;
        }
        __syncthreads();
        const int jn = (cnt + 3) & ~3;
#pragma unroll 4
        for (int j = 0; j < jn; ++j) {
            const double4 c = s_c[j];
            double w[WPAD];
#pragma unroll
            for (int q = 0; q < WPAD; q += 2) {
                const double2 t2 = *reinterpret_cast<const double2*>(&s_w[j * WPAD + q]);
                w[q] = t2.x;
                w[q + 1] = t2.y;
            }
#pragma unroll
            for (int u = 0; u < VPT; ++u) {
                const double x = fma(qx[u], c.x, fma(qy[u], c.y, fma(qz[u], c.z, c.w))) + pp[u]; // r^2 (+ R^2)
                const double ph = KERNEL == FD_KERNEL_MULTIQUADRIC ? fd_fast_sqrt64(x) : x * fd_half_log64(x, s_tab);
#pragma unroll
                for (int q = 0; q < 3 * FC; ++q) acc[u][q] = fma(w[q], ph, acc[u][q]);
            }
        }
    }
    finish_vertices<double, FC, VPT>(a, W, f0, ncol, vbase, px, py, pz, pos, acc);
}

template <int KERNEL, int FC, int VPT>
cudaError_t launch_f64(fd_ctx* ctx, const EvalArgs& a)
{
    dim3 grid((unsigned)((a.V + EVAL_THREADS * VPT - 1) / (EVAL_THREADS * VPT)), (unsigned)((a.F + FC - 1) / FC));
    k_eval_f64<KERNEL, FC, VPT><<<grid, EVAL_THREADS, 0, ctx->stream>>>(a);
    ctx->launches += 1;
    return cudaGetLastError();
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
    // FD_EVAL_AUTO with the Gaussian: the FP32 candidate and the FP64 kernel are both launched, the device word d_sel
    // (written after the solve, fd_eval64.cu) lets exactly one of them run -- no host synchronisation in between
    const int* sel = (!m->eval64 && m->auto_sel) ? m->d_sel : nullptr;
    cudaError_t e = cudaSuccess;
    if (m->use_tcx) { // wide Gaussian batches under FD_EVAL_AUTO: the exact-digit tensor-core kernel, FP64 if its bound fails
        e = fd_launch_eval_tcx(ctx, m, P, V, dist2, tu, tv, nrm, P_out, falloff_out, sel, FD_SEL_TCX);
        if (!sel || e != cudaSuccess) return e;
    }
    if (!m->eval64 && m->use_tc) {
        e = fd_launch_eval_tc(ctx, m, P, V, dist2, tu, tv, nrm, P_out, falloff_out, sel, FD_SEL_TENSOR);
        if (!sel || e != cudaSuccess) return e;
    }
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
    a.origin = m->d_rest;
    a.sel = sel;
    if (!m->eval64 && !m->use_tcx && (!m->use_tc || (sel && m->prm.eval_path != FD_PATH_TENSOR))) { // FP32 FMA/SFU
        a.sel_id = FD_SEL_SIMT;
        a.ctab = m->d_ctab32;
        a.W = m->d_W32;
        a.ldw = m->ldw32;
        if (ctx->dbg.eval_scalar_f32) { // the un-packed kernel, kept for comparison
            e = launch_kernel<float>(ctx, m->prm.kernel, a);
        } else {
            switch (m->prm.kernel) {
            case FD_KERNEL_GAUSSIAN: e = launch_f32x2_fc<FD_KERNEL_GAUSSIAN>(ctx, a); break;
            case FD_KERNEL_MULTIQUADRIC: e = launch_f32x2_fc<FD_KERNEL_MULTIQUADRIC>(ctx, a); break;
            default: e = launch_f32x2_fc<FD_KERNEL_THINPLATE>(ctx, a); break;
            }
        }
        if (!sel || e != cudaSuccess) return e;
    }
    // FP64: the DMMA kernel for wide 3F, else one thread per vertex
    a.sel_id = FD_SEL_FP64;
    if (3 * m->F >= FD_MMA64_MIN_COLUMNS)
        return fd_launch_eval64_mma(ctx, m, P, V, dist2, tu, tv, nrm, P_out, falloff_out, sel, FD_SEL_FP64);
    a.ctab = m->d_ctab64;
    a.W = m->d_W;
    a.ldw = m->ldw;
    if (m->prm.kernel == FD_KERNEL_MULTIQUADRIC)
        return a.F >= 2 ? launch_f64<FD_KERNEL_MULTIQUADRIC, 2, 2>(ctx, a) : launch_f64<FD_KERNEL_MULTIQUADRIC, 1, 4>(ctx, a);
    if (m->prm.kernel == FD_KERNEL_THINPLATE)
        return a.F >= 2 ? launch_f64<FD_KERNEL_THINPLATE, 2, 2>(ctx, a) : launch_f64<FD_KERNEL_THINPLATE, 1, 4>(ctx, a);
    return launch_kernel<double>(ctx, m->prm.kernel, a); // Gaussian in FP64
}

// the frames [f_begin, f_begin + f_count) only: a view of the model whose weight pointers start at the block's first
// column (no kernel knows about it)
cudaError_t fd_launch_eval_frames(fd_ctx* ctx, const fd_model* m, const float* P, int64_t V, const float* dist2, const float* tu,
                                  const float* tv, const float* nrm, float* P_out, float* falloff_out, int f_begin, int f_count)
{
    if (f_begin == 0 && f_count == m->F) return fd_launch_eval(ctx, m, P, V, dist2, tu, tv, nrm, P_out, falloff_out);
    fd_model view = *m;
    view.F = f_count;
    view.w_col0 = 3 * f_begin;
    if (view.d_W32) view.d_W32 = m->d_W32 + 3 * f_begin;
    view.d_W = m->d_W + 3 * f_begin; // f_begin is a multiple of 80: the 16-byte alignment of the FP64 rows is kept
    if (m->use_tcx && !fd_tcx_view_frames(m, &view, f_begin)) return cudaErrorInvalidValue;
    if (m->use_tc && !fd_tc_view_frames(m, &view, f_begin)) return cudaErrorInvalidValue;
    return fd_launch_eval(ctx, &view, P, V, dist2, tu, tv, nrm, P_out + (size_t)f_begin * (size_t)V * 3,
                          f_begin == 0 ? falloff_out : nullptr);
}
