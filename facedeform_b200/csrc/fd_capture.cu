// fd_capture.cu -- float part of ProximityCapture on the GPU (reference capture.cpp:68-99, :122).
//
//   k_nearest       GEO_PointTree::findNearestIdx per rig point (capture.cpp:122): exact brute-force nearest
//                   mesh point, ties -> lowest index, via a packed (distance bits, index) 64-bit atomicMin.
//   k_capture_dist  GU_RayIntersect::minimumPoint with GU_MinInfo(R^2) per grouped vertex (capture.cpp:76-88):
//                   closest squared distance to the rig primitives when it is < R^2, else -1; 0 for vertices
//                   outside every handle group or when dofalloff is off (capture.cpp:71-75).
//
// This file is compiled with -fmad=false: every FP32 expression below is evaluated un-fused in the written
// order, with IEEE division, so indices and distances are bit-exact against the CPU oracle, which states the
// same operation order (tests/test_gpu_capture.py).
#include "fd_internal.h"

namespace {

__device__ __forceinline__ float dot3(const float a[3], const float b[3])
{
    return (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2];
}
__device__ __forceinline__ float dist2_3(const float a[3], const float b[3])
{
    const float d[3] = {a[0] - b[0], a[1] - b[1], a[2] - b[2]};
    return dot3(d, d);
}

__device__ float point_seg_dist2(const float p[3], const float a[3], const float b[3])
{
    const float ab[3] = {b[0] - a[0], b[1] - a[1], b[2] - a[2]};
    const float ap[3] = {p[0] - a[0], p[1] - a[1], p[2] - a[2]};
    const float e = dot3(ap, ab);
    if (e <= 0.0f) return dot3(ap, ap);
    const float f = dot3(ab, ab);
    if (e >= f) return dist2_3(p, b);
    const float t = e / f;
    const float q[3] = {a[0] + t * ab[0], a[1] + t * ab[1], a[2] + t * ab[2]};
    return dist2_3(p, q);
}

// Voronoi-region closest point on a triangle (Ericson, Real-Time Collision Detection, 5.1.5)
__device__ float point_tri_dist2(const float p[3], const float a[3], const float b[3], const float c[3])
{
    const float ab[3] = {b[0] - a[0], b[1] - a[1], b[2] - a[2]};
    const float ac[3] = {c[0] - a[0], c[1] - a[1], c[2] - a[2]};
    const float ap[3] = {p[0] - a[0], p[1] - a[1], p[2] - a[2]};
    const float d1 = dot3(ab, ap), d2 = dot3(ac, ap);
    if (d1 <= 0.0f && d2 <= 0.0f) return dot3(ap, ap);
    const float bp[3] = {p[0] - b[0], p[1] - b[1], p[2] - b[2]};
    const float d3 = dot3(ab, bp), d4 = dot3(ac, bp);
    if (d3 >= 0.0f && d4 <= d3) return dot3(bp, bp);
    const float vc = d1 * d4 - d3 * d2;
    if (vc <= 0.0f && d1 >= 0.0f && d3 <= 0.0f) {
        const float t = d1 / (d1 - d3);
        const float q[3] = {a[0] + t * ab[0], a[1] + t * ab[1], a[2] + t * ab[2]};
        return dist2_3(p, q);
    }
    const float cp[3] = {p[0] - c[0], p[1] - c[1], p[2] - c[2]};
    const float d5 = dot3(ab, cp), d6 = dot3(ac, cp);
    if (d6 >= 0.0f && d5 <= d6) return dot3(cp, cp);
    const float vb = d5 * d2 - d1 * d6;
    if (vb <= 0.0f && d2 >= 0.0f && d6 <= 0.0f) {
        const float t = d2 / (d2 - d6);
        const float q[3] = {a[0] + t * ac[0], a[1] + t * ac[1], a[2] + t * ac[2]};
        return dist2_3(p, q);
    }
    const float va = d3 * d6 - d5 * d4;
    if (va <= 0.0f && (d4 - d3) >= 0.0f && (d5 - d6) >= 0.0f) {
        const float t = (d4 - d3) / ((d4 - d3) + (d5 - d6));
        const float q[3] = {b[0] + t * (c[0] - b[0]), b[1] + t * (c[1] - b[1]), b[2] + t * (c[2] - b[2])};
        return dist2_3(p, q);
    }
    const float denom = 1.0f / ((va + vb) + vc);
    const float s = vb * denom, t = vc * denom;
    const float q[3] = {(a[0] + ab[0] * s) + ac[0] * t, (a[1] + ab[1] * s) + ac[1] * t,
                        (a[2] + ab[2] * s) + ac[2] * t};
    return dist2_3(p, q);
}

__global__ void k_fill_u64(unsigned long long* p, int n, unsigned long long v)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

constexpr int NR = 64; // rig points per CTA (shared memory)

// grid: (vertex chunks, rig point tiles).  key = (bits(d2) << 32) | vertex index; d2 >= 0 so the bit pattern
// orders like the value, and the minimum key is the nearest vertex with the lowest index on ties.
__global__ void __launch_bounds__(256) k_nearest(const float* __restrict__ P, int64_t V, const float* __restrict__ rig,
                                                 int N, unsigned long long* __restrict__ keys)
{
    __shared__ float s_rig[NR][3];
    __shared__ unsigned long long s_key[NR];
    const int i0 = blockIdx.y * NR;
    const int cnt = min(NR, N - i0);
    for (int t = threadIdx.x; t < NR * 3; t += blockDim.x) s_rig[t / 3][t % 3] = (t / 3 < cnt) ? rig[3 * i0 + t] : 0.f;
    for (int t = threadIdx.x; t < NR; t += blockDim.x) s_key[t] = ~0ull;
    __syncthreads();
    const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    float p[3] = {0.f, 0.f, 0.f};
    if (v < V) {
        p[0] = P[3 * v];
        p[1] = P[3 * v + 1];
        p[2] = P[3 * v + 2];
    }
    for (int i = 0; i < cnt; ++i) {
        unsigned long long key = ~0ull;
        if (v < V) {
            const float d = dist2_3(s_rig[i], p);
            key = ((unsigned long long)__float_as_uint(d) << 32) | (unsigned long long)(unsigned)v;
        }
        for (int o = 16; o > 0; o >>= 1) {
            const unsigned long long other = __shfl_xor_sync(0xffffffffu, key, o);
            key = other < key ? other : key;
        }
        if ((threadIdx.x & 31) == 0 && key != ~0ull) atomicMin(&s_key[i], key);
    }
    __syncthreads();
    for (int t = threadIdx.x; t < cnt; t += blockDim.x)
        if (s_key[t] != ~0ull) atomicMin(&keys[i0 + t], s_key[t]);
}

__global__ void k_keys_to_idx(const unsigned long long* keys, int N, int32_t* idx)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < N) idx[i] = keys[i] == ~0ull ? -1 : (int32_t)(keys[i] & 0xffffffffull);
}

constexpr int TT = 128; // triangles per shared-memory stage

__global__ void __launch_bounds__(128) k_capture_dist(const float* __restrict__ P, int64_t V,
                                                      const uint8_t* __restrict__ member, const float* __restrict__ rig,
                                                      const int32_t* __restrict__ tri, int ntri, float radius,
                                                      int dofalloff, float* __restrict__ dist2)
{
    __shared__ float s_t[TT][9];
    __shared__ int s_seg[TT];
    const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = v < V && member[v] && dofalloff;
    float p[3] = {0.f, 0.f, 0.f};
    if (active) {
        p[0] = P[3 * v];
        p[1] = P[3 * v + 1];
        p[2] = P[3 * v + 2];
    }
    const float radius_sqrt = radius * radius; // capture.cpp:62
    float best = radius_sqrt;
    bool found = false;
    for (int t0 = 0; t0 < ntri; t0 += TT) {
        const int cnt = min(TT, ntri - t0);
        __syncthreads();
        for (int t = threadIdx.x; t < cnt; t += blockDim.x) {
            const int ia = tri[3 * (t0 + t)], ib = tri[3 * (t0 + t) + 1], ic = tri[3 * (t0 + t) + 2];
            s_seg[t] = ic < 0;
            for (int k = 0; k < 3; ++k) {
                s_t[t][k] = rig[3 * ia + k];
                s_t[t][3 + k] = rig[3 * ib + k];
                s_t[t][6 + k] = ic < 0 ? 0.f : rig[3 * ic + k];
            }
        }
        __syncthreads();
        if (active) {
            for (int t = 0; t < cnt; ++t) {
                const float d = s_seg[t] ? point_seg_dist2(p, &s_t[t][0], &s_t[t][3])
                                         : point_tri_dist2(p, &s_t[t][0], &s_t[t][3], &s_t[t][6]);
                if (d < best) {
                    best = d;
                    found = true;
                }
            }
        }
    }
    if (v < V) dist2[v] = active ? (found ? best : -1.0f) : 0.0f;
}

} // namespace

cudaError_t fd_launch_nearest(fd_ctx* ctx, const float* d_P, int64_t V, const float* d_rig, int N,
                              unsigned long long* keys, int32_t* d_nearest)
{
    if (N <= 0) return cudaSuccess;
    k_fill_u64<<<(N + 255) / 256, 256, 0, ctx->stream>>>(keys, N, ~0ull);
    ctx->launches += 1;
    if (V > 0) {
        dim3 grid((unsigned)((V + 255) / 256), (N + NR - 1) / NR);
        k_nearest<<<grid, 256, 0, ctx->stream>>>(d_P, V, d_rig, N, keys);
        ctx->launches += 1;
    }
    k_keys_to_idx<<<(N + 255) / 256, 256, 0, ctx->stream>>>(keys, N, d_nearest);
    ctx->launches += 1;
    return cudaGetLastError();
}

cudaError_t fd_launch_capture_dist(fd_ctx* ctx, const float* d_P, int64_t V, const uint8_t* d_member,
                                   const float* d_rig, const int32_t* d_tri, int ntri, float radius, int dofalloff,
                                   float* d_dist2)
{
    if (V <= 0) return cudaSuccess;
    k_capture_dist<<<(unsigned)((V + 127) / 128), 128, 0, ctx->stream>>>(d_P, V, d_member, d_rig, d_tri, ntri, radius,
                                                                        dofalloff, d_dist2);
    ctx->launches += 1;
    return cudaGetLastError();
}
