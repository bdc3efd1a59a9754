// fd_tc_common.cuh -- the PTX wrappers the tcgen05 evaluation kernels share (mbarrier, TMA, tcgen05.mma / ld, packed FP32 pairs)
// and the SOP epilogue pieces they apply per vertex.  Included inside each kernel's own namespace.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <stdint.h>

#include "fd_internal.h"

namespace tcc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}

// bounded wait: a protocol bug traps (an error the host sees) instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > (1u << 26)) __trap();
    }
}
// the same wait with its duration added to a counter (debug instantiation only)
template <bool DBG> __device__ __forceinline__ void mbar_wait_t(uint32_t bar, uint32_t parity, long long& acc)
{
    if (DBG) {
        const long long t0 = clock64();
        mbar_wait(bar, parity);
        acc += clock64() - t0;
    } else {
        mbar_wait(bar, parity);
    }
}
__device__ __forceinline__ bool elect_one()
{
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(uint32_t smem_dst, const CUtensorMap* map, uint32_t bar, int c0, int c1)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}

// the same load delivered to the same shared-memory offset of every CTA in cta_mask (each CTA's mbarrier at the same
// offset receives the bytes that landed in that CTA)
__device__ __forceinline__ void tma_load_2d_mc(uint32_t smem_dst, const CUtensorMap* map, uint32_t bar, int c0, int c1,
                                               uint16_t cta_mask)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
        ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "h"(cta_mask)
        : "memory");
}
__device__ __forceinline__ void cluster_sync_all()
{
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}

// 1-D bulk copy global -> shared, completion counted on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_load_1d(uint32_t smem_dst, const void* gsrc, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(gsrc)), "r"(bytes), "r"(bar)
                 : "memory");
}

// D[tmem] (+)= A[smem] * B[smem], M=128, K=16, FP16 inputs, FP32 accumulate
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// the arrive lands on the barrier at this offset in every CTA of cta_mask
__device__ __forceinline__ void umma_commit_mc(uint32_t bar, uint16_t cta_mask)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"(cta_mask) : "memory");
}

// shared-memory matrix descriptor, K-major, SWIZZLE_64B (rows of 64 bytes, 8-row groups 512 bytes apart):
// start address >> 4 | LBO (=1, unused for swizzled K-major) << 16 | SBO (512 B >> 4) << 32 | version 1 << 46 | layout 4 << 61
__device__ __forceinline__ uint64_t make_desc_sw64(uint32_t smem_addr)
{
    return (uint64_t)((smem_addr >> 4) & 0x3fff) | (1ull << 16) | (32ull << 32) | (1ull << 46) | (4ull << 61);
}
// instruction descriptor: D=F32 (bit 4), A=B=F16 (0), both K-major, N >> 3 at bit 17, M >> 4 at bit 24
__device__ __forceinline__ uint32_t make_idesc(int n_cols) { return (1u << 4) | ((uint32_t)(n_cols >> 3) << 17) | (8u << 24); }

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v)
{
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float* v)
{
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

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
// packed FP32 pairs (FADD2 / FMUL2 / FFMA2 on sm_100a): two basis functions per issue slot
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
    if (KERNEL == FD_KERNEL_GAUSSIAN) return ex2_approx(r2 * prm);
    if (KERNEL == FD_KERNEL_MULTIQUADRIC) return sqrt_approx(r2 + prm);
    return (0.34657359027997264f * r2) * lg2_approx(fmaxf(r2, 1e-37f));
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
// SOP_FaceDeform.hpp:28-41
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

// effective weight of row k, column c (affine rows re-expressed in the normalised coordinates)
__device__ __forceinline__ double tc_weight(const double* __restrict__ W, int ldw, int N, int np, int k, int c,
                                            const float* __restrict__ norm)
{
    if (k < N) return W[(size_t)k * ldw + c];
    if (np == 0) return 0.0;
    if (k == N) {
        double v = W[(size_t)N * ldw + c];
        if (np == 4)
            for (int a = 0; a < 3; ++a) v += W[(size_t)(N + 1 + a) * ldw + c] * (double)norm[a];
        return v;
    }
    if (np == 4 && k <= N + 3) return W[(size_t)k * ldw + c] / (double)norm[3];
    return 0.0;
}

// per column: power-of-two scale that brings max |w| into [2^(top_exp-1), 2^top_exp) (top_exp = 14: FP16 range with head-room
// for hi + lo; the exact-digit kernel asks for the width of its leading digit);
// CTA = 32 columns x 8 row groups, row-major reads stay coalesced
static __global__ void __launch_bounds__(256) k_tc_colscale(const double* __restrict__ W, int ldw, int N, int np, int ncol,
                                                     int ncol_pad, int phi_shift, int top_exp, const float* __restrict__ norm,
                                                     float* __restrict__ unscale, float* __restrict__ scale,
                                                     int* __restrict__ flags)
{
    __shared__ double s_mx[8][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + tx;
    double mx = 0.0;
    if (c < ncol) {
        // four independent running maxima: the loads of consecutive iterations overlap instead of one L2 round trip each
        double m4[4] = {0.0, 0.0, 0.0, 0.0};
        double chk = 0.0; // 0 * w stays 0 unless w is NaN or Inf (fmax would silently drop a NaN)
        int k = ty;
        for (; k + 24 < N; k += 32) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const double w = W[(size_t)(k + 8 * u) * ldw + c];
                m4[u] = fmax(m4[u], fabs(w));
                chk = fma(w, 0.0, chk);
            }
        }
        for (; k < N + 4; k += 8) {
            const double w = tc_weight(W, ldw, N, np, k, c, norm);
            m4[0] = fmax(m4[0], fabs(w));
            chk = fma(w, 0.0, chk);
        }
        mx = fmax(fmax(m4[0], m4[1]), fmax(m4[2], m4[3]));
        if (chk != 0.0) atomicExch(&flags[FD_FLAG_NONFINITE], 1); // NaN / Inf weights -> terminationtype -3
    }
    s_mx[ty][tx] = mx;
    __syncthreads();
    if (ty != 0 || c >= ncol_pad) return;
    for (int g = 1; g < 8; ++g) mx = fmax(mx, s_mx[g][tx]);
    int e = 0;
    if (mx > 0.0 && isfinite(mx)) {
        frexp(mx, &e);      // mx = m * 2^e, m in [0.5, 1)
        e = top_exp - e;    // mx * 2^e in [2^(top_exp-1), 2^top_exp)
        e = max(-60, min(60, e));
    }
    scale[c] = (float)ldexp(1.0, e);
    unscale[c] = (float)ldexp(1.0, -e - phi_shift); // also undoes the 2^phi_shift the kernel folds into Phi
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static inline EncodeTiledFn get_encode()
{
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

// P_out viewed as a [F][V*3] float tensor; one box = epi_frames frames x (32 vertices x 3 floats)
static inline bool make_out_map(CUtensorMap* map, float* P_out, int64_t V, int F, int epi_frames)
{
    EncodeTiledFn enc = get_encode();
    if (!enc) return false;
    cuuint64_t dims[2] = {(cuuint64_t)V * 3, (cuuint64_t)F};
    cuuint64_t strides[1] = {(cuuint64_t)V * 12};
    cuuint32_t box[2] = {96, (cuuint32_t)epi_frames};
    cuuint32_t estr[2] = {1, 1};
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, P_out, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static inline bool make_map(CUtensorMap* map, void* ptr, int Kpad, int rows, int box_k, int box_rows)
{
    EncodeTiledFn enc = get_encode();
    if (!enc) return false;
    cuuint64_t dims[2] = {(cuuint64_t)Kpad, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)Kpad * 2};
    cuuint32_t box[2] = {(cuuint32_t)box_k, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, ptr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

} // namespace tcc
