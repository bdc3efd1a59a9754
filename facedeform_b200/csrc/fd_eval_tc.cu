// fd_eval_tc.cu -- K3 (tensor-core path): fused vertex evaluation as a tcgen05 GEMM whose A operand is generated
// on the fly.
//
//   D[v][3f+k] = sum_j Phi[v][j] * W[j][3f+k],   Phi[v][j] = phi_j(|P[v] - c_j|)  (plus the rows [1, x', y', z']
//   that carry the affine block), followed by the SOP epilogue (gate, tangent projection, falloff, P += disp):
//   reference SOP_FaceDeform.cpp:404-439 and SOP_FaceDeform.hpp:28-41, for all F frames at once.
//
// Phi never exists in HBM.  One persistent CTA per SM walks "units" of 256 vertices x 240 columns (80 frames):
//   warp 0      TMA producer: streams the weight tile W^T[240 cols][32 k] (FP16 hi and lo parts, SWIZZLE_64B)
//   warp 1      MMA issuer  : one elected lane issues tcgen05.mma.cta_group::1.kind::f16 (M=128, N=240, K=16),
//                             FP32 accumulators in TMEM (two M tiles x 240 columns)
//   warps 2..9  Phi producers: thread t owns vertex row t; per stage it evaluates 32 basis functions on the FMA /
//                             MUFU pipes, splits each value into FP16 hi + lo and writes both into the canonical
//                             K-major SWIZZLE_64B shared-memory layout the UMMA descriptor expects;
//                             after the last stage the same warps drain TMEM (tcgen05.ld), run the epilogue and
//                             store coalesced float4 rows through a shared-memory transpose.
// Precision: hi*hi + hi*lo + lo*hi with FP16 splits keeps ~22 bits per factor (weights are pre-scaled per column by
// a power of two into FP16 range and un-scaled in the epilogue; coordinates of the affine rows are normalised to the
// control rig's bounding box), which holds the 1e-5 x bbox-diagonal tolerance of the FP32 path.
#include <cuda.h>
#include <cuda_fp16.h>

#include "fd_internal.h"

namespace tc {

constexpr int TM = 256;                       // vertices per unit (two M=128 accumulators)
constexpr int CB = 240;                       // columns per unit (multiple of 3 and of 16)
constexpr int BK = 32;                        // k per pipeline stage (64-byte rows -> SWIZZLE_64B)
constexpr int STAGES = 3;
constexpr int A_SPLIT_BYTES = TM * BK * 2;    // 16384
constexpr int B_SPLIT_BYTES = CB * BK * 2;    // 15360
constexpr int STAGE_BYTES = 2 * A_SPLIT_BYTES + 2 * B_SPLIT_BYTES; // 63488
constexpr int PRODUCER_WARPS = 8;
constexpr int THREADS = 32 * (2 + PRODUCER_WARPS);
constexpr int TMEM_COLS = 512;
constexpr int ACC1_COL = 256;                 // TMEM column of the second M tile's accumulator
constexpr int EPI_FRAMES = 16;                // frames per epilogue chunk (48 accumulator columns)
constexpr int EPI_WARP_FLOATS = EPI_FRAMES * 96; // staging floats per warp (16 frames x 32 vertices x 3)
constexpr int SMEM_BARRIERS = STAGES * STAGE_BYTES;
constexpr int SMEM_COLSCALE = SMEM_BARRIERS + 128;
constexpr int SMEM_TOTAL = SMEM_COLSCALE + CB * 4;
constexpr int SMEM_ALLOC = SMEM_TOTAL + 1024; // slack for the 1024-byte alignment of the dynamic window

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
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if ((++spins & 0x3fff) == 0 && clock64() - t0 > 20000000000ll) __trap();
    }
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

struct Args {
    const float4* ctab;     // N centres: (x, y, z, kernel parameter)
    const float* norm;      // (ox, oy, oz, s): affine-row coordinates x' = (x - o) * s
    const float* colscale;  // per column: multiply the accumulator by this to undo the FP16 pre-scaling
    int N, Kpad, F, ncb;
    const float* P;
    int64_t V;
    const float* dist2;
    const float* tu;
    const float* tv;
    const float* nrm;
    float* P_out;
    float* falloff_out;
    float radius2, falloffrate;
    int do_tangent;
    int vec_store_ok;       // V % 4 == 0 and P_out 16-byte aligned
};

// 8 basis values -> FP16 hi (RN) and lo (RN of the remainder), packed as two 16-byte chunks
__device__ __forceinline__ void split8(const float* ph, uint4& hi, uint4& lo)
{
    uint32_t h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const __half2 h2 = __floats2half2_rn(ph[2 * i], ph[2 * i + 1]);
        const float2 back = __half22float2(h2);
        const __half2 l2 = __floats2half2_rn(ph[2 * i] - back.x, ph[2 * i + 1] - back.y);
        h[i] = *reinterpret_cast<const uint32_t*>(&h2);
        l[i] = *reinterpret_cast<const uint32_t*>(&l2);
    }
    hi = make_uint4(h[0], h[1], h[2], h[3]);
    lo = make_uint4(l[0], l[1], l[2], l[3]);
}

template <int KERNEL>
__global__ void __launch_bounds__(THREADS, 1)
k_eval_tc(const Args a, const __grid_constant__ CUtensorMap map_hi, const __grid_constant__ CUtensorMap map_lo)
{
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const uint32_t smem_base = smem_u32(smem);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SMEM_BARRIERS);
    const uint32_t bar_full_a = smem_u32(bars + 0);          // [STAGES], count PRODUCER_WARPS
    const uint32_t bar_full_b = smem_u32(bars + STAGES);     // [STAGES], count 1 + tx bytes
    const uint32_t bar_empty = smem_u32(bars + 2 * STAGES);  // [STAGES], count 1 (tcgen05.commit)
    const uint32_t bar_tmem_full = smem_u32(bars + 3 * STAGES);
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 3 * STAGES + 1);
    float* s_colscale = reinterpret_cast<float*>(smem + SMEM_COLSCALE);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(bar_full_a + 8 * s, PRODUCER_WARPS);
            mbar_init(bar_full_b + 8 * s, 1);
            mbar_init(bar_empty + 8 * s, 1);
        }
        mbar_init(bar_tmem_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)), "r"(TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    const int nk = a.Kpad / BK;
    const int64_t n_vt = (a.V + TM - 1) / TM;
    const int64_t n_units = n_vt * a.ncb;

    if (warp == 0) {
        // ================= TMA producer: weight tiles =================
        if (lane == 0) {
            uint32_t it = 0;
            for (int64_t u = blockIdx.x; u < n_units; u += gridDim.x) {
                const int cb = (int)(u % a.ncb);
                for (int kb = 0; kb < nk; ++kb, ++it) {
                    const int s = it % STAGES;
                    const uint32_t ph = (it / STAGES) & 1;
                    mbar_wait(bar_empty + 8 * s, ph ^ 1);
                    const uint32_t sb = smem_base + s * STAGE_BYTES + 2 * A_SPLIT_BYTES;
                    mbar_expect_tx(bar_full_b + 8 * s, 2 * B_SPLIT_BYTES);
                    tma_load_2d(sb, &map_hi, bar_full_b + 8 * s, kb * BK, cb * CB);
                    tma_load_2d(sb + B_SPLIT_BYTES, &map_lo, bar_full_b + 8 * s, kb * BK, cb * CB);
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        if (lane == 0) {
            uint32_t it = 0;
            for (int64_t u = blockIdx.x; u < n_units; u += gridDim.x) {
                const int cb = (int)(u % a.ncb);
                const int ncols = min(CB, (3 * a.F - cb * CB + 15) & ~15);
                const uint32_t idesc = make_idesc(ncols);
                for (int kb = 0; kb < nk; ++kb, ++it) {
                    const int s = it % STAGES;
                    const uint32_t ph = (it / STAGES) & 1;
                    mbar_wait(bar_full_a + 8 * s, ph);
                    mbar_wait(bar_full_b + 8 * s, ph);
                    tc_fence_after();
                    const uint32_t sa = smem_base + s * STAGE_BYTES;
                    const uint32_t sb = sa + 2 * A_SPLIT_BYTES;
#pragma unroll
                    for (int kk = 0; kk < BK / 16; ++kk) {
#pragma unroll
                        for (int mt = 0; mt < 2; ++mt) {
                            const uint32_t a_off = mt * (128 * BK * 2) + kk * 32;
                            const uint64_t a_hi = make_desc_sw64(sa + a_off);
                            const uint64_t a_lo = make_desc_sw64(sa + A_SPLIT_BYTES + a_off);
                            const uint64_t b_hi = make_desc_sw64(sb + kk * 32);
                            const uint64_t b_lo = make_desc_sw64(sb + B_SPLIT_BYTES + kk * 32);
                            const uint32_t d = tmem_base + mt * ACC1_COL;
                            umma_f16(d, a_hi, b_hi, idesc, (kb | kk) != 0);
                            umma_f16(d, a_hi, b_lo, idesc, 1);
                            umma_f16(d, a_lo, b_hi, idesc, 1);
                        }
                    }
                    umma_commit(bar_empty + 8 * s);               // frees the stage when these MMAs have read it
                    if (kb == nk - 1) umma_commit(bar_tmem_full); // accumulators complete
                }
            }
        }
    } else {
        // ================= Phi producers, then epilogue =================
        const int pw = warp - 2;              // 0..7
        const int mt = pw >> 2;               // M tile
        const int q = warp & 3;               // TMEM lane quarter this warp may access
        const int row = mt * 128 + q * 32 + lane;
        const float4 nrm4 = *reinterpret_cast<const float4*>(a.norm);
        float* stg = reinterpret_cast<float*>(smem + (pw < 5 ? 0 : STAGE_BYTES)) + (pw < 5 ? pw : pw - 5) * EPI_WARP_FLOATS;
        uint32_t it = 0, unit_iter = 0;
        for (int64_t u = blockIdx.x; u < n_units; u += gridDim.x, ++unit_iter) {
            const int cb = (int)(u % a.ncb);
            const int64_t vt = u / a.ncb;
            const int64_t v = vt * TM + row;
            const bool valid = v < a.V;
            float px = 0.f, py = 0.f, pz = 0.f;
            if (valid) {
                px = a.P[3 * v];
                py = a.P[3 * v + 1];
                pz = a.P[3 * v + 2];
            }
            for (int t = threadIdx.x - 64; t < CB; t += 32 * PRODUCER_WARPS) s_colscale[t] = a.colscale[cb * CB + t];

            // ---- produce the Phi tile, 32 columns of K per stage ----
            for (int kb = 0; kb < nk; ++kb, ++it) {
                const int s = it % STAGES;
                const uint32_t ph = (it / STAGES) & 1;
                mbar_wait(bar_empty + 8 * s, ph ^ 1);
                uint8_t* a_hi = smem + s * STAGE_BYTES + row * (BK * 2);
                uint8_t* a_lo = a_hi + A_SPLIT_BYTES;
                const int k0 = kb * BK;
                const int swz = (row >> 1) & 3;
#pragma unroll
                for (int c16 = 0; c16 < 4; ++c16) {
                    float f[8];
                    if (k0 + c16 * 8 + 8 <= a.N) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const float4 c = __ldg(&a.ctab[k0 + c16 * 8 + j]);
                            const float dx = px - c.x, dy = py - c.y, dz = pz - c.z;
                            f[j] = phi<KERNEL>(dx * dx + dy * dy + dz * dz, c.w);
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const int k = k0 + c16 * 8 + j;
                            float val = 0.f;
                            if (k < a.N) {
                                const float4 c = __ldg(&a.ctab[k]);
                                const float dx = px - c.x, dy = py - c.y, dz = pz - c.z;
                                val = phi<KERNEL>(dx * dx + dy * dy + dz * dz, c.w);
                            } else if (k == a.N) {
                                val = 1.0f;
                            } else if (k == a.N + 1) {
                                val = (px - nrm4.x) * nrm4.w;
                            } else if (k == a.N + 2) {
                                val = (py - nrm4.y) * nrm4.w;
                            } else if (k == a.N + 3) {
                                val = (pz - nrm4.z) * nrm4.w;
                            }
                            f[j] = val;
                        }
                    }
                    uint4 hi, lo;
                    split8(f, hi, lo);
                    const int off = (c16 ^ swz) * 16;
                    *reinterpret_cast<uint4*>(a_hi + off) = hi;
                    *reinterpret_cast<uint4*>(a_lo + off) = lo;
                }
                fence_proxy_async(); // generic-proxy stores -> visible to the tensor core (async proxy)
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_full_a + 8 * s);
            }

            // ---- epilogue: TMEM -> registers -> (transpose in shared memory) -> global ----
            float fo = 0.f;
            bool skip = true;
            if (valid) {
                const float d2 = a.dist2 ? a.dist2[v] : 0.f;
                skip = d2 > a.radius2;                                  // SOP_FaceDeform.cpp:408-410
                fo = powf(1.0f - fminf(d2 / a.radius2, 1.0f), a.falloffrate); // :423-424
                if (skip) fo = 0.f;
                if (a.falloff_out && cb == 0) a.falloff_out[v] = fo;
            }
            float tu[3] = {0, 0, 0}, tv[3] = {0, 0, 0}, tn[3] = {0, 0, 0};
            if (a.do_tangent && valid) {
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
            const int64_t v_warp0 = vt * TM + mt * 128 + q * 32;
            const bool vec = a.vec_store_ok && (v_warp0 + 32 <= a.V);
            const int f_base = cb * (CB / 3);
            const int nframes = min(CB / 3, a.F - f_base);
            asm volatile("bar.sync 1, 256;" ::: "memory"); // s_colscale visible to all producer warps
            mbar_wait(bar_tmem_full, unit_iter & 1);
            tc_fence_after();
            for (int ch = 0; ch * EPI_FRAMES < nframes; ++ch) {
                float acc[48];
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + mt * ACC1_COL + ch * 48;
                tmem_ld16(taddr, acc);
                tmem_ld16(taddr + 16, acc + 16);
                tmem_ld16(taddr + 32, acc + 32);
                tmem_ld_wait();
                const int fcnt = min(EPI_FRAMES, nframes - ch * EPI_FRAMES);
#pragma unroll
                for (int i = 0; i < EPI_FRAMES; ++i) {
                    float d[3];
#pragma unroll
                    for (int k = 0; k < 3; ++k) d[k] = acc[3 * i + k] * s_colscale[ch * 48 + 3 * i + k];
                    if (a.do_tangent) project_to_tangents(tu, tv, tn, d);
                    float o[3] = {skip ? px : px + d[0] * fo, skip ? py : py + d[1] * fo, skip ? pz : pz + d[2] * fo};
                    if (vec) {
#pragma unroll
                        for (int k = 0; k < 3; ++k) stg[(i * 32 + lane) * 3 + k] = o[k];
                    } else if (valid && i < fcnt) {
                        float* dst = a.P_out + ((size_t)(f_base + ch * EPI_FRAMES + i) * (size_t)a.V + (size_t)v) * 3;
                        dst[0] = o[0];
                        dst[1] = o[1];
                        dst[2] = o[2];
                    }
                }
                if (vec) {
                    __syncwarp();
                    const float4* stg4 = reinterpret_cast<const float4*>(stg);
#pragma unroll
                    for (int r = 0; r < 12; ++r) {
                        const int q4 = r * 32 + lane;
                        const int fr = q4 / 24, w = q4 - fr * 24;
                        if (fr < fcnt) {
                            float4* dst = reinterpret_cast<float4*>(
                                a.P_out + ((size_t)(f_base + ch * EPI_FRAMES + fr) * (size_t)a.V + (size_t)v_warp0) * 3);
                            dst[w] = stg4[fr * 24 + w];
                        }
                    }
                    __syncwarp();
                }
            }
            tc_fence_before();
            asm volatile("bar.sync 1, 256;" ::: "memory"); // TMEM and the staging area are free for the next unit
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS));
    }
}

// ------------------------------------------------------------------------------------------------------------
// pack: FP64 weights -> transposed, column-scaled FP16 hi/lo tables W^T[c][k] the TMA streams
// ------------------------------------------------------------------------------------------------------------

// bounding box of the control points -> (centre, 2 / largest extent)
__global__ void __launch_bounds__(256) k_tc_norm(const float* __restrict__ rest, int N, float* __restrict__ norm)
{
    __shared__ float s_lo[3][256], s_hi[3][256];
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int i = threadIdx.x; i < N; i += 256)
        for (int k = 0; k < 3; ++k) {
            lo[k] = fminf(lo[k], rest[3 * i + k]);
            hi[k] = fmaxf(hi[k], rest[3 * i + k]);
        }
    for (int k = 0; k < 3; ++k) {
        s_lo[k][threadIdx.x] = lo[k];
        s_hi[k][threadIdx.x] = hi[k];
    }
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o)
            for (int k = 0; k < 3; ++k) {
                s_lo[k][threadIdx.x] = fminf(s_lo[k][threadIdx.x], s_lo[k][threadIdx.x + o]);
                s_hi[k][threadIdx.x] = fmaxf(s_hi[k][threadIdx.x], s_hi[k][threadIdx.x + o]);
            }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        float ext = 0.f;
        for (int k = 0; k < 3; ++k) {
            norm[k] = 0.5f * (s_lo[k][0] + s_hi[k][0]);
            ext = fmaxf(ext, s_hi[k][0] - s_lo[k][0]);
        }
        norm[3] = ext > 0.f ? 2.0f / ext : 1.0f;
    }
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

// per column: power-of-two scale that brings max |w| to <= 16384 (FP16 range with head-room for hi + lo)
__global__ void __launch_bounds__(256) k_tc_colscale(const double* __restrict__ W, int ldw, int N, int np, int ncol,
                                                     int ncol_pad, const float* __restrict__ norm,
                                                     float* __restrict__ unscale, float* __restrict__ scale)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncol_pad) return;
    double mx = 0.0;
    if (c < ncol)
        for (int k = 0; k < N + 4; ++k) mx = fmax(mx, fabs(tc_weight(W, ldw, N, np, k, c, norm)));
    int e = 0;
    if (mx > 0.0 && isfinite(mx)) {
        frexp(mx, &e);      // mx = m * 2^e, m in [0.5, 1)
        e = 14 - e;         // mx * 2^e in [8192, 16384)
        e = max(-60, min(60, e));
    }
    scale[c] = (float)ldexp(1.0, e);
    unscale[c] = (float)ldexp(1.0, -e);
}

// W^T hi/lo [ncol_pad][Kpad] FP16, k contiguous; 32x32 tile transpose through shared memory
__global__ void __launch_bounds__(256) k_tc_pack(const double* __restrict__ W, int ldw, int N, int np, int ncol,
                                                 int ncol_pad, int Kpad, const float* __restrict__ norm,
                                                 const float* __restrict__ scale, __half* __restrict__ Wt_hi,
                                                 __half* __restrict__ Wt_lo)
{
    __shared__ float s_t[32][33];
    const int c0 = blockIdx.x * 32, k0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int r = ty; r < 32; r += 8) {
        const int k = k0 + r, c = c0 + tx;
        float v = 0.f;
        if (c < ncol && k < Kpad) v = (float)(tc_weight(W, ldw, N, np, k, c, norm) * (double)scale[c]);
        s_t[r][tx] = v;
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const int c = c0 + r, k = k0 + tx;
        if (c < ncol_pad && k < Kpad) {
            const float v = s_t[tx][r];
            const __half h = __float2half_rn(v);
            const __half l = __float2half_rn(v - __half2float(h));
            Wt_hi[(size_t)c * Kpad + k] = h;
            Wt_lo[(size_t)c * Kpad + k] = l;
        }
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode()
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

static bool make_map(CUtensorMap* map, void* ptr, int Kpad, int rows)
{
    EncodeTiledFn enc = get_encode();
    if (!enc) return false;
    cuuint64_t dims[2] = {(cuuint64_t)Kpad, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)Kpad * 2};
    cuuint32_t box[2] = {BK, CB};
    cuuint32_t estr[2] = {1, 1};
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, ptr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

} // namespace tc

static_assert(sizeof(CUtensorMap) == FD_TMAP_BYTES, "fd_model reserves FD_TMAP_BYTES per tensor map");

int fd_tc_kpad(int N) { return fd_round_up(N + 4, tc::BK); }
int fd_tc_ncb(int F) { return (3 * F + tc::CB - 1) / tc::CB; }
int fd_tc_col_pad(int F) { return fd_tc_ncb(F) * tc::CB; }

// builds the tensor-path tables for the weights currently in m->d_W (called from fd_launch_pack)
cudaError_t fd_launch_pack_tc(fd_ctx* ctx, fd_model* m)
{
    cudaStream_t s = ctx->stream;
    const int ncol = 3 * m->F, ncol_pad = fd_tc_col_pad(m->F), Kpad = fd_tc_kpad(m->N);
    tc::k_tc_norm<<<1, 256, 0, s>>>(m->d_rest, m->N, m->d_tc_norm);
    tc::k_tc_colscale<<<(ncol_pad + 255) / 256, 256, 0, s>>>(m->d_W, m->ldw, m->N, m->np, ncol, ncol_pad, m->d_tc_norm,
                                                            m->d_tc_unscale, m->d_tc_scale);
    dim3 grid((ncol_pad + 31) / 32, (Kpad + 31) / 32);
    tc::k_tc_pack<<<grid, 256, 0, s>>>(m->d_W, m->ldw, m->N, m->np, ncol, ncol_pad, Kpad, m->d_tc_norm, m->d_tc_scale,
                                       (__half*)m->d_tc_wt_hi, (__half*)m->d_tc_wt_lo);
    ctx->launches += 3;
    if (!tc::make_map((CUtensorMap*)m->tc_map_hi, m->d_tc_wt_hi, Kpad, ncol_pad) ||
        !tc::make_map((CUtensorMap*)m->tc_map_lo, m->d_tc_wt_lo, Kpad, ncol_pad))
        return cudaErrorInvalidValue;
    return cudaGetLastError();
}

cudaError_t fd_launch_eval_tc(fd_ctx* ctx, const fd_model* m, const float* P, int64_t V, const float* dist2,
                              const float* tu, const float* tv, const float* nrm, float* P_out, float* falloff_out)
{
    if (V <= 0) return cudaSuccess;
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(tc::k_eval_tc<FD_KERNEL_GAUSSIAN>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SMEM_ALLOC);
        cudaFuncSetAttribute(tc::k_eval_tc<FD_KERNEL_MULTIQUADRIC>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SMEM_ALLOC);
        cudaFuncSetAttribute(tc::k_eval_tc<FD_KERNEL_THINPLATE>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SMEM_ALLOC);
        attr_set = true;
    }
    tc::Args a;
    a.ctab = m->d_ctab32;
    a.norm = m->d_tc_norm;
    a.colscale = m->d_tc_unscale;
    a.N = m->N;
    a.Kpad = fd_tc_kpad(m->N);
    a.F = m->F;
    a.ncb = fd_tc_ncb(m->F);
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
    a.do_tangent = (m->prm.tangent && tu && tv && nrm) ? 1 : 0;
    a.vec_store_ok = (V % 4 == 0) && ((reinterpret_cast<uintptr_t>(P_out) & 15) == 0);
    const int64_t n_units = ((V + tc::TM - 1) / tc::TM) * a.ncb;
    const int grid = (int)(n_units < ctx->sm_count ? n_units : ctx->sm_count);
    const CUtensorMap& mh = *(const CUtensorMap*)m->tc_map_hi;
    const CUtensorMap& ml = *(const CUtensorMap*)m->tc_map_lo;
    switch (m->prm.kernel) {
    case FD_KERNEL_GAUSSIAN:
        tc::k_eval_tc<FD_KERNEL_GAUSSIAN><<<grid, tc::THREADS, tc::SMEM_ALLOC, ctx->stream>>>(a, mh, ml);
        break;
    case FD_KERNEL_MULTIQUADRIC:
        tc::k_eval_tc<FD_KERNEL_MULTIQUADRIC><<<grid, tc::THREADS, tc::SMEM_ALLOC, ctx->stream>>>(a, mh, ml);
        break;
    default:
        tc::k_eval_tc<FD_KERNEL_THINPLATE><<<grid, tc::THREADS, tc::SMEM_ALLOC, ctx->stream>>>(a, mh, ml);
        break;
    }
    ctx->launches += 1;
    return cudaGetLastError();
}
