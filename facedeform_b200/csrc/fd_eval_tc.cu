// fd_eval_tc.cu -- K3 (tensor-core path): fused vertex evaluation as a tcgen05 GEMM whose A operand is generated
// on the fly.
//
//   D[v][3f+k] = sum_j Phi[v][j] * W[j][3f+k],   Phi[v][j] = phi_j(|P[v] - c_j|)  (plus the rows [1, x', y', z']
//   that carry the affine block), followed by the SOP epilogue (gate, tangent projection, falloff, P += disp):
//   reference SOP_FaceDeform.cpp:404-439 and SOP_FaceDeform.hpp:28-41, for all F frames at once.
//
// Phi never exists in HBM.  One persistent CTA per SM (26 warps) walks "units" of 128 vertices x 240 columns (80 frames):
//   warps 0..15  Phi producers: two groups of 8 warps alternate pipeline stages; inside a group two threads share a
//                              vertex row and evaluate 16 basis functions each per stage on the FMA / MUFU pipes
//                              (packed FP32 pairs), split each value into FP16 hi + lo and write both into the K-major
//                              SWIZZLE_64B shared-memory layout the UMMA descriptor expects
//   warps 16..23 epilogue    : drain one of the two ping-pong TMEM accumulators (tcgen05.ld) while the MMAs of the
//                              next unit fill the other, un-scale, apply falloff, add P, transpose through shared
//                              memory and hand [8 frames x 32 vertices x 3] tiles to TMA stores
//   warp 24      TMA producer: streams the weight tiles W^T[240 cols][32 k] (FP16 hi and lo, SWIZZLE_64B, 4-stage ring;
//                              CTA pairs load half a tile each and multicast it) and, in a 16-deep ring of their own,
//                              the 32-centre tiles
//   warp 25      MMA issuer  : one elected lane issues tcgen05.mma.cta_group::1.kind::f16 (M=128, N=240, K=16),
//                              FP32 accumulators in TMEM (2 x 240 of 512 columns)
// Precision: hi*hi + hi*lo + lo*hi with FP16 splits keeps ~22 bits per factor (weights are pre-scaled per column by
// a power of two into FP16 range and un-scaled in the epilogue; coordinates of the affine rows are normalised to the
// control rig's bounding box).  The error of any FP32 evaluation scales with the cancellation in sum_j w_j phi_j; the
// bound and the rule by which FD_EVAL_AUTO leaves this kernel for the FP64 one are in DESIGN.md section 2.
#include <cuda.h>
#include <cuda_fp16.h>
#include <stdlib.h>
#include <string.h>

#include "fd_internal.h"
#include "fd_tc_common.cuh"

namespace tc {

// The Gaussian basis lives in (0, 1]: its FP16 "lo" parts would fall into the FP16 subnormal range (absolute floor
// 2^-25 on Phi, times sum |w| of thousands of centres).  Phi is therefore multiplied by 2^14 before the split (one
// FMUL2 per pair; adding 14 to the exponent instead would cost the low bits of t near Phi = 1) and the factor is
// undone by the column un-scaling.
constexpr int GAUSS_SHIFT = 14;
constexpr int TM = 128;                       // vertices per unit (one M=128 accumulator; two accumulators ping-pong)
constexpr int CB = 240;                       // columns per unit (multiple of 3 and of 16)
constexpr int BK = 32;                        // k per pipeline stage (64-byte rows -> SWIZZLE_64B)
constexpr int STAGES = 4;
constexpr int A_SPLIT_BYTES = TM * BK * 2;    // 8192
constexpr int B_SPLIT_BYTES = CB * BK * 2;    // 15360
constexpr int C_TILE_BYTES = BK * 16;         // a stage's 32 centres (16 pairs of two float4)
constexpr int CDEPTH = 16;                    // centre tiles travel in their own, deeper ring: Phi production never waits
                                              // for the 30 KB weight tile of its stage, only for a free A slot
constexpr int STAGE_BYTES = 2 * A_SPLIT_BYTES + 2 * B_SPLIT_BYTES; // 47104
constexpr int PRODUCER_WARPS = 16;            // two groups of 8 warps: group g fills the stages with (iteration & 1) == g;
                                              // inside a group two threads share a vertex row, 16 basis functions each
constexpr int EPILOGUE_WARPS = 8;              // two per TMEM lane quarter (even / odd column chunks)
constexpr int THREADS = 32 * (2 + PRODUCER_WARPS + EPILOGUE_WARPS);
// Warp roles.  The SM's warp arbiter favours the highest warp ids, so the two single-lane roles every other warp
// waits on (TMA producer, MMA issuer) sit at the top: warps 0..15 Phi producers, 16..23 epilogue, 24 TMA, 25 MMA.
constexpr int WARP_EPI0 = PRODUCER_WARPS;
constexpr int WARP_TMA = PRODUCER_WARPS + EPILOGUE_WARPS;
constexpr int WARP_MMA = WARP_TMA + 1;
constexpr int TMEM_COLS = 512;
constexpr int ACC1_COL = 256;                 // TMEM column of the second (ping-pong) accumulator
constexpr int EPI_FRAMES = 8;                 // frames per epilogue chunk (24 accumulator columns)
constexpr int EPI_WARP_FLOATS = EPI_FRAMES * 96; // staging floats per warp (8 frames x 32 vertices x 3)
constexpr int EPI_COLS = EPI_FRAMES * 3;
constexpr int SMEM_EPI_STAGING = STAGES * STAGE_BYTES;                          // 4 warps x 6 KB transpose buffers
constexpr int SMEM_BARRIERS = SMEM_EPI_STAGING + EPILOGUE_WARPS * EPI_WARP_FLOATS * 4;
constexpr int SMEM_CENTRES = SMEM_BARRIERS + 512;
constexpr int SMEM_COLSCALE = SMEM_CENTRES + CDEPTH * C_TILE_BYTES;
constexpr int COLSCALE_RESIDENT_BLOCKS = 6;   // column scales of up to 6 column blocks (F <= 480) stay resident
constexpr int SMEM_TOTAL = SMEM_COLSCALE + COLSCALE_RESIDENT_BLOCKS * CB * 4;
static_assert(SMEM_TOTAL + 1024 <= 227 * 1024, "shared-memory budget");
constexpr int SMEM_ALLOC = SMEM_TOTAL + 1024; // slack for the 1024-byte alignment of the dynamic window

using namespace tcc;

struct Args {
    const float4* ctab;     // centre pairs, two float4 per pair: (x0, x1, y0, y1), (z0, z1, prm0, prm1); padded to 32 centres
    const float* norm;      // (ox, oy, oz, s): affine-row coordinates x' = (x - o) * s
    const float* colscale;  // per column: multiply the accumulator by this to undo the FP16 pre-scaling
    int N, Kpad, Ktot, F, ncb; // Ktot = N + polynomial rows
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
    long long* dbg;         // optional per-unit phase timestamps of CTA 0 (FD_TC_DEBUG=1), else NULL
    int pair;               // 1: CTAs run as clusters of two that share every weight tile (each loads half, multicast)
    const int* sel;         // FD_EVAL_AUTO: the chosen evaluation kernel (device word) or NULL; the launch returns at once
    int sel_id;             // when *sel != sel_id
    int dbg_mode;           // FD_TC_DEBUG bits (debug instantiation only; results are then garbage): 2 skip the epilogue
                            // math + stores, 4 skip the Phi computation, 8 issue one MMA of three, 16 skip the TMA stores only
};

// 8 basis values -> FP16 hi (RN) and lo (RN of the remainder), packed as two 16-byte chunks
__device__ __forceinline__ void split8(const uint64_t* pf, uint4& hi, uint4& lo)
{
    uint32_t h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float f0, f1, r0, r1;
        unpack2(pf[i], f0, f1);
        const __half2 h2 = __floats2half2_rn(f0, f1);
        const float2 back = __half22float2(h2);
        unpack2(sub2(pf[i], pack2(back.x, back.y)), r0, r1); // both remainders in one FADD2
        const __half2 l2 = __floats2half2_rn(r0, r1);
        h[i] = *reinterpret_cast<const uint32_t*>(&h2);
        l[i] = *reinterpret_cast<const uint32_t*>(&l2);
    }
    hi = make_uint4(h[0], h[1], h[2], h[3]);
    lo = make_uint4(l[0], l[1], l[2], l[3]);
}

template <int KERNEL, bool TANGENT, bool DBG>
__global__ void __launch_bounds__(THREADS, 1)
k_eval_tc(const Args a, const __grid_constant__ CUtensorMap map_hi, const __grid_constant__ CUtensorMap map_lo,
          const __grid_constant__ CUtensorMap map_out)
{
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    if (a.sel && *a.sel != a.sel_id) return; // uniform over the grid (and the cluster): nothing has been set up yet
    // align inside the shared window with pointer arithmetic only: a uintptr_t round trip would demote every
    // access below to generic LD/ST (long-scoreboard) instead of LDS/STS
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const uint32_t smem_base = smem_u32(smem);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SMEM_BARRIERS);
    const uint32_t bar_full_a = smem_u32(bars + 0);          // [STAGES], count PRODUCER_WARPS
    const uint32_t bar_full_b = smem_u32(bars + STAGES);     // [STAGES], count 1 + tx bytes
    const uint32_t bar_empty = smem_u32(bars + 2 * STAGES);  // [STAGES], count 1 (tcgen05.commit)
    const uint32_t bar_tmem_full = smem_u32(bars + 3 * STAGES);        // [2], count 1: all MMAs of the unit retired
    const uint32_t bar_tmem_empty = smem_u32(bars + 3 * STAGES + 2);   // [2], count EPILOGUE_WARPS: accumulator drained
    const uint32_t bar_cfull = smem_u32(bars + 3 * STAGES + 4);            // [CDEPTH], count 1 + tx bytes: centre tile landed
    const uint32_t bar_cempty = smem_u32(bars + 3 * STAGES + 4 + CDEPTH);  // [CDEPTH], count PRODUCER_WARPS / 2: tile consumed
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 3 * STAGES + 4 + 2 * CDEPTH);
    float* s_colscale = reinterpret_cast<float*>(smem + SMEM_COLSCALE);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(bar_full_a + 8 * s, PRODUCER_WARPS / 2);
            mbar_init(bar_full_b + 8 * s, 1);
            mbar_init(bar_empty + 8 * s, a.pair ? 2 : 1); // pair: the MMAs of both CTAs must have read the slot
        }
        mbar_init(bar_tmem_full, 1);
        mbar_init(bar_tmem_full + 8, 1);
        mbar_init(bar_tmem_empty, EPILOGUE_WARPS);
        mbar_init(bar_tmem_empty + 8, EPILOGUE_WARPS);
        for (int c = 0; c < CDEPTH; ++c) {
            mbar_init(bar_cfull + 8 * c, 1);
            mbar_init(bar_cempty + 8 * c, PRODUCER_WARPS / 2);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == WARP_TMA) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)), "r"(TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    tc_fence_before();
    __syncthreads();
    if (a.pair) cluster_sync_all(); // the peer's barriers exist before anything is multicast at them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    // Unit walk.  Alone: CTA b takes units b, b + grid, ... (unit = vertex tile x column block).  As a pair: both CTAs
    // of a cluster take the same column block of two neighbouring vertex tiles in lockstep, so one copy of every
    // weight tile serves both (half of it loaded by each CTA and multicast): half the L2 -> SM stream.
    const uint32_t crank = a.pair ? cluster_ctarank() : 0;
    const int64_t unit0 = a.pair ? (blockIdx.x >> 1) : blockIdx.x;
    const int64_t ustride = a.pair ? (gridDim.x >> 1) : gridDim.x;
    const int nk = a.Kpad / BK;
    const int tail_ksteps = (a.Ktot - (nk - 1) * BK + 15) >> 4; // K=16 steps of the last stage that hold real rows (1 or 2)
    const int64_t n_vt = (a.V + TM - 1) / TM;
    const int64_t n_units = (a.pair ? (n_vt + 1) / 2 : n_vt) * a.ncb;
#define FD_UNIT_VT(u) (a.pair ? 2 * ((u) / a.ncb) + crank : (u) / a.ncb)

    // The two issuing roles run their loops with the whole warp (waits included) and elect one lane only around the
    // issue itself: control flow and addresses stay warp-uniform, so descriptors and barrier addresses live in
    // uniform registers (a loop nested inside `if (lane == 0)` costs ~100 dependent R2UR / PLOP3 instructions per stage
    // on a single lane -- measured ~900 cycles per stage, more than the 720 cycles of the stage's MMAs).
    if (warp == WARP_TMA) {
        // ================= TMA producer: weight tiles + centre tiles =================
        long long w_ce = 0, w_e = 0;
        const long long t_begin = DBG ? clock64() : 0;
        uint32_t it = 0, ic = 0;  // weight stages / centre tiles issued so far
        int kc = 0;               // k block of centre tile ic
        const uint32_t my_units = (uint32_t)((n_units - unit0 + ustride - 1) / ustride);
        const uint32_t total = my_units * (uint32_t)nk;
        for (int64_t u = unit0; u < n_units; u += ustride) {
            const int cb = (int)(u % a.ncb);
            for (int kb = 0; kb < nk; ++kb, ++it) {
                // Centre tiles run up to CDEPTH stages ahead of the weights.  The refill never blocks: a slot that is
                // still being read is retried at the next stage, so the weight tile of this stage is issued the moment
                // its slot frees (a blocking wait here tied every weight load to the producers' progress).  The ring
                // then holds the CDEPTH oldest unconsumed tiles, which are the ones the producers need next.
                while (ic < total && ic < it + CDEPTH) {
                    const int c = ic % CDEPTH;
                    if (!mbar_try_wait(bar_cempty + 8 * c, ((ic / CDEPTH) & 1) ^ 1)) break;
                    if (elect_one()) {
                        mbar_expect_tx(bar_cfull + 8 * c, C_TILE_BYTES);
                        bulk_load_1d(smem_base + SMEM_CENTRES + c * C_TILE_BYTES, a.ctab + kc * BK, C_TILE_BYTES, bar_cfull + 8 * c);
                    }
                    ++ic;
                    if (++kc == nk) kc = 0;
                }
                const int s = it % STAGES;
                const uint32_t ph = (it / STAGES) & 1;
                mbar_wait_t<DBG>(bar_empty + 8 * s, ph ^ 1, w_e);
                if (elect_one()) {
                    const uint32_t sb = smem_base + s * STAGE_BYTES + 2 * A_SPLIT_BYTES;
                    mbar_expect_tx(bar_full_b + 8 * s, 2 * B_SPLIT_BYTES);
                    constexpr int HB = B_SPLIT_BYTES / 2; // the tensor-map box is half a tile (CB / 2 columns)
                    if (a.pair) { // this CTA's half of the tile, delivered to both CTAs
                        tma_load_2d_mc(sb + crank * HB, &map_hi, bar_full_b + 8 * s, kb * BK, cb * CB + crank * (CB / 2), 3);
                        tma_load_2d_mc(sb + B_SPLIT_BYTES + crank * HB, &map_lo, bar_full_b + 8 * s, kb * BK,
                                       cb * CB + crank * (CB / 2), 3);
                    } else {
                        tma_load_2d(sb, &map_hi, bar_full_b + 8 * s, kb * BK, cb * CB);
                        tma_load_2d(sb + HB, &map_hi, bar_full_b + 8 * s, kb * BK, cb * CB + CB / 2);
                        tma_load_2d(sb + B_SPLIT_BYTES, &map_lo, bar_full_b + 8 * s, kb * BK, cb * CB);
                        tma_load_2d(sb + B_SPLIT_BYTES + HB, &map_lo, bar_full_b + 8 * s, kb * BK, cb * CB + CB / 2);
                    }
                }
                __syncwarp();
            }
        }
        if (DBG && a.dbg && blockIdx.x == 0 && lane == 0) {
            a.dbg[210] = clock64() - t_begin;
            a.dbg[211] = w_ce;
            a.dbg[212] = w_e;
        }
    } else if (warp == WARP_MMA) {
        // ================= MMA issuer =================
        long long w_t = 0, w_a = 0, w_b = 0;
        const long long t_begin = DBG ? clock64() : 0;
        uint32_t it = 0, unit_iter = 0;
        const uint64_t desc0 = make_desc_sw64(smem_base); // A hi tile of stage 0 (the whole window is < 256 KB: no carry)
        const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0); // warp-uniform copy
        for (int64_t u = unit0; u < n_units; u += ustride, ++unit_iter) {
            const int cb = (int)(u % a.ncb);
            const int ncols = min(CB, (3 * a.F - cb * CB + 15) & ~15);
            const uint32_t idesc = make_idesc(ncols);
            const int ab = unit_iter & 1;                 // accumulator buffer of this unit
            const uint32_t d = tmem_u + ab * ACC1_COL;
            // the epilogue must have drained this buffer (two units ago)
            mbar_wait_t<DBG>(bar_tmem_empty + 8 * ab, ((unit_iter >> 1) & 1) ^ 1, w_t);
            tc_fence_after();
            for (int kb = 0; kb < nk; ++kb, ++it) {
                const int s = it % STAGES;
                const uint32_t ph = (it / STAGES) & 1;
                mbar_wait_t<DBG>(bar_full_a + 8 * s, ph, w_a);
                mbar_wait_t<DBG>(bar_full_b + 8 * s, ph, w_b);
                tc_fence_after();
                if (elect_one()) {
                    // descriptors differ from the stage-0 ones only in the start-address field (units of 16 bytes)
                    const uint64_t a_hi = desc0 + (uint64_t)(s * (STAGE_BYTES >> 4));
                    const uint64_t a_lo = a_hi + (A_SPLIT_BYTES >> 4);
                    const uint64_t b_hi = a_hi + (2 * A_SPLIT_BYTES >> 4);
                    const uint64_t b_lo = b_hi + (B_SPLIT_BYTES >> 4);
                    const bool one = DBG && (a.dbg_mode & 8);
                    // The last stage holds the tail of the centres + the polynomial rows: its K=16 steps that hold
                    // nothing but zero padding are skipped (the producers do not write them either).
                    if (kb != nk - 1 || tail_ksteps > 0) {
                        umma_f16(d, a_hi, b_hi, idesc, kb != 0);
                        if (!one) {
                            umma_f16(d, a_hi, b_lo, idesc, 1);
                            umma_f16(d, a_lo, b_hi, idesc, 1);
                        }
                    }
                    if (kb != nk - 1 || tail_ksteps > 1) {
                        umma_f16(d, a_hi + 2, b_hi + 2, idesc, 1);
                        if (!one) {
                            umma_f16(d, a_hi + 2, b_lo + 2, idesc, 1);
                            umma_f16(d, a_lo + 2, b_hi + 2, idesc, 1);
                        }
                    }
                    // frees the stage when these MMAs have read it (in both CTAs of a pair: the slot is refilled by multicast)
                    if (a.pair) umma_commit_mc(bar_empty + 8 * s, 3); else umma_commit(bar_empty + 8 * s);
                    if (kb == nk - 1) umma_commit(bar_tmem_full + 8 * ab); // accumulator complete
                }
                __syncwarp();
            }
        }
        if (DBG && a.dbg && blockIdx.x == 0 && lane == 0) {
            a.dbg[200] = clock64() - t_begin;
            a.dbg[201] = w_t;
            a.dbg[202] = w_a;
            a.dbg[203] = w_b;
        }
    } else if (warp < PRODUCER_WARPS) {
        // ================= Phi producers =================
        const int pt = threadIdx.x;             // 0..511
        const int row = pt & (TM - 1);
        const int khalf = (pt >> 7) & 1;        // which half of the stage's 32 k this thread generates
        const int grp = pt >> 8;                // producer group: stages of its parity
        const float4 nrm4 = *reinterpret_cast<const float4*>(a.norm);
        long long w_c = 0, w_e = 0;
        const long long t_begin = DBG ? clock64() : 0;
        uint32_t it = 0, unit_iter = 0;
        for (int64_t u = unit0; u < n_units; u += ustride, ++unit_iter) {
            const int64_t vt = FD_UNIT_VT(u);
            const int64_t v = vt * TM + row;
            float px = 0.f, py = 0.f, pz = 0.f;
            if (v < a.V) {
                px = a.P[3 * v];
                py = a.P[3 * v + 1];
                pz = a.P[3 * v + 2];
            }
            const uint64_t px2 = pack2(px, px), py2 = pack2(py, py), pz2 = pack2(pz, pz);
            const uint64_t shift2 = pack2((float)(1 << GAUSS_SHIFT), (float)(1 << GAUSS_SHIFT));
            const bool dbg = DBG && a.dbg && blockIdx.x == 0 && warp == 0 && lane == 0 && unit_iter < 15;
            if (dbg) a.dbg[unit_iter * 8 + 0] = clock64();
            // this group's stages of the unit: global stage counter it0 + kb with the group's parity
            const uint32_t it0 = it;
            it += nk;
            for (int kb = (int)((grp ^ it0) & 1); kb < nk; kb += 2) {
                const uint32_t itk = it0 + kb;
                const int s = itk % STAGES;
                const uint32_t ph = (itk / STAGES) & 1;
                const int cs = itk % CDEPTH;
                mbar_wait_t<DBG>(bar_cfull + 8 * cs, (itk / CDEPTH) & 1, w_c); // the stage's centre tile (its own deep ring)
                mbar_wait_t<DBG>(bar_empty + 8 * s, ph ^ 1, w_e);              // the A slot: the MMAs of stage itk - STAGES retired
                uint8_t* a_hi = smem + s * STAGE_BYTES + row * (BK * 2);
                uint8_t* a_lo = a_hi + A_SPLIT_BYTES;
                const ulonglong2* s_ctr2 = reinterpret_cast<const ulonglong2*>(smem + SMEM_CENTRES + cs * C_TILE_BYTES);
                const int k0 = kb * BK;
                const int swz = (row >> 1) & 3;
                uint64_t f2[8]; // the thread's 16 basis values as 8 packed pairs
                if (DBG && (a.dbg_mode & 4)) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) f2[j] = pack2(px, py);
                } else if (kb == nk - 1 && khalf >= tail_ksteps) {
                    // nothing but zero padding in this half of the last stage: the MMA issuer skips its K step
                } else if (k0 + BK <= a.N) {
                    // centre pairs (x0 x1 y0 y1 | z0 z1 w0 w1): the distance of two basis functions per FADD2 / FFMA2
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const ulonglong2 cxy = s_ctr2[(khalf * 8 + j) * 2];     // warp-wide broadcast LDS.128
                        const ulonglong2 czw = s_ctr2[(khalf * 8 + j) * 2 + 1];
                        const uint64_t dx = sub2(px2, cxy.x), dy = sub2(py2, cxy.y), dz = sub2(pz2, czw.x);
                        const uint64_t r2 = fma2(dz, dz, fma2(dx, dx, mul2(dy, dy)));
                        if (KERNEL == FD_KERNEL_GAUSSIAN) {
                            float t0, t1;
                            unpack2(mul2(r2, czw.y), t0, t1);
                            f2[j] = mul2(pack2(ex2_approx(t0), ex2_approx(t1)), shift2);
                        } else if (KERNEL == FD_KERNEL_MULTIQUADRIC) {
                            float t0, t1;
                            unpack2(add2(r2, czw.y), t0, t1);
                            f2[j] = pack2(sqrt_approx(t0), sqrt_approx(t1));
                        } else {
                            float t0, t1;
                            unpack2(r2, t0, t1);
                            f2[j] = pack2(phi<KERNEL>(t0, 0.f), phi<KERNEL>(t1, 0.f));
                        }
                    }
                } else { // the last stage(s): remaining centres, then the affine rows [1, x', y', z'], then zero padding
                    const float* s_ctrf = reinterpret_cast<const float*>(s_ctr2);
                    float f[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const int k = k0 + khalf * 16 + j;
                        const float* cp = s_ctrf + ((khalf * 16 + j) >> 1) * 8 + (j & 1);
                        const float dx = px - cp[0], dy = py - cp[2], dz = pz - cp[4];
                        const float r2 = dx * dx + dy * dy + dz * dz;
                        float val = KERNEL == FD_KERNEL_GAUSSIAN ? ex2_approx(r2 * cp[6]) * (float)(1 << GAUSS_SHIFT)
                                                                 : phi<KERNEL>(r2, cp[6]);
                        if (k >= a.N) {
                            const int r = k - a.N;
                            const float S = KERNEL == FD_KERNEL_GAUSSIAN ? (float)(1 << GAUSS_SHIFT) : 1.0f;
                            val = r == 0 ? S
                                : r == 1 ? (px - nrm4.x) * (nrm4.w * S)
                                : r == 2 ? (py - nrm4.y) * (nrm4.w * S)
                                : r == 3 ? (pz - nrm4.z) * (nrm4.w * S) : 0.0f;
                        }
                        f[j] = val;
                    }
#pragma unroll
                    for (int j = 0; j < 8; ++j) f2[j] = pack2(f[2 * j], f[2 * j + 1]);
                }
                if (!(kb == nk - 1 && khalf >= tail_ksteps)) {
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        uint4 hi, lo;
                        split8(f2 + 4 * h, hi, lo);
                        const int off = ((khalf * 2 + h) ^ swz) * 16;
                        *reinterpret_cast<uint4*>(a_hi + off) = hi;
                        *reinterpret_cast<uint4*>(a_lo + off) = lo;
                    }
                }
                fence_proxy_async(); // generic-proxy stores -> visible to the tensor core (async proxy)
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive(bar_full_a + 8 * s);
                    mbar_arrive(bar_cempty + 8 * cs); // this warp has read the centre tile
                }
            }
            if (dbg) a.dbg[unit_iter * 8 + 1] = clock64();
        }
        if (DBG && a.dbg && blockIdx.x == 0 && lane == 0 && (warp == 0 || warp == 8)) {
            const int o = warp == 0 ? 204 : 214;
            a.dbg[o] = clock64() - t_begin;
            a.dbg[o + 1] = w_c;
            a.dbg[o + 2] = w_e;
        }
    } else {
        // ================= epilogue warps: TMEM -> registers -> (transpose in shared memory) -> global =================
        // They run one unit behind the producers: while they drain and store unit u, Phi generation and the MMAs of
        // unit u+1 proceed (the MMA issuer re-acquires each accumulator through bar_tmem_empty[mt]).
        const int ew = warp - WARP_EPI0;
        const int q = warp & 3;               // TMEM lane quarter this warp may access
        float* stg = reinterpret_cast<float*>(smem + SMEM_EPI_STAGING) + ew * EPI_WARP_FLOATS;
        const int et = threadIdx.x - 32 * WARP_EPI0;
        uint32_t unit_iter = 0;
        long long w_f = 0;
        const long long t_begin = DBG ? clock64() : 0;
        const bool resident_cs = a.ncb <= COLSCALE_RESIDENT_BLOCKS;
        if (resident_cs) { // all column scales once, instead of a barrier + global round trip per unit
            for (int t = et; t < a.ncb * CB; t += 32 * EPILOGUE_WARPS) s_colscale[t] = a.colscale[t];
            asm volatile("bar.sync 2, 256;" ::: "memory");
        }
        for (int64_t u = unit0; u < n_units; u += ustride, ++unit_iter) {
            const int cb = (int)(u % a.ncb);
            const int64_t vt = FD_UNIT_VT(u);
            const int f_base = cb * (CB / 3);
            const int nframes = min(CB / 3, a.F - f_base);
            const float* s_cs = s_colscale + (resident_cs ? cb * CB : 0); // this unit's 240 column scales
            if (!resident_cs) {
                asm volatile("bar.sync 2, 256;" ::: "memory"); // previous unit's readers of s_colscale are done
                for (int t = et; t < CB; t += 32 * EPILOGUE_WARPS) s_colscale[t] = a.colscale[cb * CB + t];
                asm volatile("bar.sync 2, 256;" ::: "memory");
            }
            const bool dbg = DBG && a.dbg && blockIdx.x == 0 && ew == 0 && lane == 0 && unit_iter < 15;
            if (dbg) a.dbg[unit_iter * 8 + 2] = clock64();
            const int ab = unit_iter & 1;
            {
                const int mt = ab;       // TMEM column block of this unit's accumulator
                const int chunk0 = ew >> 2; // this warp drains the even (0) or the odd (1) column chunks
                const int64_t v_warp0 = vt * TM + q * 32;
                const int64_t v = v_warp0 + lane;
                const bool valid = v < a.V;
                float px = 0.f, py = 0.f, pz = 0.f, fo = 0.f;
                bool skip = true;
                if (valid) {
                    px = a.P[3 * v];
                    py = a.P[3 * v + 1];
                    pz = a.P[3 * v + 2];
                    const float d2 = a.dist2 ? a.dist2[v] : 0.f;
                    skip = d2 > a.radius2;                                        // SOP_FaceDeform.cpp:408-410
                    fo = powf(1.0f - fminf(d2 / a.radius2, 1.0f), a.falloffrate); // :423-424
                    if (skip) fo = 0.f;
                    if (a.falloff_out && cb == 0 && chunk0 == 0) a.falloff_out[v] = fo;
                }
                float tu[3] = {0, 0, 0}, tv[3] = {0, 0, 0}, tn[3] = {0, 0, 0};
                if (TANGENT && valid) {
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
                const bool vec = a.vec_store_ok != 0;
                const bool fo_one = __all_sync(0xffffffffu, fo == 1.0f || !valid);
                // everything above is independent of the accumulator: it overlaps the MMAs of this unit
                mbar_wait_t<DBG>(bar_tmem_full + 8 * ab, (unit_iter >> 1) & 1, w_f);
                tc_fence_after();
                if (dbg) a.dbg[unit_iter * 8 + 3] = clock64();
                if (vec) {
                    // fast path: out = P + acc * (colscale * falloff) (colscale is a power of two, so the product order
                    // does not change the rounding; a skipped vertex has falloff 0 and keeps P exactly), transposed
                    // through shared memory into [frame][vertex][xyz] rows that one lane hands to the bulk-copy engine
#pragma unroll 1
                    for (int ch = chunk0; ch * EPI_FRAMES < nframes; ch += 2) {
                        if (DBG && (a.dbg_mode & 2)) continue;
                        float acc[EPI_COLS];
                        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + mt * ACC1_COL + ch * EPI_COLS;
                        const bool dbg2 = DBG && dbg && unit_iter == 2 && ch == 2;
                        if (dbg2) a.dbg[120] = clock64();
                        tmem_ld16(taddr, acc);
                        tmem_ld8(taddr + 16, acc + 16);
                        float* buf = stg; // single buffer: the TMA store of the previous chunk must have read it
                        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                        if (dbg2) a.dbg[121] = clock64();
                        __syncwarp();
                        tmem_ld_wait();
                        if (dbg2) a.dbg[122] = clock64();
                        const float4* cs4 = reinterpret_cast<const float4*>(s_cs + ch * EPI_COLS);
                        float* dst = buf + lane * 3;
                        if (fo_one) { // no falloff anywhere in this warp's 32 vertices (the common cook): one FMA per value
#pragma unroll
                            for (int g = 0; g < EPI_COLS / 4; ++g) { // 4 columns at a time
                                const float4 cs = cs4[g];
                                const float c4[4] = {cs.x, cs.y, cs.z, cs.w};
#pragma unroll
                                for (int e = 0; e < 4; ++e) {
                                    const int col = 4 * g + e, i = col / 3, k = col - 3 * i;
                                    const float p = k == 0 ? px : (k == 1 ? py : pz);
                                    dst[i * 96 + k] = fmaf(acc[col], c4[e], p);
                                }
                            }
                        } else {
#pragma unroll
                            for (int g = 0; g < EPI_COLS / 4; ++g) {
                                const float4 cs = cs4[g];
                                const float c4[4] = {cs.x, cs.y, cs.z, cs.w};
#pragma unroll
                                for (int e = 0; e < 4; ++e) {
                                    const int col = 4 * g + e, i = col / 3, k = col - 3 * i;
                                    const float p = k == 0 ? px : (k == 1 ? py : pz);
                                    dst[i * 96 + k] = fmaf(acc[col], c4[e] * fo, p);
                                }
                            }
                        }
                        if (dbg2) a.dbg[123] = clock64();
                        fence_proxy_async(); // staging writes -> visible to the TMA (async proxy)
                        __syncwarp();
                        if (dbg2) a.dbg[124] = clock64();
                        if (lane == 0 && !(DBG && (a.dbg_mode & 16))) {
                            // one TMA store of the [8 frames][32 vertices x 3] tile; frames >= F and vertices >= V are clipped
                            asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                                         ::"l"(reinterpret_cast<uint64_t>(&map_out)), "r"(smem_u32(buf)),
                                           "r"((int)(v_warp0 * 3)), "r"(f_base + ch * EPI_FRAMES)
                                         : "memory");
                            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                        }
                        if (dbg2) a.dbg[125] = clock64();
                    }
                } else {
                    // general path (partial tile, V not a multiple of 4, tangent projection): per-lane scalar stores
#pragma unroll 1
                    for (int ch = chunk0; ch * EPI_FRAMES < nframes; ch += 2) {
                        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + mt * ACC1_COL + ch * EPI_COLS;
                        const int fcnt = min(EPI_FRAMES, nframes - ch * EPI_FRAMES);
#pragma unroll 1
                        for (int part = 0; part < EPI_COLS / 8; ++part) {
                            float acc[8];
                            tmem_ld8(taddr + 8 * part, acc);
                            tmem_ld_wait();
                            // frames straddle the 8-column parts: gather through the staging buffer as [col][lane]
#pragma unroll
                            for (int c = 0; c < 8; ++c) stg[(part * 8 + c) * 32 + lane] = acc[c] * s_cs[ch * EPI_COLS + part * 8 + c];
                        }
                        __syncwarp();
#pragma unroll 1
                        for (int i = 0; i < fcnt; ++i) {
                            float d[3] = {stg[(3 * i) * 32 + lane], stg[(3 * i + 1) * 32 + lane], stg[(3 * i + 2) * 32 + lane]};
                            if (TANGENT) project_to_tangents(tu, tv, tn, d);
                            if (valid) {
                                float* dst = a.P_out + ((size_t)(f_base + ch * EPI_FRAMES + i) * (size_t)a.V + (size_t)v) * 3;
                                dst[0] = skip ? px : px + d[0] * fo;
                                dst[1] = skip ? py : py + d[1] * fo;
                                dst[2] = skip ? pz : pz + d[2] * fo;
                            }
                        }
                        __syncwarp();
                    }
                }
                if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                // this warp's share of the accumulator is drained: hand it back to the MMA issuer
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_tmem_empty + 8 * mt);
            }
            if (dbg) a.dbg[unit_iter * 8 + 4] = clock64();
        }
        if (DBG && a.dbg && blockIdx.x == 0 && lane == 0 && ew == 0) {
            a.dbg[208] = clock64() - t_begin;
            a.dbg[209] = w_f;
        }
    }

#undef FD_UNIT_VT
    tc_fence_before();
    __syncthreads();
    if (a.pair) cluster_sync_all(); // no CTA leaves while its peer may still multicast into it or signal its barriers
    if (warp == WARP_TMA) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS));
    }
}

// ------------------------------------------------------------------------------------------------------------
// pack: FP64 weights -> transposed, column-scaled FP16 hi/lo tables W^T[c][k] the TMA streams
// ------------------------------------------------------------------------------------------------------------

// bounding box of the control points -> (centre, 2 / largest extent, diagonal)
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
        float ext = 0.f, d2 = 0.f;
        for (int k = 0; k < 3; ++k) {
            norm[k] = 0.5f * (s_lo[k][0] + s_hi[k][0]);
            ext = fmaxf(ext, s_hi[k][0] - s_lo[k][0]);
            d2 += (s_hi[k][0] - s_lo[k][0]) * (s_hi[k][0] - s_lo[k][0]);
        }
        norm[3] = ext > 0.f ? 2.0f / ext : 1.0f;
        norm[4] = sqrtf(d2); // the rig's bounding-box diagonal: the length FD_EVAL_AUTO's tolerance is relative to
    }
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

} // namespace tc

static_assert(sizeof(CUtensorMap) == FD_TMAP_BYTES, "fd_model reserves FD_TMAP_BYTES per tensor map");

int fd_tc_kpad(int N) { return fd_round_up(N + 4, tc::BK); }
int fd_tc_ncb(int F) { return (3 * F + tc::CB - 1) / tc::CB; }
int fd_tc_col_pad(int F) { return fd_tc_ncb(F) * tc::CB; }

// bounding box of the control points -> normalisation of the affine rows (depends on the centres only)
cudaError_t fd_launch_tc_norm(fd_ctx* ctx, fd_model* m)
{
    tc::k_tc_norm<<<1, 256, 0, ctx->stream>>>(m->d_rest, m->N, m->d_tc_norm);
    ctx->launches += 1;
    return cudaGetLastError();
}

void fd_tc_pack_args_fill(const fd_model* m, fd_tc_pack_args* pk)
{
    pk->enabled = 1;
    pk->N = m->N;
    pk->np = m->np;
    pk->Kpad = fd_tc_kpad(m->N);
    pk->ncol = 3 * m->F;
    pk->ncol_pad = fd_tc_col_pad(m->F);
    pk->phi_shift = m->prm.kernel == FD_KERNEL_GAUSSIAN ? tc::GAUSS_SHIFT : 0;
    pk->norm = m->d_tc_norm;
    pk->scale = m->d_tc_scale;
    pk->unscale = m->d_tc_unscale;
    pk->wt_hi = m->d_tc_wt_hi;
    pk->wt_lo = m->d_tc_wt_lo;
    pk->flags = m->d_flags;
}

// builds the tensor-path tables for the weights currently in m->d_W (called from fd_launch_pack); when the slab solve
// has already written them in its epilogue only the tensor maps are (re)encoded
cudaError_t fd_launch_pack_tc(fd_ctx* ctx, fd_model* m)
{
    cudaStream_t s = ctx->stream;
    const int ncol = 3 * m->F, ncol_pad = fd_tc_col_pad(m->F), Kpad = fd_tc_kpad(m->N);
    if (m->tc_packed_by_solve) {
        m->tc_packed_by_solve = false;
        if (!tc::make_map((CUtensorMap*)m->tc_map_hi, m->d_tc_wt_hi, Kpad, ncol_pad, tc::BK, tc::CB / 2) ||
            !tc::make_map((CUtensorMap*)m->tc_map_lo, m->d_tc_wt_lo, Kpad, ncol_pad, tc::BK, tc::CB / 2))
            return cudaErrorInvalidValue;
        return cudaSuccess;
    }
    tc::k_tc_colscale<<<(ncol_pad + 31) / 32, 256, 0, s>>>(fd_w_src(m), m->ldw, m->N, m->np, ncol, ncol_pad,
                                                            m->prm.kernel == FD_KERNEL_GAUSSIAN ? tc::GAUSS_SHIFT : 0, 14,
                                                            m->d_tc_norm, m->d_tc_unscale, m->d_tc_scale, m->d_flags);
    dim3 grid((ncol_pad + 31) / 32, (Kpad + 31) / 32);
    tc::k_tc_pack<<<grid, 256, 0, s>>>(fd_w_src(m), m->ldw, m->N, m->np, ncol, ncol_pad, Kpad, m->d_tc_norm, m->d_tc_scale,
                                       (__half*)m->d_tc_wt_hi, (__half*)m->d_tc_wt_lo);
    ctx->launches += 2;
    if (!tc::make_map((CUtensorMap*)m->tc_map_hi, m->d_tc_wt_hi, Kpad, ncol_pad, tc::BK, tc::CB / 2) ||
        !tc::make_map((CUtensorMap*)m->tc_map_lo, m->d_tc_wt_lo, Kpad, ncol_pad, tc::BK, tc::CB / 2))
        return cudaErrorInvalidValue;
    return cudaGetLastError();
}

// per-device function attributes of every instantiation (fd_ctx_create)
cudaError_t fd_eval_tc_setup(fd_ctx* ctx)
{
    (void)ctx;
    cudaError_t e = cudaSuccess;
#define FD_TC_ATTR(...)                                                                                                   \
    if (e == cudaSuccess) e = cudaFuncSetAttribute(tc::k_eval_tc<__VA_ARGS__>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SMEM_ALLOC)
    FD_TC_ATTR(FD_KERNEL_GAUSSIAN, false, false);
    FD_TC_ATTR(FD_KERNEL_GAUSSIAN, true, false);
    FD_TC_ATTR(FD_KERNEL_MULTIQUADRIC, false, false);
    FD_TC_ATTR(FD_KERNEL_MULTIQUADRIC, true, false);
    FD_TC_ATTR(FD_KERNEL_THINPLATE, false, false);
    FD_TC_ATTR(FD_KERNEL_THINPLATE, true, false);
    FD_TC_ATTR(FD_KERNEL_GAUSSIAN, false, true); // the debug instantiation (FD_TC_DEBUG)
#undef FD_TC_ATTR
    return e;
}

// A view of `m` that evaluates the frames [f_begin, f_begin + f_count) only: the weight tables' tensor maps re-encoded at
// the block's first column (f_begin must be a multiple of 80 frames = one 240-column block) and the column scales offset.
// Used by the host-pointer evaluation, which reads the result back block by block while the next block is evaluated.
bool fd_tc_view_frames(const fd_model* m, fd_model* view, int f_begin)
{
    const int col0 = 3 * f_begin, ncol_pad = fd_tc_col_pad(m->F), Kpad = fd_tc_kpad(m->N);
    if (col0 % tc::CB != 0 || col0 >= ncol_pad) return false;
    view->d_tc_unscale = m->d_tc_unscale + col0;
    view->d_tc_scale = m->d_tc_scale + col0;
    return tc::make_map((CUtensorMap*)view->tc_map_hi, (unsigned short*)m->d_tc_wt_hi + (size_t)col0 * Kpad, Kpad, ncol_pad - col0, tc::BK, tc::CB / 2) &&
           tc::make_map((CUtensorMap*)view->tc_map_lo, (unsigned short*)m->d_tc_wt_lo + (size_t)col0 * Kpad, Kpad, ncol_pad - col0, tc::BK, tc::CB / 2);
}

cudaError_t fd_launch_eval_tc(fd_ctx* ctx, const fd_model* m, const float* P, int64_t V, const float* dist2,
                              const float* tu, const float* tv, const float* nrm, float* P_out, float* falloff_out,
                              const int* sel, int sel_id)
{
    if (V <= 0) return cudaSuccess;
    tc::Args a;
    a.sel = sel;
    a.sel_id = sel_id;
    a.ctab = m->d_ctab_pair;
    a.norm = m->d_tc_norm;
    a.colscale = m->d_tc_unscale;
    a.N = m->N;
    a.Kpad = fd_tc_kpad(m->N);
    a.Ktot = m->N + m->np;
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
    a.vec_store_ok = (V % 4 == 0) && ((reinterpret_cast<uintptr_t>(P_out) & 15) == 0) && !a.do_tangent;
    // the tensor map of the output depends on (P_out, V, F) only: cooks that write into the same buffer reuse it
    if (a.vec_store_ok && !(ctx->tc_out_ptr == P_out && ctx->tc_out_V == V && ctx->tc_out_F == m->F)) {
        ctx->tc_out_ptr = nullptr;
        if (tc::make_out_map((CUtensorMap*)ctx->tc_out_map, P_out, V, m->F, tc::EPI_FRAMES)) {
            ctx->tc_out_ptr = P_out;
            ctx->tc_out_V = V;
            ctx->tc_out_F = m->F;
        } else {
            a.vec_store_ok = 0;
        }
    }
    const CUtensorMap& mo = *(const CUtensorMap*)ctx->tc_out_map;
    const bool want_dbg = ctx->dbg.has_tc_debug && ctx->d_tc_dbg;
    long long* d_dbg = ctx->d_tc_dbg;
    a.dbg = want_dbg ? d_dbg : nullptr;
    a.dbg_mode = want_dbg ? ctx->dbg.tc_debug : 0;
    const int64_t n_vt = (V + tc::TM - 1) / tc::TM;
    const bool no_pair = ctx->dbg.tc_nopair;
    a.pair = (!no_pair && n_vt >= 4 && ctx->sm_count >= 2) ? 1 : 0;
    const int64_t n_units = (a.pair ? (n_vt + 1) / 2 : n_vt) * a.ncb;
    const int64_t max_ctas = a.pair ? ctx->sm_count / 2 : ctx->sm_count;
    const int grid = (int)(n_units < max_ctas ? n_units : max_ctas) * (a.pair ? 2 : 1);
    const CUtensorMap& mh = *(const CUtensorMap*)m->tc_map_hi;
    const CUtensorMap& ml = *(const CUtensorMap*)m->tc_map_lo;
    const bool tang = a.do_tangent != 0;
    cudaError_t launch_err = cudaSuccess;
#define FD_TC_LAUNCH(KERNEL, TANG)                                                                                  \
    do {                                                                                                            \
        auto kfn = tc::k_eval_tc<KERNEL, TANG, false>;                                                              \
        if (want_dbg && KERNEL == FD_KERNEL_GAUSSIAN && !TANG) kfn = tc::k_eval_tc<FD_KERNEL_GAUSSIAN, false, true>; \
        cudaLaunchConfig_t cfg = {};                                                                                \
        cfg.gridDim = dim3(grid);                                                                                   \
        cfg.blockDim = dim3(tc::THREADS);                                                                           \
        cfg.dynamicSmemBytes = tc::SMEM_ALLOC;                                                                      \
        cfg.stream = ctx->stream;                                                                                   \
        cudaLaunchAttribute attr[1];                                                                                \
        attr[0].id = cudaLaunchAttributeClusterDimension;                                                           \
        attr[0].val.clusterDim.x = a.pair ? 2 : 1;                                                                  \
        attr[0].val.clusterDim.y = 1;                                                                               \
        attr[0].val.clusterDim.z = 1;                                                                               \
        cfg.attrs = attr;                                                                                           \
        cfg.numAttrs = 1;                                                                                           \
        launch_err = cudaLaunchKernelEx(&cfg, kfn, a, mh, ml, mo);                                                  \
    } while (0)
    switch (m->prm.kernel) {
    case FD_KERNEL_GAUSSIAN:
        if (tang) FD_TC_LAUNCH(FD_KERNEL_GAUSSIAN, true); else FD_TC_LAUNCH(FD_KERNEL_GAUSSIAN, false);
        break;
    case FD_KERNEL_MULTIQUADRIC:
        if (tang) FD_TC_LAUNCH(FD_KERNEL_MULTIQUADRIC, true); else FD_TC_LAUNCH(FD_KERNEL_MULTIQUADRIC, false);
        break;
    default:
        if (tang) FD_TC_LAUNCH(FD_KERNEL_THINPLATE, true); else FD_TC_LAUNCH(FD_KERNEL_THINPLATE, false);
        break;
    }
#undef FD_TC_LAUNCH
    if (want_dbg) { // development aid: per-unit phase timestamps of CTA 0 (cycles)
        long long h[256];
        cudaStreamSynchronize(ctx->stream);
        cudaMemcpy(h, d_dbg, sizeof(h), cudaMemcpyDeviceToHost);
        fprintf(stderr, "[fd_tc] CTA 0 waits (cycles): MMA lane total %lld = tmem_empty %lld + full_a %lld + full_b %lld + issue; "
                        "TMA lane total %lld: cempty %lld, empty %lld; producer w0 total %lld: cfull %lld, empty %lld; "
                        "producer w8 total %lld: cfull %lld, empty %lld; epilogue w0 total %lld: tmem_full %lld\n",
                h[200], h[201], h[202], h[203], h[210], h[211], h[212], h[204], h[205], h[206], h[214], h[215], h[216], h[208], h[209]);
        fprintf(stderr, "[fd_tc] chunk: issue-ldtm+bulkwait %lld  ldtm-wait %lld  math+sts %lld  fence %lld  bulk-issue %lld\n",
                h[121] - h[120], h[122] - h[121], h[123] - h[122], h[124] - h[123], h[125] - h[124]);
        {
            int last = 0;
            while (last + 1 < 15 && h[(last + 1) * 8]) ++last;
            fprintf(stderr, "[fd_tc] CTA 0: %d units, producer period %.0f cycles/unit, epilogue period %.0f cycles/unit\n", last + 1,
                    last ? (double)(h[last * 8] - h[0]) / last : 0.0, last ? (double)(h[last * 8 + 4] - h[4]) / last : 0.0);
        }
        for (int u = 0; u < 15 && h[u * 8]; ++u)
            fprintf(stderr, "[fd_tc] unit %d: produce %lld | epilogue: wait-tmem %lld  drain+store %lld | producer start->epilogue end %lld\n", u,
                    h[u * 8 + 1] - h[u * 8], h[u * 8 + 3] - h[u * 8 + 2], h[u * 8 + 4] - h[u * 8 + 3], h[u * 8 + 4] - h[u * 8]);
    }
    ctx->launches += 1;
    return launch_err != cudaSuccess ? launch_err : cudaGetLastError();
}
