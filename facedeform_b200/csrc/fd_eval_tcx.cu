// fd_eval_tcx.cu -- K3-TCX: the Gaussian evaluation on the FP16 tensor cores (tcgen05) WITHOUT the FP32 cancellation error.
//
//   D[v][3f+k] = sum_j Phi[v][j] * W[j][3f+k]   (reference SOP_FaceDeform.cpp:404-439 for all F frames at once, the same
//   GEMM-with-generated-A as fd_eval_tc.cu, the same SOP epilogue)
//
// Why another kernel.  The sum cancels: its terms are ~10^4 times the result, so any FP32 accumulation
// -- FMA chain or tensor-core accumulator alike -- errs by ~2^-24 x S with S = sum_j |w_j| phi_j, and FD_EVAL_AUTO had to fall
// back to FP64 (20x slower) from ~1000 control points on.  Here both operands are cut into an INTEGER leading digit of h bits
// and an FP16 mid + lo remainder:
//     Phi_k 2^(h - s_k) = a_hi + a_mid + a_lo,   W_kc 2^(e_c + h + s_k) = b_hi + b_mid + b_lo,   a_hi, b_hi integers, |mid| <= 1/2,
//     |lo| <= 2^-12.  e_c normalises column c to max |w| < 1; row k's largest weight is then 2^-r_k, and s_k = r_k / 2 moves half
//     of that deficit from Phi's digit to W's, so both digits of a row's dominant terms have h - r_k / 2 significant bits.
//   * a_hi b_hi, the only product of full size, goes to its own accumulator: a sum of integers.  While every partial sum stays
//     below 2^24 the FP32 accumulator holds it EXACTLY, whatever the order and the tensor core's internal rounding.  The digit
//     width h is chosen on the device at pack time as the widest (<= 11: FP16 holds integers up to 2^11) with
//     4 x 2^2h x max_i sum_k phi_k(c_i) rowmax_k <= 2^24 -- the bound evaluated at the control points, with a factor 4 in hand for
//     the vertices in between -- and the producers VERIFY it for every vertex they handle (sum_k a_hi 2^(h - s_k) per row; a
//     violation raises FD_FLAG_EVAL_INEXACT, reported as fd_report.eval_inexact: the result is then only FP32-accurate).
//   * the seven products a_hi b_mid, a_mid b_hi, a_hi b_lo, a_lo b_hi, a_mid b_mid, a_mid b_lo, a_lo b_mid go to a second
//     accumulator; they are 2^-h of the full size.  The tensor core truncates (rounds toward zero) when it adds into an FP32
//     accumulator, a bias that grows with the ~7 K / 16 additions: measured (and reproduced by the CPU emulation) 0.08 ... 0.17
//     x 2^-24 S at h = 8 ... 7, which is why h is taken as wide as the bound allows and not from a worst case.  What is dropped
//     (lo x lo: 2^-26 per product) and the FP16 rounding of the lo parts (2^-25 per value) are far below that.  (Without the
//     two mid x lo products the error was 7e-5 x diagonal at N = 4096, h = 5: they are 2^-14-2h of the full size.)
//   * Phi is produced in FP64 -- an FP32 Phi alone would put 2^-24 S back -- with the kernel function good to 2^-34: 15 FP64
//     instructions per basis value (expanded distance 4, exp2 by a 16-entry table + quartic 8, digit split 3) and ONE conversion (F64 <-> F32
//     conversions run at a quarter of the FP64 rate; the leading digit becomes a float by integer arithmetic).  Four values are
//     carried side by side so the dependent chains overlap.  Amortised over 120 columns.
// Net error ~2^-29 ... 2^-31 x S: measured 1e-7 x the rig's diagonal where the FP32 kernels give 5e-6 ... 2.5e-5
// (tests/tools/accuracy_probe.py; the CPU emulation of the scheme is in tests/tools/fp32_error_emulation.py).
//
// Structure (one persistent CTA per SM, 22 warps, the roles of fd_eval_tc.cu).  A unit is 128 vertices x TWO 120-column blocks
// (80 frames): the Phi tiles -- the expensive operand, FP64 arithmetic per value -- are generated once and multiplied with the
// weight tiles of both blocks (one block per unit when the batch has only one: F <= 40).
//   warps 0..15  Phi producers (two groups alternating stages), three SWIZZLE_64B A tiles per 32-centre stage; ring of 3 slots
//   warps 16..19 epilogue: acc0 + acc1 from TMEM, un-scale, falloff, P +=, transpose, TMA store (two staging buffers per warp)
//   warp 20      TMA: per stage the three weight tiles W^T[120 cols][32 k] of both blocks, back to back (240 rows per split: one
//                weight slot = 45 KB, ring of 2) + the centre tiles (FP64, own ring of 8)
//   warp 21      MMA issuer: per K=16 step 1 MMA into acc0 and 7 (5 when h >= 9) into acc1, M = 128, N = 240 -- the Phi operand
//                is read from shared memory once per instruction for both blocks (11.5 KB per 120 cycles instead of 8 KB per 64:
//                the tensor core's operand reads are the largest user of the shared-memory pipe)
// TMEM: acc0 in columns 0..239, acc1 in 256..495 (block j at +120 j); the epilogue releases it after its last tcgen05.ld, and the
// producers fill the Phi ring for the next unit meanwhile.  FD_TCX_NARROW=1 keeps one N = 128 MMA per block (3 + 5 slots),
// FD_TCX_CBU=1 the one-block units with two units in flight in TMEM (4 + 4 slots): 0.38 / 0.51 ms where the default takes 0.35
// (C2: 100k vertices x 256 control points x 240 frames; profiles/r2k_bench_fast_*.json).
#include <cuda.h>
#include <cuda_fp16.h>
#include <stdlib.h>
#include <string.h>

#include "fd_internal.h"
#include "fd_eval_common.cuh"
#include "fd_tc_common.cuh"

namespace tcx {

using namespace tcc;

constexpr int TM = 128;                       // vertices per unit
constexpr int CB = 120;                       // columns per unit (40 frames)
constexpr int NMMA = 128;                     // N of the MMAs (multiple of 16)
constexpr int BK = 32;                        // k per pipeline stage (64-byte rows -> SWIZZLE_64B)
constexpr int A_SPLIT_BYTES = TM * BK * 2;    // 8192
constexpr int B_SPLIT_BYTES = CB * BK * 2;    // 7680 = 15 swizzle atoms of 512 bytes
constexpr int A_STAGE_BYTES = 3 * A_SPLIT_BYTES; // 24576: the digit / mid / lo tiles of Phi for 32 centres
constexpr int C_TILE_BYTES = BK * 32;         // a stage's 32 centres as double4: t = q . (a, b, c) + d + |q|^2 sc (q = p - centre 0)
constexpr int S_TILE_BYTES = BK * 8;          // their sc = -log2(e) / R^2
constexpr int R_TILE_BYTES = BK * 4;          // and 2^(h - s_k) as floats (the scale of row k's digit)
constexpr int CDEPTH = 8;                     // depth of the centre-tile ring
constexpr int MAX_SA = 4, MAX_SB = 5;         // barrier slots reserved for the two rings
constexpr int PRODUCER_WARPS = 16;
constexpr int EPILOGUE_WARPS = 4;             // one per TMEM lane quarter (leaves the producers 88 registers per thread)
constexpr int THREADS = 32 * (2 + PRODUCER_WARPS + EPILOGUE_WARPS);
constexpr int WARP_EPI0 = PRODUCER_WARPS;
constexpr int WARP_TMA = PRODUCER_WARPS + EPILOGUE_WARPS;
constexpr int WARP_MMA = WARP_TMA + 1;
constexpr int TMEM_COLS = 512;
constexpr int UNIT_COLS = 256;                // TMEM columns of one column block: acc0 at +0, acc1 at +128
constexpr int ACC1_OFF = 128;
constexpr int EPI_FRAMES = 8;
constexpr int EPI_WARP_FLOATS = EPI_FRAMES * 96;
constexpr int EPI_COLS = EPI_FRAMES * 3;
constexpr int COLSCALE_RESIDENT_BLOCKS = 6;   // column scales of up to 6 column blocks (F <= 240) stay resident

// Shared-memory layout.  A unit is 128 vertices x CBU column blocks: the Phi tiles of a stage (the expensive operand: FP64
// arithmetic per value) are generated ONCE and multiplied with the weight tiles of CBU column blocks, so the two operands
// live in rings of their own -- SA slots of Phi tiles, SB slots of weight tiles.
//   CBU = 1: 4 + 4 slots, a weight slot is freed and filled together with the Phi slot of the same stage; two units in
//            flight in tensor memory (ping-pong).
//   CBU = 2: 3 + 5 slots; both column blocks' accumulators fill tensor memory (2 x 256 columns), the epilogue drains them
//            while the producers fill the Phi ring for the next unit.
//   CBU = 2, WIDE: the two blocks' weight tiles lie back to back (240 rows per split) and ONE MMA of N = 240 serves both:
//            the Phi operand is read from shared memory once per instruction instead of once per block (11.5 KB per 120
//            cycles instead of 8 KB per 64 -- the tensor core's operand reads are the largest user of the shared-memory
//            pipe), and no instruction computes 8 unused columns.  3 Phi slots + 2 weight slots of 45 KB.
template <int CBU, bool WIDE = false>
struct Lay {
    static_assert(!WIDE || CBU == 2, "the wide MMA spans two column blocks");
    static constexpr int SA = CBU == 1 ? 4 : 3;
    static constexpr int SB = CBU == 1 ? 4 : (WIDE ? 2 : 5);
    static constexpr int B_SPLIT = WIDE ? 2 * B_SPLIT_BYTES : B_SPLIT_BYTES; // one split (digit, mid or lo) of a weight slot
    static constexpr int B_SLOT = 3 * B_SPLIT;
    static constexpr int B_RING = SA * A_STAGE_BYTES;
    static constexpr int EPI_BUFS = CBU == 1 ? 1 : 2; // staging buffers per epilogue warp (two: a bulk store drains one while the next chunk fills the other)
    static constexpr int EPI_STAGING = B_RING + SB * B_SLOT;
    static constexpr int BARRIERS = EPI_STAGING + EPILOGUE_WARPS * EPI_BUFS * EPI_WARP_FLOATS * 4;
    static constexpr int CENTRES = BARRIERS + 512;
    static constexpr int SC = CENTRES + CDEPTH * C_TILE_BYTES;
    static constexpr int ROWEXP = SC + CDEPTH * S_TILE_BYTES;
    static constexpr int COLSCALE = ROWEXP + CDEPTH * R_TILE_BYTES;
    static constexpr int TOTAL = COLSCALE + COLSCALE_RESIDENT_BLOCKS * CB * 4;
    static constexpr int ALLOC = TOTAL + 1024;
    static constexpr int ACC1 = WIDE ? 256 : ACC1_OFF;          // TMEM column of the second accumulator
    static constexpr int BLOCK_COLS = WIDE ? CB : UNIT_COLS;    // TMEM column stride between the unit's column blocks
    static_assert(SA <= MAX_SA && SB <= MAX_SB, "barrier slots");
    static_assert(TOTAL + 1024 + 1024 + 128 <= 227 * 1024, "shared-memory budget (dynamic + alignment slack + static)");
    static_assert(CENTRES % 16 == 0 && A_STAGE_BYTES % 1024 == 0 && B_SLOT % 512 == 0 && B_RING % 1024 == 0, "tile alignment");
};

struct Args {
    const double4* ctab;    // (a, b, c, d + h - s_k) per centre (see C_TILE_BYTES; per solve: k_tcx_pack), Kpad entries
    const double* csc;      // sc per centre, padded
    const float* origin;    // centre 0
    const float* norm;      // (ox, oy, oz, s): affine-row coordinates x' = (x - o) * s
    const float* colscale;  // per column: 2^-e_c
    const float* pw;        // per row k: 2^(h - s_k), Kpad entries
    const int* hbits;       // device word: h (chosen at pack time from the weights)
    int* flags;             // FD_FLAG_EVAL_INEXACT is raised when a vertex's leading-digit sum may leave the exact range
    int N, Kpad, Ktot, F, ncb;
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
    const int* sel;         // FD_EVAL_AUTO: the chosen evaluation kernel (device word) or NULL; the launch returns at once
    int sel_id;             // when *sel != sel_id
    int dbg_mode;           // FD_TC_DEBUG bits (timing experiments, results are then garbage): 2 no epilogue stores, 8 one MMA
                            // of eight; 32: the exact-range check fires at 2^10 instead of 2^22 (tests its plumbing)
};

// x (|x| <= 2^11, FP64) -> the integer digit as a float (by integer arithmetic: no conversion) and the remainder |r| <= 1/2 as
// a float (one F2F)
__device__ __forceinline__ void split_digit(double x, float& hi, float& rem)
{
    const double magic = 6755399441055744.0;      // 1.5 * 2^52: the low word of (x + magic) is rint(x)
    const double xm = x + magic;
    const int k = __double2loint(xm);
    rem = __double2float_rn(x - (xm - magic));
    hi = __int_as_float(0x4B400000 + k) - 12582912.0f; // (1.5 * 2^23 + k) - 1.5 * 2^23, exact for |k| < 2^22
}

// 2^t (t <= ~30; the caller has folded the row's digit exponent h - s_k into t) with 2^-34 relative error, all FP64:
// t = k / 16 + r, 2^(k / 16) from a 16-entry table (one row of shared-memory banks: conflict free whatever entries the lanes of a
// warp pick), 2^r - 1 = r (ln2 + r (c2 + r (c3 + r c4))) for |r| <= 1 / 32 (the r^5 term: ln2^5 / 120 / 32^5 < 4e-11).  Adding
// 1.5 * 2^48 -- ulp 1 / 16 -- rounds t to a multiple of 1 / 16 and leaves k in the low word.  Below t = -200 (or for a negative
// NaN) the result is 0; the test reads the high word only.  8 FP64 instructions (the version with fma(t, 16, magic), an Estrin
// quartic, fmax(t, -200) and the exponent passed separately took 10 and twice the integer work).
__device__ __forceinline__ double exp2_digit(double t, const double* __restrict__ s_tab)
{
    const double M = 422212465065984.0;           // 1.5 * 2^48
    const double kf = t + M;
    const int k = __double2loint(kf);             // rint(16 t)
    const double r = t - (kf - M);                // in [-1/32, 1/32]
    double p = fma(r, 0.009618129107628477, 0.05550410866482158);
    p = fma(r, p, 0.2402265069591007);
    p = fma(r, p, 0.6931471805599453);
    const double T = s_tab[k & 15];
    const double s = fma(T * r, p, T);            // in [0.97, 2.05): the exponent add below cannot carry wrongly
    const int hi = __double2hiint(s) + ((k & ~15) << 16); // += floor(k / 16) << 20
    return __hiloint2double((unsigned)__double2hiint(t) > 0xC0690000u ? 0 : hi, __double2loint(s));
}

struct ProducerRow {      // what a producer thread keeps per unit: its vertex relative to centre 0
    double qx, qy, qz, pp;
    float pxf, pyf, pzf;
};

// The thread's 16 values of a stage (row `row`, k = k0 + khalf * 16 ...): Phi_k 2^(h - s_k) -> digit / mid / lo FP16 -> the three
// SWIZZLE_64B A tiles.  PLAIN: 32 centres, no affine rows, no padding.  Returns the thread's share of sum_k a_hi 2^(h - s_k).
template <bool PLAIN>
__device__ __forceinline__ float fill_stage_half(const Args& a, const ProducerRow& pr, const float4& nrm4, uint8_t* a_hi,
                                                 const double4* __restrict__ s_ctr, const double* __restrict__ s_sc,
                                                 const float* __restrict__ s_pw, const double* __restrict__ s_tab, int khalf,
                                                 int swz, int k0, float bound)
{
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        uint32_t whi[4], wmid[4], wlo[4];
        double x[8]; // eight values side by side: their dependent chains overlap
        float pw[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int kk = khalf * 16 + h * 8 + e;
            const double4 c = s_ctr[kk]; // warp-wide broadcasts
            const double t = fma(pr.qx, c.x, fma(pr.qy, c.y, fma(pr.qz, c.z, fma(pr.pp, s_sc[kk], c.w))));
            x[e] = exp2_digit(t, s_tab);
            pw[e] = s_pw[kk];
        }
        if (!PLAIN) { // the last stage(s): affine rows [1, x', y', z'] after the centres, then zero padding
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const int k = k0 + khalf * 16 + h * 8 + e;
                if (k >= a.N) {
                    const int r = k - a.N;
                    const float cf = r >= a.Ktot - a.N ? 0.f
                                   : r == 0 ? 1.f
                                   : r == 1 ? (pr.pxf - nrm4.x) * nrm4.w
                                   : r == 2 ? (pr.pyf - nrm4.y) * nrm4.w : (pr.pzf - nrm4.z) * nrm4.w;
                    x[e] = (double)cf * (double)pw[e];
                }
            }
        }
        float hf[8], rf[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            split_digit(x[e], hf[e], rf[e]);
            bound = fmaf(fabsf(hf[e]), pw[e], bound);
        }
#pragma unroll
        for (int e2 = 0; e2 < 4; ++e2) {
            const __half2 h2 = __floats2half2_rn(hf[2 * e2], hf[2 * e2 + 1]); // exact: integers <= 2048 -- always for Phi
            if (!PLAIN) { // an affine row of a vertex far outside the rig can exceed that: what the FP16 digit drops goes into
                const float2 hb = __half22float2(h2); // the remainder
                rf[2 * e2] += hf[2 * e2] - hb.x;
                rf[2 * e2 + 1] += hf[2 * e2 + 1] - hb.y;
            }
            const __half2 m2 = __floats2half2_rn(rf[2 * e2], rf[2 * e2 + 1]);
            const float2 mb = __half22float2(m2);
            const __half2 l2 = __floats2half2_rn(rf[2 * e2] - mb.x, rf[2 * e2 + 1] - mb.y);
            whi[e2] = *reinterpret_cast<const uint32_t*>(&h2);
            wmid[e2] = *reinterpret_cast<const uint32_t*>(&m2);
            wlo[e2] = *reinterpret_cast<const uint32_t*>(&l2);
        }
        const int off = ((khalf * 2 + h) ^ swz) * 16;
        *reinterpret_cast<uint4*>(a_hi + off) = make_uint4(whi[0], whi[1], whi[2], whi[3]);
        *reinterpret_cast<uint4*>(a_hi + A_SPLIT_BYTES + off) = make_uint4(wmid[0], wmid[1], wmid[2], wmid[3]);
        *reinterpret_cast<uint4*>(a_hi + 2 * A_SPLIT_BYTES + off) = make_uint4(wlo[0], wlo[1], wlo[2], wlo[3]);
    }
    return bound;
}

template <bool TANGENT, int CBU, bool WIDE>
__global__ void __launch_bounds__(THREADS, 1)
k_eval_tcx(const Args a, const __grid_constant__ CUtensorMap map_hi, const __grid_constant__ CUtensorMap map_mid,
           const __grid_constant__ CUtensorMap map_lo, const __grid_constant__ CUtensorMap map_out)
{
    using L = Lay<CBU, WIDE>;
    constexpr int SA = L::SA, SB = L::SB;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    if (a.sel && *a.sel != a.sel_id) return; // uniform over the grid: nothing has been set up yet
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const uint32_t smem_base = smem_u32(smem);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::BARRIERS);
    const uint32_t bar_full_a = smem_u32(bars + 0);                        // [SA], count PRODUCER_WARPS / 2
    const uint32_t bar_full_b = smem_u32(bars + MAX_SA);                   // [SB], count 1 + tx bytes
    const uint32_t bar_empty_a = smem_u32(bars + MAX_SA + MAX_SB);         // [SA], count 1 (tcgen05.commit)
    const uint32_t bar_empty_b = smem_u32(bars + 2 * MAX_SA + MAX_SB);     // [SB], count 1 (tcgen05.commit; CBU = 1: unused,
                                                                           //       the weight slot follows the Phi slot)
    constexpr int NB = 2 * (MAX_SA + MAX_SB);
    const uint32_t bar_tmem_full = smem_u32(bars + NB);                    // [2], count 1: all MMAs of the unit retired
    const uint32_t bar_tmem_empty = smem_u32(bars + NB + 2);               // [2], count EPILOGUE_WARPS: accumulators drained
    const uint32_t bar_cfull = smem_u32(bars + NB + 4);                    // [CDEPTH], count 1 + tx bytes
    const uint32_t bar_cempty = smem_u32(bars + NB + 4 + CDEPTH);          // [CDEPTH], count PRODUCER_WARPS / 2
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + NB + 4 + 2 * CDEPTH);
    static_assert((NB + 4 + 2 * CDEPTH + 1) * 8 <= 512, "barrier block");
    __shared__ double s_exp[16]; // 2^(i / 16); static: the look-up address needs no base register
    float* s_colscale = reinterpret_cast<float*>(smem + L::COLSCALE);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < SA; ++s) {
            mbar_init(bar_full_a + 8 * s, PRODUCER_WARPS / 2);
            mbar_init(bar_empty_a + 8 * s, 1);
        }
        for (int s = 0; s < SB; ++s) {
            mbar_init(bar_full_b + 8 * s, 1);
            mbar_init(bar_empty_b + 8 * s, 1);
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
    if (threadIdx.x >= 64 && threadIdx.x < 80) s_exp[threadIdx.x - 64] = exp2((double)(threadIdx.x - 64) / 16.0);
    if (warp == WARP_TMA) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)), "r"(TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    // unit = (vertex tile, group of CBU column blocks); ncbu groups per vertex tile, the last one may hold fewer blocks
    const int64_t unit0 = blockIdx.x, ustride = gridDim.x;
    const int nk = a.Kpad / BK;
    const int tail_ksteps = (a.Ktot - (nk - 1) * BK + 15) >> 4; // K=16 steps of the last stage that hold real rows (1 or 2)
    const int64_t n_vt = (a.V + TM - 1) / TM;
    const int ncbu = (a.ncb + CBU - 1) / CBU;
    const int64_t n_units = n_vt * ncbu;

    if (warp == WARP_TMA) {
        // ================= TMA producer: weight tiles + centre tiles =================
        uint32_t it = 0, ic = 0, ib = 0; // stages, centre tiles, weight slots issued so far
        int kc = 0;
        const uint32_t my_units = (uint32_t)((n_units - unit0 + ustride - 1) / ustride);
        const uint32_t total = my_units * (uint32_t)nk;
        for (int64_t u = unit0; u < n_units; u += ustride) {
            const int cb0 = (int)(u % ncbu) * CBU;
            const int nj = min(CBU, a.ncb - cb0);
            for (int kb = 0; kb < nk; ++kb, ++it) {
                while (ic < total && ic < it + CDEPTH) { // centre tiles run ahead; a busy slot is retried at the next stage
                    const int c = ic % CDEPTH;
                    if (!mbar_try_wait(bar_cempty + 8 * c, ((ic / CDEPTH) & 1) ^ 1)) break;
                    if (elect_one()) {
                        mbar_expect_tx(bar_cfull + 8 * c, C_TILE_BYTES + S_TILE_BYTES + R_TILE_BYTES);
                        bulk_load_1d(smem_base + L::CENTRES + c * C_TILE_BYTES, a.ctab + kc * BK, C_TILE_BYTES, bar_cfull + 8 * c);
                        bulk_load_1d(smem_base + L::SC + c * S_TILE_BYTES, a.csc + kc * BK, S_TILE_BYTES, bar_cfull + 8 * c);
                        bulk_load_1d(smem_base + L::ROWEXP + c * R_TILE_BYTES, a.pw + kc * BK, R_TILE_BYTES, bar_cfull + 8 * c);
                    }
                    ++ic;
                    if (++kc == nk) kc = 0;
                }
                if constexpr (WIDE) { // one slot holds the stage's weight tiles of both column blocks, block j at row 120 j of each split
                    const int s = ib % SB;
                    const uint32_t ph = (ib / SB) & 1;
                    mbar_wait(bar_empty_b + 8 * s, ph ^ 1);
                    if (elect_one()) {
                        const uint32_t sb = smem_base + L::B_RING + s * L::B_SLOT;
                        mbar_expect_tx(bar_full_b + 8 * s, nj * 3 * B_SPLIT_BYTES);
                        for (int j = 0; j < nj; ++j) {
                            tma_load_2d(sb + j * B_SPLIT_BYTES, &map_hi, bar_full_b + 8 * s, kb * BK, (cb0 + j) * CB);
                            tma_load_2d(sb + L::B_SPLIT + j * B_SPLIT_BYTES, &map_mid, bar_full_b + 8 * s, kb * BK, (cb0 + j) * CB);
                            tma_load_2d(sb + 2 * L::B_SPLIT + j * B_SPLIT_BYTES, &map_lo, bar_full_b + 8 * s, kb * BK, (cb0 + j) * CB);
                        }
                    }
                    __syncwarp();
                    ++ib;
                } else {
                    for (int j = 0; j < nj; ++j, ++ib) {
                        const int s = ib % SB;
                        const uint32_t ph = (ib / SB) & 1;
                        // CBU = 1: ib == it and SB == SA -- the weight slot is free when the stage's Phi slot is
                        mbar_wait((CBU == 1 ? bar_empty_a : bar_empty_b) + 8 * s, ph ^ 1);
                        if (elect_one()) {
                            const uint32_t sb = smem_base + L::B_RING + s * L::B_SLOT;
                            mbar_expect_tx(bar_full_b + 8 * s, 3 * B_SPLIT_BYTES);
                            tma_load_2d(sb, &map_hi, bar_full_b + 8 * s, kb * BK, (cb0 + j) * CB);
                            tma_load_2d(sb + B_SPLIT_BYTES, &map_mid, bar_full_b + 8 * s, kb * BK, (cb0 + j) * CB);
                            tma_load_2d(sb + 2 * B_SPLIT_BYTES, &map_lo, bar_full_b + 8 * s, kb * BK, (cb0 + j) * CB);
                        }
                        __syncwarp();
                    }
                }
            }
        }
    } else if (warp == WARP_MMA) {
        // ================= MMA issuer =================
        uint32_t it = 0, ib = 0, unit_iter = 0;
        const uint64_t desc0 = make_desc_sw64(smem_base);
        const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
        const bool wide_digit = *a.hbits >= 9; // mid x lo is then below 2^-32 of the full size: two MMAs (and their operand reads) less
        for (int64_t u = unit0; u < n_units; u += ustride, ++unit_iter) {
            const int cb0 = (int)(u % ncbu) * CBU;
            const int nj = min(CBU, a.ncb - cb0);
            // CBU = 1: two units in flight (ping-pong), the epilogue drained this one's slot two units ago;
            // CBU = 2: the unit's column blocks fill tensor memory, the epilogue drained it one unit ago
            const int ab = CBU == 1 ? (int)(unit_iter & 1) : 0;
            mbar_wait(bar_tmem_empty + 8 * ab, ((CBU == 1 ? unit_iter >> 1 : unit_iter) & 1) ^ 1);
            tc_fence_after();
            for (int kb = 0; kb < nk; ++kb, ++it) {
                const int s = it % SA;
                mbar_wait(bar_full_a + 8 * s, (it / SA) & 1);
                const int nsteps = kb != nk - 1 ? 2 : tail_ksteps; // the last stage: steps of pure zero padding are skipped
                // WIDE: one instruction of N = 120 nj' columns serves all the unit's blocks (block j at TMEM column 120 j)
                const int ngrp = WIDE ? 1 : nj;
                for (int j = 0; j < ngrp; ++j, ++ib) {
                    const int sbs = ib % SB;
                    mbar_wait(bar_full_b + 8 * sbs, (ib / SB) & 1);
                    tc_fence_after();
                    if (elect_one()) {
                        const int cb_last = WIDE ? cb0 + nj - 1 : cb0 + j;
                        const int real = min(CB, 3 * a.F - cb_last * CB); // columns of the (last) block that exist
                        const int ncols = min(WIDE ? 2 * CB : NMMA, ((WIDE ? (nj - 1) * CB : 0) + real + 15) & ~15);
                        const uint32_t idesc = make_idesc(ncols);
                        const uint32_t d0 = tmem_u + (CBU == 1 ? ab : j) * UNIT_COLS, d1 = d0 + L::ACC1;
                        const uint64_t a_hi = desc0 + (uint64_t)(s * (A_STAGE_BYTES >> 4));
                        const uint64_t a_mid = a_hi + (A_SPLIT_BYTES >> 4), a_lo = a_mid + (A_SPLIT_BYTES >> 4);
                        const uint64_t b_hi = desc0 + (uint64_t)((L::B_RING + sbs * L::B_SLOT) >> 4);
                        const uint64_t b_mid = b_hi + (L::B_SPLIT >> 4), b_lo = b_mid + (L::B_SPLIT >> 4);
                        for (int ks = 0; ks < nsteps; ++ks) {
                            const uint64_t o = 2 * ks;                     // 32 bytes = 16 FP16 along K
                            const uint32_t acc = (kb != 0 || ks != 0) ? 1u : 0u;
                            umma_f16(d0, a_hi + o, b_hi + o, idesc, acc);  // integers: exact
                            if (a.dbg_mode & 8) continue;
                            umma_f16(d1, a_hi + o, b_mid + o, idesc, acc);
                            umma_f16(d1, a_mid + o, b_hi + o, idesc, 1);
                            umma_f16(d1, a_hi + o, b_lo + o, idesc, 1);
                            umma_f16(d1, a_lo + o, b_hi + o, idesc, 1);
                            umma_f16(d1, a_mid + o, b_mid + o, idesc, 1);
                            if (wide_digit) continue;
                            umma_f16(d1, a_mid + o, b_lo + o, idesc, 1);   // 2^-14 per product: 2^-14-2h of the full size, not
                            umma_f16(d1, a_lo + o, b_mid + o, idesc, 1);   // negligible once the digit is narrow (h = 5: 2^-24)
                        }
                        if (CBU != 1) umma_commit(bar_empty_b + 8 * sbs);            // the weight slot is free once these MMAs have read it
                        if (j == ngrp - 1) {
                            umma_commit(bar_empty_a + 8 * s);                        // and the Phi slot after the last column block
                            if (kb == nk - 1) umma_commit(bar_tmem_full + 8 * ab);   // all accumulators of the unit complete
                        }
                    }
                    __syncwarp();
                }
            }
        }
    } else if (warp < PRODUCER_WARPS) {
        // ================= Phi producers =================
        const int pt = threadIdx.x;             // 0..511
        const int row = pt & (TM - 1);
        const int khalf = (pt >> 7) & 1;        // which half of the stage's 32 k this thread generates
        const int grp = pt >> 8;                // producer group: stages of its parity
        const float4 nrm4 = *reinterpret_cast<const float4*>(a.norm);
        const double ox = (double)a.origin[0], oy = (double)a.origin[1], oz = (double)a.origin[2];
        const int swz = (row >> 1) & 3;
        uint32_t it = 0;
        for (int64_t u = unit0; u < n_units; u += ustride) {
            const int64_t vt = u / ncbu;
            const int64_t v = vt * TM + row;
            ProducerRow pr;
            pr.pxf = 0.f, pr.pyf = 0.f, pr.pzf = 0.f;
            if (v < a.V) {
                pr.pxf = a.P[3 * v];
                pr.pyf = a.P[3 * v + 1];
                pr.pzf = a.P[3 * v + 2];
            }
            pr.qx = (double)pr.pxf - ox, pr.qy = (double)pr.pyf - oy, pr.qz = (double)pr.pzf - oz;
            pr.pp = fma(pr.qx, pr.qx, fma(pr.qy, pr.qy, pr.qz * pr.qz));
            float bound = 0.f; // this thread's share of sum_k a_hi 2^(h - s_k) >= sum_k |a_hi b_hi| for any column
            const uint32_t it0 = it;
            it += nk;
            for (int kb = (int)((grp ^ it0) & 1); kb < nk; kb += 2) {
                const uint32_t itk = it0 + kb;
                const int s = itk % SA;
                const uint32_t ph = (itk / SA) & 1;
                const int cs = itk % CDEPTH;
                mbar_wait(bar_cfull + 8 * cs, (itk / CDEPTH) & 1); // the stage's centre tile
                mbar_wait(bar_empty_a + 8 * s, ph ^ 1);            // the Phi slot: the MMAs of stage itk - SA retired
                uint8_t* a_hi = smem + s * A_STAGE_BYTES + row * (BK * 2);
                const double4* s_ctr = reinterpret_cast<const double4*>(smem + L::CENTRES + cs * C_TILE_BYTES);
                const double* s_sc = reinterpret_cast<const double*>(smem + L::SC + cs * S_TILE_BYTES);
                const float* s_pw = reinterpret_cast<const float*>(smem + L::ROWEXP + cs * R_TILE_BYTES);
                const int k0 = kb * BK;
                if (!(kb == nk - 1 && khalf >= tail_ksteps)) { // else: nothing but zero padding, the MMA issuer skips the step
                    if (k0 + BK <= a.N) bound = fill_stage_half<true>(a, pr, nrm4, a_hi, s_ctr, s_sc, s_pw, s_exp, khalf, swz, k0, bound);
                    else bound = fill_stage_half<false>(a, pr, nrm4, a_hi, s_ctr, s_sc, s_pw, s_exp, khalf, swz, k0, bound);
                }
                fence_proxy_async(); // generic-proxy stores -> visible to the tensor core (async proxy)
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive(bar_full_a + 8 * s);
                    mbar_arrive(bar_cempty + 8 * cs);
                }
            }
            // four threads share a row (two k halves x two groups): each share within a quarter of the range keeps the sum inside
            if (bound > ((a.dbg_mode & 32) ? 1024.0f : 4194304.0f) && v < a.V) atomicExch(&a.flags[FD_FLAG_EVAL_INEXACT], 1);
        }
    } else {
        // ================= epilogue warps: TMEM -> registers -> (transpose in shared memory) -> global =================
        const int ew = warp - WARP_EPI0;
        const int q = warp & 3;               // TMEM lane quarter this warp may access
        float* stg0 = reinterpret_cast<float*>(smem + L::EPI_STAGING) + ew * (L::EPI_BUFS * EPI_WARP_FLOATS);
        const int et = threadIdx.x - 32 * WARP_EPI0;
        uint32_t unit_iter = 0, nstore = 0;
        const bool resident_cs = a.ncb <= COLSCALE_RESIDENT_BLOCKS;
        const float hs = exp2f(-2.0f * (float)*a.hbits); // both operands carry 2^h
        if (resident_cs) {
            for (int t = et; t < a.ncb * CB; t += 32 * EPILOGUE_WARPS) s_colscale[t] = a.colscale[t];
            asm volatile("bar.sync 2, %0;" ::"n"(32 * EPILOGUE_WARPS) : "memory");
        }
        for (int64_t u = unit0; u < n_units; u += ustride, ++unit_iter) {
            const int cb0 = (int)(u % ncbu) * CBU;
            const int nj = min(CBU, a.ncb - cb0);
            const int64_t vt = u / ncbu;
            if (!resident_cs) { // the column scales of the unit's blocks
                asm volatile("bar.sync 2, %0;" ::"n"(32 * EPILOGUE_WARPS) : "memory");
                for (int t = et; t < nj * CB; t += 32 * EPILOGUE_WARPS) s_colscale[t] = a.colscale[cb0 * CB + t];
                asm volatile("bar.sync 2, %0;" ::"n"(32 * EPILOGUE_WARPS) : "memory");
            }
            const int ab = CBU == 1 ? (int)(unit_iter & 1) : 0;
            const int chunk0 = ew >> 2; // with eight epilogue warps: the even (0) or the odd (1) column chunks
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
                if (a.falloff_out && cb0 == 0 && chunk0 == 0) a.falloff_out[v] = fo;
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
            bool released = false;
            mbar_wait(bar_tmem_full + 8 * ab, (CBU == 1 ? unit_iter >> 1 : unit_iter) & 1);
            tc_fence_after();
#pragma unroll 1
            for (int j = 0; j < nj; ++j) {
                const int cb = cb0 + j;
                const int f_base = cb * (CB / 3);
                const int nframes = min(CB / 3, a.F - f_base);
                const float* s_cs = s_colscale + (resident_cs ? cb : j) * CB;
#pragma unroll 1
                for (int ch = chunk0; ch * EPI_FRAMES < nframes; ch += EPILOGUE_WARPS / 4) {
                    if (a.dbg_mode & 2) continue;
                    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (CBU == 1 ? ab * UNIT_COLS : j * L::BLOCK_COLS) + ch * EPI_COLS;
                    float acc[EPI_COLS], acc1[EPI_COLS];
                    tmem_ld16(taddr, acc);
                    tmem_ld8(taddr + 16, acc + 16);
                    tmem_ld16(taddr + L::ACC1, acc1);
                    tmem_ld8(taddr + L::ACC1 + 16, acc1 + 16);
                    float* stg = stg0 + (L::EPI_BUFS == 1 ? 0 : (nstore & 1) * EPI_WARP_FLOATS);
                    if (vec && lane == 0) { // the staging buffer is free: the bulk store that read it last has done so
                        if (L::EPI_BUFS == 1) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                        else asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                    }
                    __syncwarp();
                    tmem_ld_wait();
                    if (CBU != 1 && j == nj - 1 && (ch + EPILOGUE_WARPS / 4) * EPI_FRAMES >= nframes) {
                        // single-buffered accumulators: the MMAs of the next unit may start as soon as this warp has read its last values
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(bar_tmem_empty + 8 * ab);
                        released = true;
                    }
#pragma unroll
                    for (int c = 0; c < EPI_COLS; ++c) acc[c] = (acc[c] + acc1[c]) * (s_cs[ch * EPI_COLS + c] * hs); // powers of two: exact
                    if (vec) {
                        // out = P + disp * falloff (a skipped vertex has falloff 0 and keeps P exactly), transposed through shared
                        // memory into [frame][vertex][xyz] rows that one lane hands to the bulk-copy engine
                        float* dst = stg + lane * 3;
#pragma unroll
                        for (int col = 0; col < EPI_COLS; ++col) {
                            const int i = col / 3, k = col - 3 * i;
                            const float p = k == 0 ? px : (k == 1 ? py : pz);
                            dst[i * 96 + k] = fmaf(acc[col], fo, p);
                        }
                        fence_proxy_async();
                        __syncwarp();
                        if (lane == 0) {
                            asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                                         ::"l"(reinterpret_cast<uint64_t>(&map_out)), "r"(smem_u32(stg)),
                                           "r"((int)(v_warp0 * 3)), "r"(f_base + ch * EPI_FRAMES)
                                         : "memory");
                            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                        }
                        ++nstore;
                    } else {
                        // general path (V not a multiple of 4, unaligned output, tangent projection): per-lane stores
                        const int fcnt = min(EPI_FRAMES, nframes - ch * EPI_FRAMES);
#pragma unroll
                        for (int i = 0; i < EPI_FRAMES; ++i) {
                            if (i < fcnt) {
                                float d[3] = {acc[3 * i], acc[3 * i + 1], acc[3 * i + 2]};
                                if (TANGENT) project_to_tangents(tu, tv, tn, d);
                                if (valid) {
                                    float* dst = a.P_out + ((size_t)(f_base + ch * EPI_FRAMES + i) * (size_t)a.V + (size_t)v) * 3;
                                    dst[0] = skip ? px : px + d[0] * fo;
                                    dst[1] = skip ? py : py + d[1] * fo;
                                    dst[2] = skip ? pz : pz + d[2] * fo;
                                }
                            }
                        }
                    }
                }
            }
            if (L::EPI_BUFS == 1 && vec && lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            if (!released) {
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_tmem_empty + 8 * ab); // this warp's share of the unit is drained
            }
        }
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); // shared memory outlives the last bulk stores
    }

    tc_fence_before();
    __syncthreads();
    if (warp == WARP_TMA) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS));
    }
}

// row k: rowmax[k] = max_c |w_kc| 2^e_c in [2^-r_k-1, 2^-r_k) (r_k = 30 for a zero row) and s_k = r_k / 2 -> rowexp[k]
__global__ void __launch_bounds__(128) k_tcx_rowscale(const double* __restrict__ W, int ldw, int N, int np, int ncol,
                                                      const float* __restrict__ norm, const float* __restrict__ scale,
                                                      int* __restrict__ rowexp, float* __restrict__ rowmax)
{
    __shared__ double s_mx[4];
    const int k = blockIdx.x;
    double mx = 0.0;
    if (k < N + 4)
        for (int c = threadIdx.x; c < ncol; c += 128) {
            const double w = fabs(tc_weight(W, ldw, N, np, k, c, norm)) * (double)scale[c];
            if (w > mx) mx = w; // a NaN never wins; non-finite weights are flagged by k_tc_colscale
        }
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0) s_mx[threadIdx.x >> 5] = mx;
    __syncthreads();
    if (threadIdx.x != 0) return;
    mx = fmax(fmax(s_mx[0], s_mx[1]), fmax(s_mx[2], s_mx[3]));
    int r = 30;
    if (mx > 0.0 && isfinite(mx)) {
        int e;
        frexp(mx, &e);            // mx = m 2^e, m in [0.5, 1), e <= 0 after the column scaling
        r = min(30, max(0, -e));
    }
    rowexp[k] = r >> 1;
    rowmax[k] = isfinite(mx) ? (float)mx : 0.f;
}

// B_i = sum_j rowmax[j] phi_j(c_i): the size of the leading-digit sum (per 2^2h) at control point i, one warp per i; the maximum's
// bits (a non-negative double) -> bmax_bits.  FP32 is plenty for a bound.
__global__ void __launch_bounds__(256) k_tcx_bound(const float* __restrict__ rest, const double* __restrict__ radii, int N,
                                                   const float* __restrict__ rowmax, unsigned long long* __restrict__ bmax_bits)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int i = blockIdx.x * 8 + warp;
    if (i >= N) return;
    const float xi = rest[3 * i], yi = rest[3 * i + 1], zi = rest[3 * i + 2];
    float s = 0.f;
    for (int j = lane; j < N; j += 32) {
        const float dx = xi - rest[3 * j], dy = yi - rest[3 * j + 1], dz = zi - rest[3 * j + 2];
        const float R = (float)radii[j];
        s += rowmax[j] * __expf(-(dx * dx + dy * dy + dz * dz) / (R * R));
    }
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0 && s >= 0.f) atomicMax(bmax_bits, (unsigned long long)__double_as_longlong((double)s));
}

// h: the widest leading digit with 4 x 2^2h x B <= 2^24, between 2 and 11; B includes the affine rows (1, and coordinates taken
// up to four times the rig's half extent)
__global__ void k_tcx_hbits(unsigned long long* __restrict__ bmax_bits, const float* __restrict__ rowmax, int N,
                            int* __restrict__ hbits)
{
    double B = __longlong_as_double((long long)*bmax_bits);
    B += (double)rowmax[N] + 4.0 * ((double)rowmax[N + 1] + (double)rowmax[N + 2] + (double)rowmax[N + 3]);
    int e = 0;
    if (B > 0.0 && isfinite(B)) frexp(B, &e); // B < 2^e
    const int h = (22 - e) / 2;
    *hbits = h > 11 ? 11 : (h < 2 ? 2 : h);
    *bmax_bits = 0ull; // ready for the next pack (stream order)
}

// W^T digit / mid / lo [ncol_pad][Kpad] FP16, k contiguous; 32x32 tile transpose through shared memory
__global__ void __launch_bounds__(256) k_tcx_pack(const double* __restrict__ W, int ldw, int N, int np, int ncol, int ncol_pad,
                                                  int Kpad, const float* __restrict__ norm, const float* __restrict__ scale,
                                                  const int* __restrict__ rowexp, const int* __restrict__ hbits,
                                                  __half* __restrict__ Wt_hi, __half* __restrict__ Wt_mid, __half* __restrict__ Wt_lo,
                                                  const double4* __restrict__ ctab, double4* __restrict__ ctab_eff,
                                                  float* __restrict__ pw)
{
    __shared__ double s_t[32][33];
    const int c0 = blockIdx.x * 32, k0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int h = *hbits;
    if (blockIdx.x == 0 && ty == 0 && k0 + tx < Kpad) { // the evaluation's per-solve centre records: row k's digit exponent h - s_k
        const int k = k0 + tx, e = h - rowexp[k];       // folded into the constant term of t, and 2^(h - s_k) as a float
        double4 c = ctab[k];
        c.w += (double)e;
        ctab_eff[k] = c;
        pw[k] = __int_as_float((127 + e) << 23);
    }
    for (int r = ty; r < 32; r += 8) {
        const int k = k0 + r, c = c0 + tx;
        double v = 0.0;
        if (c < ncol && k < Kpad) v = scalbn(tc_weight(W, ldw, N, np, k, c, norm) * (double)scale[c], h + rowexp[k]);
        s_t[r][tx] = v;
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const int c = c0 + r, k = k0 + tx;
        if (c < ncol_pad && k < Kpad) {
            const double v = s_t[tx][r];
            const float hf = isfinite(v) ? rintf((float)v) : 0.f;
            const float rf = isfinite(v) ? (float)(v - (double)hf) : 0.f;
            const __half m = __float2half_rn(rf);
            Wt_hi[(size_t)c * Kpad + k] = __float2half_rn(hf);
            Wt_mid[(size_t)c * Kpad + k] = m;
            Wt_lo[(size_t)c * Kpad + k] = __float2half_rn(rf - __half2float(m));
        }
    }
}

} // namespace tcx

int fd_tcx_ncb(int F) { return (3 * F + tcx::CB - 1) / tcx::CB; }
int fd_tcx_col_pad(int F) { return fd_tcx_ncb(F) * tcx::CB; }

// the three weight tables for the weights currently in fd_w_src(m) (called from fd_launch_pack)
cudaError_t fd_launch_pack_tcx(fd_ctx* ctx, fd_model* m)
{
    cudaStream_t s = ctx->stream;
    const int ncol = 3 * m->F, ncol_pad = fd_tcx_col_pad(m->F), Kpad = fd_tc_kpad(m->N);
    // columns to max |w| in [1/2, 1), then the rows' deficits, the digit width they allow, the tables
    tcx::k_tc_colscale<<<(ncol_pad + 31) / 32, 256, 0, s>>>(fd_w_src(m), m->ldw, m->N, m->np, ncol, ncol_pad, 0, 0, m->d_tc_norm,
                                                             m->d_tc_unscale, m->d_tc_scale, m->d_flags);
    tcx::k_tcx_rowscale<<<Kpad, 128, 0, s>>>(fd_w_src(m), m->ldw, m->N, m->np, ncol, m->d_tc_norm, m->d_tc_scale,
                                             m->d_tcx_rowexp, m->d_tcx_rowmax);
    tcx::k_tcx_bound<<<(m->N + 7) / 8, 256, 0, s>>>(m->d_rest, m->d_radii, m->N, m->d_tcx_rowmax,
                                                    reinterpret_cast<unsigned long long*>(m->d_tcx_meta));
    tcx::k_tcx_hbits<<<1, 1, 0, s>>>(reinterpret_cast<unsigned long long*>(m->d_tcx_meta), m->d_tcx_rowmax, m->N,
                                     m->d_tcx_rowexp + Kpad);
    dim3 grid((ncol_pad + 31) / 32, (Kpad + 31) / 32);
    tcx::k_tcx_pack<<<grid, 256, 0, s>>>(fd_w_src(m), m->ldw, m->N, m->np, ncol, ncol_pad, Kpad, m->d_tc_norm, m->d_tc_scale,
                                         m->d_tcx_rowexp, m->d_tcx_rowexp + Kpad, (__half*)m->d_tc_wt_hi,
                                         (__half*)m->d_tcx_wt_mid, (__half*)m->d_tc_wt_lo, m->d_ctab_tcx, m->d_tcx_ctab_eff,
                                         m->d_tcx_pw);
    ctx->launches += 5;
    if (!tcx::make_map((CUtensorMap*)m->tc_map_hi, m->d_tc_wt_hi, Kpad, ncol_pad, tcx::BK, tcx::CB) ||
        !tcx::make_map((CUtensorMap*)m->tcx_map_mid, m->d_tcx_wt_mid, Kpad, ncol_pad, tcx::BK, tcx::CB) ||
        !tcx::make_map((CUtensorMap*)m->tc_map_lo, m->d_tc_wt_lo, Kpad, ncol_pad, tcx::BK, tcx::CB))
        return cudaErrorInvalidValue;
    return cudaGetLastError();
}

namespace tcx {
template <bool TANGENT, int CBU, bool WIDE>
static cudaError_t set_smem_attr()
{
    return cudaFuncSetAttribute(k_eval_tcx<TANGENT, CBU, WIDE>, cudaFuncAttributeMaxDynamicSharedMemorySize, Lay<CBU, WIDE>::ALLOC);
}
template <bool TANGENT, int CBU, bool WIDE>
static void launch(int grid, cudaStream_t st, const Args& a, const CUtensorMap& mh, const CUtensorMap& mm, const CUtensorMap& ml,
                   const CUtensorMap& mo)
{
    k_eval_tcx<TANGENT, CBU, WIDE><<<grid, THREADS, Lay<CBU, WIDE>::ALLOC, st>>>(a, mh, mm, ml, mo);
}
template <bool TANGENT>
static cudaError_t set_smem_attr_all()
{
    cudaError_t e = set_smem_attr<TANGENT, 1, false>();
    if (e == cudaSuccess) e = set_smem_attr<TANGENT, 2, false>();
    if (e == cudaSuccess) e = set_smem_attr<TANGENT, 2, true>();
    return e;
}
template <bool TANGENT>
static void launch_variant(int cbu, bool wide, int grid, cudaStream_t st, const Args& a, const CUtensorMap& mh,
                           const CUtensorMap& mm, const CUtensorMap& ml, const CUtensorMap& mo)
{
    if (cbu == 1) launch<TANGENT, 1, false>(grid, st, a, mh, mm, ml, mo);
    else if (wide) launch<TANGENT, 2, true>(grid, st, a, mh, mm, ml, mo);
    else launch<TANGENT, 2, false>(grid, st, a, mh, mm, ml, mo);
}
} // namespace tcx

cudaError_t fd_eval_tcx_setup(fd_ctx* ctx)
{
    (void)ctx;
    cudaError_t e = tcx::set_smem_attr_all<false>();
    if (e == cudaSuccess) e = tcx::set_smem_attr_all<true>();
    return e;
}

// a view of `m` that evaluates the frames from f_begin on (a multiple of 40 frames = one 120-column block)
bool fd_tcx_view_frames(const fd_model* m, fd_model* view, int f_begin)
{
    const int col0 = 3 * f_begin, ncol_pad = fd_tcx_col_pad(m->F), Kpad = fd_tc_kpad(m->N);
    if (col0 % tcx::CB != 0 || col0 >= ncol_pad) return false;
    view->d_tc_unscale = m->d_tc_unscale + col0;
    view->d_tc_scale = m->d_tc_scale + col0;
    const size_t off = (size_t)col0 * Kpad;
    return tcx::make_map((CUtensorMap*)view->tc_map_hi, (unsigned short*)m->d_tc_wt_hi + off, Kpad, ncol_pad - col0, tcx::BK, tcx::CB) &&
           tcx::make_map((CUtensorMap*)view->tcx_map_mid, (unsigned short*)m->d_tcx_wt_mid + off, Kpad, ncol_pad - col0, tcx::BK, tcx::CB) &&
           tcx::make_map((CUtensorMap*)view->tc_map_lo, (unsigned short*)m->d_tc_wt_lo + off, Kpad, ncol_pad - col0, tcx::BK, tcx::CB);
}

cudaError_t fd_launch_eval_tcx(fd_ctx* ctx, const fd_model* m, const float* P, int64_t V, const float* dist2, const float* tu,
                               const float* tv, const float* nrm, float* P_out, float* falloff_out, const int* sel, int sel_id)
{
    if (V <= 0) return cudaSuccess;
    tcx::Args a;
    a.sel = sel;
    a.sel_id = sel_id;
    a.dbg_mode = ctx->dbg.has_tc_debug ? ctx->dbg.tc_debug : 0;
    a.ctab = m->d_tcx_ctab_eff;
    a.csc = m->d_csc_tcx;
    a.origin = m->d_rest;
    a.norm = m->d_tc_norm;
    a.colscale = m->d_tc_unscale;
    a.N = m->N;
    a.Kpad = fd_tc_kpad(m->N);
    a.Ktot = m->N + m->np;
    a.F = m->F;
    a.ncb = fd_tcx_ncb(m->F);
    a.pw = m->d_tcx_pw;
    a.hbits = m->d_tcx_rowexp + a.Kpad;
    a.flags = m->d_flags;
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
    if (a.vec_store_ok && !(ctx->tc_out_ptr == P_out && ctx->tc_out_V == V && ctx->tc_out_F == m->F)) {
        ctx->tc_out_ptr = nullptr;
        if (tcx::make_out_map((CUtensorMap*)ctx->tc_out_map, P_out, V, m->F, tcx::EPI_FRAMES)) {
            ctx->tc_out_ptr = P_out;
            ctx->tc_out_V = V;
            ctx->tc_out_F = m->F;
        } else {
            a.vec_store_ok = 0;
        }
    }
    const CUtensorMap& mo = *(const CUtensorMap*)ctx->tc_out_map;
    const CUtensorMap& mh = *(const CUtensorMap*)m->tc_map_hi;
    const CUtensorMap& mm = *(const CUtensorMap*)m->tcx_map_mid;
    const CUtensorMap& ml = *(const CUtensorMap*)m->tc_map_lo;
    // two column blocks per unit (the Phi tiles are generated once for both) whenever the batch has two
    const int cbu = (a.ncb >= 2 && ctx->dbg.tcx_cbu != 1) ? 2 : 1;
    const int64_t n_units = ((V + tcx::TM - 1) / tcx::TM) * ((a.ncb + cbu - 1) / cbu);
    const int grid = (int)(n_units < ctx->sm_count ? n_units : ctx->sm_count);
    // FD_TCX_NARROW=1: one MMA of N = 128 per column block instead of one of N = 240 for the unit's two (for comparison)
    const bool wide = cbu == 2 && ctx->dbg.tcx_narrow == 0;
    if (a.do_tangent) tcx::launch_variant<true>(cbu, wide, grid, ctx->stream, a, mh, mm, ml, mo);
    else tcx::launch_variant<false>(cbu, wide, grid, ctx->stream, a, mh, mm, ml, mo);
    ctx->launches += 1;
    return cudaGetLastError();
}
