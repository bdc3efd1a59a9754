// fd_eval64.cu -- (1) the cancellation estimate behind FD_EVAL_AUTO and (2) the FP64 evaluation for wide 3F on the
// FP64 tensor pipe.
//
// (1) The evaluation sum  sum_j w_j phi_j(x)  of an RBF interpolant cancels: with a Gaussian of radius 2 x spacing the
// terms are ~10^4 times larger than the result.  Every FP32 evaluation (FMA/SFU or tensor cores; phi itself is only
// good to ~2^-22 in FP32) therefore errs by about 2^-24 S, S = the size of the cancelling terms, whatever the kernel
// does.  After each solve k_wmax + k_cancel_select measure
//        S = max_i sum_j max_c |w_jc| phi_j(c_i)            (the control points stand in for the vertices)
// and choose, on the device (no host synchronisation between solve and evaluation), the FP32 kernel the caller's
// eval_path allows with  coef x 2^-24 x S  <=  eval_tolerance x diag  (diag = the rig's bounding-box diagonal; tensor
// cores before FMA/SFU), else the FP64 evaluation.  The evaluation launches every candidate; the ones not chosen return
// at their first instruction.  The reference evaluates in FP64 (alglib::rbfcalc on double[3], SOP_FaceDeform.cpp:411-415).
//
// (2) k_eval64_mma: D[v][3f+k] = sum_j Phi[v][j] W[j][3f+k] with Phi generated on the fly in FP64 and contracted by
// mma.sync.m8n8k4.f64 (DMMA).  A CTA owns 128 vertices x 192 columns (64 frames); per stage of 32 centres its 256
// threads write the 128 x 32 Phi tile into shared memory (16 basis functions each: expanded distance, 4 DFMA;
// exp2 / sqrt / log from few DFMAs, fd_eval_common.cuh), the weight tile arrives with cp.async one stage ahead, and
// each of the 16 warps contracts its 64 x 24 sub-tile: 24 accumulator tiles, 11 fragment loads per 24 DMMAs.  Phi is computed
// once per 192 columns instead of once per 1-2 frames (k_eval_f64), and its generation for stage s + 1 is interleaved,
// two values per K = 4 step, with the DMMAs of stage s in every warp's instruction stream.  (A warp-specialised variant -- 4 producer warps for Phi, 8 consumer warps for the DMMAs -- was measured
// SLOWER, 2.98 ms against 2.41 ms at BASELINE configs[1]: DMMA and DFMA share the FP64 pipe, a DMMA holds it for 16
// cycles, and a warp of dependent DFMA chains scheduled beside DMMA warps starves; the look-ahead LU failed the same way.)
// Polynomial rows ride along as extra K rows [1, x, y, z].  Epilogue = the SOP's (gate, tangent
// projection, falloff, P += disp: SOP_FaceDeform.cpp:405-438), in FP32 after the narrowing of :415.
#include "fd_eval_common.cuh"

namespace {

// ---------------------------------------------------------------------------------------------------------------------
// (1) cancellation estimate + kernel choice
// ---------------------------------------------------------------------------------------------------------------------

// max error of an FP32 evaluation <= coef x 2^-24 x S.  coef = the largest ratio measured against the oracle over
// N = 256 / 1024 / 2048 control points x 120 frames x 4096 vertices (1.5 M values each; tests/test_gpu_round2.py,
// tests/tools/accuracy_probe.py): 1.68 tensor cores (FP16 hi/lo splits, FP32 accumulation in the tensor core),
// 0.93 FMA/SFU -- plus 20 %.  The FMA/SFU ratio falls with the number of centres (the error is a random sum over the
// terms that cancel: GPU 0.93 / 0.77 / 0.79 / 0.65 at N = 256 / 1024 / 2048 / 4096; the CPU emulation of the same
// arithmetic, tests/tools/fp32_error_emulation.py, 1.21 / 1.01 / 0.97 at N = 64 / 256 / 1024), so its coefficient is
// 1.3 at N <= 64 and decreases by 0.07 per doubling of N down to 1.1.  A kernel is eligible while its prediction stays
// within eval_tolerance x diag; the fastest eligible one runs: tensor cores, then FMA/SFU, else FP64.
constexpr double ERR_COEF_TENSOR = 2.0;
// the exact-digit tensor-core kernel (fd_eval_tcx.cu): what is left is the truncation bias of its second accumulator, measured
// 0.044 ... 0.06 x 2^-24 S at N = 256 ... 4096 (profiles/r2_accuracy.log) -- 2.5 x that
constexpr double ERR_COEF_TCX = 0.15;
__host__ __device__ inline double err_coef_simt(int N)
{
    const double c = 1.3 - 0.07 * log2(fmax((double)N, 64.0) / 64.0);
    return c < 1.1 ? 1.1 : c;
}

// wmax[j] = max_c |W[j][c]| over the nrhs solved columns (row j of the row-major weight block)
__global__ void __launch_bounds__(128) k_wmax(const double* __restrict__ W, int ldw, int nrhs, int N, float* __restrict__ wmax)
{
    __shared__ double s_m[4];
    const int j = blockIdx.x;
    if (j >= N) return;
    double m = 0.0;
    for (int c = threadIdx.x; c < nrhs; c += 128) {
        const double v = fabs(W[(size_t)j * ldw + c]);
        m = v > m ? v : m; // a NaN never wins: non-finite weights are flagged by the pack kernels
    }
    for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) s_m[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) wmax[j] = (float)fmax(fmax(s_m[0], s_m[1]), fmax(s_m[2], s_m[3]));
}

// S_i = sum_j wmax[j] phi_j(c_i), one warp per control point i; the last CTA to finish takes the maximum's bits from
// the atomic and settles the choice.  out = { sel, S, diag }.
struct SelectArgs {
    const float* rest;
    const double* radii;
    const float* wmax;
    const float* norm; // k_tc_norm: [4] = bounding-box diagonal of the control points
    int N, kernel;
    int tensor_ok;     // the tensor-core tables exist for these weights (3F wide enough, eval_path allows it)
    int simt_ok;       // eval_path allows the FMA/SFU kernel
    int tcx_ok;        // the exact-digit tensor-core tables exist for these weights
    int want;          // 0 choose; 1 / 2 / 3 forced by eval_precision / eval_path
    float tol;         // eval_tolerance
    unsigned long long* smax_bits; // max S as the bits of a non-negative double
    unsigned* done;
    int* sel;
    double* est;       // [0] S, [1] diag
};

__global__ void __launch_bounds__(256) k_cancel_select(const SelectArgs a)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int i = blockIdx.x * 8 + warp;
    if (i < a.N) {
        const float xi = a.rest[3 * i], yi = a.rest[3 * i + 1], zi = a.rest[3 * i + 2];
        float s = 0.f;
        for (int j = lane; j < a.N; j += 32) {
            const float dx = xi - a.rest[3 * j], dy = yi - a.rest[3 * j + 1], dz = zi - a.rest[3 * j + 2];
            const float r2 = dx * dx + dy * dy + dz * dz;
            const float R = (float)a.radii[j];
            float ph;
            if (a.kernel == FD_KERNEL_GAUSSIAN) ph = __expf(-r2 / (R * R));
            else if (a.kernel == FD_KERNEL_MULTIQUADRIC) ph = sqrtf(r2 + R * R);
            else ph = r2 > 0.f ? fabsf(0.5f * r2 * __logf(r2)) : 0.f;
            s += a.wmax[j] * ph;
        }
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0 && s >= 0.f) atomicMax(a.smax_bits, (unsigned long long)__double_as_longlong((double)s));
    }
    __syncthreads();
    if (threadIdx.x != 0) return;
    __threadfence();
    if (atomicAdd(a.done, 1u) != gridDim.x - 1) return;
    __threadfence();
    const double S = __longlong_as_double((long long)atomicAdd(a.smax_bits, 0ull));
    const double diag = (double)a.norm[4];
    int sel = a.want;
    if (sel == 0) {
        const double tol = a.tol > 0.f ? (double)a.tol : 1e-5;
        const double unit = 5.9604644775390625e-08 * S, lim = tol * diag;
        sel = (a.tensor_ok && ERR_COEF_TENSOR * unit <= lim) ? FD_SEL_TENSOR
            : (a.tcx_ok && ERR_COEF_TCX * unit <= lim) ? FD_SEL_TCX
            : (a.simt_ok && err_coef_simt(a.N) * unit <= lim) ? FD_SEL_SIMT : FD_SEL_FP64;
    }
    *a.sel = sel;
    a.est[0] = S;
    a.est[1] = diag;
    *a.smax_bits = 0ull; // ready for the next solve (stream order)
    *a.done = 0u;
}

// ---------------------------------------------------------------------------------------------------------------------
// (2) FP64 evaluation on the FP64 tensor pipe
// ---------------------------------------------------------------------------------------------------------------------
constexpr int E_TM = 128;          // vertices per CTA tile
constexpr int E_TN = 192;          // columns per CTA tile (64 frames)
constexpr int E_KB = 32;           // centres per stage
constexpr int E_THREADS = 512;     // 16 warps, 4 per SM sub-partition: the fixed issue delays of back-to-back DMMAs and the
                                   // DFMA chains of the basis functions need that many to overlap (2 per sub-partition left the
                                   // tensor pipe 51 % busy)
constexpr int E_LDA = E_KB + 4;    // = 4 mod 16: conflict-free A-fragment loads (see the lane map of fd_dmma884)
constexpr int E_LDB = E_TN + 4;    // = 4 mod 16: conflict-free B-fragment loads
constexpr int E_LDC = E_TN + 1;    // float staging of the accumulators for the epilogue
constexpr int E_STAGE_DOUBLES = E_TM * E_LDA + E_KB * E_LDB;
constexpr int E_SMEM_BYTES = 2 * E_STAGE_DOUBLES * 8 + 2 * E_KB * 5 * 8 + 128 * 16 + 64 * 8;
static_assert(E_TM * E_LDC * 4 <= 2 * E_STAGE_DOUBLES * 8, "the epilogue staging reuses the pipeline buffers");

struct Eval64Args {
    const double4* ctab; // Gaussian: (x, y, z, -1 / R^2); multiquadric / thin plate: (-2 (c - o), |c - o|^2 + prm)
    const float* origin; // o = centre 0
    const double* W;     // (N + np) x ldw
    int ldw, N, np, F;
    int ncol;            // readable columns from W on (ldw minus the offset of a frame-block view)
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
    const int* sel;
    int sel_id;
};

__device__ __forceinline__ void cp_async16_zfill(void* smem_dst, const void* gsrc, bool valid)
{
    const int bytes = valid ? 16 : 0; // src-size 0: the 16 destination bytes are zero-filled, nothing is read
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc),
                 "r"(bytes)
                 : "memory");
}

template <int KERNEL>
__global__ void __launch_bounds__(E_THREADS, 1) k_eval64_mma(const Eval64Args a)
{
    extern __shared__ __align__(16) unsigned char e64_smem[];
    if (a.sel && *a.sel != a.sel_id) return; // FD_EVAL_AUTO settled on an FP32 kernel
    double* s_stage = reinterpret_cast<double*>(e64_smem);                 // [2][A tile | B tile]
    double* s_ctr = s_stage + 2 * E_STAGE_DOUBLES;                         // [2][32][5]: (a, b, c, d, s) of a stage's centres
    double2* s_log = reinterpret_cast<double2*>(s_ctr + 2 * E_KB * 5);     // thin plate: fd_half_log64 table
    double* s_exp = reinterpret_cast<double*>(s_log + 128);                // Gaussian: fd_exp2_64 table
    float* s_C = reinterpret_cast<float*>(e64_smem);                       // epilogue staging, reuses the stage buffers

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (KERNEL == FD_KERNEL_THINPLATE && tid < 128) fd_half_log64_table(s_log, tid);
    if (KERNEL == FD_KERNEL_GAUSSIAN && tid < 64) fd_exp2_64_table(s_exp, tid);
    const int wm = warp & 1, wn = warp >> 1;       // warp tile: rows [64 wm, +64), columns [24 wn, +24)
    const int fr = lane >> 2, fk = lane & 3;
    const int row = tid & (E_TM - 1), kq = tid >> 7; // Phi generation: this thread's vertex row and quarter of the stage's centres
    const double ox = (double)a.origin[0], oy = (double)a.origin[1], oz = (double)a.origin[2];
    const int Ktot = a.N + a.np;
    const int nstage = (Ktot + E_KB - 1) / E_KB;
    const int ncb = (3 * a.F + E_TN - 1) / E_TN;
    const int64_t n_vt = (a.V + E_TM - 1) / E_TM;
    const int64_t n_tiles = n_vt * ncb;

    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t vt = tile / ncb;
        const int cb = (int)(tile - vt * ncb);
        const int c0 = cb * E_TN;
        const int64_t v = vt * E_TM + row;
        float pos[3] = {0.f, 0.f, 0.f};
        if (v < a.V) {
            pos[0] = a.P[3 * v];
            pos[1] = a.P[3 * v + 1];
            pos[2] = a.P[3 * v + 2];
        }
        const double px = (double)pos[0], py = (double)pos[1], pz = (double)pos[2];
        const double qx = px - ox, qy = py - oy, qz = pz - oz;
        const double pp = qx * qx + qy * qy + qz * qz;

        double acc[8][3][2];
#pragma unroll
        for (int mi = 0; mi < 8; ++mi)
#pragma unroll
            for (int ni = 0; ni < 3; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;

        // weight tile of stage s -> buffer b (cp.async, rows beyond N + np and columns beyond ldw zero-filled)
        auto load_w = [&](int s, int b) {
            double* sB = s_stage + b * E_STAGE_DOUBLES + E_TM * E_LDA;
            for (int t = tid; t < E_KB * (E_TN / 2); t += E_THREADS) {
                const int k = t / (E_TN / 2), q = t - k * (E_TN / 2);
                const int gk = s * E_KB + k, gc = c0 + 2 * q;
                const bool ok = gk < Ktot && gc < a.ncol;
                cp_async16_zfill(sB + k * E_LDB + 2 * q, a.W + (size_t)(ok ? gk : 0) * a.ldw + (ok ? gc : 0), ok);
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        // centre table of stage s in the form the distance wants: t = q . (a, b, c) + d + pp * s
        auto load_c = [&](int s) {
            if (tid < E_KB) {
                const int j = s * E_KB + tid;
                double ca = 0.0, cb_ = 0.0, cc = 0.0, cd = 0.0, cs = 0.0;
                if (j < a.N) {
                    const double4 c = a.ctab[j];
                    if (KERNEL == FD_KERNEL_GAUSSIAN) {
                        const double sc = c.w * 1.4426950408889634074; // -log2(e) / R^2
                        const double cx = c.x - ox, cy = c.y - oy, cz = c.z - oz;
                        ca = -2.0 * sc * cx, cb_ = -2.0 * sc * cy, cc = -2.0 * sc * cz;
                        cd = sc * (cx * cx + cy * cy + cz * cz);
                        cs = sc;
                    } else {
                        ca = c.x, cb_ = c.y, cc = c.z, cd = c.w, cs = 1.0;
                    }
                }
                double* d = s_ctr + (s & 1) * E_KB * 5 + tid * 5;
                d[0] = ca, d[1] = cb_, d[2] = cc, d[3] = cd, d[4] = cs;
            }
        };
        // two of this thread's 8 basis values of stage s (pair jj / 2) -> A tile of buffer s & 1
        auto gen_pair = [&](int s, int jj) {
            double* sA = s_stage + (s & 1) * E_STAGE_DOUBLES + row * E_LDA + kq * 8;
            const double* sc = s_ctr + (s & 1) * E_KB * 5;
            double ph[2];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int kk = kq * 8 + jj + e;
                const int j = s * E_KB + kk;
                const double* c = sc + kk * 5; // warp-wide broadcast reads
                const double t = fma(qx, c[0], fma(qy, c[1], fma(qz, c[2], fma(pp, c[4], c[3]))));
                double val;
                if (KERNEL == FD_KERNEL_GAUSSIAN) val = fd_exp2_64(fmin(t, 0.0), s_exp);
                else if (KERNEL == FD_KERNEL_MULTIQUADRIC) val = fd_fast_sqrt64(t);
                else val = fmax(t, 0.0) * fd_half_log64(fmax(t, 0.0), s_log);
                if (j >= a.N) { // polynomial rows [1, x, y, z], then zero padding
                    const int r = j - a.N;
                    val = r >= a.np ? 0.0 : (r == 0 ? 1.0 : (r == 1 ? px : (r == 2 ? py : pz)));
                }
                ph[e] = val;
            }
            *reinterpret_cast<double2*>(sA + jj) = make_double2(ph[0], ph[1]);
        };

        // Software pipeline: while a warp contracts stage s it also produces its share of the Phi tile of stage s + 1, two
        // basis values per two K = 4 steps, in the same instruction stream -- the dependent DFMA chains of the basis functions
        // then hide behind the DMMAs instead of forming a phase of their own (that phase left the tensor pipe 51 % busy).
        __syncthreads(); // the previous tile's epilogue has left the stage buffers
        load_c(0);
        load_w(0, 0);
        __syncthreads();
#pragma unroll
        for (int jj = 0; jj < 8; jj += 2) gen_pair(0, jj);
        if (nstage > 1) load_c(1);
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        for (int s = 0; s < nstage; ++s) {
            const int b = s & 1;
            __syncthreads(); // stage s complete (Phi + weights), the centres of s + 1 are written; buffer b ^ 1 is free
            const bool more = s + 1 < nstage;
            if (more) load_w(s + 1, b ^ 1);
            if (s + 2 < nstage) load_c(s + 2); // into the centre buffer of stage s, whose Phi tile is complete
            const double* sA = s_stage + b * E_STAGE_DOUBLES + (wm * 64 + fr) * E_LDA + fk;
            const double* sB = s_stage + b * E_STAGE_DOUBLES + E_TM * E_LDA + fk * E_LDB + wn * 24 + fr;
#pragma unroll
            for (int k4 = 0; k4 < E_KB / 4; ++k4) {
                double af[8], bf[3];
#pragma unroll
                for (int mi = 0; mi < 8; ++mi) af[mi] = sA[mi * 8 * E_LDA + k4 * 4];
#pragma unroll
                for (int ni = 0; ni < 3; ++ni) bf[ni] = sB[k4 * 4 * E_LDB + ni * 8];
#pragma unroll
                for (int mi = 0; mi < 8; ++mi)
#pragma unroll
                    for (int ni = 0; ni < 3; ++ni) fd_dmma884(acc[mi][ni][0], acc[mi][ni][1], af[mi], bf[ni]);
                if (more && (k4 & 1)) gen_pair(s + 1, k4 - 1);
            }
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        __syncthreads(); // all warps are done with the stage buffers: they become the FP32 staging of the accumulators
#pragma unroll
        for (int mi = 0; mi < 8; ++mi)
#pragma unroll
            for (int ni = 0; ni < 3; ++ni) {
                float* d = s_C + (wm * 64 + mi * 8 + fr) * E_LDC + wn * 24 + ni * 8 + 2 * fk;
                d[0] = (float)acc[mi][ni][0]; // the narrowing of SOP_FaceDeform.cpp:415
                d[1] = (float)acc[mi][ni][1];
            }
        __syncthreads();
        // epilogue: thread = (vertex row, frames kq, kq + 4, ...)
        if (v < a.V) {
            const float d2 = a.dist2 ? a.dist2[v] : 0.f;
            const bool skip = d2 > a.radius2;                                   // :408-410
            float fo = powf(1.0f - fminf(d2 / a.radius2, 1.0f), a.falloffrate); // :423-424
            if (skip) fo = 0.f;
            if (a.falloff_out && cb == 0 && kq == 0) a.falloff_out[v] = fo;
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
            const int f0 = cb * (E_TN / 3);
            const int nf = min(E_TN / 3, a.F - f0);
            for (int fi = kq; fi < nf; fi += 4) {
                const float* src = s_C + row * E_LDC + 3 * fi;
                float d[3] = {src[0], src[1], src[2]};
                if (a.do_tangent) fd_project_to_tangents(tu, tv, tn, d);
                float* o = a.P_out + ((size_t)(f0 + fi) * (size_t)a.V + (size_t)v) * 3;
#pragma unroll
                for (int k = 0; k < 3; ++k) o[k] = skip ? pos[k] : pos[k] + d[k] * fo;
            }
        }
    }
}

template <int KERNEL> cudaError_t launch_eval64(fd_ctx* ctx, const Eval64Args& a)
{
    const int ncb = (3 * a.F + E_TN - 1) / E_TN;
    const int64_t n_tiles = ((a.V + E_TM - 1) / E_TM) * ncb;
    const int grid = (int)(n_tiles < ctx->sm_count ? n_tiles : ctx->sm_count);
    k_eval64_mma<KERNEL><<<grid, E_THREADS, E_SMEM_BYTES, ctx->stream>>>(a);
    ctx->launches += 1;
    return cudaGetLastError();
}

} // namespace

cudaError_t fd_eval64_setup(fd_ctx* ctx)
{
    (void)ctx;
    cudaError_t e = cudaFuncSetAttribute(k_eval64_mma<FD_KERNEL_GAUSSIAN>, cudaFuncAttributeMaxDynamicSharedMemorySize, E_SMEM_BYTES);
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(k_eval64_mma<FD_KERNEL_MULTIQUADRIC>, cudaFuncAttributeMaxDynamicSharedMemorySize, E_SMEM_BYTES);
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(k_eval64_mma<FD_KERNEL_THINPLATE>, cudaFuncAttributeMaxDynamicSharedMemorySize, E_SMEM_BYTES);
    return e;
}

// after a solve / commit: measure the cancellation and settle the evaluation kernel on the device
cudaError_t fd_launch_cancel_select(fd_ctx* ctx, fd_model* m, int tensor_ok, int simt_ok, int tcx_ok, int want)
{
    cudaStream_t s = ctx->stream;
    k_wmax<<<m->N, 128, 0, s>>>(fd_w_src(m), m->ldw, 3 * m->F, m->N, m->d_wmax);
    SelectArgs a;
    a.rest = m->d_rest;
    a.radii = m->d_radii;
    a.wmax = m->d_wmax;
    a.norm = m->d_tc_norm;
    a.N = m->N;
    a.kernel = m->prm.kernel;
    a.tensor_ok = tensor_ok;
    a.simt_ok = simt_ok;
    a.want = want;
    a.tcx_ok = tcx_ok;
    a.tol = m->prm.eval_tolerance;
    a.smax_bits = reinterpret_cast<unsigned long long*>(m->d_est + 2);
    a.done = reinterpret_cast<unsigned*>(m->d_est + 3);
    a.sel = m->d_sel;
    a.est = m->d_est;
    k_cancel_select<<<(m->N + 7) / 8, 256, 0, s>>>(a);
    ctx->launches += 2;
    return cudaGetLastError();
}

cudaError_t fd_launch_eval64_mma(fd_ctx* ctx, const fd_model* m, const float* P, int64_t V, const float* dist2, const float* tu,
                                 const float* tv, const float* nrm, float* P_out, float* falloff_out, const int* sel, int sel_id)
{
    if (V <= 0) return cudaSuccess;
    Eval64Args a;
    a.ctab = m->d_ctab64;
    a.origin = m->d_rest;
    a.W = m->d_W;
    a.ldw = m->ldw;
    a.ncol = m->ldw - m->w_col0;
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
    a.do_tangent = (m->prm.tangent && tu && tv && nrm) ? 1 : 0;
    a.sel = sel;
    a.sel_id = sel_id;
    switch (m->prm.kernel) {
    case FD_KERNEL_GAUSSIAN: return launch_eval64<FD_KERNEL_GAUSSIAN>(ctx, a);
    case FD_KERNEL_MULTIQUADRIC: return launch_eval64<FD_KERNEL_MULTIQUADRIC>(ctx, a);
    default: return launch_eval64<FD_KERNEL_THINPLATE>(ctx, a);
    }
}
