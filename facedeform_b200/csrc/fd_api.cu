// fd_api.cu -- the C ABI of libfacedeform_gpu.so (include/facedeform_gpu.h): handles, memory, phase timing,
// host-pointer wrappers around the device-pointer entry points.  No CPU fallback anywhere: every numeric
// result comes out of the CUDA kernels in this directory.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <new>

#include "fd_internal.h"

namespace {

#define FD_SET_ERR(ctx, ...) snprintf((ctx)->err, sizeof((ctx)->err), __VA_ARGS__)

using DeviceGuard = fd_device_guard;

int stage(fd_ctx* ctx, int slot, size_t bytes, void** out)
{
    if (bytes < 256) bytes = 256;
    if (ctx->stage_bytes[slot] < bytes) {
        if (ctx->stage_dev[slot]) {
            FD_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
            FD_CUDA_OK(ctx, cudaFree(ctx->stage_dev[slot]));
            ctx->stage_dev[slot] = nullptr;
            ctx->stage_bytes[slot] = 0;
        }
        cudaError_t e = cudaMalloc(&ctx->stage_dev[slot], bytes);
        if (e != cudaSuccess) {
            FD_SET_ERR(ctx, "cudaMalloc(%zu bytes) failed: %s", bytes, cudaGetErrorString(e));
            return e == cudaErrorMemoryAllocation ? FD_E_NOMEM : FD_E_CUDA;
        }
        ctx->stage_bytes[slot] = bytes;
        if (ctx->dbg.poison) cudaMemsetAsync(ctx->stage_dev[slot], 0xFF, bytes, ctx->stream);
    }
    *out = ctx->stage_dev[slot];
    return FD_OK;
}

void phase_begin(fd_ctx* ctx, int ph) { cudaEventRecord(ctx->ev_begin[ph], ctx->stream); }
void phase_end(fd_ctx* ctx, int ph)
{
    cudaEventRecord(ctx->ev_end[ph], ctx->stream);
    ctx->phase_valid[ph] = true;
}

template <typename T> int dev_alloc(fd_ctx* ctx, T** p, size_t count)
{
    // stream-ordered pool allocation: no device-wide synchronisation when models are created and destroyed per cook
    cudaError_t e = cudaMallocAsync((void**)p, (count ? count : 1) * sizeof(T), ctx->stream);
    if (e != cudaSuccess) {
        FD_SET_ERR(ctx, "cudaMallocAsync(%zu bytes) failed: %s", count * sizeof(T), cudaGetErrorString(e));
        *p = nullptr;
        return e == cudaErrorMemoryAllocation ? FD_E_NOMEM : FD_E_CUDA;
    }
    if (ctx->dbg.poison) cudaMemsetAsync(*p, 0xFF, (count ? count : 1) * sizeof(T), ctx->stream);
    return FD_OK;
}

// the development knobs of fd_debug_opts: the only place the library reads the environment
void read_debug_opts(fd_debug_opts* o)
{
    auto flag = [](const char* name) { return getenv(name) != nullptr; };
    auto num = [](const char* name) { const char* v = getenv(name); return v ? atoi(v) : 0; };
    memset(o, 0, sizeof(*o));
    o->no_nullspace = flag("FD_NO_NULLSPACE");
    o->force_pivoted_lu = flag("FD_FORCE_PIVOTED_LU");
    o->lu_unfused = flag("FD_LU_UNFUSED");
    o->solve_dfma = flag("FD_SOLVE_DFMA");
    o->no_fused_pack = flag("FD_NO_FUSED_PACK");
    o->no_few_rhs = flag("FD_NO_FEW_RHS");
    o->no_inverse = flag("FD_NO_INVERSE");
    o->eval_scalar_f32 = flag("FD_EVAL_SCALAR_F32");
    o->tc_nopair = flag("FD_TC_NOPAIR");
    o->has_tc_debug = flag("FD_TC_DEBUG");
    o->lu_sym_off = flag("FD_LU_NOSYM");
    o->poison = flag("FD_POISON");
    o->eval_vp = num("FD_EVAL_VP");
    o->tc_debug = num("FD_TC_DEBUG");
    o->lu_debug = num("FD_LU_DEBUG");
    o->lu_nbo = num("FD_LU_NBO");
    o->lu_cluster_max_n = num("FD_LU_CLUSTER_MAX_N");
    o->lu_cluster = num("FD_LU_CLUSTER");
    o->tcx_cbu = num("FD_TCX_CBU");
    o->tcx_narrow = num("FD_TCX_NARROW");
}

int check_params(fd_ctx* ctx, const fd_params* p)
{
    if (p->model != FD_MODEL_QNN && p->model != FD_MODEL_ML) { FD_SET_ERR(ctx, "model must be 0 (QNN) or 1 (Multilayer)"); return FD_E_INVALID; }
    if (p->term < 0 || p->term > 2) { FD_SET_ERR(ctx, "term must be 0 (linear), 1 (const) or 2 (zero)"); return FD_E_INVALID; }
    if (p->kernel < 0 || p->kernel > 2) { FD_SET_ERR(ctx, "kernel must be 0 (gaussian), 1 (multiquadric) or 2 (thin plate)"); return FD_E_INVALID; }
    if (!(p->radius > 0.f) || !isfinite(p->radius)) { FD_SET_ERR(ctx, "radius must be positive"); return FD_E_INVALID; }
    if (!(p->lambda >= 0.f)) { FD_SET_ERR(ctx, "lambda must be >= 0"); return FD_E_INVALID; }
    if (p->eval_precision < 0 || p->eval_precision > 2 || p->eval_path < 0 || p->eval_path > 2) { FD_SET_ERR(ctx, "bad eval_precision / eval_path"); return FD_E_INVALID; }
    if (p->factor_precision != FD_FACTOR_FP64 && p->factor_precision != FD_FACTOR_FP32_IR) { FD_SET_ERR(ctx, "bad factor_precision"); return FD_E_INVALID; }
    if (p->fidelity != FD_FIDELITY_DENSE && p->fidelity != FD_FIDELITY_ALGLIB_V1) { FD_SET_ERR(ctx, "bad fidelity"); return FD_E_INVALID; }
    return FD_OK;
}

// allocates everything that depends only on N (fit-time state)
int model_alloc(fd_ctx* ctx, const fd_params* params, int N, bool with_factor, fd_model** out)
{
    fd_model* m = new (std::nothrow) fd_model();
    if (!m) return FD_E_NOMEM;
    memset(m, 0, sizeof(*m));
    m->ctx = ctx;
    fd_ctx_retain(ctx);
    m->prm = *params;
    m->N = N;
    m->np = fd_poly_terms(params->term);
    m->n = N + m->np;
    m->lda = fd_round_up(m->n, 32);
    m->receiver = !with_factor;
    m->eval64 = params->eval_precision == FD_EVAL_FP64 ||
                (params->eval_precision == FD_EVAL_AUTO && params->kernel != FD_KERNEL_GAUSSIAN);
    // Gaussian under FD_EVAL_AUTO: FP32 while the measured cancellation allows it, FP64 beyond (fd_eval64.cu)
    m->auto_sel = params->eval_precision == FD_EVAL_AUTO && params->kernel == FD_KERNEL_GAUSSIAN;
    int st = FD_OK;
    if (st == FD_OK) st = dev_alloc(ctx, &m->d_rest, (size_t)N * 3);
    if (st == FD_OK) st = dev_alloc(ctx, &m->d_radii, (size_t)N);
    if (st == FD_OK) st = dev_alloc(ctx, &m->d_flags, FD_NUM_FLAGS);
    if (st == FD_OK) st = dev_alloc(ctx, &m->d_pivstat, 2);
    if (st == FD_OK) st = dev_alloc(ctx, &m->d_ctab32, (size_t)fd_tc_kpad(N)); // padded: the tensor path bulk-copies 32-centre tiles
    if (st == FD_OK) st = dev_alloc(ctx, &m->d_ctab_pair, (size_t)fd_tc_kpad(N));
    if (st == FD_OK && (m->eval64 || m->auto_sel)) st = dev_alloc(ctx, &m->d_ctab64, (size_t)fd_tc_kpad(N)); // padded like d_ctab32
    if (st == FD_OK && m->auto_sel) st = dev_alloc(ctx, &m->d_ctab_tcx, (size_t)fd_tc_kpad(N));
    if (st == FD_OK && m->auto_sel) st = dev_alloc(ctx, &m->d_csc_tcx, (size_t)fd_tc_kpad(N));
    if (st == FD_OK) st = dev_alloc(ctx, &m->d_tc_norm, 8);
    if (st == FD_OK) st = dev_alloc(ctx, &m->d_sel, 2);
    if (st == FD_OK) st = dev_alloc(ctx, &m->d_est, 4);
    if (st == FD_OK) st = dev_alloc(ctx, &m->d_wmax, (size_t)N);
    if (st == FD_OK && with_factor) {
        st = dev_alloc(ctx, &m->d_A, (size_t)m->lda * m->n);
        if (st == FD_OK) st = dev_alloc(ctx, &m->d_ipiv, (size_t)m->n);
        if (st == FD_OK) st = dev_alloc(ctx, &m->d_perm, (size_t)m->n);
        if (st == FD_OK) st = dev_alloc(ctx, &m->d_win, (size_t)132);
        if (st == FD_OK) st = dev_alloc(ctx, &m->d_Tinv, (size_t)((m->n + 31) / 32) * 2 * 32 * 32);
        m->f32ir = params->factor_precision == FD_FACTOR_FP32_IR;
        if (st == FD_OK && m->f32ir) st = dev_alloc(ctx, &m->d_A32, (size_t)m->lda * m->n);
        if (st == FD_OK && m->f32ir) st = dev_alloc(ctx, &m->d_ir_norm, 2);
    }
    if (st != FD_OK) {
        fd_model_destroy(m);
        return st;
    }
    cudaMemsetAsync(m->d_flags, 0, FD_NUM_FLAGS * sizeof(int), ctx->stream);
    cudaMemsetAsync(m->d_sel, 0, 2 * sizeof(int), ctx->stream);
    cudaMemsetAsync(m->d_est, 0, 4 * sizeof(double), ctx->stream);
    *out = m;
    return FD_OK;
}

// evaluation kernel choice for F frames (fd_params.eval_path): FP32 only; FD_PATH_AUTO needs a wide 3F
bool model_wants_tc(const fd_model* m, int F)
{
    if (m->eval64 || m->prm.eval_path == FD_PATH_SIMT || !m->d_tc_wt_hi) return false;
    return m->prm.eval_path == FD_PATH_TENSOR || 3 * F >= FD_TC_MIN_COLUMNS;
}

// Which evaluation kernel a batch of F frames takes.  Gaussian under FD_EVAL_AUTO with a wide batch: the exact-digit tensor-core
// kernel (fd_eval_tcx.cu), whose error does not grow with the cancellation of the weights -- a static choice.  Otherwise the
// FP32 tensor-core kernel when asked for / wide enough, and for FD_EVAL_AUTO the device-side choice among the FP32 kernels and
// FP64 (fd_eval64.cu).
void model_choose_eval(fd_model* m, int F)
{
    m->use_tcx = m->auto_sel && m->prm.eval_path != FD_PATH_SIMT && m->d_tcx_wt_mid && 3 * F >= FD_TC_MIN_COLUMNS;
    m->use_tc = !m->use_tcx && model_wants_tc(m, F);
}

int model_reserve_frames(fd_model* m, int F)
{
    fd_ctx* ctx = m->ctx;
    if (F <= m->capF) return FD_OK;
    void* old[] = {m->d_W, m->d_W32, m->d_tc_scale, m->d_tc_unscale, m->d_tc_wt_hi, m->d_tc_wt_lo, m->d_tcx_wt_mid, m->d_B, m->d_R, m->d_D32};
    for (void* b : old)
        if (b) cudaFreeAsync(b, ctx->stream);
    m->d_W = nullptr; m->d_W32 = nullptr; m->d_tc_scale = nullptr; m->d_tc_unscale = nullptr;
    m->d_tc_wt_hi = nullptr; m->d_tc_wt_lo = nullptr; m->d_tcx_wt_mid = nullptr; m->d_B = nullptr; m->d_R = nullptr; m->d_D32 = nullptr;
    m->capF = 0;
    const int ld = fd_round_up(3 * F, 4);
    int st = dev_alloc(ctx, &m->d_W, (size_t)m->n * ld);
    if (st == FD_OK) st = dev_alloc(ctx, &m->d_W32, (size_t)m->n * ld);
    if (st == FD_OK && m->f32ir) st = dev_alloc(ctx, &m->d_B, (size_t)m->n * ld);
    if (st == FD_OK && m->f32ir) st = dev_alloc(ctx, &m->d_R, (size_t)m->n * ld);
    if (st == FD_OK && m->f32ir) st = dev_alloc(ctx, &m->d_D32, (size_t)m->n * ld);
    if (st == FD_OK && !m->eval64 && m->prm.eval_path != FD_PATH_SIMT) { // tensor-path tables (used when 3F is wide enough)
        const size_t cols = (size_t)fd_tc_col_pad(F), kpad = (size_t)fd_tc_kpad(m->N);
        unsigned short *hi = nullptr, *lo = nullptr;
        st = dev_alloc(ctx, &m->d_tc_scale, cols);
        if (st == FD_OK) st = dev_alloc(ctx, &m->d_tc_unscale, cols);
        if (st == FD_OK) st = dev_alloc(ctx, &hi, cols * kpad);
        if (st == FD_OK) st = dev_alloc(ctx, &lo, cols * kpad);
        m->d_tc_wt_hi = hi;
        m->d_tc_wt_lo = lo;
        if (st == FD_OK && m->auto_sel) { // the third table of the exact-digit kernel (its 120-column blocks fit the same padding)
            unsigned short* mid = nullptr;
            st = dev_alloc(ctx, &mid, cols * kpad);
            m->d_tcx_wt_mid = mid;
            if (st == FD_OK && !m->d_tcx_rowexp) st = dev_alloc(ctx, &m->d_tcx_rowexp, kpad + 4);
            if (st == FD_OK && !m->d_tcx_rowmax) st = dev_alloc(ctx, &m->d_tcx_rowmax, kpad + 4);
            if (st == FD_OK && !m->d_tcx_ctab_eff) st = dev_alloc(ctx, &m->d_tcx_ctab_eff, kpad);
            if (st == FD_OK && !m->d_tcx_pw) st = dev_alloc(ctx, &m->d_tcx_pw, kpad);
            if (st == FD_OK && !m->d_tcx_meta) {
                st = dev_alloc(ctx, &m->d_tcx_meta, 2);
                if (st == FD_OK) cudaMemsetAsync(m->d_tcx_meta, 0, 2 * sizeof(double), ctx->stream);
            }
        }
    }
    if (st != FD_OK) return st;
    m->capF = F;
    return FD_OK;
}

// ---- FD_FIDELITY_ALGLIB_V1: two-stage polynomial + Gaussian layers (oracle: fdo_fit_v1) ---------------
// The parent model owns one ordinary factored sub-model per layer (term = zero: no side conditions) and, after a
// solve, a stacked evaluation model with N * layers centres that the ordinary evaluation kernels take.
int v1_fit(fd_ctx* ctx, const fd_params* params, const float* rest_dev, int32_t N, fd_model** out)
{
    if (params->kernel != FD_KERNEL_GAUSSIAN) { FD_SET_ERR(ctx, "the ALGLIB v1 formulation is Gaussian only"); return FD_E_UNSUPPORTED; }
    if (params->factor_precision != FD_FACTOR_FP64) { FD_SET_ERR(ctx, "the ALGLIB v1 formulation factors in FP64"); return FD_E_UNSUPPORTED; }
    const int L = params->model == FD_MODEL_QNN ? 1 : (params->layers < 1 ? 1 : params->layers);
    if (L > FD_V1_MAX_LAYERS) { FD_SET_ERR(ctx, "at most %d layers", FD_V1_MAX_LAYERS); return FD_E_INVALID; }
    fd_model* m = nullptr;
    int st = model_alloc(ctx, params, N, false, &m);
    if (st != FD_OK) return st;
    m->receiver = false;
    cudaError_t e = cudaMemcpyAsync(m->d_rest, rest_dev, (size_t)N * 3 * sizeof(float), cudaMemcpyDefault, ctx->stream);
    if (e != cudaSuccess) { FD_SET_ERR(ctx, "fit: %s", cudaGetErrorString(e)); fd_model_destroy(m); return FD_E_CUDA; }
    for (int k = 0; k < L; ++k) {
        fd_params pk = *params;
        pk.fidelity = FD_FIDELITY_DENSE;
        pk.term = FD_TERM_ZERO;
        pk.layers = 1;
        pk.eval_path = FD_PATH_SIMT; // the layers are never evaluated on their own
        if (params->model != FD_MODEL_QNN) pk.radius = params->radius / (float)(1 << k); // R, R/2, R/4 ...
        st = fd_rbf_fit_dev(ctx, &pk, m->d_rest, N, &m->v1_layer[k]);
        if (st != FD_OK) { fd_model_destroy(m); return st; }
        m->v1_layers = k + 1;
    }
    m->fitted = true;
    *out = m;
    return FD_OK;
}

int v1_solve(fd_model* m, const float* deform_dev, int32_t F)
{
    fd_ctx* ctx = m->ctx;
    const int N = m->N, np = m->np, L = m->v1_layers, nrhs = 3 * F, ldw = fd_round_up(3 * F, 4);
    cudaStream_t s = ctx->stream;
    int st = FD_OK;
    if (F > m->capF) {
        if (m->d_v1_R) cudaFreeAsync(m->d_v1_R, s);
        if (m->d_v1_V) cudaFreeAsync(m->d_v1_V, s);
        m->d_v1_R = m->d_v1_V = nullptr;
        st = dev_alloc(ctx, &m->d_v1_R, (size_t)N * ldw);
        if (st == FD_OK) st = dev_alloc(ctx, &m->d_v1_V, (size_t)4 * ldw);
        if (st != FD_OK) return st;
        m->capF = F;
    }
    const int lda = m->v1_layer[0]->lda;
    if (!m->d_v1_K && (st = dev_alloc(ctx, &m->d_v1_K, (size_t)lda * N)) != FD_OK) return st;
    if (!m->d_v1_stack && (st = dev_alloc(ctx, &m->d_v1_stack, (size_t)N * L * 3)) != FD_OK) return st;
    m->F = F;
    m->ldw = ldw;
    phase_begin(ctx, FD_PH_SOLVE);
    cudaError_t e = fd_launch_v1_rhs_poly(ctx, m->d_rest, deform_dev, N, F, np, m->d_v1_R, m->d_v1_V, ldw, m->d_flags);
    for (int k = 0; k < L && e == cudaSuccess; ++k) {
        fd_model* l = m->v1_layer[k];
        st = model_reserve_frames(l, F);
        if (st != FD_OK) return st;
        l->F = F;
        l->ldw = l->ldw32 = ldw;
        l->use_tc = false;
        e = fd_launch_v1_gather(ctx, m->d_v1_R, l->d_perm, N, ldw, l->d_W);
        if (e == cudaSuccess) e = fd_launch_solve_prebuilt(ctx, l, nrhs);
        fd_params p0 = l->prm;
        p0.lambda = 0.f; // the fitted function uses the kernel matrix itself, the shift only damps the solve
        if (e == cudaSuccess) e = fd_launch_assemble(ctx, p0, l->d_rest, l->d_radii, N, 0, m->d_v1_K, lda);
        if (e == cudaSuccess) e = fd_launch_gemm_sub(ctx, m->d_v1_K, lda, N, N, l->d_W, m->d_v1_R, ldw, nrhs);
    }
    if (e != cudaSuccess) { FD_SET_ERR(ctx, "solve: %s", cudaGetErrorString(e)); return FD_E_CUDA; }
    // the stacked model the evaluation kernels take: layer k's centres, radii and weights at rows [k N, (k + 1) N)
    if (m->v1_eval) { fd_model_destroy(m->v1_eval); m->v1_eval = nullptr; }
    for (int k = 0; k < L; ++k)
        cudaMemcpyAsync(m->d_v1_stack + (size_t)k * N * 3, m->d_rest, (size_t)N * 3 * sizeof(float), cudaMemcpyDeviceToDevice, s);
    fd_params pe = m->prm;
    pe.fidelity = FD_FIDELITY_DENSE;
    st = fd_model_create_receiver(ctx, &pe, m->d_v1_stack, N * L, F, &m->v1_eval);
    if (st != FD_OK) return st;
    fd_model* ev = m->v1_eval;
    for (int k = 0; k < L; ++k) {
        fd_model* l = m->v1_layer[k];
        cudaMemcpyAsync(ev->d_radii + (size_t)k * N, l->d_radii, (size_t)N * sizeof(double), cudaMemcpyDeviceToDevice, s);
        cudaMemcpyAsync(ev->d_W + (size_t)k * N * ldw, l->d_W, (size_t)N * ldw * sizeof(double), cudaMemcpyDeviceToDevice, s);
    }
    if (np > 0)
        cudaMemcpyAsync(ev->d_W + (size_t)N * L * ldw, m->d_v1_V, (size_t)np * ldw * sizeof(double), cudaMemcpyDeviceToDevice, s);
    st = fd_model_commit_weights(ev);
    phase_end(ctx, FD_PH_SOLVE);
    if (st != FD_OK) return st;
    m->solved = true;
    return FD_OK;
}

} // namespace

extern "C" {

int fd_abi_version(void) { return FD_ABI_VERSION; }

const char* fd_status_string(int status)
{
    switch (status) {
    case FD_OK: return "ok";
    case FD_E_INVALID: return "invalid argument";
    case FD_E_MISMATCH_POINT: return "Rest and deform geometry should match.";
    case FD_E_BUILD: return "Can't build RBF model.";
    case FD_E_SINGULAR: return "Can't solve the problem.";
    case FD_E_CUDA: return "CUDA error";
    case FD_E_NOMEM: return "out of device memory";
    case FD_E_CAPTURE: return "Can't capture geometry with a rig!";
    case FD_E_UNSUPPORTED: return "unsupported";
    case FD_E_STATE: return "call order";
    default: return "unknown status";
    }
}

// defaults of the parameter templates, SOP_FaceDeform.cpp:117-137
void fd_params_default(fd_params* p)
{
    memset(p, 0, sizeof(*p));
    p->model = FD_MODEL_QNN;
    p->term = FD_TERM_LINEAR;
    p->kernel = FD_KERNEL_GAUSSIAN;
    p->qcoef = 1.0f;
    p->zcoef = 5.0f;
    p->radius = 1.0f;
    p->layers = 4;
    p->lambda = 0.1f;
    p->maxedges = 4;
    p->weightrange[0] = 0.0f;
    p->weightrange[1] = 1.0f;
    p->falloffradius = 1.0f;
    p->falloffrate = 1.0f;
    p->eval_precision = FD_EVAL_AUTO;
    p->eval_path = FD_PATH_AUTO;
    p->factor_precision = FD_FACTOR_FP64;
    p->fidelity = FD_FIDELITY_DENSE;
    p->strict_reference = 0;
    p->eval_tolerance = 1e-5f;
    p->group[0] = 0;
}

// the fields a fit depends on (everything else is consumed by the evaluation epilogue, the capture or the host mirror)
int fd_params_fit_equal(const fd_params* a, const fd_params* b)
{
    if (!a || !b) return 0;
    if (a->model != b->model || a->term != b->term || a->kernel != b->kernel || a->lambda != b->lambda ||
        a->eval_precision != b->eval_precision || a->eval_path != b->eval_path || a->factor_precision != b->factor_precision ||
        a->fidelity != b->fidelity || a->eval_tolerance != b->eval_tolerance)
        return 0;
    if (a->model == FD_MODEL_QNN) {
        if (a->qcoef != b->qcoef || a->zcoef != b->zcoef) return 0; // `radius` is only the capture / falloff radius here
    } else if (a->radius != b->radius) {
        return 0;
    }
    if (a->fidelity == FD_FIDELITY_ALGLIB_V1 && a->model == FD_MODEL_ML && a->layers != b->layers) return 0;
    return 1;
}

// SYSmax clamps of cookMySop, SOP_FaceDeform.cpp:249-257
void fd_params_clamp(fd_params* p)
{
    if (p->qcoef < 0.1f) p->qcoef = 0.1f;
    if (p->zcoef < 0.1f) p->zcoef = 0.1f;
    if (p->radius < 0.01f) p->radius = 0.01f;
    if (p->layers < 1) p->layers = 1;
    if (p->lambda < 0.01f) p->lambda = 0.01f;
    if (p->maxedges < 1) p->maxedges = 1;
}

int fd_ctx_create(fd_ctx** out, int device, void* stream)
{
    if (!out) return FD_E_INVALID;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) return FD_E_CUDA; // no GPU: fail loudly, no fallback
    if (device < 0 && cudaGetDevice(&device) != cudaSuccess) return FD_E_CUDA;
    if (device >= count) return FD_E_INVALID;
    DeviceGuard g(device); // the caller's current device is restored on every return path
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return FD_E_CUDA;
    if (prop.major != 10) return FD_E_UNSUPPORTED; // built for sm_100a only
    fd_ctx* ctx = new (std::nothrow) fd_ctx();
    if (!ctx) return FD_E_NOMEM;
    memset(ctx, 0, sizeof(*ctx));
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    read_debug_opts(&ctx->dbg);
    if (stream) {
        ctx->stream = (cudaStream_t)stream;
    } else {
        if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return FD_E_CUDA; }
        ctx->own_stream = true;
    }
    for (int i = 0; i < FD_PH_COUNT; ++i) {
        cudaEventCreate(&ctx->ev_begin[i]);
        cudaEventCreate(&ctx->ev_end[i]);
    }
    if (cudaMalloc(&ctx->d_sync, 256) != cudaSuccess || cudaMemset(ctx->d_sync, 0, 256) != cudaSuccess) {
        fd_ctx_destroy(ctx);
        return FD_E_CUDA;
    }
    // function attributes are per device: every ctx sets them for its own (one process may hold one ctx per GPU)
    if (fd_solve_setup(ctx) != cudaSuccess || fd_factor_setup(ctx) != cudaSuccess || fd_eval_tc_setup(ctx) != cudaSuccess ||
        fd_eval64_setup(ctx) != cudaSuccess || fd_eval_tcx_setup(ctx) != cudaSuccess) {
        fd_ctx_destroy(ctx);
        return FD_E_CUDA;
    }
    if (ctx->dbg.has_tc_debug && (cudaMalloc(&ctx->d_tc_dbg, 256 * sizeof(long long)) != cudaSuccess ||
                                  cudaMemset(ctx->d_tc_dbg, 0, 256 * sizeof(long long)) != cudaSuccess))
        ctx->d_tc_dbg = nullptr;
    if (ctx->dbg.lu_debug && cudaMalloc(&ctx->d_lu_dbg, 64) != cudaSuccess) ctx->d_lu_dbg = nullptr;
    { // keep freed blocks cached in the device's default pool instead of returning them to the driver
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
            unsigned long long keep = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
    }
    *out = ctx;
    return FD_OK;
}

static void ctx_teardown(fd_ctx* ctx);

void fd_ctx_destroy(fd_ctx* ctx)
{
    if (!ctx) return;
    if (__atomic_load_n(&ctx->refs, __ATOMIC_ACQUIRE) > 0) { // live models / dbse handles: they finish the teardown (fd_ctx_release)
        ctx->destroy_requested = true;
        return;
    }
    ctx_teardown(ctx);
}

static void ctx_teardown(fd_ctx* ctx)
{
    DeviceGuard g(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (int i = 0; i < FD_NUM_STAGE; ++i)
        if (ctx->stage_dev[i]) cudaFree(ctx->stage_dev[i]);
    if (ctx->d_sync) cudaFree(ctx->d_sync);
    if (ctx->copy_stream) {
        cudaStreamSynchronize(ctx->copy_stream);
        cudaStreamDestroy(ctx->copy_stream);
        for (int i = 0; i < 8; ++i) cudaEventDestroy(ctx->ev_block[i]);
        cudaEventDestroy(ctx->ev_copied);
    }
    if (ctx->d_tc_dbg) cudaFree(ctx->d_tc_dbg);
    if (ctx->d_lu_dbg) cudaFree(ctx->d_lu_dbg);
    for (int i = 0; i < FD_PH_COUNT; ++i) {
        cudaEventDestroy(ctx->ev_begin[i]);
        cudaEventDestroy(ctx->ev_end[i]);
    }
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

int fd_ctx_synchronize(fd_ctx* ctx)
{
    if (!ctx) return FD_E_INVALID;
    FD_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
    return FD_OK;
}

const char* fd_last_error(const fd_ctx* ctx) { return ctx ? ctx->err : "null ctx"; }

float fd_ctx_phase_ms(fd_ctx* ctx, int phase)
{
    if (!ctx || phase < 0 || phase >= FD_PH_COUNT || !ctx->phase_valid[phase]) return -1.f;
    float ms = -1.f;
    if (cudaEventSynchronize(ctx->ev_end[phase]) != cudaSuccess) return -1.f;
    if (cudaEventElapsedTime(&ms, ctx->ev_begin[phase], ctx->ev_end[phase]) != cudaSuccess) return -1.f;
    return ms;
}

int64_t fd_ctx_launch_count(const fd_ctx* ctx) { return ctx ? ctx->launches : 0; }

void fd_model_destroy(fd_model* m)
{
    if (!m) return;
    fd_ctx* owner = m->ctx;
    for (int k = 0; k < m->v1_layers; ++k) fd_model_destroy(m->v1_layer[k]);
    if (m->v1_eval) fd_model_destroy(m->v1_eval);
    {
    DeviceGuard g(m->ctx->device);
    cudaStream_t s = m->ctx->stream; // stream-ordered frees: later work on the stream may reuse the blocks safely
    void* blocks[] = {m->d_rest, m->d_radii, m->d_A, m->d_ipiv, m->d_perm, m->d_W, m->d_flags, m->d_pivstat,
                      m->d_ctab32, m->d_W32, m->d_ctab64, m->d_tc_norm, m->d_tc_scale, m->d_tc_unscale,
                      m->d_tc_wt_hi, m->d_tc_wt_lo, m->d_Tinv, m->d_win, m->d_ctab_pair, m->d_A32, m->d_B, m->d_R, m->d_D32,
                      m->d_ir_norm, m->d_v1_R, m->d_v1_K, m->d_v1_V, m->d_v1_stack, m->d_ns, m->d_sel, m->d_est, m->d_wmax, m->d_inv, m->d_inv_rhs,
                      m->d_tcx_wt_mid, m->d_tcx_rowexp, m->d_tcx_rowmax, m->d_tcx_meta, m->d_ctab_tcx, m->d_csc_tcx,
                      m->d_tcx_ctab_eff, m->d_tcx_pw};
    for (void* b : blocks)
        if (b) cudaFreeAsync(b, s);
    delete m;
    }
    fd_ctx_release(owner);
}

// rbfcreate + rbfsetpoints + rbfsetalgo* + rbfset*term + the factorisation half of rbfbuildmodel
int fd_rbf_fit_dev(fd_ctx* ctx, const fd_params* params, const float* rest_ctrl_dev, int32_t n_ctrl, fd_model** out)
{
    if (!ctx || !params || !out) return FD_E_INVALID;
    *out = nullptr;
    DeviceGuard g(ctx->device);
    int st = check_params(ctx, params);
    if (st != FD_OK) return st;
    if (n_ctrl < 1 || !rest_ctrl_dev) { FD_SET_ERR(ctx, "Can't build RBF model: no control points"); return FD_E_BUILD; }
    if (params->fidelity == FD_FIDELITY_ALGLIB_V1) return v1_fit(ctx, params, rest_ctrl_dev, n_ctrl, out);
    if ((size_t)(n_ctrl + 4) * sizeof(int) > 64 * 1024) { FD_SET_ERR(ctx, "more than 16380 control points are not supported"); return FD_E_UNSUPPORTED; }
    fd_model* m = nullptr;
    st = model_alloc(ctx, params, n_ctrl, true, &m);
    if (st != FD_OK) return st;
    void* scratch;
    st = stage(ctx, FD_STAGE_MISC, 256, &scratch);
    if (st != FD_OK) { fd_model_destroy(m); return st; }
    cudaError_t e = cudaMemcpyAsync(m->d_rest, rest_ctrl_dev, (size_t)n_ctrl * 3 * sizeof(float), cudaMemcpyDefault, ctx->stream);
    phase_begin(ctx, FD_PH_ASSEMBLE);
    if (e == cudaSuccess) e = fd_launch_radii(ctx, m->prm, m->d_rest, m->N, m->d_radii, m->d_flags);
    // Multiquadric / thin plate with a uniform radius and the linear term: the null-space transform makes the system
    // definite, so it takes the fused no-pivot LU instead of the pivoted one (fd_nullspace.cu).  The multiquadric's
    // reduced matrix is NEGATIVE definite and its smoothing shift is -lambda (k_assemble), which keeps it so.
    m->ns = m->prm.kernel != FD_KERNEL_GAUSSIAN && m->prm.model == FD_MODEL_ML && m->np == 4 && m->N >= 8 &&
            m->prm.factor_precision == FD_FACTOR_FP64 && !ctx->dbg.no_nullspace && !ctx->dbg.force_pivoted_lu;
    if (m->ns && dev_alloc(ctx, &m->d_ns, (size_t)m->N * 5 + 32) != FD_OK) m->ns = false;
    if (e == cudaSuccess)
        e = fd_launch_assemble(ctx, m->prm, m->d_rest, m->d_radii, m->N, m->ns ? 0 : m->np, m->d_A, m->lda);
    if (e == cudaSuccess && m->ns) e = fd_launch_ns_transform(ctx, m);
    phase_end(ctx, FD_PH_ASSEMBLE);
    phase_begin(ctx, FD_PH_FACTOR);
    // Gaussian kernel with one radius: K + lambda I is symmetric positive definite -> no pivot search needed
    const bool spd = m->prm.kernel == FD_KERNEL_GAUSSIAN && (m->prm.model == FD_MODEL_ML || m->N == 1) && !ctx->dbg.force_pivoted_lu;
    const bool unfused = ctx->dbg.lu_unfused; // per-block-column launches (kept for comparison)
    if (e == cudaSuccess && m->ns) {
        // the definite (N - 4) x (N - 4) block of Q^T K Q, in place
        e = fd_launch_lu_nopivot_fused(ctx, m->d_A + (size_t)4 * m->lda + 4, m->lda, m->N - 4, m->d_ipiv, m->d_perm, m->d_flags,
                                       m->d_pivstat, m->d_Tinv, 1);
    } else if (e == cudaSuccess && m->f32ir) {
        // FP32 factorisation of fl32(A); d_A keeps the FP64 system for the residuals of the refinement (fd_refine.cu)
        e = fd_launch_to_f32(ctx, m->d_A, m->d_A32, (size_t)m->lda * m->n);
        if (e == cudaSuccess)
            e = spd ? fd_launch_lu_nopivot_f32(ctx, m->d_A32, m->lda, m->n, m->d_ipiv, m->d_perm, m->d_flags, m->d_pivstat)
                    : fd_launch_lu_f32(ctx, m->d_A32, m->lda, m->n, m->d_ipiv, m->d_perm, m->d_flags, m->d_pivstat, m->d_win);
    } else if (e == cudaSuccess) {
        if (spd && !unfused) {
            e = fd_launch_lu_nopivot_fused(ctx, m->d_A, m->lda, m->n, m->d_ipiv, m->d_perm, m->d_flags, m->d_pivstat, m->d_Tinv, 1);
        } else {
            e = spd ? fd_launch_lu_nopivot(ctx, m->d_A, m->lda, m->n, m->d_ipiv, m->d_perm, m->d_flags, m->d_pivstat)
                    : fd_launch_lu(ctx, m->d_A, m->lda, m->n, m->d_ipiv, m->d_perm, m->d_flags, m->d_pivstat, m->d_win);
            if (e == cudaSuccess) e = fd_launch_invdiag(ctx, m);
        }
    }
    phase_end(ctx, FD_PH_FACTOR);
    if (e == cudaSuccess) e = fd_launch_pack_tables(ctx, m); // centre tables: once per fit, not per solve
    if (e != cudaSuccess) {
        FD_SET_ERR(ctx, "fit: %s", cudaGetErrorString(e));
        fd_model_destroy(m);
        return FD_E_CUDA;
    }
    m->fitted = true;
    *out = m;
    return FD_OK;
}

int fd_rbf_fit(fd_ctx* ctx, const fd_params* params, const float* rest_ctrl, int32_t n_ctrl, fd_model** out,
               fd_report* report)
{
    // rest_ctrl is a host pointer: cudaMemcpyDefault inside fit_dev copies it (pageable memory is staged by the driver)
    int st = fd_rbf_fit_dev(ctx, params, rest_ctrl, n_ctrl, out);
    if (st != FD_OK) return st;
    st = fd_model_report(*out, report);
    if (st != FD_OK) {
        fd_model_destroy(*out);
        *out = nullptr;
    }
    return st;
}

int fd_rbf_solve_dev(fd_model* m, const float* deform_ctrl_dev, int32_t n_ctrl, int32_t frames)
{
    if (!m || !deform_ctrl_dev) return FD_E_INVALID;
    fd_ctx* ctx = m->ctx;
    DeviceGuard g(ctx->device);
    if (m->receiver || !m->fitted) { FD_SET_ERR(ctx, "solve: the model holds no factorisation"); return FD_E_STATE; }
    if (n_ctrl != m->N) { FD_SET_ERR(ctx, "%s", fd_status_string(FD_E_MISMATCH_POINT)); return FD_E_MISMATCH_POINT; } // :231-234
    if (frames < 1) { FD_SET_ERR(ctx, "frames must be >= 1"); return FD_E_INVALID; }
    if (m->v1_layers) return v1_solve(m, deform_ctrl_dev, frames);
    int st = model_reserve_frames(m, frames);
    if (st != FD_OK) return st;
    m->F = frames;
    m->ldw = fd_round_up(3 * frames, 4);
    m->ldw32 = m->ldw;
    model_choose_eval(m, frames);
    phase_begin(ctx, FD_PH_SOLVE);
    // FD_FLAG_NONFINITE describes the weights of THIS solve (a NaN in one frame's rig must not poison later cooks of a
    // cached model); singular / zero-radius are properties of the fit and stay
    cudaError_t e = cudaMemsetAsync(m->d_flags + FD_FLAG_NONFINITE, 0, 2 * sizeof(int), ctx->stream);
    if (e != cudaSuccess) { FD_SET_ERR(ctx, "solve: %s", cudaGetErrorString(e)); return FD_E_CUDA; }
    if (m->ns) { // D' = Q^T D, z = S^-1 D'[4:], a = R^-1 (D'[:4] - K'[:4, 4:] z), w = Q [0; z]
        m->tc_packed_by_solve = false;
        e = fd_launch_ns_rhs(ctx, m, deform_ctrl_dev, frames);
        cudaError_t e_inv = cudaSuccess;
        if (e == cudaSuccess && fd_try_inverse_solve(ctx, m, m->d_A + (size_t)4 * m->lda + 4, m->lda, m->N - 4, m->d_perm, m->d_Tinv,
                                                     nullptr, frames, m->d_W + (size_t)4 * m->ldw, m->ldw, true, &e_inv))
            e = e_inv; // per-cook fast path: one pass over the explicit inverse of the reduced block
        else if (e == cudaSuccess)
            e = fd_launch_solve_sub(ctx, m->d_A + (size_t)4 * m->lda + 4, m->lda, m->N - 4, m->d_perm, m->d_Tinv,
                                    m->d_W + (size_t)4 * m->ldw, m->ldw, 3 * frames);
        if (e == cudaSuccess) e = fd_launch_ns_finish(ctx, m, frames);
    } else {
        e = m->f32ir ? fd_refine_solve(ctx, m, deform_ctrl_dev, frames) : fd_launch_solve(ctx, m, deform_ctrl_dev, frames);
    }
    if (e == cudaSuccess) e = fd_launch_pack(ctx, m);
    phase_end(ctx, FD_PH_SOLVE);
    if (e != cudaSuccess) { FD_SET_ERR(ctx, "solve: %s", cudaGetErrorString(e)); return FD_E_CUDA; }
    m->solved = true;
    return FD_OK;
}

int fd_rbf_solve(fd_model* m, const float* deform_ctrl, int32_t n_ctrl, int32_t frames, fd_report* report)
{
    if (!m || !deform_ctrl) return FD_E_INVALID;
    fd_ctx* ctx = m->ctx;
    DeviceGuard g(ctx->device);
    if (n_ctrl != m->N) { FD_SET_ERR(ctx, "%s", fd_status_string(FD_E_MISMATCH_POINT)); return FD_E_MISMATCH_POINT; }
    if (frames < 1) { FD_SET_ERR(ctx, "frames must be >= 1"); return FD_E_INVALID; }
    const size_t bytes = (size_t)frames * n_ctrl * 3 * sizeof(float);
    void* d_def;
    int st = stage(ctx, FD_STAGE_P, bytes, &d_def);
    if (st != FD_OK) return st;
    FD_CUDA_OK(ctx, cudaMemcpyAsync(d_def, deform_ctrl, bytes, cudaMemcpyHostToDevice, ctx->stream));
    st = fd_rbf_solve_dev(m, (const float*)d_def, n_ctrl, frames);
    if (st != FD_OK) return st;
    return fd_model_report(m, report);
}

int fd_model_set_epilogue(fd_model* m, const fd_params* p)
{
    if (!m || !p) return FD_E_INVALID;
    if (!fd_params_fit_equal(&m->prm, p)) { FD_SET_ERR(m->ctx, "set_epilogue: the parameters differ in a field the fit depends on"); return FD_E_INVALID; }
    auto apply = [&](fd_model* t) {
        t->prm.tangent = p->tangent;
        t->prm.dofalloff = p->dofalloff;
        t->prm.falloffrate = p->falloffrate;
        t->prm.falloffradius = p->falloffradius;
        t->prm.maxedges = p->maxedges;
        t->prm.morphspace = p->morphspace;
        t->prm.doclampweight = p->doclampweight;
        t->prm.weightrange[0] = p->weightrange[0];
        t->prm.weightrange[1] = p->weightrange[1];
        t->prm.strict_reference = p->strict_reference;
        memcpy(t->prm.group, p->group, sizeof(t->prm.group));
        t->prm.radius = p->radius; // equal already unless model = QNN, where it is the capture / falloff radius only
    };
    apply(m);
    if (m->v1_eval) apply(m->v1_eval);
    return FD_OK;
}

int fd_model_report(fd_model* m, fd_report* report)
{
    if (!m) return FD_E_INVALID;
    fd_ctx* ctx = m->ctx;
    DeviceGuard g(ctx->device);
    if (m->v1_layers) { // layered fit: every layer's factorisation, then the stacked weights, must be sound
        for (int k = 0; k < m->v1_layers; ++k) {
            const int st = fd_model_report(m->v1_layer[k], report);
            if (st != FD_OK) return st;
        }
        fd_report last;
        if (report) last = *report;
        if (m->v1_eval) {
            const int st = fd_model_report(m->v1_eval, report);
            if (st != FD_OK) return st;
        }
        // the parent's own flag: a singular Gram matrix of the polynomial fit
        int flags[FD_NUM_FLAGS];
        FD_CUDA_OK(ctx, cudaMemcpyAsync(flags, m->d_flags, sizeof(flags), cudaMemcpyDeviceToHost, ctx->stream));
        FD_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
        const int term = flags[FD_FLAG_SINGULAR] ? -3 : 1;
        if (report) {
            const double canc = report->cancellation; // of the stacked evaluation model
            const int ek = report->eval_kernel, inexact = report->eval_inexact;
            *report = last; // pivots of the last layer
            if (m->v1_eval) { report->cancellation = canc; report->eval_kernel = ek; report->eval_inexact = inexact; }
            report->terminationtype = term;
            report->n = m->N * m->v1_layers;
            report->npoly = m->np;
            report->frames = m->F;
        }
        if (term != 1) {
            FD_SET_ERR(ctx, "%s (terminationtype %d)", fd_status_string(FD_E_SINGULAR), term);
            return FD_E_SINGULAR;
        }
        return FD_OK;
    }
    int flags[FD_NUM_FLAGS];
    double piv[2] = {0, 0}, est[2] = {0, 0};
    int sel = 0;
    FD_CUDA_OK(ctx, cudaMemcpyAsync(flags, m->d_flags, sizeof(flags), cudaMemcpyDeviceToHost, ctx->stream));
    FD_CUDA_OK(ctx, cudaMemcpyAsync(piv, m->d_pivstat, sizeof(piv), cudaMemcpyDeviceToHost, ctx->stream));
    FD_CUDA_OK(ctx, cudaMemcpyAsync(est, m->d_est, sizeof(est), cudaMemcpyDeviceToHost, ctx->stream));
    FD_CUDA_OK(ctx, cudaMemcpyAsync(&sel, m->d_sel, sizeof(sel), cudaMemcpyDeviceToHost, ctx->stream));
    FD_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
    int term = 1;
    if (flags[FD_FLAG_ZERO_RADIUS]) term = -5;
    else if (flags[FD_FLAG_SINGULAR] || flags[FD_FLAG_NONFINITE]) term = -3;
    else if (m->f32ir && m->solved && !m->ir_converged) term = -4;
    if (report) {
        report->terminationtype = term;
        report->iterationscount = m->f32ir ? m->ir_sweeps : 0;
        report->residual = m->f32ir ? m->ir_residual : 0.0;
        report->n = m->N;
        report->npoly = m->np;
        report->frames = m->F;
        report->reserved = flags[FD_FLAG_SINGULAR];
        report->min_pivot = piv[0];
        report->max_pivot = piv[1];
        report->cancellation = est[0];
        report->eval_kernel = m->solved ? (m->eval64 ? FD_SEL_FP64 : (sel ? sel : (m->use_tc ? FD_SEL_TENSOR : FD_SEL_SIMT))) : 0;
        report->eval_inexact = flags[FD_FLAG_EVAL_INEXACT];
    }
    if (term != 1) { // SOP_FaceDeform.cpp:365-368
        FD_SET_ERR(ctx, "%s (terminationtype %d)", fd_status_string(FD_E_SINGULAR), term);
        return FD_E_SINGULAR;
    }
    return FD_OK;
}

int fd_rbf_eval_dev(fd_model* m, const float* P, int64_t n_vtx, const float* dist2, const float* tangentu,
                    const float* tangentv, const float* normal, float* P_out, float* falloff_out)
{
    if (!m || (n_vtx > 0 && (!P || !P_out)) || n_vtx < 0) return FD_E_INVALID;
    if (m->v1_layers && m->solved) m = m->v1_eval; // the stacked model of the layered fit
    fd_ctx* ctx = m->ctx;
    DeviceGuard g(ctx->device);
    if (!m->solved) { FD_SET_ERR(ctx, "eval: no weights (call fd_rbf_solve or fd_model_commit_weights first)"); return FD_E_STATE; }
    phase_begin(ctx, FD_PH_EVAL);
    cudaError_t e = fd_launch_eval(ctx, m, P, n_vtx, dist2, tangentu, tangentv, normal, P_out, falloff_out);
    phase_end(ctx, FD_PH_EVAL);
    if (e != cudaSuccess) { FD_SET_ERR(ctx, "eval: %s", cudaGetErrorString(e)); return FD_E_CUDA; }
    return FD_OK;
}

int fd_rbf_eval(fd_model* m, const float* P, int64_t n_vtx, const float* dist2, const float* tangentu,
                const float* tangentv, const float* normal, float* P_out, float* falloff_out)
{
    return fd_eval_host_strided(m, P, n_vtx, dist2, tangentu, tangentv, normal, P_out, (size_t)n_vtx * 3 * sizeof(float),
                                falloff_out);
}

// ---- multi-GPU plumbing ------------------------------------------------------------------------------------

int fd_model_create_receiver(fd_ctx* ctx, const fd_params* params, const float* rest_ctrl, int32_t n_ctrl,
                             int32_t frames, fd_model** out)
{
    if (!ctx || !params || !out || !rest_ctrl || n_ctrl < 1 || frames < 1) return FD_E_INVALID;
    *out = nullptr;
    DeviceGuard g(ctx->device);
    int st = check_params(ctx, params);
    if (st != FD_OK) return st;
    if (params->fidelity == FD_FIDELITY_ALGLIB_V1) {
        // the layered fit's weights describe N * layers stacked centres (v1_solve): the receiver is an ordinary one of
        // that size, its centres are the control points repeated once per layer
        if (params->kernel != FD_KERNEL_GAUSSIAN) { FD_SET_ERR(ctx, "the ALGLIB v1 formulation is Gaussian only"); return FD_E_UNSUPPORTED; }
        const int L = params->model == FD_MODEL_QNN ? 1 : (params->layers < 1 ? 1 : params->layers);
        if (L > FD_V1_MAX_LAYERS) { FD_SET_ERR(ctx, "at most %d layers", FD_V1_MAX_LAYERS); return FD_E_INVALID; }
        float* d_stack = nullptr;
        st = dev_alloc(ctx, &d_stack, (size_t)n_ctrl * L * 3);
        if (st != FD_OK) return st;
        for (int k = 0; k < L; ++k)
            cudaMemcpyAsync(d_stack + (size_t)k * n_ctrl * 3, rest_ctrl, (size_t)n_ctrl * 3 * sizeof(float), cudaMemcpyDefault, ctx->stream);
        fd_params pe = *params;
        pe.fidelity = FD_FIDELITY_DENSE;
        st = fd_model_create_receiver(ctx, &pe, d_stack, n_ctrl * L, frames, out);
        cudaFreeAsync(d_stack, ctx->stream); // stream-ordered: the receiver's copy of the centres is enqueued before it
        return st;
    }
    fd_model* m = nullptr;
    st = model_alloc(ctx, params, n_ctrl, false, &m);
    if (st != FD_OK) return st;
    st = model_reserve_frames(m, frames);
    if (st != FD_OK) { fd_model_destroy(m); return st; }
    m->F = frames;
    m->ldw = fd_round_up(3 * frames, 4);
    m->ldw32 = m->ldw;
    model_choose_eval(m, frames);
    // no host synchronisation: pageable sources are staged by the driver before the call returns, pinned or device
    // sources must stay valid until the ctx stream has consumed them (like every *_dev entry point)
    cudaError_t e = cudaMemcpyAsync(m->d_rest, rest_ctrl, (size_t)n_ctrl * 3 * sizeof(float), cudaMemcpyDefault, ctx->stream);
    if (e != cudaSuccess) { FD_SET_ERR(ctx, "receiver: %s", cudaGetErrorString(e)); fd_model_destroy(m); return FD_E_CUDA; }
    *out = m;
    return FD_OK;
}

int fd_model_weights_dev(fd_model* m, void** ptr, size_t* bytes)
{
    if (!m || !ptr || !bytes) return FD_E_INVALID;
    if (m->v1_layers && m->v1_eval) m = m->v1_eval;
    if (!m->d_W || m->F < 1) { FD_SET_ERR(m->ctx, "weights: nothing solved or reserved yet"); return FD_E_STATE; }
    *ptr = m->d_W;
    *bytes = (size_t)m->n * m->ldw * sizeof(double);
    return FD_OK;
}

int fd_model_radii_dev(fd_model* m, void** ptr, size_t* bytes)
{
    if (!m || !ptr || !bytes) return FD_E_INVALID;
    if (m->v1_layers && m->v1_eval) m = m->v1_eval;
    *ptr = m->d_radii;
    *bytes = (size_t)m->N * sizeof(double);
    return FD_OK;
}

int fd_model_commit_weights(fd_model* m)
{
    if (!m) return FD_E_INVALID;
    fd_ctx* ctx = m->ctx;
    DeviceGuard g(ctx->device);
    if (!m->d_W || m->F < 1) { FD_SET_ERR(ctx, "commit: no weight block reserved"); return FD_E_STATE; }
    cudaError_t e = cudaMemsetAsync(m->d_flags + FD_FLAG_NONFINITE, 0, 2 * sizeof(int), ctx->stream); // per commit, like a solve
    if (e == cudaSuccess) e = fd_launch_pack(ctx, m);
    if (e != cudaSuccess) { FD_SET_ERR(ctx, "commit: %s", cudaGetErrorString(e)); return FD_E_CUDA; }
    m->solved = true;
    return FD_OK;
}

int fd_model_info(const fd_model* m, int32_t* n_ctrl, int32_t* npoly, int32_t* frames, int32_t* weights_ld)
{
    if (!m) return FD_E_INVALID;
    if (m->v1_layers && m->v1_eval) m = m->v1_eval; // N * layers centres
    if (n_ctrl) *n_ctrl = m->N;
    if (npoly) *npoly = m->np;
    if (frames) *frames = m->F;
    if (weights_ld) *weights_ld = m->ldw;
    return FD_OK;
}

int fd_model_get_weights(fd_model* m, double* weights, double* radii)
{
    if (!m) return FD_E_INVALID;
    if (m->v1_layers && m->v1_eval) m = m->v1_eval;
    fd_ctx* ctx = m->ctx;
    DeviceGuard g(ctx->device);
    if (weights) {
        if (!m->solved) { FD_SET_ERR(ctx, "get_weights: nothing solved yet"); return FD_E_STATE; }
        FD_CUDA_OK(ctx, cudaMemcpy2DAsync(weights, (size_t)3 * m->F * sizeof(double), m->d_W, (size_t)m->ldw * sizeof(double),
                                          (size_t)3 * m->F * sizeof(double), m->n, cudaMemcpyDeviceToHost, ctx->stream));
    }
    if (radii) FD_CUDA_OK(ctx, cudaMemcpyAsync(radii, m->d_radii, (size_t)m->N * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    FD_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
    return FD_OK;
}

} // extern "C"

// handles may be created / destroyed from different host threads (one ctx per node): the count is atomic
void fd_ctx_retain(fd_ctx* ctx) { __atomic_add_fetch(&ctx->refs, 1, __ATOMIC_ACQ_REL); }
void fd_ctx_release(fd_ctx* ctx)
{
    if (__atomic_sub_fetch(&ctx->refs, 1, __ATOMIC_ACQ_REL) == 0 && ctx->destroy_requested) ctx_teardown(ctx);
}

// host-pointer evaluation; frame f of the result lands at P_out + f * out_pitch_bytes (fd_rbf_eval: the frames are
// contiguous; fd_mgpu_eval: every device writes its vertex range into the caller's F x V x 3 array)
int fd_eval_host_strided(fd_model* m, const float* P, int64_t n_vtx, const float* dist2, const float* tangentu,
                         const float* tangentv, const float* normal, float* P_out, size_t out_pitch_bytes, float* falloff_out)
{
    if (!m || (n_vtx > 0 && (!P || !P_out)) || n_vtx < 0) return FD_E_INVALID;
    if (m->v1_layers && m->solved) m = m->v1_eval;
    fd_ctx* ctx = m->ctx;
    DeviceGuard g(ctx->device);
    if (!m->solved) { FD_SET_ERR(ctx, "eval: no weights (call fd_rbf_solve or fd_model_commit_weights first)"); return FD_E_STATE; }
    if (n_vtx == 0) return FD_OK;
    const size_t v3 = (size_t)n_vtx * 3 * sizeof(float), v1 = (size_t)n_vtx * sizeof(float);
    if (out_pitch_bytes < v3) { FD_SET_ERR(ctx, "eval: output pitch smaller than a frame"); return FD_E_INVALID; }
    void *dP = nullptr, *dD = nullptr, *dU = nullptr, *dV = nullptr, *dN = nullptr, *dO = nullptr, *dF = nullptr;
    int st = stage(ctx, FD_STAGE_P, v3, &dP);
    if (st == FD_OK && dist2) st = stage(ctx, FD_STAGE_DIST, v1, &dD);
    const bool tang = m->prm.tangent && tangentu && tangentv && normal;
    if (st == FD_OK && tang) st = stage(ctx, FD_STAGE_TU, v3, &dU);
    if (st == FD_OK && tang) st = stage(ctx, FD_STAGE_TV, v3, &dV);
    if (st == FD_OK && tang) st = stage(ctx, FD_STAGE_N, v3, &dN);
    if (st == FD_OK) st = stage(ctx, FD_STAGE_OUT, v3 * (size_t)m->F, &dO);
    if (st == FD_OK && falloff_out) st = stage(ctx, FD_STAGE_FALLOFF, v1, &dF);
    if (st != FD_OK) return st;
    cudaStream_t s = ctx->stream;
    FD_CUDA_OK(ctx, cudaMemcpyAsync(dP, P, v3, cudaMemcpyHostToDevice, s));
    if (dD) FD_CUDA_OK(ctx, cudaMemcpyAsync(dD, dist2, v1, cudaMemcpyHostToDevice, s));
    if (tang) {
        FD_CUDA_OK(ctx, cudaMemcpyAsync(dU, tangentu, v3, cudaMemcpyHostToDevice, s));
        FD_CUDA_OK(ctx, cudaMemcpyAsync(dV, tangentv, v3, cudaMemcpyHostToDevice, s));
        FD_CUDA_OK(ctx, cudaMemcpyAsync(dN, normal, v3, cudaMemcpyHostToDevice, s));
    }
    // Wide batches: evaluate in blocks of 80 frames (one 240-column block of the tensor path) and read block i back on a
    // second stream while block i + 1 is evaluated -- the D2H copy bounds this entry point (288 MB at C2), the kernels hide
    // behind it.  Needs pinned host memory to overlap; pageable buffers still work, serialised by the driver.
    constexpr int FB = 80;
    const int F = m->F;
    const int nblk = (F >= 2 * FB && v3 * (size_t)F >= ((size_t)32 << 20)) ? (F + FB - 1) / FB : 1;
    if (nblk > 1 && !ctx->copy_stream) {
        FD_CUDA_OK(ctx, cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
        for (int i = 0; i < 8; ++i) FD_CUDA_OK(ctx, cudaEventCreateWithFlags(&ctx->ev_block[i], cudaEventDisableTiming));
        FD_CUDA_OK(ctx, cudaEventCreateWithFlags(&ctx->ev_copied, cudaEventDisableTiming));
    }
    if (nblk == 1) {
        st = fd_rbf_eval_dev(m, (const float*)dP, n_vtx, (const float*)dD, (const float*)dU, (const float*)dV,
                             (const float*)dN, (float*)dO, (float*)dF);
        if (st != FD_OK) return st;
        if (out_pitch_bytes == v3)
            FD_CUDA_OK(ctx, cudaMemcpyAsync(P_out, dO, v3 * (size_t)F, cudaMemcpyDeviceToHost, s));
        else
            FD_CUDA_OK(ctx, cudaMemcpy2DAsync(P_out, out_pitch_bytes, dO, v3, v3, (size_t)F, cudaMemcpyDeviceToHost, s));
    } else {
        phase_begin(ctx, FD_PH_EVAL);
        for (int b = 0; b < nblk; ++b) {
            const int f0 = b * FB, fc = F - f0 < FB ? F - f0 : FB;
            cudaError_t e = fd_launch_eval_frames(ctx, m, (const float*)dP, n_vtx, (const float*)dD, (const float*)dU, (const float*)dV,
                                                  (const float*)dN, (float*)dO, (float*)dF, f0, fc);
            if (e != cudaSuccess) { FD_SET_ERR(ctx, "eval: %s", cudaGetErrorString(e)); return FD_E_CUDA; }
            FD_CUDA_OK(ctx, cudaEventRecord(ctx->ev_block[b & 7], s));
            FD_CUDA_OK(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_block[b & 7], 0));
            FD_CUDA_OK(ctx, cudaMemcpy2DAsync((char*)P_out + (size_t)f0 * out_pitch_bytes, out_pitch_bytes, (const char*)dO + (size_t)f0 * v3,
                                              v3, v3, (size_t)fc, cudaMemcpyDeviceToHost, ctx->copy_stream));
        }
        phase_end(ctx, FD_PH_EVAL);
        FD_CUDA_OK(ctx, cudaEventRecord(ctx->ev_copied, ctx->copy_stream));
        FD_CUDA_OK(ctx, cudaStreamWaitEvent(s, ctx->ev_copied, 0)); // later work on the ctx stream may reuse the staging buffer
    }
    if (dF) FD_CUDA_OK(ctx, cudaMemcpyAsync(falloff_out, dF, v1, cudaMemcpyDeviceToHost, s));
    FD_CUDA_OK(ctx, cudaStreamSynchronize(s));
    return FD_OK;
}

// fd_mgpu's p2p transport: this (receiver) model builds its evaluation tables straight from the weight block of the root
// device -- the table builders read it through peer loads over NVLink, so the broadcast and the pack are one pass -- and
// pulls the FP64 block itself only when the FP64 evaluation will run.  Enqueued on this model's stream; the caller has
// ordered it after the root's solve (event).
int fd_model_commit_from_peer(fd_model* m, const double* peer_W, const double* peer_radii, int peer_device)
{
    if (!m || !peer_W || !peer_radii) return FD_E_INVALID;
    fd_ctx* ctx = m->ctx;
    DeviceGuard g(ctx->device);
    if (!m->d_W || m->F < 1) { FD_SET_ERR(ctx, "commit: no weight block reserved"); return FD_E_STATE; }
    cudaError_t e = cudaMemcpyPeerAsync(m->d_radii, ctx->device, peer_radii, peer_device, (size_t)m->N * sizeof(double), ctx->stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(m->d_flags + FD_FLAG_NONFINITE, 0, 2 * sizeof(int), ctx->stream);
    m->d_W_src = peer_W;
    if (e == cudaSuccess) e = fd_launch_pack(ctx, m);
    if (e == cudaSuccess) e = fd_launch_pull_weights(ctx, m);
    m->d_W_src = nullptr;
    if (e != cudaSuccess) { FD_SET_ERR(ctx, "commit (p2p): %s", cudaGetErrorString(e)); return FD_E_CUDA; }
    m->solved = true;
    return FD_OK;
}

// ---- serialise (the reference's intent: alglib::rbfserialize, SOP_FaceDeform.cpp:377) --------------------------------------
namespace {
struct SaveHeader {
    char magic[8];          // "FDMODEL"
    uint32_t abi, header_bytes;
    fd_params prm;
    int32_t N, np, n, lda, F, ldw;
    int32_t has_factor, ns, reserved0, reserved1;
    uint64_t total_bytes;
};
struct Section { void** dev; size_t bytes; };

// the device blocks a saved model consists of, in file order
int save_sections(fd_model* m, bool has_factor, Section* sec)
{
    int k = 0;
    sec[k++] = {(void**)&m->d_rest, (size_t)m->N * 3 * sizeof(float)};
    sec[k++] = {(void**)&m->d_radii, (size_t)m->N * sizeof(double)};
    if (m->F > 0) sec[k++] = {(void**)&m->d_W, (size_t)m->n * m->ldw * sizeof(double)};
    if (has_factor) {
        sec[k++] = {(void**)&m->d_A, (size_t)m->lda * m->n * sizeof(double)};
        sec[k++] = {(void**)&m->d_ipiv, (size_t)m->n * sizeof(int)};
        sec[k++] = {(void**)&m->d_perm, (size_t)m->n * sizeof(int)};
        sec[k++] = {(void**)&m->d_Tinv, (size_t)((m->n + 31) / 32) * 2 * 32 * 32 * sizeof(double)};
        sec[k++] = {(void**)&m->d_pivstat, 2 * sizeof(double)};
        sec[k++] = {(void**)&m->d_flags, FD_NUM_FLAGS * sizeof(int)};
        if (m->ns) sec[k++] = {(void**)&m->d_ns, ((size_t)m->N * 5 + 32) * sizeof(double)};
    }
    return k;
}
} // namespace

extern "C" int fd_model_save(fd_model* m, void* buf, size_t cap, size_t* bytes)
{
    if (!m || !bytes) return FD_E_INVALID;
    if (m->v1_layers) { // the layered fit is saved as what evaluates it: the stacked model of its last solve
        if (!m->v1_eval) { FD_SET_ERR(m->ctx, "save: a layered (ALGLIB v1) model is saved after a solve"); return FD_E_STATE; }
        m = m->v1_eval;
    }
    fd_ctx* ctx = m->ctx;
    DeviceGuard g(ctx->device);
    const bool has_factor = m->fitted && !m->receiver && !m->f32ir && m->d_A;
    if (!has_factor && !(m->solved && m->F > 0)) { FD_SET_ERR(ctx, "save: the model holds neither a factorisation nor weights"); return FD_E_STATE; }
    SaveHeader h;
    memset(&h, 0, sizeof(h));
    memcpy(h.magic, "FDMODEL", 8);
    h.abi = FD_ABI_VERSION;
    h.header_bytes = sizeof(h);
    h.prm = m->prm;
    h.N = m->N, h.np = m->np, h.n = m->n, h.lda = m->lda;
    h.F = m->solved ? m->F : 0;
    h.ldw = m->solved ? m->ldw : 0;
    h.has_factor = has_factor ? 1 : 0;
    h.ns = (has_factor && m->ns) ? 1 : 0;
    const int keepF = m->F;
    if (!m->solved) m->F = 0;
    Section sec[12];
    const int nsec = save_sections(m, has_factor, sec);
    m->F = keepF;
    size_t total = sizeof(h);
    for (int k = 0; k < nsec; ++k) total += (sec[k].bytes + 15) & ~(size_t)15;
    h.total_bytes = total;
    *bytes = total;
    if (!buf) return FD_OK;
    if (cap < total) { FD_SET_ERR(ctx, "save: buffer of %zu bytes, %zu needed", cap, total); return FD_E_INVALID; }
    unsigned char* p = (unsigned char*)buf;
    memcpy(p, &h, sizeof(h));
    p += sizeof(h);
    for (int k = 0; k < nsec; ++k) {
        FD_CUDA_OK(ctx, cudaMemcpyAsync(p, *sec[k].dev, sec[k].bytes, cudaMemcpyDeviceToHost, ctx->stream));
        p += (sec[k].bytes + 15) & ~(size_t)15;
    }
    FD_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
    return FD_OK;
}

extern "C" int fd_model_load(fd_ctx* ctx, const void* buf, size_t bytes, fd_model** out)
{
    if (!ctx || !buf || !out) return FD_E_INVALID;
    *out = nullptr;
    DeviceGuard g(ctx->device);
    SaveHeader h;
    if (bytes < sizeof(h)) { FD_SET_ERR(ctx, "load: truncated header"); return FD_E_INVALID; }
    memcpy(&h, buf, sizeof(h));
    if (memcmp(h.magic, "FDMODEL", 8) != 0 || h.abi != FD_ABI_VERSION || h.header_bytes != sizeof(h) || h.total_bytes > bytes) {
        FD_SET_ERR(ctx, "load: not a model saved by this ABI version (%d)", FD_ABI_VERSION);
        return FD_E_INVALID;
    }
    int st = check_params(ctx, &h.prm);
    if (st != FD_OK) return st;
    if (h.N < 1 || h.np != fd_poly_terms(h.prm.term) || h.n != h.N + h.np || h.lda != fd_round_up(h.n, 32) || h.F < 0 ||
        (h.F > 0 && h.ldw != fd_round_up(3 * h.F, 4)) || h.prm.fidelity != FD_FIDELITY_DENSE) {
        FD_SET_ERR(ctx, "load: inconsistent header");
        return FD_E_INVALID;
    }
    fd_model* m = nullptr;
    st = model_alloc(ctx, &h.prm, h.N, h.has_factor != 0, &m);
    if (st != FD_OK) return st;
    m->f32ir = false;
    m->ns = h.ns != 0;
    if (m->ns && (st = dev_alloc(ctx, &m->d_ns, (size_t)m->N * 5 + 32)) != FD_OK) { fd_model_destroy(m); return st; }
    if (h.F > 0) {
        st = model_reserve_frames(m, h.F);
        if (st != FD_OK) { fd_model_destroy(m); return st; }
        m->F = h.F;
        m->ldw = m->ldw32 = h.ldw;
        model_choose_eval(m, h.F);
    }
    Section sec[12];
    const int nsec = save_sections(m, h.has_factor != 0, sec);
    size_t need = sizeof(h);
    for (int k = 0; k < nsec; ++k) need += (sec[k].bytes + 15) & ~(size_t)15;
    if (need != h.total_bytes) { FD_SET_ERR(ctx, "load: size mismatch (%zu expected, %llu stored)", need, (unsigned long long)h.total_bytes); fd_model_destroy(m); return FD_E_INVALID; }
    const unsigned char* p = (const unsigned char*)buf + sizeof(h);
    cudaError_t e = cudaSuccess;
    for (int k = 0; k < nsec && e == cudaSuccess; ++k) {
        e = cudaMemcpyAsync(*sec[k].dev, p, sec[k].bytes, cudaMemcpyHostToDevice, ctx->stream);
        p += (sec[k].bytes + 15) & ~(size_t)15;
    }
    m->fitted = h.has_factor != 0;
    m->receiver = !m->fitted;
    if (e == cudaSuccess && h.F > 0) {
        m->tc_packed_by_solve = false;
        e = fd_launch_pack(ctx, m);
        m->solved = e == cudaSuccess;
    } else if (e == cudaSuccess) {
        e = fd_launch_pack_tables(ctx, m);
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream); // `buf` may be pageable and freed by the caller on return
    if (e != cudaSuccess) { FD_SET_ERR(ctx, "load: %s", cudaGetErrorString(e)); fd_model_destroy(m); return FD_E_CUDA; }
    *out = m;
    return FD_OK;
}

// exposed to fd_capture_host.cu
int fd_stage(fd_ctx* ctx, int slot, size_t bytes, void** out) { return stage(ctx, slot, bytes, out); }
