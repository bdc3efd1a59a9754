// fd_internal.h -- shared declarations of libfacedeform_gpu.so (not part of the C ABI).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "facedeform_gpu.h"

#define FD_NUM_FLAGS 8
#define FD_TMAP_BYTES 128
#define FD_V1_MAX_LAYERS 16      // FD_FIDELITY_ALGLIB_V1: layers kept per model
#define FD_IR_MAX_SWEEPS 40     // FD_FACTOR_FP32_IR: refinement sweeps at most
#define FD_IR_TOLERANCE 1e-12   // converged: max |B - A X| <= this * max |B|
#define FD_IR_FLOOR_OK 1e-9     // a residual that stagnates below this sits at the FP64 floor of the system: accepted
#define FD_TC_MIN_COLUMNS 48 // FD_PATH_AUTO takes the tensor-core evaluation from 3F >= 48 columns
// device-side status words (ints): written by kernels, read by fd_model_report
#define FD_FLAG_ZERO_RADIUS 0 // != 0: a QNN radius was zero (duplicate centres)      -> terminationtype -5
#define FD_FLAG_SINGULAR 1    // k+1 of the first zero pivot                           -> terminationtype -3
#define FD_FLAG_NONFINITE 2   // weights contain NaN/Inf                               -> terminationtype -3
#define FD_FLAG_EVAL_INEXACT 3 // exact-digit tensor-core evaluation: a vertex left the exact range -> fd_report.eval_inexact

// device staging slots owned by the ctx (host-pointer entry points copy through them)
enum fd_stage_slot {
    FD_STAGE_P = 0, FD_STAGE_DIST = 1, FD_STAGE_TU = 2, FD_STAGE_TV = 3, FD_STAGE_N = 4, FD_STAGE_MISC = 5,
    FD_STAGE_OUT = 6, FD_STAGE_FALLOFF = 7, FD_NUM_STAGE = 8
};
int fd_stage(struct fd_ctx* ctx, int slot, size_t bytes, void** out); // grows the slot on demand (fd_api.cu)

enum fd_phase { FD_PH_ASSEMBLE = 0, FD_PH_FACTOR = 1, FD_PH_SOLVE = 2, FD_PH_EVAL = 3, FD_PH_COUNT = 4 };

// Development knobs.  The environment is read ONCE, by fd_ctx_create (fd_api.cu: read_debug_opts); no entry point on
// a cook's path calls getenv.  They select older kernels for comparison or switch on the cycle probes of the debug
// instantiations; none of them changes a result of the default path.
struct fd_debug_opts {
    bool no_nullspace;      // FD_NO_NULLSPACE: multiquadric / thin plate take the pivoted LU
    bool force_pivoted_lu;  // FD_FORCE_PIVOTED_LU
    bool lu_unfused;        // FD_LU_UNFUSED: per-block-column launches instead of the fused LU
    bool solve_dfma;        // FD_SOLVE_DFMA: the DFMA slab solve instead of the DMMA one
    bool no_fused_pack;     // FD_NO_FUSED_PACK: separate tensor-table pack kernels after the slab solve
    bool no_few_rhs;        // FD_NO_FEW_RHS
    bool no_inverse;        // FD_NO_INVERSE: per-cook solves never build / use the explicit inverse
    bool eval_scalar_f32;   // FD_EVAL_SCALAR_F32: the un-packed FP32 evaluation kernel
    bool tc_nopair;         // FD_TC_NOPAIR: no CTA pairs in the tensor evaluation
    bool has_tc_debug;      // FD_TC_DEBUG set
    bool lu_sym_off;        // FD_LU_NOSYM: the fused LU ignores symmetry
    int eval_vp;            // FD_EVAL_VP
    int tc_debug;           // FD_TC_DEBUG bits
    int lu_debug;           // FD_LU_DEBUG step
    int lu_nbo;             // FD_LU_NBO
    int lu_cluster_max_n;   // FD_LU_CLUSTER_MAX_N
    int lu_cluster;         // FD_LU_CLUSTER
    int tcx_cbu;            // FD_TCX_CBU: 1 = the exact-digit evaluation takes one column block per unit (default: two)
    int tcx_narrow;         // FD_TCX_NARROW: one MMA of N = 128 per column block instead of one of N = 240 for the unit's two
    bool poison;            // FD_POISON: every device allocation starts as 0xFF bytes (NaN / -1), so a read of memory the
                            // library never wrote shows in the results instead of depending on what the pool held before
};

struct fd_ctx {
    int device;
    int sm_count;
    cudaStream_t stream;
    bool own_stream;
    char err[512];
    cudaEvent_t ev_begin[FD_PH_COUNT];
    cudaEvent_t ev_end[FD_PH_COUNT];
    bool phase_valid[FD_PH_COUNT];
    int64_t launches;
    // host<->device staging owned by the ctx (grown on demand)
    void* stage_dev[FD_NUM_STAGE];
    size_t stage_bytes[FD_NUM_STAGE];
    // grid-barrier counter of the persistent factorisation kernels: monotonic, the host tracks its value (sync_base)
    unsigned* d_sync;
    unsigned sync_base;
    // handles created from this ctx (fd_model, fd_dbse) keep it alive: fd_ctx_destroy with live handles only marks
    // the ctx, the teardown happens when the last handle is destroyed (any destruction order is safe for the caller)
    int refs;
    bool destroy_requested;
    fd_debug_opts dbg;       // environment knobs, read once at creation
    long long* d_tc_dbg;     // cycle probes of the debug instantiations (allocated at creation when the knob is set)
    long long* d_lu_dbg;
    // tensor map of the last output buffer the tensor evaluation wrote (fd_eval_tc.cu): re-encoded only when it changes
    alignas(64) unsigned char tc_out_map[FD_TMAP_BYTES];
    const void* tc_out_ptr;
    int64_t tc_out_V;
    int tc_out_F;
    // host-pointer evaluation of wide batches: the read-back of frame block i runs on copy_stream while block i + 1 is
    // evaluated (created on first use)
    cudaStream_t copy_stream;
    cudaEvent_t ev_block[8];
    cudaEvent_t ev_copied;
};
// per-device kernel attributes (dynamic shared-memory limits, non-portable cluster sizes): function attributes are per
// device, so every ctx sets them for its own device at creation (never behind a process-wide flag)
cudaError_t fd_solve_setup(fd_ctx* ctx);     // fd_solve.cu
cudaError_t fd_factor_setup(fd_ctx* ctx);    // fd_factor.cu
cudaError_t fd_eval_tc_setup(fd_ctx* ctx);   // fd_eval_tc.cu
cudaError_t fd_eval64_setup(fd_ctx* ctx);    // fd_eval64.cu
cudaError_t fd_eval_tcx_setup(fd_ctx* ctx);   // fd_eval_tcx.cu
void fd_ctx_retain(fd_ctx* ctx);
void fd_ctx_release(fd_ctx* ctx); // fd_api.cu

struct fd_model {
    fd_ctx* ctx;
    fd_params prm;
    int N;    // control points
    int np;   // polynomial terms (4 / 1 / 0)
    int n;    // N + np
    int lda;  // column stride of the FP64 system (column-major)
    int F;    // frames of the last solve (0: none)
    int capF; // allocated frames
    int ldw;  // row stride (doubles) of the FP64 weight block, row-major n x ldw
    bool receiver;
    bool fitted;
    bool solved;
    bool eval64;     // the evaluation is FP64 whatever the weights (eval_precision FP64; AUTO with multiquadric / thin plate)
    bool auto_sel;   // Gaussian + FD_EVAL_AUTO: FP32 or FP64 is settled per solve on the device (d_sel, fd_eval64.cu)
    int* d_sel;      // the evaluation kernel chosen for the current weights: 1 FMA/SFU, 2 tensor cores, 3 FP64
    double* d_est;   // [0] cancellation S, [1] bounding-box diagonal of the rig, [2], [3] scratch of the estimate
    float* d_wmax;   // N: max_c |w_jc|
    float* d_rest;   // N x 3
    double* d_radii; // N
    double* d_A;     // lda x n, LU in place
    int* d_ipiv;     // n (global row index swapped with row k)
    int* d_perm;     // n: row i of P*A is row perm[i] of A
    int* d_win;      // 1 + 4*32 ints: the rows the current panel's interchanges touch and their composition
    double* d_Tinv;  // [ceil(n/32)][2][32x32]: inverses of the diagonal blocks of L and U
    double* d_W;     // n x ldw weights (solve in place over the right-hand sides)
    int w_col0;            // frame-block views only (fd_launch_eval_frames): the first column the weight pointers were advanced to
    const double* d_W_src; // where the table builders read the weights from: NULL = d_W; fd_mgpu's p2p transport points it
                           // at the ROOT device's weight block, so the tables are built through peer loads over NVLink
    int* d_flags;    // FD_NUM_FLAGS
    double* d_pivstat; // [min |u_kk|, max |u_kk|]
    // Per-cook solves (a handful of right-hand sides against a cached factorisation, the reference's one frame per cook,
    // SOP_FaceDeform.cpp:215): from the second such solve on the explicit inverse of the factored block is kept and a
    // solve is one pass over it (fd_solve.cu: k_inv_apply) instead of 2 n / 32 dependent block steps.
    double* d_inv;      // n_f x ld_inv, row-major; n_f = the factored block (n, or N - 4 on the null-space path)
    double* d_inv_rhs;  // n_f x 8 right-hand sides in original row order
    int ld_inv;
    int small_solves;   // solves with <= 8 right-hand sides seen so far
    // null-space path of multiquadric / thin plate (fd_nullspace.cu): d_A holds Q^T K Q, its [4:, 4:] block LU-factored;
    // d_ns = reflectors V (N x 4), tau (4), R (4 x 4), scratch (N)
    bool ns;
    double* d_ns;
    // FD_FIDELITY_ALGLIB_V1 (fd_api.cu): the parent holds one factored sub-model per layer and a stacked evaluation
    // model (N * layers centres); d_v1_* are the residual, a scratch for the layer's kernel matrix and the polynomial
    int v1_layers;
    fd_model* v1_layer[FD_V1_MAX_LAYERS];
    fd_model* v1_eval;
    double* d_v1_R;
    double* d_v1_K;
    double* d_v1_V;
    float* d_v1_stack;
    // FD_FACTOR_FP32_IR (fd_refine.cu): d_A stays the assembled FP64 system, the LU lives in d_A32
    bool f32ir;
    float* d_A32;      // lda x n FP32, LU in place
    double* d_B;       // n x ldw right-hand sides
    double* d_R;       // n x ldw residual
    float* d_D32;      // n x ldw correction
    double* d_ir_norm; // [max |B|, max |R|]
    int ir_sweeps;
    double ir_residual;
    bool ir_converged;
    // evaluation tables (built by pack)
    float4* d_ctab32;  // N: (cx, cy, cz, kernel parameter)
    float4* d_ctab_pair; // the same table with pairs of centres interleaved (tensor path, fd_eval_tc.cu)
    float* d_W32;      // n x ldw32 floats
    int ldw32;
    double4* d_ctab64; // N (only when eval64)
    // tensor-core evaluation tables (fd_eval_tc.cu), present when use_tc
    bool tc_packed_by_solve; // the slab solve wrote the tensor path's weight tables itself (fused epilogue)
    bool tables_packed;  // centre tables / normalisation built for the current centres and radii
    bool use_tc;
    float* d_tc_norm;    // (ox, oy, oz, s)
    float* d_tc_scale;   // per padded column: 2^e
    float* d_tc_unscale; // per padded column: 2^-e
    void* d_tc_wt_hi;    // __half [col_pad][Kpad]
    void* d_tc_wt_lo;
    alignas(64) unsigned char tc_map_hi[FD_TMAP_BYTES]; // CUtensorMap
    alignas(64) unsigned char tc_map_lo[FD_TMAP_BYTES];
    // exact-digit tensor-core evaluation (fd_eval_tcx.cu; Gaussian under FD_EVAL_AUTO with 3F >= 48): digit / mid / lo tables in
    // d_tc_wt_hi / d_tcx_wt_mid / d_tc_wt_lo, 120-column blocks
    bool use_tcx;
    void* d_tcx_wt_mid;
    double4* d_ctab_tcx; // per centre (Kpad): (a, b, c, d) of t = q . (a, b, c) + d + |q|^2 sc, q = p - centre 0 (per fit)
    double* d_csc_tcx;   // per centre (Kpad): sc = -log2(e) / R^2
    int* d_tcx_rowexp;   // per row: s_k (Kpad entries), then the digit width h chosen at pack time
    float* d_tcx_rowmax; // per row: max_c |w_kc| 2^e_c
    double* d_tcx_meta;  // scratch of the pack: the bits of max_i sum_k phi_k(c_i) rowmax_k
    double4* d_tcx_ctab_eff; // per solve (Kpad): d_ctab_tcx with the row's digit exponent h - s_k added to d -- what the kernel reads
    float* d_tcx_pw;     // per solve (Kpad): 2^(h - s_k)
    alignas(64) unsigned char tcx_map_mid[FD_TMAP_BYTES];
};

#define FD_CUDA_OK(ctx, call)                                                                      \
    do {                                                                                           \
        cudaError_t e__ = (call);                                                                  \
        if (e__ != cudaSuccess) {                                                                  \
            snprintf((ctx)->err, sizeof((ctx)->err), "%s:%d %s: %s", __FILE__, __LINE__, #call,    \
                     cudaGetErrorString(e__));                                                     \
            return FD_E_CUDA;                                                                      \
        }                                                                                          \
    } while (0)

// makes `dev` current for the scope and restores the caller's device on every exit path
struct fd_device_guard {
    int prev = -1;
    explicit fd_device_guard(int dev) { cudaGetDevice(&prev); if (prev != dev) cudaSetDevice(dev); else prev = -1; }
    ~fd_device_guard() { if (prev >= 0) cudaSetDevice(prev); }
    fd_device_guard(const fd_device_guard&) = delete;
    fd_device_guard& operator=(const fd_device_guard&) = delete;
};

static inline int fd_poly_terms(int term) { return term == FD_TERM_LINEAR ? 4 : (term == FD_TERM_CONST ? 1 : 0); }
static inline int fd_round_up(int x, int m) { return (x + m - 1) / m * m; }

// ---- kernel launchers (each returns a cudaError_t from the launch; all asynchronous on ctx->stream) -------
// fd_assemble.cu
cudaError_t fd_launch_radii(fd_ctx* ctx, const fd_params& prm, const float* d_rest, int N, double* d_radii, int* d_flags);
cudaError_t fd_launch_assemble(fd_ctx* ctx, const fd_params& prm, const float* d_rest, const double* d_radii, int N,
                               int np, double* d_A, int lda);
// fd_factor.cu
cudaError_t fd_launch_lu(fd_ctx* ctx, double* d_A, int lda, int n, int* d_ipiv, int* d_perm, int* d_flags,
                         double* d_pivstat, int* d_win);
cudaError_t fd_launch_lu_nopivot(fd_ctx* ctx, double* d_A, int lda, int n, int* d_ipiv, int* d_perm, int* d_flags,
                                 double* d_pivstat);
// one persistent launch; also writes the inverted diagonal blocks (d_Tinv) the slab solve uses
cudaError_t fd_launch_lu_nopivot_fused(fd_ctx* ctx, double* d_A, int lda, int n, int* d_ipiv, int* d_perm, int* d_flags,
                                       double* d_pivstat, double* d_Tinv, int sym);
// fd_solve.cu
cudaError_t fd_launch_solve(fd_ctx* ctx, fd_model* m, const float* d_deform, int F);
cudaError_t fd_launch_solve_prebuilt(fd_ctx* ctx, fd_model* m, int nrhs);
cudaError_t fd_launch_solve_sub(fd_ctx* ctx, const double* d_A, int lda, int n, const int* d_perm, const double* d_Tinv,
                                double* d_W, int ldw, int nrhs);
// per-cook fast path: true when the solve of `nrhs` right-hand sides was done through the explicit inverse (built on the
// way when this is the second small solve).  rhs_in_W: the right-hand sides sit in W (original row order) instead of
// being built from rest / deform.
bool fd_try_inverse_solve(fd_ctx* ctx, fd_model* m, const double* d_A, int lda, int n_f, const int* d_perm, const double* d_Tinv,
                          const float* d_deform, int F, double* d_W, int ldw, bool rhs_in_W, cudaError_t* err);
// fd_nullspace.cu
cudaError_t fd_launch_ns_transform(fd_ctx* ctx, fd_model* m);
cudaError_t fd_launch_ns_rhs(fd_ctx* ctx, fd_model* m, const float* d_deform, int F);
cudaError_t fd_launch_ns_finish(fd_ctx* ctx, fd_model* m, int F);
cudaError_t fd_launch_v1_rhs_poly(fd_ctx* ctx, const float* d_rest, const float* d_deform, int N, int F, int np, double* d_R,
                                  double* d_V, int ldw, int* d_flags);
cudaError_t fd_launch_v1_gather(fd_ctx* ctx, const double* d_R, const int* d_perm, int N, int ldw, double* d_W);
cudaError_t fd_launch_gemm_sub(fd_ctx* ctx, const double* d_A, int lda, int rows, int K, const double* d_X, double* d_C, int ldw,
                               int nrhs);
cudaError_t fd_launch_pack(fd_ctx* ctx, fd_model* m);
static inline const double* fd_w_src(const fd_model* m) { return m->d_W_src ? m->d_W_src : m->d_W; }
// d_W <- d_W_src when the evaluation needs the FP64 block locally (eval64, or FD_EVAL_AUTO chose FP64 on the device)
cudaError_t fd_launch_pull_weights(fd_ctx* ctx, fd_model* m);
// internals shared with fd_mgpu.cu / the serialiser (fd_api.cu)
int fd_model_commit_from_peer(fd_model* m, const double* peer_W, const double* peer_radii, int peer_device);
int fd_eval_host_strided(fd_model* m, const float* P, int64_t n_vtx, const float* dist2, const float* tangentu,
                         const float* tangentv, const float* normal, float* P_out, size_t out_pitch_bytes, float* falloff_out);
cudaError_t fd_launch_pack_tables(fd_ctx* ctx, fd_model* m);
cudaError_t fd_launch_invdiag(fd_ctx* ctx, fd_model* m);
// fd_refine.cu
cudaError_t fd_launch_to_f32(fd_ctx* ctx, const double* d_A, float* d_A32, size_t count);
cudaError_t fd_refine_solve(fd_ctx* ctx, fd_model* m, const float* d_deform, int F);
cudaError_t fd_launch_lu_f32(fd_ctx* ctx, float* d_A, int lda, int n, int* d_ipiv, int* d_perm, int* d_flags,
                             double* d_pivstat, int* d_win);
cudaError_t fd_launch_lu_nopivot_f32(fd_ctx* ctx, float* d_A, int lda, int n, int* d_ipiv, int* d_perm, int* d_flags,
                                     double* d_pivstat);
// fd_eval.cu
// the frames [f_begin, f_begin + f_count) of the solved batch into P_out + f_begin * V * 3 (f_begin a multiple of 80)
cudaError_t fd_launch_eval_frames(fd_ctx* ctx, const fd_model* m, const float* P, int64_t V, const float* dist2, const float* tu,
                                  const float* tv, const float* nrm, float* P_out, float* falloff_out, int f_begin, int f_count);
bool fd_tc_view_frames(const fd_model* m, fd_model* view, int f_begin); // fd_eval_tc.cu
cudaError_t fd_launch_eval(fd_ctx* ctx, const fd_model* m, const float* P, int64_t V, const float* dist2,
                           const float* tu, const float* tv, const float* nrm, float* P_out, float* falloff_out);
// fd_eval_tc.cu
int fd_tc_kpad(int N);
int fd_tc_ncb(int F);
int fd_tc_col_pad(int F);
cudaError_t fd_launch_pack_tc(fd_ctx* ctx, fd_model* m);
// what the fused pack epilogue of the slab solve needs (fd_solve.cu: k_solve_slab8); enabled = 0: plain solve
struct fd_tc_pack_args {
    int enabled;
    int N, np, Kpad, ncol, ncol_pad, phi_shift;
    const float* norm;  // (ox, oy, oz, s)
    float* scale;       // per padded column
    float* unscale;
    void* wt_hi;        // __half [ncol_pad][Kpad]
    void* wt_lo;
    int* flags;
};
void fd_tc_pack_args_fill(const fd_model* m, fd_tc_pack_args* pk); // fd_eval_tc.cu
cudaError_t fd_launch_tc_norm(fd_ctx* ctx, fd_model* m);
cudaError_t fd_launch_eval_tc(fd_ctx* ctx, const fd_model* m, const float* P, int64_t V, const float* dist2,
                              const float* tu, const float* tv, const float* nrm, float* P_out, float* falloff_out,
                              const int* sel, int sel_id);
// fd_eval_tcx.cu
int fd_tcx_ncb(int F);
int fd_tcx_col_pad(int F);
cudaError_t fd_launch_pack_tcx(fd_ctx* ctx, fd_model* m);
bool fd_tcx_view_frames(const fd_model* m, fd_model* view, int f_begin);
cudaError_t fd_launch_eval_tcx(fd_ctx* ctx, const fd_model* m, const float* P, int64_t V, const float* dist2, const float* tu,
                               const float* tv, const float* nrm, float* P_out, float* falloff_out, const int* sel, int sel_id);
// fd_eval64.cu
#define FD_SEL_SIMT 1
#define FD_SEL_TENSOR 2
#define FD_SEL_FP64 3
#define FD_SEL_TCX 4 // tensor cores, exact leading digit (fd_eval_tcx.cu)
#define FD_MMA64_MIN_COLUMNS 48 // the FP64 evaluation takes the DMMA kernel from 3F >= 48 columns
cudaError_t fd_launch_cancel_select(fd_ctx* ctx, fd_model* m, int tensor_ok, int simt_ok, int tcx_ok, int want);
cudaError_t fd_launch_eval64_mma(fd_ctx* ctx, const fd_model* m, const float* P, int64_t V, const float* dist2, const float* tu,
                                 const float* tv, const float* nrm, float* P_out, float* falloff_out, const int* sel, int sel_id);
// fd_capture.cu
cudaError_t fd_launch_nearest(fd_ctx* ctx, const float* d_P, int64_t V, const float* d_rig, int N,
                              unsigned long long* d_keys /* N scratch */, int32_t* d_nearest);
cudaError_t fd_launch_capture_dist(fd_ctx* ctx, const float* d_P, int64_t V, const uint8_t* d_member,
                                   const float* d_rig, const int32_t* d_tri /* ntri x 3, -1 in [2] = segment */,
                                   int ntri, float radius, int dofalloff, float* d_dist2);
