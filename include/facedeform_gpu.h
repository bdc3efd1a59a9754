/*
 * facedeform_gpu.h -- C ABI of libfacedeform_gpu.so, the B200 (sm_100a) replacement for the RBF deformation
 * path of symek/facedeform.
 *
 * Every entry point replaces a region of the reference's SOP cook (paths are under the reference's src/):
 *
 *   fd_params / fd_params_default / fd_params_clamp
 *        the cook parameters and their clamps          SOP_FaceDeform.cpp:99-137 (templates), :244-263 (clamps)
 *   fd_rbf_fit        rbfcreate + rbfsetpoints + rbfsetalgo* + rbfset*term + rbfbuildmodel
 *                                                      SOP_FaceDeform.cpp:331-363 (the centres half of :268-287)
 *   fd_rbf_solve      the delta half of the pack loop + the solve inside rbfbuildmodel, for F frames at once
 *                                                      SOP_FaceDeform.cpp:268-287, :363-368
 *   fd_rbf_eval       the evaluation loop: gate, rbfcalc, project_to_tangents, falloff, P += disp
 *                                                      SOP_FaceDeform.cpp:384-439, SOP_FaceDeform.hpp:28-41
 *   fd_capture        ProximityCapture::init / capture / findIslands
 *                                                      capture.cpp:10-44, :46-105, :107-141 (capture.hpp:21-27)
 *   fd_dbse_*         DirectBSEdit::init / computeWeights / displaceVector / getWeights (the "morph space" post-pass)
 *                                                      dbse.cpp:9-87 (dbse.hpp:16-22), SOP_FaceDeform.cpp:444-482
 *   fd_model_report   alglib::rbfreport checked at     SOP_FaceDeform.cpp:365-373
 *   fd_last_error     addError / addWarning strings    SOP_FaceDeform.cpp:231-234, :337-340, :365-368
 *
 * Conventions: plain pointers and sizes only; the caller owns every host pointer, the
 * library owns device memory; every function returns an fd_status and never throws or aborts; handles carry all
 * state (no globals), so different handles may be used from different host threads.  One fd_ctx drives one GPU:
 * multi-GPU runs use either fd_mgpu_* (one process, the library owns the devices and the weight broadcast) or one
 * process / thread and one ctx per GPU, vertex ranges sharded by the caller, weights moved with fd_model_weights_dev +
 * fd_model_commit_weights around the caller's broadcast (torch.distributed in facedeform_b200/shard.py).
 *
 * The `*_dev` variants take device pointers, enqueue on the ctx stream and return without synchronising.
 */
#ifndef FACEDEFORM_GPU_H
#define FACEDEFORM_GPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FD_ABI_VERSION 3

typedef struct fd_ctx fd_ctx;     /* one GPU + stream + scratch; replaces the node-member state, SOP_FaceDeform.hpp:108-113 */
typedef struct fd_model fd_model; /* replaces alglib::rbfmodel (SOP_FaceDeform.cpp:332): centres, radii, LU factors, weights */

typedef enum fd_status {
    FD_OK = 0,
    FD_E_INVALID = 1,        /* bad argument */
    FD_E_MISMATCH_POINT = 2, /* SOP_ERR_MISMATCH_POINT, SOP_FaceDeform.cpp:231-234 */
    FD_E_BUILD = 3,          /* "Can't build RBF model." (alglib::ap_error), SOP_FaceDeform.cpp:337-340 */
    FD_E_SINGULAR = 4,       /* "Can't solve the problem." (terminationtype != 1), SOP_FaceDeform.cpp:365-368 */
    FD_E_CUDA = 5,
    FD_E_NOMEM = 6,
    FD_E_CAPTURE = 7,        /* "Can't capture geometry with a rig!", SOP_FaceDeform.cpp:318-321 */
    FD_E_UNSUPPORTED = 8,
    FD_E_STATE = 9           /* call order (eval before solve, ...) */
} fd_status;

/* menu values, SOP_FaceDeform.hpp:13-18 */
#define FD_MODEL_QNN 0        /* per-centre radii from the nearest-neighbour rule (qcoef, zcoef) */
#define FD_MODEL_ML 1         /* uniform `radius` */
#define FD_TERM_LINEAR 0
#define FD_TERM_CONST 1
#define FD_TERM_ZERO 2
/* north_star "kernel type" (the reference itself is Gaussian only) */
#define FD_KERNEL_GAUSSIAN 0     /* exp(-r^2 / R^2) */
#define FD_KERNEL_MULTIQUADRIC 1 /* sqrt(r^2 + R^2) */
#define FD_KERNEL_THINPLATE 2    /* r^2 log r */
/* arithmetic of the vertex evaluation */
#define FD_EVAL_AUTO 0   /* FP32 for the Gaussian, FP64 for multiquadric / thin plate (cancellation) */
#define FD_EVAL_FP32 1
#define FD_EVAL_FP64 2
/* arithmetic of the factorisation (BASELINE.json config 4: "FP64 vs FP32+refinement tolerance study") */
#define FD_FACTOR_FP64 0    /* FP64 LU + FP64 triangular solves (default) */
#define FD_FACTOR_FP32_IR 1 /* FP32 LU, solutions refined in FP64 until the residual stops shrinking (fd_report) */
/* formulation of the fit */
#define FD_FIDELITY_DENSE 0     /* north_star: the augmented saddle-point system [[K + lambda I, P], [P^T, 0]] (default) */
#define FD_FIDELITY_ALGLIB_V1 1 /* what the SOP's two rbfsetalgo* calls mean in ALGLIB's v1 unit (SURVEY appendix B,
                                   unverifiable here): polynomial term fitted first by least squares, then Gaussian layers
                                   on the residual -- QNN: one layer, per-centre radii; Multilayer: `layers` layers with
                                   radius R, R/2, R/4 ...  The model then holds N * layers centres (weights, radii). */
/* evaluation kernel */
#define FD_PATH_AUTO 0   /* tensor cores when 3F is wide enough, else FMA/SFU */
#define FD_PATH_SIMT 1
#define FD_PATH_TENSOR 2

/* The SOP's parameter surface, names = the reference parm tokens (SOP_FaceDeform.cpp:99-115). */
typedef struct fd_params {
    int32_t model;         /* "model"  default 0 (QNN)        :48-53, :121 */
    int32_t term;          /* "term"   default 0 (linear)     :55-61, :122 */
    int32_t kernel;        /* extension, default gaussian */
    float qcoef;           /* default 1, clamp >= 0.1         :123, :249 */
    float zcoef;           /* default 5, clamp >= 0.1         :124, :250 */
    float radius;          /* default 1, clamp >= 0.01        :125, :251  (RBF radius AND capture/falloff radius :318, :402) */
    int32_t layers;        /* default 4, clamp >= 1           :126, :252  (solved with fidelity = FD_FIDELITY_ALGLIB_V1) */
    float lambda;          /* default 0.1, clamp >= 0.01      :128, :253  (K + lambda I; K - lambda I for the conditionally
                              negative definite multiquadric) */
    int32_t tangent;       /* default 0                       :129 */
    int32_t maxedges;      /* default 4, clamp >= 1           :127, :257 */
    int32_t morphspace;    /* default 0                       :130  (dbse post-pass: fd_dbse_*, FaceDeformOp::setBlendshapes) */
    int32_t doclampweight; /* default 0                       :131 */
    float weightrange[2];  /* default (0, 1)                  :132 */
    int32_t dofalloff;     /* default 0                       :133 */
    float falloffradius;   /* default 1                       :134 */
    float falloffrate;     /* default 1                       :135 */
    int32_t eval_precision;/* FD_EVAL_* */
    int32_t eval_path;     /* FD_PATH_* */
    int32_t factor_precision; /* FD_FACTOR_* */
    int32_t fidelity;         /* FD_FIDELITY_* */
    int32_t strict_reference; /* default 0.  1: reproduce the reference's quirks that this library otherwise normalises:
                                 the point `group` never restricts the loop (it only gates the data-ID bump, :380-381,
                                 :485-486) and the morph-space weights are computed only while !isComputed(), later cooks
                                 skip the pass with the reference's warning (:446-452).  0: the group restricts the
                                 deformation, the weights follow every cooked frame. */
    float eval_tolerance;     /* default 1e-5: FD_EVAL_AUTO runs the fastest evaluation whose predicted maximum error stays
                                 within eval_tolerance x the control rig's bounding-box diagonal.  Gaussian, >= 16 frames: the
                                 tensor cores with an exact leading digit (0.15 x 2^-24 S), else FP64; fewer frames: FMA/SFU
                                 FP32 (1.3 ... 1.1 x 2^-24 S, falling with N), else FP64.  S = fd_report.cancellation
                                  */
    char group[64];           /* "group"  default ""  (all points)  :119-120  point-group pattern: "*", "7", "3-40",
                                 "0-100:2", "^5" (remove), space separated; resolved by the host mirror (facedeform_sop.hpp) */
} fd_params;

/* analogue of alglib::rbfreport (SOP_FaceDeform.cpp:333, :365-373) */
typedef struct fd_report {
    int32_t terminationtype; /* 1 = ok; -3 = singular / non-finite weights; -4 = refinement did not converge;
                                -5 = zero radius (duplicate centres) */
    int32_t iterationscount; /* refinement sweeps of FD_FACTOR_FP32_IR (0: direct FP64 solve) */
    int32_t n;               /* control points */
    int32_t npoly;           /* polynomial terms 4 / 1 / 0 */
    int32_t frames;          /* F of the last solve */
    int32_t reserved;
    double min_pivot;        /* min |u_kk| of the LU */
    double max_pivot;
    double residual;         /* FD_FACTOR_FP32_IR: max |B - A X| / max |B| after the last sweep (0 otherwise) */
    double cancellation;     /* after a solve: S = max_i sum_j max_c |w_jc| phi_j(c_i), the size of the terms that cancel in
                                the evaluation sum; an FP32 evaluation errs by about 2^-24 S */
    int32_t eval_kernel;     /* after a solve: the evaluation FD_EVAL_AUTO / FD_PATH_AUTO settled on: 1 FMA/SFU FP32,
                                2 tensor cores (FP16 hi/lo splits), 3 FP64, 4 tensor cores with an exact leading digit
                                (FP64-class accuracy; Gaussian under FD_EVAL_AUTO from 16 frames on) */
    int32_t eval_inexact;    /* != 0: an evaluation with kernel 4 met a vertex whose leading-digit sum may have left the exact
                                range (a mesh far outside the rig, or weights unlike those at the control points): that result
                                is FP32-accurate only; re-evaluate with eval_precision = FD_EVAL_FP64 */

} fd_report;

int fd_abi_version(void);
const char* fd_status_string(int status);

void fd_params_default(fd_params* p);
void fd_params_clamp(fd_params* p); /* SOP_FaceDeform.cpp:249-257 */
/* 1 when a model fitted with `a` serves a cook with `b` without a refit: only epilogue parameters differ (tangent,
 * dofalloff, falloffrate, falloffradius, maxedges, the morph-space and group parameters, and `radius` when it is only
 * the capture / falloff radius, i.e. model = QNN).  The reference refits every cook (SOP_FaceDeform.cpp:331-363). */
int fd_params_fit_equal(const fd_params* a, const fd_params* b);

/* device < 0: current CUDA device.  stream: a cudaStream_t (NULL = the ctx creates its own). */
int fd_ctx_create(fd_ctx** out, int device, void* stream);
void fd_ctx_destroy(fd_ctx* ctx);
int fd_ctx_synchronize(fd_ctx* ctx);
const char* fd_last_error(const fd_ctx* ctx);

/* ---- fit: assemble + factor, once per rest pose ---------------------------------------------------------- */
int fd_rbf_fit(fd_ctx* ctx, const fd_params* params, const float* rest_ctrl /* N x 3 host */, int32_t n_ctrl,
               fd_model** out, fd_report* report /* may be NULL */);
int fd_rbf_fit_dev(fd_ctx* ctx, const fd_params* params, const float* rest_ctrl_dev, int32_t n_ctrl, fd_model** out);
void fd_model_destroy(fd_model* m);

/* ---- solve: weights for F frames at once (multi-RHS) ------------------------------------------------------ */
int fd_rbf_solve(fd_model* m, const float* deform_ctrl /* F x N x 3 host */, int32_t n_ctrl, int32_t frames,
                 fd_report* report /* may be NULL */);
int fd_rbf_solve_dev(fd_model* m, const float* deform_ctrl_dev, int32_t n_ctrl, int32_t frames);

/* the epilogue parameters of later fd_rbf_eval calls (see fd_params_fit_equal); FD_E_INVALID when `p` needs a refit */
int fd_model_set_epilogue(fd_model* m, const fd_params* p);

/* synchronises the ctx stream and reads the status flags of the last fit/solve; returns FD_OK or FD_E_SINGULAR */
int fd_model_report(fd_model* m, fd_report* report);

/* ---- eval: P_out[f][v] = P[v] + falloff(v) * tangent_project(rbf_f(P[v])) ------------------------------- */
int fd_rbf_eval(fd_model* m, const float* P /* V x 3 */, int64_t n_vtx,
                const float* dist2 /* V or NULL */, const float* tangentu, const float* tangentv,
                const float* normal /* V x 3 each or NULL */,
                float* P_out /* F x V x 3 */, float* falloff_out /* V or NULL */);
int fd_rbf_eval_dev(fd_model* m, const float* P, int64_t n_vtx, const float* dist2, const float* tangentu,
                    const float* tangentv, const float* normal, float* P_out, float* falloff_out);

/* ---- multi-GPU plumbing: a model that receives weights instead of solving for them ------------------------ */
int fd_model_create_receiver(fd_ctx* ctx, const fd_params* params, const float* rest_ctrl /* host */, int32_t n_ctrl,
                             int32_t frames, fd_model** out);
/* device pointer + size of the FP64 weight block ((N + npoly) x ld doubles, ld = fd_model_weights_ld) */
int fd_model_weights_dev(fd_model* m, void** ptr, size_t* bytes);
int fd_model_radii_dev(fd_model* m, void** ptr, size_t* bytes); /* N doubles */
/* after the caller has written (e.g. NCCL-broadcast) the weight and radii blocks: build the evaluation tables */
int fd_model_commit_weights(fd_model* m);
int fd_model_info(const fd_model* m, int32_t* n_ctrl, int32_t* npoly, int32_t* frames, int32_t* weights_ld);
/* copies weights to the host as (N + npoly) x 3F row-major doubles, radii as N doubles (either may be NULL) */
int fd_model_get_weights(fd_model* m, double* weights, double* radii);

/* ---- serialise: the reference's intent at SOP_FaceDeform.cpp:377 (alglib::rbfserialize, commented out) ------
 * fd_model_save writes parameters, centres, radii, the factorisation (FP64 dense fits) and the weights of the last
 * solve into `buf`; buf == NULL queries the size.  fd_model_load rebuilds a model on `ctx` that solves (when the
 * factorisation was saved) and evaluates bit-identically to the saved one. */
int fd_model_save(fd_model* m, void* buf, size_t cap, size_t* bytes);
int fd_model_load(fd_ctx* ctx, const void* buf, size_t bytes, fd_model** out);

/* ---- several GPUs of one box behind one handle (single process) -------------------------------------
 * The evaluation loop SOP_FaceDeform.cpp:404-439 has no cross-vertex dependence, so it is partitioned by contiguous
 * vertex range over the devices; device devices[0] assembles, factors and solves, then the weights cross NVLink once:
 * transport "nccl" (ncclBroadcast of the FP64 weight block, libnccl.so.2 loaded at run time), or "p2p" (each device
 * builds its evaluation tables straight from the root's weight block through peer loads -- broadcast and pack in one
 * kernel -- falling back to cudaMemcpyPeerAsync where the evaluation needs the FP64 block itself).
 * Host buffers in, host buffers out (each device copies its vertex range), results in the caller's vertex order. */
typedef struct fd_mgpu fd_mgpu;
#define FD_MGPU_AUTO 0 /* p2p when every device reaches the root by peer access, else nccl */
#define FD_MGPU_NCCL 1
#define FD_MGPU_P2P 2
int fd_mgpu_create(fd_mgpu** out, const int* devices, int32_t ndev, int32_t transport);
void fd_mgpu_destroy(fd_mgpu* g);
int fd_mgpu_fit(fd_mgpu* g, const fd_params* params, const float* rest_ctrl, int32_t n_ctrl, fd_report* report);
int fd_mgpu_solve(fd_mgpu* g, const float* deform_ctrl, int32_t n_ctrl, int32_t frames, fd_report* report);
int fd_mgpu_eval(fd_mgpu* g, const float* P, int64_t n_vtx, const float* dist2, const float* tangentu,
                 const float* tangentv, const float* normal, float* P_out, float* falloff_out);
/* ndev, transport in use (FD_MGPU_NCCL / FD_MGPU_P2P), bytes that crossed NVLink in the last solve, and the device
 * time (ms, max over devices) from the root's solve end to the last device's tables being ready */
int fd_mgpu_info(const fd_mgpu* g, int32_t* ndev, int32_t* transport, int64_t* bcast_bytes, float* bcast_ms);
/* vertex range [begin, end) of device slot `i` for n_vtx vertices (the partition fd_mgpu_eval uses) */
int fd_mgpu_range(const fd_mgpu* g, int32_t i, int64_t n_vtx, int64_t* begin, int64_t* end);
fd_ctx* fd_mgpu_ctx(fd_mgpu* g, int32_t i);
const char* fd_mgpu_last_error(const fd_mgpu* g);

/* ---- capture ------------------------------------------------------------------------------------------------
 * mesh P + polygons (CSR) and rig points + primitives (CSR; 2 vertices = segment, >= 3 = fan-triangulated
 * polygon) + optional class attribute.  Outputs (host): nearest_idx[N]; member[V]; dist2[V]; groups as CSR
 * (grp_class[G] ascending, grp_off[G+1], grp_idx[] ascending; grp_idx may be NULL to query sizes).
 * *n_groups receives G; G == 0 returns FD_E_CAPTURE like the reference (capture.cpp:54-56). */
int fd_capture(fd_ctx* ctx, const float* P, int64_t n_vtx, const int32_t* poly_off, const int32_t* poly_vtx,
               int32_t n_poly, const float* rig_P, int32_t n_rig, const int32_t* rig_off, const int32_t* rig_vtx,
               int32_t n_rig_prim, const int32_t* rig_class /* or NULL */, int32_t max_edges, float radius,
               int32_t dofalloff, int32_t* nearest_idx, uint8_t* member, float* dist2, int32_t* n_groups,
               int32_t* grp_class, int64_t* grp_off, int32_t* grp_idx, int32_t grp_cap, int64_t idx_cap);

/* ---- DirectBSEdit: the "morph space" post-pass (dbse.hpp:16-22) ---------------------------------------------------
 * init: shapes matrix M (3P x S) of FP32 (shape - rest) deltas + its Householder QR (dbse.cpp:9-35); `shapes` is
 * S x P x 3 floats.  compute_weights: w_s = sum_i (pos - rest)_i * QR(i, s) with the PACKED QR storage, as the
 * reference does (matrixQR(), dbse.cpp:53-54).  displace: displaceVector for every point + the SOP's write
 * (dbse.cpp:60-75, SOP_FaceDeform.cpp:460-472): P_out = rest + sum_s M_s * clamp((float)(3 w_s)) [+ (pos - rest) *
 * falloffradius when dofalloff and falloffradius != 0].  get_weights fails with FD_E_STATE before compute_weights
 * (getWeights returns false, dbse.cpp:79-81).  get_qr copies the packed QR (3P x S column-major) and tau. */
typedef struct fd_dbse fd_dbse;
int fd_dbse_init(fd_ctx* ctx, const float* rest_P, int64_t n_pts, const float* shapes, int32_t n_shapes, fd_dbse** out);
int fd_dbse_compute_weights(fd_dbse* h, const float* pos, const float* rest, double* weights_out /* S or NULL */);
int fd_dbse_displace(fd_dbse* h, const float* pos, const float* rest, int32_t doclamp, const float* weightrange /* 2 */,
                     int32_t dofalloff, float falloffradius, float* P_out /* P x 3 */);
int fd_dbse_get_weights(fd_dbse* h, double* weights /* S */);
int fd_dbse_get_qr(fd_dbse* h, double* qr, double* tau);
int fd_dbse_info(const fd_dbse* h, int64_t* n_pts, int32_t* n_shapes, int32_t* computed);
void fd_dbse_destroy(fd_dbse* h);

/* timing of the last call on this ctx, measured with CUDA events on the ctx stream (ms); phase: 0 assemble,
 * 1 factor, 2 solve, 3 eval.  Returns < 0 when the phase has not run. */
float fd_ctx_phase_ms(fd_ctx* ctx, int phase);
/* number of kernels this library launched on the ctx since creation */
int64_t fd_ctx_launch_count(const fd_ctx* ctx);

/* ---- C shim over the C++ operator mirror (include/facedeform_sop.hpp: fd::FaceDeformOp = cookMySop,
 * SOP_FaceDeform.cpp:215-489, with ProximityCapture and the data-ID change tracking of SOP_FaceDeform.hpp:47-63).
 * fd_sop_cook returns 0 ok / 1 warnings / 2 errors; the texts are the reference's addError/addWarning strings. */
typedef struct fd_sop fd_sop;
fd_sop* fd_sop_create(int device);
void fd_sop_destroy(fd_sop* s);
fd_params* fd_sop_params(fd_sop* s);
int fd_sop_cook(fd_sop* s, const float* mesh_P, int64_t n_vtx, const int32_t* poly_off, const int32_t* poly_vtx,
                int32_t n_poly, const float* tangentu, const float* tangentv, const float* normal, int64_t mesh_p_id,
                int64_t mesh_topo_id, const float* rest_rig_P, int32_t n_rig, const int32_t* rig_off,
                const int32_t* rig_vtx, int32_t n_rig_prim, const int32_t* rig_class, int64_t rig_p_id,
                int64_t rig_topo_id, const float* deform_rig_P, int32_t n_deform, int32_t frames, float* P_out,
                float* falloff_out);
const char* fd_sop_messages(fd_sop* s, int kind); /* 0 errors, 1 warnings, 2 messages */
int fd_sop_fit_count(const fd_sop* s);
/* 1 when the last cook bumped P's data ID: `!myGroup || !myGroup->isEmpty()` (SOP_FaceDeform.cpp:485-486) */
int fd_sop_positions_bumped(const fd_sop* s);
/* inputs >= 3 of the SOP (blendshapes of the morph-space pass, `morphspace` parm): n_shapes x n_pts x 3 floats, copied */
int fd_sop_set_blendshapes(fd_sop* s, const float* shapes, int32_t n_shapes, int64_t n_pts, int64_t data_id);
/* the "weights" detail attribute of the last cook (SOP_FaceDeform.cpp:474-480); returns their count */
int fd_sop_blend_weights(fd_sop* s, double* weights, int32_t cap);

#ifdef __cplusplus
}
#endif
#endif /* FACEDEFORM_GPU_H */
