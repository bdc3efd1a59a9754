/*
 * fd_oracle.h -- CPU oracle for the FaceDeform RBF deformation path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load it, and only as the checker or the timed
 * CPU baseline.  The product path (libfacedeform_gpu.so) never calls it.
 *
 * PARITY UNPINNED: the reference (symek/facedeform) ships no tests, golden
 * vectors or fixtures, and its arithmetic lives in un-vendored, un-pinned
 * third-party code (ALGLIB rbf unit, Houdini HDK, Eigen) that is absent from
 * /root/reference.  This file restates the reference's call sites
 * (src/SOP_FaceDeform.cpp, src/SOP_FaceDeform.hpp, src/capture.cpp) around the
 * dense RBF formulation named by BASELINE.json's north_star; the math is
 * pinned against scipy.interpolate.RBFInterpolator and analytic properties
 * (tests/test_oracle_*.py), not against reference outputs.
 */
#ifndef FD_ORACLE_H
#define FD_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* menu indices: SOP_FaceDeform.hpp:13-18 */
#define FDO_MODEL_QNN 0
#define FDO_MODEL_ML 1
#define FDO_TERM_LINEAR 0
#define FDO_TERM_CONST 1
#define FDO_TERM_ZERO 2
/* north_star "kernel type" extension (the reference is Gaussian only) */
#define FDO_KERNEL_GAUSSIAN 0
#define FDO_KERNEL_MULTIQUADRIC 1
#define FDO_KERNEL_THINPLATE 2

typedef struct fdo_params {
    int32_t model;   /* SOP_FaceDeform.cpp:48-53 */
    int32_t term;    /* :55-61 */
    int32_t kernel;  /* extension */
    float qcoef;     /* :123, clamp :249 */
    float zcoef;     /* :124, clamp :250 */
    float radius;    /* :125, clamp :251 */
    int32_t layers;  /* :126, clamp :252 (carried; layers>1 is not part of the dense formulation) */
    float lambda;    /* :128, clamp :253 (only fdo_clamp_params applies the clamp) */
    int32_t tangent; /* :129 */
    int32_t maxedges;/* :127, clamp :257 */
    int32_t dofalloff;     /* :133 */
    float falloffradius;   /* :134 */
    float falloffrate;     /* :135 */
} fdo_params;

void fdo_params_default(fdo_params* p);
/* parameter clamps of cookMySop, SOP_FaceDeform.cpp:249-257 */
void fdo_clamp_params(fdo_params* p);

int fdo_poly_terms(int term); /* 4 / 1 / 0 */

/* a1: pack loop SOP_FaceDeform.cpp:268-287. out is N x 6 doubles. */
void fdo_pack(const float* rest, const float* deform, int N, double* out6);

/* a3: per-centre radii. QNN rule [recollection, SURVEY appendix B] or the uniform `radius`.
 * returns 0, or -5 if a radius is zero (duplicate centres). */
int fdo_radii(const fdo_params* p, const float* rest, int N, double* radii);

/* a6 (dense restatement): assemble the (N+p)^2 system, row-major doubles. */
void fdo_assemble(const fdo_params* p, const float* rest, const double* radii, int N, double* A);

/* pivoted LU (partial pivoting, row-major, in place). returns 0 ok, k+1 if zero pivot at column k. */
int fdo_lu_factor(double* A, int n, int32_t* piv);
/* B is n x nrhs row-major, overwritten with the solution */
void fdo_lu_solve(const double* LU, const int32_t* piv, int n, double* B, int nrhs);

/* a1+a2+a6: fit F frames at once. deform is F x N x 3 floats. weights out: (N+p) x 3F row-major doubles.
 * returns a terminationtype-like status (SOP_FaceDeform.cpp:365): 1 ok, -5 zero radius, -3 singular. */
int fdo_fit(const fdo_params* p, const float* rest, const float* deform, int N, int F,
            double* radii_out /* N */, double* weights_out /* (N+p) x 3F */);

/* a7..a10: evaluation loop SOP_FaceDeform.cpp:404-439 (+ SOP_FaceDeform.hpp:28-41).
 * P V x 3; dist2 V or NULL; tu/tv/nrm V x 3 or NULL (all three are needed for the tangent projection);
 * P_out F x V x 3; falloff_out V or NULL. nthreads<=1 => serial like the reference (NO_RBF_THREADS). */
void fdo_eval(const fdo_params* p, const float* rest, const double* radii, const double* weights,
              int N, int F, const float* P, int64_t V, const float* dist2,
              const float* tu, const float* tv, const float* nrm,
              float* P_out, float* falloff_out, int nthreads);

/* raw RBF value in double (no epilogue): out V x 3F doubles. */
void fdo_eval_raw(const fdo_params* p, const float* rest, const double* radii, const double* weights,
                  int N, int F, const float* P, int64_t V, double* out, int nthreads);

/* project_to_tangents, SOP_FaceDeform.hpp:28-41 (FP32, row-vector convention); u,v,n normalised by the caller. */
void fdo_project_to_tangents(const float u[3], const float v[3], const float n[3], float disp[3]);

/* a11: ProximityCapture (capture.cpp:46-141) on plain arrays.
 * mesh: P V x 3 and polygons in CSR form (poly_off[npoly+1], poly_vtx).  rig: points N x 3, primitives in CSR
 * form (2 vertices = segment, >=3 = polygon, fan-triangulated), rig_class[N] or NULL.
 * outputs: nearest_idx[N] (capture.cpp:122); member[V] (1 if the vertex is in any handle group);
 * dist2[V]: 0 for ungrouped vertices or !dofalloff, the closest squared distance to the rig primitives when
 * it is < radius^2, else -1 (capture.cpp:71-88);  groups as CSR: grp_class[G] ascending, grp_off[G+1],
 * grp_idx[] ascending vertex indices (skipped when grp_idx == NULL or idx_cap is too small).
 * returns the number of groups G (0 => capture fails, capture.cpp:54-56), or -1 when grp_cap is too small. */
int fdo_capture(const float* P, int64_t V, const int32_t* poly_off, const int32_t* poly_vtx, int32_t npoly,
                const float* rigP, int32_t N, const int32_t* rig_off, const int32_t* rig_vtx, int32_t nrigprim,
                const int32_t* rig_class, int32_t max_edges, float radius, int32_t dofalloff,
                int32_t* nearest_idx, uint8_t* member, float* dist2,
                int32_t* grp_class, int64_t* grp_off, int32_t* grp_idx, int32_t grp_cap, int64_t idx_cap);

/* closest squared distances in FP32 with a fixed operation order (the GPU kernel states the same order). */
float fdo_point_tri_dist2(const float p[3], const float a[3], const float b[3], const float c[3]);
float fdo_point_seg_dist2(const float p[3], const float a[3], const float b[3]);

/* ---- "ALGLIB v1 like" fit -- UNVERIFIABLE here: ALGLIB is absent and unpinned,
 * this documents what the SOP's two algorithm settings mean (SOP_FaceDeform.cpp:342-361), with dense LU in place of
 * ALGLIB's sparse LSQR [recollection]:
 *   polynomial term first (rbfsetlinterm / constterm / zeroterm, :351-361): least squares of the deltas on [1 x y z]
 *   (or their mean / nothing), the RBF part is fitted to the de-trended residual without side conditions;
 *   model QNN (:344): one Gaussian layer with per-centre radii R_j (fdo_radii);
 *   model Multilayer (:347): `layers` successive Gaussian layers, radius R, R/2, R/4 ..., each solving
 *   (K_k + lambda I) w_k = residual and passing residual - K_k w_k on (lambda as a diagonal shift: this repo's
 *   definition of the damping).
 * Outputs describe one stacked model the ordinary evaluation takes: centres (N * L x 3, `rest` repeated per layer),
 * radii (N * L), weights ((N * L + npoly) x 3F, polynomial rows last).  L = layers for Multilayer, 1 for QNN.
 * Returns 1, -5 (zero QNN radius) or -3 (singular). */
int fdo_fit_v1(const fdo_params* p, const float* rest, const float* deform, int N, int F,
               float* centres_out, double* radii_out, double* weights_out);

/* ---- DirectBSEdit, the "morph space" post-pass (reference src/dbse.cpp, caller SOP_FaceDeform.cpp:444-482) ----------
 * Eigen (absent, unpinned) supplies HouseholderQR there; restated here as the unblocked Householder QR with Eigen's /
 * LAPACK dgeqr2's conventions [recollection for Eigen; pinned against LAPACK through scipy.linalg.qr(mode="raw")]:
 * beta = -sign(alpha) |x|, v = x / (alpha - beta) with v0 = 1 implied, tau = (beta - alpha) / beta; packed storage =
 * R in the upper triangle, the essential parts of v below the diagonal (what Eigen's matrixQR() returns). */
/* dbse.cpp:9-35: M (3P x S, column-major doubles) = shape - rest, subtracted in FP32 (UT_Vector3) and widened */
void fdo_dbse_shapes_matrix(const float* rest, const float* shapes /* S x P x 3 */, int64_t P, int32_t S, double* M);
/* in place: A (m x n column-major, lda = m) -> packed QR; tau[n] */
void fdo_householder_qr(double* A, int64_t m, int32_t n, double* tau);
/* dbse.cpp:37-58: delta = pos - rest (FP32 subtract, widened); weights[s] = sum_i delta_i * QR(i, s) */
void fdo_dbse_weights(const double* QR, int64_t P, int32_t S, const float* pos, const float* rest, double* weights);
/* dbse.cpp:60-75 for every point + the SOP's write SOP_FaceDeform.cpp:460-472 (all FP32, un-fused, columns in order):
 * disp = sum_s (float)M[3p..3p+2][s] * clamp((float)(w_s * 3)); disp += (pos - rest) * falloffradius when dofalloff
 * and falloffradius != 0; P_out = rest + disp.  weightrange == NULL: no clamping. */
void fdo_dbse_displace(const double* M, int64_t P, int32_t S, const double* weights, const float* weightrange,
                       int32_t dofalloff, float falloffradius, const float* pos, const float* rest, float* P_out);

int fdo_num_threads(void);

#ifdef __cplusplus
}
#endif
#endif
