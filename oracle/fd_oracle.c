/*
 * fd_oracle.c -- CPU oracle (FP64) for the FaceDeform RBF deformation path.
 * TEST INFRASTRUCTURE ONLY; PARITY UNPINNED -- see fd_oracle.h.
 *
 * Every function cites the reference site it restates (paths under
 * /root/reference/src).  Compile with -ffp-contract=off: the FP32 geometry
 * helpers are defined with a fixed, un-fused operation order so the GPU
 * capture kernels can be bit-exact against them.
 */
#include "fd_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

int fdo_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* defaults of the parameter templates, SOP_FaceDeform.cpp:117-137 */
void fdo_params_default(fdo_params* p)
{
    p->model = FDO_MODEL_QNN;
    p->term = FDO_TERM_LINEAR;
    p->kernel = FDO_KERNEL_GAUSSIAN;
    p->qcoef = 1.0f;
    p->zcoef = 5.0f;
    p->radius = 1.0f;
    p->layers = 4;
    p->lambda = 0.1f;
    p->tangent = 0;
    p->maxedges = 4;
    p->dofalloff = 0;
    p->falloffradius = 1.0f;
    p->falloffrate = 1.0f;
}

/* SOP_FaceDeform.cpp:249-257 (SYSmax clamps) */
void fdo_clamp_params(fdo_params* p)
{
    if (p->qcoef < 0.1f) p->qcoef = 0.1f;
    if (p->zcoef < 0.1f) p->zcoef = 0.1f;
    if (p->radius < 0.01f) p->radius = 0.01f;
    if (p->layers < 1) p->layers = 1;
    if (p->lambda < 0.01f) p->lambda = 0.01f;
    if (p->maxedges < 1) p->maxedges = 1;
}

/* rbfsetlinterm / rbfsetconstterm / rbfsetzeroterm, SOP_FaceDeform.cpp:351-361 */
int fdo_poly_terms(int term)
{
    return term == FDO_TERM_LINEAR ? 4 : (term == FDO_TERM_CONST ? 1 : 0);
}

/* SOP_FaceDeform.cpp:268-287: delta subtracted in FP32 (UT_Vector3), then widened to FP64 */
void fdo_pack(const float* rest, const float* deform, int N, double* out6)
{
    for (int i = 0; i < N; ++i) {
        for (int k = 0; k < 3; ++k) {
            const float d = deform[3 * i + k] - rest[3 * i + k];
            out6[6 * i + k] = (double)rest[3 * i + k];
            out6[6 * i + 3 + k] = (double)d;
        }
    }
}

static int cmp_double(const void* a, const void* b)
{
    const double x = *(const double*)a, y = *(const double*)b;
    return (x > y) - (x < y);
}

/*
 * rbfsetalgoqnn(model, q, z) / rbfsetalgomultilayer(model, radius, ...), SOP_FaceDeform.cpp:342-349.
 * QNN [recollection]: R_i = q * (distance to the nearest other centre), then R_i = min(R_i, z * median(R)).
 * Our definition of the median: sorted[N/2].  ML: every centre uses `radius`.
 */
int fdo_radii(const fdo_params* p, const float* rest, int N, double* radii)
{
    if (p->model != FDO_MODEL_QNN) {
        for (int i = 0; i < N; ++i) radii[i] = (double)p->radius;
        return 0;
    }
    if (N == 1) {
        radii[0] = (double)p->radius;
        return 0;
    }
#pragma omp parallel for schedule(static)
    for (int i = 0; i < N; ++i) {
        double best = INFINITY;
        const double xi = rest[3 * i], yi = rest[3 * i + 1], zi = rest[3 * i + 2];
        for (int j = 0; j < N; ++j) {
            if (j == i) continue;
            const double dx = xi - rest[3 * j], dy = yi - rest[3 * j + 1], dz = zi - rest[3 * j + 2];
            const double d2 = dx * dx + dy * dy + dz * dz;
            if (d2 < best) best = d2;
        }
        radii[i] = (double)p->qcoef * sqrt(best);
    }
    double* tmp = (double*)malloc(sizeof(double) * (size_t)N);
    memcpy(tmp, radii, sizeof(double) * (size_t)N);
    qsort(tmp, (size_t)N, sizeof(double), cmp_double);
    const double cap = (double)p->zcoef * tmp[N / 2];
    free(tmp);
    int status = 0;
    for (int i = 0; i < N; ++i) {
        if (radii[i] > cap) radii[i] = cap;
        if (!(radii[i] > 0.0)) status = -5;
    }
    return status;
}

/* basis function of centre j at squared distance r2 (FP64) */
static inline double fdo_phi(int kernel, double r2, double R)
{
    switch (kernel) {
    case FDO_KERNEL_GAUSSIAN:
        return exp(-r2 / (R * R));
    case FDO_KERNEL_MULTIQUADRIC:
        return sqrt(r2 + R * R);
    default: /* thin plate: r^2 log r = 0.5 r^2 log r^2 */
        return r2 > 0.0 ? 0.5 * r2 * log(r2) : 0.0;
    }
}

/*
 * Dense restatement of the system rbfbuildmodel solves (SOP_FaceDeform.cpp:363), in the saddle-point form
 * north_star names:  [[K + lambda I, P], [P^T, 0]] [w; a] = [delta; 0],  K_ij = phi_j(|c_i - c_j|),
 * P_i = [1, x, y, z] (linear) / [1] (const) / nothing (zero).
 */
void fdo_assemble(const fdo_params* p, const float* rest, const double* radii, int N, double* A)
{
    const int np = fdo_poly_terms(p->term);
    const int n = N + np;
    const double lambda = (double)p->lambda;
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n; ++i) {
        double* row = A + (size_t)i * n;
        if (i < N) {
            const double xi = rest[3 * i], yi = rest[3 * i + 1], zi = rest[3 * i + 2];
            for (int j = 0; j < N; ++j) {
                const double dx = xi - rest[3 * j], dy = yi - rest[3 * j + 1], dz = zi - rest[3 * j + 2];
                row[j] = fdo_phi(p->kernel, dx * dx + dy * dy + dz * dz, radii[j]);
            }
            /* smoothing: the multiquadric +sqrt(r^2 + R^2) is conditionally NEGATIVE definite, so its diagonal shift
             * carries the opposite sign (the same system scipy solves with kernel -sqrt(1 + (r/R)^2), smoothing lambda/R) */
            row[i] += p->kernel == FDO_KERNEL_MULTIQUADRIC ? -lambda : lambda;
            if (np >= 1) row[N] = 1.0;
            if (np == 4) {
                row[N + 1] = xi;
                row[N + 2] = yi;
                row[N + 3] = zi;
            }
        } else {
            const int k = i - N;
            for (int j = 0; j < N; ++j) row[j] = (k == 0) ? 1.0 : (double)rest[3 * j + (k - 1)];
            for (int j = N; j < n; ++j) row[j] = 0.0;
        }
    }
}

int fdo_lu_factor(double* A, int n, int32_t* piv)
{
    for (int k = 0; k < n; ++k) {
        int pr = k;
        double best = fabs(A[(size_t)k * n + k]);
        for (int i = k + 1; i < n; ++i) {
            const double v = fabs(A[(size_t)i * n + k]);
            if (v > best) {
                best = v;
                pr = i;
            }
        }
        piv[k] = pr;
        if (best == 0.0 || best != best) return k + 1;
        if (pr != k) {
            double* a = A + (size_t)k * n;
            double* b = A + (size_t)pr * n;
            for (int j = 0; j < n; ++j) {
                const double t = a[j];
                a[j] = b[j];
                b[j] = t;
            }
        }
        const double inv = 1.0 / A[(size_t)k * n + k];
        const double* rk = A + (size_t)k * n;
#pragma omp parallel for schedule(static) if (n - k > 256)
        for (int i = k + 1; i < n; ++i) {
            double* ri = A + (size_t)i * n;
            const double l = ri[k] * inv;
            ri[k] = l;
            for (int j = k + 1; j < n; ++j) ri[j] -= l * rk[j];
        }
    }
    return 0;
}

void fdo_lu_solve(const double* LU, const int32_t* piv, int n, double* B, int nrhs)
{
    for (int k = 0; k < n; ++k) {
        if (piv[k] != k) {
            double* a = B + (size_t)k * nrhs;
            double* b = B + (size_t)piv[k] * nrhs;
            for (int j = 0; j < nrhs; ++j) {
                const double t = a[j];
                a[j] = b[j];
                b[j] = t;
            }
        }
    }
    for (int k = 0; k < n; ++k) { /* forward, unit lower */
        const double* bk = B + (size_t)k * nrhs;
#pragma omp parallel for schedule(static) if ((size_t)(n - k) * nrhs > 65536)
        for (int i = k + 1; i < n; ++i) {
            const double l = LU[(size_t)i * n + k];
            double* bi = B + (size_t)i * nrhs;
            for (int j = 0; j < nrhs; ++j) bi[j] -= l * bk[j];
        }
    }
    for (int k = n - 1; k >= 0; --k) { /* backward */
        double* bk = B + (size_t)k * nrhs;
        const double inv = 1.0 / LU[(size_t)k * n + k];
        for (int j = 0; j < nrhs; ++j) bk[j] *= inv;
#pragma omp parallel for schedule(static) if ((size_t)k * nrhs > 65536)
        for (int i = 0; i < k; ++i) {
            const double u = LU[(size_t)i * n + k];
            double* bi = B + (size_t)i * nrhs;
            for (int j = 0; j < nrhs; ++j) bi[j] -= u * bk[j];
        }
    }
}

/* rbfcreate + rbfsetpoints + rbfbuildmodel for F frames, SOP_FaceDeform.cpp:331-368 */
int fdo_fit(const fdo_params* p, const float* rest, const float* deform, int N, int F,
            double* radii_out, double* weights_out)
{
    const int np = fdo_poly_terms(p->term);
    const int n = N + np;
    const int nrhs = 3 * F;
    if (fdo_radii(p, rest, N, radii_out) != 0) return -5;
    double* A = (double*)malloc(sizeof(double) * (size_t)n * n);
    int32_t* piv = (int32_t*)malloc(sizeof(int32_t) * (size_t)n);
    fdo_assemble(p, rest, radii_out, N, A);
    if (fdo_lu_factor(A, n, piv) != 0) {
        free(A);
        free(piv);
        return -3;
    }
    /* right-hand sides: delta subtracted in FP32 then widened (SOP_FaceDeform.cpp:276-284) */
    for (int i = 0; i < N; ++i)
        for (int f = 0; f < F; ++f)
            for (int k = 0; k < 3; ++k) {
                const float d = deform[((size_t)f * N + i) * 3 + k] - rest[3 * i + k];
                weights_out[(size_t)i * nrhs + 3 * f + k] = (double)d;
            }
    for (int i = N; i < n; ++i)
        for (int j = 0; j < nrhs; ++j) weights_out[(size_t)i * nrhs + j] = 0.0;
    fdo_lu_solve(A, piv, n, weights_out, nrhs);
    free(A);
    free(piv);
    for (size_t i = 0; i < (size_t)n * nrhs; ++i)
        if (weights_out[i] != weights_out[i] || isinf(weights_out[i])) return -3;
    return 1;
}

/* rbfcalc (SOP_FaceDeform.cpp:414) in the dense formulation, one vertex, all 3F outputs */
static void fdo_calc(const fdo_params* p, const float* rest, const double* radii, const double* weights,
                     int N, int nrhs, const double x[3], double* y)
{
    const int np = fdo_poly_terms(p->term);
    for (int c = 0; c < nrhs; ++c) y[c] = 0.0;
    for (int j = 0; j < N; ++j) {
        const double dx = x[0] - rest[3 * j], dy = x[1] - rest[3 * j + 1], dz = x[2] - rest[3 * j + 2];
        const double phi = fdo_phi(p->kernel, dx * dx + dy * dy + dz * dz, radii[j]);
        const double* w = weights + (size_t)j * nrhs;
        for (int c = 0; c < nrhs; ++c) y[c] += w[c] * phi;
    }
    if (np >= 1) {
        const double* a0 = weights + (size_t)N * nrhs;
        for (int c = 0; c < nrhs; ++c) y[c] += a0[c];
    }
    if (np == 4) {
        for (int k = 0; k < 3; ++k) {
            const double* a = weights + (size_t)(N + 1 + k) * nrhs;
            for (int c = 0; c < nrhs; ++c) y[c] += a[c] * x[k];
        }
    }
}

void fdo_eval_raw(const fdo_params* p, const float* rest, const double* radii, const double* weights,
                  int N, int F, const float* P, int64_t V, double* out, int nthreads)
{
    const int nrhs = 3 * F;
    (void)nthreads;
#pragma omp parallel for schedule(static) num_threads(nthreads > 1 ? nthreads : 1)
    for (int64_t v = 0; v < V; ++v) {
        const double x[3] = {P[3 * v], P[3 * v + 1], P[3 * v + 2]};
        fdo_calc(p, rest, radii, weights, N, nrhs, x, out + (size_t)v * nrhs);
    }
}

static inline void fdo_normalize3(float a[3])
{
    /* UT_Vector3::normalize(): scale by 1/length when the length is non-zero */
    const float len = sqrtf(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]);
    if (len > 0.0f) {
        const float inv = 1.0f / len;
        a[0] *= inv;
        a[1] *= inv;
        a[2] *= inv;
    }
}

/*
 * SOP_FaceDeform.hpp:28-41.  M has rows (u, v, n); B = M^T * M; a1 = normalize(u * B), a2 = normalize(v * B)
 * with the row-vector-times-matrix convention of UT_Vector3 * UT_Matrix3; disp = a1 (disp.a1) + a2 (disp.a2).
 */
void fdo_project_to_tangents(const float u[3], const float v[3], const float n[3], float disp[3])
{
    float B[3][3];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) B[i][j] = u[i] * u[j] + v[i] * v[j] + n[i] * n[j];
    float a1[3], a2[3];
    for (int j = 0; j < 3; ++j) {
        a1[j] = u[0] * B[0][j] + u[1] * B[1][j] + u[2] * B[2][j];
        a2[j] = v[0] * B[0][j] + v[1] * B[1][j] + v[2] * B[2][j];
    }
    fdo_normalize3(a1);
    fdo_normalize3(a2);
    const float da1 = disp[0] * a1[0] + disp[1] * a1[1] + disp[2] * a1[2];
    const float da2 = disp[0] * a2[0] + disp[1] * a2[1] + disp[2] * a2[2];
    for (int k = 0; k < 3; ++k) disp[k] = a1[k] * da1 + a2[k] * da2;
}

/*
 * The evaluation loop, SOP_FaceDeform.cpp:404-439, for F frames at once:
 *   d2 = dist_attr (0 when the attribute is invalid)                          :405-407
 *   d2 > R^2  -> vertex skipped, P unchanged (fd_falloff keeps its default 0) :408-410
 *   pos f32 -> f64, rbfcalc, f64 -> f32                                       :411-415
 *   optional tangent projection with u, v, n normalised                       :416-422
 *   falloff = pow(1 - min(d2/R^2, 1), rate); fd_falloff = falloff             :423-425
 *   P = pos + displace * falloff (FP32)                                       :437-438
 */
void fdo_eval(const fdo_params* p, const float* rest, const double* radii, const double* weights,
              int N, int F, const float* P, int64_t V, const float* dist2,
              const float* tu, const float* tv, const float* nrm,
              float* P_out, float* falloff_out, int nthreads)
{
    const int nrhs = 3 * F;
    const float radius_sqrt = p->radius * p->radius; /* :402 */
    const int do_tangent = p->tangent && tu && tv && nrm; /* :293-294 */
    (void)nthreads;
#pragma omp parallel num_threads(nthreads > 1 ? nthreads : 1)
    {
        double* y = (double*)malloc(sizeof(double) * (size_t)nrhs);
#pragma omp for schedule(static)
        for (int64_t v = 0; v < V; ++v) {
            float distance_sqrt = 0.0f;
            if (dist2) distance_sqrt = dist2[v];
            const float pos[3] = {P[3 * v], P[3 * v + 1], P[3 * v + 2]};
            if (distance_sqrt > radius_sqrt) {
                for (int f = 0; f < F; ++f)
                    for (int k = 0; k < 3; ++k) P_out[((size_t)f * V + v) * 3 + k] = pos[k];
                if (falloff_out) falloff_out[v] = 0.0f;
                continue;
            }
            const double x[3] = {pos[0], pos[1], pos[2]};
            fdo_calc(p, rest, radii, weights, N, nrhs, x, y);
            float falloff = distance_sqrt / radius_sqrt;
            if (falloff > 1.0f) falloff = 1.0f;
            falloff = powf(1.0f - falloff, p->falloffrate);
            if (falloff_out) falloff_out[v] = falloff;
            float u[3], w[3], n[3];
            if (do_tangent) {
                for (int k = 0; k < 3; ++k) {
                    u[k] = tu[3 * v + k];
                    w[k] = tv[3 * v + k];
                    n[k] = nrm[3 * v + k];
                }
                fdo_normalize3(u);
                fdo_normalize3(w);
                fdo_normalize3(n);
            }
            for (int f = 0; f < F; ++f) {
                float disp[3] = {(float)y[3 * f], (float)y[3 * f + 1], (float)y[3 * f + 2]};
                if (do_tangent) fdo_project_to_tangents(u, w, n, disp);
                for (int k = 0; k < 3; ++k) {
                    const float d = disp[k] * falloff;
                    P_out[((size_t)f * V + v) * 3 + k] = pos[k] + d;
                }
            }
        }
        free(y);
    }
}

/* ---------------------------------------------------------------------------------------------------
 * ProximityCapture restatement (capture.cpp).  FP32 geometry with a fixed, un-fused operation order.
 * ------------------------------------------------------------------------------------------------- */

static inline float dot3(const float a[3], const float b[3])
{
    return (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2];
}

static inline float dist2_3(const float a[3], const float b[3])
{
    const float d[3] = {a[0] - b[0], a[1] - b[1], a[2] - b[2]};
    return dot3(d, d);
}

float fdo_point_seg_dist2(const float p[3], const float a[3], const float b[3])
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

/* closest point on a triangle by Voronoi-region classification (Ericson, Real-Time Collision Detection 5.1.5) */
float fdo_point_tri_dist2(const float p[3], const float a[3], const float b[3], const float c[3])
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

/* GEO_PointTree::findNearestIdx (capture.cpp:122): exact nearest, our tie-break = lowest index */
static int32_t nearest_point(const float* P, int64_t V, const float q[3])
{
    int32_t best = -1;
    float bd = INFINITY;
    for (int64_t v = 0; v < V; ++v) {
        const float d = dist2_3(q, P + 3 * v);
        if (d < bd) {
            bd = d;
            best = (int32_t)v;
        }
    }
    return best;
}

static int cmp_i64(const void* a, const void* b)
{
    const int64_t x = *(const int64_t*)a, y = *(const int64_t*)b;
    return (x > y) - (x < y);
}

static int cmp_i32(const void* a, const void* b)
{
    const int32_t x = *(const int32_t*)a, y = *(const int32_t*)b;
    return (x > y) - (x < y);
}

int fdo_capture(const float* P, int64_t V, const int32_t* poly_off, const int32_t* poly_vtx, int32_t npoly,
                const float* rigP, int32_t N, const int32_t* rig_off, const int32_t* rig_vtx, int32_t nrigprim,
                const int32_t* rig_class, int32_t max_edges, float radius, int32_t dofalloff,
                int32_t* nearest_idx, uint8_t* member, float* dist2,
                int32_t* grp_class, int64_t* grp_off, int32_t* grp_idx, int32_t grp_cap, int64_t idx_cap)
{
    /* --- GQ_Detail edge structure (capture.cpp:24): undirected polygon edges as a CSR adjacency --- */
    int64_t* adj_off = (int64_t*)calloc((size_t)V + 1, sizeof(int64_t));
    for (int32_t f = 0; f < npoly; ++f) {
        const int32_t b = poly_off[f], e = poly_off[f + 1], m = e - b;
        if (m < 2) continue;
        for (int32_t k = 0; k < m; ++k) {
            const int32_t u = poly_vtx[b + k], w = poly_vtx[b + (k + 1) % m];
            if (m == 2 && k == 1) break;
            adj_off[u + 1]++;
            adj_off[w + 1]++;
        }
    }
    for (int64_t v = 0; v < V; ++v) adj_off[v + 1] += adj_off[v];
    int32_t* adj = (int32_t*)malloc(sizeof(int32_t) * (size_t)(adj_off[V] > 0 ? adj_off[V] : 1));
    int64_t* fill = (int64_t*)malloc(sizeof(int64_t) * ((size_t)V + 1));
    memcpy(fill, adj_off, sizeof(int64_t) * ((size_t)V + 1));
    for (int32_t f = 0; f < npoly; ++f) {
        const int32_t b = poly_off[f], e = poly_off[f + 1], m = e - b;
        if (m < 2) continue;
        for (int32_t k = 0; k < m; ++k) {
            const int32_t u = poly_vtx[b + k], w = poly_vtx[b + (k + 1) % m];
            if (m == 2 && k == 1) break;
            adj[fill[u]++] = w;
            adj[fill[w]++] = u;
        }
    }
    free(fill);

    /* --- findIslands (capture.cpp:107-141) --- */
    /* distinct class ids, ascending (HandlerGroupMap is unordered; we fix ascending order) */
    int32_t ngrp = 0;
    int32_t* classes = (int32_t*)malloc(sizeof(int32_t) * ((size_t)N + 1));
    if (!rig_class) {
        classes[ngrp++] = 0; /* capture.cpp:114-118 */
    } else {
        for (int32_t i = 0; i < N; ++i) classes[i] = rig_class[i];
        qsort(classes, (size_t)N, sizeof(int32_t), cmp_i32);
        for (int32_t i = 0; i < N; ++i)
            if (i == 0 || classes[i] != classes[i - 1]) classes[ngrp++] = classes[i];
    }
    if (ngrp > grp_cap) {
        free(classes);
        free(adj);
        free(adj_off);
        return -1;
    }
    /* (group, vertex) pairs collected per rig point, de-duplicated below (GA_PointGroup::combine, :135-137) */
    size_t pair_cap = 1024, npairs = 0;
    int64_t* pairs = (int64_t*)malloc(sizeof(int64_t) * pair_cap);
    int32_t* depth = (int32_t*)malloc(sizeof(int32_t) * (size_t)V);
    int32_t* queue = (int32_t*)malloc(sizeof(int32_t) * (size_t)V);
    for (int64_t v = 0; v < V; ++v) depth[v] = -1;
    memset(member, 0, (size_t)V);
#pragma omp parallel for schedule(dynamic, 1)
    for (int32_t i = 0; i < N; ++i) nearest_idx[i] = nearest_point(P, V, rigP + 3 * i); /* :122 */
    for (int32_t i = 0; i < N; ++i) { /* GA_FOR_ALL_PTOFF(m_rig, ptoff), capture.cpp:120 */
        const int32_t target = nearest_idx[i];
        if (target < 0) continue;
        int32_t g = 0;
        if (rig_class) {
            const int32_t* hit = (const int32_t*)bsearch(&rig_class[i], classes, (size_t)ngrp, sizeof(int32_t), cmp_i32);
            g = (int32_t)(hit - classes);
        }
        /* groupEdgePoints(target, max_edges, partial) (capture.cpp:134): our definition = the seed plus
         * every vertex within max_edges edge hops (breadth-first rings). */
        int32_t head = 0, tail = 0;
        queue[tail++] = target;
        depth[target] = 0;
        while (head < tail) {
            const int32_t u = queue[head++];
            if (depth[u] >= max_edges) continue;
            for (int64_t e = adj_off[u]; e < adj_off[u + 1]; ++e) {
                const int32_t w = adj[e];
                if (depth[w] < 0) {
                    depth[w] = depth[u] + 1;
                    queue[tail++] = w;
                }
            }
        }
        if (npairs + (size_t)tail > pair_cap) {
            while (npairs + (size_t)tail > pair_cap) pair_cap *= 2;
            pairs = (int64_t*)realloc(pairs, sizeof(int64_t) * pair_cap);
        }
        for (int32_t t = 0; t < tail; ++t) {
            const int32_t u = queue[t];
            depth[u] = -1;
            member[u] = 1;
            pairs[npairs++] = ((int64_t)g << 32) | (int64_t)u;
        }
    }
    free(depth);
    free(queue);
    qsort(pairs, npairs, sizeof(int64_t), cmp_i64);

    /* --- capture (capture.cpp:46-105): distance attribute, default 0 (detached attr, :31) --- */
    const float radius_sqrt = radius * radius; /* :62 */
#pragma omp parallel for schedule(dynamic, 1024)
    for (int64_t v = 0; v < V; ++v) {
        if (!member[v] || !dofalloff) { /* :71-75, and ungrouped vertices keep the default */
            dist2[v] = 0.0f;
            continue;
        }
        /* GU_RayIntersect::minimumPoint with GU_MinInfo(R^2) (:77-86): found iff d^2 < R^2 (our definition) */
        float best = radius_sqrt;
        int found = 0;
        const float* p = P + 3 * v;
        for (int32_t f = 0; f < nrigprim; ++f) {
            const int32_t b = rig_off[f], m = rig_off[f + 1] - b;
            if (m == 2) {
                const float d = fdo_point_seg_dist2(p, rigP + 3 * rig_vtx[b], rigP + 3 * rig_vtx[b + 1]);
                if (d < best) {
                    best = d;
                    found = 1;
                }
            } else {
                for (int32_t k = 1; k + 1 < m; ++k) {
                    const float d = fdo_point_tri_dist2(p, rigP + 3 * rig_vtx[b], rigP + 3 * rig_vtx[b + k],
                                                        rigP + 3 * rig_vtx[b + k + 1]);
                    if (d < best) {
                        best = d;
                        found = 1;
                    }
                }
            }
        }
        dist2[v] = found ? best : -1.0f; /* :76, :86-88 */
    }

    /* groups as CSR (ascending class, ascending vertex index) */
    int64_t total = 0;
    size_t k = 0;
    for (int32_t g = 0; g < ngrp; ++g) {
        grp_class[g] = classes[g];
        grp_off[g] = total;
        int64_t prev = -1;
        while (k < npairs && (int32_t)(pairs[k] >> 32) == g) {
            if (pairs[k] != prev) {
                if (grp_idx && total < idx_cap) grp_idx[total] = (int32_t)(pairs[k] & 0xffffffff);
                ++total;
                prev = pairs[k];
            }
            ++k;
        }
    }
    grp_off[ngrp] = total;
    free(pairs);
    free(classes);
    free(adj);
    free(adj_off);
    return ngrp;
}


/* ======================================================================================================
 * DirectBSEdit (reference src/dbse.cpp), see fd_oracle.h
 * ====================================================================================================== */

/* dbse.cpp:9-35 */
void fdo_dbse_shapes_matrix(const float* rest, const float* shapes, int64_t P, int32_t S, double* M)
{
    const int64_t m = 3 * P;
    for (int32_t s = 0; s < S; ++s)
        for (int64_t p = 0; p < P; ++p)
            for (int k = 0; k < 3; ++k) {
                const float d = shapes[((int64_t)s * P + p) * 3 + k] - rest[3 * p + k]; /* UT_Vector3 subtraction, :24 */
                M[(int64_t)s * m + 3 * p + k] = (double)d;                               /* :25-27 */
            }
}

/* Eigen::HouseholderQR<MatrixXd> (dbse.cpp:31), unblocked, LAPACK dgeqr2 conventions */
void fdo_householder_qr(double* A, int64_t m, int32_t n, double* tau)
{
    for (int32_t j = 0; j < n && j < m; ++j) {
        double* x = A + (int64_t)j * m;
        const double alpha = x[j];
        double ss = 0.0;
        for (int64_t i = j + 1; i < m; ++i) ss += x[i] * x[i];
        if (ss == 0.0) { /* tail is zero: H = I (Eigen: tau = 0, beta = c0) */
            tau[j] = 0.0;
            continue;
        }
        double beta = sqrt(alpha * alpha + ss);
        if (alpha >= 0.0) beta = -beta;
        const double t = (beta - alpha) / beta;
        const double scale = 1.0 / (alpha - beta);
        for (int64_t i = j + 1; i < m; ++i) x[i] *= scale;
        x[j] = beta;
        tau[j] = t;
        for (int32_t c = j + 1; c < n; ++c) { /* apply H = I - tau v v^T to the trailing columns */
            double* y = A + (int64_t)c * m;
            double w = y[j];
            for (int64_t i = j + 1; i < m; ++i) w += x[i] * y[i];
            w *= t;
            y[j] -= w;
            for (int64_t i = j + 1; i < m; ++i) y[i] -= w * x[i];
        }
    }
}

/* dbse.cpp:37-58 */
void fdo_dbse_weights(const double* QR, int64_t P, int32_t S, const float* pos, const float* rest, double* weights)
{
    const int64_t m = 3 * P;
    for (int32_t s = 0; s < S; ++s) {
        double w = 0.0;
        for (int64_t i = 0; i < m; ++i) {
            const float d = pos[i] - rest[i]; /* :46-48 */
            w += (double)d * QR[(int64_t)s * m + i]; /* (delta.asDiagonal() * matrixQR()).colwise().sum(), :53-54 */
        }
        weights[s] = w;
    }
}

/* dbse.cpp:60-75 + SOP_FaceDeform.cpp:460-472 */
void fdo_dbse_displace(const double* M, int64_t P, int32_t S, const double* weights, const float* weightrange,
                       int32_t dofalloff, float falloffradius, const float* pos, const float* rest, float* P_out)
{
    const int64_t m = 3 * P;
    for (int64_t p = 0; p < P; ++p) {
        float disp[3] = {0.f, 0.f, 0.f};
        for (int32_t s = 0; s < S; ++s) {
            const float w = (float)(weights[s] * 3); /* :69 "magic number" */
            float cw = w;
            if (weightrange) cw = w < weightrange[0] ? weightrange[0] : (w > weightrange[1] ? weightrange[1] : w); /* SYSclamp :71 */
            for (int k = 0; k < 3; ++k) {
                const float d = (float)M[(int64_t)s * m + 3 * p + k]; /* :65-67 */
                const float prod = d * cw;
                disp[k] = disp[k] + prod; /* :72 */
            }
        }
        for (int k = 0; k < 3; ++k) {
            if (dofalloff && falloffradius != 0.f) { /* SOP_FaceDeform.cpp:467-470 */
                const float delta = pos[3 * p + k] - rest[3 * p + k];
                const float prod = delta * falloffradius;
                disp[k] = disp[k] + prod;
            }
            P_out[3 * p + k] = rest[3 * p + k] + disp[k]; /* :471 */
        }
    }
}


/* ======================================================================================================
 * "ALGLIB v1 like" fit, see fd_oracle.h (unverifiable: documentation of SOP_FaceDeform.cpp:342-361)
 * ====================================================================================================== */
int fdo_fit_v1(const fdo_params* p, const float* rest, const float* deform, int N, int F,
               float* centres_out, double* radii_out, double* weights_out)
{
    const int np = fdo_poly_terms(p->term);
    const int nrhs = 3 * F;
    const int L = p->model == FDO_MODEL_QNN ? 1 : (p->layers < 1 ? 1 : p->layers);
    const int NL = N * L;
    int status = 1;
    double* R = (double*)malloc(sizeof(double) * (size_t)N * nrhs);   /* residual */
    double* A = (double*)malloc(sizeof(double) * (size_t)N * N);
    double* K0 = (double*)malloc(sizeof(double) * (size_t)N * N);
    double* Wk = (double*)malloc(sizeof(double) * (size_t)N * nrhs);
    int32_t* piv = (int32_t*)malloc(sizeof(int32_t) * (size_t)N);
    double* rad = (double*)malloc(sizeof(double) * (size_t)N);
    /* deltas: FP32 subtract, widened (SOP_FaceDeform.cpp:276-284) */
    for (int i = 0; i < N; ++i)
        for (int f = 0; f < F; ++f)
            for (int k = 0; k < 3; ++k)
                R[(size_t)i * nrhs + 3 * f + k] = (double)(deform[((size_t)f * N + i) * 3 + k] - rest[3 * i + k]);
    /* polynomial term first: normal equations of [1 x y z] (np = 4), the mean (np = 1) or nothing */
    double* v = weights_out + (size_t)NL * nrhs;
    if (np > 0) {
        double G[16], piv4[4];
        (void)piv4;
        for (int a = 0; a < np; ++a)
            for (int b = 0; b < np; ++b) {
                double s = 0.0;
                for (int i = 0; i < N; ++i) {
                    const double pa = a == 0 ? 1.0 : (double)rest[3 * i + a - 1], pb = b == 0 ? 1.0 : (double)rest[3 * i + b - 1];
                    s += pa * pb;
                }
                G[a * np + b] = s;
            }
        for (int c = 0; c < nrhs; ++c) {
            double M[16], y[4];
            for (int a = 0; a < np * np; ++a) M[a] = G[a];
            for (int a = 0; a < np; ++a) {
                double s = 0.0;
                for (int i = 0; i < N; ++i) s += (a == 0 ? 1.0 : (double)rest[3 * i + a - 1]) * R[(size_t)i * nrhs + c];
                y[a] = s;
            }
            /* Gaussian elimination without pivoting on the symmetric positive (semi)definite Gram matrix */
            for (int k = 0; k < np; ++k) {
                const double d = M[k * np + k];
                if (!(d > 0.0)) { status = -3; break; }
                for (int r = k + 1; r < np; ++r) {
                    const double l = M[r * np + k] / d;
                    for (int q = k; q < np; ++q) M[r * np + q] -= l * M[k * np + q];
                    y[r] -= l * y[k];
                }
            }
            if (status != 1) break;
            for (int k = np - 1; k >= 0; --k) {
                double s = y[k];
                for (int q = k + 1; q < np; ++q) s -= M[k * np + q] * y[q];
                y[k] = s / M[k * np + k];
            }
            for (int a = 0; a < np; ++a) v[(size_t)a * nrhs + c] = y[a];
        }
        if (status == 1)
            for (int i = 0; i < N; ++i)
                for (int c = 0; c < nrhs; ++c) {
                    double t = v[c];
                    if (np == 4)
                        for (int k = 0; k < 3; ++k) t += v[(size_t)(1 + k) * nrhs + c] * (double)rest[3 * i + k];
                    R[(size_t)i * nrhs + c] -= t;
                }
    }
    fdo_params q = *p;
    q.term = FDO_TERM_ZERO; /* the layers carry no side conditions */
    for (int k = 0; k < L && status == 1; ++k) {
        if (p->model == FDO_MODEL_QNN) {
            if (fdo_radii(p, rest, N, rad) != 0) { status = -5; break; }
        } else {
            const double Rk = (double)p->radius / (double)(1 << k);
            for (int i = 0; i < N; ++i) rad[i] = Rk;
        }
        for (int i = 0; i < N; ++i) {
            radii_out[(size_t)k * N + i] = rad[i];
            for (int c = 0; c < 3; ++c) centres_out[((size_t)k * N + i) * 3 + c] = rest[3 * i + c];
        }
        q.lambda = 0.0f;
        fdo_assemble(&q, rest, rad, N, K0);           /* the layer's kernel matrix */
        for (size_t t = 0; t < (size_t)N * N; ++t) A[t] = K0[t];
        for (int i = 0; i < N; ++i) A[(size_t)i * N + i] += (double)p->lambda;
        if (fdo_lu_factor(A, N, piv) != 0) { status = -3; break; }
        for (size_t t = 0; t < (size_t)N * nrhs; ++t) Wk[t] = R[t];
        fdo_lu_solve(A, piv, N, Wk, nrhs);
        for (size_t t = 0; t < (size_t)N * nrhs; ++t) weights_out[(size_t)k * N * nrhs + t] = Wk[t];
        for (int i = 0; i < N; ++i)                    /* residual -= K_k w_k */
            for (int j = 0; j < N; ++j) {
                const double kij = K0[(size_t)i * N + j];
                const double* w = Wk + (size_t)j * nrhs;
                double* r = R + (size_t)i * nrhs;
                for (int c = 0; c < nrhs; ++c) r[c] -= kij * w[c];
            }
    }
    if (status == 1)
        for (size_t t = 0; t < (size_t)(NL + np) * nrhs; ++t)
            if (weights_out[t] != weights_out[t] || isinf(weights_out[t])) { status = -3; break; }
    free(R); free(A); free(K0); free(Wk); free(piv); free(rad);
    return status;
}
