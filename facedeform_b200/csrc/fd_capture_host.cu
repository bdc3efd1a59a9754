// fd_capture_host.cu -- fd_capture: ProximityCapture on plain arrays (reference capture.cpp:10-141).
//
// Integer part on the host, float part on the GPU:
//   init         (capture.cpp:10-44)   edge structure (GQ_Detail) = CSR adjacency from the polygon list
//   findIslands  (capture.cpp:107-141) nearest mesh point per rig point [GPU, k_nearest], then a breadth-first
//                                      ring walk of depth <= max_edges per rig point, united per `class`
//   capture      (capture.cpp:46-105)  closest distance to the rig primitives per grouped vertex [GPU]
// Definitions where the HDK leaves them open (documented in DESIGN.md): nearest ties -> lowest index; a
// group holds the seed and every vertex within max_edges edge hops; found iff d^2 < R^2; groups are reported
// in ascending class order with ascending vertex indices.
#include <algorithm>
#include <vector>

#include "fd_internal.h"

extern "C" int fd_capture(fd_ctx* ctx, const float* P, int64_t n_vtx, const int32_t* poly_off, const int32_t* poly_vtx,
                          int32_t n_poly, const float* rig_P, int32_t n_rig, const int32_t* rig_off,
                          const int32_t* rig_vtx, int32_t n_rig_prim, const int32_t* rig_class, int32_t max_edges,
                          float radius, int32_t dofalloff, int32_t* nearest_idx, uint8_t* member, float* dist2,
                          int32_t* n_groups, int32_t* grp_class, int64_t* grp_off, int32_t* grp_idx, int32_t grp_cap,
                          int64_t idx_cap)
{
    if (!ctx || !n_groups || !member || !dist2 || n_vtx < 0 || n_rig < 0 || (n_vtx > 0 && !P) ||
        (n_rig > 0 && (!rig_P || !nearest_idx)) || (n_poly > 0 && (!poly_off || !poly_vtx)) ||
        (n_rig_prim > 0 && (!rig_off || !rig_vtx)) || !grp_class || !grp_off)
        return FD_E_INVALID;
    if (n_vtx > 0x7fffffff) { snprintf(ctx->err, sizeof(ctx->err), "capture: more than 2^31 vertices"); return FD_E_UNSUPPORTED; }
    const int64_t V = n_vtx;
    const int N = n_rig;
    if (max_edges < 1) max_edges = 1; // SOP_FaceDeform.cpp:257
    // The topology comes from the caller: every offset array must be non-decreasing from 0 and every vertex index in
    // range before anything is indexed with it (host adjacency lists below, rig[] on the device).
    auto bad = [&](const char* what, long long where, long long value) {
        snprintf(ctx->err, sizeof(ctx->err), "capture: %s[%lld] = %lld is out of range", what, where, value);
        return FD_E_INVALID;
    };
    if (n_poly < 0 || n_rig_prim < 0 || grp_cap < 0 || idx_cap < 0) return bad("count", 0, -1);
    if (n_poly > 0 && poly_off[0] != 0) return bad("poly_off", 0, poly_off[0]);
    for (int32_t f = 0; f < n_poly; ++f)
        if (poly_off[f + 1] < poly_off[f]) return bad("poly_off", f + 1, poly_off[f + 1]);
    for (int64_t t = 0, e = n_poly > 0 ? poly_off[n_poly] : 0; t < e; ++t)
        if (poly_vtx[t] < 0 || poly_vtx[t] >= V) return bad("poly_vtx", t, poly_vtx[t]);
    if (n_rig_prim > 0 && rig_off[0] != 0) return bad("rig_off", 0, rig_off[0]);
    for (int32_t f = 0; f < n_rig_prim; ++f)
        if (rig_off[f + 1] < rig_off[f]) return bad("rig_off", f + 1, rig_off[f + 1]);
    for (int64_t t = 0, e = n_rig_prim > 0 ? rig_off[n_rig_prim] : 0; t < e; ++t)
        if (rig_vtx[t] < 0 || rig_vtx[t] >= N) return bad("rig_vtx", t, rig_vtx[t]);
    fd_device_guard guard(ctx->device); // restores the caller's device on every return below
    cudaStream_t s = ctx->stream;
    int st = FD_OK;

    // ---- device copies: mesh points, rig points --------------------------------------------------------------
    void *dP = nullptr, *dRig = nullptr, *dKeys = nullptr, *dNear = nullptr;
    st = fd_stage(ctx, FD_STAGE_P, (size_t)V * 12, &dP);
    if (st == FD_OK) st = fd_stage(ctx, FD_STAGE_TU, (size_t)N * 12, &dRig);
    if (st == FD_OK) st = fd_stage(ctx, FD_STAGE_MISC, (size_t)N * 8, &dKeys);
    if (st == FD_OK) st = fd_stage(ctx, FD_STAGE_TV, (size_t)N * 4, &dNear);
    if (st != FD_OK) return st;
    if (V > 0) FD_CUDA_OK(ctx, cudaMemcpyAsync(dP, P, (size_t)V * 12, cudaMemcpyHostToDevice, s));
    if (N > 0) FD_CUDA_OK(ctx, cudaMemcpyAsync(dRig, rig_P, (size_t)N * 12, cudaMemcpyHostToDevice, s));

    // ---- findIslands: nearest mesh point per rig point on the GPU (capture.cpp:122) ---------------------------
    if (N > 0) {
        FD_CUDA_OK(ctx, fd_launch_nearest(ctx, (const float*)dP, V, (const float*)dRig, N, (unsigned long long*)dKeys,
                                          (int32_t*)dNear));
        FD_CUDA_OK(ctx, cudaMemcpyAsync(nearest_idx, dNear, (size_t)N * 4, cudaMemcpyDeviceToHost, s));
    }

    // ---- init: edge structure while the GPU searches (capture.cpp:24) -----------------------------------------
    std::vector<int64_t> adj_off((size_t)V + 1, 0);
    for (int32_t f = 0; f < n_poly; ++f) {
        const int32_t b = poly_off[f], m = poly_off[f + 1] - b;
        if (m < 2) continue;
        for (int32_t k = 0; k < m; ++k) {
            if (m == 2 && k == 1) break;
            adj_off[(size_t)poly_vtx[b + k] + 1]++;
            adj_off[(size_t)poly_vtx[b + (k + 1) % m] + 1]++;
        }
    }
    for (int64_t v = 0; v < V; ++v) adj_off[v + 1] += adj_off[v];
    std::vector<int32_t> adj((size_t)adj_off[V]);
    {
        std::vector<int64_t> fill(adj_off.begin(), adj_off.end() - 1);
        for (int32_t f = 0; f < n_poly; ++f) {
            const int32_t b = poly_off[f], m = poly_off[f + 1] - b;
            if (m < 2) continue;
            for (int32_t k = 0; k < m; ++k) {
                if (m == 2 && k == 1) break;
                const int32_t u = poly_vtx[b + k], w = poly_vtx[b + (k + 1) % m];
                adj[(size_t)fill[u]++] = w;
                adj[(size_t)fill[w]++] = u;
            }
        }
    }
    // handle ids: ascending distinct classes; without a class attribute the single group 0 (capture.cpp:113-118)
    std::vector<int32_t> classes;
    if (!rig_class) {
        classes.push_back(0);
    } else {
        classes.assign(rig_class, rig_class + N);
        std::sort(classes.begin(), classes.end());
        classes.erase(std::unique(classes.begin(), classes.end()), classes.end());
    }
    const int32_t G = (int32_t)classes.size();
    *n_groups = G;
    if (G > grp_cap) { snprintf(ctx->err, sizeof(ctx->err), "capture: %d groups exceed grp_cap %d", G, grp_cap); return FD_E_INVALID; }

    FD_CUDA_OK(ctx, cudaStreamSynchronize(s)); // nearest_idx is on the host now

    // ---- findIslands: ring walk per rig point, union per class (capture.cpp:120-138) --------------------------
    std::fill(member, member + V, (uint8_t)0);
    std::vector<int64_t> pairs; // (group << 32) | vertex
    std::vector<int32_t> depth((size_t)V, -1), queue;
    queue.reserve(1024);
    for (int32_t i = 0; i < N; ++i) {
        const int32_t target = nearest_idx[i];
        if (target < 0) continue;
        int32_t g = 0;
        if (rig_class) g = (int32_t)(std::lower_bound(classes.begin(), classes.end(), rig_class[i]) - classes.begin());
        queue.clear();
        queue.push_back(target);
        depth[target] = 0;
        for (size_t head = 0; head < queue.size(); ++head) {
            const int32_t u = queue[head];
            if (depth[u] >= max_edges) continue;
            for (int64_t e = adj_off[u]; e < adj_off[(size_t)u + 1]; ++e) {
                const int32_t w = adj[(size_t)e];
                if (depth[w] < 0) {
                    depth[w] = depth[u] + 1;
                    queue.push_back(w);
                }
            }
        }
        for (int32_t u : queue) {
            depth[u] = -1;
            member[u] = 1;
            pairs.push_back(((int64_t)g << 32) | (int64_t)u);
        }
    }
    std::sort(pairs.begin(), pairs.end());
    pairs.erase(std::unique(pairs.begin(), pairs.end()), pairs.end());
    {
        size_t k = 0;
        int64_t total = 0;
        for (int32_t g = 0; g < G; ++g) {
            grp_class[g] = classes[g];
            grp_off[g] = total;
            for (; k < pairs.size() && (int32_t)(pairs[k] >> 32) == g; ++k, ++total)
                if (grp_idx && total < idx_cap) grp_idx[total] = (int32_t)(pairs[k] & 0xffffffff);
        }
        grp_off[G] = total;
    }
    if (G == 0) { // capture.cpp:54-56
        std::fill(dist2, dist2 + V, 0.0f);
        snprintf(ctx->err, sizeof(ctx->err), "%s", fd_status_string(FD_E_CAPTURE));
        return FD_E_CAPTURE;
    }

    // ---- capture: distance attribute on the GPU (capture.cpp:68-99) -------------------------------------------
    std::vector<int32_t> tri; // fan triangulation in primitive order; third index -1 marks a segment
    for (int32_t f = 0; f < n_rig_prim; ++f) {
        const int32_t b = rig_off[f], m = rig_off[f + 1] - b;
        if (m == 2) {
            tri.insert(tri.end(), {rig_vtx[b], rig_vtx[b + 1], -1});
        } else {
            for (int32_t k = 1; k + 1 < m; ++k) tri.insert(tri.end(), {rig_vtx[b], rig_vtx[b + k], rig_vtx[b + k + 1]});
        }
    }
    const int ntri = (int)(tri.size() / 3);
    void *dMember = nullptr, *dTri = nullptr, *dDist = nullptr;
    st = fd_stage(ctx, FD_STAGE_DIST, (size_t)V, &dMember);
    if (st == FD_OK) st = fd_stage(ctx, FD_STAGE_N, tri.size() * 4, &dTri);
    if (st == FD_OK) st = fd_stage(ctx, FD_STAGE_FALLOFF, (size_t)V * 4, &dDist);
    if (st != FD_OK) return st;
    if (V > 0) {
        FD_CUDA_OK(ctx, cudaMemcpyAsync(dMember, member, (size_t)V, cudaMemcpyHostToDevice, s));
        if (ntri > 0) FD_CUDA_OK(ctx, cudaMemcpyAsync(dTri, tri.data(), tri.size() * 4, cudaMemcpyHostToDevice, s));
        FD_CUDA_OK(ctx, fd_launch_capture_dist(ctx, (const float*)dP, V, (const uint8_t*)dMember, (const float*)dRig,
                                               (const int32_t*)dTri, ntri, radius, dofalloff, (float*)dDist));
        FD_CUDA_OK(ctx, cudaMemcpyAsync(dist2, dDist, (size_t)V * 4, cudaMemcpyDeviceToHost, s));
    }
    FD_CUDA_OK(ctx, cudaStreamSynchronize(s));
    return FD_OK;
}
