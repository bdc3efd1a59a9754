// facedeform_sop.cpp -- host-side mirror of SOP_FaceDeform::cookMySop and ProximityCapture over the C ABI
// (see include/facedeform_sop.hpp).  Pure host code: every number comes from libfacedeform_gpu's kernels.
#include "facedeform_sop.hpp"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace fd {

// Point-group pattern of the `group` parm (SOP_FaceDeform.cpp:119-120, resolved by cookInputPointGroups :155-173 in
// Houdini).  The numeric subset of Houdini's syntax: space separated tokens "*", "n", "a-b", "a-b:step", each optionally
// prefixed by "^" (remove from what was selected so far; a pattern that starts with "^" starts from all points).
bool resolvePointGroup(const char* pattern, int64_t npoints, std::vector<uint8_t>& mask)
{
    mask.assign((size_t)npoints, 0);
    const char* s = pattern;
    bool first = true;
    while (*s) {
        while (*s == ' ' || *s == '\t' || *s == ',') ++s;
        if (!*s) break;
        bool remove = false;
        if (*s == '^') { remove = true; ++s; }
        if (first && remove) mask.assign((size_t)npoints, 1);
        first = false;
        int64_t a = 0, b = 0, step = 1;
        if (*s == '*') {
            a = 0, b = npoints - 1;
            ++s;
        } else {
            char* end = nullptr;
            a = std::strtoll(s, &end, 10);
            if (end == s || a < 0) return false;
            s = end;
            b = a;
            if (*s == '-') {
                ++s;
                b = std::strtoll(s, &end, 10);
                if (end == s || b < a) return false;
                s = end;
                if (*s == ':') {
                    ++s;
                    step = std::strtoll(s, &end, 10);
                    if (end == s || step < 1) return false;
                    s = end;
                }
            }
        }
        if (*s && *s != ' ' && *s != '\t' && *s != ',') return false;
        for (int64_t i = a; i <= b && i < npoints; i += step) mask[(size_t)i] = remove ? 0 : 1;
    }
    return true;
}

bool ProximityCapture::init(const Geo& mesh, const Geo& rig)
{
    // capture.cpp:10-44: the point trees, ray cache and edge structure are built inside fd_capture; here only
    // the inputs are latched and the detached attributes (re)allocated.
    m_mesh = mesh;
    m_rig = rig;
    m_dist.assign((size_t)mesh.npoints, 0.0f); // dist_a, default 0 (capture.cpp:31)
    m_member.assign((size_t)mesh.npoints, 0);
    m_nearest.assign((size_t)rig.npoints, -1);
    if (!m_ctx || (mesh.npoints > 0 && !mesh.P)) {
        m_init = false;
        m_capture = false;
        init_counter = 0;
        capture_counter = 0;
    } else {
        init_counter++;
        m_init = true;
        m_capture = false;
    }
    return m_init;
}

bool ProximityCapture::capture(const int& max_edges, const float& radius, const int& dofalloff, const float& /*falloffrate*/)
{
    if (!m_init) return false; // capture.cpp:50-52
    std::vector<int32_t> grp_class((size_t)m_rig.npoints + 1);
    std::vector<int64_t> grp_off((size_t)m_rig.npoints + 2);
    int32_t ngroups = 0;
    const int st = fd_capture(m_ctx, m_mesh.P, m_mesh.npoints, m_mesh.prim_off, m_mesh.prim_vtx, m_mesh.nprims, m_rig.P,
                              (int32_t)m_rig.npoints, m_rig.prim_off, m_rig.prim_vtx, m_rig.nprims, m_rig.cls, max_edges,
                              radius, dofalloff, m_nearest.data(), m_member.data(), m_dist.data(), &ngroups,
                              grp_class.data(), grp_off.data(), nullptr, (int32_t)grp_class.size(), 0);
    m_groups = ngroups;
    if (st != FD_OK) return false; // at least one island should be found (capture.cpp:54-56)
    capture_counter++;
    m_capture = true;
    return m_capture;
}

FaceDeformOp::FaceDeformOp(int device) : m_mesh_capture(nullptr)
{
    fd_params_default(&parms);
    if (fd_ctx_create(&m_ctx, device, nullptr) != FD_OK) m_ctx = nullptr;
    m_mesh_capture = ProximityCapture(m_ctx);
    m_direct_blends.reset(new DirectBSEdit(m_ctx));
}

// ---- DirectBSEdit (reference src/dbse.cpp) over fd_dbse_* ----------------------------------------------------------
DirectBSEdit::~DirectBSEdit()
{
    if (m_h) fd_dbse_destroy(m_h);
}

bool DirectBSEdit::init(const Geo& gdp, const ShapesVector& shapes)
{
    if (m_h) {
        fd_dbse_destroy(m_h);
        m_h = nullptr;
    }
    m_computed = false; // we need to recompute weights (dbse.cpp:33)
    m_shapes = (int)shapes.size();
    if (!m_ctx || shapes.empty() || gdp.npoints < 1) return false;
    // the C ABI takes the S shapes contiguously (S x P x 3)
    std::vector<float> packed((size_t)shapes.size() * gdp.npoints * 3);
    for (size_t s = 0; s < shapes.size(); ++s)
        std::memcpy(packed.data() + s * gdp.npoints * 3, shapes[s], (size_t)gdp.npoints * 3 * sizeof(float));
    return fd_dbse_init(m_ctx, gdp.P, gdp.npoints, packed.data(), (int32_t)shapes.size(), &m_h) == FD_OK;
}

bool DirectBSEdit::computeWeights(const float* pos, const float* rest)
{
    if (!m_h || !pos || !rest) return false; // rest_h.isInvalid(), dbse.cpp:39-41
    m_computed = fd_dbse_compute_weights(m_h, pos, rest, nullptr) == FD_OK;
    return m_computed;
}

bool DirectBSEdit::displace(const float* pos, const float* rest, const float* clamp, int dofalloff, float falloffradius,
                            float* P_out)
{
    return m_h && m_computed &&
           fd_dbse_displace(m_h, pos, rest, clamp ? 1 : 0, clamp, dofalloff, falloffradius, P_out) == FD_OK;
}

bool DirectBSEdit::getWeights(std::vector<double>& weights_array)
{
    if (!m_computed) return false; // dbse.cpp:79-81
    weights_array.resize((size_t)m_shapes);
    return fd_dbse_get_weights(m_h, weights_array.data()) == FD_OK;
}

void FaceDeformOp::setBlendshapes(const std::vector<const float*>& shapes, const std::vector<int64_t>& npoints,
                                  int64_t data_id)
{
    m_blend_shapes = shapes;
    m_blend_npoints = npoints;
    m_blend_id = data_id;
}

FaceDeformOp::~FaceDeformOp()
{
    if (m_model) fd_model_destroy(m_model);
    if (m_ctx) fd_ctx_destroy(m_ctx);
}

CookStatus FaceDeformOp::cook(const Geo& mesh, const Geo& rest_rig, const float* deform_rig_P, int64_t deform_npoints,
                              int frames, float* P_out, float* falloff_out)
{
    m_errors.clear();
    m_warnings.clear();
    m_messages.clear();
    if (!m_ctx) {
        addError("No CUDA device: the GPU deformation path has no CPU fallback.");
        return COOK_ERROR;
    }
    // Points count in control rig should match (SOP_FaceDeform.cpp:231-234)
    if (rest_rig.npoints != deform_npoints) {
        addError("Rest and deform geometry should match.");
        return COOK_ERROR;
    }
    // parameter clamps (SOP_FaceDeform.cpp:249-257)
    fd_params p = parms;
    fd_params_clamp(&p);
    // Do we do tangent projection? (:289-298)
    const bool do_tangent_disp = p.tangent && mesh.tangentu && mesh.tangentv && mesh.N;
    if (p.tangent && !do_tangent_disp)
        addWarning("Append PolyFrameSOP and enable tangent[u/v] and N attribute to allow tangent displacement.");
    // change tracking (:222-223, :236-241, :301-305)
    const bool rest_pose_changed = m_mesh_ids[0] != mesh.p_data_id || m_mesh_ids[1] != mesh.topo_data_id ||
                                   mesh.p_data_id < 0;
    const bool rest_rig_changed = m_rig_ids[0] != rest_rig.p_data_id || m_rig_ids[1] != rest_rig.topo_data_id ||
                                  rest_rig.p_data_id < 0;
    m_mesh_ids[0] = mesh.p_data_id;
    m_mesh_ids[1] = mesh.topo_data_id;
    m_rig_ids[0] = rest_rig.p_data_id;
    m_rig_ids[1] = rest_rig.topo_data_id;
    // Proximity capture (:310-322).  The reference re-captures only when the mesh or the rig changed and carries a FIXME
    // for the parameters the capture depends on (:310: radius, max_edges; also dofalloff); here a change of any of them
    // re-captures too.
    const bool capture_parms_changed = m_cap_maxedges != p.maxedges || m_cap_radius != p.radius || m_cap_dofalloff != p.dofalloff;
    if (rest_pose_changed || rest_rig_changed || capture_parms_changed || !m_mesh_capture.isInitialized() ||
        !m_mesh_capture.isCaptured()) {
        if (!m_mesh_capture.init(mesh, rest_rig)) {
            addError("Can't initialize geometry to capture with a rig!");
            return COOK_ERROR;
        }
        if (!m_mesh_capture.capture(p.maxedges, p.radius, p.dofalloff, p.falloffrate)) {
            addError("Can't capture geometry with a rig!");
            return COOK_ERROR;
        }
        m_cap_maxedges = p.maxedges;
        m_cap_radius = p.radius;
        m_cap_dofalloff = p.dofalloff;
    }
    // Create / build the model (:331-368).  The reference rebuilds it every cook; here the factorisation is kept
    // while the rest rig and the fit parameters are unchanged ("once per rest pose").
    // A parameter the fit does not depend on (tangent, falloff, maxedges, morph space, group ...) keeps the factorisation.
    const bool refit = !m_model || rest_rig_changed || !fd_params_fit_equal(&m_fit_parms, &p);
    if (refit) {
        if (m_model) {
            fd_model_destroy(m_model);
            m_model = nullptr;
        }
        fd_report rep;
        const int st = fd_rbf_fit(m_ctx, &p, rest_rig.P, (int32_t)rest_rig.npoints, &m_model, &rep);
        if (st == FD_E_SINGULAR) {
            addError("Can't solve the problem."); // :365-368
            return COOK_ERROR;
        }
        if (st != FD_OK) {
            addError("Can't build RBF model."); // :337-340
            return COOK_ERROR;
        }
        ++m_fit_counter;
    } else if (fd_model_set_epilogue(m_model, &p) != FD_OK) {
        addError(fd_last_error(m_ctx));
        return COOK_ERROR;
    }
    m_fit_parms = p;
    fd_report rep;
    int st = fd_rbf_solve(m_model, deform_rig_P, (int32_t)deform_npoints, frames, &rep);
    if (st != FD_OK) {
        addError(st == FD_E_SINGULAR ? "Can't solve the problem." : fd_last_error(m_ctx));
        return COOK_ERROR;
    }
    char info_buffer[200];
    std::snprintf(info_buffer, sizeof(info_buffer), "Termination type: %d, Iterations: %d", rep.terminationtype,
                  rep.iterationscount); // :370-373
    addMessage(info_buffer);
    // Here we determine which groups we have to work on (cookInputGroups, :380-381).  The reference resolves the group
    // and then never consults it inside the loop (:404-439): it only gates the data-ID bump (:485-486).
    // strict_reference = 1 reproduces that; otherwise the points outside the group are left where they are.
    std::vector<uint8_t> gmask;
    const bool have_group = p.group[0] != 0;
    if (have_group && !resolvePointGroup(p.group, mesh.npoints, gmask)) {
        addError("Invalid point group pattern.");
        return COOK_ERROR;
    }
    bool group_empty = have_group;
    for (size_t i = 0; i < gmask.size() && group_empty; ++i) group_empty = gmask[i] == 0;
    m_p_bumped = !have_group || !group_empty;
    const float* dist = m_mesh_capture.getDistanceAttribute();
    if (!dist) addWarning("Can't find distance capture attribute. Won't apply radius nor falloff."); // :396-398
    if (have_group && !p.strict_reference) {
        // outside the group: a distance beyond every radius, which the gate of :408-410 skips (P stays, falloff 0)
        m_group_dist.resize((size_t)mesh.npoints);
        for (int64_t i = 0; i < mesh.npoints; ++i) m_group_dist[(size_t)i] = gmask[(size_t)i] ? (dist ? dist[i] : 0.0f) : INFINITY;
        dist = m_group_dist.data();
    }
    st = fd_rbf_eval(m_model, mesh.P, mesh.npoints, dist, do_tangent_disp ? mesh.tangentu : nullptr,
                     do_tangent_disp ? mesh.tangentv : nullptr, do_tangent_disp ? mesh.N : nullptr, P_out, falloff_out);
    if (st != FD_OK) {
        addError(fd_last_error(m_ctx));
        return COOK_ERROR;
    }
    // Any inputs above 2 are morph targets: project the deformation onto them (SOP_FaceDeform.cpp:325-329, :444-482)
    if (p.morphspace && !m_blend_shapes.empty()) {
        // setupBlends (:175-213): the rest attribute is the undeformed mesh; re-init when the blendshapes changed
        if (rest_pose_changed || m_blend_built_id != m_blend_id || !m_direct_blends->isInitialized()) {
            DirectBSEdit::ShapesVector shapes;
            for (size_t i = 0; i < m_blend_shapes.size(); ++i) {
                if (i >= m_blend_npoints.size() || m_blend_npoints[i] != mesh.npoints) {
                    addWarning("Some blendshapes don't match rest pose point count. Ignoring them."); // :201-205
                    continue;
                }
                shapes.push_back(m_blend_shapes[i]);
            }
            if (!m_direct_blends->init(mesh, shapes))
                addWarning("Can't proceed with morph space deformation. Ingoring it."); // :209-211
            m_blend_built_id = m_blend_id;
        }
        if (m_direct_blends->isInitialized()) {
            const float* clamp = p.doclampweight ? p.weightrange : nullptr; // :455-458
            for (int f = 0; f < frames; ++f) {
                float* Pf = P_out + (size_t)f * mesh.npoints * 3;
                // The reference computes the weights only while !isComputed() and warns otherwise (:446-452): after the
                // first cook that followed an init() the pass is skipped.  strict_reference = 1 reproduces that; the
                // default recomputes the weights for every cooked frame.
                bool weights_done = false;
                if (!p.strict_reference || !m_direct_blends->isComputed()) weights_done = m_direct_blends->computeWeights(Pf, mesh.P);
                if (!weights_done) {
                    addWarning("Can't compute weights for morphspace deformation. Ingoring it."); // :451-452
                    break;
                }
                m_direct_blends->displace(Pf, mesh.P, clamp, p.dofalloff, p.falloffradius, Pf); // :460-472
            }
            m_direct_blends->getWeights(m_blend_weights); // the "weights" detail attribute, :474-480
        }
    }
    return m_warnings.empty() ? COOK_OK : COOK_WARNING;
}

} // namespace fd

// ---- C shim so the operator mirror can be driven (and tested) through the same C ABI -----------------------------
extern "C" {

struct fd_sop {
    fd::FaceDeformOp op;
    std::string joined;
    std::vector<float> blends; // copy of the blendshape inputs (the shim owns what the operator points at)
    explicit fd_sop(int device) : op(device) {}
};

fd_sop* fd_sop_create(int device) { return new fd_sop(device); }
void fd_sop_destroy(fd_sop* s) { delete s; }
fd_params* fd_sop_params(fd_sop* s) { return s ? &s->op.parms : nullptr; }

int fd_sop_cook(fd_sop* s, const float* mesh_P, int64_t n_vtx, const int32_t* poly_off, const int32_t* poly_vtx,
                int32_t n_poly, const float* tangentu, const float* tangentv, const float* normal, int64_t mesh_p_id,
                int64_t mesh_topo_id, const float* rest_rig_P, int32_t n_rig, const int32_t* rig_off,
                const int32_t* rig_vtx, int32_t n_rig_prim, const int32_t* rig_class, int64_t rig_p_id,
                int64_t rig_topo_id, const float* deform_rig_P, int32_t n_deform, int32_t frames, float* P_out,
                float* falloff_out)
{
    if (!s) return fd::COOK_ERROR;
    fd::Geo mesh, rig;
    mesh.P = mesh_P; mesh.npoints = n_vtx; mesh.prim_off = poly_off; mesh.prim_vtx = poly_vtx; mesh.nprims = n_poly;
    mesh.tangentu = tangentu; mesh.tangentv = tangentv; mesh.N = normal;
    mesh.p_data_id = mesh_p_id; mesh.topo_data_id = mesh_topo_id;
    rig.P = rest_rig_P; rig.npoints = n_rig; rig.prim_off = rig_off; rig.prim_vtx = rig_vtx; rig.nprims = n_rig_prim;
    rig.cls = rig_class; rig.p_data_id = rig_p_id; rig.topo_data_id = rig_topo_id;
    return s->op.cook(mesh, rig, deform_rig_P, n_deform, frames, P_out, falloff_out);
}

// kind: 0 errors, 1 warnings, 2 messages; entries joined by '\n'
const char* fd_sop_messages(fd_sop* s, int kind)
{
    if (!s) return "";
    const std::vector<std::string>& v = kind == 0 ? s->op.errors() : (kind == 1 ? s->op.warnings() : s->op.messages());
    s->joined.clear();
    for (size_t i = 0; i < v.size(); ++i) s->joined += (i ? "\n" : "") + v[i];
    return s->joined.c_str();
}

int fd_sop_fit_count(const fd_sop* s) { return s ? s->op.fits() : 0; }
int fd_sop_positions_bumped(const fd_sop* s) { return (s && s->op.positionsBumped()) ? 1 : 0; }

int fd_sop_set_blendshapes(fd_sop* s, const float* shapes, int32_t n_shapes, int64_t n_pts, int64_t data_id)
{
    if (!s || n_shapes < 0 || (n_shapes > 0 && (!shapes || n_pts < 1))) return FD_E_INVALID;
    s->blends.assign(shapes, shapes + (size_t)n_shapes * n_pts * 3);
    std::vector<const float*> ptrs;
    std::vector<int64_t> np;
    for (int32_t i = 0; i < n_shapes; ++i) {
        ptrs.push_back(s->blends.data() + (size_t)i * n_pts * 3);
        np.push_back(n_pts);
    }
    s->op.setBlendshapes(ptrs, np, data_id);
    return FD_OK;
}

int fd_sop_blend_weights(fd_sop* s, double* weights, int32_t cap)
{
    if (!s) return -1;
    const std::vector<double>& w = s->op.blendWeights();
    if (weights)
        for (size_t i = 0; i < w.size() && (int32_t)i < cap; ++i) weights[i] = w[i];
    return (int)w.size();
}

} // extern "C"
