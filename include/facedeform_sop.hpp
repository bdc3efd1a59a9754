// facedeform_sop.hpp -- C++ host-side mirror of the reference operator for the RBF path, over the C ABI.
//
// The reference is one Houdini SOP (src/SOP_FaceDeform.{hpp,cpp}) plus the ProximityCapture helper
// (src/capture.{hpp,cpp}).  Houdini's HDK is absent, so geometry arrives as plain arrays (fd::Geo) instead of
// GU_Detail; everything else keeps the reference's names, argument meaning and error behaviour:
//   fd::ProximityCapture::{init, capture, isInitialized, isCaptured, getDistanceAttribute}   capture.hpp:21-27
//   fd::DirectBSEdit::{init, isInitialized, computeWeights, isComputed, displace, getWeights}   dbse.hpp:16-22
//   fd::FaceDeformOp::cook                                          SOP_FaceDeform::cookMySop, SOP_FaceDeform.cpp:215-489
//   parameter accessors MODEL/TERM/QCOEF/ZCOEF/RADIUS/LAYERS/LAMBDA/TANGENT/MAXEDGES/DOFALLOFF/FALLOFFRATE
//                                                                   SOP_FaceDeform.hpp:84-101
// A Houdini shim would keep SOP_FaceDeform's registration/templates and replace the body of cookMySop between
// :268 and :439 by FaceDeformOp::cook.
#pragma once

#include <cstdint>
#include <memory>
#include <string>
#include <vector>

#include "facedeform_gpu.h"

namespace fd {

// what the SOP reads from a GU_Detail (Appendix A of SURVEY.md): points, polygons (CSR), optional point attributes
struct Geo {
    const float* P = nullptr;          // npoints x 3
    int64_t npoints = 0;
    const int32_t* prim_off = nullptr; // nprims + 1
    const int32_t* prim_vtx = nullptr;
    int32_t nprims = 0;
    const int32_t* cls = nullptr;      // rig "class" attribute (capture.cpp:113) or null
    const float* tangentu = nullptr;   // mesh "tangentu" / "tangentv" / "N" (SOP_FaceDeform.cpp:289-291) or null
    const float* tangentv = nullptr;
    const float* N = nullptr;
    int64_t p_data_id = -1;            // getP()->getDataId()        (InputGeoID, SOP_FaceDeform.hpp:47-63)
    int64_t topo_data_id = -1;         // getTopology().getDataId()
};

class ProximityCapture {
public:
    explicit ProximityCapture(fd_ctx* ctx) : m_ctx(ctx) {}
    bool init(const Geo& mesh, const Geo& rig);                                               // capture.cpp:10-44
    bool capture(const int& max_edges, const float& radius, const int& dofalloff, const float& falloffrate); // :46-105
    bool isInitialized() const { return m_init; }
    bool isCaptured() const { return m_capture; }
    const float* getDistanceAttribute() const { return m_capture ? m_dist.data() : nullptr; }
    const std::vector<uint8_t>& getMembership() const { return m_member; }
    const std::vector<int32_t>& getNearestIndices() const { return m_nearest; }
    int groups() const { return m_groups; }

private:
    fd_ctx* m_ctx;
    Geo m_mesh, m_rig;
    bool m_init = false, m_capture = false;
    int init_counter = 0, capture_counter = 0;
    int m_groups = 0;
    std::vector<float> m_dist;
    std::vector<uint8_t> m_member;
    std::vector<int32_t> m_nearest;
};

// the "morph space" post-pass (reference src/dbse.{hpp,cpp}); blendshapes arrive as plain P x 3 float arrays
class DirectBSEdit {
public:
    typedef std::vector<const float*> ShapesVector;
    explicit DirectBSEdit(fd_ctx* ctx) : m_ctx(ctx) {}
    ~DirectBSEdit();
    DirectBSEdit(const DirectBSEdit&) = delete;
    DirectBSEdit& operator=(const DirectBSEdit&) = delete;
    bool init(const Geo& gdp, const ShapesVector& shapes);                               // dbse.cpp:9-35
    bool isInitialized() const { return m_h != nullptr; }
    bool computeWeights(const float* pos, const float* rest);                            // dbse.cpp:37-58
    bool isComputed() const { return m_computed; }
    // displaceVector for every point + the position write of SOP_FaceDeform.cpp:460-472 (clamp: nullptr = no clamping)
    bool displace(const float* pos, const float* rest, const float* clamp, int dofalloff, float falloffradius, float* P_out);
    bool getWeights(std::vector<double>& weights_array);                                 // dbse.cpp:77-87

private:
    fd_ctx* m_ctx;
    fd_dbse* m_h = nullptr;
    bool m_computed = false;
    int m_shapes = 0;
};

// the `group` parm's pattern -> per-point membership (cookInputPointGroups, SOP_FaceDeform.cpp:155-173); false = syntax error
bool resolvePointGroup(const char* pattern, int64_t npoints, std::vector<uint8_t>& mask);

// status of a cook, the analogue of OP_ERROR + the node's message lists
enum CookStatus { COOK_OK = 0, COOK_WARNING = 1, COOK_ERROR = 2 };

class FaceDeformOp {
public:
    explicit FaceDeformOp(int device = -1);
    ~FaceDeformOp();
    FaceDeformOp(const FaceDeformOp&) = delete;
    FaceDeformOp& operator=(const FaceDeformOp&) = delete;

    fd_params parms; // raw parm values (defaults SOP_FaceDeform.cpp:117-137); cook() applies the clamps of :249-257

    // accessors with the reference's names (SOP_FaceDeform.hpp:84-101)
    int MODEL() const { return parms.model; }
    int TERM() const { return parms.term; }
    float QCOEF() const { return parms.qcoef; }
    float ZCOEF() const { return parms.zcoef; }
    float RADIUS() const { return parms.radius; }
    int LAYERS() const { return parms.layers; }
    float LAMBDA() const { return parms.lambda; }
    int TANGENT() const { return parms.tangent; }
    int MAXEDGES() const { return parms.maxedges; }
    int DOFALLOFF() const { return parms.dofalloff; }
    float FALLOFFRADIUS() const { return parms.falloffradius; }
    float FALLOFFRATE() const { return parms.falloffrate; }

    // cookMySop: mesh = input 0, rest_rig = input 1, deform_rig = input 2 (F frames of the deformed rig, F x N x 3).
    // P_out receives F x V x 3 positions, falloff_out (may be null) the fd_falloff attribute.
    CookStatus cook(const Geo& mesh, const Geo& rest_rig, const float* deform_rig_P, int64_t deform_npoints,
                    int frames, float* P_out, float* falloff_out);

    // inputs >= 3 of the SOP: the blendshapes of the morph-space pass (setupBlends, SOP_FaceDeform.cpp:175-213);
    // npoints[i] != mesh point count => the shape is ignored with the reference's warning.  data_id tracks changes.
    void setBlendshapes(const std::vector<const float*>& shapes, const std::vector<int64_t>& npoints, int64_t data_id);
    const std::vector<double>& blendWeights() const { return m_blend_weights; } // the "weights" detail attribute, :474-480

    const std::vector<std::string>& errors() const { return m_errors; }
    const std::vector<std::string>& warnings() const { return m_warnings; }
    const std::vector<std::string>& messages() const { return m_messages; }
    const ProximityCapture& capture() const { return m_mesh_capture; }
    fd_ctx* ctx() const { return m_ctx; }
    int fits() const { return m_fit_counter; } // how many times the system was factored (once per rest pose)
    // whether the last cook bumped P's data ID: `!myGroup || !myGroup->isEmpty()` (SOP_FaceDeform.cpp:485-486)
    bool positionsBumped() const { return m_p_bumped; }

private:
    void addError(const std::string& s) { m_errors.push_back(s); }
    void addWarning(const std::string& s) { m_warnings.push_back(s); }
    void addMessage(const std::string& s) { m_messages.push_back(s); }

    fd_ctx* m_ctx = nullptr;
    fd_model* m_model = nullptr;
    ProximityCapture m_mesh_capture;
    std::unique_ptr<DirectBSEdit> m_direct_blends;
    std::vector<const float*> m_blend_shapes;
    std::vector<int64_t> m_blend_npoints;
    int64_t m_blend_id = -2, m_blend_built_id = -3;
    std::vector<double> m_blend_weights;
    // InputGeoID trackers (SOP_FaceDeform.hpp:47-63): mesh, rest rig
    int64_t m_mesh_ids[2] = {-2, -2};
    int64_t m_rig_ids[2] = {-2, -2};
    fd_params m_fit_parms{};
    int m_cap_maxedges = -1, m_cap_dofalloff = -1; // the parameters of the last capture (FIXME of SOP_FaceDeform.cpp:310)
    float m_cap_radius = -1.f;
    int m_fit_counter = 0;
    bool m_p_bumped = true;
    std::vector<float> m_group_dist; // capture distances with the points outside `group` pushed beyond every radius
    std::vector<std::string> m_errors, m_warnings, m_messages;
};

} // namespace fd
