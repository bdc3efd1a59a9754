// SOP_FaceDeform_gpu.cpp -- Houdini-side shim: what a maintainer of symek/facedeform links into SOP_FaceDeform.so so
// that the node's RBF path runs on the GPU library.
//
// It is compiled ONLY when the Houdini toolkit is present (make HT=/path/to/hfs/toolkit; FD_WITH_HDK is then defined):
// this image has no HDK, so here the file is documentation that the build skips.  Nothing in it is arithmetic -- it
// moves GU_Detail data into the plain arrays fd::FaceDeformOp::cook takes (include/facedeform_sop.hpp) and the result
// back into gdp.  In the reference the call replaces the body of SOP_FaceDeform::cookMySop between the parameter
// reads (SOP_FaceDeform.cpp:244-263) and the data-ID bump (:483-486), i.e. the pack loop :268-287, the capture
// :310-322, the model build :331-373, the evaluation loop :384-439 and the morph-space pass :444-482; registration,
// parameter templates, input locking and cookInputGroups stay as they are.  See INTEGRATION.md for the diff.
#ifdef FD_WITH_HDK

#include <GA/GA_Handle.h>
#include <GA/GA_Iterator.h>
#include <GEO/GEO_Primitive.h>
#include <GU/GU_Detail.h>
#include <SOP/SOP_Node.h>
#include <UT/UT_Vector3.h>

#include <vector>

#include "facedeform_sop.hpp"

namespace fdgpu {

// points (by point INDEX, the order the reference packs by: SOP_FaceDeform.cpp:281), polygons as CSR, data IDs
struct DetailArrays {
    std::vector<float> P, tu, tv, nrm;
    std::vector<int32_t> prim_off, prim_vtx, cls;
    fd::Geo geo;
};

static void gather(const GU_Detail* gdp, bool with_prims, bool with_tangents, bool with_class, DetailArrays& out)
{
    const GA_Size npts = gdp->getNumPoints();
    out.P.resize((size_t)npts * 3);
    GA_Offset ptoff;
    GA_FOR_ALL_PTOFF(gdp, ptoff)
    {
        const GA_Index i = gdp->pointIndex(ptoff);
        const UT_Vector3 p = gdp->getPos3(ptoff);
        out.P[3 * i] = p.x(), out.P[3 * i + 1] = p.y(), out.P[3 * i + 2] = p.z();
    }
    out.geo = fd::Geo();
    out.geo.P = out.P.data();
    out.geo.npoints = npts;
    out.geo.p_data_id = gdp->getP()->getDataId();                 // InputGeoID, SOP_FaceDeform.hpp:47-63
    out.geo.topo_data_id = gdp->getTopology().getDataId();
    if (with_prims) {
        out.prim_off.assign(1, 0);
        for (GA_Iterator it(gdp->getPrimitiveRange()); !it.atEnd(); ++it) {
            const GEO_Primitive* prim = gdp->getGEOPrimitive(*it);
            for (GA_Size v = 0, nv = prim->getVertexCount(); v < nv; ++v)
                out.prim_vtx.push_back((int32_t)gdp->pointIndex(prim->getPointOffset(v)));
            out.prim_off.push_back((int32_t)out.prim_vtx.size());
        }
        out.geo.prim_off = out.prim_off.data();
        out.geo.prim_vtx = out.prim_vtx.data();
        out.geo.nprims = (int32_t)out.prim_off.size() - 1;
    }
    if (with_tangents) { // "tangentu" / "tangentv" / "N": SOP_FaceDeform.cpp:289-294
        GA_ROHandleV3 hu(gdp, GA_ATTRIB_POINT, "tangentu"), hv(gdp, GA_ATTRIB_POINT, "tangentv"), hn(gdp, GA_ATTRIB_POINT, "N");
        if (hu.isValid() && hv.isValid() && hn.isValid()) {
            out.tu.resize((size_t)npts * 3), out.tv.resize((size_t)npts * 3), out.nrm.resize((size_t)npts * 3);
            GA_FOR_ALL_PTOFF(gdp, ptoff)
            {
                const GA_Index i = gdp->pointIndex(ptoff);
                const UT_Vector3 u = hu.get(ptoff), v = hv.get(ptoff), n = hn.get(ptoff);
                for (int k = 0; k < 3; ++k) out.tu[3 * i + k] = u(k), out.tv[3 * i + k] = v(k), out.nrm[3 * i + k] = n(k);
            }
            out.geo.tangentu = out.tu.data(), out.geo.tangentv = out.tv.data(), out.geo.N = out.nrm.data();
        }
    }
    if (with_class) { // the rig's "class" attribute groups the handles: capture.cpp:113-118
        GA_ROHandleI hc(gdp, GA_ATTRIB_POINT, "class");
        if (hc.isValid()) {
            out.cls.resize((size_t)npts);
            GA_FOR_ALL_PTOFF(gdp, ptoff) out.cls[gdp->pointIndex(ptoff)] = hc.get(ptoff);
            out.geo.cls = out.cls.data();
        }
    }
}

// One cook of the RBF path.  `op` is a node member (it owns the fd_ctx, the cached factorisation, the capture and the
// blendshape projector -- the GPU-side counterpart of m_mesh_capture / m_direct_blends, SOP_FaceDeform.hpp:108-113);
// its `parms` were filled from the node's parameters by the caller.  blends: inputs 3.. (SOP_FaceDeform.cpp:199).
// Returns the worst severity and forwards the messages with the reference's texts.
OP_ERROR cookRbfPath(SOP_Node& node, GU_Detail* gdp, const GU_Detail* rest_rig, const GU_Detail* deform_rig,
                     const std::vector<const GU_Detail*>& blends, int64_t blends_data_id, fd::FaceDeformOp& op)
{
    DetailArrays mesh, rest, deform;
    gather(gdp, true, op.parms.tangent != 0, false, mesh);
    gather(rest_rig, true, false, true, rest);
    gather(deform_rig, false, false, false, deform);
    std::vector<DetailArrays> shapes(blends.size());
    if (op.parms.morphspace && !blends.empty()) {
        std::vector<const float*> ptrs;
        std::vector<int64_t> counts;
        for (size_t i = 0; i < blends.size(); ++i) {
            gather(blends[i], false, false, false, shapes[i]);
            ptrs.push_back(shapes[i].P.data());
            counts.push_back(shapes[i].geo.npoints);
        }
        op.setBlendshapes(ptrs, counts, blends_data_id);
    }
    std::vector<float> P_out(mesh.P.size()), falloff((size_t)mesh.geo.npoints);
    const fd::CookStatus st = op.cook(mesh.geo, rest.geo, deform.P.data(), deform.geo.npoints, 1, P_out.data(), falloff.data());
    for (const std::string& m : op.errors()) node.addError(SOP_MESSAGE, m.c_str());
    for (const std::string& m : op.warnings()) node.addWarning(SOP_MESSAGE, m.c_str());
    for (const std::string& m : op.messages()) node.addMessage(SOP_MESSAGE, m.c_str());
    if (st == fd::COOK_ERROR) return node.error();
    GA_RWHandleF fall_h(gdp->addFloatTuple(GA_ATTRIB_POINT, "fd_falloff", 1)); // SOP_FaceDeform.cpp:401
    GA_Offset ptoff;
    GA_FOR_ALL_PTOFF(gdp, ptoff)
    {
        const GA_Index i = gdp->pointIndex(ptoff);
        gdp->setPos3(ptoff, UT_Vector3(P_out[3 * i], P_out[3 * i + 1], P_out[3 * i + 2]));
        fall_h.set(ptoff, falloff[i]);
    }
    if (op.parms.morphspace && !op.blendWeights().empty()) { // the "weights" detail attribute, SOP_FaceDeform.cpp:474-480
        GA_Attribute* w_attrib = gdp->addFloatArray(GA_ATTRIB_DETAIL, "weights", 1);
        if (const GA_AIFNumericArray* aif = w_attrib ? w_attrib->getAIFNumericArray() : nullptr) {
            UT_FprealArray w;
            for (double v : op.blendWeights()) w.append(v);
            aif->set(w_attrib, 0, w);
            w_attrib->bumpDataId();
        }
    }
    if (op.positionsBumped()) gdp->getP()->bumpDataId(); // :485-486
    return node.error();
}

} // namespace fdgpu

#endif // FD_WITH_HDK
