"""The C++ operator mirror (fd::FaceDeformOp = SOP_FaceDeform::cookMySop, include/facedeform_sop.hpp) driven through
its C shim: the cook order, error / warning texts and caching of the reference, results against the oracle."""
import ctypes as C

import numpy as np
import pytest

from facedeform_b200 import synth

pytestmark = pytest.mark.gpu


class Sop:
    def __init__(self):
        from facedeform_b200 import _lib
        self.L = _lib.load()
        self.h = self.L.fd_sop_create(-1)
        self.parms = self.L.fd_sop_params(self.h).contents

    def cook(self, mesh, rig, deform, mesh_ids=(1, 1), rig_ids=(1, 1), tangents=False, cls=None, n_deform=None):
        F, V = deform.shape[0], mesh.P.shape[0]
        out = np.empty((F, V, 3), np.float32)
        fall = np.zeros(V, np.float32)
        p = lambda a: None if a is None else a.ctypes.data
        tu, tv, nn = (mesh.tangentu, mesh.tangentv, mesh.N) if tangents else (None, None, None)
        st = self.L.fd_sop_cook(self.h, p(mesh.P), V, p(mesh.poly_off), p(mesh.poly_vtx), len(mesh.poly_off) - 1,
                                p(tu), p(tv), p(nn), mesh_ids[0], mesh_ids[1], p(rig.rest), rig.rest.shape[0],
                                p(rig.prim_off), p(rig.prim_vtx), len(rig.prim_off) - 1, p(cls), rig_ids[0], rig_ids[1],
                                p(deform), deform.shape[1] if n_deform is None else n_deform, F, p(out), p(fall))
        return st, out, fall

    def msgs(self, kind):
        return self.L.fd_sop_messages(self.h, kind).decode()

    def close(self):
        self.L.fd_sop_destroy(self.h)


def test_cook_matches_oracle_and_caches_the_factorisation(oracle):
    mesh = synth.face_mesh(10_000)
    rig = synth.control_rig(64, prims=True)
    deform = synth.deformed_rig(rig, 3)
    sop = Sop()
    sop.parms.model, sop.parms.radius, sop.parms.dofalloff, sop.parms.falloffrate = 1, 2 * rig.spacing, 1, 2.0
    sop.parms.lambda_, sop.parms.maxedges = 0.0, 5           # lambda is clamped to 0.01 by the cook (:253)
    st, out, fall = sop.cook(mesh, rig, deform)
    assert st == 0, sop.msgs(0)
    assert "Termination type: 1, Iterations: 0" in sop.msgs(2)            # SOP_FaceDeform.cpp:370-373
    op = oracle.make_params(model=1, radius=2 * rig.spacing, dofalloff=1, falloffrate=2.0, maxedges=5, **{"lambda": 0.0})
    oracle.clamp_params(op)
    assert op.lambda_ == pytest.approx(0.01)
    cap = oracle.capture(mesh.P, mesh.poly_off, mesh.poly_vtx, rig.rest, rig.prim_off, rig.prim_vtx, None, 5, op.radius, 1)
    s, rad, W = oracle.fit(op, rig.rest, deform)
    ref, rfall = oracle.evaluate(op, rig.rest, rad, W, mesh.P, cap["dist2"])
    amp = np.maximum(rfall, 1.0)[None, :, None]
    assert (np.abs(out - ref) / amp).max() <= 1e-5 * mesh.bbox_diag
    np.testing.assert_allclose(fall, rfall, rtol=2e-6, atol=1e-7)
    assert sop.L.fd_sop_fit_count(sop.h) == 1
    # same rest rig (same data ids), new deformed rig: solve + eval only, the factorisation is reused
    st, out2, _ = sop.cook(mesh, rig, synth.deformed_rig(rig, 3, seed=9))
    assert st == 0 and sop.L.fd_sop_fit_count(sop.h) == 1 and not np.array_equal(out, out2)
    # rest rig data id bumped: re-capture and re-fit
    st, out3, _ = sop.cook(mesh, rig, deform, rig_ids=(2, 1))
    assert st == 0 and sop.L.fd_sop_fit_count(sop.h) == 2 and np.array_equal(out3, out)
    sop.close()


def test_cook_errors_and_warnings_quote_the_reference():
    mesh = synth.face_mesh(2_000)
    rig = synth.control_rig(16, prims=True)
    deform = synth.deformed_rig(rig, 1)
    sop = Sop()
    sop.parms.model, sop.parms.radius = 1, 0.3
    st, _, _ = sop.cook(mesh, rig, deform[:, :15].copy())
    assert st == 2 and sop.msgs(0) == "Rest and deform geometry should match."        # :231-234
    sop.parms.tangent = 1
    st, _, _ = sop.cook(mesh, rig, deform, tangents=False)
    assert st == 1 and sop.msgs(1).startswith("Append PolyFrameSOP")                   # :295-298
    st, out_t, _ = sop.cook(mesh, rig, deform, tangents=True)
    assert st == 0
    rest = rig.rest.copy()
    rest[3] = rest[2]
    dup = synth.Rig(rest, rig.normals, rig.prim_off, rig.prim_vtx, rig.spacing)
    # duplicate centres: with the Multilayer model the clamped lambda >= 0.01 (:253) regularises the system, so the
    # failure shows with the QNN model, whose nearest-neighbour radius becomes zero (terminationtype -5)
    sop.parms.tangent, sop.parms.model = 0, 0
    st, _, _ = sop.cook(mesh, dup, deform, rig_ids=(5, 5))
    assert st == 2 and sop.msgs(0) == "Can't solve the problem."                       # :365-368
    empty = synth.Rig(np.zeros((0, 3), np.float32), rig.normals[:0], rig.prim_off[:1], rig.prim_vtx[:0], 1.0)
    sop.parms.model = 1
    st, _, _ = sop.cook(mesh, empty, np.zeros((1, 0, 3), np.float32), rig_ids=(6, 6), cls=np.zeros(0, np.int32))
    assert st == 2 and sop.msgs(0) == "Can't capture geometry with a rig!"             # :318-321
    sop.close()


def test_cook_morph_space_pass(oracle):
    """morphspace = 1 with blendshape inputs: after the RBF evaluation the cook projects the deformation onto the
    blendshapes (SOP_FaceDeform.cpp:444-482, dbse.cpp) -- same result as the oracle's DirectBSEdit fed the cook's own
    RBF output, the reference's warnings for mismatched shapes, and the "weights" detail attribute."""
    mesh = synth.face_mesh(4_000)
    rig = synth.control_rig(32, prims=True)
    deform = synth.deformed_rig(rig, 2)
    rng = np.random.default_rng(4)
    V, S = mesh.P.shape[0], 5
    shapes = (mesh.P[None] + 0.05 * rng.standard_normal((S, V, 3))).astype(np.float32)
    sop = Sop()
    sop.parms.model, sop.parms.radius, sop.parms.lambda_ = 1, 2 * rig.spacing, 0.0
    st, plain, _ = sop.cook(mesh, rig, deform)                     # morphspace off: the RBF result
    assert st == 0
    sop.parms.morphspace, sop.parms.doclampweight = 1, 1
    sop.parms.weightrange[0], sop.parms.weightrange[1] = -0.5, 0.75
    assert sop.L.fd_sop_set_blendshapes(sop.h, shapes.ctypes.data, S, V, 7) == 0
    st, out, _ = sop.cook(mesh, rig, deform)
    assert st == 0, sop.msgs(1)
    M = oracle.dbse_shapes_matrix(mesh.P, shapes)
    QR, _ = oracle.householder_qr(M)
    for f in range(2):
        w = oracle.dbse_weights(QR, plain[f], mesh.P)
        ref = oracle.dbse_displace(M, w, plain[f], mesh.P, weightrange=(-0.5, 0.75))
        np.testing.assert_allclose(out[f], ref, rtol=0, atol=3e-6)
    wts = np.zeros(S)
    assert sop.L.fd_sop_blend_weights(sop.h, wts.ctypes.data, S) == S
    np.testing.assert_allclose(wts, w, rtol=0, atol=1e-10 * np.abs(w).max())   # weights of the last cooked frame
    # a blendshape set with the wrong point count is ignored with the reference's warnings (:201-205, :209-211)
    bad = shapes[:, :-1].copy()
    assert sop.L.fd_sop_set_blendshapes(sop.h, bad.ctypes.data, S, V - 1, 8) == 0
    st, out_bad, _ = sop.cook(mesh, rig, deform)
    assert st == 1 and "Some blendshapes don't match rest pose point count. Ignoring them." in sop.msgs(1)
    assert "Can't proceed with morph space deformation. Ingoring it." in sop.msgs(1)
    np.testing.assert_array_equal(out_bad, plain)
    sop.close()


def test_cook_recaptures_when_the_capture_parameters_change():
    """the reference's FIXME (SOP_FaceDeform.cpp:310): radius / max_edges changes did not re-capture; here they do."""
    mesh = synth.face_mesh(6_000)
    rig = synth.control_rig(32, prims=True)
    deform = synth.deformed_rig(rig, 1)
    sop = Sop()
    sop.parms.model, sop.parms.radius, sop.parms.dofalloff, sop.parms.maxedges = 1, 2 * rig.spacing, 1, 3
    st, out_a, fall_a = sop.cook(mesh, rig, deform)
    assert st == 0
    sop.parms.maxedges = 12                                  # wider rings: more vertices grouped, more of them fall off
    st, out_b, fall_b = sop.cook(mesh, rig, deform)
    assert st == 0 and not np.array_equal(fall_a, fall_b)
    assert sop.L.fd_sop_fit_count(sop.h) == 1               # maxedges is a capture parameter: no re-fit
    sop.parms.radius = 4 * rig.spacing                      # also the RBF radius (model = Multilayer): re-fit and re-capture
    st, out_c, fall_c = sop.cook(mesh, rig, deform)
    assert st == 0 and not np.array_equal(fall_b, fall_c) and sop.L.fd_sop_fit_count(sop.h) == 2
    sop.close()
