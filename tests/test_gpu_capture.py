"""GPU capture (fd_capture) vs the oracle: indices, group membership and the distance attribute are BIT-EXACT
(capture.cpp:46-141; the GPU kernels and the oracle state the same un-fused FP32 operation order)."""
import numpy as np
import pytest

from facedeform_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from facedeform_b200 import Context
    c = Context()
    yield c
    c.close()


def _same(a, b):
    assert a["ngroups"] == b["ngroups"]
    for k in ("nearest_idx", "member", "grp_class", "grp_off", "grp_idx"):
        assert np.array_equal(a[k], b[k]), k
    assert np.array_equal(a["dist2"].view(np.uint32), b["dist2"].view(np.uint32)), "dist2 not bit-exact"


@pytest.mark.parametrize("V,N,edges,with_class,falloff", [
    (2500, 12, 3, False, 1), (2500, 12, 3, True, 1), (10_000, 64, 4, False, 1), (10_000, 64, 1, True, 0),
    (40_000, 256, 4, True, 1), (999, 7, 6, False, 1),
])
def test_capture_bit_exact(ctx, oracle, V, N, edges, with_class, falloff):
    mesh = synth.face_mesh(V)
    rig = synth.control_rig(N, prims=True)
    cls = (np.random.default_rng(4).integers(-2, 5, N)).astype(np.int32) if with_class else None
    R = 1.5 * rig.spacing
    args = (mesh.P, mesh.poly_off, mesh.poly_vtx, rig.rest, rig.prim_off, rig.prim_vtx, cls, edges, R, falloff)
    got = ctx.capture(*args)
    want = oracle.capture(*args)
    _same(got, want)
    if falloff:
        d = got["dist2"][got["member"]]
        assert (d == -1).any() or (d > 0).any()


def test_capture_segments_and_quads(ctx, oracle):
    mesh = synth.face_mesh(3000)
    rig = synth.control_rig(16)
    # a rig made of a polyline (segments) and two quads
    seg = np.array([[0, 1], [1, 2], [2, 3]], np.int32)
    quad = np.array([[4, 5, 9, 8], [5, 6, 10, 9]], np.int32)
    off = np.array([0, 2, 4, 6, 10, 14], np.int32)
    vtx = np.concatenate([seg.reshape(-1), quad.reshape(-1)]).astype(np.int32)
    args = (mesh.P, mesh.poly_off, mesh.poly_vtx, rig.rest, off, vtx, None, 5, 0.4, 1)
    _same(ctx.capture(*args), oracle.capture(*args))


def test_capture_edge_cases(ctx, oracle):
    from facedeform_b200 import FdError
    mesh = synth.face_mesh(400)
    rig = synth.control_rig(5)
    args = (mesh.P, mesh.poly_off, mesh.poly_vtx, rig.rest, None, None, None, 2, 0.5, 1)
    got = ctx.capture(*args)                      # no primitives: every grouped vertex gets -1
    _same(got, oracle.capture(*args))
    assert np.all(got["dist2"][got["member"]] == -1)
    with pytest.raises(FdError) as e:             # empty rig with a class attribute: no groups (capture.cpp:54-56)
        ctx.capture(mesh.P, mesh.poly_off, mesh.poly_vtx, np.zeros((0, 3), np.float32), None, None,
                    np.zeros(0, np.int32), 2, 0.5, 1)
    assert e.value.status == 7
    got = ctx.capture(mesh.P, mesh.poly_off, mesh.poly_vtx, np.zeros((0, 3), np.float32), None, None, None, 2, 0.5, 1)
    assert got["ngroups"] == 1 and not got["member"].any()


def test_capture_feeds_eval(ctx, oracle):
    """cook order of the SOP: capture -> fit -> eval with the captured distance attribute (SOP_FaceDeform.cpp:311-439)."""
    from facedeform_b200 import make_params
    mesh = synth.face_mesh(10_000)
    rig = synth.control_rig(64, prims=True)
    deform = synth.deformed_rig(rig, 1)
    R = 2 * rig.spacing
    p = make_params(clamp=True, model=1, radius=R, dofalloff=1, falloffrate=2.0, maxedges=6, **{"lambda": 0.01})
    cap = ctx.capture(mesh.P, mesh.poly_off, mesh.poly_vtx, rig.rest, rig.prim_off, rig.prim_vtx, None,
                      p.maxedges, p.radius, p.dofalloff)
    m = ctx.fit(p, rig.rest).solve(deform)
    out, fall = m.eval(mesh.P, cap["dist2"])
    op = oracle.make_params(model=1, radius=p.radius, dofalloff=1, falloffrate=2.0, maxedges=6, **{"lambda": p.lambda_})
    st, rad, W = oracle.fit(op, rig.rest, deform)
    ocap = oracle.capture(mesh.P, mesh.poly_off, mesh.poly_vtx, rig.rest, rig.prim_off, rig.prim_vtx, None, 6, p.radius, 1)
    ref, rfall = oracle.evaluate(op, rig.rest, rad, W, mesh.P, ocap["dist2"])
    assert np.abs(out - ref).max() <= 1e-5 * mesh.bbox_diag
    np.testing.assert_allclose(fall, rfall, rtol=2e-6, atol=1e-7)
    m.close()
