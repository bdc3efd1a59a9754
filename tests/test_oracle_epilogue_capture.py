"""Oracle epilogue (tangent projection, falloff, gate) and ProximityCapture restatement vs numpy re-derivations."""
import numpy as np
import pytest

from facedeform_b200 import synth


def _np_project(u, v, n, d):
    u, v, n = (x / np.linalg.norm(x) for x in (u, v, n))
    M = np.stack([u, v, n])
    B = M.T @ M
    a1 = u @ B
    a1 /= np.linalg.norm(a1)
    a2 = v @ B
    a2 /= np.linalg.norm(a2)
    return a1 * (d @ a1) + a2 * (d @ a2)


def test_project_to_tangents(oracle):
    rng = np.random.default_rng(0)
    for _ in range(20):
        u, v, n, d = rng.standard_normal((4, 3))
        un, vn, nn = (x / np.linalg.norm(x) for x in (u, v, n))
        got = oracle.project_to_tangents(un.astype(np.float32), vn.astype(np.float32), nn.astype(np.float32), d)
        np.testing.assert_allclose(got, _np_project(u, v, n, d), rtol=2e-5, atol=2e-6)
    # orthonormal frame: the projection removes exactly the normal component
    got = oracle.project_to_tangents([1, 0, 0], [0, 1, 0], [0, 0, 1], [0.3, -0.2, 0.9])
    np.testing.assert_allclose(got, [0.3, -0.2, 0.0], atol=1e-7)


def test_eval_epilogue(oracle):
    rig = synth.control_rig(32)
    deform = synth.deformed_rig(rig, 2)
    mesh = synth.face_mesh(400, topology=False)
    V = 400
    R = 0.5
    p = oracle.make_params(model=1, term=0, kernel=0, radius=R, tangent=1, dofalloff=1, falloffrate=2.5,
                           **{"lambda": 0.0})
    st, rad, W = oracle.fit(p, rig.rest, deform)
    rng = np.random.default_rng(5)
    dist2 = rng.uniform(0, 0.3, V).astype(np.float32)
    dist2[::7] = -1.0          # "grouped but farther than R" sentinel of capture.cpp:76
    dist2[3::11] = 0.26        # > R^2 = 0.25 -> skipped
    out, fall = oracle.evaluate(p, rig.rest, rad, W, mesh.P, dist2, mesh.tangentu, mesh.tangentv, mesh.N)
    raw = oracle.evaluate_raw(p, rig.rest, rad, W, mesh.P)
    r2 = np.float32(R) * np.float32(R)
    for v in range(V):
        if dist2[v] > r2:
            assert np.array_equal(out[:, v], np.stack([mesh.P[v]] * 2)) and fall[v] == 0
            continue
        fo = np.float32(1) - min(dist2[v] / r2, np.float32(1))
        fo = np.float32(fo) ** np.float32(2.5)
        assert fall[v] == pytest.approx(fo, rel=1e-6)
        if dist2[v] < 0:
            assert fall[v] > 1.0   # reference quirk: pow(1 + 1/R^2, rate)
        for f in range(2):
            d = _np_project(mesh.tangentu[v].astype(np.float64), mesh.tangentv[v].astype(np.float64),
                            mesh.N[v].astype(np.float64), raw[v, 3 * f:3 * f + 3])
            np.testing.assert_allclose(out[f, v], mesh.P[v] + d * fo, rtol=0, atol=3e-6)
    # no tangents supplied -> plain displacement (do_tangent_disp false, SOP_FaceDeform.cpp:293-298)
    out2, _ = oracle.evaluate(p, rig.rest, rad, W, mesh.P, dist2)
    keep = dist2 <= r2
    want = mesh.P[None] + raw.reshape(V, 2, 3).transpose(1, 0, 2).astype(np.float32) * fall[None, :, None]
    np.testing.assert_allclose(out2[:, keep], want[:, keep], atol=1e-6)
    # threaded == serial bit for bit
    out3, fall3 = oracle.evaluate(p, rig.rest, rad, W, mesh.P, dist2, mesh.tangentu, mesh.tangentv, mesh.N, nthreads=4)
    assert np.array_equal(out, out3) and np.array_equal(fall, fall3)


def _np_tri_dist2(p, a, b, c):
    # dense sampling-free reference: project onto plane, clamp via barycentric + edges (float64)
    def seg(p, a, b):
        ab = b - a
        t = np.clip((p - a) @ ab / (ab @ ab), 0, 1)
        return ((p - (a + t * ab)) ** 2).sum()
    n = np.cross(b - a, c - a)
    nn = n @ n
    if nn > 0:
        q = p - n * ((p - a) @ n) / nn
        T = np.stack([b - a, c - a], axis=1)
        st, *_ = np.linalg.lstsq(T, q - a, rcond=None)
        if st[0] >= 0 and st[1] >= 0 and st.sum() <= 1:
            return ((p - q) ** 2).sum()
    return min(seg(p, a, b), seg(p, b, c), seg(p, c, a))


def test_point_triangle_distance(oracle):
    rng = np.random.default_rng(11)
    for _ in range(300):
        a, b, c = rng.standard_normal((3, 3))
        p = rng.standard_normal(3) * 2
        got = oracle.point_tri_dist2(p, a, b, c)
        f32 = [x.astype(np.float32).astype(np.float64) for x in (p, a, b, c)]
        assert got == pytest.approx(_np_tri_dist2(*f32), rel=2e-4, abs=1e-6)
    assert oracle.point_seg_dist2([0, 1, 0], [-1, 0, 0], [1, 0, 0]) == pytest.approx(1.0)
    assert oracle.point_seg_dist2([3, 0, 0], [-1, 0, 0], [1, 0, 0]) == pytest.approx(4.0)


def _np_capture(mesh, rig, cls, max_edges, radius, dofalloff):
    V = mesh.P.shape[0]
    quads = mesh.poly_vtx.reshape(-1, 4)
    nbr = [set() for _ in range(V)]
    for q in quads:
        for k in range(4):
            a, b = int(q[k]), int(q[(k + 1) % 4])
            nbr[a].add(b)
            nbr[b].add(a)
    groups = {}
    if cls is None:
        groups[0] = set()
    nearest = []
    for i, c in enumerate(rig.rest):
        d = ((mesh.P - c) ** 2)
        d2 = (d[:, 0] + d[:, 1]) + d[:, 2]
        t = int(np.argmin(d2))
        nearest.append(t)
        ring, frontier = {t}, {t}
        for _ in range(max_edges):
            frontier = {w for u in frontier for w in nbr[u]} - ring
            ring |= frontier
        groups.setdefault(0 if cls is None else int(cls[i]), set()).update(ring)
    return np.array(nearest), groups


@pytest.mark.parametrize("with_class", [False, True])
def test_capture_groups_and_distance(oracle, with_class):
    mesh = synth.face_mesh(2500)
    rig = synth.control_rig(12, prims=True)
    cls = np.array([3, 3, 1, 1, 1, 7, 7, 7, 7, 2, 2, 2], np.int32) if with_class else None
    R = 0.06
    res = oracle.capture(mesh.P, mesh.poly_off, mesh.poly_vtx, rig.rest, rig.prim_off, rig.prim_vtx, cls, 3, R, 1)
    nearest, groups = _np_capture(mesh, rig, cls, 3, R, 1)
    assert np.array_equal(res["nearest_idx"], nearest)
    assert res["ngroups"] == len(groups)
    assert list(res["grp_class"]) == sorted(groups)
    for g, c in enumerate(res["grp_class"]):
        idx = res["grp_idx"][res["grp_off"][g]:res["grp_off"][g + 1]]
        assert list(idx) == sorted(groups[int(c)])
    member = np.zeros(len(mesh.P), bool)
    for s in groups.values():
        member[list(s)] = True
    assert np.array_equal(res["member"], member)
    # distance attribute: default 0 outside groups, closest triangle distance (< R^2) or -1 inside
    tris = rig.prim_vtx.reshape(-1, 3)
    d2 = res["dist2"]
    assert np.all(d2[~member] == 0)
    r2 = np.float32(R) * np.float32(R)
    for v in np.flatnonzero(member)[::5]:
        best = min(oracle.point_tri_dist2(mesh.P[v], *rig.rest[t]) for t in tris)
        assert d2[v] == (np.float32(best) if best < r2 else np.float32(-1))
    assert (d2[member] == -1).any() and (d2[member] > 0).any()
    # dofalloff == 0 => every grouped vertex gets 0 (capture.cpp:71-75)
    res0 = oracle.capture(mesh.P, mesh.poly_off, mesh.poly_vtx, rig.rest, rig.prim_off, rig.prim_vtx, cls, 3, R, 0)
    assert np.all(res0["dist2"] == 0) and np.array_equal(res0["member"], member)


def test_capture_edge_cases(oracle):
    mesh = synth.face_mesh(400)
    rig = synth.control_rig(5, prims=False)
    # a rig without primitives: minimumPoint finds nothing => -1 for grouped vertices
    res = oracle.capture(mesh.P, mesh.poly_off, mesh.poly_vtx, rig.rest, None, None, None, 2, 0.5, 1)
    assert res["ngroups"] == 1 and np.all(res["dist2"][res["member"]] == -1)
    # an empty rig with a class attribute => no groups => capture fails (capture.cpp:54-56)
    res = oracle.capture(mesh.P, mesh.poly_off, mesh.poly_vtx, np.zeros((0, 3), np.float32), None, None,
                         np.zeros(0, np.int32), 2, 0.5, 1)
    assert res["ngroups"] == 0
    # an empty rig without a class attribute still has the single default group (capture.cpp:114-118)
    res = oracle.capture(mesh.P, mesh.poly_off, mesh.poly_vtx, np.zeros((0, 3), np.float32), None, None,
                         None, 2, 0.5, 1)
    assert res["ngroups"] == 1 and not res["member"].any()
