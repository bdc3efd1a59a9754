"""DirectBSEdit on the GPU (fd_dbse_*, csrc/fd_dbse.cu) against the oracle's restatement of reference src/dbse.cpp.
Stated tolerances: packed QR and weights 1e-11 relative to their largest entry (FP64, different but fixed summation
orders); displaced positions bit-exact when fed the same weights' FP32 factors, else within 2 ulp of the coordinate
scale (the FP32 factor (float)(3 w) may round differently when w differs in its last bits)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from facedeform_b200 import Context
    c = Context()
    yield c
    c.close()


def _case(seed, P, S):
    rng = np.random.default_rng(seed)
    rest = rng.standard_normal((P, 3)).astype(np.float32)
    shapes = (rest[None] + 0.1 * rng.standard_normal((S, P, 3))).astype(np.float32)
    pos = (rest + 0.05 * rng.standard_normal((P, 3))).astype(np.float32)
    return rest, shapes, pos


@pytest.mark.parametrize("P,S", [(1, 1), (7, 3), (1365, 1), (1366, 9), (5000, 40), (3, 6)])
def test_qr_weights_displace_match_the_oracle(ctx, oracle, P, S):
    from facedeform_b200 import DirectBSEdit, FdError
    rest, shapes, pos = _case(11, P, S)
    b = DirectBSEdit(ctx, rest, shapes)
    assert b.is_initialized() and not b.is_computed()
    with pytest.raises(FdError) as e:                  # getWeights before computeWeights: false, dbse.cpp:79-81
        b.get_weights()
    assert e.value.status == 9
    M = oracle.dbse_shapes_matrix(rest, shapes)
    QR_o, tau_o = oracle.householder_qr(M)
    QR, tau = b.packed_qr()
    np.testing.assert_allclose(QR, QR_o, rtol=0, atol=1e-11 * max(np.abs(QR_o).max(), 1e-300))
    np.testing.assert_allclose(tau, tau_o, rtol=0, atol=1e-11)
    w = b.compute_weights(pos, rest)
    w_o = oracle.dbse_weights(QR_o, pos, rest)
    np.testing.assert_allclose(w, w_o, rtol=0, atol=1e-11 * max(np.abs(w_o).max(), 1e-300))
    assert b.is_computed()
    np.testing.assert_array_equal(b.get_weights(), w)
    for wr, dofall, fr in ((None, 0, 1.0), ((0.0, 1.0), 1, 0.5), ((-0.2, 0.3), 1, 0.0)):
        out = b.displace(pos, rest, weightrange=wr, dofalloff=dofall, falloffradius=fr)
        same_w = oracle.dbse_displace(M, w, pos, rest, weightrange=wr, dofalloff=dofall, falloffradius=fr)
        np.testing.assert_array_equal(out, same_w)     # same weights in: bit-exact FP32 loop
        ref = oracle.dbse_displace(M, w_o, pos, rest, weightrange=wr, dofalloff=dofall, falloffradius=fr)
        np.testing.assert_allclose(out, ref, rtol=0, atol=3e-6)
    b.close()


def test_rest_pose_is_a_fixed_point(ctx):
    from facedeform_b200 import DirectBSEdit
    rest, shapes, _ = _case(5, 2000, 12)
    b = DirectBSEdit(ctx, rest, shapes)
    w = b.compute_weights(rest, rest)
    assert np.all(w == 0.0)
    np.testing.assert_array_equal(b.displace(rest, rest), rest)
    b.close()


def test_face_sized_blendshape_set(ctx, oracle):
    """100k points x 48 blendshapes (a 115 MB shapes matrix): QR columns and weights agree with the oracle."""
    from facedeform_b200 import DirectBSEdit
    rest, shapes, pos = _case(7, 100_000, 48)
    b = DirectBSEdit(ctx, rest, shapes)
    M = oracle.dbse_shapes_matrix(rest, shapes)
    QR_o, _ = oracle.householder_qr(M)
    w = b.compute_weights(pos, rest)
    w_o = oracle.dbse_weights(QR_o, pos, rest)
    np.testing.assert_allclose(w, w_o, rtol=0, atol=1e-10 * np.abs(w_o).max())
    out = b.displace(pos, rest, weightrange=(0.0, 1.0))
    np.testing.assert_array_equal(out, oracle.dbse_displace(M, w, pos, rest, weightrange=(0.0, 1.0)))
    b.close()


def test_handles_outlive_their_ctx_safely():
    """fd_ctx_destroy with live handles defers the teardown to the last handle (any destruction order is safe: a
    garbage collector may finalise the ctx before the models it created)."""
    import ctypes as C
    from facedeform_b200 import _lib, make_params
    L = _lib.load()
    ctx = C.c_void_p()
    assert L.fd_ctx_create(C.byref(ctx), -1, None) == 0
    rest, shapes, pos = _case(3, 500, 4)
    h, m = C.c_void_p(), C.c_void_p()
    assert L.fd_dbse_init(ctx, rest.ctypes.data, 500, shapes.ctypes.data, 4, C.byref(h)) == 0
    p = make_params(model=1, radius=0.5, **{"lambda": 0.0})
    assert L.fd_rbf_fit(ctx, C.byref(p), rest.ctypes.data, 40, C.byref(m), None) == 0
    L.fd_ctx_destroy(ctx)                       # both handles still alive
    w = np.empty(4)
    assert L.fd_dbse_compute_weights(h, pos.ctypes.data, rest.ctypes.data, w.ctypes.data) == 0 and np.isfinite(w).all()
    L.fd_model_destroy(m)
    L.fd_dbse_destroy(h)                        # the last handle tears the ctx down
