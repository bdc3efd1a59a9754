"""Pins the CPU oracle (oracle/fd_oracle.c): the reference has no golden vectors (parity unpinned), so the
dense RBF math is cross-checked against scipy.interpolate.RBFInterpolator and analytic properties
(SURVEY.md section 4, items 1-2)."""
import numpy as np
import pytest
from scipy.interpolate import RBFInterpolator

from facedeform_b200 import synth

SCIPY_KERNEL = {0: "gaussian", 1: "multiquadric", 2: "thin_plate_spline"}
DEGREE = {0: 1, 1: 0, 2: -1}


def _problem(N=48, F=2, V=500):
    rig = synth.control_rig(N)
    deform = synth.deformed_rig(rig, F)
    mesh = synth.face_mesh(V, topology=False)
    return rig, deform, mesh


@pytest.mark.filterwarnings("ignore:.degree. should not be below")
@pytest.mark.parametrize("kernel", [0, 1, 2])
@pytest.mark.parametrize("term", [0, 1, 2])
def test_matches_scipy(oracle, kernel, term):
    if kernel == 2 and term == 2:
        pytest.skip("thin plate without a polynomial block is not guaranteed non-singular")
    rig, deform, mesh = _problem()
    R = synth.default_radius(SCIPY_KERNEL[kernel].replace("_spline", ""), rig.spacing)
    p = oracle.make_params(model=oracle.MODEL_ML, term=term, kernel=kernel, radius=R, **{"lambda": 0.0})
    st, rad, W = oracle.fit(p, rig.rest, deform)
    assert st == 1
    ours = oracle.evaluate_raw(p, rig.rest, rad, W, mesh.P)
    delta = (deform - rig.rest[None]).astype(np.float64)          # float32 subtract, like the SOP
    y = rig.rest.astype(np.float64)
    for f in range(deform.shape[0]):
        # scipy's multiquadric is -sqrt(1+(eps r)^2) = -(1/R) sqrt(r^2+R^2): scaling/sign leave the interpolant
        # unchanged when the polynomial block is the same; with degree -1 (term zero) it also holds.
        ip = RBFInterpolator(y, delta[f], kernel=SCIPY_KERNEL[kernel], epsilon=1.0 / R, degree=DEGREE[term])
        ref = ip(mesh.P.astype(np.float64))
        np.testing.assert_allclose(ours[:, 3 * f:3 * f + 3], ref, rtol=0, atol=2e-8)


def test_smoothing_matches_scipy(oracle):
    rig, deform, mesh = _problem()
    R = 2 * rig.spacing
    lam = np.float32(0.05)
    p = oracle.make_params(model=1, term=0, kernel=0, radius=R, **{"lambda": float(lam)})
    st, rad, W = oracle.fit(p, rig.rest, deform)
    assert st == 1
    ours = oracle.evaluate_raw(p, rig.rest, rad, W, mesh.P)
    delta = (deform - rig.rest[None]).astype(np.float64)
    ip = RBFInterpolator(rig.rest.astype(np.float64), delta[0], kernel="gaussian", epsilon=1.0 / R, degree=1,
                         smoothing=float(lam))
    np.testing.assert_allclose(ours[:, :3], ip(mesh.P.astype(np.float64)), atol=2e-8)


@pytest.mark.parametrize("kernel", [1, 2])
def test_smoothing_of_the_conditionally_definite_kernels_matches_scipy(oracle, kernel):
    """multiquadric: +sqrt(r^2 + R^2) is conditionally negative definite, the shift is -lambda -- the system scipy solves
    with its kernel -sqrt(1 + (r/R)^2) and smoothing lambda / R; thin plate (conditionally positive): + lambda."""
    rig, deform, mesh = _problem()
    R = float(np.float32(1.5 * rig.spacing))
    lam = float(np.float32(0.02))
    p = oracle.make_params(model=1, term=0, kernel=kernel, radius=R, **{"lambda": lam})
    st, rad, W = oracle.fit(p, rig.rest, deform)
    assert st == 1
    ours = oracle.evaluate_raw(p, rig.rest, rad, W, mesh.P)
    delta = (deform - rig.rest[None]).astype(np.float64)
    if kernel == 1:
        ip = RBFInterpolator(rig.rest.astype(np.float64), delta[0], kernel="multiquadric", epsilon=1.0 / R, degree=1,
                             smoothing=lam / R)
    else:
        ip = RBFInterpolator(rig.rest.astype(np.float64), delta[0], kernel="thin_plate_spline", degree=1, smoothing=lam)
    np.testing.assert_allclose(ours[:, :3], ip(mesh.P.astype(np.float64)), atol=5e-8)


@pytest.mark.parametrize("kernel", [0, 1, 2])
def test_interpolates_control_points(oracle, kernel):
    rig, deform, _ = _problem(N=100, F=3)
    p = oracle.make_params(model=1, term=0, kernel=kernel, radius=1.5 * rig.spacing, **{"lambda": 0.0})
    st, rad, W = oracle.fit(p, rig.rest, deform)
    assert st == 1
    at = oracle.evaluate_raw(p, rig.rest, rad, W, rig.rest)
    delta = (deform - rig.rest[None]).transpose(1, 0, 2).reshape(rig.rest.shape[0], -1)
    np.testing.assert_allclose(at, delta, atol=1e-9)


def test_zero_delta_is_identity(oracle):
    rig, _, mesh = _problem()
    p = oracle.make_params(model=1, term=0, kernel=0, radius=2 * rig.spacing, **{"lambda": 0.0})
    st, rad, W = oracle.fit(p, rig.rest, rig.rest[None].copy())
    assert st == 1 and np.all(W == 0)
    out, fall = oracle.evaluate(p, rig.rest, rad, W, mesh.P)
    assert np.array_equal(out[0], mesh.P) and np.all(fall == 1.0)


@pytest.mark.parametrize("kernel", [0, 1, 2])
def test_affine_reproduction(oracle, kernel):
    """d_i = A c_i + b with the linear term => w = 0 and f(v) = A v + b everywhere."""
    rig, _, mesh = _problem(N=60)
    rng = np.random.default_rng(7)
    A = 0.1 * rng.standard_normal((3, 3))
    b = 0.1 * rng.standard_normal(3)
    c = rig.rest.astype(np.float64)
    deform = (c + c @ A.T + b).astype(np.float32)[None]
    p = oracle.make_params(model=1, term=0, kernel=kernel, radius=2 * rig.spacing, **{"lambda": 0.0})
    st, rad, W = oracle.fit(p, rig.rest, deform)
    assert st == 1
    delta = (deform[0] - rig.rest).astype(np.float64)             # what the SOP actually fits (float32 subtract)
    Afit, *_ = np.linalg.lstsq(np.c_[np.ones(len(c)), c], delta, rcond=None)
    ours = oracle.evaluate_raw(p, rig.rest, rad, W, mesh.P)
    want = np.c_[np.ones(len(mesh.P)), mesh.P.astype(np.float64)] @ Afit
    np.testing.assert_allclose(ours, want, atol=5e-6)             # residual of the float32 subtraction only
    assert np.abs(W[:len(c)]).max() < 1e-3


def test_permutation_equivariance_and_frame_batching(oracle):
    rig, deform, mesh = _problem(N=40, F=4)
    p = oracle.make_params(model=1, term=0, kernel=1, radius=rig.spacing, **{"lambda": 0.0})
    st, rad, W = oracle.fit(p, rig.rest, deform)
    base = oracle.evaluate_raw(p, rig.rest, rad, W, mesh.P)
    perm = np.random.default_rng(0).permutation(rig.rest.shape[0])
    st2, rad2, W2 = oracle.fit(p, rig.rest[perm], deform[:, perm])
    np.testing.assert_allclose(oracle.evaluate_raw(p, rig.rest[perm], rad2, W2, mesh.P), base, atol=1e-9)
    for f in range(4):                                            # batched F frames == F independent solves
        _, _, Wf = oracle.fit(p, rig.rest, deform[f])
        np.testing.assert_allclose(Wf, W[:, 3 * f:3 * f + 3], rtol=0, atol=1e-12 * max(1.0, np.abs(W).max()))


def test_qnn_radii_rule(oracle):
    rig, deform, mesh = _problem(N=80)
    p = oracle.make_params(model=oracle.MODEL_QNN, term=0, kernel=0, qcoef=1.5, zcoef=1.25, **{"lambda": 0.0})
    st, rad = oracle.radii(p, rig.rest)
    c = rig.rest.astype(np.float64)
    d = np.linalg.norm(c[:, None] - c[None], axis=-1)
    np.fill_diagonal(d, np.inf)
    want = 1.5 * d.min(axis=1)
    want = np.minimum(want, 1.25 * np.sort(want)[len(want) // 2])
    assert st == 0
    np.testing.assert_allclose(rad, want, rtol=1e-13)
    # per-centre radii give a non-symmetric system K_ij = exp(-d_ij^2 / R_j^2); it still interpolates
    A = oracle.assemble(p, rig.rest, rad)
    assert not np.allclose(A[:80, :80], A[:80, :80].T)
    st, rad, W = oracle.fit(p, rig.rest, deform)
    assert st == 1
    at = oracle.evaluate_raw(p, rig.rest, rad, W, rig.rest)
    delta = (deform - rig.rest[None]).transpose(1, 0, 2).reshape(80, -1)
    np.testing.assert_allclose(at, delta, atol=1e-9)


def test_duplicate_centres_fail(oracle):
    rig, deform, _ = _problem(N=20)
    rest = rig.rest.copy()
    rest[5] = rest[4]
    p = oracle.make_params(model=oracle.MODEL_QNN, term=0, kernel=0)
    st, _, _ = oracle.fit(p, rest, deform)
    assert st == -5                                               # zero QNN radius
    p = oracle.make_params(model=oracle.MODEL_ML, term=0, kernel=0, radius=0.3, **{"lambda": 0.0})
    st, _, _ = oracle.fit(p, rest, deform)
    assert st == -3                                               # singular ("Can't solve the problem.")


def test_lu_against_numpy(oracle):
    rng = np.random.default_rng(3)
    A = rng.standard_normal((70, 70))
    B = rng.standard_normal((70, 5))
    st, LU, piv = oracle.lu_factor(A)
    assert st == 0
    X = oracle.lu_solve(LU, piv, B)
    np.testing.assert_allclose(X, np.linalg.solve(A, B), rtol=1e-9, atol=1e-11)


def test_pack_and_clamps(oracle):
    rig, deform, _ = _problem(N=10)
    packed = oracle.pack(rig.rest, deform[0])
    assert np.array_equal(packed[:, :3], rig.rest.astype(np.float64))
    assert np.array_equal(packed[:, 3:], (deform[0] - rig.rest).astype(np.float64))
    p = oracle.make_params(qcoef=0.0, zcoef=-1.0, radius=0.0, layers=0, maxedges=0, **{"lambda": 0.0})
    oracle.clamp_params(p)
    assert (p.qcoef, p.zcoef, p.radius, p.layers, p.lambda_, p.maxedges) == pytest.approx((0.1, 0.1, 0.01, 1, 0.01, 1))
    d = oracle.make_params()
    assert (d.model, d.term, d.qcoef, d.zcoef, d.radius, d.layers, d.maxedges) == (0, 0, 1.0, 5.0, 1.0, 4, 4)
