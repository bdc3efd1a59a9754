"""The oracle's "ALGLIB v1 like" fit (fdo_fit_v1): what the SOP's rbfsetalgoqnn / rbfsetalgomultilayer + rbfset*term calls
mean (SOP_FaceDeform.cpp:342-361) in the dense restatement.  ALGLIB is absent and unpinned, so these tests pin the
restatement's own definition (two-stage polynomial, layered residual fitting) against numpy, not against ALGLIB."""
import numpy as np
import pytest

from facedeform_b200 import synth


def _numpy_v1(p, rest, deform, layers, radii_of_layer):
    N, F = rest.shape[0], deform.shape[0]
    D = (deform - rest[None]).astype(np.float64).transpose(1, 0, 2).reshape(N, 3 * F)   # FP32 subtract, widened
    np_ = {0: 4, 1: 1, 2: 0}[p.term]
    Pm = np.c_[np.ones(N), rest.astype(np.float64)][:, :np_]
    v = np.linalg.lstsq(Pm, D, rcond=None)[0] if np_ else np.zeros((0, 3 * F))
    R = D - Pm @ v
    Ws = []
    x = rest.astype(np.float64)
    d2 = ((x[:, None, :] - x[None, :, :]) ** 2).sum(-1)
    for k in range(layers):
        rad = radii_of_layer(k)
        K0 = np.exp(-d2 / rad[None, :] ** 2)
        W = np.linalg.solve(K0 + float(p.lambda_) * np.eye(N), R)
        R = R - K0 @ W
        Ws.append(W)
    return np.vstack(Ws + [v]), R


@pytest.mark.parametrize("term", [0, 1, 2])
@pytest.mark.parametrize("layers", [1, 3])
def test_multilayer_matches_the_numpy_statement(oracle, term, layers):
    rig = synth.control_rig(60)
    deform = synth.deformed_rig(rig, 2)
    p = oracle.make_params(model=1, term=term, kernel=0, radius=3 * rig.spacing, layers=layers, **{"lambda": 0.01})
    st, cen, rad, W = oracle.fit_v1(p, rig.rest, deform)
    assert st == 1 and cen.shape == (60 * layers, 3) and W.shape == (60 * layers + {0: 4, 1: 1, 2: 0}[term], 6)
    R0 = float(np.float32(3 * rig.spacing))
    ref, _ = _numpy_v1(p, rig.rest, deform, layers, lambda k: np.full(60, R0 / 2 ** k))
    np.testing.assert_allclose(W, ref, rtol=0, atol=1e-8 * np.abs(ref).max())
    np.testing.assert_array_equal(cen, np.tile(rig.rest, (layers, 1)))
    np.testing.assert_allclose(rad, np.repeat(R0 / 2.0 ** np.arange(layers), 60), rtol=1e-15)


def test_qnn_is_one_layer_with_per_centre_radii(oracle):
    rig = synth.control_rig(50)
    deform = synth.deformed_rig(rig, 1)
    p = oracle.make_params(model=0, term=0, kernel=0, qcoef=1.0, zcoef=5.0, **{"lambda": 0.01})
    st, cen, rad, W = oracle.fit_v1(p, rig.rest, deform)
    assert st == 1 and cen.shape == (50, 3)
    s2, rq = oracle.radii(p, rig.rest)
    np.testing.assert_allclose(rad, rq, rtol=1e-15)
    ref, _ = _numpy_v1(p, rig.rest, deform, 1, lambda k: rq)
    np.testing.assert_allclose(W, ref, rtol=0, atol=1e-8 * np.abs(ref).max())


def test_layers_shrink_the_residual_at_the_control_points(oracle):
    rig = synth.control_rig(80)
    deform = synth.deformed_rig(rig, 1)
    errs = []
    for layers in (1, 2, 4):
        p = oracle.make_params(model=1, term=0, kernel=0, radius=3 * rig.spacing, layers=layers, **{"lambda": 0.01})
        st, cen, rad, W = oracle.fit_v1(p, rig.rest, deform)
        out, _ = oracle.evaluate(p, cen, rad, W, rig.rest)         # the stacked model through the ordinary evaluation
        errs.append(np.abs(out - deform).max())
    assert errs[0] > errs[1] > errs[2]


def test_affine_deltas_are_taken_by_the_polynomial_stage(oracle):
    """deltas = A c + b: the least-squares polynomial reproduces them, every layer sees a zero residual."""
    rig = synth.control_rig(40)
    A = np.array([[0.02, 0.01, 0.0], [0.0, -0.03, 0.01], [0.01, 0.0, 0.02]], np.float32)
    deform = (rig.rest + rig.rest @ A.T + np.float32(0.05))[None]
    p = oracle.make_params(model=1, term=0, kernel=0, radius=2 * rig.spacing, layers=2, **{"lambda": 0.01})
    st, cen, rad, W = oracle.fit_v1(p, rig.rest, deform)
    assert st == 1
    assert np.abs(W[:80]).max() <= 1e-5 and np.abs(W[80] - 0.05).max() <= 1e-6
