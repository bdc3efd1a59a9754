"""Every kernel of the path is a function of its input: the same fit / solve / evaluation repeated on one context returns
the same bits (no atomically accumulated floating point, no result that depends on which CTA ran first).  A difference here
is a race -- this is the test that would have caught the fused LU's diagonal write-back race of round 2.
Models are closed between repetitions, so the pool hands the same blocks out again in a different role."""
import numpy as np
import pytest

from facedeform_b200 import synth

pytestmark = pytest.mark.gpu
KERNELS = ["gaussian", "multiquadric", "thin_plate"]


@pytest.fixture(scope="module")
def ctx():
    from facedeform_b200 import Context
    c = Context()
    yield c
    c.close()


def _run(ctx, N, F, V, kernel, reps, model=1, **kw):
    from facedeform_b200 import make_params
    rig = synth.control_rig(N)
    deform = synth.deformed_rig(rig, F)
    mesh = synth.face_mesh(V, topology=False)
    R = synth.default_radius(KERNELS[kernel], rig.spacing)
    lam = kw.pop("lam", 0.0)
    p = make_params(model=model, term=kw.pop("term", 0), kernel=kernel, radius=R, **{"lambda": lam}, **kw)
    first = None
    for _ in range(reps):
        m = ctx.fit(p, rig.rest).solve(deform)
        out, fall = m.eval(mesh.P)
        W = m.weights()[0]
        kern = m.report().eval_kernel
        m.close()
        if first is None:
            first = (out, fall, W, kern)
            assert np.isfinite(out).all()
        else:
            assert kern == first[3]
            np.testing.assert_array_equal(W, first[2])
            np.testing.assert_array_equal(out, first[0])
            np.testing.assert_array_equal(fall, first[1])


# N: cluster LU (<= 512), cooperative LU with outer blocks of 32 / 128 / 256; F: one frame, the slab solve, the blocked solve
@pytest.mark.parametrize("N,F,V", [(64, 3, 3000), (300, 1, 20_000), (300, 240, 20_000), (1500, 40, 10_000), (2500, 16, 5000),
                                   (4096, 120, 20_000)])
@pytest.mark.parametrize("prec,path", [(0, 0), (1, 2), (1, 1), (2, 0)])
def test_gaussian_pipeline_repeats_bit_for_bit(ctx, N, F, V, prec, path):
    _run(ctx, N, F, V, 0, 4 if N < 4096 else 6, eval_precision=prec, eval_path=path)


@pytest.mark.parametrize("kernel", [1, 2])
@pytest.mark.parametrize("N,F,V", [(300, 1, 20_000), (700, 30, 10_000), (3600, 20, 5000)])
def test_null_space_pipeline_repeats_bit_for_bit(ctx, kernel, N, F, V):
    _run(ctx, N, F, V, kernel, 4)


@pytest.mark.parametrize("kw", [dict(model=0), dict(lam=1e-3), dict(term=1), dict(term=2), dict(factor_precision=1)])
def test_other_factorisations_repeat_bit_for_bit(ctx, kw):
    """QNN radii (pivoted LU), smoothing, constant / no polynomial term, FP32 factorisation + refinement"""
    kw = dict(kw)
    _run(ctx, 1100, 12, 5000, 0, 4, model=kw.pop("model", 1), **kw)


def test_layered_fit_repeats_bit_for_bit(ctx):
    _run(ctx, 500, 6, 5000, 0, 3, fidelity=1, layers=3)


def test_per_cook_inverse_repeats_bit_for_bit(ctx):
    """<= 8 right-hand sides against a cached factorisation: the explicit inverse is built on the second solve"""
    from facedeform_b200 import make_params
    rig = synth.control_rig(3600)
    R = synth.default_radius("gaussian", rig.spacing)
    p = make_params(model=1, term=0, kernel=0, radius=R, eval_precision=2, **{"lambda": 0.0})
    runs = []
    for _ in range(3):
        m = ctx.fit(p, rig.rest)
        ws = []
        for cook in range(4):
            m.solve(synth.deformed_rig(rig, 1, seed=cook))
            ws.append(m.weights()[0])
        m.close()
        runs.append(ws)
    for ws in runs[1:]:
        for a, b in zip(ws, runs[0]):
            np.testing.assert_array_equal(a, b)


@pytest.mark.parametrize("P,S", [(5000, 40), (200_000, 24)])
def test_direct_blendshape_edit_repeats_bit_for_bit(ctx, P, S):
    from facedeform_b200 import DirectBSEdit
    rng = np.random.default_rng(3)
    rest = rng.standard_normal((P, 3)).astype(np.float32)
    shapes = (rest[None] + 0.1 * rng.standard_normal((S, P, 3))).astype(np.float32)
    pos = (rest + 0.05 * rng.standard_normal((P, 3))).astype(np.float32)
    first = None
    for _ in range(4):
        d = DirectBSEdit(ctx, rest, shapes)
        w = d.compute_weights(pos, rest)
        out = d.displace(pos, rest, weightrange=(0.0, 1.0))
        d.close()
        if first is None:
            first = (w, out)
        else:
            np.testing.assert_array_equal(w, first[0])
            np.testing.assert_array_equal(out, first[1])
