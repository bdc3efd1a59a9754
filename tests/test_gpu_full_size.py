"""BASELINE.json's full-size configurations on the GPU, checked through size-independent properties and oracle
subsamples (the oracle evaluates a few thousand vertices in seconds; the fit at N = 2048 takes ~1 s):
  C2  256 control points, 100k vertices, 240 frames (tensor-core path)
  C3  2048 control points, 1M vertices, multiquadric / thin plate + affine block (FP64 evaluation)
  C4  8192 control points, factorisation-dominated: interpolation property of the FP64 LU + solve
  C5  (shape only, scaled to one GPU's test time) 4096 control points, wide frame batch, vertex-sharded
"""
import numpy as np
import pytest

from facedeform_b200 import shard, synth

pytestmark = pytest.mark.gpu
TOL = 1e-5


@pytest.fixture(scope="module")
def ctx():
    from facedeform_b200 import Context
    c = Context()
    yield c
    c.close()


def _oracle_subsample(oracle, p, rig, deform, mesh, out, idx, frames=None):
    op = oracle.make_params(model=p.model, term=p.term, kernel=p.kernel, radius=p.radius, **{"lambda": p.lambda_})
    d = deform if frames is None else deform[frames]
    st, rad, W = oracle.fit(op, rig.rest, d)
    assert st == 1
    ref, _ = oracle.evaluate(op, rig.rest, rad, W, mesh.P[idx], nthreads=8)
    got = out[:, idx] if frames is None else out[frames][:, idx]
    return np.abs(got.astype(np.float64) - ref).max()


def test_c2_full(ctx, oracle):
    from facedeform_b200 import make_params
    cfg = synth.CONFIGS["C2"]
    rig = synth.control_rig(cfg["N"])
    deform = synth.deformed_rig(rig, cfg["F"])
    mesh = synth.face_mesh(cfg["V"], topology=False)
    p = make_params(model=1, radius=synth.default_radius("gaussian", rig.spacing), **{"lambda": 0.0})
    m = ctx.fit(p, rig.rest).solve(deform)
    out, fall = m.eval(mesh.P)
    assert out.shape == (240, 100_000, 3) and np.isfinite(out).all() and np.all(fall == 1.0)
    idx = np.random.default_rng(0).choice(cfg["V"], 1500, replace=False)
    frames = [0, 79, 80, 159, 160, 239]                       # both sides of every 80-frame column block
    assert _oracle_subsample(oracle, p, rig, deform, mesh, out, idx, frames) <= TOL * mesh.bbox_diag
    # vertex-range sharding: evaluating the ranges separately and concatenating is bit-identical
    parts = []
    for r in range(3):
        b, e = shard.vertex_range(cfg["V"], r, 3)
        parts.append(m.eval(mesh.P[b:e])[0])
    assert np.array_equal(np.concatenate(parts, axis=1), out)
    # interpolation at the control points
    at, _ = m.eval(rig.rest)
    assert np.abs(at - deform).max() <= TOL * mesh.bbox_diag
    m.close()


@pytest.mark.parametrize("kernel", ["multiquadric", "thin_plate"])
def test_c3_full(ctx, oracle, kernel):
    from facedeform_b200 import make_params
    cfg = synth.CONFIGS["C3"]
    rig = synth.control_rig(cfg["N"])
    deform = synth.deformed_rig(rig, 1)
    mesh = synth.face_mesh(cfg["V"], topology=False)
    p = make_params(model=1, term=0, kernel=synth.KERNELS[kernel], radius=synth.default_radius(kernel, rig.spacing),
                    **{"lambda": 0.0})
    m = ctx.fit(p, rig.rest).solve(deform)
    out, _ = m.eval(mesh.P)
    assert out.shape == (1, 1_000_000, 3) and np.isfinite(out).all()
    idx = np.random.default_rng(1).choice(cfg["V"], 1500, replace=False)
    assert _oracle_subsample(oracle, p, rig, deform, mesh, out, idx) <= TOL * mesh.bbox_diag
    at, _ = m.eval(rig.rest)
    assert np.abs(at - deform).max() <= TOL * mesh.bbox_diag
    b, e = shard.vertex_range(cfg["V"], 5, 8)                 # one of eight ranks
    assert np.array_equal(m.eval(mesh.P[b:e])[0], out[:, b:e])
    m.close()


def test_c4_factorisation_8192(ctx):
    """8192 control points + affine block: the FP64 LU (16-CTA cluster panels) and the solve reproduce the control
    displacements (interpolation), the pivots are sane and the Gaussian weights are finite."""
    from facedeform_b200 import make_params
    rig = synth.control_rig(8192)
    deform = synth.deformed_rig(rig, 1)
    p = make_params(model=1, radius=1.5 * rig.spacing, **{"lambda": 0.0})
    m = ctx.fit(p, rig.rest)
    rep = m.report()
    assert rep.terminationtype == 1 and rep.n == 8192 and rep.npoly == 4 and 0 < rep.min_pivot <= rep.max_pivot
    m.solve(deform)
    W, R = m.weights()
    assert np.isfinite(W).all() and W.shape == (8196, 3)
    at, _ = m.eval(rig.rest)
    assert np.abs(at - deform).max() <= TOL * 2.86
    # side conditions of the saddle-point system: P^T w = 0
    c = np.c_[np.ones(8192), rig.rest.astype(np.float64)]
    assert np.abs(c.T @ W[:8192]).max() <= 1e-7 * np.abs(W).max() * 8192
    m.close()


def test_c5_shape_sharded(ctx, oracle):
    """C5's shape at a size one test can afford: 4096 control points, 120 frames (1.5 column blocks), 200k vertices
    evaluated as 8 vertex ranges like 8 GPUs would; oracle check on a subsample of two frames."""
    from facedeform_b200 import make_params
    rig = synth.control_rig(4096)
    deform = synth.deformed_rig(rig, 120)
    mesh = synth.face_mesh(200_000, topology=False)
    p = make_params(model=1, radius=synth.default_radius("gaussian", rig.spacing), **{"lambda": 0.0})
    m = ctx.fit(p, rig.rest).solve(deform)
    full, _ = m.eval(mesh.P)
    parts = [m.eval(mesh.P[slice(*shard.vertex_range(200_000, r, 8))])[0] for r in range(8)]
    assert np.array_equal(np.concatenate(parts, axis=1), full)
    idx = np.random.default_rng(2).choice(200_000, 600, replace=False)
    assert _oracle_subsample(oracle, p, rig, deform, mesh, full, idx, [0, 119]) <= TOL * mesh.bbox_diag
    m.close()
