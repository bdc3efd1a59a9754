"""GPU parity: the CUDA path through the C ABI vs the CPU oracle on the same seeded inputs.

Stated tolerance (north_star): max |P_gpu - P_oracle| <= 1e-5 x bounding-box diagonal of the mesh.  Weights of
the FP64 factor/solve are compared with a conditioning-aware bound.  Indices are bit-exact (test_gpu_capture).
"""
import numpy as np
import pytest

from facedeform_b200 import synth

pytestmark = pytest.mark.gpu

REL_TOL = 1e-5  # of the bounding-box diagonal


@pytest.fixture(scope="module")
def ctx():
    from facedeform_b200 import Context
    c = Context()
    yield c
    c.close()


def _oracle_params(oracle, p):
    return oracle.make_params(model=p.model, term=p.term, kernel=p.kernel, qcoef=p.qcoef, zcoef=p.zcoef,
                              radius=p.radius, layers=p.layers, tangent=p.tangent, maxedges=p.maxedges,
                              dofalloff=p.dofalloff, falloffradius=p.falloffradius, falloffrate=p.falloffrate,
                              **{"lambda": p.lambda_})


def _run(ctx, oracle, N, V, F, kernel, term, model=1, lam=0.0, radius=None, tangent=0, falloff=False,
         eval_precision=0, seed_shift=0, tol=REL_TOL, eval_path=0):
    from facedeform_b200 import make_params
    rig = synth.control_rig(N, seed=synth.SEED_CTRL + seed_shift)
    deform = synth.deformed_rig(rig, F)
    mesh = synth.face_mesh(V, topology=False)
    R = radius if radius is not None else synth.default_radius(["gaussian", "multiquadric", "thin_plate"][kernel], rig.spacing)
    p = make_params(model=model, term=term, kernel=kernel, radius=R, tangent=tangent, dofalloff=int(falloff),
                    falloffrate=1.7, eval_precision=eval_precision, eval_path=eval_path, **{"lambda": lam})
    dist2 = None
    if falloff:
        rng = np.random.default_rng(9)
        dist2 = rng.uniform(0, 1.2 * R * R, V).astype(np.float32)
        dist2[::13] = -1.0
    tu, tv, nn = (mesh.tangentu, mesh.tangentv, mesh.N) if tangent else (None, None, None)
    model_h = ctx.fit(p, rig.rest)
    model_h.solve(deform)
    out, fall = model_h.eval(mesh.P, dist2, tu, tv, nn)
    Wg, Rg = model_h.weights()
    op = _oracle_params(oracle, p)
    st, rad, W = oracle.fit(op, rig.rest, deform)
    assert st == 1
    ref, rfall = oracle.evaluate(op, rig.rest, rad, W, mesh.P, dist2, tu, tv, nn, nthreads=8)
    np.testing.assert_allclose(Rg, rad, rtol=1e-12)
    diag = max(mesh.bbox_diag, 1.0)              # tiny ragged meshes: fall back to the unit scale of the domain
    # the reference's -1 sentinel yields falloff = pow(1 + 1/R^2, rate) >> 1: the displacement and
    # its rounding error are both multiplied by it, so the bound is on the un-amplified displacement
    amp = np.maximum(rfall, 1.0)[None, :, None]
    err = (np.abs(out.astype(np.float64) - ref.astype(np.float64)) / amp).max()
    assert err <= tol * diag, f"max err {err:.3e} > {tol * diag:.3e}"
    np.testing.assert_allclose(fall, rfall, rtol=2e-6, atol=1e-7)
    model_h.close()
    return err / diag, Wg, W


@pytest.mark.parametrize("kernel", [0, 1, 2])
@pytest.mark.parametrize("term", [0, 1, 2])
def test_small_all_kernels_terms(ctx, oracle, kernel, term):
    if kernel == 2 and term == 2:
        pytest.skip("thin plate needs a polynomial block")
    rel, Wg, W = _run(ctx, oracle, N=50, V=3000, F=3, kernel=kernel, term=term)
    scale = np.abs(W).max()
    np.testing.assert_allclose(Wg, W, rtol=0, atol=1e-7 * scale)


def test_config_c1(ctx, oracle):
    """BASELINE.json configs[0]: 64 control points, 10k vertices, Gaussian, single frame."""
    rel, Wg, W = _run(ctx, oracle, N=64, V=10_000, F=1, kernel=0, term=0)
    assert rel < 1e-5


@pytest.mark.parametrize("F", [1, 2, 3, 4, 5, 9])
def test_frame_chunk_edges(ctx, oracle, F):
    _run(ctx, oracle, N=70, V=1000, F=F, kernel=0, term=0)


@pytest.mark.parametrize("V", [1, 31, 255, 256, 257, 511, 512, 513, 1025])
def test_ragged_vertex_counts(ctx, oracle, V):
    _run(ctx, oracle, N=33, V=V, F=2, kernel=0, term=0)


@pytest.mark.parametrize("N", [1, 2, 5, 31, 32, 33, 255, 256, 257, 300])
def test_ragged_control_counts(ctx, oracle, N):
    # N below the polynomial degree of freedom makes the linear block singular: use the constant term there
    term = 0 if N >= 5 else 1
    kernel = 0
    _run(ctx, oracle, N=N, V=700, F=2, kernel=kernel, term=term, radius=0.5 if N < 5 else None)


def test_epilogue_tangent_falloff(ctx, oracle):
    _run(ctx, oracle, N=64, V=5000, F=2, kernel=0, term=0, tangent=1, falloff=True)
    _run(ctx, oracle, N=64, V=5000, F=1, kernel=1, term=0, tangent=1, falloff=True)


def test_qnn_model_and_lambda(ctx, oracle):
    _run(ctx, oracle, N=120, V=4000, F=2, kernel=0, term=0, model=0)
    _run(ctx, oracle, N=120, V=4000, F=2, kernel=0, term=0, model=1, lam=0.05)


def test_mid_size_c2_shape_subsample(ctx, oracle):
    """C2 control rig (N=256) and frame batching at a vertex count the oracle finishes in seconds."""
    _run(ctx, oracle, N=256, V=4096, F=24, kernel=0, term=0)


@pytest.mark.parametrize("kernel", [1, 2])
def test_mid_size_c3_kernels_fp64_eval(ctx, oracle, kernel):
    """C3 kernels (multiquadric / thin plate + affine block), N=1024: FP64 evaluation keeps 1e-5 of the diagonal."""
    _run(ctx, oracle, N=1024, V=3000, F=1, kernel=kernel, term=0)


def test_fp32_eval_of_multiquadric_states_its_looser_tolerance(ctx, oracle):
    """FD_EVAL_FP32 on a globally supported kernel suffers cancellation: stated tolerance 5e-4 of the diagonal."""
    _run(ctx, oracle, N=1024, V=3000, F=1, kernel=1, term=0, eval_precision=1, tol=5e-4)


def test_errors_mirror_the_sop(ctx):
    from facedeform_b200 import FdError, make_params
    rig = synth.control_rig(20)
    deform = synth.deformed_rig(rig, 1)
    p = make_params(model=1, radius=0.3, **{"lambda": 0.0})
    m = ctx.fit(p, rig.rest)
    with pytest.raises(FdError) as e:           # point-count mismatch, SOP_FaceDeform.cpp:231-234
        m.solve(deform[:, :19])
    assert e.value.status == 2
    with pytest.raises(FdError) as e:           # eval before solve
        m.eval(rig.rest)
    assert e.value.status == 9
    rest = rig.rest.copy()
    rest[7] = rest[3]
    with pytest.raises(FdError) as e:           # duplicate centres: singular system, :365-368
        ctx.fit(p, rest)
    assert e.value.status == 4
    with pytest.raises(FdError) as e:           # QNN zero radius
        ctx.fit(make_params(model=0), rest)
    assert e.value.status == 4
    with pytest.raises(FdError) as e:
        ctx.fit(make_params(kernel=7), rig.rest)
    assert e.value.status == 1


def test_zero_delta_is_identity_and_interpolation(ctx):
    from facedeform_b200 import make_params
    rig = synth.control_rig(90)
    deform = synth.deformed_rig(rig, 2)
    p = make_params(model=1, radius=2 * rig.spacing, **{"lambda": 0.0})
    m = ctx.fit(p, rig.rest)
    m.solve(rig.rest[None].copy())
    mesh = synth.face_mesh(2000, topology=False)
    out, fall = m.eval(mesh.P)
    assert np.array_equal(out[0], mesh.P) and np.all(fall == 1.0)
    m.solve(deform)                               # same factorisation, new frames: f(c_i) = c_i + delta_i
    out, _ = m.eval(rig.rest)
    np.testing.assert_allclose(out, deform, atol=1e-5 * 2.86)
    m.close()


def test_device_pointer_entry_points(ctx, oracle):
    """fd_rbf_*_dev on torch CUDA tensors gives the same bits as the host-pointer entry points."""
    import torch
    from facedeform_b200 import Context, make_params
    rig = synth.control_rig(100)
    deform = synth.deformed_rig(rig, 3)
    mesh = synth.face_mesh(5000, topology=False)
    p = make_params(model=1, radius=2 * rig.spacing, **{"lambda": 0.0})
    m = ctx.fit(p, rig.rest)
    m.solve(deform)
    host_out, host_fall = m.eval(mesh.P)
    c2 = Context(stream=torch.cuda.current_stream().cuda_stream)
    dm = c2.fit(p, torch.from_numpy(rig.rest).cuda())
    dm.solve(torch.from_numpy(deform).cuda())
    dm.report()
    dout, dfall = dm.eval(torch.from_numpy(mesh.P).cuda())
    torch.cuda.synchronize()
    assert np.array_equal(dout.cpu().numpy(), host_out) and np.array_equal(dfall.cpu().numpy(), host_fall)
    dm.close()
    m.close()
    c2.close()


# ---- tensor-core evaluation path (tcgen05, fd_eval_tc.cu): same oracle, same tolerance --------------------------

@pytest.mark.parametrize("F", [1, 5, 16, 79, 80, 81, 160, 161])
def test_tensor_path_frame_blocks(ctx, oracle, F):
    """column blocks of 240 = 80 frames: partial, exact and multi-block cases; forced FD_PATH_TENSOR."""
    _run(ctx, oracle, N=70, V=1024, F=F, kernel=0, term=0, eval_path=2)


@pytest.mark.parametrize("V", [1, 255, 256, 257, 1023, 4098])
def test_tensor_path_ragged_vertices(ctx, oracle, V):
    """V % 4 != 0 takes the scalar store path, partial 256-vertex tiles are masked."""
    _run(ctx, oracle, N=40, V=V, F=20, kernel=0, term=0, eval_path=2)


@pytest.mark.parametrize("N", [1, 5, 27, 28, 29, 60, 61, 255, 256, 257])
def test_tensor_path_ragged_centres(ctx, oracle, N):
    """K padding: centres + 4 affine rows rounded up to 32 per stage."""
    term = 0 if N >= 5 else 1
    _run(ctx, oracle, N=N, V=1024, F=18, kernel=0, term=term, radius=0.5 if N < 5 else None, eval_path=2)


@pytest.mark.parametrize("term", [0, 1, 2])
def test_tensor_path_terms_and_epilogue(ctx, oracle, term):
    _run(ctx, oracle, N=64, V=4096, F=30, kernel=0, term=term, tangent=1, falloff=True, eval_path=2)


def test_tensor_path_c2_shape(ctx, oracle):
    """C2: 256 control points x 240 frames (FD_PATH_AUTO selects the tensor path), vertices subsampled for the oracle."""
    rel, _, _ = _run(ctx, oracle, N=256, V=4096, F=240, kernel=0, term=0)
    assert rel < 1e-5


def test_tensor_path_matches_simt_closely(ctx):
    from facedeform_b200 import make_params
    rig = synth.control_rig(256)
    deform = synth.deformed_rig(rig, 48)
    mesh = synth.face_mesh(20_000, topology=False)
    outs = []
    for path in (1, 2):
        p = make_params(model=1, radius=2 * rig.spacing, eval_path=path, **{"lambda": 0.0})
        m = ctx.fit(p, rig.rest).solve(deform)
        outs.append(m.eval(mesh.P)[0])
        m.close()
    assert np.abs(outs[0] - outs[1]).max() <= 1e-5 * mesh.bbox_diag


def test_tensor_path_other_kernels_fp32(ctx, oracle):
    """multiquadric / thin plate through the tensor path in FP32 mode: the stated looser FP32 tolerance applies."""
    _run(ctx, oracle, N=256, V=2048, F=20, kernel=1, term=0, eval_precision=1, eval_path=2, tol=5e-4)
    _run(ctx, oracle, N=256, V=2048, F=20, kernel=2, term=0, eval_precision=1, eval_path=2, tol=5e-4)


def test_receiver_model_matches_root_bit_for_bit(ctx):
    """multi-GPU plumbing on one GPU: a receiver model fed the root's weight and radii blocks (what NCCL broadcasts in
    shard.broadcast_model) evaluates bit-identically to the root."""
    import torch
    from facedeform_b200 import make_params, shard
    rig = synth.control_rig(96)
    deform = synth.deformed_rig(rig, 20)
    mesh = synth.face_mesh(3000, topology=False)
    for model in (0, 1):
        p = make_params(model=model, radius=2 * rig.spacing, **{"lambda": 0.0})
        root = ctx.fit(p, rig.rest).solve(deform)
        want, wfall = root.eval(mesh.P)
        recv = ctx.receiver(p, rig.rest, 20)
        assert recv.info() == root.info()
        for get in ("weights_dev", "radii_dev"):
            (sp, sb), (dp, db) = getattr(root, get)(), getattr(recv, get)()
            assert sb == db
            ctx.synchronize()
            shard.device_view(dp, db).copy_(shard.device_view(sp, sb))
        torch.cuda.synchronize()
        recv.commit_weights()
        got, gfall = recv.eval(mesh.P)
        assert np.array_equal(got, want) and np.array_equal(gfall, wfall)
        from facedeform_b200 import FdError
        with pytest.raises(FdError) as e:           # a receiver holds no factorisation
            recv.solve(deform)
        assert e.value.status == 9
        root.close()
        recv.close()


# ---- FP32 factorisation + FP64 iterative refinement (fd_params.factor_precision = 1, BASELINE config 4) -------------

@pytest.mark.parametrize("kernel,N,rscale", [(0, 512, 1.5), (0, 1000, 2.0), (1, 300, 1.0), (2, 257, 1.0)])
def test_fp32_refinement_reaches_the_fp64_solution(ctx, kernel, N, rscale):
    """FP32 LU (pivoted for multiquadric / thin plate, unpivoted for the SPD Gaussian) + FP64 residual sweeps: the
    weights agree with the FP64 factorisation and the report carries the sweep count and the final residual."""
    from facedeform_b200 import make_params
    rig = synth.control_rig(N)
    deform = synth.deformed_rig(rig, 3)
    W = []
    for fp in (0, 1):
        p = make_params(model=1, term=0, kernel=kernel, radius=rscale * rig.spacing, factor_precision=fp,
                        **{"lambda": 0.0})
        m = ctx.fit(p, rig.rest).solve(deform)
        rep = m.report()
        assert rep.terminationtype == 1
        if fp:
            assert 1 <= rep.iterationscount <= 40 and 0.0 <= rep.residual <= 1e-9
        else:
            assert rep.iterationscount == 0 and rep.residual == 0.0
        W.append(m.weights()[0])
        out, _ = m.eval(rig.rest)
        assert np.abs(out - deform).max() <= 1e-5 * 2.9      # interpolation at the control points
        m.close()
    assert np.abs(W[0] - W[1]).max() <= 1e-6 * np.abs(W[0]).max()


def test_fp32_refinement_reports_non_convergence(ctx):
    """a Gaussian system far beyond cond ~ 2^24 cannot be refined from FP32 factors: the solve must say so
    (terminationtype -4 -> FD_E_SINGULAR, "Can't solve the problem."), never return unconverged weights silently."""
    from facedeform_b200 import FdError, make_params
    rig = synth.control_rig(1024)
    deform = synth.deformed_rig(rig, 1)
    p = make_params(model=1, term=0, kernel=0, radius=8.0 * rig.spacing, factor_precision=1, **{"lambda": 0.0})
    m = ctx.fit(p, rig.rest)
    try:
        m.solve(deform)
        rep = m.report()
        assert rep.terminationtype == 1 and rep.residual <= 1e-9   # it converged after all: then it must be accurate
    except FdError as e:
        assert e.status == 4
        assert m.last_report.terminationtype in (-3, -4)
    finally:
        m.close()


# ---- fidelity = FD_FIDELITY_ALGLIB_V1: two-stage polynomial + Gaussian layers (unverifiable vs ALGLIB) -------

@pytest.mark.parametrize("model,layers,term,F", [(1, 1, 0, 2), (1, 3, 0, 2), (1, 4, 1, 20), (1, 2, 2, 1), (0, 1, 0, 3)])
def test_alglib_v1_like_mode_matches_the_oracle(ctx, oracle, model, layers, term, F):
    from facedeform_b200 import make_params
    rig = synth.control_rig(120)
    deform = synth.deformed_rig(rig, F)
    mesh = synth.face_mesh(3000, topology=False)
    kw = dict(model=model, term=term, kernel=0, radius=3 * rig.spacing, layers=layers, qcoef=1.0, zcoef=5.0)
    p = make_params(fidelity=1, **kw, **{"lambda": 0.01})
    m = ctx.fit(p, rig.rest).solve(deform)
    rep = m.report()
    L = 1 if model == 0 else layers
    assert rep.terminationtype == 1 and rep.n == 120 * L
    out, _ = m.eval(mesh.P)
    Wg, Rg = m.weights()
    op = oracle.make_params(**kw, **{"lambda": 0.01})
    st, cen, rad, W = oracle.fit_v1(op, rig.rest, deform)
    assert st == 1 and Wg.shape == W.shape
    np.testing.assert_allclose(Rg, rad, rtol=1e-12)
    np.testing.assert_allclose(Wg, W, rtol=0, atol=1e-7 * np.abs(W).max())
    ref, _ = oracle.evaluate(op, cen, rad, W, mesh.P, nthreads=8)
    assert np.abs(out.astype(np.float64) - ref).max() <= REL_TOL * mesh.bbox_diag
    m.close()


def test_alglib_v1_like_mode_limits(ctx):
    from facedeform_b200 import FdError, make_params
    rig = synth.control_rig(30)
    with pytest.raises(FdError) as e:      # ALGLIB's v1 unit is Gaussian only
        ctx.fit(make_params(fidelity=1, model=1, kernel=1, radius=0.3), rig.rest)
    assert e.value.status == 8


def test_alglib_v1_like_receiver_matches_root_bit_for_bit(ctx):
    """the layered fit across GPUs: a receiver of N * layers stacked centres fed the root's blocks evaluates identically."""
    import torch
    from facedeform_b200 import make_params, shard
    rig = synth.control_rig(64)
    deform = synth.deformed_rig(rig, 20)
    mesh = synth.face_mesh(2000, topology=False)
    p = make_params(fidelity=1, model=1, term=0, radius=3 * rig.spacing, layers=3, **{"lambda": 0.01})
    root = ctx.fit(p, rig.rest).solve(deform)
    want, _ = root.eval(mesh.P)
    recv = ctx.receiver(p, rig.rest, 20)
    assert recv.info() == root.info() and recv.info()["n_ctrl"] == 192
    for get in ("weights_dev", "radii_dev"):
        (sp, sb), (dp, db) = getattr(root, get)(), getattr(recv, get)()
        assert sb == db
        ctx.synchronize()
        shard.device_view(dp, db).copy_(shard.device_view(sp, sb))
    torch.cuda.synchronize()
    recv.commit_weights()
    got, _ = recv.eval(mesh.P)
    assert np.array_equal(got, want)
    root.close()
    recv.close()


@pytest.mark.parametrize("kernel", [1, 2])
def test_smoothed_multiquadric_and_thin_plate_take_the_nullspace_path(ctx, oracle, kernel):
    """lambda > 0 (the SOP clamps it to >= 0.01): -lambda on the multiquadric's diagonal, +lambda on the thin plate's,
    both keep the reduced matrix definite, so the fused no-pivot LU applies; weights and positions against the oracle."""
    rel, Wg, W = _run(ctx, oracle, N=300, V=2000, F=2, kernel=kernel, term=0, lam=0.02)
    np.testing.assert_allclose(Wg, W, rtol=0, atol=1e-5 * np.abs(W).max())
