"""GPU tests of the round-2 work: the error bound behind FD_EVAL_AUTO and the sizes the first suite missed
(Gaussian, N = 1024 / 2048), the FP64 tensor-pipe evaluation, the per-solve NaN flag, serialisation, epilogue-only
parameter changes, input validation of fd_capture, several contexts / devices in one process and fd_mgpu_*."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

from facedeform_b200 import synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def ctx():
    from facedeform_b200 import Context
    c = Context()
    yield c
    c.close()


def _oparams(oracle, p):
    return oracle.make_params(model=p.model, term=p.term, kernel=p.kernel, qcoef=p.qcoef, zcoef=p.zcoef, radius=p.radius,
                              layers=p.layers, tangent=p.tangent, maxedges=p.maxedges, dofalloff=p.dofalloff,
                              falloffradius=p.falloffradius, falloffrate=p.falloffrate, **{"lambda": p.lambda_})


_ORACLE_CACHE = {}


def _case(oracle, N, F, V, kernel=0):
    """rig, frames, a V-vertex sample of the mesh and the oracle's positions for it (cached per shape)."""
    key = (N, F, V, kernel)
    if key not in _ORACLE_CACHE:
        from facedeform_b200 import make_params
        rig = synth.control_rig(N)
        deform = synth.deformed_rig(rig, F)
        mesh = synth.face_mesh(100_000, topology=False)
        idx = np.sort(np.random.default_rng(11).choice(mesh.P.shape[0], V, replace=False))
        P = np.ascontiguousarray(mesh.P[idx])
        R = synth.default_radius(["gaussian", "multiquadric", "thin_plate"][kernel], rig.spacing)
        p = make_params(model=1, term=0, kernel=kernel, radius=R, **{"lambda": 0.0})
        st, rad, W = oracle.fit(_oparams(oracle, p), rig.rest, deform)
        assert st == 1
        ref, _ = oracle.evaluate(_oparams(oracle, p), rig.rest, rad, W, P, nthreads=8)
        _ORACLE_CACHE[key] = (rig, deform, P, R, ref, mesh.bbox_diag)
    return _ORACLE_CACHE[key]


# ---- the sizes VERDICT r1 found untested: Gaussian at N = 1024 / 2048, one frame and wide batches -----------------------
@pytest.mark.parametrize("N", [256, 1024, 2048])
@pytest.mark.parametrize("F", [1, 120, 240])
def test_gaussian_auto_meets_the_tolerance_with_margin(ctx, oracle, N, F):
    """FD_EVAL_AUTO: max |P_gpu - P_oracle| over 4096 vertices x all frames stays within the stated 1e-5 x bbox diagonal
    with a margin (<= 0.75e-5 asserted), whichever kernel the measured cancellation selects; wide batches take the exact-digit
    tensor-core kernel and stay within 1e-6."""
    from facedeform_b200 import make_params
    rig, deform, P, R, ref, diag = _case(oracle, N, F, 4096)
    m = ctx.fit(make_params(model=1, term=0, kernel=0, radius=R, **{"lambda": 0.0}), rig.rest).solve(deform)
    out, _ = m.eval(P)
    rep = m.report()
    err = float(np.abs(out.astype(np.float64) - ref).max()) / diag
    print(f"N={N} F={F}: AUTO kernel {rep.eval_kernel} cancellation {rep.cancellation:.3e} err/diag {err:.3e}")
    assert rep.eval_kernel in (1, 3, 4) and rep.eval_inexact == 0
    assert err <= 0.75e-5, f"err/diag {err:.3e} with kernel {rep.eval_kernel}"
    if F >= 16:
        # BASELINE configs[1] among them: the FP32 tensor-core kernel's bound (2.0 x 2^-24 S) exceeds the tolerance there; the
        # exact-digit kernel's (0.15 x 2^-24 S) does not at any of these sizes
        assert rep.eval_kernel == 4 and err <= 1e-6, f"err/diag {err:.3e} with kernel {rep.eval_kernel}"
    m.close()


@pytest.mark.parametrize("N", [256, 1024, 2048])
@pytest.mark.parametrize("path", [1, 2])
def test_fp32_error_follows_the_cancellation_model(ctx, oracle, N, path):
    """forced FP32 (FMA/SFU and tensor cores): the maximum error over 4096 vertices x 120 frames stays below the bound
    coef x 2^-24 x S that FD_EVAL_AUTO decides with (S = fd_report.cancellation), and within the stated 1e-5 at the
    benchmark's N = 256."""
    from facedeform_b200 import make_params
    F = 120
    rig, deform, P, R, ref, diag = _case(oracle, N, F, 4096)
    m = ctx.fit(make_params(model=1, term=0, kernel=0, radius=R, eval_precision=1, eval_path=path, **{"lambda": 0.0}),
                rig.rest).solve(deform)
    out, _ = m.eval(P)
    rep = m.report()
    assert rep.eval_kernel == path
    err = float(np.abs(out.astype(np.float64) - ref).max())
    coef = max(1.1, 1.3 - 0.07 * np.log2(max(N, 64) / 64)) if path == 1 else 2.0
    model = coef * 2.0 ** -24 * rep.cancellation
    print(f"N={N} path={path}: err/diag {err / diag:.3e} model/diag {model / diag:.3e} ratio {err / model:.2f}")
    assert err <= model  # the coefficients are the largest ratios of the calibration runs + 20 %
    if N == 256:
        assert err <= 1e-5 * diag
    m.close()


def test_auto_on_well_conditioned_weights(ctx, oracle):
    """well-conditioned weights (radius = spacing: little cancellation): FD_EVAL_AUTO takes the exact-digit tensor-core kernel
    for a wide batch like everywhere else, and with FD_PATH_TENSOR + FD_EVAL_FP32 the FP16 hi/lo kernel is within the tolerance
    too."""
    from facedeform_b200 import make_params
    N, F, V = 256, 120, 4096
    rig = synth.control_rig(N)
    deform = synth.deformed_rig(rig, F)
    P = np.ascontiguousarray(synth.face_mesh(100_000, topology=False).P[:V])
    p = make_params(model=1, term=0, kernel=0, radius=rig.spacing, **{"lambda": 0.0})
    m = ctx.fit(p, rig.rest).solve(deform)
    out, _ = m.eval(P)
    rep = m.report()
    st, rad, W = oracle.fit(_oparams(oracle, p), rig.rest, deform)
    ref, _ = oracle.evaluate(_oparams(oracle, p), rig.rest, rad, W, P, nthreads=8)
    err = float(np.abs(out.astype(np.float64) - ref).max()) / 2.9
    print(f"radius = spacing: kernel {rep.eval_kernel} cancellation {rep.cancellation:.3e} err/diag {err:.3e}")
    assert rep.eval_kernel == 4 and err <= 2e-7
    m.close()
    p2 = make_params(model=1, term=0, kernel=0, radius=rig.spacing, eval_precision=1, eval_path=2, **{"lambda": 0.0})
    m = ctx.fit(p2, rig.rest).solve(deform)
    out, _ = m.eval(P)
    assert m.report().eval_kernel == 2 and float(np.abs(out.astype(np.float64) - ref).max()) / 2.9 <= 0.75e-5
    m.close()


@pytest.mark.parametrize("kernel", [0, 1, 2])
@pytest.mark.parametrize("F", [16, 40])
def test_fp64_tensor_pipe_evaluation(ctx, oracle, kernel, F):
    """eval_precision = FP64 with 3F >= 48 columns runs k_eval64_mma (DMMA): FP64-accurate against the oracle, and for
    every epilogue option identical to the per-vertex FP64 kernel's semantics."""
    from facedeform_b200 import make_params
    rig, deform, P, R, ref, diag = _case(oracle, 300, F, 2000, kernel)
    m = ctx.fit(make_params(model=1, term=0, kernel=kernel, radius=R, eval_precision=2, **{"lambda": 0.0}), rig.rest).solve(deform)
    out, _ = m.eval(P)
    assert m.report().eval_kernel == 3
    err = float(np.abs(out.astype(np.float64) - ref).max()) / diag
    assert err <= 2e-7, err  # FP32 output rounding: 2^-24 x |P|
    m.close()


def test_fp64_tensor_pipe_epilogue_matches_the_oracle(ctx, oracle):
    from facedeform_b200 import make_params
    N, F, V = 100, 20, 3001  # ragged vertex count, tangent projection, falloff with the -1 sentinel
    rig = synth.control_rig(N)
    deform = synth.deformed_rig(rig, F)
    mesh = synth.face_mesh(V, topology=False)
    R = synth.default_radius("multiquadric", rig.spacing)
    p = make_params(model=1, term=0, kernel=1, radius=R, tangent=1, dofalloff=1, falloffrate=1.7, **{"lambda": 0.0})
    d2 = np.random.default_rng(9).uniform(0, 1.2 * R * R, V).astype(np.float32)
    d2[::13] = -1.0
    m = ctx.fit(p, rig.rest).solve(deform)
    out, fall = m.eval(mesh.P, d2, mesh.tangentu, mesh.tangentv, mesh.N)
    st, rad, W = oracle.fit(_oparams(oracle, p), rig.rest, deform)
    ref, rfall = oracle.evaluate(_oparams(oracle, p), rig.rest, rad, W, mesh.P, d2, mesh.tangentu, mesh.tangentv, mesh.N)
    amp = np.maximum(rfall, 1.0)[None, :, None]
    assert (np.abs(out.astype(np.float64) - ref) / amp).max() <= 1e-6 * max(mesh.bbox_diag, 1.0)
    np.testing.assert_allclose(fall, rfall, rtol=2e-6, atol=1e-7)
    m.close()


# ---- the NaN flag belongs to a solve, not to the model ---------------------------------------------------
@pytest.mark.parametrize("F", [1, 40])
def test_nan_frame_does_not_poison_later_solves(ctx, F):
    from facedeform_b200 import FdError, make_params
    rig = synth.control_rig(96)
    deform = synth.deformed_rig(rig, F)
    m = ctx.fit(make_params(model=1, radius=2 * rig.spacing, **{"lambda": 0.0}), rig.rest)
    bad = deform.copy()
    bad[0, 5, 1] = np.nan
    with pytest.raises(FdError) as e:
        m.solve(bad)
    assert e.value.status == 4 and m.last_report.terminationtype == -3
    m.solve(deform)                      # the same cached factorisation, clean input
    assert m.last_report.terminationtype == 1
    m.close()


# ---- serialisation ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("kernel", [0, 1])
def test_save_load_round_trip_is_bit_identical(ctx, kernel):
    from facedeform_b200 import make_params
    rig = synth.control_rig(200)
    mesh = synth.face_mesh(5000, topology=False)
    d1, d2 = synth.deformed_rig(rig, 4), synth.deformed_rig(rig, 4, seed=77)
    R = synth.default_radius(["gaussian", "multiquadric"][kernel], rig.spacing)
    m = ctx.fit(make_params(model=1, term=0, kernel=kernel, radius=R, **{"lambda": 0.0}), rig.rest).solve(d1)
    out1, _ = m.eval(mesh.P)
    blob = m.save()
    m2 = ctx.load_model(blob)
    out2, _ = m2.eval(mesh.P)            # the saved weights evaluate without a solve
    np.testing.assert_array_equal(out1, out2)
    a, _ = m.solve(d2).eval(mesh.P)      # and the saved factorisation solves new frames identically
    b, _ = m2.solve(d2).eval(mesh.P)
    np.testing.assert_array_equal(a, b)
    m.close()
    m2.close()
    from facedeform_b200 import FdError
    with pytest.raises(FdError):
        ctx.load_model(blob[:100])


def test_epilogue_parameters_change_without_a_refit(ctx, oracle):
    from facedeform_b200 import FdError, make_params
    rig = synth.control_rig(80)
    deform = synth.deformed_rig(rig, 2)
    mesh = synth.face_mesh(4000, topology=False)
    R = 2 * rig.spacing
    p = make_params(model=1, radius=R, **{"lambda": 0.0})
    m = ctx.fit(p, rig.rest).solve(deform)
    q = make_params(model=1, radius=R, tangent=1, dofalloff=1, falloffrate=2.5, **{"lambda": 0.0})
    m.set_epilogue(q)
    d2 = np.random.default_rng(3).uniform(0, R * R, 4000).astype(np.float32)
    out, fall = m.eval(mesh.P, d2, mesh.tangentu, mesh.tangentv, mesh.N)
    st, rad, W = oracle.fit(_oparams(oracle, q), rig.rest, deform)
    ref, rfall = oracle.evaluate(_oparams(oracle, q), rig.rest, rad, W, mesh.P, d2, mesh.tangentu, mesh.tangentv, mesh.N)
    assert np.abs(out - ref).max() <= 1e-5 * mesh.bbox_diag
    np.testing.assert_allclose(fall, rfall, rtol=2e-6, atol=1e-7)
    with pytest.raises(FdError):
        m.set_epilogue(make_params(model=1, radius=1.5 * R, **{"lambda": 0.0}))  # the kernel radius is a fit parameter
    m.close()


# ---- fd_capture validates the caller's topology ---------------------------------------------------------------
def test_capture_rejects_out_of_range_topology(ctx):
    from facedeform_b200 import FdError
    mesh = synth.face_mesh(400)
    rig = synth.control_rig(8, prims=True)
    bad_vtx = mesh.poly_vtx.copy()
    bad_vtx[7] = 400
    with pytest.raises(FdError) as e:
        ctx.capture(mesh.P, mesh.poly_off, bad_vtx, rig.rest, rig.prim_off, rig.prim_vtx)
    assert e.value.status == 1 and "poly_vtx" in str(e.value)
    neg = mesh.poly_vtx.copy()
    neg[0] = -3
    with pytest.raises(FdError):
        ctx.capture(mesh.P, mesh.poly_off, neg, rig.rest, rig.prim_off, rig.prim_vtx)
    bad_rig = rig.prim_vtx.copy()
    bad_rig[0] = 8
    with pytest.raises(FdError) as e:
        ctx.capture(mesh.P, mesh.poly_off, mesh.poly_vtx, rig.rest, rig.prim_off, bad_rig)
    assert "rig_vtx" in str(e.value)
    off = mesh.poly_off.copy()
    off[3] = off[2] - 1
    with pytest.raises(FdError) as e:
        ctx.capture(mesh.P, off, mesh.poly_vtx, rig.rest, rig.prim_off, rig.prim_vtx)
    assert "poly_off" in str(e.value)
    ok = ctx.capture(mesh.P, mesh.poly_off, mesh.poly_vtx, rig.rest, rig.prim_off, rig.prim_vtx)  # and still works after the errors
    assert ok["ngroups"] == 1


# ---- several contexts in one process -------------------------------------------------------------------------------------
def _two_ctx_results(devs):
    from facedeform_b200 import Context, make_params
    rig = synth.control_rig(600)  # > 512 control points: the cooperative-grid LU; 240 frames: tensor path + slab solve
    deform = synth.deformed_rig(rig, 48)
    mesh = synth.face_mesh(4096, topology=False)
    p = make_params(model=1, radius=2 * rig.spacing, **{"lambda": 0.0})
    outs = []
    ctxs = [Context(d) for d in devs]   # all created first: nothing may depend on a process-wide "already set up" flag
    for c in ctxs:
        m = c.fit(p, rig.rest).solve(deform)
        outs.append(m.eval(mesh.P)[0])
        m.close()
    for c in ctxs:
        c.close()
    return outs


def test_two_contexts_on_one_device_agree():
    a, b = _two_ctx_results([0, 0])
    np.testing.assert_array_equal(a, b)


def test_two_contexts_on_two_devices_agree():
    """function attributes are per device -- the second device's ctx must set its own."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    a, b = _two_ctx_results([0, 1])
    np.testing.assert_array_equal(a, b)


# ---- fd_mgpu_*: one handle, several devices ----------------------------------------------------------------------------
def _mgpu_vs_single(devices, transport, kernel=0, eval_precision=0, F=48):
    from facedeform_b200 import Context, MultiGpu, make_params
    rig = synth.control_rig(300)
    deform = synth.deformed_rig(rig, F)
    mesh = synth.face_mesh(10_001, topology=False)
    R = synth.default_radius(["gaussian", "multiquadric"][kernel], rig.spacing)
    p = make_params(model=1, kernel=kernel, radius=R, eval_precision=eval_precision, dofalloff=1, falloffrate=1.3,
                    **{"lambda": 0.0})
    d2 = np.random.default_rng(5).uniform(0, R * R, mesh.P.shape[0]).astype(np.float32)
    c = Context(devices[0])
    m = c.fit(p, rig.rest).solve(deform)
    ref, rfall = m.eval(mesh.P, d2)
    m.close()
    c.close()
    g = MultiGpu(devices, transport)
    g.fit(p, rig.rest).solve(deform)
    out, fall = g.eval(mesh.P, d2)
    info = g.info()
    # a second solve through the same handle (receivers are reused)
    deform2 = synth.deformed_rig(rig, F, seed=5)
    g.solve(deform2)
    out2, _ = g.eval(mesh.P, d2)
    g.close()
    return ref, rfall, out, fall, info, out2


def test_mgpu_single_device_equals_the_plain_path():
    ref, rfall, out, fall, info, out2 = _mgpu_vs_single([0], 0)
    np.testing.assert_array_equal(out, ref)
    np.testing.assert_array_equal(fall, rfall)
    assert info["ndev"] == 1 and info["bcast_bytes"] == 0
    assert np.abs(out2 - ref).max() > 0  # the second solve produced new positions


@pytest.mark.parametrize("transport", [1, 2])
@pytest.mark.parametrize("case", [(0, 0, 48), (0, 1, 4), (1, 0, 48), (0, 2, 48)])
def test_mgpu_matches_one_gpu_bit_for_bit(transport, case):
    """G-device result == 1-device result, bit for bit (same kernels on disjoint vertex ranges), for the
    NCCL and the peer-load transports, the tensor / FMA / FP64 evaluations."""
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs (run with gpurun --gpus 2)")
    kernel, prec, F = case
    ref, rfall, out, fall, info, _ = _mgpu_vs_single(list(range(n)), transport, kernel, prec, F)
    np.testing.assert_array_equal(out, ref)
    np.testing.assert_array_equal(fall, rfall)
    assert info["transport"] == ("nccl" if transport == 1 else "p2p") and info["bcast_bytes"] > 0 and info["bcast_ms"] >= 0


def test_torchrun_ranks_match_one_gpu_bit_for_bit(tmp_path):
    """one process per GPU (torchrun, NCCL broadcast of the weights through facedeform_b200.shard): the concatenated
    shards equal the single-GPU result bit for bit."""
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs (run with gpurun --gpus 2)")
    out = tmp_path / "ranks.npz"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", "29631", os.path.join(ROOT, "tests", "tools", "shard_worker.py"), "--backend", "nccl", "--out", str(out)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    z = np.load(out)
    np.testing.assert_array_equal(z["sharded"], z["single"])
    assert int(z["world"]) == n


# ---- the operator mirror: group / strict_reference ------------------------------------------------------------------
def test_sop_group_and_strict_reference():
    from test_gpu_sop import Sop
    mesh = synth.face_mesh(2500)
    rig = synth.control_rig(32, prims=True)
    deform = synth.deformed_rig(rig, 1)
    sop = Sop()
    sop.parms.model, sop.parms.radius = 1, 2 * rig.spacing
    st, full, _ = sop.cook(mesh, rig, deform)
    assert st == 0 and sop.L.fd_sop_positions_bumped(sop.h) == 1
    sop.parms.group = b"0-999 ^500-599"
    st, part, _ = sop.cook(mesh, rig, deform)
    assert st == 0, sop.msgs(0)
    assert sop.L.fd_sop_fit_count(sop.h) == 1          # `group` is not a fit parameter: no refit
    inside = np.zeros(2500, bool)
    inside[0:1000] = True
    inside[500:600] = False
    np.testing.assert_array_equal(part[0][inside], full[0][inside])
    np.testing.assert_array_equal(part[0][~inside], mesh.P[~inside])   # outside the group: untouched
    sop.parms.strict_reference = 1                      # the reference never consults the group in its loop (:404-439)
    st, strict, _ = sop.cook(mesh, rig, deform)
    np.testing.assert_array_equal(strict, full)
    sop.parms.group = b"^*"                             # an empty group only withholds the data-ID bump (:485-486)
    st, _, _ = sop.cook(mesh, rig, deform)
    assert sop.L.fd_sop_positions_bumped(sop.h) == 0
    sop.parms.group = b"12-"
    st, _, _ = sop.cook(mesh, rig, deform)
    assert st == 2 and "group" in sop.msgs(0)
    sop.close()


def test_sop_strict_reference_morph_space_quirk():
    """strict_reference = 1: the weights are computed only while !isComputed(); the cook after that skips the pass with
    the reference's warning (SOP_FaceDeform.cpp:446-452).  Default: every cook runs the pass."""
    from test_gpu_sop import Sop
    V, S = 1500, 6
    mesh = synth.face_mesh(V)
    rig = synth.control_rig(24, prims=True)
    deform = synth.deformed_rig(rig, 1)
    rng = np.random.default_rng(4)
    shapes = (mesh.P[None] + 0.05 * rng.standard_normal((S, V, 3))).astype(np.float32)
    sop = Sop()
    sop.parms.model, sop.parms.radius = 1, 2 * rig.spacing
    st, plain, _ = sop.cook(mesh, rig, deform)
    sop.parms.morphspace = 1
    assert sop.L.fd_sop_set_blendshapes(sop.h, shapes.ctypes.data, S, V, 3) == 0
    st, a, _ = sop.cook(mesh, rig, deform)
    st, b, _ = sop.cook(mesh, rig, deform)
    assert st == 0 and np.array_equal(a, b) and not np.array_equal(a, plain)   # default: the pass runs every cook
    sop.parms.strict_reference = 1
    assert sop.L.fd_sop_set_blendshapes(sop.h, shapes.ctypes.data, S, V, 4) == 0  # changed input: init() again
    st, c, _ = sop.cook(mesh, rig, deform)
    assert st == 0 and np.array_equal(c, a)
    st, d, _ = sop.cook(mesh, rig, deform)                                      # isComputed(): skipped with the warning
    assert st == 1 and "Can't compute weights for morphspace deformation" in sop.msgs(1)
    np.testing.assert_array_equal(d, plain)
    sop.close()


@pytest.mark.parametrize("kernel,N", [(0, 600), (1, 600), (0, 2048)])
def test_per_cook_solves_through_the_explicit_inverse(ctx, oracle, kernel, N):
    """one frame per cook against a cached factorisation (the reference's usage, SOP_FaceDeform.cpp:215): from the second
    small solve on the model applies its explicit inverse; the weights match the sweeps and the oracle."""
    from facedeform_b200 import make_params
    rig = synth.control_rig(N)
    R = synth.default_radius(["gaussian", "multiquadric"][kernel], rig.spacing)
    p = make_params(model=1, term=0, kernel=kernel, radius=R, **{"lambda": 0.0})
    m = ctx.fit(p, rig.rest)
    frames = [synth.deformed_rig(rig, 1, seed=40 + i) for i in range(3)]
    m.solve(frames[0])
    W_sweeps, _ = m.weights()                 # first small solve: block sweeps
    m.solve(frames[1])                        # second: builds the inverse and applies it
    m.solve(frames[0])                        # third: the inverse again, same input as the first
    W_inv, _ = m.weights()
    scale = np.abs(W_sweeps).max()
    np.testing.assert_allclose(W_inv, W_sweeps, rtol=0, atol=1e-8 * scale)
    st, rad, W = oracle.fit(_oparams(oracle, p), rig.rest, frames[0])
    np.testing.assert_allclose(W_inv, W, rtol=0, atol=1e-7 * np.abs(W).max())
    two = synth.deformed_rig(rig, 2, seed=50)   # 6 right-hand sides also take the inverse
    m.solve(two)
    W2, _ = m.weights()
    st, rad, Wo = oracle.fit(_oparams(oracle, p), rig.rest, two)
    np.testing.assert_allclose(W2, Wo, rtol=0, atol=1e-7 * np.abs(Wo).max())
    m.close()


@pytest.mark.parametrize("prec,F", [(0, 240), (1, 240), (2, 200), (1, 170)])
def test_host_eval_in_frame_blocks_equals_one_launch(ctx, prec, F):
    """the host-pointer evaluation of a wide batch runs in 80-frame blocks whose read-back overlaps the next block's
    kernel; bit-identical to the single launch of the device-pointer entry (FMA/SFU under AUTO, tensor cores, FP64)."""
    import torch
    from facedeform_b200 import make_params
    N, V = 256, 20_004
    rig = synth.control_rig(N)
    deform = synth.deformed_rig(rig, F)
    mesh = synth.face_mesh(V, topology=False)
    p = make_params(model=1, radius=2 * rig.spacing, eval_precision=prec, dofalloff=1, falloffrate=1.5, **{"lambda": 0.0})
    d2 = np.random.default_rng(8).uniform(0, (2 * rig.spacing) ** 2, V).astype(np.float32)
    m = ctx.fit(p, rig.rest).solve(deform)
    host, hfall = m.eval(mesh.P, d2)                                   # host pointers: 3 blocks of 80 frames
    dev, dfall = m.eval(torch.from_numpy(mesh.P).cuda(), torch.from_numpy(d2).cuda())   # device pointers: one launch
    ctx.synchronize()                                                  # the *_dev entry points only enqueue on the ctx stream
    np.testing.assert_array_equal(host, dev.cpu().numpy())
    np.testing.assert_array_equal(hfall, dfall.cpu().numpy())
    m.close()


# ---- the persistent LU is a function of its input: repeated factorisations are bit-identical ----------------------------
# (round 2 found CTA 0 writing the factored diagonal block back while a CTA that left the grid barrier late still read the
#  raw block: ~30 % of the fits at N = 4096 came out different, a few of them wrong by 1e-1)
@pytest.mark.parametrize("N,kernel,reps", [(4096, 0, 12), (1500, 0, 6), (2048, 1, 6), (300, 0, 6)])
def test_repeated_fits_are_bit_identical(ctx, N, kernel, reps):
    from facedeform_b200 import make_params
    rig = synth.control_rig(N)
    deform = synth.deformed_rig(rig, 2)
    R = synth.default_radius(["gaussian", "multiquadric"][kernel], rig.spacing)
    p = make_params(model=1, term=0, kernel=kernel, radius=R, eval_precision=2, **{"lambda": 0.0})
    first = None
    for _ in range(reps):
        m = ctx.fit(p, rig.rest).solve(deform)
        W = m.weights()[0]
        m.close()
        if first is None:
            first = W
        else:
            np.testing.assert_array_equal(W, first)


def test_fused_lu_agrees_with_the_per_step_launches():
    """the one-launch LU against the per-block-column kernels (FD_LU_UNFUSED, read at context creation) at a size with
    two-level blocking (outer blocks of 256)"""
    from facedeform_b200 import Context, make_params
    rig = synth.control_rig(4096)
    deform = synth.deformed_rig(rig, 2)
    R = synth.default_radius("gaussian", rig.spacing)
    p = make_params(model=1, term=0, kernel=0, radius=R, eval_precision=2, **{"lambda": 0.0})
    Ws = []
    for unfused in (False, True):
        if unfused:
            os.environ["FD_LU_UNFUSED"] = "1"
        try:
            c = Context()
        finally:
            os.environ.pop("FD_LU_UNFUSED", None)
        m = c.fit(p, rig.rest).solve(deform)
        Ws.append(m.weights()[0])
        m.close()
        c.close()
    assert np.abs(Ws[0] - Ws[1]).max() <= 1e-8 * np.abs(Ws[1]).max()


def test_poisoned_allocations_change_nothing():
    """FD_POISON fills every device allocation with 0xFF bytes: a kernel that reads memory the library never wrote would
    turn the weights into NaN (or trip the non-finite flag) instead of depending on what the pool held before"""
    from facedeform_b200 import Context, make_params
    outs = []
    for poison in (False, True):
        if poison:
            os.environ["FD_POISON"] = "1"
        try:
            c = Context()
        finally:
            os.environ.pop("FD_POISON", None)
        for N, kernel, F in ((3500, 0, 40), (700, 2, 3)):
            rig = synth.control_rig(N)
            deform = synth.deformed_rig(rig, F)
            mesh = synth.face_mesh(5_000, topology=False)
            R = synth.default_radius(["gaussian", "multiquadric", "thin_plate"][kernel], rig.spacing)
            m = c.fit(make_params(model=1, term=0, kernel=kernel, radius=R, **{"lambda": 0.0}), rig.rest).solve(deform)
            outs.append(m.eval(mesh.P)[0])
            assert m.report().terminationtype == 1
            m.close()
        c.close()
    np.testing.assert_array_equal(outs[0], outs[2])
    np.testing.assert_array_equal(outs[1], outs[3])


# ---- the exact-digit tensor-core kernel (fd_eval_tcx.cu) ------------------------------------------------------------------
@pytest.mark.parametrize("N,F,V", [(64, 16, 1000), (256, 240, 4096), (1024, 120, 4096), (2048, 40, 3001), (4096, 120, 4096)])
def test_exact_digit_kernel_is_fp64_class(ctx, oracle, N, F, V):
    """Gaussian, FD_EVAL_AUTO, >= 16 frames: tcgen05 with an exact integer leading digit.  Against the oracle the error stays
    below 0.15 x 2^-24 S (the bound FD_EVAL_AUTO admits it with) and below 2e-6 x diagonal up to 4096 control points, where the
    FP32 kernels are at 1.4e-5 ... 2.5e-5; the runtime check of the exact range stays quiet."""
    from facedeform_b200 import make_params
    rig, deform, P, R, ref, diag = _case(oracle, N, F, V)
    m = ctx.fit(make_params(model=1, term=0, kernel=0, radius=R, **{"lambda": 0.0}), rig.rest).solve(deform)
    out, _ = m.eval(P)
    rep = m.report()
    err = float(np.abs(out.astype(np.float64) - ref).max())
    print(f"N={N} F={F}: err/diag {err / diag:.3e} = {err / (2.0 ** -24 * rep.cancellation):.3f} x 2^-24 S")
    assert rep.eval_kernel == 4 and rep.eval_inexact == 0
    assert err <= 0.15 * 2.0 ** -24 * rep.cancellation + 1.5e-7 * diag  # + the FP32 rounding of the output positions
    assert err <= 2e-6 * diag
    m.close()


def test_exact_digit_kernel_epilogue_and_ragged_shapes(ctx, oracle):
    """falloff gate, tangent projection, V not a multiple of 4 (per-lane stores), one 120-column block not full"""
    from facedeform_b200 import make_params
    N, F, V = 100, 21, 3001
    rig = synth.control_rig(N)
    deform = synth.deformed_rig(rig, F)
    mesh = synth.face_mesh(V, topology=False)
    R = synth.default_radius("gaussian", rig.spacing)
    rng = np.random.default_rng(2)
    d2 = rng.uniform(0, 1.5 * R * R, V).astype(np.float32)
    tu = rng.standard_normal((V, 3)).astype(np.float32)
    tv = rng.standard_normal((V, 3)).astype(np.float32)
    nr = rng.standard_normal((V, 3)).astype(np.float32)
    for tangent in (0, 1):
        p = make_params(model=1, term=0, kernel=0, radius=R, tangent=tangent, dofalloff=1, falloffrate=1.7, **{"lambda": 0.0})
        m = ctx.fit(p, rig.rest).solve(deform)
        args = (mesh.P, d2) + ((tu, tv, nr) if tangent else ())
        out, fall = m.eval(*args)
        assert m.report().eval_kernel == 4
        m.close()
        p64 = make_params(model=1, term=0, kernel=0, radius=R, tangent=tangent, dofalloff=1, falloffrate=1.7, eval_precision=2,
                          **{"lambda": 0.0})
        m = ctx.fit(p64, rig.rest).solve(deform)
        ref, rfall = m.eval(*args)
        m.close()
        np.testing.assert_array_equal(fall, rfall)
        skip = d2 > np.float32(R) * np.float32(R)
        np.testing.assert_array_equal(out[:, skip], np.broadcast_to(mesh.P[skip], out[:, skip].shape))  # gated vertices keep P
        assert np.abs(out - ref).max() <= 1e-6 * mesh.bbox_diag


@pytest.mark.parametrize("kw", [dict(model=0), dict(term=1), dict(term=2), dict(lam=1e-3), dict(fidelity=1, layers=2)])
def test_exact_digit_kernel_on_other_systems(ctx, kw):
    """QNN radii (per-centre radii in the bound and in Phi), constant / no polynomial term (fewer affine rows), smoothing, the
    layered fit (stacked centres): FD_EVAL_AUTO's kernel 4 against the FP64 evaluation of the same weights"""
    from facedeform_b200 import make_params
    kw = dict(kw)
    lam = kw.pop("lam", 0.0)
    N, F, V = 700, 24, 6000
    rig = synth.control_rig(N)
    deform = synth.deformed_rig(rig, F)
    mesh = synth.face_mesh(V, topology=False)
    R = synth.default_radius("gaussian", rig.spacing)
    outs = []
    for prec in (0, 2):
        base = dict(model=1, term=0, kernel=0, radius=R, eval_precision=prec)
        base.update(kw)
        m = ctx.fit(make_params(**base, **{"lambda": lam}), rig.rest).solve(deform)
        out, _ = m.eval(mesh.P)
        rep = m.report()
        assert rep.eval_kernel == (4 if prec == 0 else 3) and rep.eval_inexact == 0
        outs.append(out)
        m.close()
    assert np.abs(outs[0].astype(np.float64) - outs[1]).max() <= 1e-6 * mesh.bbox_diag


def test_exact_digit_kernel_far_mesh_and_range_check():
    """a mesh blown up 40x around the rig: the basis sums vanish, the affine rows grow -- their weights are small, so the leading
    digits still sum exactly: no flag, FP64-class result.  The plumbing of the runtime check (fd_report.eval_inexact) is
    exercised with its threshold lowered by a development knob (FD_TC_DEBUG=32, read at context creation)."""
    from facedeform_b200 import Context, make_params
    N, F, V = 256, 40, 4096
    rig = synth.control_rig(N)
    deform = synth.deformed_rig(rig, F)
    mesh = synth.face_mesh(V, topology=False)
    centre = rig.rest.mean(0)
    P_far = ((mesh.P - centre) * 40.0 + centre).astype(np.float32)
    R = synth.default_radius("gaussian", rig.spacing)
    for knob in (False, True):
        if knob:
            os.environ["FD_TC_DEBUG"] = "32"
        try:
            c = Context()
        finally:
            os.environ.pop("FD_TC_DEBUG", None)
        outs = []
        for prec in (0, 2):
            m = c.fit(make_params(model=1, term=0, kernel=0, radius=R, eval_precision=prec, **{"lambda": 0.0}), rig.rest).solve(deform)
            out, _ = m.eval(P_far)
            rep = m.report()
            if prec == 0:
                assert rep.eval_kernel == 4 and (rep.eval_inexact != 0) == knob
            outs.append(out)
            m.close()
        c.close()
        assert np.abs(outs[0].astype(np.float64) - outs[1]).max() <= 1e-6 * np.abs(outs[1]).max()
