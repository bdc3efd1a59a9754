"""Shapes and forms of the exact-digit tensor-core evaluation (csrc/fd_eval_tcx.cu) beyond tests/test_gpu_round2.py: the
development forms of the kernel against the shipped one, and batches of more than 240 frames in one launch (the vertex loop of
SOP_FaceDeform.cpp:404-439 for all frames at once).  Through the C ABI (ctypes), against the FP64 evaluation of the same
weights."""
import os

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


def test_exact_digit_kernel_variants_agree():
    """The shipped kernel takes two 120-column blocks per unit with one N = 240 MMA for both; the development knobs
    FD_TCX_NARROW=1 (one N = 128 MMA per block) and FD_TCX_CBU=1 (one block per unit, two units in flight in tensor memory)
    select the earlier forms (read at context creation).  All three see the same digits and sum them in the same order along
    K: the results agree far inside the FP64-class bound -- on a batch with an odd number of blocks and a partly filled last
    one (100 frames = 2.5 blocks), with and without the vectorised store path (V % 4)."""
    from facedeform_b200 import Context, make_params
    N, F = 300, 100
    rig = synth.control_rig(N)
    deform = synth.deformed_rig(rig, F)
    R = synth.default_radius("gaussian", rig.spacing)
    for V in (12800, 4097):
        mesh = synth.face_mesh(V, topology=False)
        outs = {}
        for name, env in (("default", {}), ("narrow", {"FD_TCX_NARROW": "1"}), ("one_block", {"FD_TCX_CBU": "1"})):
            os.environ.update(env)
            try:
                c = Context()
            finally:
                for k in env:
                    os.environ.pop(k, None)
            m = c.fit(make_params(model=1, term=0, kernel=0, radius=R, **{"lambda": 0.0}), rig.rest).solve(deform)
            out, _ = m.eval(mesh.P)
            rep = m.report()
            assert rep.eval_kernel == 4 and rep.eval_inexact == 0
            outs[name] = out.astype(np.float64)
            m.close()
            c.close()
        for name in ("narrow", "one_block"):
            assert np.abs(outs[name] - outs["default"]).max() <= 1e-6 * mesh.bbox_diag, (name, V)


@pytest.mark.parametrize("N,F,V", [(64, 250, 4096), (300, 270, 3001)])
def test_exact_digit_kernel_more_than_six_column_blocks(ctx, N, F, V):
    """More than 240 frames in ONE launch (the device-pointer entry; C5 runs 250-frame chunks): seven 120-column blocks -- an
    odd count, so the last unit of a vertex tile holds a single, partly filled block -- and the column scales no longer stay
    resident in shared memory but are reloaded per unit.  Checked against the FP64 evaluation of the same weights."""
    import torch
    from facedeform_b200 import make_params
    rig = synth.control_rig(N)
    deform = synth.deformed_rig(rig, F)
    mesh = synth.face_mesh(V, topology=False)
    R = synth.default_radius("gaussian", rig.spacing)
    m = ctx.fit(make_params(model=1, term=0, kernel=0, radius=R, **{"lambda": 0.0}), rig.rest).solve(deform)
    dev, _ = m.eval(torch.from_numpy(mesh.P).cuda())   # device pointers: one launch over all the frames
    ctx.synchronize()                                  # the *_dev entry points only enqueue on the ctx stream
    rep = m.report()
    assert rep.eval_kernel == 4 and rep.eval_inexact == 0
    out = dev.cpu().numpy().astype(np.float64)
    m.close()
    m = ctx.fit(make_params(model=1, term=0, kernel=0, radius=R, eval_precision=2, **{"lambda": 0.0}), rig.rest).solve(deform)
    ref, _ = m.eval(mesh.P)
    m.close()
    assert np.isfinite(out).all()
    assert np.abs(out - ref).max() <= 1e-6 * mesh.bbox_diag
