"""Small end-to-end case for compute-sanitizer memcheck: every kernel family once (SIMT + tensor eval, cluster LU with
one and several CTAs, slab and blocked solve, capture)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

from facedeform_b200 import Context, make_params, synth  # noqa: E402

ctx = Context(0)
for N, F, V, kernel, path in [(70, 3, 700, 0, 1), (70, 30, 1000, 0, 2), (300, 2, 500, 1, 0), (600, 1, 300, 2, 0)]:
    rig = synth.control_rig(N, prims=True)
    deform = synth.deformed_rig(rig, F)
    mesh = synth.face_mesh(V)
    p = make_params(model=1, kernel=kernel, radius=2 * rig.spacing, eval_path=path, **{"lambda": 0.0})
    m = ctx.fit(p, rig.rest).solve(deform)
    out, fall = m.eval(mesh.P)
    assert np.isfinite(out).all()
    cap = ctx.capture(mesh.P, mesh.poly_off, mesh.poly_vtx, rig.rest, rig.prim_off, rig.prim_vtx, None, 3, 0.3, 1)
    m.close()
    print("ok", N, F, V, kernel, path, flush=True)
# FP32 LU + refinement, the fused no-pivot LU beyond one cluster (cooperative grid), DirectBSEdit
from facedeform_b200 import DirectBSEdit  # noqa: E402
for N, fp in [(200, 1), (600, 0)]:
    rig = synth.control_rig(N)
    deform = synth.deformed_rig(rig, 2)
    p = make_params(model=1, kernel=0, radius=1.5 * rig.spacing, factor_precision=fp, **{"lambda": 0.0})
    m = ctx.fit(p, rig.rest).solve(deform)
    assert m.report().terminationtype == 1
    out, _ = m.eval(rig.rest)
    assert np.abs(out - deform).max() < 1e-4
    m.close()
    print("ok refine/fused", N, fp, flush=True)
rng = np.random.default_rng(0)
rest = rng.standard_normal((3000, 3)).astype(np.float32)
shapes = (rest[None] + 0.1 * rng.standard_normal((6, 3000, 3))).astype(np.float32)
b = DirectBSEdit(ctx, rest, shapes)
w = b.compute_weights(rest + 0.01, rest)
out = b.displace(rest + 0.01, rest, weightrange=(0, 1), dofalloff=1, falloffradius=0.5)
assert np.isfinite(out).all() and np.isfinite(w).all()
b.close()
print("ok dbse", flush=True)
ctx.close()
