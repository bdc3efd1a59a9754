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
ctx.close()
