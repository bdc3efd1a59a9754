"""FD_EVAL_AUTO (the exact-digit tensor-core kernel for batches of >= 16 frames) against the FP64 evaluation of the same weights on
the full mesh, over shapes that exercise the kernel's edges: one partly filled 120-column block, vertex counts that are not a
multiple of 128 / 4, K = N + 4 with a short last stage, N up to 4096.  Prints max |difference| / bbox diagonal.
Usage: python profiles/tools/tcx_check.py [number of cases]"""
import sys, time, numpy as np
sys.path.insert(0, ".")
from facedeform_b200 import Context, make_params, synth
ctx = Context(0)
cases = [(256, 16, 1000), (256, 40, 4096), (64, 240, 10000), (256, 240, 100000), (1024, 120, 20001), (300, 100, 12800), (4096, 120, 20000)]
if len(sys.argv) > 1:
    cases = cases[:int(sys.argv[1])]
for N, F, V in cases:
    rig = synth.control_rig(N); deform = synth.deformed_rig(rig, F)
    mesh = synth.face_mesh(V, topology=False)
    R = synth.default_radius("gaussian", rig.spacing)
    outs = {}
    for prec in (0, 2):
        p = make_params(model=1, term=0, kernel=0, radius=R, eval_precision=prec, **{"lambda": 0.0})
        m = ctx.fit(p, rig.rest).solve(deform)
        out, _ = m.eval(mesh.P)
        t0 = time.perf_counter()
        out, _ = m.eval(mesh.P)
        dt = time.perf_counter() - t0
        rep = m.report()
        outs[prec] = out
        m.close()
        print(f"  N={N} F={F} V={V} prec={prec}: kernel {rep.eval_kernel} inexact {rep.eval_inexact} host eval {dt*1e3:.2f} ms  finite {np.isfinite(out).all()}", flush=True)
    d = np.abs(outs[0].astype(np.float64) - outs[2]) / mesh.bbox_diag
    bad = np.argwhere(d > 1e-6)
    print(f"N={N} F={F} V={V}: max |tcx - fp64| / diag = {d.max():.3e}   (> 1e-6: {len(bad)}; first {bad[:3].tolist()})", flush=True)
ctx.close()
