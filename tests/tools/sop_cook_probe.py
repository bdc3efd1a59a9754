"""BASELINE.json configs[0] through the operator mirror: one cook of fd::FaceDeformOp (64 control points, 10k-vertex mesh,
Gaussian, single frame, host pointers in and out) beside the CPU oracle doing what the reference does per cook
(fit + serial evaluation, 1 thread: NO_RBF_THREADS, SOP_FaceDeform.hpp:11).  Three cook kinds: the rest rig changed
(capture + fit + solve + eval), only the deformed rig changed (solve + eval: the factorisation is cached), and the
reference's behaviour of refitting every cook.  Usage: python tests/tools/sop_cook_probe.py"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from facedeform_b200 import _lib, synth  # noqa: E402
from oracle import fd_oracle as o  # noqa: E402

L = _lib.load()
mesh = synth.face_mesh(10_000)
rig = synth.control_rig(64, prims=True)
deform = synth.deformed_rig(rig, 1)
h = L.fd_sop_create(-1)
prm = L.fd_sop_params(h).contents
prm.model, prm.radius, prm.lambda_ = 1, 2 * rig.spacing, 0.0
V = mesh.P.shape[0]
out = np.empty((1, V, 3), np.float32)
fall = np.zeros(V, np.float32)
p = lambda a: None if a is None else a.ctypes.data


def cook(rig_id):
    return L.fd_sop_cook(h, p(mesh.P), V, p(mesh.poly_off), p(mesh.poly_vtx), len(mesh.poly_off) - 1, None, None, None, 1, 1,
                         p(rig.rest), 64, p(rig.prim_off), p(rig.prim_vtx), len(rig.prim_off) - 1, None, rig_id, 1,
                         p(deform), 64, 1, p(out), p(fall))


def best(fn, reps=20):
    t = 1e9
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        t = min(t, time.perf_counter() - t0)
    return t * 1e3


assert cook(1) in (0, 1)
ms_cached = best(lambda: cook(1))
ids = iter(range(2, 10_000))
ms_recapture = best(lambda: cook(next(ids)), reps=10)
op = o.make_params(model=1, radius=2 * rig.spacing, **{"lambda": 0.01})
t0 = time.perf_counter()
st, rad, W = o.fit(op, rig.rest, deform)
ref, _ = o.evaluate(op, rig.rest, rad, W, mesh.P, nthreads=1)
ms_cpu = (time.perf_counter() - t0) * 1e3
print(f"C1 cook (64 control points, 10k vertices, 1 frame, host pointers): solve + eval {ms_cached:.3f} ms; "
      f"capture + fit + solve + eval {ms_recapture:.3f} ms; CPU oracle fit + eval, 1 thread {ms_cpu:.2f} ms; "
      f"max |gpu - cpu| {np.abs(out[0] - ref[0]).max():.2e}")
L.fd_sop_destroy(h)
