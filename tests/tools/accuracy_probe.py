"""max |P_gpu - P_oracle| / bbox diagonal of the FP32 evaluation paths (tensor-core and FMA/SFU) by control-point count.
Usage: python tests/tools/accuracy_probe.py [N ...]   (the oracle fit at N = 4096 takes ~15 s of CPU)"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from facedeform_b200 import Context, make_params, synth  # noqa: E402
from oracle import fd_oracle as o  # noqa: E402


def main():
    Ns = [int(a) for a in sys.argv[1:]] or [256, 1024, 4096]
    ctx = Context(0)
    for N in Ns:
        F = 120
        rig = synth.control_rig(N)
        deform = synth.deformed_rig(rig, F)
        mesh = synth.face_mesh(50_000, topology=False)
        R = synth.default_radius("gaussian", rig.spacing)
        idx = np.random.default_rng(2).choice(50_000, 1000, replace=False)
        frames = [0, 79, 80, F - 1]
        op = o.make_params(model=1, term=0, kernel=0, radius=R, **{"lambda": 0.0})
        st, rad, W = o.fit(op, rig.rest, deform[frames])
        ref, _ = o.evaluate(op, rig.rest, rad, W, mesh.P[idx], nthreads=o.num_threads())
        row = {}
        for name, path in (("tensor", 2), ("simt", 1)):
            p = make_params(model=1, term=0, kernel=0, radius=R, eval_path=path, **{"lambda": 0.0})
            m = ctx.fit(p, rig.rest).solve(deform)
            out, _ = m.eval(mesh.P)
            row[name] = float(np.abs(out[frames][:, idx].astype(np.float64) - ref).max() / mesh.bbox_diag)
            m.close()
        print(f"N={N:5d} F={F} max err / bbox diag: tensor {row['tensor']:.3e}  simt {row['simt']:.3e}", flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
