"""max |P_gpu - P_oracle| / bbox diagonal of the evaluation paths (FP32 tensor-core and FMA/SFU forced, FD_EVAL_AUTO, FP64) by
control-point count, next to the cancellation estimate S (fd_report.cancellation) and the error model 2^-24 S that
FD_EVAL_AUTO decides with -- the calibration of ERR_COEF_* in csrc/fd_eval64.cu.
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
        idx = np.sort(np.random.default_rng(2).choice(50_000, 4096, replace=False))
        frames = list(range(F))  # every frame: the maximum over ~1.5 M values (a sparse sample misses the tail)
        op = o.make_params(model=1, term=0, kernel=0, radius=R, **{"lambda": 0.0})
        st, rad, W = o.fit(op, rig.rest, deform[frames])
        ref, _ = o.evaluate(op, rig.rest, rad, W, mesh.P[idx], nthreads=o.num_threads())
        row = {}
        S = 0.0
        for name, path, prec in (("tensor", 2, 1), ("simt", 1, 1), ("auto", 0, 0), ("fp64", 0, 2)):
            p = make_params(model=1, term=0, kernel=0, radius=R, eval_path=path, eval_precision=prec, **{"lambda": 0.0})
            m = ctx.fit(p, rig.rest).solve(deform)
            out, _ = m.eval(mesh.P)
            rep = m.report()
            S = max(S, rep.cancellation)
            row[name] = float(np.abs(out[frames][:, idx].astype(np.float64) - ref).max() / mesh.bbox_diag)
            row[name + "_kernel"] = rep.eval_kernel
            m.close()
        unit = 2.0 ** -24 * S / mesh.bbox_diag
        print(f"N={N:5d} F={F} max err / bbox diag: tensor {row['tensor']:.3e} ({row['tensor'] / unit:.2f} x 2^-24 S)  "
              f"simt {row['simt']:.3e} ({row['simt'] / unit:.2f} x)  auto {row['auto']:.3e} [kernel {row['auto_kernel']}]  "
              f"fp64 {row['fp64']:.3e}   S = {S:.3e}, 2^-24 S / diag = {unit:.3e}", flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
