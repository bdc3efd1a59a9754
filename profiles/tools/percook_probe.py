"""per-cook latency against a cached factorisation (one frame per cook, the reference's usage, SOP_FaceDeform.cpp:215):
solve ms of the first cook (block sweeps), the second (builds the explicit inverse) and the following ones, eval ms.
Usage: python profiles/tools/percook_probe.py  -> one JSON line per configuration"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from facedeform_b200 import Context, make_params, synth  # noqa: E402


def main():
    import torch
    ctx = Context(0)
    for (N, V, kern) in ((2048, 1_000_000, "gaussian"), (2048, 1_000_000, "multiquadric"), (2048, 1_000_000, "thin_plate"),
                         (4096, 1_000_000, "gaussian"), (8192, 1_000_000, "gaussian")):
        rig = synth.control_rig(N)
        d_rest = torch.from_numpy(rig.rest).cuda()
        d_P = torch.from_numpy(synth.face_mesh(V, topology=False).P).cuda()
        d_out = torch.empty((1, V, 3), dtype=torch.float32, device="cuda")
        p = make_params(model=1, term=0, kernel=synth.KERNELS[kern], radius=synth.default_radius(kern, rig.spacing), **{"lambda": 0.0})
        m = ctx.fit(p, d_rest)
        solve, ev = [], []
        for i in range(6):
            d_def = torch.from_numpy(synth.deformed_rig(rig, 1, seed=10 + i)).cuda()
            m.solve(d_def)
            m.eval(d_P, out=d_out)
            ctx.synchronize()
            solve.append(round(ctx.phase_ms("solve"), 4))
            ev.append(round(ctx.phase_ms("eval"), 4))
        rep = m.report()
        print(json.dumps(dict(N=N, V=V, kernel=kern, fit_ms=round(ctx.phase_ms("assemble") + ctx.phase_ms("factor"), 3),
                              solve_ms_by_cook=solve, eval_ms=float(np.median(ev)), eval_kernel=rep.eval_kernel,
                              cook_ms_steady=round(float(np.median(solve[2:])) + float(np.median(ev)), 4))), flush=True)
        m.close()
    ctx.close()


if __name__ == "__main__":
    main()
