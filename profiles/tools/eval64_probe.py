"""eval time of the FP64 tensor-pipe kernel (k_eval64_mma) against the FP32 kernels at the shapes FD_EVAL_AUTO switches on.
Usage: python profiles/tools/eval64_probe.py  -> one JSON line per (shape, precision)"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from facedeform_b200 import Context, make_params, synth  # noqa: E402


def main():
    import torch
    ctx = Context(0)
    for (N, V, F, kern) in ((256, 100_000, 240, "gaussian"), (1024, 100_000, 120, "gaussian"), (2048, 100_000, 240, "gaussian"),
                            (4096, 65_536, 250, "gaussian"), (2048, 100_000, 240, "multiquadric")):
        rig = synth.control_rig(N)
        d_rest = torch.from_numpy(rig.rest).cuda()
        d_def = torch.from_numpy(synth.deformed_rig(rig, F)).cuda()
        d_P = torch.from_numpy(synth.face_mesh(V, topology=False).P).cuda()
        d_out = torch.empty((F, V, 3), dtype=torch.float32, device="cuda")
        for prec in (1, 0, 2):
            if kern != "gaussian" and prec == 1:
                continue
            p = make_params(model=1, term=0, kernel=synth.KERNELS[kern], radius=synth.default_radius(kern, rig.spacing),
                            eval_precision=prec, **{"lambda": 0.0})
            m = ctx.fit(p, d_rest).solve(d_def)
            ts = []
            for i in range(6):
                m.eval(d_P, out=d_out)
                ctx.synchronize()
                if i:
                    ts.append(ctx.phase_ms("eval"))
            rep = m.report()
            ms = float(np.median(ts))
            flops = float(V) * N * (6.0 * F + 10.0)
            print(json.dumps(dict(N=N, V=V, F=F, kernel=kern, eval_precision=["AUTO", "FP32", "FP64"][prec], eval_kernel=rep.eval_kernel,
                                  cancellation=rep.cancellation, eval_ms=round(ms, 4), alg_tflops=round(flops / ms / 1e9, 2),
                                  solve_ms=round(ctx.phase_ms("solve"), 4))), flush=True)
            m.close()
    ctx.close()


if __name__ == "__main__":
    main()
