"""BASELINE.json config 4: FP64 LU vs FP32 LU + FP64 iterative refinement, Gaussian kernel, radius walked across the
conditioning range (cond grows steeply with radius / spacing).  One JSON line per (N, radius) point:
factor ms of both modes, refinement sweeps, final relative residual, weight agreement with the FP64 solve.
Usage: python profiles/tools/ir_study.py [--n 2048,8192]"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from facedeform_b200 import Context, FdError, make_params, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", default="2048,8192")
    ap.add_argument("--scales", default="1.0,1.5,2.0,2.5,3.0,4.0")
    ap.add_argument("--kernel", default="gaussian")
    a = ap.parse_args()
    ctx = Context(0)
    for N in [int(x) for x in a.n.split(",")]:
        rig = synth.control_rig(N)
        deform = synth.deformed_rig(rig, 1)
        for sc in [float(x) for x in a.scales.split(",")]:
            row = dict(N=N, kernel=a.kernel, radius_over_spacing=sc)
            W = {}
            for name, fp in (("fp64", 0), ("fp32_ir", 1)):
                p = make_params(model=1, term=0, kernel=synth.KERNELS[a.kernel], radius=sc * rig.spacing,
                                factor_precision=fp, **{"lambda": 0.0})
                ts = []
                for rep_i in range(3):
                    m = ctx.fit(p, rig.rest)
                    ctx.synchronize()
                    ts.append(ctx.phase_ms("factor"))
                    if rep_i < 2:
                        m.close()
                row[name + "_factor_ms"] = round(float(np.median(ts)), 3)
                rep = None
                try:
                    m.solve(deform)
                    rep = m.report()
                    row[name + "_solve_ms"] = round(ctx.phase_ms("solve"), 3)
                    W[name] = m.weights()[0]
                except FdError:
                    rep = m.last_report
                if rep is not None:
                    row[name + "_term"] = rep.terminationtype
                    row[name + "_pivot_ratio"] = float(rep.max_pivot / rep.min_pivot) if rep.min_pivot > 0 else None
                    if fp:
                        row["sweeps"], row["residual"] = rep.iterationscount, float(rep.residual)
                m.close()
            if len(W) == 2:
                row["max_weight_diff_rel"] = float(np.abs(W["fp64"] - W["fp32_ir"]).max() / np.abs(W["fp64"]).max())
            print(json.dumps(row), flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
