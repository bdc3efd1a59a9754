"""Fused persistent no-pivot LU vs the per-block-column launches: weight agreement and factor-phase time by N.
Usage: python profiles/tools/lu_fused_probe.py [N ...]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from facedeform_b200 import Context, make_params, synth  # noqa: E402


def fit_once(ctx, p, rig, deform, reps):
    best = 1e9
    W = None
    for _ in range(reps):
        m = ctx.fit(p, rig.rest)
        ctx.synchronize()
        best = min(best, ctx.phase_ms("factor"))
        if W is None:
            m.solve(deform)
            W, _ = m.weights()
            rep = m.report()
        m.close()
    return best, W, rep


def main():
    Ns = [int(a) for a in sys.argv[1:]] or [1, 5, 31, 32, 33, 64, 100, 256, 300, 511, 513, 700, 1024, 1600, 2048, 4096]
    ctx = Context(0)
    for N in Ns:
        rig = synth.control_rig(N)
        deform = synth.deformed_rig(rig, 2)
        p = make_params(model=1, term=0, kernel=0, radius=2.0 * rig.spacing, **{"lambda": 0.0})
        os.environ.pop("FD_LU_UNFUSED", None)
        t0 = time.time()
        try:
            tf, Wf, rf = fit_once(ctx, p, rig, deform, 4)
            os.environ["FD_LU_UNFUSED"] = "1"
            tu, Wu, ru = fit_once(ctx, p, rig, deform, 4)
        except Exception as e:  # noqa: BLE001
            print(f"N={N:5d} failed ({'unfused' if 'FD_LU_UNFUSED' in os.environ else 'fused'}): {e}", flush=True)
            continue
        finally:
            os.environ.pop("FD_LU_UNFUSED", None)
        d = float(np.abs(Wf - Wu).max() / max(np.abs(Wu).max(), 1e-300))
        print(f"N={N:5d} fused {tf:8.3f} ms  unfused {tu:8.3f} ms  x{tu / tf:5.2f}  max|dW|/max|W| {d:.2e}  "
              f"term {rf.terminationtype}/{ru.terminationtype} pivots [{rf.min_pivot:.3e},{rf.max_pivot:.3e}] vs "
              f"[{ru.min_pivot:.3e},{ru.max_pivot:.3e}]  ({time.time() - t0:.1f}s)", flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
