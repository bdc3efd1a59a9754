"""cycle breakdown of one block step of the fused LU (FD_LU_DEBUG=<step> prints it from the kernel) and factor ms.
Usage: FD_LU_DEBUG=3 python profiles/tools/lu_step_probe.py [N ...]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from facedeform_b200 import Context, make_params, synth  # noqa: E402

ctx = Context(0)
for N in [int(a) for a in sys.argv[1:]] or [256]:
    rig = synth.control_rig(N)
    p = make_params(model=1, term=0, kernel=0, radius=2.0 * rig.spacing, **{"lambda": 0.0})
    for i in range(3):
        m = ctx.fit(p, rig.rest)
        ctx.synchronize()
        print("N", N, "factor ms", round(ctx.phase_ms("factor"), 4), flush=True)
        m.close()
ctx.close()
