import os, sys, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from facedeform_b200 import Context, make_params, synth
import torch
ctx = Context(0)
N, V = 2048, 1_000_000
rig = synth.control_rig(N)
d_rest = torch.from_numpy(rig.rest).cuda()
d_P = torch.from_numpy(synth.face_mesh(V, topology=False).P).cuda()
d_out = torch.empty((1, V, 3), dtype=torch.float32, device="cuda")
p = make_params(model=1, term=0, kernel=0, radius=2 * rig.spacing, **{"lambda": 0.0})
for mode in ("persistent model", "fresh model per cook"):
    ev = []
    m = ctx.fit(p, d_rest)
    for i in range(6):
        if mode != "persistent model" and i:
            m.close(); m = ctx.fit(p, d_rest)
        m.solve(torch.from_numpy(synth.deformed_rig(rig, 1, seed=10 + i)).cuda())
        m.eval(d_P, out=d_out); ctx.synchronize()
        ev.append(round(ctx.phase_ms("eval"), 3))
        rep = m.report()
        ev.append((rep.eval_kernel, round(rep.cancellation, 1)))
    m.close()
    print(mode, ev, flush=True)
