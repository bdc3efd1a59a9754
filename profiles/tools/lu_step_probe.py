import os, sys
sys.path.insert(0, '/root/repo')
from facedeform_b200 import Context, make_params, synth
ctx = Context(0)
rig = synth.control_rig(256)
p = make_params(model=1, term=0, kernel=0, radius=2.0 * rig.spacing, **{"lambda": 0.0})
for i in range(3):
    m = ctx.fit(p, rig.rest); ctx.synchronize(); print("factor ms", ctx.phase_ms("factor"), flush=True); m.close()
