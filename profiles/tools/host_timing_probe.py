"""Development probe: host-side and device-side time of each C-ABI call of one C2 step."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from facedeform_b200 import Context, make_params, synth  # noqa: E402

N, V, F = int(os.environ.get("N", 256)), int(os.environ.get("V", 100000)), int(os.environ.get("F", 240))
rig = synth.control_rig(N)
deform = synth.deformed_rig(rig, F)
mesh = synth.face_mesh(V, topology=False)
s = torch.cuda.Stream()
torch.cuda.set_stream(s)
ctx = Context(0, stream=s.cuda_stream)
p = make_params(model=1, radius=2 * rig.spacing, **{"lambda": 0.0})
d_rest, d_def, P = torch.from_numpy(rig.rest).cuda(), torch.from_numpy(deform).cuda(), torch.from_numpy(mesh.P).cuda()
out = torch.empty((F, V, 3), device="cuda")
for it in range(4):
    torch.cuda.synchronize()
    t = [time.perf_counter()]
    m = ctx.fit(p, d_rest); t.append(time.perf_counter())
    torch.cuda.synchronize(); t.append(time.perf_counter())
    m.solve(d_def); t.append(time.perf_counter())
    torch.cuda.synchronize(); t.append(time.perf_counter())
    m.eval(P, out=out); t.append(time.perf_counter())
    torch.cuda.synchronize(); t.append(time.perf_counter())
    m.close(); t.append(time.perf_counter())
    d = [(b - a) * 1e3 for a, b in zip(t, t[1:])]
    print("iter %d ms: fit host %.3f +sync %.3f | solve host %.3f +sync %.3f | eval host %.3f +sync %.3f | close %.3f || dev phases a=%.3f f=%.3f s=%.3f e=%.3f"
          % (it, *d, ctx.phase_ms("assemble"), ctx.phase_ms("factor"), ctx.phase_ms("solve"), ctx.phase_ms("eval")))
