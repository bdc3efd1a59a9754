"""Development probe: run the C2 evaluation with FD_TC_DEBUG=1 to print per-unit phase cycle counts of CTA 0."""
import os
import sys

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
p = make_params(model=1, radius=2 * rig.spacing, eval_path=int(os.environ.get("PATH_", 0)), **{"lambda": 0.0})
m = ctx.fit(p, torch.from_numpy(rig.rest).cuda())
m.solve(torch.from_numpy(deform).cuda())
P = torch.from_numpy(mesh.P).cuda()
out = torch.empty((F, V, 3), device="cuda")
for i in range(3):
    m.eval(P, out=out)
    print("eval ms", ctx.phase_ms("eval"), file=sys.stderr)
