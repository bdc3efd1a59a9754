import sys; sys.path.insert(0, '.')
import torch
from facedeform_b200 import Context, make_params, synth
ctx = Context(0)
rig = synth.control_rig(2048); deform = synth.deformed_rig(rig, 1); mesh = synth.face_mesh(1_000_000, topology=False)
p = make_params(model=1, term=0, kernel=2, radius=synth.default_radius("thin_plate", rig.spacing), **{"lambda": 0.0})
m = ctx.fit(p, torch.from_numpy(rig.rest).cuda()); m.solve(torch.from_numpy(deform).cuda())
P = torch.from_numpy(mesh.P).cuda(); out = torch.empty((1, 1_000_000, 3), device="cuda")
for i in range(4): m.eval(P, out=out)
ctx.synchronize()
