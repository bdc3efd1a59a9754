"""Phase times of BASELINE.json's other configurations on one B200, device resident (C2 is bench.py's line).
    C1   64 control points, 10k vertices, Gaussian, 1 frame
    C3   2048 control points, 1M vertices, multiquadric / thin plate + affine block, 1 frame (FP64 evaluation)
    C3g  the same shape with the Gaussian kernel (FP32 FMA/SFU evaluation)
    C5s  one GPU's slice of C5 at a size that fits a probe: 4096 control points, 1000 frames, V vertices
Usage: python profiles/tools/configs_probe.py [--v5 131072] [--reps 5]  -> one JSON line per configuration"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from facedeform_b200 import Context, make_params, synth  # noqa: E402


def run(ctx, name, N, V, F, kernel, reps, **extra):
    import torch
    rig = synth.control_rig(N)
    deform = synth.deformed_rig(rig, F)
    mesh = synth.face_mesh(V, topology=False)
    p = make_params(model=1, term=0, kernel=synth.KERNELS[kernel], radius=synth.default_radius(kernel, rig.spacing),
                    **{"lambda": 0.0}, **extra)
    d_rest = torch.from_numpy(rig.rest).cuda()
    d_def = torch.from_numpy(deform).cuda()
    d_P = torch.from_numpy(mesh.P).cuda()
    d_out = torch.empty((F, V, 3), dtype=torch.float32, device="cuda")
    ph = {k: [] for k in ("assemble", "factor", "solve", "eval")}
    for i in range(reps + 1):
        m = ctx.fit(p, d_rest)
        m.solve(d_def)
        m.eval(d_P, out=d_out)
        ctx.synchronize()
        if i:
            for k in ph:
                ph[k].append(ctx.phase_ms(k))
        m.close()
    med = {k: float(np.median(v)) for k, v in ph.items()}
    pairs = float(V) * N
    row = dict(config=name, N=N, V=V, F=F, kernel=kernel, **{k + "_ms": round(v, 4) for k, v in med.items()},
               step_ms=round(sum(med.values()), 4), eval_vertex_frames_per_s=V * F / (med["eval"] * 1e-3),
               eval_pairs_per_s=pairs / (med["eval"] * 1e-3),
               eval_alg_tflops=pairs * (6.0 * F + 10.0) / (med["eval"] * 1e-3) / 1e12)
    print(json.dumps(row), flush=True)
    del d_out
    torch.cuda.empty_cache()
    return row


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--v5", type=int, default=131072)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--only", default="")
    a = ap.parse_args()
    ctx = Context(0)
    todo = [("C1", 64, 10_000, 1, "gaussian"), ("C3", 2048, 1_000_000, 1, "multiquadric"),
            ("C3t", 2048, 1_000_000, 1, "thin_plate"), ("C3g", 2048, 1_000_000, 1, "gaussian"),
            ("C5s", 4096, a.v5, 1000, "gaussian")]
    for name, N, V, F, kern in todo:
        if a.only and name not in a.only.split(","):
            continue
        run(ctx, name, N, V, F, kern, a.reps)
    ctx.close()


if __name__ == "__main__":
    main()
