"""Generates tests/golden/rbf_golden.npz.

The reference (symek/facedeform) holds no golden vectors and cannot be imported or compiled here (HDK / ALGLIB /
Eigen are absent), so these fixtures come from an INDEPENDENT implementation of the same dense RBF interpolant,
scipy.interpolate.RBFInterpolator (scipy 1.18.1, float64), on seeded inputs.  They pin the oracle (CPU tests) and
the CUDA path (GPU tests) to values that exist outside this repo's own code.  Re-run: python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np
from scipy.interpolate import RBFInterpolator

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from facedeform_b200 import synth  # noqa: E402

KERNELS = {0: "gaussian", 1: "multiquadric", 2: "thin_plate_spline"}
DEGREE = {0: 1, 1: 0, 2: -1}


def main():
    rig = synth.control_rig(24, seed=11)
    deform = synth.deformed_rig(rig, 2, seed=12)
    mesh = synth.face_mesh(64, seed=13, topology=False)
    out = dict(rest=rig.rest, deform=deform, P=mesh.P, spacing=np.float64(rig.spacing))
    y = rig.rest.astype(np.float64)
    delta = (deform - rig.rest[None]).astype(np.float64)   # FP32 subtract like SOP_FaceDeform.cpp:276-278
    for kernel, name in KERNELS.items():
        R = float(np.float32(synth.default_radius(name.replace("_spline", ""), rig.spacing)))
        for term, degree in DEGREE.items():
            if kernel == 2 and term == 2:
                continue
            disp = np.stack([RBFInterpolator(y, delta[f], kernel=name, epsilon=1.0 / R, degree=degree)(
                mesh.P.astype(np.float64)) for f in range(2)])
            out[f"disp_k{kernel}_t{term}"] = disp
            out[f"radius_k{kernel}"] = np.float32(R)
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "rbf_golden.npz"), **out)
    print("wrote rbf_golden.npz with", sorted(out))


if __name__ == "__main__":
    import warnings
    warnings.simplefilter("ignore")
    main()
