"""Generates tests/golden/dbse_golden.npz: DirectBSEdit fixtures from an INDEPENDENT implementation -- LAPACK's
Householder QR through scipy.linalg.qr(mode="raw") (same packed storage and sign conventions as Eigen's
HouseholderQR::matrixQR(), which the reference uses at dbse.cpp:31 and :53 but which is absent here) and plain numpy
for the weight / displacement formulas of dbse.cpp:37-75 and SOP_FaceDeform.cpp:460-472.
Re-run: python tests/golden/make_golden_dbse.py"""
import os

import numpy as np
import scipy.linalg


def main():
    rng = np.random.default_rng(21)
    P, S = 96, 5
    rest = rng.standard_normal((P, 3)).astype(np.float32)
    shapes = (rest[None] + 0.1 * rng.standard_normal((S, P, 3))).astype(np.float32)
    pos = (rest + 0.05 * rng.standard_normal((P, 3))).astype(np.float32)
    M = (shapes - rest[None]).reshape(S, -1).T.astype(np.float64)          # FP32 subtract, widened (dbse.cpp:24-27)
    (qr, tau), _ = scipy.linalg.qr(np.asfortranarray(M), mode="raw")
    delta = (pos - rest).astype(np.float64).reshape(-1)                    # :46-48
    w = delta @ qr                                                         # :53-54
    lo, hi, fr = np.float32(-0.25), np.float32(0.5), np.float32(0.5)
    disp = np.zeros((P, 3), np.float32)
    for s in range(S):                                                     # :63-73 (FP32, column order)
        cw = np.float32(min(max(np.float32(w[s] * 3), lo), hi))
        disp = (disp + (M[:, s].astype(np.float32).reshape(P, 3) * cw).astype(np.float32)).astype(np.float32)
    disp = (disp + ((pos - rest) * fr).astype(np.float32)).astype(np.float32)   # SOP_FaceDeform.cpp:467-470
    out = (rest + disp).astype(np.float32)                                      # :471
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "dbse_golden.npz"), rest=rest,
                        shapes=shapes, pos=pos, qr=qr, tau=tau, weights=w, weightrange=np.array([lo, hi], np.float32),
                        falloffradius=fr, P_out=out)
    print("wrote dbse_golden.npz")


if __name__ == "__main__":
    main()
