"""DirectBSEdit timings through the host-pointer C ABI (H2D / D2H included), beside the CPU oracle (1 thread, like Eigen's
HouseholderQR in the reference).  Usage: python tests/tools/dbse_probe.py [P S]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from facedeform_b200 import Context, DirectBSEdit  # noqa: E402
from oracle import fd_oracle as o  # noqa: E402

P, S = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (100_000, 48)
rng = np.random.default_rng(0)
rest = rng.standard_normal((P, 3)).astype(np.float32)
shapes = (rest[None] + 0.1 * rng.standard_normal((S, P, 3))).astype(np.float32)
pos = (rest + 0.05 * rng.standard_normal((P, 3))).astype(np.float32)
ctx = Context(0)
DirectBSEdit(ctx, rest[:100], shapes[:, :100]).close()  # warm-up


def t(fn, reps=3):
    best = 1e9
    for _ in range(reps):
        t0 = time.perf_counter()
        r = fn()
        best = min(best, time.perf_counter() - t0)
    return best * 1e3, r


ms_init, b = t(lambda: DirectBSEdit(ctx, rest, shapes))
ms_w, w = t(lambda: b.compute_weights(pos, rest))
ms_d, out = t(lambda: b.displace(pos, rest, weightrange=(0.0, 1.0)))
t0 = time.perf_counter()
M = o.dbse_shapes_matrix(rest, shapes)
QR, _ = o.householder_qr(M)
c_init = (time.perf_counter() - t0) * 1e3
t0 = time.perf_counter()
w_o = o.dbse_weights(QR, pos, rest)
c_w = (time.perf_counter() - t0) * 1e3
t0 = time.perf_counter()
ref = o.dbse_displace(M, w, pos, rest, weightrange=(0.0, 1.0))
c_d = (time.perf_counter() - t0) * 1e3
print(f"P={P} S={S}: init (QR) gpu {ms_init:.2f} ms / cpu {c_init:.1f} ms; weights {ms_w:.3f} / {c_w:.1f} ms; "
      f"displace {ms_d:.3f} / {c_d:.1f} ms; max|w - w_cpu| {np.abs(w - w_o).max():.2e}; displace bit-exact {np.array_equal(out, ref)}")
