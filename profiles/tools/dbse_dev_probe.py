"""DirectBSEdit device times (CUDA events around the kernels, no host copies): QR of the 3P x S shapes matrix, the weight
pass and the displacement pass.  Usage: python profiles/tools/dbse_dev_probe.py [P S]  -> one JSON line"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from facedeform_b200 import Context, DirectBSEdit  # noqa: E402


def run(ctx, P, S):
    rng = np.random.default_rng(0)
    rest = rng.standard_normal((P, 3)).astype(np.float32)
    shapes = (rest[None] + 0.1 * rng.standard_normal((S, P, 3), dtype=np.float32)).astype(np.float32)
    pos = (rest + 0.05 * rng.standard_normal((P, 3), dtype=np.float32)).astype(np.float32)
    qr = []
    for _ in range(2):
        b = DirectBSEdit(ctx, rest, shapes)
        qr.append(ctx.phase_ms("factor"))
        b.compute_weights(pos, rest)
        w_ms = ctx.phase_ms("solve")
        b.displace(pos, rest, weightrange=(0.0, 1.0))
        d_ms = ctx.phase_ms("eval")
        b.close()
    return dict(P=P, S=S, qr_ms=round(min(qr), 3), weights_ms=round(w_ms, 4), displace_ms=round(d_ms, 4),
                qr_bytes_min=3 * P * S * 8 * 2, note="device time of the kernels; qr_bytes_min = one read + one write of the FP64 matrix")


if __name__ == "__main__":
    P, S = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (1_000_000, 32)
    c = Context(0)
    print(json.dumps(run(c, P, S)), flush=True)
    c.close()
