"""Run-to-run differences of the factorisation: fits the same rig `iters` times, reads the LU factors back through
fd_model_save and reports, for every fit that differs from the most common result, the first block step whose blocks
differ (which phase of the fused LU raced).  This is how the diagonal write-back race of round 2 was located (differences
began at the first block step after an interior update, in L21 / U12 of single row groups).
Usage: python profiles/tools/lu_determinism_probe.py N iters   (FD_NO_NULLSPACE=1 / FD_LU_NOSYM=1 / FD_LU_UNFUSED=1 to bisect)"""
import sys, struct, numpy as np
sys.path.insert(0, ".")
from facedeform_b200 import Context, make_params, synth
from collections import Counter
N = int(sys.argv[1]); iters = int(sys.argv[2])
ctx = Context(0)
rig = synth.control_rig(N)
R = synth.default_radius("gaussian", rig.spacing)
def pad(b): return (b + 15) & ~15
As = []
for it in range(iters):
    p = make_params(model=1, term=0, kernel=0, radius=R, eval_path=1, eval_precision=1, **{"lambda": 0.0})
    m = ctx.fit(p, rig.rest)
    buf = m.save()
    m.close()
    raw = bytes(buf)
    hb = struct.unpack_from("<I", raw, 12)[0]
    # N, np, n, lda, F, ldw follow fd_params inside the header: read them from the tail of the header
    for sh in (0, 4):
        Nn, npoly, n, lda, F, ldw, has_factor, ns = struct.unpack_from("<8i", raw, hb - 8 - 10 * 4 - sh)
        if Nn == N: break
    assert Nn == N, (Nn, hb)
    off = hb + pad(Nn * 3 * 4) + pad(Nn * 8)
    A = np.frombuffer(raw, np.float64, lda * n, off).reshape(n, lda).T[:n, :n].copy()  # A[r, c]
    As.append(A)
print("n", n, "lda", lda, "ns", ns)
keys = [a.tobytes() for a in As]
cnt = Counter(keys)
ref = As[keys.index(cnt.most_common(1)[0][0])]
print("mode count", cnt.most_common(1)[0][1], "of", iters)
for it, a in enumerate(As):
    d = a != ref
    if not d.any():
        continue
    rr, cc = np.nonzero(d)
    step = np.minimum(rr, cc) // 32
    s0 = step.min()
    sel = step == s0
    br, bc = rr[sel] // 32, cc[sel] // 32
    rel = np.abs(a - ref)[d].max() / np.abs(ref).max()
    blocks = sorted(set(zip(br.tolist(), bc.tolist())))
    print(f"it {it}: {d.sum()} entries differ (max {rel:.2e} rel); first step {s0} (k0={32*s0}); blocks at that step: {len(blocks)} {blocks[:12]}"
          f"  rows {rr[sel].min()}..{rr[sel].max()} cols {cc[sel].min()}..{cc[sel].max()}", flush=True)
ctx.close()
