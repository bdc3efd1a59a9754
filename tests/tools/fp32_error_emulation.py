"""CPU emulation of the FP32 evaluation paths' arithmetic against the FP64 sum, in units of 2^-24 S (S = the cancellation
estimate of fd_report.cancellation): the sequential FP32 FMA chain (FMA/SFU kernel), the same with a correctly rounded
phi (how much of the error is phi's own FP32 error), and FP16-split tensor-core schemes with 2-way (3 or 4 MMAs) and 3-way
(6 MMAs) splits, FP32 accumulation in K = 16 blocks rounded to nearest or toward zero.  It is the experiment behind the
error model of FD_EVAL_AUTO; the GPU's measured coefficients are in tests/test_gpu_round2.py.
Usage: python tests/tools/fp32_error_emulation.py N      (uses the oracle for the weights; a minute at N = 1024)"""
import sys, numpy as np
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from facedeform_b200 import synth
from oracle import fd_oracle as o
N, F, V = int(sys.argv[1]), 120, 4096
rig = synth.control_rig(N); deform = synth.deformed_rig(rig, F)
mesh = synth.face_mesh(100_000, topology=False)
idx = np.sort(np.random.default_rng(11).choice(mesh.P.shape[0], V, replace=False)); P = np.ascontiguousarray(mesh.P[idx])
R = 2.0 * rig.spacing
p = o.make_params(model=1, term=0, kernel=0, radius=R, **{"lambda": 0.0})
st, rad, W = o.fit(p, rig.rest, deform)          # W: (N+4, 3F) float64
C = rig.rest.astype(np.float64)
# exact (float64) reference incl. polynomial part
d2 = ((P[:, None, :].astype(np.float64) - C[None]) ** 2).sum(-1)
Phi64 = np.exp(-d2 / (R * R))
A64 = np.concatenate([Phi64, np.ones((V, 1)), P.astype(np.float64)], 1)   # [Phi | 1 x y z]
ref = A64 @ W
S = (np.abs(W[:N]).max(1)[None, :] * np.exp(-((C[:, None] - C[None]) ** 2).sum(-1) / (R * R))).sum(1).max()
unit = 2.0 ** -24 * S
# FP32 phi: float32 distance, float32 argument, exp2 with ~1.5 ulp noise (MUFU.EX2)
Pf, Cf = P.astype(np.float32), rig.rest.astype(np.float32)
dx = Pf[:, None, :] - Cf[None]
r2 = (dx[..., 0] * dx[..., 0] + dx[..., 1] * dx[..., 1]).astype(np.float32) + dx[..., 2] * dx[..., 2]
prm = np.float32(-1.4426950408889634 / (R * R))
t = (r2 * prm).astype(np.float32)
rng = np.random.default_rng(0)
phi32 = np.exp2(t.astype(np.float64)) * (1.0 + rng.uniform(-1.5, 1.5, t.shape) * 2.0 ** -24)
phi32 = phi32.astype(np.float32)
A32 = np.concatenate([phi32, np.ones((V, 1), np.float32), Pf], 1)          # float32 operand incl. affine rows (unnormalised; fine)
W32 = W.astype(np.float32)
def err(x): return np.abs(x.astype(np.float64) - ref).max() / unit
# SIMT: sequential fma in float32
acc = np.zeros((V, 3 * F), np.float32)
for k in range(N + 4):
    acc = (acc.astype(np.float64) + A32[:, k:k + 1].astype(np.float64) * W32[k:k + 1].astype(np.float64)).astype(np.float32)
print(f"N={N}: S={S:.1f}  SIMT fma chain: {err(acc):.2f} x 2^-24 S")
# exact-phi SIMT (isolates phi error)
acc = np.zeros((V, 3 * F), np.float32)
A32e = A64.astype(np.float32)
for k in range(N + 4):
    acc = (acc.astype(np.float64) + A32e[:, k:k + 1].astype(np.float64) * W32[k:k + 1].astype(np.float64)).astype(np.float32)
print(f"       SIMT with correctly rounded phi: {err(acc):.2f}")
# tensor variants
def split(x, parts):
    out, rem = [], x.astype(np.float64)
    for _ in range(parts):
        h = rem.astype(np.float16).astype(np.float64); out.append(h); rem = rem - h
    return out
colscale = 2.0 ** (14 - np.ceil(np.log2(np.abs(W).max(0) + 1e-300)))     # max|w| * scale in [8192, 16384)
Ws = W * colscale[None]
def rz32(x):  # round toward zero to float32
    y = x.astype(np.float32); bad = np.abs(y.astype(np.float64)) > np.abs(x)
    y[bad] = np.nextafter(y[bad], np.float32(0)); return y
def tensor(pa, pb, terms, trunc):
    Ap = split(A32.astype(np.float64) * 2.0 ** 14, pa); Bp = split(Ws, pb)
    acc = np.zeros((V, 3 * F), np.float32)
    K = N + 4
    for k0 in range(0, K, 16):
        for (i, j) in terms:
            part = Ap[i][:, k0:k0 + 16] @ Bp[j][k0:k0 + 16]                # exact products, ~exact 16-term sums in float64
            tot = acc.astype(np.float64) + part
            acc = rz32(tot) if trunc else tot.astype(np.float32)
    return acc.astype(np.float64) / colscale[None] / 2.0 ** 14
for name, pa, pb, terms in (("tensor3 (hh hl lh)", 2, 2, [(0, 0), (0, 1), (1, 0)]), ("tensor4 (+ll)", 2, 2, [(0, 0), (0, 1), (1, 0), (1, 1)]),
                            ("tensor6 (3-way)", 3, 3, [(0, 0), (0, 1), (1, 0), (0, 2), (2, 0), (1, 1)])):
    for trunc in (False, True):
        print(f"       {name:22s} accumulate {'RZ' if trunc else 'RN'}: {err(tensor(pa, pb, terms, trunc)):.2f}")


# ---- error-free slices (Ozaki scheme): what DESIGN.md section 9 item 1 proposes ---------------------------------------------
# Phi (computed in FP64) and the column-scaled weights are cut into SL-bit integer slices held in FP16; a slice-pair product is
# an exact integer, and K_BLK of them add exactly in an FP32 accumulator while SL_A + SL_B + log2(K_BLK) <= 24.  The pair sums
# are combined in FP64 (the epilogue).  The emulation checks the exactness claim (the largest |partial sum| in bits) and reports
# the end-to-end error, which is only the truncation of Phi and W to their slices.
def sliced(nsl_a, nsl_b, sl=8, k_blk=256, max_order=None):
    K = N + 4
    A = A64.copy()
    A[:, N + 1:] = A64[:, N + 1:]                     # affine columns: coordinates (|x| < 4 here); scaled like Phi below
    amax = np.abs(A).max(0)                           # per-K-row scale of the A operand (Phi <= 1; coordinates by their range)
    a_scale = 2.0 ** -np.ceil(np.log2(np.maximum(amax, 1e-300)))  # A * a_scale in (-1, 1]
    An = A * a_scale[None]
    Wn = (W / a_scale[:, None])                       # keep the product unchanged
    cs = 2.0 ** -np.ceil(np.log2(np.abs(Wn).max(0) + 1e-300))      # per-column power of two: |W| * cs <= 1
    Wn = Wn * cs[None]
    def cut(x, n):                                    # signed digits: x ~ sum_i d_i 2^(-sl (i+1)), |d_i| <= 2^(sl-1)
        out, rem = [], x.copy()
        for i in range(n):
            d = np.rint(rem * 2.0 ** (sl * (i + 1)))
            out.append(d)
            rem = rem - d * 2.0 ** (-sl * (i + 1))
        return out
    Ad, Wd = cut(An, nsl_a), cut(Wn, nsl_b)
    order = max_order if max_order is not None else max(nsl_a, nsl_b) - 1
    acc = np.zeros((V, 3 * F))
    worst_bits = 0.0
    for i in range(nsl_a):
        for j in range(nsl_b):
            if i + j > order:
                continue
            for k0 in range(0, K, k_blk):
                part = Ad[i][:, k0:k0 + k_blk] @ Wd[j][k0:k0 + k_blk]            # integers: exact in float64
                worst_bits = max(worst_bits, float(np.log2(np.abs(part).max() + 1)))
                acc += part * 2.0 ** (-sl * (i + j + 2))
    return acc / cs[None], worst_bits, sum(1 for i in range(nsl_a) for j in range(nsl_b) if i + j <= order)


for na, nb, sl in ((3, 3, 8), (4, 4, 8), (4, 4, 7)):
    out, bits, pairs = sliced(na, nb, sl=sl, k_blk=256)
    print(f"       sliced {na}x{nb} x {sl}-bit, K blocks of 256 ({pairs} MMAs): {err(out):.4f} x 2^-24 S   "
          f"(largest partial sum {bits:.1f} bits: {'exact' if bits <= 24 else 'NOT exact'} in an FP32 accumulator)")
