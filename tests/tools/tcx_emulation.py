"""CPU emulation of the exact-digit tensor-core evaluation (csrc/fd_eval_tcx.cu): column / row scaling, the digit width h from the
bound at the control points, integer leading digits + FP16 mid / lo remainders, the leading product summed exactly, the seven
remaining products accumulated in FP32 per K = 16 step either rounded to nearest or TRUNCATED.  The truncating variant reproduces
what the GPU measures (0.08 ... 0.17 x 2^-24 S at h = 8 ... 7 before the wider digit): tcgen05 adds into its FP32 accumulator
with round-toward-zero, and that bias -- not the representation -- is what is left of the error.
Usage: python tests/tools/tcx_emulation.py N   (N = 256: seconds; 1024: a minute)"""
import sys, numpy as np, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from facedeform_b200 import synth
from oracle import fd_oracle as o
N, F, V = int(sys.argv[1]), 40, 2048
rig = synth.control_rig(N); deform = synth.deformed_rig(rig, F)
mesh = synth.face_mesh(20000, topology=False)
P = np.ascontiguousarray(mesh.P[:V])
R = synth.default_radius("gaussian", rig.spacing)
p = o.make_params(model=1, term=0, kernel=0, radius=R, **{"lambda": 0.0})
st, rad, W = o.fit(p, rig.rest, deform)
C = rig.rest.astype(np.float64)
d2 = ((P[:, None, :].astype(np.float64) - C[None]) ** 2).sum(-1)
Phi = np.exp(-d2 / (R * R))
lo_, hi_ = C.min(0), C.max(0); oc = 0.5 * (lo_ + hi_); sc = 2.0 / (hi_ - lo_).max()
A = np.concatenate([Phi, np.ones((V, 1)), (P.astype(np.float64) - oc) * sc], 1)
Weff = W.copy()
Weff[N] = W[N] + (W[N + 1:N + 4] * oc[:, None]).sum(0)
Weff[N + 1:N + 4] = W[N + 1:N + 4] / sc
ref = A @ Weff
S = (np.abs(W[:N]).max(1)[None, :] * np.exp(-((C[:, None] - C[None]) ** 2).sum(-1) / (R * R))).sum(1).max()
unit = 2.0 ** -24 * S
K = N + 4
colmax = np.abs(Weff).max(0); ec = -np.ceil(np.log2(colmax)) ; ec = np.where(colmax * 2.0 ** ec >= 1.0, ec - 1, ec)
Wn = Weff * 2.0 ** ec[None]
rowmax = np.abs(Wn).max(1)
r = np.clip(-np.floor(np.log2(rowmax)) - 1, 0, 30).astype(int)   # rowmax in [2^-r-1, 2^-r)
s = r >> 1
PhiC = np.exp(-((C[:, None] - C[None]) ** 2).sum(-1) / (R * R))          # the bound at the control points (k_tcx_bound)
B = (PhiC * rowmax[None, :N]).sum(1).max() + rowmax[N] + 4.0 * rowmax[N + 1:N + 4].sum()
e = int(np.floor(np.log2(B))) + 1
h = min(11, max(2, (22 - e) // 2))                                       # 4 x 2^2h x B <= 2^24 (k_tcx_hbits)
print(f"N={N}: S={S:.1f} B={B:.2f} h={h}  r median {np.median(r)} min {r.min()} max {r.max()}")
Ax = A * 2.0 ** (h - s)[None, :]
Bx = Wn * 2.0 ** (h + s)[:, None]
def split(x):
    hi = np.rint(x.astype(np.float32)).astype(np.float64)
    rf = (x - hi).astype(np.float32)
    mid = rf.astype(np.float16)
    lo = (rf - mid.astype(np.float32)).astype(np.float16)
    return hi, mid.astype(np.float64), lo.astype(np.float64)
ah, am, al = split(Ax); bh, bm, bl = split(Bx)
acc0 = ah @ bh
print("  max |acc0 partial| bits:", np.log2(np.abs(np.cumsum(ah[:64, :, None] * bh[None, :, :8], 1)).max()))
def rz32(x):
    y = x.astype(np.float32); bad = np.abs(y.astype(np.float64)) > np.abs(x)
    y[bad] = np.nextafter(y[bad], np.float32(0)); return y
for trunc in (False, True):
    acc1 = np.zeros((V, 3 * F), np.float32)
    for k0 in range(0, K, 16):
        sl = slice(k0, k0 + 16)
        for (x, y) in ((ah, bm), (am, bh), (ah, bl), (al, bh), (am, bm)) + (((am, bl), (al, bm)) if h < 9 else ()):
            tot_ = acc1.astype(np.float64) + x[:, sl] @ y[sl]
            acc1 = rz32(tot_) if trunc else tot_.astype(np.float32)
    out = (acc0.astype(np.float32) + acc1).astype(np.float64) * 2.0 ** (-ec[None] - 2 * h)
    exact1 = (ah @ bm + am @ bh + ah @ bl + al @ bh + am @ bm) + ((am @ bl + al @ bm) if h < 9 else 0.0)
    out_e = (acc0 + exact1) * 2.0 ** (-ec[None] - 2 * h)
    print(f"  accumulate {'RZ' if trunc else 'RN'}: err {np.abs(out - ref).max() / unit:.4f} x 2^-24 S;  with exact acc1: {np.abs(out_e - ref).max() / unit:.5f};  |acc1| max {np.abs(exact1).max():.1f} vs S_units {S * 2.0 ** (ec.max() + 2 * h):.3g}")
