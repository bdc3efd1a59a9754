"""ctypes binding of the CPU oracle (oracle/libfd_oracle.so).

TEST INFRASTRUCTURE ONLY -- see oracle/fd_oracle.h.  Importable from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs; the product package (facedeform_b200) never imports it.
PARITY UNPINNED: the reference holds no golden vectors; the oracle is pinned against scipy and analytic
properties in tests/test_oracle_*.py.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libfd_oracle.so")

MODEL_QNN, MODEL_ML = 0, 1
TERM_LINEAR, TERM_CONST, TERM_ZERO = 0, 1, 2
KERNEL_GAUSSIAN, KERNEL_MULTIQUADRIC, KERNEL_THINPLATE = 0, 1, 2


class Params(C.Structure):
    _fields_ = [
        ("model", C.c_int32), ("term", C.c_int32), ("kernel", C.c_int32),
        ("qcoef", C.c_float), ("zcoef", C.c_float), ("radius", C.c_float),
        ("layers", C.c_int32), ("lambda_", C.c_float),
        ("tangent", C.c_int32), ("maxedges", C.c_int32),
        ("dofalloff", C.c_int32), ("falloffradius", C.c_float), ("falloffrate", C.c_float),
    ]


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "fd_oracle.c")
    hdr = os.path.join(_HERE, "fd_oracle.h")
    stale = (not os.path.exists(_LIB_PATH)) or any(
        os.path.exists(s) and os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in (src, hdr))
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "libfd_oracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        fp = C.POINTER(C.c_float)
        dp = C.POINTER(C.c_double)
        ip = C.POINTER(C.c_int32)
        lp = C.POINTER(C.c_int64)
        bp = C.POINTER(C.c_uint8)
        pp = C.POINTER(Params)
        L.fdo_params_default.argtypes = [pp]
        L.fdo_clamp_params.argtypes = [pp]
        L.fdo_poly_terms.argtypes = [C.c_int]
        L.fdo_poly_terms.restype = C.c_int
        L.fdo_pack.argtypes = [fp, fp, C.c_int, dp]
        L.fdo_radii.argtypes = [pp, fp, C.c_int, dp]
        L.fdo_radii.restype = C.c_int
        L.fdo_assemble.argtypes = [pp, fp, dp, C.c_int, dp]
        L.fdo_lu_factor.argtypes = [dp, C.c_int, ip]
        L.fdo_lu_factor.restype = C.c_int
        L.fdo_lu_solve.argtypes = [dp, ip, C.c_int, dp, C.c_int]
        L.fdo_fit.argtypes = [pp, fp, fp, C.c_int, C.c_int, dp, dp]
        L.fdo_fit.restype = C.c_int
        L.fdo_eval.argtypes = [pp, fp, dp, dp, C.c_int, C.c_int, fp, C.c_int64, fp, fp, fp, fp, fp, fp, C.c_int]
        L.fdo_eval_raw.argtypes = [pp, fp, dp, dp, C.c_int, C.c_int, fp, C.c_int64, dp, C.c_int]
        L.fdo_project_to_tangents.argtypes = [fp, fp, fp, fp]
        L.fdo_capture.argtypes = [fp, C.c_int64, ip, ip, C.c_int32, fp, C.c_int32, ip, ip, C.c_int32, ip,
                                  C.c_int32, C.c_float, C.c_int32, ip, bp, fp, ip, lp, ip, C.c_int32, C.c_int64]
        L.fdo_capture.restype = C.c_int
        L.fdo_point_tri_dist2.argtypes = [fp, fp, fp, fp]
        L.fdo_point_tri_dist2.restype = C.c_float
        L.fdo_point_seg_dist2.argtypes = [fp, fp, fp]
        L.fdo_point_seg_dist2.restype = C.c_float
        L.fdo_num_threads.restype = C.c_int
        L.fdo_fit_v1.argtypes = [pp, fp, fp, C.c_int, C.c_int, fp, dp, dp]
        L.fdo_fit_v1.restype = C.c_int
        L.fdo_dbse_shapes_matrix.argtypes = [fp, fp, C.c_int64, C.c_int32, dp]
        L.fdo_householder_qr.argtypes = [dp, C.c_int64, C.c_int32, dp]
        L.fdo_dbse_weights.argtypes = [dp, C.c_int64, C.c_int32, fp, fp, dp]
        L.fdo_dbse_displace.argtypes = [dp, C.c_int64, C.c_int32, dp, fp, C.c_int32, C.c_float, fp, fp, fp]
        _lib = L
    return _lib


def _f(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_float))


def _d(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _i(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_int32))


def _cf(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float32)


def make_params(**kw) -> Params:
    p = Params()
    lib().fdo_params_default(C.byref(p))
    for k, v in kw.items():
        if k == "lambda":
            k = "lambda_"
        if not hasattr(p, k):
            raise AttributeError(k)
        setattr(p, k, v)
    return p


def clamp_params(p: Params) -> Params:
    lib().fdo_clamp_params(C.byref(p))
    return p


def poly_terms(term: int) -> int:
    return lib().fdo_poly_terms(term)


def num_threads() -> int:
    return lib().fdo_num_threads()


def pack(rest, deform):
    rest, deform = _cf(rest), _cf(deform)
    n = rest.shape[0]
    out = np.empty((n, 6), np.float64)
    lib().fdo_pack(_f(rest), _f(deform), n, _d(out))
    return out


def radii(p: Params, rest):
    rest = _cf(rest)
    out = np.empty(rest.shape[0], np.float64)
    st = lib().fdo_radii(C.byref(p), _f(rest), rest.shape[0], _d(out))
    return st, out


def assemble(p: Params, rest, rad):
    rest = _cf(rest)
    n = rest.shape[0] + poly_terms(p.term)
    A = np.empty((n, n), np.float64)
    lib().fdo_assemble(C.byref(p), _f(rest), _d(rad), rest.shape[0], _d(A))
    return A


def lu_factor(A):
    A = np.array(A, dtype=np.float64, order="C")
    piv = np.empty(A.shape[0], np.int32)
    st = lib().fdo_lu_factor(_d(A), A.shape[0], _i(piv))
    return st, A, piv


def lu_solve(LU, piv, B):
    B = np.array(B, dtype=np.float64, order="C")
    lib().fdo_lu_solve(_d(LU), _i(piv), LU.shape[0], _d(B), B.shape[1])
    return B


def fit(p: Params, rest, deform):
    """deform: (F, N, 3) or (N, 3). returns (status, radii[N], weights[(N+p), 3F])."""
    rest, deform = _cf(rest), _cf(deform)
    if deform.ndim == 2:
        deform = deform[None]
    F, N = deform.shape[0], rest.shape[0]
    assert deform.shape == (F, N, 3)
    n = N + poly_terms(p.term)
    rad = np.empty(N, np.float64)
    W = np.empty((n, 3 * F), np.float64)
    st = lib().fdo_fit(C.byref(p), _f(rest), _f(deform), N, F, _d(rad), _d(W))
    return st, rad, W


def evaluate(p: Params, rest, rad, W, P, dist2=None, tu=None, tv=None, nrm=None, nthreads=1):
    """returns (P_out[F, V, 3] f32, falloff[V] f32) -- the loop at SOP_FaceDeform.cpp:404-439."""
    rest, P = _cf(rest), _cf(P)
    dist2, tu, tv, nrm = _cf(dist2), _cf(tu), _cf(tv), _cf(nrm)
    N, V, F = rest.shape[0], P.shape[0], W.shape[1] // 3
    W = np.ascontiguousarray(W, np.float64)
    out = np.empty((F, V, 3), np.float32)
    fo = np.empty(V, np.float32)
    lib().fdo_eval(C.byref(p), _f(rest), _d(rad), _d(W), N, F, _f(P), V, _f(dist2), _f(tu), _f(tv), _f(nrm),
                   _f(out), _f(fo), int(nthreads))
    return out, fo


def evaluate_raw(p: Params, rest, rad, W, P, nthreads=1):
    rest, P = _cf(rest), _cf(P)
    N, V, F = rest.shape[0], P.shape[0], W.shape[1] // 3
    W = np.ascontiguousarray(W, np.float64)
    out = np.empty((V, 3 * F), np.float64)
    lib().fdo_eval_raw(C.byref(p), _f(rest), _d(rad), _d(W), N, F, _f(P), V, _d(out), int(nthreads))
    return out


def project_to_tangents(u, v, n, disp):
    u, v, n = _cf(u), _cf(v), _cf(n)
    d = np.array(disp, dtype=np.float32)
    lib().fdo_project_to_tangents(_f(u), _f(v), _f(n), _f(d))
    return d


def point_tri_dist2(p, a, b, c):
    p, a, b, c = _cf(p), _cf(a), _cf(b), _cf(c)
    return float(lib().fdo_point_tri_dist2(_f(p), _f(a), _f(b), _f(c)))


def point_seg_dist2(p, a, b):
    p, a, b = _cf(p), _cf(a), _cf(b)
    return float(lib().fdo_point_seg_dist2(_f(p), _f(a), _f(b)))


def capture(P, poly_off, poly_vtx, rigP, rig_off, rig_vtx, rig_class, max_edges, radius, dofalloff):
    """returns dict(ngroups, nearest_idx, member, dist2, grp_class, grp_off, grp_idx)."""
    P, rigP = _cf(P), _cf(rigP)
    V, N = P.shape[0], rigP.shape[0]
    poly_off = np.ascontiguousarray(poly_off, np.int32)
    poly_vtx = np.ascontiguousarray(poly_vtx, np.int32)
    rig_off = np.ascontiguousarray(rig_off if rig_off is not None else [0], np.int32)
    rig_vtx = np.ascontiguousarray(rig_vtx if rig_vtx is not None else [], np.int32)
    if rig_vtx.size == 0:
        rig_vtx = np.zeros(1, np.int32)
    rc = None if rig_class is None else np.ascontiguousarray(rig_class, np.int32)
    nearest = np.empty(N, np.int32)
    member = np.empty(V, np.uint8)
    dist2 = np.empty(V, np.float32)
    cap = N + 1
    gclass = np.empty(cap, np.int32)
    goff = np.empty(cap + 1, np.int64)
    lp = C.POINTER(C.c_int64)
    bp = C.POINTER(C.c_uint8)

    def call(gidx, idx_cap):
        return lib().fdo_capture(_f(P), V, _i(poly_off), _i(poly_vtx), len(poly_off) - 1, _f(rigP), N,
                                 _i(rig_off), _i(rig_vtx), len(rig_off) - 1, _i(rc), int(max_edges),
                                 float(radius), int(dofalloff), _i(nearest), member.ctypes.data_as(bp), _f(dist2),
                                 _i(gclass), goff.ctypes.data_as(lp), _i(gidx), cap, idx_cap)

    g = call(None, 0)
    if g < 0:
        raise RuntimeError("fdo_capture: group capacity")
    total = int(goff[g]) if g > 0 else 0
    gidx = np.empty(max(total, 1), np.int32)
    g = call(gidx, total)
    return dict(ngroups=g, nearest_idx=nearest, member=member.astype(bool), dist2=dist2,
                grp_class=gclass[:g].copy(), grp_off=goff[:g + 1].copy(), grp_idx=gidx[:total].copy())


# ---- DirectBSEdit (reference src/dbse.cpp) --------------------------------------------------------------------------

def dbse_shapes_matrix(rest, shapes):
    """(3P, S) matrix of FP32 shape deltas widened to FP64 (dbse.cpp:9-35); returned column-major (Fortran order)."""
    rest = np.ascontiguousarray(rest, dtype=np.float32)
    shapes = np.ascontiguousarray(shapes, dtype=np.float32)
    S, P = shapes.shape[0], rest.shape[0]
    M = np.empty((3 * P, S), dtype=np.float64, order="F")
    lib().fdo_dbse_shapes_matrix(_f(rest), _f(shapes), P, S, M.ctypes.data_as(C.POINTER(C.c_double)))
    return M


def householder_qr(M):
    """packed Householder QR (R above, essential reflector parts below the diagonal) and tau, LAPACK conventions."""
    A = np.array(M, dtype=np.float64, order="F", copy=True)
    tau = np.zeros(A.shape[1], dtype=np.float64)
    lib().fdo_householder_qr(A.ctypes.data_as(C.POINTER(C.c_double)), A.shape[0], A.shape[1], _d(tau))
    return A, tau


def dbse_weights(QR, pos, rest):
    pos = np.ascontiguousarray(pos, dtype=np.float32)
    rest = np.ascontiguousarray(rest, dtype=np.float32)
    QR = np.asfortranarray(QR, dtype=np.float64)
    w = np.zeros(QR.shape[1], dtype=np.float64)
    lib().fdo_dbse_weights(QR.ctypes.data_as(C.POINTER(C.c_double)), rest.shape[0], QR.shape[1], _f(pos), _f(rest), _d(w))
    return w


def dbse_displace(M, weights, pos, rest, weightrange=None, dofalloff=0, falloffradius=1.0):
    pos = np.ascontiguousarray(pos, dtype=np.float32)
    rest = np.ascontiguousarray(rest, dtype=np.float32)
    M = np.asfortranarray(M, dtype=np.float64)
    weights = np.ascontiguousarray(weights, dtype=np.float64)
    out = np.empty_like(rest)
    wr = None if weightrange is None else np.ascontiguousarray(weightrange, dtype=np.float32)
    lib().fdo_dbse_displace(M.ctypes.data_as(C.POINTER(C.c_double)), rest.shape[0], M.shape[1], _d(weights),
                            None if wr is None else _f(wr), int(dofalloff), float(falloffradius), _f(pos), _f(rest), _f(out))
    return out


# ---- "ALGLIB v1 like" fit (unverifiable documentation mode, fd_oracle.h) -----------------------------------------------

def fit_v1(p: Params, rest, deform):
    """two-stage polynomial + QNN single layer / Multilayer `layers` halving radii.
    returns (status, centres[N*L, 3], radii[N*L], weights[(N*L + npoly), 3F]): a stacked model for evaluate()."""
    rest, deform = _cf(rest), _cf(deform)
    if deform.ndim == 2:
        deform = deform[None]
    F, N = deform.shape[0], rest.shape[0]
    L = 1 if p.model == MODEL_QNN else max(1, int(p.layers))
    cen = np.empty((N * L, 3), np.float32)
    rad = np.empty(N * L, np.float64)
    W = np.empty((N * L + poly_terms(p.term), 3 * F), np.float64)
    st = lib().fdo_fit_v1(C.byref(p), _f(rest), _f(deform), N, F, _f(cen), _d(rad), _d(W))
    return st, cen, rad, W
