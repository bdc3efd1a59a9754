"""Python mirror of the C ABI (include/facedeform_gpu.h): thin, no arithmetic of its own.

Host arrays (numpy) go through the host-pointer entry points (fd_rbf_fit / fd_rbf_solve / fd_rbf_eval: the
library copies in and out); torch CUDA tensors go through the `*_dev` entry points on the ctx stream.
The names follow the reference's cook (SOP_FaceDeform.cpp:215-489): fit = rbfcreate..rbfbuildmodel,
solve = the per-frame deltas, eval = the vertex loop, capture = ProximityCapture.
"""
from __future__ import annotations

import ctypes as C
import weakref

import numpy as np

from . import _lib
from ._lib import FdParams, FdReport

FD_OK, FD_E_INVALID, FD_E_MISMATCH_POINT, FD_E_BUILD, FD_E_SINGULAR = 0, 1, 2, 3, 4
FD_E_CUDA, FD_E_NOMEM, FD_E_CAPTURE, FD_E_UNSUPPORTED, FD_E_STATE = 5, 6, 7, 8, 9
MODEL_QNN, MODEL_ML = 0, 1
TERM_LINEAR, TERM_CONST, TERM_ZERO = 0, 1, 2
KERNEL_GAUSSIAN, KERNEL_MULTIQUADRIC, KERNEL_THINPLATE = 0, 1, 2
EVAL_AUTO, EVAL_FP32, EVAL_FP64 = 0, 1, 2
PATH_AUTO, PATH_SIMT, PATH_TENSOR = 0, 1, 2
PHASES = {"assemble": 0, "factor": 1, "solve": 2, "eval": 3}


class FdError(RuntimeError):
    """A non-zero fd_status; `.status` is the code, the message is fd_last_error()."""

    def __init__(self, status: int, message: str):
        super().__init__(f"[fd_status {status}: {_lib.load().fd_status_string(status).decode()}] {message}")
        self.status = status


def make_params(clamp: bool = False, **kw) -> FdParams:
    """fd_params with the SOP defaults (SOP_FaceDeform.cpp:117-137); clamp=True applies :249-257."""
    p = FdParams()
    _lib.load().fd_params_default(C.byref(p))
    for k, v in kw.items():
        if k == "lambda":
            k = "lambda_"
        if k == "weightrange":
            p.weightrange[0], p.weightrange[1] = v
            continue
        if k == "group":
            p.group = v.encode() if isinstance(v, str) else v
            continue
        if not hasattr(p, k):
            raise AttributeError(k)
        setattr(p, k, v)
    if clamp:
        _lib.load().fd_params_clamp(C.byref(p))
    return p


def _ptr(a):
    """address of a numpy array or torch tensor (None -> NULL)."""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    return a.data_ptr()


def _host_f32(a, shape_last=None):
    if a is None:
        return None
    a = np.ascontiguousarray(a, dtype=np.float32)
    if shape_last is not None and (a.ndim == 0 or a.shape[-1] != shape_last):
        raise ValueError(f"expected (..., {shape_last}) array, got {a.shape}")
    return a


def _is_torch_cuda(a) -> bool:
    return a is not None and not isinstance(a, np.ndarray) and hasattr(a, "is_cuda") and a.is_cuda


class Context:
    """fd_ctx: one GPU + stream.  stream: a raw cudaStream_t value (e.g. torch.cuda.current_stream().cuda_stream)."""

    def __init__(self, device: int = -1, stream: int | None = None):
        self._L = _lib.load()
        h = C.c_void_p()
        st = self._L.fd_ctx_create(C.byref(h), int(device), C.c_void_p(stream) if stream else None)
        if st != FD_OK:
            raise FdError(st, "fd_ctx_create failed (no B200 visible?)")
        self._h = h
        self._models = weakref.WeakSet()   # live models: destroyed before the ctx they point into

    def close(self):
        if getattr(self, "_h", None):
            for m in list(self._models):
                m.close()
            self._L.fd_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, st: int):
        if st != FD_OK:
            raise FdError(st, self._L.fd_last_error(self._h).decode(errors="replace"))

    def synchronize(self):
        self._check(self._L.fd_ctx_synchronize(self._h))

    def phase_ms(self, phase: str) -> float:
        return float(self._L.fd_ctx_phase_ms(self._h, PHASES[phase]))

    @property
    def launch_count(self) -> int:
        return int(self._L.fd_ctx_launch_count(self._h))

    # ---- fit ------------------------------------------------------------------------------------------------
    def fit(self, params: FdParams, rest_ctrl) -> "RbfModel":
        """assemble + factor once per rest pose (replaces SOP_FaceDeform.cpp:331-363)."""
        h = C.c_void_p()
        if _is_torch_cuda(rest_ctrl):
            n = rest_ctrl.shape[0]
            self._check(self._L.fd_rbf_fit_dev(self._h, C.byref(params), _ptr(rest_ctrl), n, C.byref(h)))
            return RbfModel(self, h, n, params, keep=rest_ctrl)
        rest = _host_f32(rest_ctrl, 3)
        rep = FdReport()
        self._check(self._L.fd_rbf_fit(self._h, C.byref(params), _ptr(rest), rest.shape[0], C.byref(h), C.byref(rep)))
        m = RbfModel(self, h, rest.shape[0], params)
        m.last_report = rep
        return m

    def load_model(self, blob) -> "RbfModel":
        """fd_model_load: a model saved with RbfModel.save (alglib::rbfunserialize in the reference's dead threaded path,
        SOP_FaceDeform.hpp:150-152)."""
        buf = np.frombuffer(bytes(blob), dtype=np.uint8)
        h = C.c_void_p()
        self._check(self._L.fd_model_load(self._h, buf.ctypes.data, buf.size, C.byref(h)))
        n, f = C.c_int32(), C.c_int32()
        self._L.fd_model_info(h, C.byref(n), None, C.byref(f), None)
        m = RbfModel(self, h, n.value, None)
        m.frames = f.value
        return m

    def receiver(self, params: FdParams, rest_ctrl, frames: int) -> "RbfModel":
        """a model that receives broadcast weights instead of solving (multi-GPU ranks != root)."""
        rest = rest_ctrl if _is_torch_cuda(rest_ctrl) else _host_f32(rest_ctrl, 3)
        h = C.c_void_p()
        self._check(self._L.fd_model_create_receiver(self._h, C.byref(params), _ptr(rest), rest.shape[0], int(frames),
                                                     C.byref(h)))
        m = RbfModel(self, h, rest.shape[0], params)
        m.frames = int(frames)
        return m

    # ---- capture --------------------------------------------------------------------------------------------
    def capture(self, P, poly_off, poly_vtx, rig_P, rig_off=None, rig_vtx=None, rig_class=None, max_edges=4,
                radius=1.0, dofalloff=0):
        """ProximityCapture::init + capture (capture.cpp:10-141).  Raises FdError(FD_E_CAPTURE) when no group forms."""
        P = _host_f32(P, 3)
        rig_P = _host_f32(np.zeros((0, 3)) if rig_P is None else rig_P, 3)
        V, N = P.shape[0], rig_P.shape[0]
        i32 = lambda a, d: np.ascontiguousarray(d if a is None else a, dtype=np.int32)
        poly_off, poly_vtx = i32(poly_off, [0]), i32(poly_vtx, [])
        rig_off, rig_vtx = i32(rig_off, [0]), i32(rig_vtx, [])
        rc = None if rig_class is None else np.ascontiguousarray(rig_class, dtype=np.int32)
        nearest = np.empty(max(N, 1), np.int32)
        member = np.empty(max(V, 1), np.uint8)
        dist2 = np.empty(max(V, 1), np.float32)
        cap = N + 1
        gclass = np.empty(cap, np.int32)
        goff = np.empty(cap + 1, np.int64)
        ng = C.c_int32(0)

        def call(gidx, idx_cap):
            return self._L.fd_capture(self._h, _ptr(P), V, _ptr(poly_off), _ptr(poly_vtx), len(poly_off) - 1,
                                      _ptr(rig_P), N, _ptr(rig_off), _ptr(rig_vtx), len(rig_off) - 1, _ptr(rc),
                                      int(max_edges), float(radius), int(dofalloff), _ptr(nearest), _ptr(member),
                                      _ptr(dist2), C.byref(ng), _ptr(gclass), _ptr(goff), _ptr(gidx), cap, idx_cap)

        self._check(call(None, 0))
        g = ng.value
        total = int(goff[g])
        gidx = np.empty(max(total, 1), np.int32)
        self._check(call(gidx, total))
        return dict(ngroups=g, nearest_idx=nearest[:N].copy(), member=member[:V].astype(bool), dist2=dist2[:V].copy(),
                    grp_class=gclass[:g].copy(), grp_off=goff[:g + 1].copy(), grp_idx=gidx[:total].copy())


class DirectBSEdit:
    """fd_dbse: the "morph space" post-pass of the SOP (reference src/dbse.{hpp,cpp}, same entry-point names).

    init happens in the constructor (DirectBSEdit::init, dbse.cpp:9-35): `rest` (P, 3) and `shapes` (S, P, 3)."""

    def __init__(self, ctx: Context, rest, shapes):
        self.ctx, self._L = ctx, ctx._L
        rest = _host_f32(rest, 3)
        shapes = np.ascontiguousarray(shapes, dtype=np.float32)
        if shapes.ndim != 3 or shapes.shape[1:] != rest.shape:
            raise ValueError("shapes must be (S, P, 3) matching rest (P, 3)")
        h = C.c_void_p()
        ctx._check(self._L.fd_dbse_init(ctx._h, _ptr(rest), rest.shape[0], _ptr(shapes), shapes.shape[0], C.byref(h)))
        self._h, self.n_pts, self.n_shapes = h, rest.shape[0], shapes.shape[0]
        ctx._models.add(self)

    def close(self):
        if getattr(self, "_h", None):
            self._L.fd_dbse_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def is_initialized(self) -> bool:                      # DirectBSEdit::isInitialized
        return bool(self._h)

    def is_computed(self) -> bool:                         # DirectBSEdit::isComputed
        c = C.c_int32()
        self._L.fd_dbse_info(self._h, None, None, C.byref(c))
        return bool(c.value)

    def compute_weights(self, pos, rest):
        """DirectBSEdit::computeWeights (dbse.cpp:37-58); returns the S weights."""
        pos, rest = _host_f32(pos, 3), _host_f32(rest, 3)
        w = np.empty(self.n_shapes, dtype=np.float64)
        self.ctx._check(self._L.fd_dbse_compute_weights(self._h, _ptr(pos), _ptr(rest), _ptr(w)))
        return w

    def displace(self, pos, rest, weightrange=None, dofalloff=0, falloffradius=1.0):
        """displaceVector over all points + the SOP's position write (dbse.cpp:60-75, SOP_FaceDeform.cpp:460-472)."""
        pos, rest = _host_f32(pos, 3), _host_f32(rest, 3)
        out = np.empty_like(rest)
        wr = None if weightrange is None else np.ascontiguousarray(weightrange, dtype=np.float32)
        self.ctx._check(self._L.fd_dbse_displace(self._h, _ptr(pos), _ptr(rest), 0 if wr is None else 1, _ptr(wr),
                                                 int(dofalloff), float(falloffradius), _ptr(out)))
        return out

    def get_weights(self):
        """DirectBSEdit::getWeights (dbse.cpp:77-87): raises FdError(FD_E_STATE) before compute_weights."""
        w = np.empty(self.n_shapes, dtype=np.float64)
        self.ctx._check(self._L.fd_dbse_get_weights(self._h, _ptr(w)))
        return w

    def packed_qr(self):
        qr = np.empty((3 * self.n_pts, self.n_shapes), dtype=np.float64, order="F")
        tau = np.empty(self.n_shapes, dtype=np.float64)
        self.ctx._check(self._L.fd_dbse_get_qr(self._h, _ptr(qr), _ptr(tau)))
        return qr, tau


class RbfModel:
    """fd_model: centres, radii, LU factors and the weights of the last solve (replaces alglib::rbfmodel)."""

    def __init__(self, ctx: Context, handle, n_ctrl: int, params: FdParams, keep=None):
        self.ctx, self._h, self.n_ctrl, self.params = ctx, handle, n_ctrl, params
        self._L = ctx._L
        self.frames = 0
        self.last_report = None
        self._keep = keep
        ctx._models.add(self)

    def close(self):
        if getattr(self, "_h", None):
            self._L.fd_model_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def report(self) -> FdReport:
        """synchronises and returns the status of the last fit/solve; raises FdError(FD_E_SINGULAR) like :365-368."""
        rep = FdReport()
        st = self._L.fd_model_report(self._h, C.byref(rep))
        self.last_report = rep
        self.ctx._check(st)
        return rep

    def solve(self, deform_ctrl):
        """weights for F frames at once; deform_ctrl (F, N, 3) or (N, 3) (replaces :268-287 deltas + :363)."""
        if _is_torch_cuda(deform_ctrl):
            d = deform_ctrl if deform_ctrl.dim() == 3 else deform_ctrl[None]
            F, n = d.shape[0], d.shape[1]
            self.ctx._check(self._L.fd_rbf_solve_dev(self._h, _ptr(d), n, F))
            self.frames = F
            self._keep_def = d
            return self
        d = _host_f32(deform_ctrl, 3)
        if d.ndim == 2:
            d = d[None]
        F, n = d.shape[0], d.shape[1]
        rep = FdReport()
        st = self._L.fd_rbf_solve(self._h, _ptr(d), n, F, C.byref(rep))
        self.last_report = rep
        self.ctx._check(st)
        self.frames = F
        return self

    def eval(self, P, dist2=None, tangentu=None, tangentv=None, normal=None, out=None, falloff_out=None,
             want_falloff=True):
        """the vertex loop (:384-439).  numpy in -> (P_out[F, V, 3], falloff[V]) numpy; torch CUDA in -> torch out."""
        if _is_torch_cuda(P):
            import torch
            V = P.shape[0]
            if out is None:
                out = torch.empty((self.frames, V, 3), dtype=torch.float32, device=P.device)
            if falloff_out is None and want_falloff:
                falloff_out = torch.empty((V,), dtype=torch.float32, device=P.device)
            self.ctx._check(self._L.fd_rbf_eval_dev(self._h, _ptr(P), V, _ptr(dist2), _ptr(tangentu), _ptr(tangentv),
                                                    _ptr(normal), _ptr(out), _ptr(falloff_out)))
            return out, falloff_out
        P = _host_f32(P, 3)
        V = P.shape[0]
        dist2 = _host_f32(dist2)
        tangentu, tangentv, normal = _host_f32(tangentu, 3), _host_f32(tangentv, 3), _host_f32(normal, 3)
        if out is None:
            out = np.empty((self.frames, V, 3), np.float32)
        if falloff_out is None and want_falloff:
            falloff_out = np.empty(V, np.float32)
        self.ctx._check(self._L.fd_rbf_eval(self._h, _ptr(P), V, _ptr(dist2), _ptr(tangentu), _ptr(tangentv),
                                            _ptr(normal), _ptr(out), _ptr(falloff_out)))
        return out, falloff_out

    # ---- weights ----------------------------------------------------------------------------------------------
    def info(self):
        n, p, f, ld = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32()
        self.ctx._check(self._L.fd_model_info(self._h, C.byref(n), C.byref(p), C.byref(f), C.byref(ld)))
        return dict(n_ctrl=n.value, npoly=p.value, frames=f.value, weights_ld=ld.value)

    def weights(self):
        """(weights[(N + npoly), 3F] float64, radii[N] float64) copied to the host."""
        i = self.info()
        W = np.empty((i["n_ctrl"] + i["npoly"], 3 * i["frames"]), np.float64)
        R = np.empty(i["n_ctrl"], np.float64)
        self.ctx._check(self._L.fd_model_get_weights(self._h, _ptr(W), _ptr(R)))
        return W, R

    def weights_dev(self):
        """(device address, bytes) of the FP64 weight block, for the caller's broadcast."""
        p, b = C.c_void_p(), C.c_size_t()
        self.ctx._check(self._L.fd_model_weights_dev(self._h, C.byref(p), C.byref(b)))
        return p.value, b.value

    def radii_dev(self):
        p, b = C.c_void_p(), C.c_size_t()
        self.ctx._check(self._L.fd_model_radii_dev(self._h, C.byref(p), C.byref(b)))
        return p.value, b.value

    def commit_weights(self):
        self.ctx._check(self._L.fd_model_commit_weights(self._h))
        return self

    def set_epilogue(self, params: FdParams):
        """epilogue-only parameter changes (tangent, falloff ...) without a refit; raises when `params` needs one."""
        self.ctx._check(self._L.fd_model_set_epilogue(self._h, C.byref(params)))
        self.params = params
        return self

    def save(self) -> bytes:
        """fd_model_save: parameters, centres, radii, factorisation and weights (the reference's rbfserialize, :377)."""
        n = C.c_size_t()
        self.ctx._check(self._L.fd_model_save(self._h, None, 0, C.byref(n)))
        buf = np.empty(n.value, np.uint8)
        self.ctx._check(self._L.fd_model_save(self._h, buf.ctypes.data, buf.size, C.byref(n)))
        return buf.tobytes()


MGPU_AUTO, MGPU_NCCL, MGPU_P2P = 0, 1, 2


class MultiGpu:
    """fd_mgpu: several GPUs of one box behind one handle (single process); vertex ranges per device, the root's weights
    cross NVLink once per solve (transport "p2p": tables built through peer loads; "nccl": ncclBroadcast)."""

    def __init__(self, devices=None, transport: int = MGPU_AUTO):
        self._L = _lib.load()
        h = C.c_void_p()
        if devices is None:
            import torch
            devices = list(range(torch.cuda.device_count()))
        arr = np.ascontiguousarray(devices, dtype=np.int32)
        st = self._L.fd_mgpu_create(C.byref(h), arr.ctypes.data, len(arr), int(transport))
        if st != FD_OK:
            raise FdError(st, f"fd_mgpu_create({list(devices)}, transport={transport}) failed")
        self._h, self.devices, self.frames = h, list(devices), 0

    def close(self):
        if getattr(self, "_h", None):
            self._L.fd_mgpu_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, st):
        if st != FD_OK:
            raise FdError(st, self._L.fd_mgpu_last_error(self._h).decode(errors="replace"))

    def fit(self, params: FdParams, rest_ctrl):
        rest = _host_f32(rest_ctrl, 3)
        rep = FdReport()
        self._check(self._L.fd_mgpu_fit(self._h, C.byref(params), _ptr(rest), rest.shape[0], C.byref(rep)))
        self.last_report = rep
        return self

    def solve(self, deform_ctrl):
        d = _host_f32(deform_ctrl, 3)
        if d.ndim == 2:
            d = d[None]
        rep = FdReport()
        self._check(self._L.fd_mgpu_solve(self._h, _ptr(d), d.shape[1], d.shape[0], C.byref(rep)))
        self.last_report, self.frames = rep, d.shape[0]
        return self

    def eval(self, P, dist2=None, tangentu=None, tangentv=None, normal=None, out=None, falloff_out=None):
        P = _host_f32(P, 3)
        V = P.shape[0]
        dist2 = _host_f32(dist2)
        tangentu, tangentv, normal = _host_f32(tangentu, 3), _host_f32(tangentv, 3), _host_f32(normal, 3)
        if out is None:
            out = np.empty((self.frames, V, 3), np.float32)
        if falloff_out is None:
            falloff_out = np.empty(V, np.float32)
        self._check(self._L.fd_mgpu_eval(self._h, _ptr(P), V, _ptr(dist2), _ptr(tangentu), _ptr(tangentv), _ptr(normal),
                                         _ptr(out), _ptr(falloff_out)))
        return out, falloff_out

    def info(self):
        n, t, b, ms = C.c_int32(), C.c_int32(), C.c_int64(), C.c_float()
        self._check(self._L.fd_mgpu_info(self._h, C.byref(n), C.byref(t), C.byref(b), C.byref(ms)))
        return dict(ndev=n.value, transport={MGPU_NCCL: "nccl", MGPU_P2P: "p2p"}.get(t.value, "?"), bcast_bytes=b.value,
                    bcast_ms=ms.value)

    def vertex_range(self, i: int, n_vtx: int):
        b, e = C.c_int64(), C.c_int64()
        self._check(self._L.fd_mgpu_range(self._h, i, n_vtx, C.byref(b), C.byref(e)))
        return b.value, e.value
