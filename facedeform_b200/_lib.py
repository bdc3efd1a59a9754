"""Loader of libfacedeform_gpu.so (the C ABI in include/facedeform_gpu.h).

There is no CPU fallback: if the CUDA library is missing or cannot be loaded this module raises, and every
numeric entry point of the package fails with it.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libfacedeform_gpu.so")

# every symbol include/facedeform_gpu.h declares (tests/test_abi.py checks the header against this list)
EXPORTS = [
    "fd_abi_version", "fd_status_string", "fd_params_default", "fd_params_clamp", "fd_params_fit_equal",
    "fd_ctx_create", "fd_ctx_destroy", "fd_ctx_synchronize", "fd_last_error",
    "fd_rbf_fit", "fd_rbf_fit_dev", "fd_model_destroy",
    "fd_rbf_solve", "fd_rbf_solve_dev", "fd_model_report", "fd_model_set_epilogue", "fd_model_save", "fd_model_load",
    "fd_mgpu_create", "fd_mgpu_destroy", "fd_mgpu_fit", "fd_mgpu_solve", "fd_mgpu_eval", "fd_mgpu_info", "fd_mgpu_range",
    "fd_mgpu_ctx", "fd_mgpu_last_error",
    "fd_rbf_eval", "fd_rbf_eval_dev",
    "fd_model_create_receiver", "fd_model_weights_dev", "fd_model_radii_dev", "fd_model_commit_weights",
    "fd_model_info", "fd_model_get_weights",
    "fd_capture", "fd_ctx_phase_ms", "fd_ctx_launch_count",
    "fd_dbse_init", "fd_dbse_compute_weights", "fd_dbse_displace", "fd_dbse_get_weights", "fd_dbse_get_qr", "fd_dbse_info",
    "fd_dbse_destroy",
    "fd_sop_create", "fd_sop_destroy", "fd_sop_params", "fd_sop_cook", "fd_sop_messages", "fd_sop_fit_count", "fd_sop_positions_bumped",
    "fd_sop_set_blendshapes", "fd_sop_blend_weights",
]


class FdParams(C.Structure):
    """struct fd_params -- the SOP's parameter surface (SOP_FaceDeform.cpp:99-137)."""
    _fields_ = [
        ("model", C.c_int32), ("term", C.c_int32), ("kernel", C.c_int32),
        ("qcoef", C.c_float), ("zcoef", C.c_float), ("radius", C.c_float),
        ("layers", C.c_int32), ("lambda_", C.c_float),
        ("tangent", C.c_int32), ("maxedges", C.c_int32), ("morphspace", C.c_int32), ("doclampweight", C.c_int32),
        ("weightrange", C.c_float * 2),
        ("dofalloff", C.c_int32), ("falloffradius", C.c_float), ("falloffrate", C.c_float),
        ("eval_precision", C.c_int32), ("eval_path", C.c_int32), ("factor_precision", C.c_int32),
        ("fidelity", C.c_int32), ("strict_reference", C.c_int32), ("eval_tolerance", C.c_float), ("group", C.c_char * 64),
    ]


class FdReport(C.Structure):
    """struct fd_report -- analogue of alglib::rbfreport (SOP_FaceDeform.cpp:365-373)."""
    _fields_ = [
        ("terminationtype", C.c_int32), ("iterationscount", C.c_int32), ("n", C.c_int32), ("npoly", C.c_int32),
        ("frames", C.c_int32), ("reserved", C.c_int32), ("min_pivot", C.c_double), ("max_pivot", C.c_double),
        ("residual", C.c_double), ("cancellation", C.c_double), ("eval_kernel", C.c_int32), ("eval_inexact", C.c_int32),
    ]


_lib = None


def load() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        # the library is built in-tree by __graft_entry__.build(); on a checkout without it, try once with nvcc
        import shutil
        import subprocess
        if shutil.which("make") and (shutil.which("nvcc") or os.path.exists("/usr/local/cuda/bin/nvcc")):
            subprocess.run(["make", "-C", os.path.join(_HERE, "csrc"), "-j8", "all"], check=False,
                           stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `make -C facedeform_b200/csrc` (or __graft_entry__.build()). "
            "facedeform_b200 has no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp, fp, ip = C.c_void_p, C.c_void_p, C.c_void_p  # raw addresses: host numpy or device pointers alike
    pp, rp = C.POINTER(FdParams), C.POINTER(FdReport)
    L.fd_abi_version.restype = C.c_int
    L.fd_status_string.argtypes = [C.c_int]
    L.fd_status_string.restype = C.c_char_p
    L.fd_params_default.argtypes = [pp]
    L.fd_params_default.restype = None
    L.fd_params_clamp.argtypes = [pp]
    L.fd_params_clamp.restype = None
    L.fd_params_fit_equal.argtypes = [pp, pp]
    L.fd_model_set_epilogue.argtypes = [vp, pp]
    L.fd_model_save.argtypes = [vp, vp, C.c_size_t, C.POINTER(C.c_size_t)]
    L.fd_model_load.argtypes = [vp, vp, C.c_size_t, C.POINTER(vp)]
    L.fd_mgpu_create.argtypes = [C.POINTER(vp), ip, C.c_int32, C.c_int32]
    L.fd_mgpu_destroy.argtypes = [vp]
    L.fd_mgpu_destroy.restype = None
    L.fd_mgpu_fit.argtypes = [vp, pp, fp, C.c_int32, rp]
    L.fd_mgpu_solve.argtypes = [vp, fp, C.c_int32, C.c_int32, rp]
    L.fd_mgpu_eval.argtypes = [vp, fp, C.c_int64, fp, fp, fp, fp, fp, fp]
    L.fd_mgpu_info.argtypes = [vp, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int64), C.POINTER(C.c_float)]
    L.fd_mgpu_range.argtypes = [vp, C.c_int32, C.c_int64, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    L.fd_mgpu_ctx.argtypes = [vp, C.c_int32]
    L.fd_mgpu_ctx.restype = vp
    L.fd_mgpu_last_error.argtypes = [vp]
    L.fd_mgpu_last_error.restype = C.c_char_p
    L.fd_ctx_create.argtypes = [C.POINTER(vp), C.c_int, vp]
    L.fd_ctx_destroy.argtypes = [vp]
    L.fd_ctx_destroy.restype = None
    L.fd_ctx_synchronize.argtypes = [vp]
    L.fd_last_error.argtypes = [vp]
    L.fd_last_error.restype = C.c_char_p
    L.fd_rbf_fit.argtypes = [vp, pp, fp, C.c_int32, C.POINTER(vp), rp]
    L.fd_rbf_fit_dev.argtypes = [vp, pp, fp, C.c_int32, C.POINTER(vp)]
    L.fd_model_destroy.argtypes = [vp]
    L.fd_model_destroy.restype = None
    L.fd_rbf_solve.argtypes = [vp, fp, C.c_int32, C.c_int32, rp]
    L.fd_rbf_solve_dev.argtypes = [vp, fp, C.c_int32, C.c_int32]
    L.fd_model_report.argtypes = [vp, rp]
    L.fd_rbf_eval.argtypes = [vp, fp, C.c_int64, fp, fp, fp, fp, fp, fp]
    L.fd_rbf_eval_dev.argtypes = [vp, fp, C.c_int64, fp, fp, fp, fp, fp, fp]
    L.fd_model_create_receiver.argtypes = [vp, pp, fp, C.c_int32, C.c_int32, C.POINTER(vp)]
    L.fd_model_weights_dev.argtypes = [vp, C.POINTER(vp), C.POINTER(C.c_size_t)]
    L.fd_model_radii_dev.argtypes = [vp, C.POINTER(vp), C.POINTER(C.c_size_t)]
    L.fd_model_commit_weights.argtypes = [vp]
    L.fd_model_info.argtypes = [vp, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
    L.fd_model_get_weights.argtypes = [vp, fp, fp]
    L.fd_capture.argtypes = [vp, fp, C.c_int64, ip, ip, C.c_int32, fp, C.c_int32, ip, ip, C.c_int32, ip, C.c_int32,
                             C.c_float, C.c_int32, ip, vp, fp, C.POINTER(C.c_int32), ip, vp, ip, C.c_int32, C.c_int64]
    L.fd_ctx_phase_ms.argtypes = [vp, C.c_int]
    L.fd_ctx_phase_ms.restype = C.c_float
    L.fd_ctx_launch_count.argtypes = [vp]
    L.fd_ctx_launch_count.restype = C.c_int64
    L.fd_dbse_init.argtypes = [vp, fp, C.c_int64, fp, C.c_int32, C.POINTER(vp)]
    L.fd_dbse_compute_weights.argtypes = [vp, fp, fp, fp]
    L.fd_dbse_displace.argtypes = [vp, fp, fp, C.c_int32, fp, C.c_int32, C.c_float, fp]
    L.fd_dbse_get_weights.argtypes = [vp, fp]
    L.fd_dbse_get_qr.argtypes = [vp, fp, fp]
    L.fd_dbse_info.argtypes = [vp, C.POINTER(C.c_int64), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
    L.fd_dbse_destroy.argtypes = [vp]
    L.fd_dbse_destroy.restype = None
    L.fd_sop_set_blendshapes.argtypes = [vp, fp, C.c_int32, C.c_int64, C.c_int64]
    L.fd_sop_blend_weights.argtypes = [vp, fp, C.c_int32]
    L.fd_sop_create.argtypes = [C.c_int]
    L.fd_sop_create.restype = vp
    L.fd_sop_destroy.argtypes = [vp]
    L.fd_sop_destroy.restype = None
    L.fd_sop_params.argtypes = [vp]
    L.fd_sop_params.restype = pp
    L.fd_sop_cook.argtypes = [vp, fp, C.c_int64, ip, ip, C.c_int32, fp, fp, fp, C.c_int64, C.c_int64, fp, C.c_int32,
                              ip, ip, C.c_int32, ip, C.c_int64, C.c_int64, fp, C.c_int32, C.c_int32, fp, fp]
    L.fd_sop_messages.argtypes = [vp, C.c_int]
    L.fd_sop_messages.restype = C.c_char_p
    L.fd_sop_fit_count.argtypes = [vp]
    L.fd_sop_positions_bumped.argtypes = [vp]
    _lib = L
    return L
