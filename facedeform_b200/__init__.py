"""facedeform_b200 -- B200-native (sm_100a) RBF deformation path of symek/facedeform behind a C ABI.

Only what the hot path needs lives here: csrc/ (CUDA kernels + the C ABI, built in-tree into
libfacedeform_gpu.so), the ctypes mirror of that ABI (api.py), vertex sharding helpers (shard.py) and the
seeded synthetic inputs (synth.py).  There is no CPU fallback.
"""
from .api import (Context, RbfModel, DirectBSEdit, MultiGpu, FdError, make_params,  # noqa: F401
                  MGPU_AUTO, MGPU_NCCL, MGPU_P2P,
                  MODEL_QNN, MODEL_ML, TERM_LINEAR, TERM_CONST, TERM_ZERO,
                  KERNEL_GAUSSIAN, KERNEL_MULTIQUADRIC, KERNEL_THINPLATE,
                  EVAL_AUTO, EVAL_FP32, EVAL_FP64, PATH_AUTO, PATH_SIMT, PATH_TENSOR)
