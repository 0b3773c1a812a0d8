"""aimet_b200 -- a B200-native (sm_100a) implementation of AIMET's quantization-simulation hot path.

Layout
  csrc/                  hand-written CUDA kernels + the C ABI (include/aimet_b200.h) -> lib/libaimet_b200.so
  _lib.py, ops.py        ctypes binding and torch custom ops (namespace ``aimet_b200``) over that ABI
  libpymo.py             drop-in for the in-scope part of ``aimet_common._libpymo``
  tensor_quantizer_op.py drop-in for ``aimet_common.AimetTensorQuantizer``
  install.py             registers both under the reference's module names
  quantsim/              host mirror of the reference's Python layers on this path (quantizers, wrappers,
                         QuantizationSimModel.compute_encodings, encodings export), batched for the GPU
  distributed.py         calibration sharded by batch across GPUs with an exact NCCL merge

Importing this package loads the CUDA library; if it has not been built the import fails (there is no fallback).
"""
import sys as _sys

from . import _lib

# `python -m aimet_b200._build` is how the library gets built in the first place: only then the package may be imported
# without it (nothing but the build script runs).
_BUILDING = "aimet_b200._build" in getattr(_sys, "orig_argv", [])[1:3]
if not _BUILDING:
    _lib.load()

    from . import ops  # noqa: E402,F401  (registers torch.ops.aimet_b200.*)
    from . import libpymo  # noqa: E402,F401
    from .tensor_quantizer_op import AimetTensorQuantizer  # noqa: E402,F401

__version__ = "0.1.0"
