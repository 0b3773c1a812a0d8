"""ctypes binding of the C ABI declared in include/aimet_b200.h.

The shared library is built in-tree by aimet_b200/_build.py (nvcc, sm_100a). There is no fallback of any kind: if the
library is missing or cannot be loaded, importing this module raises, and every device entry point raises when CUDA
reports an error (including "no device").
"""
import ctypes as C
import os

from . import _build

_LIB = None


class AbError(RuntimeError):
    """A C-ABI call returned a negative ab_status."""


class Encoding(C.Structure):
    """ab_encoding == DlQuantization::TfEncoding (reference Quantization.hpp:113-120)."""
    _fields_ = [("min", C.c_double), ("max", C.c_double), ("delta", C.c_double), ("offset", C.c_double),
                ("bw", C.c_int)]


class StatsSegment(C.Structure):
    """ab_stats_segment: one (tensor, record) pair of a multi-tensor statistics call."""
    _fields_ = [("data", C.c_void_p), ("count", C.c_int64), ("state_index", C.c_int32), ("reserved", C.c_int32)]


class QcQuantizeInfo(C.Structure):
    """ab_qc_quantize_info: what the reference's ONNX op keeps per quantizer (QcQuantizeInfo.h:47-73)."""
    _fields_ = [("states", C.c_void_p), ("encodings", C.POINTER(Encoding)), ("num_encodings", C.c_int), ("op_mode", C.c_int),
                ("quant_mode", C.c_int), ("use_symmetric_encoding", C.c_int), ("enabled", C.c_int),
                ("is_int_data_type", C.c_int), ("use_per_channel_mode", C.c_int), ("channel_axis", C.c_int),
                ("block_axis", C.c_int), ("block_size", C.c_int)]


OP_UPDATE_STATS, OP_ONE_SHOT_QDQ, OP_QDQ, OP_PASS_THROUGH = 0, 1, 2, 3


class RefreshItem(C.Structure):
    """ab_refresh_item: one parameter tensor of a multi-tensor encoding refresh."""
    _fields_ = [("data", C.c_void_p), ("num_segments", C.c_int64), ("segment_len", C.c_int64), ("first_record", C.c_int64)]


PDF_SIZE = 512
REFRESH_MULTI_MAX_ITEMS = 96
STATS_MULTI_MAX_SEGMENTS = 128
STATS_MULTI_LOG_ONLY = 1
AB_OK, AB_ERR_INVALID, AB_ERR_CUDA, AB_ERR_UNSUPPORTED = 0, -1, -2, -3
AB_F32, AB_BF16 = 0, 1

_vp, _i64, _u64, _int, _dbl, _flt = C.c_void_p, C.c_int64, C.c_uint64, C.c_int, C.c_double, C.c_float
_encp, _dblp, _fltp = C.POINTER(Encoding), C.POINTER(C.c_double), C.POINTER(C.c_float)

# name -> (restype, argtypes); every symbol include/aimet_b200.h declares
PROTOTYPES = {
    "ab_last_error": (C.c_char_p, []),
    "ab_version": (_int, []),
    "ab_device_count": (_int, []),
    "ab_stats_state_bytes": (C.c_size_t, []),
    "ab_gate_min_max": (_int, [_dblp, _dblp]),
    "ab_fill_encoding_info": (_int, [_int, _dbl, _dbl, _encp]),
    "ab_tf_compute_encoding": (_int, [_int, _dbl, _dbl, _int, _int, _int, _encp]),
    "ab_tf_analyzer_encoding": (_int, [_int, _dbl, _dbl, _int, _int, _int, _encp]),
    "ab_compute_partial_encoding": (_int, [_int, _encp, _int, _int, _int]),
    "ab_per_channel_params": (_int, [_dblp, _dblp, _int, _int, _fltp]),
    "ab_qdq_per_tensor_fwd": (_int, [_vp, _vp, _i64, _int, _dbl, _dbl, _int, _int, _u64, _vp]),
    "ab_qdq_per_tensor_fwd_dev": (_int, [_vp, _vp, _i64, _int, _vp, _int, _u64, _vp]),
    "ab_quantize_to_grid": (_int, [_vp, _vp, _i64, _int, _dbl, _dbl, _int, _int, _int, _u64, _vp]),
    "ab_quantize_to_packed": (_int, [_vp, _vp, _i64, _int, _dbl, _dbl, _int, _int, _vp]),
    "ab_qdq_per_channel_fwd": (_int, [_vp, _vp, _i64, _i64, _i64, _int, _vp, _int, _u64, _vp]),
    "ab_per_channel_params_dev": (_int, [_vp, _i64, _int, _vp, _vp]),
    "ab_qdq_broadcast_fwd": (_int, [_vp, _vp, _i64, _int, C.POINTER(C.c_int64), C.POINTER(C.c_int64), _vp, _vp, _vp, _vp,
                                    _int, _vp]),
    "ab_qdq_ste_bwd": (_int, [_vp, _vp, _vp, _i64, _int, _flt, _flt, _vp]),
    "ab_qdq_ste_bwd_per_channel": (_int, [_vp, _vp, _vp, _i64, _i64, _i64, _int, _vp, _vp, _vp]),
    "ab_qdq_ste_bwd_enc5": (_int, [_vp, _vp, _vp, _i64, _i64, _i64, _int, _vp, _int, _vp]),
    "ab_stats_reset": (_int, [_vp, _i64, _vp]),
    "ab_stats_update": (_int, [_vp, _i64, _int, _int, _vp, _vp, _int, _vp]),
    "ab_stats_update_segmented": (_int, [_vp, _i64, _i64, _int, _int, _vp, _vp]),
    "ab_stats_update_multi": (_int, [C.POINTER(StatsSegment), _int, _int, _vp, _vp, _int, _vp]),
    "ab_compute_encodings": (_int, [_vp, _i64, _int, _int, _int, _int, _int, _vp, _vp, _vp]),
    "ab_compute_encodings_percentile": (_int, [_vp, _i64, _flt, _int, _int, _int, _int, _vp, _vp, _vp]),
    "ab_debug_hist_timer": (_i64, [_vp, _i64]),
    "ab_stats_refresh_encodings": (_int, [_vp, _i64, _i64, _int, _int, _vp, _int, _int, _int, _int, _vp, _vp, _vp, _vp]),
    "ab_stats_refresh_encodings_multi": (_int, [C.POINTER(RefreshItem), _int, _int, _int, _vp, _int, _int, _int, _int, _vp, _vp,
                                                _vp, _vp]),
    "ab_stats_init_range": (_int, [_vp, _i64, _vp, _vp]),
    "ab_stats_fold_batches": (_int, [_vp, _i64, _vp, _vp, _i64, _vp]),
    "ab_stats_fold_log": (_int, [_vp, _i64, _vp, _vp, _vp, _vp]),
    "ab_qc_quantize_op_workspace_bytes": (C.c_size_t, [_int]),
    "ab_qc_quantize_op_compute": (_int, [C.POINTER(QcQuantizeInfo), _vp, _vp, C.POINTER(C.c_int64), _int, _int, _vp, _vp]),
    "ab_entropy_update": (_int, [_vp, _i64, _int, _vp, _vp]),
    "ab_entropy_compute_encoding": (_int, [_vp, _int, _int, _int, _int, _encp, _vp]),
    "ab_entropy_histogram": (_int, [_vp, _dblp, _dblp, C.POINTER(C.c_int), _vp]),
    "ab_lg_workspace_bytes": (_i64, [_i64]),
    "ab_lg_qdq_fwd": (_int, [_vp, _vp, _i64, _i64, _i64, _int, _vp, _vp, _int, _int, _int, _int, _vp, _vp]),
    "ab_lg_qdq_bwd": (_int, [_vp, _vp, _vp, _i64, _i64, _i64, _int, _vp, _vp, _int, _int, _int, _vp, _vp, _vp, _vp]),
}


def lib_path():
    return _build.LIB_PATH


def load():
    """dlopen libaimet_b200.so (once). Raises if it has not been built: there is no other implementation."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = lib_path()
    if not os.path.exists(path):
        raise ImportError(
            f"{path} not found. Build it with `python -m aimet_b200._build` (needs nvcc); aimet_b200 has no "
            "CPU or pure-PyTorch fallback.")
    lib = C.CDLL(path)
    for name, (restype, argtypes) in PROTOTYPES.items():
        fn = getattr(lib, name)          # AttributeError here == the library is stale / incomplete
        fn.restype = restype
        fn.argtypes = argtypes
    _LIB = lib
    return lib


def check(rc):
    if rc != AB_OK:
        msg = load().ab_last_error().decode("utf-8", "replace")
        if rc == AB_ERR_INVALID:
            raise ValueError(msg)
        raise AbError(f"aimet_b200 C-ABI call failed (status {rc}): {msg}")


def call(name, *args):
    check(getattr(load(), name)(*args))
