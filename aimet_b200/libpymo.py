"""Drop-in for the in-scope part of the reference's pybind11 module ``aimet_common._libpymo``.

Mirrors (names, arity, argument meaning, error behaviour) the bindings in
ModelOptimizations/PyModelOptimizations/PyModelOptimizations.cpp:148-259 that sit on the quantization-simulation hot
path: ``TfEncoding``, the enums, ``TensorQuantizer`` (PyTensorQuantizer.cpp:48-77 + DlQuantization/src/TensorQuantizer.cpp),
``EncodingAnalyzerForPython`` (DlQuantization/src/EncodingAnalyzerForPython.cpp) and ``TensorQuantizationSimForPython``
(DlQuantization/src/TensorQuantizationSimForPython.cpp). The SVD / equalization / BN-fold bindings are out of scope.

numpy arrays are host data: they are staged onto the current CUDA device, processed by the sm_100a kernels, and copied
back. `useCuda` is accepted for signature compatibility; computation always runs on the GPU (there is no CPU path).
"""
import enum

import numpy as np
import torch

from . import _lib
from . import ops
from .state import StateArena


class ComputationMode(enum.IntEnum):
    COMP_MODE_CPU = 0
    COMP_MODE_GPU = 1


class QuantizationMode(enum.IntEnum):
    """DlQuantization/include/DlQuantization/Quantization.hpp:83-107"""
    QUANTIZATION_TF = 0
    QUANTIZATION_TF_ENHANCED = 1
    QUANTIZATION_RANGE_LEARNING = 2
    QUANTIZATION_PERCENTILE = 3
    QUANTIZATION_MSE = 4
    QUANTIZATION_ENTROPY = 5


class RoundingMode(enum.IntEnum):
    ROUND_NEAREST = 0
    ROUND_STOCHASTIC = 1


class TensorQuantizerOpMode(enum.IntEnum):
    """DlQuantization/include/DlQuantization/TensorQuantizerOpFacade.h"""
    updateStats = 0
    oneShotQuantizeDequantize = 1
    quantizeDequantize = 2
    passThrough = 3


# pybind's .export_values()
COMP_MODE_CPU, COMP_MODE_GPU = ComputationMode.COMP_MODE_CPU, ComputationMode.COMP_MODE_GPU
QUANTIZATION_TF = QuantizationMode.QUANTIZATION_TF
QUANTIZATION_TF_ENHANCED = QuantizationMode.QUANTIZATION_TF_ENHANCED
QUANTIZATION_RANGE_LEARNING = QuantizationMode.QUANTIZATION_RANGE_LEARNING
QUANTIZATION_PERCENTILE = QuantizationMode.QUANTIZATION_PERCENTILE
QUANTIZATION_MSE = QuantizationMode.QUANTIZATION_MSE
QUANTIZATION_ENTROPY = QuantizationMode.QUANTIZATION_ENTROPY
ROUND_NEAREST, ROUND_STOCHASTIC = RoundingMode.ROUND_NEAREST, RoundingMode.ROUND_STOCHASTIC


_ENCODING_EPOCH = [0]   # bumped by every write to any TfEncoding field: lets device-side caches of encodings stay valid


def encoding_epoch() -> int:
    return _ENCODING_EPOCH[0]


_FIELDS = ("min", "max", "delta", "offset", "bw")


class TfEncoding:
    """DlQuantization::TfEncoding (Quantization.hpp:113-120): read/write fields min, max, delta, offset, bw.

    An encoding handed out by a DEFERRED getEncoding (tensor_quantizer_op.DeferredAimetTensorQuantizer) has its five
    slots unset and `_lazy` pointing at the queue that owes the result: the first read of any field lands in
    `__getattr__`, which has the queue executed (one batched launch + one read-back for every encoding it owes) and the
    slots filled. An ordinary encoding never reaches `__getattr__`."""
    __slots__ = _FIELDS + ("_lazy",)

    def __init__(self):
        _set = object.__setattr__
        _set(self, "min", 0.0)
        _set(self, "max", 0.0)
        _set(self, "delta", 0.0)
        _set(self, "offset", 0.0)
        _set(self, "bw", 0)
        _set(self, "_lazy", None)
        _ENCODING_EPOCH[0] += 1

    def __getattr__(self, name):
        # only reached when the slot is unset
        if name in _FIELDS:
            try:
                lazy = object.__getattribute__(self, "_lazy")
            except AttributeError:
                lazy = None
            if lazy is not None:
                lazy.resolve()
                return object.__getattribute__(self, name)
        raise AttributeError(name)

    def __setattr__(self, name, value):
        if getattr(self, "_lazy", None) is not None:
            self._lazy.resolve()          # the other fields must exist before one of them is overwritten
        object.__setattr__(self, name, value)
        _ENCODING_EPOCH[0] += 1

    @classmethod
    def _from_values(cls, mn, mx, delta, offset, bw):
        e = cls.__new__(cls)
        _set = object.__setattr__
        _set(e, "min", float(mn))
        _set(e, "max", float(mx))
        _set(e, "delta", float(delta))
        _set(e, "offset", float(offset))
        _set(e, "bw", int(bw))
        _set(e, "_lazy", None)
        _ENCODING_EPOCH[0] += 1
        return e

    @classmethod
    def _deferred(cls, lazy):
        """An encoding whose values `lazy.resolve()` will fill in (see the class docstring)."""
        e = cls.__new__(cls)
        object.__setattr__(e, "_lazy", lazy)
        _ENCODING_EPOCH[0] += 1
        return e

    def _fill(self, row):
        _set = object.__setattr__
        _set(self, "min", row[0])
        _set(self, "max", row[1])
        _set(self, "delta", row[2])
        _set(self, "offset", row[3])
        _set(self, "bw", int(row[4]))
        _set(self, "_lazy", None)

    @classmethod
    def _from_c(cls, c):
        return cls._from_values(c.min, c.max, c.delta, c.offset, c.bw)

    def _to_c(self):
        return _lib.Encoding(float(self.min), float(self.max), float(self.delta), float(self.offset), int(self.bw))

    def __getstate__(self):
        return (self.min, self.max, self.delta, self.offset, self.bw)

    def __setstate__(self, s):
        object.__setattr__(self, "_lazy", None)
        for k, v in zip(_FIELDS, s):
            object.__setattr__(self, k, v)
        _ENCODING_EPOCH[0] += 1

    def __repr__(self):
        return (f"TfEncoding(min={self.min!r}, max={self.max!r}, delta={self.delta!r}, offset={self.offset!r}, "
                f"bw={self.bw!r})")


def scheme_code(quant_scheme) -> int:
    """Map a QuantizationMode to the kind of statistics the C ABI keeps for it (ab_quant_mode). tf, tf_enhanced and
    percentile are on the hot path; percentile keeps the tf_enhanced statistics (PercentileEncodingAnalyzer.cpp:69-75)
    and differs only in how the encoding is closed (see is_percentile). The reference's factory
    (QuantizerFactory.cpp:72-103) falls back to the TF analyzer for anything it does not know, range learning included,
    and so do we. The entropy analyzer keeps a different, rescaling histogram behind its own entry points (ab_entropy_*)."""
    mode = QuantizationMode(int(quant_scheme))
    if mode in (QuantizationMode.QUANTIZATION_TF_ENHANCED, QuantizationMode.QUANTIZATION_PERCENTILE):
        return ops.QUANTIZATION_TF_ENHANCED
    if mode == QuantizationMode.QUANTIZATION_MSE:
        return ops.QUANTIZATION_MSE        # tf_enhanced statistics, MseEncodingAnalyzer's closing search
    if mode in (QuantizationMode.QUANTIZATION_TF, QuantizationMode.QUANTIZATION_RANGE_LEARNING):
        return ops.QUANTIZATION_TF
    if mode == QuantizationMode.QUANTIZATION_ENTROPY:
        return ops.QUANTIZATION_ENTROPY
    raise NotImplementedError(f"{mode.name} is not a statistics scheme")


def is_percentile(quant_scheme) -> bool:
    return QuantizationMode(int(quant_scheme)) == QuantizationMode.QUANTIZATION_PERCENTILE


def _default_device() -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError("aimet_b200 needs a CUDA device: there is no CPU path")
    return torch.device("cuda", torch.cuda.current_device())


def _stage(array) -> torch.Tensor:
    """numpy (or any array-like) -> contiguous fp32 tensor on the current CUDA device"""
    a = np.ascontiguousarray(array, dtype=np.float32)
    return torch.from_numpy(a).to(_default_device())


class _Analyzer:
    """Device-resident statistics of one quantizer + the host flags the reference keeps next to it."""

    def __init__(self, quant_scheme):
        self._scheme = QuantizationMode(int(quant_scheme))
        self._code = scheme_code(self._scheme)
        self._slot = None
        self.percentile = 100.0   # PercentileEncodingAnalyzer.h:100

    def reset(self):
        if self._slot is not None:
            self._slot.reset()

    def update(self, tensor: torch.Tensor):
        if self._slot is None or self._slot.device != tensor.device:
            self._slot = StateArena.for_device(tensor.device).allocate(1)
        if self._code == ops.QUANTIZATION_ENTROPY:
            ops.entropy_update_impl(tensor, self._slot.arena, self._slot.first)
            return
        ops.stats_update_impl(tensor, self._slot.arena, self._slot.first, self._code, None, 0)

    def compute(self, bw, sym, strict, unsigned_sym) -> TfEncoding:
        if self._slot is None:
            return TfEncoding()   # no statistics at all: the zero encoding
        if self._code == ops.QUANTIZATION_ENTROPY:
            return TfEncoding._from_c(ops.entropy_compute_impl(self._slot.arena, self._slot.first, bw, sym, strict,
                                                               unsigned_sym))
        enc, _ = ops.compute_encodings_impl(self._slot.arena, self._slot.first, 1, self._code, bw, sym, strict,
                                            unsigned_sym,
                                            percentile=self.percentile if is_percentile(self._scheme) else None)
        v = enc[0].tolist()
        return TfEncoding._from_values(v[0], v[1], v[2], v[3], int(v[4]))

    def histogram(self):
        if self._code == ops.QUANTIZATION_ENTROPY:
            # the reference's EntropyEncodingAnalyzer::getStatsHistogram trips its own size assertion (:54-78)
            raise AssertionError("pdf.xLeft.size() == pdf.pdf.size()")
        if not ops.keeps_histogram(self._code):
            raise AssertionError("No real histogram data is kept for TF Encoding analyzer")   # TfEncodingAnalyzer.cpp:53-57
        if self._slot is None:
            return []
        return self._slot.histogram(0)


class EncodingAnalyzerForPython:
    """DlQuantization/src/EncodingAnalyzerForPython.cpp:47-92"""

    def __init__(self, quantization_scheme):
        self._analyzer = _Analyzer(quantization_scheme)
        self._valid = False

    def updateStats(self, input, use_cuda):   # pylint: disable=invalid-name,redefined-builtin
        self._valid = True
        self._analyzer.update(_stage(input))

    def computeEncoding(self, bitwidth, use_symmetric_encodings, use_strict_symmetric, use_unsigned_symmetric):   # pylint: disable=invalid-name
        if not self._valid:
            return TfEncoding(), False
        return self._analyzer.compute(bitwidth, use_symmetric_encodings, use_strict_symmetric,
                                      use_unsigned_symmetric), True


class TensorQuantizationSimForPython:
    """DlQuantization/src/TensorQuantizationSimForPython.cpp:47-79"""

    def quantizeDequantize(self, input, encoding, rounding_mode, *args):   # pylint: disable=invalid-name,redefined-builtin
        # overloads: (input, encoding, roundingMode, bitwidth, use_cuda) and (input, encoding, roundingMode, use_cuda)
        if len(args) == 2:
            bitwidth = int(args[0])
        elif len(args) == 1:
            bitwidth = int(encoding.bw)
        else:
            raise TypeError("quantizeDequantize(input, encoding, roundingMode, [bitwidth,] use_cuda)")
        arr = np.asarray(input)
        out = ops.qdq_per_tensor_impl(_stage(arr), encoding.min, encoding.max, bitwidth, int(rounding_mode), 0)
        return out.cpu().numpy().reshape(arr.shape)


class TensorQuantizer:
    """libpymo.TensorQuantizer == DlQuantization::PyTensorQuantizer (PyTensorQuantizer.cpp:48-77) over
    DlQuantization::TensorQuantizer (DlQuantization/src/TensorQuantizer.cpp:49-343)."""

    def __init__(self, quant_scheme, rounding_mode):
        self._quant_scheme = QuantizationMode(int(quant_scheme))
        self.roundingMode = RoundingMode(int(rounding_mode))
        self.isEncodingValid = False
        self._strict = False
        self._unsigned = False
        self._valid_stats = False
        self._analyzer = _Analyzer(self._quant_scheme)
        self._percentile = 100.0

    # -- flags: every setter resets the statistics (TensorQuantizer.cpp:66-103) ------------------------------------
    def getQuantScheme(self):
        return self._quant_scheme

    def setQuantScheme(self, quant_scheme):
        self._quant_scheme = QuantizationMode(int(quant_scheme))
        self.resetEncodingStats()

    def getStrictSymmetric(self):
        return self._strict

    def setStrictSymmetric(self, use_strict_symmetric):
        self._strict = bool(use_strict_symmetric)
        self.resetEncodingStats()

    def getUnsignedSymmetric(self):
        return self._unsigned

    def setUnsignedSymmetric(self, use_unsigned_symmetric):
        self._unsigned = bool(use_unsigned_symmetric)
        self.resetEncodingStats()

    def resetEncodingStats(self):
        self._valid_stats = False
        self.isEncodingValid = False
        self._analyzer = _Analyzer(self._quant_scheme)   # a new analyzer: the percentile is back at its default too

    # -- statistics / encodings ---------------------------------------------------------------------------------
    def updateStats(self, tensor, use_cuda):
        self._valid_stats = True
        self._analyzer.update(_stage(tensor))

    def computeEncoding(self, bitwidth, use_symmetric_encoding):
        if not self._valid_stats:
            return TfEncoding()
        enc = self._analyzer.compute(bitwidth, use_symmetric_encoding, self._strict, self._unsigned)
        self.isEncodingValid = True
        return enc

    def quantizeDequantize(self, input_tensor, output_tensor, encoding_min, encoding_max, bitwidth, use_cuda):
        assert self.isEncodingValid   # the reference's assert is live (no NDEBUG): TensorQuantizer.cpp:172
        arr = np.asarray(input_tensor)
        out = ops.qdq_per_tensor_impl(_stage(arr), encoding_min, encoding_max, bitwidth, int(self.roundingMode), 0)
        np.copyto(output_tensor, out.cpu().numpy().reshape(arr.shape))

    def getStatsHistogram(self):
        return self._analyzer.histogram()

    def setPercentileValue(self, percentile):
        # only meaningful for the percentile scheme (TensorQuantizer.cpp:239-246); lives in the analyzer, as there
        if self._quant_scheme == QuantizationMode.QUANTIZATION_PERCENTILE:
            self._analyzer.percentile = float(percentile)

    def getPercentileValue(self):
        if self._quant_scheme == QuantizationMode.QUANTIZATION_PERCENTILE:
            return self._analyzer.percentile
        raise RuntimeError("Percentile Value only exists in case of percentile quant scheme.")

    def computePartialEncoding(self, bw, encoding, use_symmetric_encodings, use_unsigned_symmetric,
                               use_strict_symmetric):
        """In-place on `encoding`, like the reference (it takes TfEncoding&): TensorQuantizer.cpp:327-343."""
        c = encoding._to_c()
        import ctypes
        rc = _lib.load().ab_compute_partial_encoding(int(bw), ctypes.byref(c), int(bool(use_symmetric_encodings)),
                                                     int(bool(use_unsigned_symmetric)),
                                                     int(bool(use_strict_symmetric)))
        if rc != _lib.AB_OK:
            msg = _lib.load().ab_last_error().decode()
            if "Cannot determine" in msg:
                raise RuntimeError(msg)      # std::runtime_error
            raise ValueError(msg)            # std::invalid_argument
        encoding.min, encoding.max, encoding.delta, encoding.offset = c.min, c.max, c.delta, c.offset
        encoding.bw = c.bw


def PtrToInt64(ptr):   # pylint: disable=invalid-name
    return int(ptr)
