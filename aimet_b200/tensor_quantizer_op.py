"""Drop-in for the reference's torch extension module ``aimet_common.AimetTensorQuantizer``.

Same class name, same nine methods, same argument meaning as the pybind11 class in
TrainingExtensions/torch/src/AimetTensorQuantizer.cpp:75-331 -- what aimet_torch/v1/tensor_quantizer.py calls as
``self._cppOp[i].updateStats / getEncoding / quantizeDequantize / quantizeDequantizePerChannel / quantize``.

Differences a caller can observe, all deliberate:
  * computation always runs on a CUDA device through the sm_100a kernels. A CPU tensor (the reference's
    `use_cuda=False` path, e.g. the two-element tensor built for encoding_min_max_fixed_vals) is staged to the current
    device and the result copied back; without a CUDA device every method raises. There is no CPU fallback.
  * bfloat16 tensors are accepted natively (the reference reads `input.data<float>()` only).
  * `updateStats` never synchronises: statistics are device resident; the only host round trip is in `getEncoding`
    / `getStatsHistogram`.
"""
import torch

from . import libpymo
from . import ops
from .state import StateArena, field_index


def _to_device_tensor(t: torch.Tensor):
    """-> (fp32/bf16 CUDA tensor, original device)"""
    orig = t.device
    if not t.is_cuda:
        if not torch.cuda.is_available():
            raise RuntimeError("aimet_b200 needs a CUDA device: there is no CPU path")
        t = t.to(torch.device("cuda", torch.cuda.current_device()))
    if t.dtype not in (torch.float32, torch.bfloat16):
        t = t.to(torch.float32)   # the reference upcasts half tensors before the native call (tensor_quantizer.py:477)
    return t, orig


class ValidityGroup:
    """`isEncodingValid` shared by the ops of one per-channel quantizer: the host layer updates and resets all channels
    together (one launch), so it flips ONE flag instead of looping over up to thousands of ops. An op that is written
    individually leaves the group (`detached` counts those), and the host layer falls back to looking at every op."""
    __slots__ = ("valid", "detached")

    def __init__(self):
        self.valid = False
        self.detached = 0


class AimetTensorQuantizer:
    """One encoding analyzer + one quantize-dequantize simulator (AimetTensorQuantizer.cpp:75-83)."""

    @property
    def _is_encoding_valid(self):
        group = self._group
        return group.valid if group is not None else self._valid

    @_is_encoding_valid.setter
    def _is_encoding_valid(self, value):
        if self._group is not None:
            self._group.detached += 1
            self._group = None
        self._valid = bool(value)

    def __init__(self, quantization_scheme):
        self._group = None
        self._valid = False
        self._scheme = libpymo.QuantizationMode(int(quantization_scheme))
        self._code = libpymo.scheme_code(self._scheme)
        self._percentile = 100.0 if libpymo.is_percentile(self._scheme) else None   # None: not the percentile scheme
        self._is_encoding_valid = False
        self._block = None            # StateBlock with one record, allocated on first use
        self._index = 0
        self._pc_cache = None         # (key, device tensor) for the per-channel parameter block
        self._range_fixed = False     # host knowledge that the device record's histogram range is fixed
        self._probe = None            # (pinned int32 tensor, event) of an in-flight read-back of `initialized`
        self._updates = 0             # updateStats calls since the last reset / bind

    # ---- wiring used by the batched host layer (aimet_b200.quantsim): share one contiguous block per weight -------
    def _bind(self, block, index, group=None):
        self._block, self._index = block, index
        self._range_fixed, self._probe, self._updates = False, None, 0
        if group is not None:
            self._group = group

    def _ensure_state(self, device):
        if self._block is None or self._block.device != device:
            self._block, self._index = StateArena.for_device(device).allocate(1), 0

    # ---- reference API ---------------------------------------------------------------------------------------
    def _reset_host_state(self):
        """The host half of resetEncodingStats; the caller has reset (or is about to reset) the record itself, together with
        its neighbours, in one launch."""
        self._is_encoding_valid = False
        if self._percentile is not None:
            self._percentile = 100.0
        self._range_fixed, self._probe, self._updates = False, None, 0

    def resetEncodingStats(self):
        """AimetTensorQuantizer.cpp:85-92 (a new analyzer: the percentile value falls back to its default as well)"""
        self._reset_host_state()
        if self._block is not None:
            ops.stats_reset_impl(self._block.arena, self._block.first + self._index, 1)

    def updateStats(self, input, use_cuda):   # pylint: disable=redefined-builtin
        """AimetTensorQuantizer.cpp:94-126"""
        t, _ = _to_device_tensor(input)
        self._is_encoding_valid = True
        self._ensure_state(t.device)
        if self._code == ops.QUANTIZATION_ENTROPY:
            ops.entropy_update_impl(t, self._block.arena, self._block.first + self._index)
            return
        ops.stats_update_impl(t, self._block.arena, self._block.first + self._index, self._code, None, 0,
                              ops.STATS_RANGE_FIXED if self._range_fixed else 0)
        self._updates += 1
        # The probe only pays off when the same record is updated again and again (calibration). A record that is reset
        # before every update -- a weight quantizer in training mode -- never gets there, and a probe costs a pinned
        # buffer, a copy and an event (2.4 ms per MobileNet-v2 QAT step): start with the second update after a reset.
        if ops.keeps_histogram(self._code) and not self._range_fixed and self._updates >= 2:
            self._poll_range_fixed()

    _INITIALIZED_WORD = field_index("initialized", 4)      # ab_stats_state.initialized as an int32 index

    def _poll_range_fixed(self):
        """Learn, without ever blocking, that the record's range got fixed: a 4-byte asynchronous read-back of the
        `initialized` flag into pinned memory, looked at on a later call once its event has completed."""
        if self._probe is not None:
            flag, event = self._probe
            if not event.query():
                return
            self._probe = None
            if int(flag[0]) == 1:
                self._range_fixed = True
                return
        if torch.cuda.is_current_stream_capturing():
            return
        rec = self._block.arena[(self._block.first + self._index) * ops.STATE_BYTES:
                                (self._block.first + self._index + 1) * ops.STATE_BYTES].view(torch.int32)
        flag = torch.empty(1, dtype=torch.int32, pin_memory=True)
        flag.copy_(rec[self._INITIALIZED_WORD:self._INITIALIZED_WORD + 1], non_blocking=True)
        event = torch.cuda.Event()
        event.record(torch.cuda.current_stream(self._block.device))
        self._probe = (flag, event)

    def getEncoding(self, bitwidth, use_symmetric_encodings, use_strict_symmetric, use_unsigned_symmetric):
        """AimetTensorQuantizer.cpp:180-192 -> (TfEncoding, is_valid)"""
        if not self._is_encoding_valid or self._block is None:
            return libpymo.TfEncoding(), self._is_encoding_valid
        if self._code == ops.QUANTIZATION_ENTROPY:
            return libpymo.TfEncoding._from_c(ops.entropy_compute_impl(
                self._block.arena, self._block.first + self._index, bitwidth, use_symmetric_encodings, use_strict_symmetric,
                use_unsigned_symmetric)), True
        if use_symmetric_encodings and self._code == ops.QUANTIZATION_TF:
            assert not (use_strict_symmetric and use_unsigned_symmetric)   # TfEncodingAnalyzer.cpp:85-86
        enc, _ = ops.compute_encodings_impl(self._block.arena, self._block.first + self._index, 1, self._code,
                                            bitwidth, use_symmetric_encodings, use_strict_symmetric,
                                            use_unsigned_symmetric, percentile=self._percentile)
        v = enc[0].tolist()
        return libpymo.TfEncoding._from_values(v[0], v[1], v[2], v[3], int(v[4])), True

    def quantizeDequantize(self, input, encoding, rounding_mode, use_cuda):   # pylint: disable=redefined-builtin
        """AimetTensorQuantizer.cpp:129-153: uses only encoding.min / max / bw."""
        t, orig = _to_device_tensor(input)
        fn = torch.ops.aimet_b200.qdq_per_tensor if (t.requires_grad and torch.is_grad_enabled()) else \
            ops.qdq_per_tensor_impl
        out = fn(t, float(encoding.min), float(encoding.max), int(encoding.bw), int(rounding_mode), _next_seed())
        return _restore(out, input, orig)

    def quantize(self, input, encoding, rounding_mode, use_cuda, shift_to_signed):   # pylint: disable=redefined-builtin
        """AimetTensorQuantizer.cpp:155-178"""
        t, orig = _to_device_tensor(input)
        out = torch.ops.aimet_b200.quantize_to_grid(t, float(encoding.min), float(encoding.max), int(encoding.bw),
                                                    int(rounding_mode), bool(shift_to_signed), _next_seed())
        return _restore(out, input, orig)

    def quantizeDequantizePerChannel(self, input, encodings, num_channel, num_element, num_element_per_channel,   # pylint: disable=redefined-builtin
                                     rounding_mode, use_cuda):
        """AimetTensorQuantizer.cpp:256-307"""
        t, orig = _to_device_tensor(input)
        params = self._per_channel_params(encodings, t.device)
        fn = torch.ops.aimet_b200.qdq_per_channel if (t.requires_grad and torch.is_grad_enabled()) else \
            ops.qdq_per_channel_impl
        out = fn(t, params, int(num_channel), int(num_element_per_channel), int(rounding_mode), _next_seed())
        return _restore(out, input, orig)

    def makeDeltaOffsetTensor(self, device, encodings):
        """AimetTensorQuantizer.cpp:209-234 -> (delta, offset) fp32 tensors on `device`"""
        n = len(encodings)
        host = torch.tensor([[e.delta for e in encodings], [e.offset for e in encodings]], dtype=torch.float32)
        host = host.reshape(2, n).to(device)
        return host[0], host[1]

    def getStatsHistogram(self):
        """AimetTensorQuantizer.cpp:194-198"""
        if not ops.keeps_histogram(self._code):
            raise AssertionError("No real histogram data is kept for TF Encoding analyzer")
        if self._block is None:
            return []
        return self._block.histogram(self._index)

    def setPercentileValue(self, percentile):
        """AimetTensorQuantizer.cpp:200-207: a no-op unless the scheme is percentile."""
        if self._percentile is not None:
            self._percentile = float(percentile)

    # ---- helpers ---------------------------------------------------------------------------------------------
    def _per_channel_params(self, encodings, device):
        # valid while the very same list object is passed and no TfEncoding anywhere has been written since
        key = (device, id(encodings), len(encodings), libpymo.encoding_epoch())
        if self._pc_cache is not None and self._pc_cache[0] == key and self._pc_cache[2] is encodings:
            return self._pc_cache[1]
        if torch.cuda.is_current_stream_capturing():
            raise RuntimeError("per-channel parameters changed while a CUDA graph was being captured")
        host = ops.per_channel_params([e.min for e in encodings], [e.max for e in encodings], encodings[0].bw)
        dev = host.pin_memory().to(device, non_blocking=True) if torch.cuda.is_available() else host.to(device)
        self._pc_cache = (key, dev, encodings)
        return dev


_SEED = [0]


def _next_seed():
    _SEED[0] = (_SEED[0] + 0x9E3779B97F4A7C15) & (2**63 - 1)
    return _SEED[0]


def _restore(out, original, orig_device):
    """Give the result the dtype/device the reference would: same dtype as the (fp32) input, same device."""
    if out.dtype != original.dtype and original.dtype in (torch.float16, torch.float64):
        out = out.to(original.dtype)
    if out.device != orig_device:
        out = out.to(orig_device)
    return out
