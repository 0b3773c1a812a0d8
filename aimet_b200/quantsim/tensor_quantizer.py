"""Host mirror of the reference's static-grid tensor quantizers and their autograd functions.

Reference: TrainingExtensions/torch/src/python/aimet_torch/v1/tensor_quantizer.py
  StaticGridTensorQuantizer :143-421, StaticGridPerTensorQuantizer :408-480, StaticGridPerChannelQuantizer :483-570,
  QuantizeDequantize :1098-1212, Quantize :1215-1282.
Same attribute and method names, same state machine. What changes is underneath `_cppOp`:
  * the native op is aimet_b200.AimetTensorQuantizer (sm_100a kernels, device-resident statistics);
  * a per-channel quantizer owns ONE contiguous block of statistics records, so `update_encoding_stats` is one segmented
    launch and `compute_encoding` one batched grid search + one device->host copy, instead of the reference's Python loop
    of num_channels native calls each (:567-570, :296-299);
  * bfloat16 tensors go to the kernels as they are (the reference only knows how to upcast float16).
The product has ONE native op class (AimetTensorQuantizer over the sm_100a library); `_set_op_class_for_testing` is a
private hook with which tests/ and bench.py's CPU-baseline leg drive this same host layer with the CPU oracle -- it is
not a backend-dispatch mechanism and nothing in aimet_b200/ calls it.
"""
import functools
import math
import os
import struct
from typing import List, Optional, Tuple, Union

import torch

from .. import libpymo, ops
from ..state import StateArena
from ..tensor_quantizer_op import AimetTensorQuantizer, ValidityGroup
from .defs import MAP_QUANT_SCHEME_TO_PYMO, QuantizationDataType, QuantScheme

_DEFAULT_OP_FACTORY = AimetTensorQuantizer
# `_encoding is _LAZY`: the encodings were computed on the device and still live only there (`_enc_dev`, a [n, 5] float64
# tensor, plus the ready-made kernel parameters). The hot paths consume those directly; TfEncoding objects are built --
# one device->host copy -- the first time anybody asks for `.encoding`. For per-channel ResNet-50 this takes 53 host
# synchronisations and 26 560 Python objects out of every calibration job.
_LAZY = object()


FUSED_REFRESH = os.environ.get("AB_FUSED_REFRESH", "1") != "0"   # A/B switch for refresh_encoding_from


def _set_op_class_for_testing(factory):
    """PRIVATE, test infrastructure only: replace the class used for `_cppOp` objects (tests/ and bench.py's CPU-baseline
    leg put the CPU oracle underneath this host layer to check it). Returns the previous one."""
    global _DEFAULT_OP_FACTORY
    prev, _DEFAULT_OP_FACTORY = _DEFAULT_OP_FACTORY, factory
    return prev


def _is_native(op) -> bool:
    return isinstance(op, AimetTensorQuantizer)


class StaticGridTensorQuantizer:
    """Simulates quantization for a tensor with a grid fixed by calibration (reference :143-421)."""

    def __init__(self, bitwidth: int, round_mode, quant_scheme: QuantScheme, use_symmetric_encodings: bool,
                 enabled_by_default: bool, data_type: QuantizationDataType = QuantizationDataType.int):
        if data_type != QuantizationDataType.int:
            raise NotImplementedError("float (fp16 / fp8) simulation is outside the aimet_b200 hot path")
        self.round_mode = round_mode
        self._quant_scheme = quant_scheme
        self.use_symmetric_encodings = use_symmetric_encodings
        self.use_strict_symmetric = False
        self.use_unsigned_symmetric = False
        self.is_unsigned_symmetric = False
        self.bitwidth = bitwidth
        self.enabled = enabled_by_default
        self.data_type = data_type
        self.is_const = False
        self._encoding_min_max_fixed_vals = None
        self._is_encoding_frozen = False
        self._cppOp = None
        self._encoding = None
        self._enc_dev = self._qdq4_dev = self._params_dev = None
        self._lazy_ok = False         # parameter quantizers switch this on (their encodings are consumed on the device)
        self._stats_dirty = True      # statistics changed since the encoding was last computed
        self._op_factory = _DEFAULT_OP_FACTORY

    # ---- properties ----------------------------------------------------------------------------------------------
    @property
    def quant_scheme(self) -> QuantScheme:
        return self._quant_scheme

    @quant_scheme.setter
    def quant_scheme(self, quant_scheme: QuantScheme):
        # changing the scheme re-creates the native objects, which also clears their statistics (reference :236-249)
        self._quant_scheme = quant_scheme
        assert self._cppOp
        self._cppOp = [self._op_factory(MAP_QUANT_SCHEME_TO_PYMO[quant_scheme]) for _ in self._cppOp]
        self._block = None

    @property
    def is_encoding_frozen(self) -> bool:
        return self._is_encoding_frozen

    @property
    def channel_axis(self):
        return None

    @property
    def encoding(self):
        self._materialize()
        return self._encoding

    @encoding.setter
    def encoding(self, encoding):
        if self._is_encoding_frozen:
            raise RuntimeError("Encoding can be set only when it is not frozen.")
        self._drop_device_encoding()
        self._encoding = encoding

    @property
    def encoding_min_max_fixed_vals(self) -> Optional[Tuple[float, float]]:
        return self._encoding_min_max_fixed_vals

    @encoding_min_max_fixed_vals.setter
    def encoding_min_max_fixed_vals(self, min_max_vals: Tuple[float, float]):
        assert isinstance(min_max_vals, tuple) and len(min_max_vals) == 2, "Min max vals must be a tuple of two values"
        assert min_max_vals[0] < min_max_vals[1], f"Min value {min_max_vals[0]} is not less than max val {min_max_vals[1]}"
        if self.quant_scheme != QuantScheme.post_training_tf:
            self.quant_scheme = QuantScheme.post_training_tf
        self._encoding_min_max_fixed_vals = min_max_vals

    # ---- encodings -----------------------------------------------------------------------------------------------
    def _has_encoding(self) -> bool:
        return self._encoding is _LAZY or bool(self._encoding)

    def _drop_device_encoding(self):
        self._enc_dev = self._qdq4_dev = self._params_dev = None

    def _materialize(self):
        """Build the host-side TfEncoding objects of a device-resident result (one synchronising copy)."""
        if self._encoding is _LAZY:
            rows = self._enc_dev.cpu().tolist()
            self._encoding = [libpymo.TfEncoding._from_values(r[0], r[1], r[2], r[3], int(r[4])) for r in rows]   # pylint: disable=protected-access
            self._host_epoch = libpymo.encoding_epoch()

    def _device_columns(self):
        """(mins, maxs, deltas, offsets, bitwidths) as Python lists for a result that lives only on the device (None
        otherwise): one copy, five vectorised conversions -- the exporter builds 26 560 dictionaries per ResNet-50 job."""
        if self._encoding is _LAZY:
            a = self._enc_dev.cpu().numpy()
            return (a[:, 0].tolist(), a[:, 1].tolist(), a[:, 2].tolist(), a[:, 3].astype("int64").tolist(),
                    a[:, 4].astype("int64").tolist())
        return None

    def _device_encoding_valid(self) -> bool:
        """The device-side copy still describes the encodings: nothing has written to a TfEncoding since it was made."""
        if getattr(self, "_enc_dev", None) is None:
            return False
        return self._encoding is _LAZY or getattr(self, "_host_epoch", -1) == libpymo.encoding_epoch()

    def _native_block(self):
        """(arena, first, count, code) when all `_cppOp` are native ops over one contiguous block with valid statistics."""
        ops_ = self._cppOp
        op0 = ops_[0]
        if not _is_native(op0) or op0._block is None:   # pylint: disable=protected-access
            return None
        blk, idx0 = op0._block, op0._index              # pylint: disable=protected-access
        group = getattr(self, "_group", None)
        if group is not None and group.detached == 0 and blk is getattr(self, "_block", None) and idx0 == 0:
            # all channels were bound together by _ensure_block and nobody has touched an op individually since
            return (blk.arena, blk.first, len(ops_), op0._code) if group.valid else None   # pylint: disable=protected-access
        for i, op in enumerate(ops_):
            if op._block is not blk or op._index != idx0 + i or not op._is_encoding_valid:   # pylint: disable=protected-access
                return None
        return blk.arena, blk.first + idx0, len(ops_), op0._code   # pylint: disable=protected-access

    def _compute_encoding_on_device(self) -> bool:
        """Search on the device, keep the result there. False if this quantizer cannot take that route."""
        nb = self._native_block()
        if nb is None or (self.use_symmetric_encodings and self.use_unsigned_symmetric):
            return False        # is_unsigned_symmetric needs the values on the host anyway
        arena, first, count, code = nb
        per_tensor = self.channel_axis is None
        enc, qdq4 = ops.compute_encodings_impl(arena, first, count, code, self.bitwidth, self.use_symmetric_encodings,
                                               self.use_strict_symmetric, self.use_unsigned_symmetric,
                                               want_qdq4=per_tensor, percentile=self._cppOp[0]._percentile)   # pylint: disable=protected-access
        self._enc_dev = enc
        self._qdq4_dev = qdq4
        self._params_dev = None if per_tensor else ops.per_channel_params_dev(enc, self.bitwidth)
        self._encoding = _LAZY
        self.is_unsigned_symmetric = False
        return True

    def _collect_encodings(self):
        """[(TfEncoding, is_valid)] for every native op: the generic, one-call-per-op route (reference :296-299)."""
        return [op.getEncoding(self.bitwidth, self.use_symmetric_encodings, self.use_strict_symmetric,
                               self.use_unsigned_symmetric) for op in self._cppOp]

    def compute_encoding(self):
        """reference :280-321. Recomputing from unchanged statistics returns the same encoding (the reference does it for
        every parameter after calibration, 26 560 extra native calls for per-channel ResNet-50); that case is skipped."""
        if self.enabled and not self._is_encoding_frozen:
            if self._has_encoding() and not self._stats_dirty:
                return
            self._stats_dirty = False
            self._drop_device_encoding()
            self._encoding = []
            if self.bitwidth == 32:
                self._encoding = None
                return
            if self._lazy_ok and self._compute_encoding_on_device():
                return
            for encoding, is_valid in self._collect_encodings():
                if not is_valid:
                    self.enabled = False
                else:
                    self._encoding.append(encoding)
            self.is_unsigned_symmetric = self.use_symmetric_encodings and self.use_unsigned_symmetric and \
                all(enc.min >= 0 and enc.max >= 0 for enc in self._encoding)
            if not self.enabled and self._encoding:
                raise AssertionError("At least one encoding for a multi-encoding quantizer is invalid.")

    def refresh_encoding_from(self, tensor: torch.Tensor) -> bool:
        """reset_encoding_stats(); update_encoding_stats(tensor); compute_encoding() -- what the wrapper does with every
        parameter before a training-mode forward (reference qc_quantize_op.py:753-798) -- as ONE native call that enqueues
        the same kernels (ab_stats_refresh_encodings). Returns False when this quantizer cannot take that route (not on the
        device-resident path, frozen, fixed min/max, a calibration hook, the percentile scheme, ...): the caller then
        makes the three calls."""
        if not (FUSED_REFRESH and self.enabled and self._lazy_ok and not self._is_encoding_frozen and self.bitwidth != 32):
            return False
        if getattr(self, "_calib_hook", None) is not None or self.encoding_min_max_fixed_vals is not None:
            return False
        if self.use_symmetric_encodings and self.use_unsigned_symmetric:
            return False        # is_unsigned_symmetric needs the values on the host
        op0 = self._cppOp[0]
        if not _is_native(op0) or op0._percentile is not None or not tensor.is_cuda or \
                tensor.dtype not in (torch.float32, torch.bfloat16):   # pylint: disable=protected-access
            return False
        n = len(self._cppOp)
        if n == 1:
            if self.channel_axis is not None:
                return False
            op0._ensure_state(tensor.device)                                       # pylint: disable=protected-access
            arena, first = op0._block.arena, op0._block.first + op0._index          # pylint: disable=protected-access
            data = tensor if (tensor.is_contiguous() or tensor.is_contiguous(memory_format=torch.channels_last)) \
                else tensor.contiguous()
            seg_len = data.numel()
        else:
            self._ensure_block(tensor.device)
            if self._group is None or self._group.detached != 0:
                return False
            arena, first = self._block.arena, self._block.first
            data = tensor if self._ch_axis == 0 else tensor.movedim(self._ch_axis, 0)
            data = data.contiguous(memory_format=torch.contiguous_format)
            seg_len = data.numel() // n
        self.__dict__.pop("_reset_is_pending", None)     # this call resets the records itself
        enc, qdq4, params = ops.stats_refresh_encodings_impl(data, arena, first, n, seg_len, op0._code, self.bitwidth,   # pylint: disable=protected-access
                                                             self.use_symmetric_encodings, self.use_strict_symmetric,
                                                             self.use_unsigned_symmetric)
        # the bookkeeping of resetEncodingStats + updateStats on the native ops ...
        if n == 1:
            op0._is_encoding_valid = True                                          # pylint: disable=protected-access
            op0._range_fixed, op0._probe, op0._updates = False, None, 1            # pylint: disable=protected-access
        else:
            self._group.valid = True
            for op in self._cppOp:
                op._range_fixed, op._probe, op._updates = False, None, 1           # pylint: disable=protected-access
        # ... and of reset_encoding_stats + update_encoding_stats + compute_encoding on this quantizer
        self._enc_dev, self._qdq4_dev, self._params_dev = enc, qdq4, params
        self._encoding = _LAZY
        self.is_unsigned_symmetric = False
        self._stats_dirty = False
        return True

    def quantize_dequantize(self, tensor: torch.Tensor, round_mode) -> torch.Tensor:
        if not (torch.is_grad_enabled() and tensor.requires_grad):
            return QuantizeDequantize.run(tensor, self, round_mode)
        return QuantizeDequantize.apply(tensor, self, round_mode)

    def quantize(self, tensor: torch.Tensor, round_mode) -> torch.Tensor:
        return Quantize.apply(tensor, self, round_mode)

    def reset_encoding_stats(self):
        if not self._is_encoding_frozen:
            if self.__dict__.get("_reset_is_pending") or self.__dict__.pop("_reset_done_blockwide", False):
                # planned parameter quantizer (quantsim.param_plan): the refresh that follows resets the whole block of
                # records in one launch; activation quantizer of a sim: prepare_sim_for_compute_encodings has just reset
                # the sim's whole activation block in one launch -- only the host-side bookkeeping happens here
                self._mark_ops_invalid()
            else:
                self._reset_ops()
            self._encoding = None
            self._drop_device_encoding()
            self._stats_dirty = True

    def _mark_ops_invalid(self):
        group = getattr(self, "_group", None)
        if group is not None and group.detached == 0:
            group.valid = False
        else:
            for op in self._cppOp:
                op._reset_host_state()   # pylint: disable=protected-access

    def _reset_ops(self):
        for op in self._cppOp:
            op.resetEncodingStats()

    def get_stats_histogram(self) -> List[List]:
        if self._quant_scheme != QuantScheme.post_training_tf_enhanced:
            raise RuntimeError("get_stats_histogram() can be invoked only when quantization scheme is TF-Enhanced.")
        if not self._has_encoding():
            raise RuntimeError("get_stats_histogram() can be invoked only when encoding is computed.")
        return [op.getStatsHistogram() for op in self._cppOp]

    def freeze_encoding(self):
        if not self._has_encoding():
            raise RuntimeError("Encoding can be frozen only when it is not None.")
        self._is_encoding_frozen = True

    def set_percentile_value(self, percentile_value: float):
        for op in self._cppOp:
            op.setPercentileValue(percentile_value)

    # ---- pickling: native objects are re-created, statistics are dropped (reference :128-220) --------------------
    def __getstate__(self):
        self._materialize()
        state = self.__dict__.copy()
        state["_cppOp"] = len(self._cppOp)
        state.pop("_block", None)
        state.pop("_group", None)
        state.pop("_op_factory", None)
        for k in ("_enc_dev", "_qdq4_dev", "_params_dev"):
            state[k] = None
        return state

    def __setstate__(self, state):
        n = state.pop("_cppOp")
        self.__dict__.update(state)
        self._op_factory = _DEFAULT_OP_FACTORY
        self._block = None
        self._group = None
        self._cppOp = [self._op_factory(MAP_QUANT_SCHEME_TO_PYMO[self._quant_scheme]) for _ in range(n)]


class StaticGridPerTensorQuantizer(StaticGridTensorQuantizer):
    """reference :408-480"""

    def __init__(self, bitwidth, round_mode, quant_scheme, use_symmetric_encodings, enabled_by_default,
                 data_type=QuantizationDataType.int):
        super().__init__(bitwidth, round_mode, quant_scheme, use_symmetric_encodings, enabled_by_default, data_type)
        self._cppOp = [self._op_factory(MAP_QUANT_SCHEME_TO_PYMO[quant_scheme])]
        self._block = None

    @property
    def encoding(self):
        self._materialize()
        return self._encoding[0] if self._encoding else None

    @encoding.setter
    def encoding(self, encoding):
        if self._is_encoding_frozen:
            raise RuntimeError("Encoding can be set only when it is not frozen.")
        self._drop_device_encoding()
        self._encoding = encoding if isinstance(encoding, list) and len(encoding) == 1 else [encoding]

    def update_encoding_stats(self, tensor: torch.Tensor):
        """reference :452-480"""
        if self.enabled and not self._is_encoding_frozen:
            if self.bitwidth == 32:
                return
            if self.__dict__.pop("_reset_is_pending", None):
                self._reset_ops()        # the block-wide reset never came: do this quantizer's own now
            hook = getattr(self, "_calib_hook", None)
            if hook is not None and self.encoding_min_max_fixed_vals is None:
                hook(tensor)        # sharded calibration (aimet_b200.distributed) records / logs the call itself
                return
            self._stats_dirty = True
            if self.encoding_min_max_fixed_vals is not None:
                tensor = torch.tensor([self.encoding_min_max_fixed_vals[0], self.encoding_min_max_fixed_vals[1]])
            for op in self._cppOp:
                if tensor.dtype == torch.float16:
                    tensor = tensor.to(torch.float32)
                op.updateStats(tensor, tensor.is_cuda)


class StaticGridPerChannelQuantizer(StaticGridTensorQuantizer):
    """reference :483-570"""

    def __init__(self, bitwidth, round_mode, quant_scheme, use_symmetric_encodings, num_channels, enabled_by_default,
                 ch_axis: int = 0, data_type=QuantizationDataType.int):
        super().__init__(bitwidth, round_mode, quant_scheme, use_symmetric_encodings, enabled_by_default, data_type)
        self._cppOp = [self._op_factory(MAP_QUANT_SCHEME_TO_PYMO[quant_scheme]) for _ in range(num_channels)]
        self._ch_axis = ch_axis
        self._block = None          # one contiguous block of statistics records shared by all channels (native ops)
        self._group = None          # their shared isEncodingValid flag (tensor_quantizer_op.ValidityGroup)
        self._params_cache = None

    @property
    def channel_axis(self) -> int:
        return self._ch_axis

    def _ensure_block(self, device):
        if self._block is None or self._block.device != device:
            self._block = StateArena.for_device(device).allocate(len(self._cppOp))
            self._group = ValidityGroup()
            for i, op in enumerate(self._cppOp):
                op._bind(self._block, i, self._group)   # pylint: disable=protected-access

    def _reset_ops(self):
        if self._block is not None and _is_native(self._cppOp[0]):
            self._block.reset()
            if self._group is not None and self._group.detached == 0:
                self._group.valid = False
            else:
                for op in self._cppOp:
                    op._is_encoding_valid = False   # pylint: disable=protected-access
        elif _is_native(self._cppOp[0]) and not self.enabled and not getattr(self, "_ops_used_individually", False):
            # switched off, never bound to a block, its per-channel ops never driven one by one: nothing to clear. (The bias
            # quantizers of a per-channel model own one op per output channel -- 26 560 for ResNet-50 -- and walking them on
            # every reset was 5 of the 6.6 ms a calibration job spent in prepare_sim_for_compute_encodings.)
            return
        else:
            super()._reset_ops()

    def update_encoding_stats(self, tensor: torch.Tensor):
        """reference :537-570"""
        if self.enabled and not self._is_encoding_frozen:
            if self.bitwidth == 32:
                return
            if self.__dict__.pop("_reset_is_pending", None):
                self._reset_ops()        # the block-wide reset never came: do this quantizer's own now
            self._stats_dirty = True
            if self.encoding_min_max_fixed_vals is not None:
                tensor = torch.tensor([self.encoding_min_max_fixed_vals[0], self.encoding_min_max_fixed_vals[1]])
                self._ops_used_individually = True
                for op in self._cppOp:
                    op.updateStats(tensor, tensor.is_cuda)
                return
            if tensor.dtype == torch.float16:
                tensor = tensor.to(torch.float32)
            if _is_native(self._cppOp[0]) and tensor.is_cuda and tensor.dtype in (torch.float32, torch.bfloat16):
                # all channels in one launch: channel-major contiguous view, one segment per channel
                n_ch = len(self._cppOp)
                moved = tensor if self._ch_axis == 0 else tensor.movedim(self._ch_axis, 0)
                moved = moved.contiguous(memory_format=torch.contiguous_format)
                self._ensure_block(tensor.device)
                ops.stats_update_segmented_impl(moved, self._block.arena, self._block.first, n_ch,
                                                moved.numel() // n_ch, self._cppOp[0]._code)   # pylint: disable=protected-access
                if self._group is not None and self._group.detached == 0:
                    self._group.valid = True
                else:
                    for op in self._cppOp:
                        op._is_encoding_valid = True   # pylint: disable=protected-access
                return
            self._ops_used_individually = True
            for channel_idx, op in enumerate(self._cppOp):
                tensor_slice = tensor.select(self._ch_axis, channel_idx).contiguous(memory_format=torch.contiguous_format)
                op.updateStats(tensor_slice, tensor.is_cuda)

    def _collect_encodings(self):
        op0 = self._cppOp[0]
        if _is_native(op0) and self._block is not None and all(op._block is self._block for op in self._cppOp):   # pylint: disable=protected-access
            if not all(op._is_encoding_valid for op in self._cppOp):   # pylint: disable=protected-access
                return [(libpymo.TfEncoding(), op._is_encoding_valid) for op in self._cppOp]   # pylint: disable=protected-access
            enc, _ = ops.compute_encodings_impl(self._block.arena, self._block.first, len(self._cppOp), op0._code,   # pylint: disable=protected-access
                                                self.bitwidth, self.use_symmetric_encodings, self.use_strict_symmetric,
                                                self.use_unsigned_symmetric, percentile=op0._percentile)   # pylint: disable=protected-access
            rows = enc.cpu().tolist()
            return [(libpymo.TfEncoding._from_values(r[0], r[1], r[2], r[3], int(r[4])), True) for r in rows]   # pylint: disable=protected-access
        return super()._collect_encodings()


# ---------------------------------------------------------------------------------------------------------------------
# autograd functions
# ---------------------------------------------------------------------------------------------------------------------
def scalar_in_dtype(value, dtype) -> float:
    """float(torch.tensor(float(value), dtype=torch.float32).to(dtype)) without building tensors: the value narrowed to
    float32 (round to nearest even, overflow to infinity), then to bfloat16 / float16 the same way. The backward of every
    activation quantizer needs two of these per step."""
    value = float(value)
    try:
        f32 = struct.unpack("<f", struct.pack("<f", value))[0]
    except OverflowError:
        f32 = math.copysign(math.inf, value)
    if dtype == torch.float32 or f32 != f32 or math.isinf(f32):
        return f32
    if dtype == torch.bfloat16:
        bits = struct.unpack("<I", struct.pack("<f", f32))[0]
        bits = (bits + 0x7fff + ((bits >> 16) & 1)) & 0xffff0000
        return struct.unpack("<f", struct.pack("<I", bits))[0]
    if dtype == torch.float16:
        try:
            return struct.unpack("<e", struct.pack("<e", f32))[0]
        except OverflowError:
            return math.copysign(math.inf, f32)
    raise TypeError(dtype)


def compute_dloss_by_dx(x, grad, encoding_min, encoding_max, ch_axis=0):
    """Straight-through estimator, reference quantsim_straight_through_grad.py:91-118: grad * [min <= x <= max].
    On CUDA tensors this is one fused kernel (3 tensors of traffic) instead of three element-wise torch kernels."""
    if not x.is_cuda:
        raise RuntimeError("aimet_b200 has no CPU path")
    if isinstance(encoding_min, (list, tuple)):
        mins = torch.tensor(encoding_min).to(x.device)       # float32, as torch.tensor(list of python floats) gives
        maxs = torch.tensor(encoding_max).to(x.device)
        if mins.numel() > 1:
            n_ch = x.shape[ch_axis]
            assert mins.numel() == n_ch
            per_channel = 1
            for d in x.shape[ch_axis + 1:]:
                per_channel *= d
            if x.dtype == torch.float16:      # same route as the per-tensor branch and both forwards: upcast, cast back
                return ops.ste_bwd_per_channel_impl(x.float(), grad.float(), mins, maxs, n_ch,
                                                    per_channel).to(torch.float16)
            return ops.ste_bwd_per_channel_impl(x, grad, mins, maxs, n_ch, per_channel)
        encoding_min, encoding_max = float(mins), float(maxs)
    # torch.tensor(python float) is a 0-dim float32 tensor; compared with a lower-precision tensor it does not promote,
    # i.e. the reference compares in x's dtype with the range rounded to that dtype
    cmp_dtype = x.dtype if x.dtype in (torch.bfloat16, torch.float16) else torch.float32
    lo, hi = scalar_in_dtype(encoding_min, cmp_dtype), scalar_in_dtype(encoding_max, cmp_dtype)
    if x.dtype == torch.float16:
        return ops.ste_bwd_impl(x.float(), grad.float(), lo, hi).to(torch.float16)
    return ops.ste_bwd_impl(x, grad, lo, hi)


class QuantizeDequantize(torch.autograd.Function):
    """reference :1098-1212"""

    @staticmethod
    def _per_tensor(tensor, tensor_quantizer, round_mode):
        dtype = tensor.dtype
        if dtype not in (torch.float32, torch.bfloat16):
            tensor = tensor.to(torch.float32)
        q = tensor_quantizer
        if q._qdq4_dev is not None and tensor.is_cuda and not tensor.requires_grad and q._device_encoding_valid():   # pylint: disable=protected-access
            out = ops.qdq_per_tensor_dev_impl(tensor, q._qdq4_dev[0], int(round_mode), 0)   # pylint: disable=protected-access
        else:
            out = q._cppOp[0].quantizeDequantize(tensor, q.encoding, round_mode, tensor.is_cuda)   # pylint: disable=protected-access
        return out if out.dtype == dtype else out.to(dtype)

    @staticmethod
    def _per_channel(tensor, tensor_quantizer, round_mode):
        dtype = tensor.dtype
        if dtype not in (torch.float32, torch.bfloat16):
            tensor = tensor.to(torch.float32)
        sizes = [*tensor.shape, 1]
        num_channel = sizes[tensor_quantizer.channel_axis]
        num_element = functools.reduce(lambda x, y: x * y, sizes)
        num_element_per_channel = functools.reduce(lambda x, y: x * y, sizes[tensor_quantizer.channel_axis + 1:])
        q = tensor_quantizer
        if q._params_dev is not None and tensor.is_cuda and not tensor.requires_grad and q._device_encoding_valid():   # pylint: disable=protected-access
            out = ops.qdq_per_channel_impl(tensor, q._params_dev, num_channel, num_element_per_channel,   # pylint: disable=protected-access
                                           int(round_mode), 0)
        else:
            out = q._cppOp[0].quantizeDequantizePerChannel(tensor, q.encoding, num_channel, num_element,   # pylint: disable=protected-access
                                                           num_element_per_channel, round_mode, tensor.is_cuda)
        return out if out.dtype == dtype else out.to(dtype)

    @staticmethod
    def run(tensor, tensor_quantizer, round_mode):
        """The forward computation without autograd bookkeeping."""
        if tensor_quantizer.enabled and tensor_quantizer.bitwidth != 32:
            if isinstance(tensor_quantizer, StaticGridPerChannelQuantizer):
                return QuantizeDequantize._per_channel(tensor, tensor_quantizer, round_mode)
            return QuantizeDequantize._per_tensor(tensor, tensor_quantizer, round_mode)
        return tensor

    @staticmethod
    def forward(ctx, tensor, tensor_quantizer, round_mode):   # pylint: disable=arguments-differ
        out = QuantizeDequantize.run(tensor, tensor_quantizer, round_mode)
        if out is not tensor:
            ctx.save_for_backward(tensor)
        ctx.tensor_quantizer = tensor_quantizer
        return out

    @staticmethod
    def backward(ctx, output_grad):   # pylint: disable=arguments-differ
        q = ctx.tensor_quantizer
        if q.enabled and q.data_type == QuantizationDataType.int and q.bitwidth != 32:
            (tensor,) = ctx.saved_tensors
            if isinstance(q, StaticGridPerChannelQuantizer):
                # (the reference reads `.encoding.min` here, which only exists per tensor; the per-channel weights get
                # their gradient gate from SteGatingFuncForParameters instead -- we gate consistently in both places)
                grad = compute_dloss_by_dx(tensor, output_grad, [e.min for e in q.encoding],
                                           [e.max for e in q.encoding], q.channel_axis)
            else:
                grad = compute_dloss_by_dx(tensor, output_grad, q.encoding.min, q.encoding.max, 0)
        else:
            grad = output_grad
        return grad, None, None


class Quantize(torch.autograd.Function):
    """Quantize-only to the integer grid, reference :1215-1282."""

    @staticmethod
    def forward(ctx, tensor, tensor_quantizer, round_mode):   # pylint: disable=arguments-differ
        assert tensor_quantizer.enabled, "Tensor quantizer must be enabled to perform quantize only."
        assert tensor_quantizer.bitwidth != 32, "Tensor quantizer bitwidth must be < 32 to perform quantize only."
        assert tensor_quantizer.encoding is not None, "Tensor quantizer encoding must be valid to perform quantize only."
        shift_to_signed = not (tensor_quantizer.use_symmetric_encodings and tensor_quantizer.use_unsigned_symmetric)
        if isinstance(tensor_quantizer, StaticGridPerChannelQuantizer):
            outs = []
            for index, op in enumerate(tensor_quantizer._cppOp):   # pylint: disable=protected-access
                tensor_slice = tensor.select(tensor_quantizer.channel_axis, index).contiguous(
                    memory_format=torch.contiguous_format)
                outs.append(op.quantize(tensor_slice, tensor_quantizer._encoding[index], round_mode, tensor.is_cuda,   # pylint: disable=protected-access
                                        shift_to_signed))
            return torch.stack(tuple(outs), dim=tensor_quantizer.channel_axis)
        return tensor_quantizer._cppOp[0].quantize(tensor, tensor_quantizer.encoding, round_mode, tensor.is_cuda,   # pylint: disable=protected-access
                                                   shift_to_signed)

    @staticmethod
    def backward(ctx, _output_grad):   # pylint: disable=arguments-differ
        raise AssertionError("Backward pass for quantize only not implemented")
