"""Host mirror of the reference's static-grid quantization wrapper.

Reference: TrainingExtensions/torch/src/python/aimet_torch/v1/qc_quantize_op.py -- QcQuantizeOpMode :63-70,
QcQuantizeWrapper :82-677, StaticGridQuantWrapper :679-924, SteGatingFuncForParameters :1314-1366.
"""
import enum
from typing import Dict, List, Optional, Tuple

import torch
from torch import nn

from .. import libpymo
from .defs import MAP_ROUND_MODE_TO_PYMO, QuantizationDataType, QuantScheme
from .tensor_quantizer import (_LAZY, StaticGridPerChannelQuantizer, StaticGridPerTensorQuantizer, compute_dloss_by_dx)

import os

# reference :75-77: AIMET_TFE_USE_DOWNSAMPLING=1 strides the tensor handed to tf_enhanced statistics
TF_ENHANCED_USE_DOWNSAMPLING = bool(int(os.environ.get("AIMET_TFE_USE_DOWNSAMPLING", "0")))
TF_ENHANCED_OFFSET_FACTOR = 0
TF_ENHANCED_STRIDE_FACTOR = 2

_IGNORED_DTYPES = (torch.int, torch.int8, torch.int16, torch.int32, torch.int64, torch.bool, torch.uint8)


class QcQuantizeOpMode(enum.Enum):
    PASSTHROUGH = 1
    ANALYSIS = 2
    ACTIVE = 3
    LEARN_ENCODINGS = 4


def export_quantizer_encoding(quantizer) -> Optional[List[Dict]]:
    """reference :1529-1545 + aimet_torch/utils.py:1156-1185 (create_encoding_dict)"""
    if not quantizer.enabled:
        return None
    if quantizer.data_type == QuantizationDataType.int and quantizer.bitwidth == 32:
        return None
    # dictionaries built while the calibration forwards were running (quantsim._ParamExportPrefetch): served once, and only
    # if the encodings still are the device table they were made from
    cache = quantizer.__dict__.pop("_export_cache", None)
    if cache is not None and hasattr(quantizer, "_device_columns"):
        from .tensor_quantizer import _LAZY
        table, sym, dicts = cache
        if quantizer._encoding is _LAZY and table is quantizer._enc_dev and sym == str(quantizer.use_symmetric_encodings):   # pylint: disable=protected-access
            return dicts
    cols = quantizer._device_columns() if hasattr(quantizer, "_device_columns") else None   # pylint: disable=protected-access
    if cols is not None:
        # encodings that still live only on the device: one copy, dictionaries built straight from the columns (no
        # intermediate TfEncoding objects or per-row lists -- 26 560 of them for per-channel ResNet-50)
        sym = str(quantizer.use_symmetric_encodings)
        return [{"min": mn, "max": mx, "scale": sc, "offset": off, "bitwidth": bw, "is_symmetric": sym, "dtype": "int"}
                for mn, mx, sc, off, bw in zip(*cols)]
    # a learned-grid quantizer exports its effective encoding (reference get_encoding_by_quantizer :1514-1526)
    encoding = quantizer.get_effective_encoding() if hasattr(quantizer, "get_effective_encoding") else quantizer.encoding

    def to_dict(enc):
        if not enc:
            return None
        return {"min": enc.min, "max": enc.max, "scale": enc.delta, "offset": int(enc.offset), "bitwidth": enc.bw,
                "is_symmetric": str(quantizer.use_symmetric_encodings), "dtype": "int"}

    if isinstance(encoding, list):
        return [to_dict(e) for e in encoding]
    d = to_dict(encoding)
    return [d] if d else None


def ste_for_quantizer(x, grad, q):
    """compute_dloss_by_dx with the quantizer's current encoding range; when that encoding lives on the device (see
    tensor_quantizer._LAZY) the range is taken from there -- float32(min), float32(max), exactly what
    torch.tensor(python floats) yields in the reference -- and no host synchronisation happens."""
    from .. import ops
    if q._device_encoding_valid() and x.is_cuda and x.dtype in (torch.float32, torch.bfloat16):   # pylint: disable=protected-access
        enc5 = q._enc_dev                                                                           # pylint: disable=protected-access
        n_ch = enc5.shape[0]
        axis = q.channel_axis if q.channel_axis is not None else 0
        per_channel = 1
        for d in (x.shape[axis + 1:] if n_ch > 1 else x.shape):
            per_channel *= d
        # a per-tensor bound is a 0-dim tensor in the reference: compared in x's dtype
        return ops.ste_bwd_enc5_impl(x, grad, enc5, n_ch, per_channel,
                                     range_in_bf16=x.dtype == torch.bfloat16 and q.channel_axis is None)
    enc = q.encoding
    if isinstance(enc, list):
        return compute_dloss_by_dx(x, grad, [e.min for e in enc], [e.max for e in enc], q.channel_axis)
    return compute_dloss_by_dx(x, grad, enc.min, enc.max)


_NEVER_IN_PLACE = (nn.modules.conv._ConvNd, nn.Linear, nn.modules.batchnorm._BatchNorm, nn.LayerNorm, nn.GroupNorm,   # pylint: disable=protected-access
                   nn.Embedding, nn.modules.pooling._AvgPoolNd, nn.modules.pooling._MaxPoolNd,                        # pylint: disable=protected-access
                   nn.modules.pooling._AdaptiveAvgPoolNd, nn.modules.pooling._AdaptiveMaxPoolNd, nn.Flatten, nn.Identity)  # pylint: disable=protected-access


class ForwardToken:
    """One model forward in flight (set by the sim's forward hooks): parameter encodings stamped with `current` while
    `active` were refreshed for this very forward by quantsim.param_plan, all parameters of the model in one native call."""
    current = 0
    active = False


_FORWARD = ForwardToken


class CalibrationJob:
    """Set by QuantizationSimModel.compute_encodings (and the sharded calibrator) around the user's calibration callback."""
    active = False

    def __init__(self, sim):
        self.sim = sim

    def __enter__(self):
        self.previous, CalibrationJob.active = CalibrationJob.active, True
        return self

    def __exit__(self, *exc):
        CalibrationJob.active = self.previous
        for _, w in self.sim.quant_wrappers():
            for q in getattr(w, "param_quantizers", {}).values():
                q.__dict__.pop("_qdq_cache", None)
ALWAYS_GATE_AND_CLONE = False   # test hook: the reference's unconditional gating + clone in every wrapper


def _leaves_its_input_alone(module: nn.Module) -> bool:
    """True for torch.nn modules whose forward never writes to its input tensor."""
    if type(module).__module__.startswith("torch.nn."):
        if isinstance(module, _NEVER_IN_PLACE):
            return True
        return getattr(module, "inplace", None) is False
    return False


class SteGatingFuncForParameters(torch.autograd.Function):
    """Gates the parameter gradients with the straight-through estimator after the wrapped module's backward
    (reference :1314-1366)."""

    @staticmethod
    def forward(ctx, quant_wrapper_ref, *quantized_inputs):   # pylint: disable=arguments-differ
        ctx.quantization_wrapper_ref = quant_wrapper_ref
        return quantized_inputs

    @staticmethod
    def backward(ctx, *output_grad):   # pylint: disable=arguments-differ
        wrapper = ctx.quantization_wrapper_ref
        for name, param in wrapper.get_named_parameters():
            q = wrapper.param_quantizers[name]
            if q.bitwidth == 32 or q.data_type == QuantizationDataType.float:
                continue
            if q.enabled and param.grad is not None:
                param.grad = ste_for_quantizer(param, param.grad, q)
        return (None, *output_grad)


class EncodingImportMixin:
    """import_param_encodings / import_input_encodings / import_output_encodings of the reference's QcQuantizeWrapper
    (:499-677), which both StaticGridQuantWrapper and LearnedGridQuantWrapper inherit there."""

    @staticmethod
    def _encoding_from_dict(quantizer, enc_dict):
        """utils.create_encoding_from_dict + compute_partial_encoding (reference :1548-1570): fill whatever half of
        (min, max) / (scale, offset) the dictionary leaves out through the native computePartialEncoding."""
        enc = libpymo.TfEncoding()
        enc.bw = int(enc_dict["bitwidth"])
        enc.min = float(enc_dict.get("min", 0.0))
        enc.max = float(enc_dict.get("max", 0.0))
        enc.delta = float(enc_dict.get("scale", 0.0))
        enc.offset = float(enc_dict.get("offset", 0.0))
        if (enc.min == 0 and enc.max == 0) or enc.delta == 0:
            tq = libpymo.TensorQuantizer(libpymo.QuantizationMode.QUANTIZATION_TF, libpymo.RoundingMode.ROUND_NEAREST)
            tq.computePartialEncoding(enc.bw, enc, quantizer.use_symmetric_encodings,
                                      quantizer.use_unsigned_symmetric, quantizer.use_strict_symmetric)
        return enc

    def _import_quantizer(self, quantizer, enc_dicts, strict, partial, requires_grad, allow_overwrite):
        if quantizer.is_encoding_frozen:
            return
        if not enc_dicts:
            if not partial:
                quantizer.enabled = False       # dangling quantizers are removed when the encodings are not partial
            return
        if not quantizer.enabled:
            if strict:
                raise RuntimeError("The quantsim passed for loading encodings does not have the same "
                                   "configuration as the quantsim which was used to export the encodings")
            return
        if isinstance(enc_dicts, dict):
            enc_dicts = [enc_dicts]
        if enc_dicts[0].get("dtype", "int") != "int":
            raise NotImplementedError("float encodings are outside the aimet_b200 hot path")
        is_symmetric = enc_dicts[0]["is_symmetric"] == "True"
        quantizer.use_symmetric_encodings = is_symmetric
        if not is_symmetric:
            quantizer.use_unsigned_symmetric = False
            quantizer.use_strict_symmetric = False
        encodings = [self._encoding_from_dict(quantizer, d) for d in enc_dicts]
        quantizer.bitwidth = encodings[0].bw
        learned = getattr(quantizer, "wrapper_ref", None) is not None      # LearnedGridTensorQuantizer
        if learned:
            # the setter re-creates the <name>_encoding_min / _max parameters of the wrapper (reference :851-883)
            quantizer.encoding = encodings if len(encodings) > 1 else encodings[0]
            if requires_grad is not None:
                for suffix in ("_encoding_min", "_encoding_max"):
                    getattr(quantizer.wrapper_ref, quantizer.name + suffix).requires_grad_(requires_grad)
        elif hasattr(quantizer, "_ch_axis"):
            if len(encodings) != len(quantizer._cppOp):   # pylint: disable=protected-access
                raise RuntimeError("number of per-channel encodings does not match the number of channels")
            quantizer.encoding = encodings
        else:
            quantizer.encoding = encodings[0]
        if hasattr(quantizer, "_stats_dirty"):
            quantizer._stats_dirty = False                 # pylint: disable=protected-access
        if allow_overwrite is False and quantizer.encoding is not None:
            quantizer.freeze_encoding()

    def import_input_encodings(self, encodings, strict, partial, requires_grad, allow_overwrite):
        for i, q in enumerate(self.input_quantizers):
            self._import_quantizer(q, encodings.get(str(i), encodings.get(i)), strict, partial, requires_grad,
                                   allow_overwrite)

    def import_output_encodings(self, encodings, strict, partial, requires_grad, allow_overwrite):
        for i, q in enumerate(self.output_quantizers):
            self._import_quantizer(q, encodings.get(str(i), encodings.get(i)), strict, partial, requires_grad,
                                   allow_overwrite)

    def import_param_encodings(self, encodings, strict, partial, requires_grad, allow_overwrite):
        for name, q in self.param_quantizers.items():
            self._import_quantizer(q, encodings.get(name), strict, partial, requires_grad, allow_overwrite)


class StaticGridQuantWrapper(EncodingImportMixin, nn.Module):
    """Wraps one leaf module: quantizes its inputs, parameters and outputs around the wrapped forward."""

    def __init__(self, module_to_wrap: nn.Module, weight_bw: int, activation_bw: int, round_mode, quant_scheme,
                 is_output_quantized=True, is_symmetric=False, num_inputs=1, num_outputs=1,
                 data_type: QuantizationDataType = QuantizationDataType.int):
        super().__init__()
        if isinstance(round_mode, str):
            round_mode = MAP_ROUND_MODE_TO_PYMO[round_mode]
        if isinstance(quant_scheme, str):
            quant_scheme = QuantScheme.from_str(quant_scheme)
        self._module_to_wrap = module_to_wrap
        self._mode = QcQuantizeOpMode.ANALYSIS
        self._quant_scheme = quant_scheme
        self.output_quantizers = [StaticGridPerTensorQuantizer(activation_bw, round_mode, quant_scheme, is_symmetric,
                                                               enabled_by_default=is_output_quantized,
                                                               data_type=data_type) for _ in range(num_outputs)]
        self.input_quantizers = [StaticGridPerTensorQuantizer(activation_bw, round_mode, quant_scheme, is_symmetric,
                                                              enabled_by_default=False, data_type=data_type)
                                 for _ in range(num_inputs)]
        self.param_quantizers = {}
        for name, _ in module_to_wrap.named_parameters():
            self.param_quantizers[name] = StaticGridPerTensorQuantizer(weight_bw, round_mode, quant_scheme,
                                                                       is_symmetric, enabled_by_default=True,
                                                                       data_type=data_type)
            self.param_quantizers[name]._lazy_ok = True   # pylint: disable=protected-access

    # ---- accessors the reference exposes -------------------------------------------------------------------------
    @property
    def output_quantizer(self):
        return self.output_quantizers[0]

    @property
    def input_quantizer(self):
        return self.input_quantizers[0]

    def get_original_module(self) -> nn.Module:
        return self._module_to_wrap

    def get_named_parameters(self):
        """reference :257-268. A replica made by torch.nn.DataParallel keeps its parameters in `_former_parameters` (plain
        tensors, views of the broadcast copies) and carries `_is_replica`."""
        if getattr(self, "_is_replica", False):
            return list(getattr(self._module_to_wrap, "_former_parameters", {}).items())
        params = self._module_to_wrap._parameters   # pylint: disable=protected-access
        if not self._module_to_wrap._modules:        # leaf module: its own parameters are all there is
            return [(k, v) for k, v in params.items() if v is not None]
        return list(self._module_to_wrap.named_parameters())

    def set_mode(self, mode: QcQuantizeOpMode):
        self._mode = mode

    def reset_encodings(self):
        """reference :231-243"""
        for q in self._all_quantizers():
            q.reset_encoding_stats()

    def _all_quantizers(self):
        return list(self.input_quantizers) + list(self.param_quantizers.values()) + list(self.output_quantizers)

    def enable_per_channel_quantization(self):
        """reference :899-920"""
        new = {}
        for name, param in self._module_to_wrap.named_parameters():
            q = self.param_quantizers[name]
            axis = 0
            if isinstance(self._module_to_wrap, (nn.ConvTranspose1d, nn.ConvTranspose2d, nn.ConvTranspose3d)) and \
                    len(param.shape) > 1:
                axis = 1
            pcq = StaticGridPerChannelQuantizer(q.bitwidth, q.round_mode, q.quant_scheme, q.use_symmetric_encodings,
                                                num_channels=param.shape[axis], enabled_by_default=q.enabled,
                                                ch_axis=axis, data_type=q.data_type)
            pcq.use_strict_symmetric = q.use_strict_symmetric
            pcq.use_unsigned_symmetric = q.use_unsigned_symmetric
            pcq._lazy_ok = True   # pylint: disable=protected-access
            new[name] = pcq
        self.param_quantizers = new

    # ---- forward -------------------------------------------------------------------------------------------------
    def forward(self, *inputs, **kwargs):
        """reference :705-745"""
        quantized_inputs = self._quantize_activation(self.input_quantizers, list(inputs))
        shadow_params = self._quantize_dequantize_params()
        # The reference routes the inputs through SteGatingFuncForParameters whenever grad mode is on (:716-726): in ITS
        # backward -- which runs after the wrapped module's -- that function overwrites `param.grad` of every quantized
        # parameter with grad * [min <= w <= max], and it clones the inputs to protect its aliased outputs from in-place
        # modules. Here the same gate is a gradient hook on the parameter itself (`_gate_parameter_gradients`): it acts on
        # the gradient BEFORE it is accumulated into `.grad`, which (i) yields the same `.grad`, (ii) needs no alias of the
        # inputs, hence no clone per layer and step, and (iii) is what makes the gate survive DistributedDataParallel, whose
        # reducer copies a gradient into its bucket the moment it is accumulated -- a later write to `param.grad` is
        # overwritten by the all-reduced bucket. ALWAYS_GATE_AND_CLONE restores the reference's form (test hook).
        if torch.is_grad_enabled():
            if ALWAYS_GATE_AND_CLONE:
                quantized_inputs = SteGatingFuncForParameters.apply(self, *quantized_inputs)
                quantized_inputs = [inp.clone() if isinstance(inp, torch.Tensor) and inp.requires_grad else inp
                                    for inp in quantized_inputs]
            elif shadow_params:
                # The reference's gating function only ever runs its backward when one of the wrapper's inputs requires
                # grad (an autograd function without a differentiable input is not part of the graph): the parameters of a
                # model's FIRST layer, fed by the data tensor, are therefore not gated there. Kept, for identical gradients.
                self._gate_armed = any(isinstance(t, torch.Tensor) and t.requires_grad for t in quantized_inputs)
                if self._gate_armed:
                    self._gate_parameter_gradients()
        wrapped_output = self._module_to_wrap(*quantized_inputs, **kwargs)
        self._restore_shadow_params(shadow_params)
        is_seq = isinstance(wrapped_output, (list, tuple))
        outputs = self._quantize_activation(self.output_quantizers, list(wrapped_output) if is_seq else [wrapped_output])
        return outputs[0] if len(outputs) == 1 else outputs

    def _gate_parameter_gradients(self):
        """Make sure every quantized parameter of the wrapped module carries the straight-through gate as a gradient hook
        (registered once per parameter tensor; a no-op while its quantizer is disabled)."""
        import weakref
        hooks = self.__dict__.setdefault("_ste_hooks", {})
        for name, param in self.get_named_parameters():
            if not param.requires_grad:
                continue
            known = hooks.get(name)
            live = getattr(param, "_backward_hooks", None) or {}
            if known is not None and known[0] is param and known[1].id in live:
                continue
            ref = weakref.ref(self)

            def gate(grad, name=name, ref=ref, pref=weakref.ref(param)):
                wrapper, p = ref(), pref()
                if wrapper is None or p is None or ALWAYS_GATE_AND_CLONE or not wrapper.__dict__.get("_gate_armed"):
                    return grad
                q = wrapper.param_quantizers.get(name)
                if q is None or not q.enabled or q.bitwidth == 32 or q.data_type == QuantizationDataType.float or \
                        not q._has_encoding():   # pylint: disable=protected-access
                    return grad
                return ste_for_quantizer(p.data, grad, q)

            hooks[name] = (param, param.register_hook(gate))

    def __getstate__(self):
        state = self.__dict__.copy()
        state.pop("_ste_hooks", None)          # gradient hooks are not serialised with the tensors: re-registered on use
        return state

    def _restore_shadow_params(self, shadow_params):
        for name, param in self.get_named_parameters():
            if name in shadow_params:
                param.data = shadow_params[name]

    def _quantize_dequantize_params(self):
        """reference :753-798. The reference clones every parameter and later copies it back in place; holding the
        original tensor aside and re-pointing `.data` gives the same values with two fewer passes over the weights."""
        shadow_params = {}
        for name, param in self.get_named_parameters():
            q = self.param_quantizers[name]
            if q.enabled and q.bitwidth != 32:
                shadow_params[name] = param.data
                if self._module_to_wrap.training or not q._has_encoding():   # pylint: disable=protected-access
                    if _FORWARD.active and q.__dict__.get("_fresh_token") == _FORWARD.current and q._has_encoding():   # pylint: disable=protected-access
                        pass   # refreshed for this very forward, together with all other parameters (param_plan)
                    elif not q.refresh_encoding_from(param.data):              # one native call where possible
                        q.reset_encoding_stats()
                        q.update_encoding_stats(param.data)
                        q.compute_encoding()
                round_mode = q.round_mode if self.training else libpymo.RoundingMode.ROUND_NEAREST
                if CalibrationJob.active and not self.training and not torch.is_grad_enabled() and \
                        q._encoding is _LAZY and round_mode == libpymo.RoundingMode.ROUND_NEAREST:   # pylint: disable=protected-access
                    # Inside one calibration job (eval mode, no_grad) neither the weights nor their encodings change from
                    # batch to batch: the quantize-dequantized weight of the first batch serves the others (the reference
                    # re-quantizes all weights on every forward; 53 launches per ResNet-50 step). Keyed on the parameter's
                    # storage and version and on the identity of the device-resident encoding table.
                    key = (param.data.data_ptr(), param._version, id(q._enc_dev))   # pylint: disable=protected-access
                    cached = q.__dict__.get("_qdq_cache")
                    if cached is None or cached[0] != key:
                        cached = q._qdq_cache = (key, q.quantize_dequantize(param.data, round_mode))   # pylint: disable=protected-access
                    param.data = cached[1]
                elif getattr(self, "_is_replica", False):
                    param.data = q.quantize_dequantize(param.data.clone(), round_mode)   # reference :785-786
                else:
                    param.data = q.quantize_dequantize(param.data, round_mode)
        return shadow_params

    def ensure_param_encodings(self):
        """Derive the encodings of this wrapper's parameters without running the wrapped module: what the first forward of a
        calibration job does on its way (reference :753-798). Sharded calibration needs it on a rank that was dealt no
        batch (the weights, hence these encodings, are the same on every rank)."""
        for name, param in self.get_named_parameters():
            q = self.param_quantizers[name]
            if q.enabled and q.bitwidth != 32 and not q._has_encoding():   # pylint: disable=protected-access
                if not q.refresh_encoding_from(param.data):
                    q.reset_encoding_stats()
                    q.update_encoding_stats(param.data)
                    q.compute_encoding()

    def compute_encoding(self):
        """reference :811-826"""
        for q in self.input_quantizers:
            q.compute_encoding()
        for q in self.param_quantizers.values():
            q.compute_encoding()
        for q in self.output_quantizers:
            q.compute_encoding()

    def set_percentile_value(self, percentile_value: float):
        """reference :827-835: the activation quantizers only"""
        for q in self.input_quantizers + self.output_quantizers:
            q.set_percentile_value(percentile_value)

    @staticmethod
    def should_perform_quant_dequant(tensor, tensor_quantizer) -> bool:
        """reference :451-473"""
        if not isinstance(tensor, torch.Tensor) or tensor.dtype in _IGNORED_DTYPES or \
                (tensor_quantizer.is_const and torch.numel(tensor) == 1) or not tensor_quantizer.enabled:
            tensor_quantizer.enabled = False
            return False
        return True

    def _quantize_activation(self, tensor_quantizers, tensors_to_quantize):
        """reference :837-897"""

        def inner(t, index):
            if isinstance(t, (list, tuple)):
                return [inner(x, index) for x in t]
            q = tensor_quantizers[index]
            if not self.should_perform_quant_dequant(t, q):
                return t
            if self._mode is QcQuantizeOpMode.ANALYSIS and not q.is_encoding_frozen:
                if TF_ENHANCED_USE_DOWNSAMPLING and q.quant_scheme == QuantScheme.post_training_tf_enhanced:
                    flat = t.reshape(-1)
                    q.update_encoding_stats(flat[TF_ENHANCED_OFFSET_FACTOR::TF_ENHANCED_STRIDE_FACTOR].contiguous())
                else:
                    q.update_encoding_stats(t)
                return t
            if self._mode is QcQuantizeOpMode.ACTIVE or (self._mode is QcQuantizeOpMode.ANALYSIS and
                                                         q.is_encoding_frozen):
                round_mode = q.round_mode if self.training else libpymo.RoundingMode.ROUND_NEAREST
                return q.quantize_dequantize(t, round_mode)
            return t

        outputs = []
        for index, t in enumerate(tensors_to_quantize):
            assert len(tensor_quantizers) > index, f"Not enough tensor quantizers ({len(tensor_quantizers)}) allocated"
            outputs.append(inner(t, index))
        return outputs

    # ---- export --------------------------------------------------------------------------------------------------
    def export_param_encodings(self):
        return {name: export_quantizer_encoding(q) for name, q in self.param_quantizers.items()}

    def export_output_encodings(self):
        return [export_quantizer_encoding(q) for q in self.output_quantizers]

    def export_input_encodings(self):
        return [export_quantizer_encoding(q) for q in self.input_quantizers]


# names the reference exports for the same class
QcQuantizeWrapper = StaticGridQuantWrapper
QcPostTrainingWrapper = StaticGridQuantWrapper
