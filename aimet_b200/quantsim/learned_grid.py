"""Host mirror of the reference's range-learning ("learned grid") quantizer and wrapper.

Reference: TrainingExtensions/torch/src/python/aimet_torch/v1/
  tensor_quantizer.py  LearnedGridTensorQuantizer :573-851, QuantizeDequantizeFunc :854-963,
                       initialize_learned_grid_quantizer_attributes :1285-1344,
                       set_encoding_min_max_gating_threshold :1347-1359
  qc_quantize_op.py    LearnedGridQuantWrapper :947-1158, _patch_param :1369-1404
  quantsim.py          _construct_and_initialize_trainable_wrapper :786-831

Same class, attribute, parameter (`<name>_encoding_min` / `<name>_encoding_max`) and method names. Underneath, one
quantize-dequantize is ONE kernel forward (gating of the two parameters included) and ONE kernel backward
(aimet_b200.ops.LearnedGridQdq over ab_lg_qdq_fwd / ab_lg_qdq_bwd) instead of ~18 + ~14 torch launches, and only the
input tensor is kept for the backward. The function can be swapped with `set_qdq_function` (the CPU test-suite injects
the oracle that way); there is no CPU implementation in the product.
"""
import math
from typing import List, Optional, Tuple, Union

import torch
from torch import nn

from .. import libpymo, ops
from .defs import MAP_ROUND_MODE_TO_PYMO, QuantizationDataType, QuantScheme
from .qc_quantize_op import EncodingImportMixin

_IGNORED_DTYPES = (torch.int, torch.int8, torch.int16, torch.int32, torch.int64, torch.bool, torch.uint8)


def _fused_qdq(tensor, encoding_min, encoding_max, quantizer, gate):
    mode = ops.lg_symmetry_mode(quantizer.use_symmetric_encodings, quantizer.is_unsigned_symmetric)
    return ops.LearnedGridQdq.apply(tensor, encoding_min, encoding_max, quantizer.bitwidth, mode,
                                    quantizer.use_strict_symmetric, quantizer.channel_axis, gate)


_QDQ_FUNCTION = _fused_qdq


def set_qdq_function(fn):
    """Replace the quantize-dequantize function `(tensor, enc_min, enc_max, quantizer, gate) -> tensor` (test hook).
    Returns the previous one."""
    global _QDQ_FUNCTION
    prev, _QDQ_FUNCTION = _QDQ_FUNCTION, (fn or _fused_qdq)
    return prev


def is_non_strict_symmetric(use_symmetric_encodings, use_strict_symmetric, is_unsigned_symmetric) -> bool:
    """reference aimet_common/quantsim.py:84-96"""
    return use_symmetric_encodings and not use_strict_symmetric and not is_unsigned_symmetric


def set_encoding_min_max_gating_threshold(encoding_min: nn.Parameter, encoding_max: nn.Parameter):
    """reference v1/tensor_quantizer.py:1347-1359 (the fused forward does the same on the device with gate=True)."""
    zero = torch.zeros((), dtype=encoding_min.dtype, device=encoding_min.device)
    eps = torch.tensor(1e-5, dtype=encoding_min.dtype, device=encoding_min.device)
    with torch.no_grad():
        encoding_min.clamp_(max=zero)
        encoding_max.clamp_(min=zero)
        encoding_max.clamp_(min=encoding_min.data + eps)


def get_computed_encodings(bitwidth, encoding_min, encoding_max, use_symmetric_encodings, use_strict_symmetric,
                           is_unsigned_symmetric):
    """(delta, offset, num_steps) as small tensors, reference quantsim_straight_through_grad.py:121-160. Used for
    export / inspection only: the kernels derive the same values on the device."""
    num_steps = 2 ** bitwidth - 1
    if use_symmetric_encodings and use_strict_symmetric:
        num_steps -= 1
    like = dict(dtype=encoding_min.dtype, device=encoding_min.device)
    steps = torch.tensor(num_steps, **like)
    if use_symmetric_encodings and not is_unsigned_symmetric:
        delta = encoding_max / torch.tensor(math.floor(num_steps / 2), **like)
        offset = -torch.tensor(math.ceil(num_steps / 2), **like)
    else:
        delta = (encoding_max - encoding_min) / steps
        if use_symmetric_encodings:
            offset = encoding_min / delta
        else:
            b_zero = torch.round(-encoding_min / delta)
            b_zero = torch.min(steps, torch.max(torch.zeros((), **like), b_zero))
            offset = -b_zero
    return delta, offset, steps


class LearnedGridTensorQuantizer:
    """Quantizer whose (min, max) are trainable parameters of its wrapper (reference :573-851)."""

    def __init__(self, bitwidth: int, round_mode, quant_scheme: QuantScheme, use_symmetric_encodings: bool,
                 enabled_by_default: bool, data_type: QuantizationDataType = QuantizationDataType.int):
        if data_type != QuantizationDataType.int:
            raise ValueError("Only QuantizationDataType.int is supported for LearnedGridTensorQuantizer")
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
        self.wrapper_ref = None
        self.name = None
        self.device = None
        self._ch_axis = 0

    # ---- properties ----------------------------------------------------------------------------------------------
    @property
    def quant_scheme(self):
        return self._quant_scheme

    @property
    def is_encoding_frozen(self) -> bool:
        return self._is_encoding_frozen

    @property
    def channel_axis(self) -> int:
        return self._ch_axis

    @property
    def encoding_min_max_fixed_vals(self) -> Optional[Tuple[float, float]]:
        return self._encoding_min_max_fixed_vals

    @encoding_min_max_fixed_vals.setter
    def encoding_min_max_fixed_vals(self, min_max_vals):
        self._encoding_min_max_fixed_vals = min_max_vals

    def _params(self):
        return (getattr(self.wrapper_ref, self.name + "_encoding_min"),
                getattr(self.wrapper_ref, self.name + "_encoding_max"))

    @property
    def encoding(self) -> Union[None, libpymo.TfEncoding, List[libpymo.TfEncoding]]:
        """The up-to-date (learned) encoding, computed from the two parameters (reference :699-716, :803-849)."""
        if not self.enabled or self.bitwidth == 32 or self.data_type == QuantizationDataType.float:
            return None
        return self._compute_updated_encoding()

    @encoding.setter
    def encoding(self, encoding):
        if not self.enabled or self.bitwidth == 32 or self.data_type == QuantizationDataType.float:
            return
        if encoding is None:
            raise RuntimeError("Encodings cannot be None if Quantizer is enabled.")
        bitwidth = encoding[0].bw if isinstance(encoding, list) else encoding.bw
        if bitwidth != self.bitwidth:
            raise RuntimeError(f"Bitwidth mismatched. The bitwidth for quantizer is {self.bitwidth}, but the bitwidth "
                               f"in encodings is {bitwidth}. If the intent is to change the bitwidth, please set "
                               f"quantizer bitwidth to {bitwidth} first.")
        if self._is_encoding_frozen:
            raise RuntimeError("Encoding can be set only when it is not frozen.")
        self._set_encoding_min_max_parameters(encoding)

    def _set_encoding_min_max_parameters(self, encodings):
        """reference :851-883: the parameters are created as float32 on the wrapper's device."""
        if isinstance(encodings, list):
            mins, maxs = [e.min for e in encodings], [e.max for e in encodings]
        else:
            mins, maxs = [encodings.min], [encodings.max]
        params = self.wrapper_ref._parameters   # pylint: disable=protected-access
        params[self.name + "_encoding_min"] = nn.Parameter(torch.FloatTensor(mins).to(self.wrapper_ref.device),
                                                           requires_grad=True)
        params[self.name + "_encoding_max"] = nn.Parameter(torch.FloatTensor(maxs).to(self.wrapper_ref.device),
                                                           requires_grad=True)

    def compute_scaling_offset(self, encoding_min, encoding_max):
        if encoding_min is None or encoding_max is None:
            return None, None
        scaling, offset, _ = get_computed_encodings(self.bitwidth, encoding_min, encoding_max,
                                                    self.use_symmetric_encodings, self.use_strict_symmetric,
                                                    self.is_unsigned_symmetric)
        return scaling, offset

    def _compute_updated_encoding(self):
        encoding_min, encoding_max = self._params()
        if encoding_min is None or encoding_max is None:
            return None
        with torch.no_grad():
            encoding_min, encoding_max = encoding_min.detach().float(), encoding_max.detach().float()
            scale, offset = self.compute_scaling_offset(encoding_min, encoding_max)
            scale, offset = scale.expand_as(encoding_min), offset.expand_as(encoding_min)
            if not self.use_symmetric_encodings or self.is_unsigned_symmetric:
                # zero must stay exactly representable: min / max follow the rounded offset (reference :822-832)
                adjusted_min = scale * offset
                encoding_max = encoding_max - encoding_min + adjusted_min
                encoding_min = adjusted_min
            rows = torch.stack([encoding_min, encoding_max, scale, offset]).cpu().tolist()
        encodings = [libpymo.TfEncoding._from_values(mn, mx, dl, of, self.bitwidth)   # pylint: disable=protected-access
                     for mn, mx, dl, of in zip(*rows)]
        return encodings[0] if len(encodings) == 1 else encodings

    def get_effective_encoding(self):
        """Encoding faithful to the configured scheme: a non-strict symmetric grid has one more bin below (reference
        :640-690)."""
        if not self.enabled:
            return None
        encodings = self.encoding
        if not encodings:
            return None
        if isinstance(encodings, libpymo.TfEncoding):
            encodings = [encodings]
        out = []
        for e in encodings:
            if is_non_strict_symmetric(self.use_symmetric_encodings, self.use_strict_symmetric,
                                       self.is_unsigned_symmetric):
                out.append(libpymo.TfEncoding._from_values(e.min - e.delta, e.max, e.delta, e.offset, e.bw))   # pylint: disable=protected-access
            else:
                out.append(e)
        return out[0] if len(out) == 1 else out

    def quantize_dequantize(self, tensor, encoding_min, encoding_max, gate: bool = False):
        """reference :786-801. `gate=True` additionally clamps the two parameters in place inside the same kernel
        (what LearnedGridQuantWrapper.apply_gating_logic does in the reference before every forward)."""
        if not self.enabled or self.bitwidth == 32:
            return tensor
        if encoding_min is None or encoding_max is None:
            raise RuntimeError("Forward pass used for compute_encodings differs from forward pass used during training")
        if tensor.dtype not in (torch.float32, torch.float16, torch.bfloat16):
            raise RuntimeError("Invalid input data type. Expected torch.float32 or torch.float16. "
                               f"Got {tensor.dtype}.")
        if gate and self._is_encoding_frozen:
            pass   # gating a frozen pair is still what the reference does (clamp_ under no_grad); keep it
        return _QDQ_FUNCTION(tensor, encoding_min, encoding_max, self, gate)

    def freeze_encoding(self):
        params = self.wrapper_ref._parameters   # pylint: disable=protected-access
        mn, mx = params[self.name + "_encoding_min"], params[self.name + "_encoding_max"]
        if mn is None and mx is None:
            raise RuntimeError("Encoding can be frozen only when it is not None.")
        self._is_encoding_frozen = True
        mn.requires_grad = False
        mx.requires_grad = False

    def __str__(self):
        lines = ["LearnedGrid TensorQuantizer:",
                 f"    quant-scheme:{self._quant_scheme}, round_mode={self.round_mode}, bitwidth={self.bitwidth}, "
                 f"enabled={self.enabled}"]
        enc = self.get_effective_encoding() if self.encoding else None
        if enc:
            for e in ([enc] if isinstance(enc, libpymo.TfEncoding) else enc):
                lines.append(f"    min:{e.min}, max={e.max}, delta={e.delta}, offset={e.offset}")
        else:
            lines.append("    no encoding")
        return "\n".join(lines) + "\n"


def initialize_learned_grid_quantizer_attributes(new_quantizer: LearnedGridTensorQuantizer, old_quantizer):
    """reference :1285-1344 (including its in-place symmetrisation of the old quantizer's encodings)."""
    new_quantizer.enabled = old_quantizer.enabled
    new_quantizer.bitwidth = old_quantizer.bitwidth
    new_quantizer.data_type = old_quantizer.data_type
    new_quantizer.use_symmetric_encodings = old_quantizer.use_symmetric_encodings
    new_quantizer.use_strict_symmetric = old_quantizer.use_strict_symmetric
    new_quantizer.use_unsigned_symmetric = old_quantizer.use_unsigned_symmetric
    new_quantizer.is_unsigned_symmetric = False      # range learning never runs the unsigned variant
    new_quantizer.encoding_min_max_fixed_vals = old_quantizer.encoding_min_max_fixed_vals
    new_quantizer.is_const = old_quantizer.is_const
    if new_quantizer.data_type == QuantizationDataType.float or new_quantizer.bitwidth == 32:
        new_quantizer.encoding = None
        return
    old_encoding = old_quantizer.encoding if old_quantizer.enabled else None
    as_list = old_encoding if isinstance(old_encoding, list) else ([old_encoding] if old_encoding else [])
    if old_quantizer.enabled and old_quantizer.use_symmetric_encodings and not old_quantizer.is_unsigned_symmetric:
        for e in as_list:
            e.min = -e.max
    if old_quantizer.enabled and old_quantizer.is_unsigned_symmetric:
        half = (2 ** old_quantizer.bitwidth - 1) / 2
        for e in as_list:
            e.min = -e.max
            e.delta = e.max / math.floor(half)
            e.offset = -math.ceil(half)
    new_quantizer.encoding = old_encoding


class _PatchedParams:
    """`with` scope in which getattr(module, name) returns the quantized tensor (reference _patch_param :1369-1404)."""

    def __init__(self, module):
        self.module = module
        self.undo = []

    def patch(self, name, quantized):
        module = self.module
        original = getattr(module, name)
        assert original.shape == quantized.shape
        if name in module.__dict__:
            self.undo.append(lambda: module.__dict__.update({name: original}))
        else:
            self.undo.append(lambda: module.__dict__.pop(name))
        module.__dict__[name] = quantized

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        for fn in reversed(self.undo):
            fn()
        self.undo = []


class LearnedGridQuantWrapper(EncodingImportMixin, nn.Module):
    """Learns min and max of every enabled quantizer of one wrapped layer (reference qc_quantize_op.py:947-1158)."""

    def __init__(self, module_to_wrap: nn.Module, weight_bw: int, activation_bw: int, round_mode, quant_scheme,
                 device, is_output_quantized: bool = True, is_symmetric: bool = False, num_inputs: int = 1,
                 num_outputs: int = 1, data_type: QuantizationDataType = QuantizationDataType.int):
        super().__init__()
        if data_type != QuantizationDataType.int:
            raise ValueError("Only QuantizationDataType.int is supported for LearnedGridQuantWrapper")
        if isinstance(round_mode, str):
            round_mode = MAP_ROUND_MODE_TO_PYMO[round_mode]
        if isinstance(quant_scheme, str):
            quant_scheme = QuantScheme.from_str(quant_scheme)
        self._module_to_wrap = module_to_wrap
        self._quant_scheme = quant_scheme
        self.device = device
        mk = lambda bw, enabled: LearnedGridTensorQuantizer(bw, round_mode, quant_scheme, is_symmetric, enabled,   # noqa: E731
                                                            data_type)
        self.output_quantizers = [mk(activation_bw, is_output_quantized) for _ in range(num_outputs)]
        self.input_quantizers = [mk(activation_bw, False) for _ in range(num_inputs)]
        self.param_quantizers = {name: mk(weight_bw, True) for name, _ in module_to_wrap.named_parameters()}
        self._initialize_trainable_parameters_and_tensor_quantizers(num_inputs, num_outputs)

    def _initialize_trainable_parameters_and_tensor_quantizers(self, num_inputs, num_outputs):
        def attach(q, name):
            self.register_parameter(name + "_encoding_min", None)
            self.register_parameter(name + "_encoding_max", None)
            q.name, q.wrapper_ref, q.device = name, self, self.device

        for i in range(num_inputs):
            attach(self.input_quantizers[i], f"input{i}")
        for i in range(num_outputs):
            attach(self.output_quantizers[i], f"output{i}")
        for name, param in self.get_named_parameters():
            attach(self.param_quantizers[name], name)
            axis = 0
            if isinstance(self._module_to_wrap, (nn.ConvTranspose1d, nn.ConvTranspose2d, nn.ConvTranspose3d)) and \
                    len(param.shape) > 1:
                axis = 1
            self.param_quantizers[name]._ch_axis = axis   # pylint: disable=protected-access

    # ---- accessors -----------------------------------------------------------------------------------------------
    @property
    def output_quantizer(self):
        return self.output_quantizers[0]

    @property
    def input_quantizer(self):
        return self.input_quantizers[0]

    def get_original_module(self) -> nn.Module:
        return self._module_to_wrap

    def get_named_parameters(self):
        return list(self._module_to_wrap.named_parameters())

    def set_mode(self, mode):
        """Learned-grid wrappers always quantize (the reference keeps the attribute for interface parity)."""
        self._mode = mode

    def _enabled_pairs(self):
        named = [(f"input{i}", q) for i, q in enumerate(self.input_quantizers)] + \
                [(f"output{i}", q) for i, q in enumerate(self.output_quantizers)] + \
                [(name, self.param_quantizers[name]) for name, _ in self._module_to_wrap.named_parameters()]
        for name, q in named:
            if q.enabled and q.bitwidth != 32 and q.data_type != QuantizationDataType.float:
                yield getattr(self, name + "_encoding_min"), getattr(self, name + "_encoding_max")

    def apply_gating_logic(self):
        """reference :1019-1052. forward() does not call this: every fused quantize-dequantize gates its own pair on the
        device. Kept for callers that want the parameters gated without running a forward."""
        for mn, mx in self._enabled_pairs():
            if mn is not None and mx is not None:
                set_encoding_min_max_gating_threshold(mn, mx)

    # ---- forward -------------------------------------------------------------------------------------------------
    def forward(self, *inputs, **kwargs):
        quantized_inputs = self._quantize_activation(list(inputs), self.input_quantizers, "input")
        with self._quantize_params():
            wrapped_output = self._module_to_wrap(*quantized_inputs, **kwargs)
        if not isinstance(wrapped_output, (list, tuple)):
            wrapped_output = [wrapped_output]
        output = self._quantize_activation(list(wrapped_output), self.output_quantizers, "output")
        return output[0] if len(output) == 1 else output

    def _quantize_params(self):
        patched = _PatchedParams(self._module_to_wrap)
        try:
            for name, _ in self.get_named_parameters():
                q = self.param_quantizers[name]
                if not q.enabled:
                    continue
                original = getattr(self._module_to_wrap, name)
                quantized = q.quantize_dequantize(original, getattr(self, name + "_encoding_min"),
                                                  getattr(self, name + "_encoding_max"), gate=True)
                patched.patch(name, quantized)
        except Exception:
            patched.__exit__(None, None, None)
            raise
        return patched

    @staticmethod
    def should_perform_quant_dequant(tensor, tensor_quantizer) -> bool:
        if not isinstance(tensor, torch.Tensor) or tensor.dtype in _IGNORED_DTYPES or \
                (tensor_quantizer.is_const and torch.numel(tensor) == 1) or not tensor_quantizer.enabled:
            tensor_quantizer.enabled = False
            return False
        return True

    def _quantize_activation(self, tensors_to_quantize, tensor_quantizers, type_of_quantizer: str):
        def inner(t, index):
            if isinstance(t, (list, tuple)):
                return [inner(x, index) for x in t]
            q = tensor_quantizers[index]
            if not self.should_perform_quant_dequant(t, q):
                return t
            return q.quantize_dequantize(t, getattr(self, f"{type_of_quantizer}{index}_encoding_min"),
                                         getattr(self, f"{type_of_quantizer}{index}_encoding_max"), gate=True)

        out = []
        for index, t in enumerate(tensors_to_quantize):
            assert len(tensor_quantizers) > index, f"Not enough tensor quantizers ({len(tensor_quantizers)}) allocated"
            out.append(inner(t, index))
        return out

    # ---- export --------------------------------------------------------------------------------------------------
    def export_param_encodings(self):
        from .qc_quantize_op import export_quantizer_encoding
        return {name: export_quantizer_encoding(q) for name, q in self.param_quantizers.items()}

    def export_output_encodings(self):
        from .qc_quantize_op import export_quantizer_encoding
        return [export_quantizer_encoding(q) for q in self.output_quantizers]

    def export_input_encodings(self):
        from .qc_quantize_op import export_quantizer_encoding
        return [export_quantizer_encoding(q) for q in self.input_quantizers]


def construct_and_initialize_trainable_wrapper(post_training_module, device, default_param_bw, default_output_bw,
                                               rounding_mode, quant_scheme) -> LearnedGridQuantWrapper:
    """reference quantsim.py:786-831"""
    module = post_training_module._module_to_wrap   # pylint: disable=protected-access
    trainable = LearnedGridQuantWrapper(module, default_param_bw, default_output_bw, rounding_mode, quant_scheme,
                                        device=device, num_inputs=len(post_training_module.input_quantizers),
                                        num_outputs=len(post_training_module.output_quantizers))
    pairs = list(zip(trainable.output_quantizers, post_training_module.output_quantizers)) + \
        list(zip(trainable.input_quantizers, post_training_module.input_quantizers)) + \
        [(trainable.param_quantizers[name], q) for name, q in post_training_module.param_quantizers.items()]
    for new_q, old_q in pairs:
        initialize_learned_grid_quantizer_attributes(new_q, old_q)
        if new_q.encoding_min_max_fixed_vals is not None:
            new_q.freeze_encoding()
    return trainable
