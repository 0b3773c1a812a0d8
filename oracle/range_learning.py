"""CPU oracle for the range-learning ("learned grid") QDQ -- TEST INFRASTRUCTURE, never on the product path.

The reference implements this path entirely with torch element-wise ops
(aimet_torch/v1/quantsim_straight_through_grad.py:121-346, v1/tensor_quantizer.py:854-963, 1347-1359), so the
restatement below is torch-on-CPU as well: every operation is written out one at a time in the tensor's own dtype, which
reproduces the reference's rounding (bf16 tensors below 16 bit are processed in bf16). Pinned against the reference's
unmodified functions by tests/golden/range_learning.npz (tests/golden/make_range_learning_golden.py generated it here by
importing the reference) -- see tests/test_range_learning_oracle.py.

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module.
"""
import math

import torch

ASYMMETRIC, SIGNED_SYMMETRIC, UNSIGNED_SYMMETRIC = 0, 1, 2


def symmetry_mode(use_symmetric, is_unsigned_symmetric):
    """The branch get_computed_encodings takes (:145-158)."""
    if use_symmetric and not is_unsigned_symmetric:
        return SIGNED_SYMMETRIC
    return UNSIGNED_SYMMETRIC if use_symmetric else ASYMMETRIC


def gate(enc_min, enc_max):
    """set_encoding_min_max_gating_threshold (v1/tensor_quantizer.py:1347-1359), in place, in the parameters' dtype."""
    zero = torch.zeros((), dtype=enc_min.dtype, device=enc_min.device)   # constant_like: same dtype AND device
    eps = torch.tensor(1e-5, dtype=enc_min.dtype, device=enc_min.device)
    with torch.no_grad():
        enc_min.clamp_(max=zero)
        enc_max.clamp_(min=zero)
        enc_max.clamp_(min=enc_min + eps)


def _grid(enc_min, enc_max, bw, mode, strict):
    """(delta, offset, steps) as tensors of enc_min's dtype (get_computed_encodings, :121-160)."""
    n = 2 ** bw - 1
    if mode != ASYMMETRIC and strict:
        n -= 1
    like = dict(dtype=enc_min.dtype, device=enc_min.device)
    steps = torch.tensor(n, **like)
    if mode == SIGNED_SYMMETRIC:
        delta = enc_max / torch.tensor(math.floor(n / 2), **like)
        offset = -torch.tensor(math.ceil(n / 2), **like)
    else:
        delta = (enc_max - enc_min) / steps
        if mode == UNSIGNED_SYMMETRIC:
            offset = enc_min / delta
        else:
            zero_point = torch.round(-enc_min / delta)
            zero_point = torch.min(steps, torch.max(torch.zeros((), **like), zero_point))
            offset = -zero_point
    return delta, offset, steps


def _view(t, x, ch_axis):
    """broadcast_to_tensor (:70-92)."""
    if t.numel() == 1:
        return t
    shape = [1] * x.dim()
    shape[ch_axis] = x.shape[ch_axis]
    return t.view(shape)


def _arith(x, enc_min, enc_max, bw):
    if x.dtype in (torch.float16, torch.bfloat16) and bw >= 16:   # :211-214
        return x.float(), enc_min.float(), enc_max.float()
    return x, enc_min, enc_max


def forward(x, enc_min, enc_max, bw, mode, strict=False, ch_axis=0):
    """calculate_forward_pass (:183-247): returns (y, saved) where saved feeds backward()."""
    if bw >= 32:
        raise RuntimeError(f"Invalid bitwidth: {bw}")
    out_dtype = x.dtype
    x, enc_min, enc_max = _arith(x, enc_min, enc_max, bw)
    delta, offset, steps = _grid(enc_min, enc_max, bw, mode, strict)
    delta_b, offset_b = _view(delta, x, ch_axis), _view(offset, x, ch_axis)
    zero = torch.zeros_like(steps)
    position = torch.round(x / delta_b) - offset_b
    x_quant = position.clamp(zero, steps)
    y = (x_quant + offset_b) * delta_b
    mask = position.ge(zero) * position.le(steps)
    saved = dict(x=x, x_quant=x_quant, mask=mask, delta=delta_b, offset=offset_b, steps=steps, enc_min=enc_min,
                 enc_max=enc_max, mode=mode, ch_axis=ch_axis)
    return y.to(out_dtype), saved


def backward(grad, saved):
    """QuantizeDequantizeFunc.backward (v1/tensor_quantizer.py:927-963) with asymmetric_gradients / symmetric_gradients
    (:250-330). Returns (grad_x, grad_min, grad_max)."""
    x, x_quant, mask = saved["x"], saved["x_quant"], saved["mask"]
    delta, offset, steps = saved["delta"], saved["offset"], saved["steps"]
    enc_min, enc_max, ch_axis = saved["enc_min"], saved["enc_max"], saved["ch_axis"]
    grad = grad.to(x.dtype)
    grad_x = mask * grad
    per_channel = delta.numel() > 1
    elementwise = per_channel and x.dim() == 1
    dims = [d for d in range(x.dim()) if not (per_channel and d == ch_axis)]

    def total(t):
        return t if elementwise else t.sum(dim=dims)

    if saved["mode"] == ASYMMETRIC:
        to_scale = (x_quant + offset - x * mask / delta) * grad
        to_offset = (delta * grad) * (~mask)
        first = total(to_scale) / steps
        second = steps / (enc_max - enc_min) ** 2 * total(to_offset)
        grad_min = -first + enc_max * second
        grad_max = first - enc_min * second
    else:
        diff = total((x_quant + offset) * grad) - total(mask * (x / delta) * grad)
        grad_max = diff / torch.div(steps, 2, rounding_mode="floor")
        grad_min = -grad_max
    return grad_x, grad_min.view_as(enc_min), grad_max.view_as(enc_max)
