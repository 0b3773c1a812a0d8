"""Known answers of the reference's own PYTHON unit tests, driven through the host mirror's
quantizer classes exactly as the reference's tests drive the originals:

    TEt/test/python/test_qc_quantize_op.py:177-260           quantize-only integer grids (asymmetric, symmetric signed,
                                                             symmetric unsigned), torch.equal
    TEt/test/python/test_per_channel_quantization.py:67-186  per-channel QDQ (symmetric / asymmetric) and per-channel TF
                                                             encodings from data
    TEt/test/python/test_quantizer.py:1692-1775              straight-through gradient masks (scalar and per-channel range)

Each case runs with the CPU oracle as the native op on CPU tensors (not gpu) and with the CUDA ops on CUDA tensors (gpu).
"""
import pytest
import torch

from aimet_b200 import libpymo
from aimet_b200.quantsim.defs import MAP_ROUND_MODE_TO_PYMO, QuantizationDataType, QuantScheme
from aimet_b200.quantsim.tensor_quantizer import StaticGridPerChannelQuantizer, StaticGridPerTensorQuantizer

NEAREST = MAP_ROUND_MODE_TO_PYMO["nearest"]


@pytest.fixture(params=["oracle", pytest.param("cuda", marks=pytest.mark.gpu)])
def device(request, oracle):
    if request.param == "cuda":
        yield torch.device("cuda")
        return
    from aimet_b200.quantsim import tensor_quantizer
    from tests.oracle_backend import OracleTensorQuantizer
    prev = tensor_quantizer._set_op_class_for_testing(OracleTensorQuantizer)   # pylint: disable=protected-access
    yield torch.device("cpu")
    tensor_quantizer._set_op_class_for_testing(prev)                           # pylint: disable=protected-access


def encoding(mn, mx, offset, delta=None, bw=8):
    e = libpymo.TfEncoding()
    e.bw, e.max, e.min, e.offset = bw, mx, mn, offset
    if delta is not None:
        e.delta = delta
    return e


def per_tensor(sym):
    return StaticGridPerTensorQuantizer(bitwidth=8, round_mode="nearest", quant_scheme=QuantScheme.post_training_tf,
                                        use_symmetric_encodings=sym, enabled_by_default=True,
                                        data_type=QuantizationDataType.int)


def per_channel(sym):
    return StaticGridPerChannelQuantizer(bitwidth=8, round_mode="nearest", quant_scheme=QuantScheme.post_training_tf,
                                         use_symmetric_encodings=sym, enabled_by_default=True, num_channels=4)


# ---- test_qc_quantize_op.py: quantize-only ------------------------------------------------------------------------------
def test_quantize_only_asymmetric(device):                                               # :177-195
    q = per_tensor(False)
    q.encoding = encoding(-5.19, 2.23, -178)
    out = q.quantize(torch.tensor([-7, -5, -3, 0, .1, 2.5], device=device), NEAREST)
    assert torch.equal(out.cpu(), torch.tensor([-128., -122., -53., 50., 53., 127.]))


def test_quantize_only_symmetric_signed(device):                                         # :197-217
    q = per_tensor(True)
    q.encoding = encoding(-5.20, 5.19, -128)
    out = q.quantize(torch.tensor([-7, -5, -3, 0, .1, 2.5], device=device), NEAREST)
    assert torch.equal(out.cpu(), torch.tensor([-128., -123., -74., 0., 2., 61.]))


def test_quantize_only_symmetric_unsigned(device):                                       # :238-260
    q = per_tensor(True)
    q.use_unsigned_symmetric = True
    q.encoding = encoding(0.0, 5.19, 0)
    out = q.quantize(torch.tensor([0, 1.2, 1.5, 4.0, 4.9, 5.3], device=device), NEAREST)
    assert torch.equal(out.cpu(), torch.tensor([0., 59., 74., 197., 241., 255.]))


# ---- test_per_channel_quantization.py -----------------------------------------------------------------------------------
ROWS = [[-7, -5, -3, 0, .1, 2.5]] * 4


def test_per_channel_symmetric_qdq(device):                                              # :67-102
    q = per_channel(True)
    q.encoding = [encoding(-3.84, 3.81, -128, 0.03)] * 3 + [encoding(-6.4, 6.35, -128, 0.05)]
    out = q.quantize_dequantize(torch.tensor(ROWS, dtype=torch.float32, device=device), NEAREST)
    expected = torch.tensor([[-3.84, -3.84, -3, 0, .089999996, 2.49]] * 3 + [[-6.4, -5, -3, 0, .1, 2.5]])
    assert torch.allclose(out.cpu(), expected, atol=1e-5)


def test_per_channel_asymmetric_qdq(device):                                             # :104-139
    q = per_channel(False)
    q.encoding = [encoding(-2.9999934, 1.9999956, -153, 0.0196078)] * 3 + [encoding(-5.995262, 2.404693, -182, 0.032941)]
    out = q.quantize_dequantize(torch.tensor(ROWS, dtype=torch.float32, device=device), NEAREST)
    expected = torch.tensor([[-3.0, -3.0, -3.0, 0, .098, 2.0]] * 3 + [[-5.9953, -5.0070, -2.9976, 0, .09888, 2.4047]])
    assert torch.allclose(out.cpu(), expected, atol=0.0001)


DATA = [[-7, -5, -3, 0, .1, 2.5], [-5, -5, -3, 0, .1, 2.7], [-6, -5, -3, 0, .1, 2.8], [-5, -5, -3, 0, .1, 2]]


def test_per_channel_symmetric_compute_encodings(device):                                # :141-162
    q = per_channel(True)
    q.update_encoding_stats(torch.tensor(DATA, device=device))
    q.compute_encoding()
    assert len(q.encoding) == 4
    assert q.encoding[0].max == 7 and round(q.encoding[0].min, 2) == -7.06
    assert q.encoding[3].max == 5 and round(q.encoding[3].min, 2) == -5.04


def test_per_channel_asymmetric_compute_encodings(device):                               # :164-186
    q = per_channel(False)
    q.update_encoding_stats(torch.tensor(DATA, device=device))
    q.compute_encoding()
    assert len(q.encoding) == 4
    assert round(q.encoding[0].max, 3) == 2.496 and round(q.encoding[0].min, 3) == -7.004
    assert round(q.encoding[3].max, 3) == 2.004 and round(q.encoding[3].min, 3) == -4.996


# ---- test_quantizer.py: straight-through gradient -----------------------------------------------------------------------
STE_CASES = [                                                                            # (input, expected gradient)
    ([[1.0, 1.5], [0.125, -0.12]], [[1.0, 0.0], [1.0, 1.0]]),                            # input > max
    ([[1.0, 0.5], [0.125, -0.30]], [[1.0, 1.0], [1.0, 0.0]]),                            # input < min
    ([[1.0, 0.5], [0.125, -0.25]], [[1.0, 1.0], [1.0, 1.0]]),                            # both bounds are inclusive
]


@pytest.mark.parametrize("x,expected", STE_CASES)
def test_ste_gradient_math_oracle(oracle, x, expected):                                  # :1692-1719, :1750-1775 (oracle)
    import numpy as np
    x = np.array(x, np.float32)
    grad = np.ones_like(x)
    assert oracle.ste_bwd(x.reshape(-1), grad.reshape(-1), -0.25, 1.0).reshape(2, 2).tolist() == expected
    per_ch = oracle.ste_bwd_per_channel(x.reshape(-1), grad.reshape(-1), 2, 2, np.array([-0.25, -0.25], np.float32),
                                        np.array([1.0, 1.0], np.float32))
    assert per_ch.reshape(2, 2).tolist() == expected


@pytest.mark.gpu
@pytest.mark.parametrize("x,expected", STE_CASES)
def test_ste_gradient_math_cuda(x, expected):                                            # :1721-1748
    from aimet_b200.quantsim.tensor_quantizer import compute_dloss_by_dx
    x = torch.tensor(x, device="cuda")
    grad = torch.ones(2, 2, device="cuda")
    assert torch.equal(compute_dloss_by_dx(x, grad, -0.25, 1.0).cpu(), torch.tensor(expected))
    assert torch.equal(compute_dloss_by_dx(x, grad, [-0.25, -0.25], [1.0, 1.0]).cpu(), torch.tensor(expected))
