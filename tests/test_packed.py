"""quantizeTensorPacked (DlQ/src/TensorQuantizationSim.cpp:128-139, trim_functions.cpp:221-388): the C restatement against
the reference's known answer, the golden vectors generated from the reference's C++, and the reference live; the CUDA kernel
against the restatement (bytes identical)."""
import os

import numpy as np
import pytest

from tests.conftest import GOLDEN


def cases():
    g = np.load(os.path.join(GOLDEN, "packed.npz"))
    return g, [(k, int(b), bool(int(s)), float(mn), float(mx)) for k, b, s, mn, mx in g["cases"]]


def test_oracle_packed_reference_known_answer(oracle):
    # DlQ/test/TestTensorQuantizationSim.cpp:240-262
    x = np.array([-0.5, -0.25, 0, 0.25, 0.5, 0.75], dtype=np.float32)
    assert oracle.quantize_packed(x, -0.46, 0.72, 8, False).tolist() == [0, 45, 99, 153, 207, 255]


def test_oracle_packed_matches_reference_goldens(oracle):
    g, cs = cases()
    assert len(cs) == 48
    for key, bw, signed, mn, mx in cs:
        assert np.array_equal(oracle.quantize_packed(g["x"], mn, mx, bw, signed), g[key]), key
    assert oracle.quantize_packed(g["x"], -1.0, 1.0, 3, False) is None          # the reference throws


def test_oracle_packed_matches_reference_live(oracle, reference):
    rng = np.random.default_rng(5)
    for trial in range(6):
        x = (rng.standard_normal(20011) * rng.uniform(0.1, 30)).astype(np.float32)
        mn, mx = -abs(rng.normal()) * 4, abs(rng.normal()) * 4 + 0.01
        for bw in (1, 2, 4, 8, 16, 32):
            for signed in (False, True):
                assert np.array_equal(oracle.quantize_packed(x, mn, mx, bw, signed),
                                      reference.quantize_packed(x, mn, mx, bw, signed)), (trial, bw, signed)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", ["float32", "bfloat16"])
def test_cuda_packed_is_bytewise_identical(oracle, dtype):
    import torch
    from aimet_b200 import ops
    g, cs = cases()
    td = getattr(torch, dtype)
    for n, offset in ((1031, 0), (1030, 1), (5, 0)):                 # aligned, misaligned base pointer, shorter than a vector
        base = torch.from_numpy(g["x"].copy()).cuda().to(td)
        x = base[offset:offset + n]
        xh = x.float().cpu().numpy()
        for key, bw, signed, mn, mx in cs:
            got = ops.quantize_to_packed_impl(x, mn, mx, bw, signed).cpu().numpy()
            assert np.array_equal(got, oracle.quantize_packed(xh, mn, mx, bw, signed)), (key, n, offset)
    big = (torch.randn(3_000_017, device="cuda") * 3).to(td)
    for bw, signed in ((4, False), (8, True), (16, False), (32, True)):
        got = ops.quantize_to_packed_impl(big, -2.5, 7.0, bw, signed).cpu().numpy()
        assert np.array_equal(got, oracle.quantize_packed(big.float().cpu().numpy(), -2.5, 7.0, bw, signed)), (bw, signed)
    view = ops.quantize_to_packed_impl(big[:1000], -2.5, 7.0, 16, True).view(torch.int16)
    assert view.shape == (1000,) and int(view.min()) >= -32768
    with pytest.raises(ValueError):
        ops.quantize_to_packed_impl(big[:16], -1.0, 1.0, 3, False)
