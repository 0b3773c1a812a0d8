"""The C-ABI library loads and exports every symbol include/aimet_b200.h declares; host-only helpers are bit-identical
to the oracle. No device compute is called here (this file runs in the CPU-only container)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from tests.conftest import GOLDEN, ROOT
from tests.test_oracle_pin import same_f32

VARIANTS = [(0, 0, 0), (1, 0, 0), (1, 1, 0), (1, 0, 1)]


@pytest.fixture(scope="module")
def lib():
    from aimet_b200 import _build, _lib
    _build.build()
    return _lib.load()


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "aimet_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ab_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported(lib):
    from aimet_b200 import _lib
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/aimet_b200.h but not exported"
        assert n in _lib.PROTOTYPES, f"{n} has no ctypes prototype"
    assert sorted(_lib.PROTOTYPES) == names


def test_state_layout_matches_header(lib):
    from aimet_b200.state import STATE_DTYPE
    assert lib.ab_stats_state_bytes() == STATE_DTYPE.itemsize == 8288
    assert STATE_DTYPE.itemsize % 16 == 0


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from aimet_b200 import _build, _lib
    monkeypatch.setattr(_lib, "_LIB", None)
    monkeypatch.setattr(_build, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(ImportError, match="no CPU or pure-PyTorch fallback"):
        _lib.load()


def test_device_entry_points_fail_without_a_device(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    from aimet_b200 import _lib
    assert lib.ab_device_count() == 0
    buf = (C.c_float * 8)()
    rc = lib.ab_qdq_per_tensor_fwd(C.addressof(buf), C.addressof(buf), 8, 0, -1.0, 1.0, 8, 0, 0, None)
    assert rc == _lib.AB_ERR_CUDA
    assert b"CUDA error" in lib.ab_last_error()


def enc_tuple(e):
    return (e.min, e.max, e.delta, e.offset, e.bw)


def test_host_encoding_math_matches_oracle(lib, oracle):
    from aimet_b200 import _lib
    rng = np.random.default_rng(3)
    for _ in range(400):
        bw = int(rng.choice([2, 4, 8, 12, 16, 31]))
        mn = float(rng.standard_normal() * 10 ** rng.uniform(-6, 6))
        mx = float(rng.standard_normal() * 10 ** rng.uniform(-6, 6))
        if rng.random() < 0.2:
            mn = -mx
        if rng.random() < 0.1:
            mn = 0.0
        e = _lib.Encoding()
        assert lib.ab_fill_encoding_info(bw, mn, mx, C.byref(e)) == 0
        assert enc_tuple(e) == oracle.fill_encoding_info(bw, mn, mx)
        lo, hi = min(mn, mx), max(mn, mx)
        for (s, st, u) in VARIANTS:
            assert lib.ab_tf_compute_encoding(bw, lo, hi, s, st, u, C.byref(e)) == 0
            assert enc_tuple(e) == oracle.tf_encoding(bw, lo, hi, s, st, u)
    for (lo, hi) in ((-np.inf, np.inf), (0.0, np.inf), (-np.inf, 0.0), (-3.4e38, 3.4e38)):
        e = _lib.Encoding()
        lib.ab_tf_compute_encoding(8, lo, hi, 0, 0, 0, C.byref(e))
        assert enc_tuple(e) == oracle.tf_encoding(8, lo, hi)


def test_host_partial_encoding_matches_golden(lib):
    from aimet_b200 import _lib
    g = np.load(os.path.join(GOLDEN, "partial.npz"))
    for inp, out in zip(g["inputs"], g["outputs"]):
        bw, s, u, st, mn, mx, delta, offset = inp
        e = _lib.Encoding(mn, mx, delta, offset, int(bw))
        rc = lib.ab_compute_partial_encoding(int(bw), C.byref(e), int(s), int(u), int(st))
        assert (rc != 0) == (int(out[0]) != 0), inp
        if rc == 0:
            assert enc_tuple(e)[:4] == tuple(out[1:5]), inp


def test_host_per_channel_params_match_oracle(lib, oracle):
    rng = np.random.default_rng(9)
    for bw in (4, 8, 16):
        for sym0 in (False, True):
            c = 37
            mins = -np.abs(rng.standard_normal(c)) * 3
            maxs = np.abs(rng.standard_normal(c)) * 3
            mins[1], maxs[1] = 0.4, 0.4
            if sym0:
                mins[0] = -maxs[0]
            out = np.empty(4 * c, np.float32)
            dp = C.POINTER(C.c_double)
            assert lib.ab_per_channel_params(mins.ctypes.data_as(dp), maxs.ctypes.data_as(dp), c, bw,
                                             out.ctypes.data_as(C.POINTER(C.c_float))) == 0
            exp = np.concatenate(oracle.per_channel_prepare(mins, maxs, bw))
            assert same_f32(out, exp)


def test_kernel_math_headers_match_oracle_on_host(hostmath, oracle):
    """tfe_math.h / encoding_math.h are what the CUDA kernels execute; compiled for the host they must agree with the
    oracle bit for bit: histogram range, xLeft, and the full grid search."""
    from oracle.bindings import OracleMse, OraclePercentile, OracleTfe
    rng = np.random.default_rng(13)
    dp = C.POINTER(C.c_double)
    for t in range(120):
        first = (rng.standard_normal(int(rng.integers(2, 6000))) * rng.uniform(1e-3, 100) + rng.uniform(-4, 4)).astype(
            np.float32)
        if t % 3 == 0:
            first = np.maximum(first, 0)
        a = OracleTfe(oracle)
        a.update(first)
        for _ in range(int(rng.integers(0, 3))):
            a.update((first * np.float32(rng.uniform(0.2, 1.5))).astype(np.float32))
        mn, mx = oracle.get_min_max(first)
        x0, bd, bf, of = C.c_double(), C.c_double(), C.c_float(), C.c_float()
        hostmath.ht_init_pdf_range(mn, mx, C.byref(x0), C.byref(bd), C.byref(bf), C.byref(of))
        xl, pdf = a.histogram()
        mine = np.array([hostmath.ht_x_left(x0.value, bd.value, i) for i in range(512)])
        assert np.array_equal(mine, xl)
        assert (bf.value, of.value) == a.bucket_params()
        for bw in (4, 8, 16):
            for (s, st, u) in VARIANTS:
                out = np.zeros(5)
                hostmath.ht_tfe_compute(pdf.ctypes.data_as(dp), x0.value, bd.value, 1, 1, bw, s, st, u,
                                        out.ctypes.data_as(dp))
                assert tuple(out[:4]) == a.compute(bw, s, st, u)[:4], (t, bw, s, st, u)
        if t % 4 == 0:
            # the percentile and MSE closings (percentile_math.h, mse_math.h) on the same statistics
            pct, mse = OraclePercentile(oracle, 99.0 + (t % 7) * 0.1), OracleMse(oracle)
            pct.s, mse.s = a.s, a.s
            for bw in (4, 8):
                for (s, st, u) in VARIANTS:
                    out = np.zeros(5)
                    hostmath.ht_percentile_compute(pdf.ctypes.data_as(dp), x0.value, bd.value, pct.percentile, bw, s, st, u,
                                                   out.ctypes.data_as(dp))
                    assert tuple(out[:4]) == pct.compute(bw, s, st, u)[:4], ("percentile", t, bw, s, st, u)
                    hostmath.ht_mse_compute(pdf.ctypes.data_as(dp), x0.value, bd.value, bw, s, st, u,
                                            out.ctypes.data_as(dp))
                    assert tuple(out[:4]) == mse.compute(bw, s, st, u)[:4], ("mse", t, bw, s, st, u)
    out = np.zeros(5)
    z = np.zeros(512)
    hostmath.ht_tfe_compute(z.ctypes.data_as(dp), 0.0, 0.0, 0, 1, 8, 0, 0, 0, out.ctypes.data_as(dp))
    a = OracleTfe(oracle)
    a.update(np.zeros(10, np.float32))
    assert tuple(out[:4]) == a.compute(8)[:4]
