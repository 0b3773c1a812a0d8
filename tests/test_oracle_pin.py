"""Pins the CPU oracle (oracle/qsim_oracle.c) to the reference: against the reference's own known answers, against the
committed golden fixtures generated from the reference's unmodified C++ (tests/golden/make_golden.py), and -- where
oracle/_ref was built -- against that C++ run live on fresh seeded inputs. Bit-exact everywhere.
"""
import os

import numpy as np
import pytest

from oracle.bindings import (QUANTIZATION_TF, QUANTIZATION_TF_ENHANCED, OracleTf, OracleTfe, RefAnalyzer)
from tests import kat
from tests.conftest import GOLDEN

VARIANTS = [(0, 0, 0), (1, 0, 0), (1, 1, 0), (1, 0, 1)]


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def same_f32(a, b):
    """bit-identical, except that any NaN equals any NaN"""
    a, b = np.asarray(a, np.float32), np.asarray(b, np.float32)
    return bool(np.all((bits(a) == bits(b)) | (np.isnan(a) & np.isnan(b))))


def ulp_diff(a, b):
    a = bits(a).astype(np.int64)
    b = bits(b).astype(np.int64)
    a = np.where(a < 0x80000000, a, 0x80000000 - a)
    b = np.where(b < 0x80000000, b, 0x80000000 - b)
    return np.abs(a - b)


# ---- reference known answers ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("mn,mx,bw,expected", kat.QDQ_KATS)
def test_qdq_known_answers(oracle, mn, mx, bw, expected):
    out = oracle.qdq(kat.SIX, mn, mx, bw)
    assert ulp_diff(out, np.array(expected, np.float32)).max() <= 4   # EXPECT_FLOAT_EQ


@pytest.mark.parametrize("mn,mx,bw,signed,expected", kat.GRID_KATS)
def test_grid_known_answers(oracle, mn, mx, bw, signed, expected):
    assert oracle.quantize(kat.SIX, mn, mx, bw, signed).tolist() == expected


def test_tfe_all_zero_known_answer(oracle):
    a = OracleTfe(oracle)
    a.update(np.zeros(100, np.float32))
    mn, mx, delta, offset, bw = a.compute(8)
    k = kat.TFE_ALL_ZERO
    assert abs(mn - k["min"]) < k["tol"] and abs(mx - k["max"]) < k["tol"] and offset == k["offset"] and bw == 8


def test_tfe_reference_fixture_known_answer(oracle):
    d = np.load(os.path.join(GOLDEN, "kat_n22.npz"))
    a = OracleTfe(oracle)
    a.update(d["data4"])
    enc = a.compute(8)
    k = kat.TFE_N22
    assert abs(enc[0] - (-6.52711)) < 0.001 and abs(enc[1] - 8.88412) < 0.001     # the reference test's own bar
    assert np.float32(enc[0]) == np.float32(k["min"]) and np.float32(enc[1]) == np.float32(k["max"])
    assert np.float32(enc[2]) == np.float32(k["delta"]) and enc[3] == k["offset"]
    assert tuple(enc) == tuple(d["enc"][:4]) + (8,)
    q = oracle.qdq(np.full(16, 5.0, np.float32), enc[0], enc[1], 8)
    assert same_f32(q, d["qdq5"]) and abs(q[0] - 5.0162) < 0.001


# ---- golden fixtures -----------------------------------------------------------------------------------------------
def test_golden_qdq(oracle):
    g = np.load(os.path.join(GOLDEN, "qdq.npz"))
    for k in range(int(g["count"])):
        mn, mx, bw = g[f"meta{k}"]
        x = g[f"x{k}"]
        assert same_f32(oracle.qdq(x, mn, mx, int(bw)), g[f"qdq{k}"]), k
        assert same_f32(oracle.quantize(x, mn, mx, int(bw), False), g[f"grid_u{k}"]), k
        assert same_f32(oracle.quantize(x, mn, mx, int(bw), True), g[f"grid_s{k}"]), k
        assert oracle.fill_encoding_info(int(bw), mn, mx)[:4] == tuple(g[f"enc{k}"][:4]), k


def test_golden_per_channel(oracle):
    g = np.load(os.path.join(GOLDEN, "per_channel.npz"))
    for k in range(int(g["count"])):
        c, per, _ = g[f"geom{k}"]
        p = g[f"params{k}"]
        out = oracle.qdq_per_channel(g[f"x{k}"], int(c), int(per), *[np.ascontiguousarray(r) for r in p])
        assert same_f32(out, g[f"out{k}"]), k


def test_golden_analyzers(oracle):
    g = np.load(os.path.join(GOLDEN, "analyzers.npz"))
    for k in range(int(g["count"])):
        tfe, tf = OracleTfe(oracle), OracleTf(oracle)
        for i in range(int(g[f"nbatch{k}"])):
            tfe.update(g[f"batch{k}_{i}"])
            tf.update(g[f"batch{k}_{i}"])
        xl, pdf = tfe.histogram()
        assert np.array_equal(xl, g[f"xleft{k}"]) and np.array_equal(pdf, g[f"pdf{k}"]), k
        j = 0
        for bw in (4, 8, 16):
            for (s, st, u) in VARIANTS:
                assert tfe.compute(bw, s, st, u) == tuple(g[f"tfe{k}"][j][:4]) + (bw,), (k, bw, s, st, u)
                assert tf.compute(bw, s, st, u) == tuple(g[f"tf{k}"][j][:4]) + (bw,), (k, bw, s, st, u)
                j += 1
    a = OracleTfe(oracle)
    a.update(np.zeros(100, np.float32))
    a.update(g["zero_then_data_batch"])
    assert np.array_equal(a.histogram()[1], g["zero_then_data_pdf"])
    assert a.compute(8)[:4] == tuple(g["zero_then_data_enc"][:4])
    a = OracleTfe(oracle)
    a.update(np.zeros(100, np.float32))
    for i, bw in enumerate((4, 8, 16)):
        assert a.compute(bw)[:4] == tuple(g["zeros_only_enc"][i][:4])


def test_golden_partial_encodings(oracle):
    g = np.load(os.path.join(GOLDEN, "partial.npz"))
    for inp, out in zip(g["inputs"], g["outputs"]):
        bw, s, u, st, mn, mx, delta, offset = inp
        rc, enc = oracle.partial_encoding(int(bw), (mn, mx, delta, offset, int(bw)), int(s), int(u), int(st))
        assert rc == int(out[0])
        if rc == 0:
            assert enc[:4] == tuple(out[1:5]), inp


# ---- live against the compiled reference ---------------------------------------------------------------------------
def test_live_qdq_and_grid(oracle, reference):
    rng = np.random.default_rng(7)
    for t in range(60):
        x = (rng.standard_normal(4099) * rng.uniform(1e-3, 1e3) + rng.uniform(-3, 3)).astype(np.float32)
        if t % 5 == 0:
            x[::50] = np.nan
            x[1::77] = np.inf
        bw = int(rng.choice([2, 4, 8, 12, 16]))
        mn, mx = float(np.nanmin(x[np.isfinite(x)])) * rng.uniform(0.3, 1.2), float(
            np.nanmax(x[np.isfinite(x)])) * rng.uniform(0.3, 1.2)
        if t % 7 == 0:
            mn = -mx
        assert same_f32(oracle.qdq(x, mn, mx, bw), reference.qdq(x, mn, mx, bw))
        for signed in (False, True):
            assert same_f32(oracle.quantize(x, mn, mx, bw, signed), reference.quantize(x, mn, mx, bw, signed))
        assert oracle.fill_encoding_info(bw, mn, mx) == reference.fill_encoding_info(bw, mn, mx)


def test_live_extreme_values(oracle, reference):
    # TEt/test/python/test_tensor_quantizer.py:418-446 exercises +-3.4e38
    x = np.array([3.4e38, -3.4e38, 1e-45, -1e-45, 0.0, -0.0, 1.0, np.nan, np.inf, -np.inf], np.float32)
    for (mn, mx) in ((-3.4e38, 3.4e38), (-1e-30, 1e-30), (0.0, 3.4e38), (-1.0, 1.0)):
        for bw in (4, 8, 16):
            assert same_f32(oracle.qdq(x, mn, mx, bw), reference.qdq(x, mn, mx, bw)), (mn, mx, bw)


def test_live_analyzers(oracle, reference):
    rng = np.random.default_rng(11)
    for t in range(80):
        ra, oa = RefAnalyzer(reference, QUANTIZATION_TF_ENHANCED), OracleTfe(oracle)
        rt, ot = RefAnalyzer(reference, QUANTIZATION_TF), OracleTf(oracle)
        for _ in range(int(rng.integers(1, 4))):
            d = (rng.standard_normal(int(rng.integers(1, 9000))) * rng.uniform(0.01, 50) + rng.uniform(-5, 5)).astype(
                np.float32)
            if t % 4 == 0:
                d = np.maximum(d, 0)
            if t % 9 == 0:
                d[::13] = np.nan
            for a in (ra, oa, rt, ot):
                a.update(d)
        h_ref, h_or = ra.histogram(), oa.histogram()
        assert np.array_equal(h_ref[0], h_or[0]) and np.array_equal(h_ref[1], h_or[1])
        for bw in (4, 8, 16):
            for (s, st, u) in VARIANTS:
                assert ra.compute(bw, s, st, u) == oa.compute(bw, s, st, u), (t, bw, s, st, u)
                assert rt.compute(bw, s, st, u) == ot.compute(bw, s, st, u), (t, bw, s, st, u)


def test_per_channel_prepare_matches_torch_cpu_ops(oracle):
    """qo_per_channel_prepare restates torch fp32 CPU ops (AimetTensorQuantizer.cpp:236-299): check against torch."""
    import torch
    rng = np.random.default_rng(5)
    for bw in (4, 8, 16):
        for sym0 in (False, True):
            mins = -np.abs(rng.standard_normal(64)) * rng.uniform(0.01, 10)
            maxs = np.abs(rng.standard_normal(64)) * rng.uniform(0.01, 10)
            mins[3], maxs[3] = 0.5, 0.5          # gated
            mins[4], maxs[4] = 0.2, 0.9          # min > 0
            if sym0:
                mins[0] = -maxs[0]
            o_min, o_max, o_delta, o_offset = oracle.per_channel_prepare(mins, maxs, bw)
            t = torch.tensor(np.stack([mins, maxs]).astype(np.float32))
            e_min, e_max = t[0], t[1]
            steps = 2.0 ** bw - 1 - (1 if mins[0] == -maxs[0] else 0)
            zero = torch.zeros(1)
            e_min = torch.minimum(e_min, zero)
            e_max = torch.maximum(e_max, zero)
            e_max = torch.maximum(e_max, e_min + 1e-5)
            delta = (e_max - e_min) / steps
            offset = torch.round(e_min / delta)
            assert same_f32(o_min, e_min.numpy()) and same_f32(o_max, e_max.numpy())
            assert same_f32(o_delta, delta.numpy()) and same_f32(o_offset, offset.numpy())


def test_ste_matches_torch_expression(oracle):
    import torch
    rng = np.random.default_rng(2)
    x = (rng.standard_normal(5000) * 2).astype(np.float32)
    g = rng.standard_normal(5000).astype(np.float32)
    g[::17] = np.inf
    x[::19] = np.nan
    mn, mx = np.float32(-1.3), np.float32(0.77)
    xt, gt = torch.from_numpy(x), torch.from_numpy(g)
    mask = (torch.tensor(float(mn)) <= xt).logical_and(xt <= torch.tensor(float(mx)))
    assert same_f32(oracle.ste_bwd(x, g, mn, mx), (gt * mask).numpy())
