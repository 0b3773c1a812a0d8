"""GPU parity of the percentile scheme: statistics through ab_stats_update (mode AB_QUANTIZATION_PERCENTILE) and the
encoding through ab_compute_encodings_percentile, against the goldens generated from the reference's own
PercentileEncodingAnalyzer and against the C oracle -- encodings BIT-EXACT."""
import os

import numpy as np
import pytest
import torch

from oracle import bindings
from tests.conftest import GOLDEN
from tests.golden.make_percentile_cases import ANALYZER_CASES, analyzer_batches

pytestmark = pytest.mark.gpu
PERCENTILE = 3


@pytest.fixture(scope="module")
def ops():
    from aimet_b200 import ops as o
    return o


def device_encodings(ops, batches, pct, variant, dtype=torch.float32):
    from aimet_b200.state import StateArena
    blk = StateArena.for_device(torch.device("cuda", 0)).allocate(1)
    for b in batches:
        ops.stats_update_impl(torch.from_numpy(b).cuda().to(dtype), blk.arena, blk.first, PERCENTILE, None, 0)
    bw, sym, strict, unsigned = variant
    enc, qdq4 = ops.compute_encodings_impl(blk.arena, blk.first, 1, PERCENTILE, bw, sym, strict, unsigned,
                                           want_qdq4=True, percentile=pct)
    return enc[0].cpu().numpy(), qdq4[0].cpu().numpy()


@pytest.mark.parametrize("name", list(ANALYZER_CASES))
def test_device_matches_reference_goldens(ops, name):
    gold = np.load(os.path.join(GOLDEN, "percentile.npz"))[name]
    spec = ANALYZER_CASES[name]
    batches = analyzer_batches(name)
    row = 0
    for pct in spec["percentiles"]:
        for variant in spec["variants"]:
            enc, qdq4 = device_encodings(ops, batches, pct, variant)
            assert np.array_equal(enc, gold[row]), (name, pct, variant, enc, gold[row])
            if gold[row][4] != 0:
                e = ops.fill_encoding_info(variant[0], gold[row][0], gold[row][1])
                assert np.array_equal(qdq4, np.array([e.min, e.max, e.delta, e.offset], dtype=np.float32))
            row += 1


def test_batched_records_and_bf16(ops, oracle):
    """Several records in one launch, each with its own statistics; bf16 inputs are widened exactly."""
    from aimet_b200.state import StateArena
    rng = np.random.default_rng(3)
    n_rec = 37
    blk = StateArena.for_device(torch.device("cuda", 0)).allocate(n_rec)
    ports = []
    for i in range(n_rec):
        a = bindings.OraclePercentile(oracle, 99.5)
        for _ in range(1 + i % 3):
            x = (rng.standard_normal(3000 + 17 * i) * (0.5 + i % 5) + (i % 4 - 1)).astype(np.float32)
            xb = torch.from_numpy(x).to(torch.bfloat16)
            ops.stats_update_impl(xb.cuda(), blk.arena, blk.first + i, PERCENTILE, None, 0)
            a.update(xb.float().numpy())
        ports.append(a)
    for (bw, sym, strict, unsigned) in ((8, 0, 0, 0), (8, 1, 0, 0), (4, 1, 1, 0)):
        enc, _ = ops.compute_encodings_impl(blk.arena, blk.first, n_rec, PERCENTILE, bw, sym, strict, unsigned,
                                            percentile=99.5)
        enc = enc.cpu().numpy()
        for i, a in enumerate(ports):
            assert tuple(enc[i]) == tuple(float(v) for v in a.compute(bw, sym, strict, unsigned)), i


def test_python_api(ops):
    """AimetTensorQuantizer / libpymo.TensorQuantizer with the percentile scheme (AimetTensorQuantizer.cpp:200-207,
    TensorQuantizer.cpp:239-256)."""
    from aimet_b200 import AimetTensorQuantizer, libpymo
    x = torch.randn(50000, device="cuda") * 2
    x[::4000] = 60.0
    q = AimetTensorQuantizer(libpymo.QuantizationMode.QUANTIZATION_PERCENTILE)
    q.updateStats(x, True)
    full, ok = q.getEncoding(8, False, False, False)
    assert ok and full.max > 50
    q.setPercentileValue(99.0)
    clipped, _ = q.getEncoding(8, False, False, False)
    assert clipped.max < 10 < full.max
    q.resetEncodingStats()                       # a new analyzer: back to 100
    q.updateStats(x, True)
    again, _ = q.getEncoding(8, False, False, False)
    assert (again.min, again.max) == (full.min, full.max)
    tq = libpymo.TensorQuantizer(libpymo.QuantizationMode.QUANTIZATION_PERCENTILE, libpymo.RoundingMode.ROUND_NEAREST)
    assert tq.getPercentileValue() == 100.0
    tq.setPercentileValue(99.0)
    tq.updateStats(x.cpu().numpy(), False)
    e = tq.computeEncoding(8, False)
    assert (e.min, e.max) == (clipped.min, clipped.max)
    tf = libpymo.TensorQuantizer(libpymo.QuantizationMode.QUANTIZATION_TF, libpymo.RoundingMode.ROUND_NEAREST)
    with pytest.raises(RuntimeError):
        tf.getPercentileValue()
    with pytest.raises(ValueError):              # the plain entry point refuses the percentile mode
        from aimet_b200.state import StateArena
        blk = StateArena.for_device(torch.device("cuda", 0)).allocate(1)
        ops.compute_encodings_impl(blk.arena, blk.first, 1, PERCENTILE, 8, False, False, False)


# ---- MSE scheme ------------------------------------------------------------------------------------------------------
MSE = 4


@pytest.mark.parametrize("name", list(ANALYZER_CASES))
def test_mse_device_matches_reference_goldens(ops, name):
    from aimet_b200.state import StateArena
    gold = np.load(os.path.join(GOLDEN, "mse.npz"))[name]
    batches = analyzer_batches(name)
    blk = StateArena.for_device(torch.device("cuda", 0)).allocate(1)
    for b in batches:
        ops.stats_update_impl(torch.from_numpy(b).cuda(), blk.arena, blk.first, MSE, None, 0)
    for row, (bw, sym, strict, unsigned) in enumerate(ANALYZER_CASES[name]["variants"]):
        enc, _ = ops.compute_encodings_impl(blk.arena, blk.first, 1, MSE, bw, sym, strict, unsigned)
        assert np.array_equal(enc[0].cpu().numpy(), gold[row]), (name, bw, sym, strict, unsigned)


def test_mse_batched_records_and_python_api(ops, oracle):
    from aimet_b200 import AimetTensorQuantizer, libpymo
    from aimet_b200.state import StateArena
    rng = np.random.default_rng(5)
    n_rec = 19
    blk = StateArena.for_device(torch.device("cuda", 0)).allocate(n_rec)
    ports = []
    for i in range(n_rec):
        a = bindings.OracleMse(oracle)
        for _ in range(1 + i % 2):
            x = (rng.standard_normal(2000 + 31 * i) * (0.3 + i % 4) + (i % 3 - 1)).astype(np.float32)
            if i % 5 == 0:
                x = np.maximum(x, 0)
            ops.stats_update_impl(torch.from_numpy(x).cuda(), blk.arena, blk.first + i, MSE, None, 0)
            a.update(x)
        ports.append(a)
    for (bw, sym, strict, unsigned) in ((8, 0, 0, 0), (8, 1, 0, 0), (4, 1, 1, 0), (8, 1, 0, 1)):
        enc, _ = ops.compute_encodings_impl(blk.arena, blk.first, n_rec, MSE, bw, sym, strict, unsigned)
        enc = enc.cpu().numpy()
        for i, a in enumerate(ports):
            assert tuple(enc[i]) == tuple(float(v) for v in a.compute(bw, sym, strict, unsigned)), (i, bw, sym)
    # the class-level drop-ins
    x = torch.randn(30000, device="cuda") * 2 + 0.5
    q = AimetTensorQuantizer(libpymo.QuantizationMode.QUANTIZATION_MSE)
    q.updateStats(x, True)
    enc, ok = q.getEncoding(8, False, False, False)
    port = bindings.OracleMse(oracle)
    port.update(x.cpu().numpy())
    assert ok and (enc.min, enc.max, enc.delta, enc.offset, enc.bw) == port.compute(8)
    assert len(q.getStatsHistogram()) == 512
    tq = libpymo.TensorQuantizer(libpymo.QuantizationMode.QUANTIZATION_MSE, libpymo.RoundingMode.ROUND_NEAREST)
    tq.updateStats(x.cpu().numpy(), False)
    e = tq.computeEncoding(8, False)
    assert (e.min, e.max) == (enc.min, enc.max)
