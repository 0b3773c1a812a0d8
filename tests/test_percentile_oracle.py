"""Pin the percentile restatement (oracle/qsim_oracle.c qo_percentile_compute) to the reference's own
PercentileEncodingAnalyzer: against tests/golden/percentile.npz (generated from oracle/_ref by
tests/golden/make_percentile_golden.py) and, where oracle/_ref is present, against the reference live on fresh inputs."""
import os

import numpy as np
import pytest

from oracle import bindings
from tests.conftest import GOLDEN
from tests.golden.make_percentile_cases import ANALYZER_CASES, analyzer_batches


@pytest.mark.parametrize("name", list(ANALYZER_CASES))
def test_port_matches_golden(oracle, name):
    gold = np.load(os.path.join(GOLDEN, "percentile.npz"))[name]
    spec = ANALYZER_CASES[name]
    batches = analyzer_batches(name)
    row = 0
    for pct in spec["percentiles"]:
        for (bw, sym, strict, unsigned) in spec["variants"]:
            a = bindings.OraclePercentile(oracle, pct)
            for b in batches:
                a.update(b)
            got = np.array(a.compute(bw, sym, strict, unsigned), dtype=np.float64)
            assert np.array_equal(got, gold[row]), (name, pct, bw, sym, strict, unsigned, got, gold[row])
            row += 1
    assert row == len(gold)


def test_port_matches_reference_live(oracle, reference):
    rng = np.random.default_rng(11)
    for trial in range(12):
        batches = [rng.standard_normal(4000).astype(np.float32) * (1 + trial % 4) + (trial % 3) for _ in range(2)]
        for pct in (100.0, 99.95, 99.0, 91.5):
            for (bw, sym, strict, unsigned) in ((8, 0, 0, 0), (8, 1, 0, 0), (4, 1, 1, 0), (8, 1, 0, 1)):
                a = bindings.OraclePercentile(oracle, pct)
                r = bindings.RefAnalyzer(reference, 3)
                r.set_percentile(pct)
                for b in batches:
                    a.update(b)
                    r.update(b)
                assert a.compute(bw, sym, strict, unsigned) == tuple(r.compute(bw, sym, strict, unsigned))
    # no statistics at all: the zero encoding
    assert bindings.OraclePercentile(oracle, 99.0).compute(8) == tuple(bindings.RefAnalyzer(reference, 3).compute(8))


# ---- MSE analyzer (QuantizationMode 4): same statistics, least-MSE (min, max) among the bin edges ------------------------
@pytest.mark.parametrize("name", list(ANALYZER_CASES))
def test_mse_port_matches_golden(oracle, name):
    gold = np.load(os.path.join(GOLDEN, "mse.npz"))[name]
    batches = analyzer_batches(name)
    for row, (bw, sym, strict, unsigned) in enumerate(ANALYZER_CASES[name]["variants"]):
        a = bindings.OracleMse(oracle)
        for b in batches:
            a.update(b)
        got = np.array(a.compute(bw, sym, strict, unsigned), dtype=np.float64)
        assert np.array_equal(got, gold[row]), (name, bw, sym, strict, unsigned, got, gold[row])


def test_mse_port_matches_reference_live(oracle, reference):
    rng = np.random.default_rng(12)
    for trial in range(10):
        batches = [rng.standard_normal(3000).astype(np.float32) * (1 + trial % 4) + (trial % 3) for _ in range(2)]
        if trial % 5 == 4:
            batches = [np.abs(b) for b in batches]
        for (bw, sym, strict, unsigned) in ((8, 0, 0, 0), (8, 1, 0, 0), (4, 1, 1, 0), (8, 1, 0, 1)):
            a = bindings.OracleMse(oracle)
            r = bindings.RefAnalyzer(reference, 4)
            for b in batches:
                a.update(b)
                r.update(b)
            assert a.compute(bw, sym, strict, unsigned) == tuple(r.compute(bw, sym, strict, unsigned))
    assert bindings.OracleMse(oracle).compute(8) == tuple(bindings.RefAnalyzer(reference, 4).compute(8))
