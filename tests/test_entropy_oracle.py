"""The entropy scheme's oracle (qo_entropy_* in oracle/qsim_oracle.c) pinned to the reference's own C++: against the committed
golden (tests/golden/entropy.npz, from oracle/_ref) and, where oracle/_ref is present, live on fresh seeded batches -- raw
histogram, range and iteration count and the encodings of every variant, bit for bit. The reference's tests hold no known
answer for this analyzer (DlQ/test has none for EntropyEncodingAnalyzer), so the compiled reference is the pin."""
import os

import numpy as np
import pytest

from oracle.bindings import OracleEntropy
from tests.conftest import GOLDEN
from tests.golden.make_entropy_cases import NUM_CASES, VARIANTS, batches


@pytest.mark.parametrize("case", range(NUM_CASES))
def test_oracle_reproduces_reference_golden(oracle, case):
    gold = np.load(os.path.join(GOLDEN, "entropy.npz"))
    a = OracleEntropy(oracle)
    for x in batches(case):
        a.update(x)
    hist, mn, mx, it = a.raw()
    want = gold[f"c{case}.hist"]
    if want.size == 0:
        assert hist is None
    else:
        assert np.array_equal(hist, want)
        assert [mn, mx, it] == gold[f"c{case}.range"].tolist()
    enc = np.array([a.compute(bw, s, st, u) for bw, s, st, u in VARIANTS], dtype=np.float64)
    assert np.array_equal(enc, gold[f"c{case}.enc"])


def test_oracle_equals_reference_live(oracle, reference):
    from oracle.bindings import RefAnalyzer, RefTensorHistogram
    rng = np.random.default_rng(5)
    for trial in range(30):
        a, r, t = OracleEntropy(oracle), RefAnalyzer(reference, 5), RefTensorHistogram(reference)
        for x in batches(500 + trial):
            x = x * np.float32(rng.uniform(0.5, 2))
            a.update(x), r.update(x), t.update(x)
        h1, h2 = a.raw(), t.raw()
        assert (h1[0] is None) == (h2[0] is None)
        if h1[0] is not None:
            assert np.array_equal(h1[0], h2[0]) and h1[1:] == h2[1:]
        for bw, s, st, u in VARIANTS:
            assert tuple(a.compute(bw, s, st, u)) == tuple(r.compute(bw, s, st, u)), (trial, bw, s, st, u)


def test_no_statistics_and_all_zero(oracle):
    a = OracleEntropy(oracle)
    assert a.compute(8) == (0.0, 0.0, 0.0, 0.0, 0)                       # never updated: the failure indicator
    a.update(np.zeros(100, np.float32))
    mn, mx, delta, offset, bw = a.compute(8)                            # only zeros seen: a valid encoding around 0
    assert bw == 8 and mn < 0 < mx and delta == 2 / 255 and offset == np.floor(-1 / delta)
