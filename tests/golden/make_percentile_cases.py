"""Cases and seeded inputs of the percentile goldens (shared by the generator and the tests; no reference imports)."""
import numpy as np

VARIANTS = [(8, 0, 0, 0), (8, 1, 0, 0), (8, 1, 1, 0), (8, 1, 0, 1), (4, 0, 0, 0), (16, 1, 0, 0)]
ANALYZER_CASES = {
    "normal": dict(percentiles=[100.0, 99.99, 99.9, 99.0, 95.0, 90.0], variants=VARIANTS),
    "shifted": dict(percentiles=[100.0, 99.9, 99.0, 90.0], variants=VARIANTS),
    "relu": dict(percentiles=[100.0, 99.9, 99.0, 90.0], variants=VARIANTS),
    "outliers": dict(percentiles=[100.0, 99.99, 99.9, 99.0], variants=VARIANTS),
    "zeros_first": dict(percentiles=[99.9, 99.0], variants=VARIANTS),
    "single_bin": dict(percentiles=[99.9, 50.0], variants=VARIANTS[:2]),
    "all_zero": dict(percentiles=[99.0], variants=VARIANTS),
}


def analyzer_batches(name):
    rng = np.random.default_rng(abs(hash(name)) % 1000 if False else sum(map(ord, name)))
    mk = lambda n: rng.standard_normal(n).astype(np.float32)   # noqa: E731
    if name == "normal":
        return [mk(20000), mk(20000) * 1.5, mk(8000)]
    if name == "shifted":
        return [mk(20000) * 2 + 2, mk(20000) * 2 + 2]
    if name == "relu":
        return [np.maximum(mk(30000), 0), np.maximum(mk(30000) * 2, 0)]
    if name == "outliers":
        a = mk(50000)
        a[::5000] = 40.0
        a[1::7000] = -25.0
        return [a, mk(50000)]
    if name == "zeros_first":
        return [np.zeros(1000, np.float32), mk(10000), mk(10000) * 3]
    if name == "single_bin":
        a = np.full(5000, 0.37, np.float32)
        a[0], a[1] = -1.0, 2.0
        return [a]
    if name == "all_zero":
        return [np.zeros(4096, np.float32)]
    raise KeyError(name)
