"""Seeded batches for the entropy-scheme checks, shared by the golden generator (reference C++) and the tests."""
import numpy as np

VARIANTS = [(8, 0, 0, 0), (8, 1, 0, 0), (8, 1, 1, 0), (8, 1, 0, 1), (4, 0, 0, 0), (16, 1, 0, 0)]   # bw, sym, strict, unsigned


def batches(case: int):
    """A list of float32 arrays: ranges that grow, shrink, all-zero and constant tensors, one-sided data."""
    rng = np.random.default_rng(1000 + case)
    out = []
    for _ in range(int(rng.integers(1, 6))):
        n = int(rng.integers(1, 30000))
        kind = int(rng.integers(0, 7))
        x = (rng.standard_normal(n) * rng.uniform(0.05, 6) + rng.uniform(-3, 3)).astype(np.float32)
        if kind == 0:
            x = np.abs(x)
        elif kind == 1:
            x = np.zeros(n, np.float32)
        elif kind == 2:
            x = np.full(n, np.float32(rng.uniform(-2, 2)), np.float32)
        elif kind == 3:
            x = -np.abs(x)
        elif kind == 4:
            x = (x * 100).astype(np.float32)
        out.append(x)
    return out


NUM_CASES = 24
