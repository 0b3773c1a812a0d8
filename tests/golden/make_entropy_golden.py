"""tests/golden/entropy.npz: the reference's own C++ (oracle/_ref: updateTensorHistogram + EntropyEncodingAnalyzer<float>) on the
seeded batches of make_entropy_cases.py -- raw histogram, min, max, iterations and the encodings of six variants per case.
    python tests/golden/make_entropy_golden.py      (build container only)"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle.bindings import RefAnalyzer, Reference, RefTensorHistogram  # noqa: E402
from tests.golden.make_entropy_cases import NUM_CASES, VARIANTS, batches  # noqa: E402

ref = Reference()
out = {}
for case in range(NUM_CASES):
    a, t = RefAnalyzer(ref, 5), RefTensorHistogram(ref)
    for x in batches(case):
        a.update(x)
        t.update(x)
    hist, mn, mx, it = t.raw()
    out[f"c{case}.hist"] = np.zeros(0) if hist is None else hist
    out[f"c{case}.range"] = np.array([mn, mx, it], dtype=np.float64)
    out[f"c{case}.enc"] = np.array([a.compute(bw, s, st, u) for bw, s, st, u in VARIANTS], dtype=np.float64)
np.savez_compressed(os.path.join(os.environ.get("GOLDEN_OUT", HERE), "entropy.npz"), **out)
print("cases", NUM_CASES)
