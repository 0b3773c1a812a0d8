"""Golden vectors for quantizeTensorPacked, generated from the reference's own C++ (oracle/_ref/libaimet_ref.so).
    python tests/golden/make_packed_golden.py     (build container only)"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle.bindings import Reference, build  # noqa: E402

build(with_ref=True)
ref = Reference()
rng = np.random.default_rng(11)
x = (rng.standard_normal(1031) * 2).astype(np.float32)
x[::97], x[1::101], x[2::103], x[3::107], x[4::109] = np.nan, np.inf, -np.inf, 0.0, -0.0
out = {"x": x}
cases = []
for bw in (1, 2, 4, 8, 16, 32):
    for signed in (0, 1):
        for mn, mx in ((-3.0, 5.0), (-4.0, 4.0), (0.0, 6.0), (-1e-3, 1e-3)):
            key = f"bw{bw}_s{signed}_{len(cases)}"
            out[key] = ref.quantize_packed(x, mn, mx, bw, bool(signed))
            cases.append((key, bw, signed, mn, mx))
out["cases"] = np.array([(k, str(b), str(s), repr(mn), repr(mx)) for k, b, s, mn, mx in cases])
path = os.path.join(os.environ.get("GOLDEN_OUT", os.path.dirname(os.path.abspath(__file__))), "packed.npz")
np.savez_compressed(path, **out)
print(path, len(cases), "cases")
