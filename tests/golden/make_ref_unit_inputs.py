"""Inputs of the reference's own unit-test fixtures: std::normal_distribution<float>(mean, stddev) driven by
std::mt19937(seed) (libstdc++-specific, hence generated through oracle/_ref's ref_kat_normal, which runs exactly that loop)
-> tests/golden/ref_unit_inputs.npz.     python tests/golden/make_ref_unit_inputs.py

    n22_seed1     DlQ/test/TestTensorQuantizer.cpp:88-103, TestTfEnhancedEncodingAnalyzer.cpp:96-112
    n22_seed10    DlQ/test/TestTfEncodingAnalyzer.cpp:62-75
    n22_seed100   DlQ/test/TestTfEncodingAnalyzer.cpp:104-117 (and the three cases after it)
    nm21_seed1    DlQ/test/TestTfEnhancedEncodingAnalyzer.cpp:199-211 (N(-2, 1)); the Percentile / Mse analyzer tests too
    nm12_seed1    DlQ/test/TestPercentileEncodingAnalyzer.cpp:262-285 (N(-1, 2))
    n22_seed1_100k  DlQ/test/TestPercentileEncodingAnalyzer.cpp:56-86, TestMseEncodingAnalyzer.cpp:56-86 (100 000 samples)
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import bindings  # noqa: E402

bindings.build(with_ref=True)
L = C.CDLL(bindings.REF_SO)
L.ref_kat_normal.argtypes = [C.c_uint, C.c_float, C.c_float, C.c_uint, C.POINTER(C.c_float)]


def normal(seed, mean, stddev, n=6000):
    out = np.empty(n, np.float32)
    L.ref_kat_normal(seed, mean, stddev, n, out.ctypes.data_as(C.POINTER(C.c_float)))
    return out


out = {"n22_seed1": normal(1, 2, 2), "n22_seed10": normal(10, 2, 2), "n22_seed100": normal(100, 2, 2),
       "nm21_seed1": normal(1, -2, 1), "nm12_seed1": normal(1, -1, 2), "n22_seed1_100k": normal(1, 2, 2, 100000)}
assert np.array_equal(out["n22_seed1_100k"][:6000], out["n22_seed1"])
assert np.array_equal(out["n22_seed1"], np.load(os.path.join(ROOT, "tests", "golden", "kat_n22.npz"))["data4"])
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "ref_unit_inputs.npz"), **out)
print({k: (float(v.min()), float(v.max())) for k, v in out.items()})
