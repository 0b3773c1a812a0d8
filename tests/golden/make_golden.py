"""Generates tests/golden/*.npz from the REFERENCE ITSELF (its unmodified C++ compiled into oracle/_ref by
oracle/Makefile). Run in the build container, where /root/reference exists:

    make -C oracle ref && python tests/golden/make_golden.py

The fixtures are small seeded input/output vectors; they travel with the repo so that the oracle (and through it the
CUDA path) stays pinned to the reference on machines that do not have the reference checkout.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.bindings import (QUANTIZATION_TF, QUANTIZATION_TF_ENHANCED, RefAnalyzer, Reference)  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
VARIANTS = [(0, 0, 0), (1, 0, 0), (1, 1, 0), (1, 0, 1)]   # (symmetric, strict, unsigned)


def make_tensor(rng, n, kind):
    x = rng.standard_normal(n).astype(np.float32)
    if kind == "normal":
        return x
    if kind == "shifted":           # the reference's own test distribution N(2, 2)
        return (x * 2 + 2).astype(np.float32)
    if kind == "relu":
        return np.maximum(x, 0).astype(np.float32)
    if kind == "positive":
        return (np.abs(x) + 0.5).astype(np.float32)
    if kind == "tiny":
        return (x * 1e-4).astype(np.float32)
    if kind == "wide":
        return (x * 1e4 - 300).astype(np.float32)
    if kind == "special":
        x[::97] = np.nan
        x[1::193] = np.inf
        x[2::211] = -np.inf
        x[3::89] = 0.0
        x[4::101] = -0.0
        return x
    raise ValueError(kind)


def main():
    ref = Reference()
    rng = np.random.default_rng(20261018)

    # ---- quantize-dequantize / quantize-only (TensorQuantizationSim) ----
    qdq = {}
    k = 0
    for kind in ("normal", "shifted", "relu", "tiny", "wide", "special"):
        for bw in (4, 8, 16):
            x = make_tensor(rng, 1201, kind)
            finite = x[np.isfinite(x)]
            lo, hi = float(finite.min()), float(finite.max())
            ranges = [(lo, hi), (lo * 0.5, hi * 0.5), (-max(abs(lo), abs(hi)), max(abs(lo), abs(hi))), (0.3, 0.3)]
            for (mn, mx) in ranges:
                qdq[f"x{k}"] = x
                qdq[f"meta{k}"] = np.array([mn, mx, bw], dtype=np.float64)
                qdq[f"qdq{k}"] = ref.qdq(x, mn, mx, bw)
                qdq[f"grid_u{k}"] = ref.quantize(x, mn, mx, bw, False)
                qdq[f"grid_s{k}"] = ref.quantize(x, mn, mx, bw, True)
                qdq[f"enc{k}"] = np.array(ref.fill_encoding_info(bw, mn, mx), dtype=np.float64)
                k += 1
    qdq["count"] = np.array(k)
    np.savez_compressed(os.path.join(HERE, "qdq.npz"), **qdq)

    # ---- per-channel kernel (parameters supplied, as the reference's L1 entry point takes them) ----
    pc = {}
    k = 0
    for (c, per) in ((8, 27), (16, 64), (5, 1), (3, 1000), (64, 9)):
        for bw in (4, 8):
            x = (rng.standard_normal(c * per) * rng.uniform(0.1, 3)).astype(np.float32)
            mx = np.abs(x.reshape(c, per)).max(axis=1).astype(np.float32) + 1e-3
            mn = (-mx * rng.uniform(0.2, 1.0, c)).astype(np.float32)
            steps = np.float32(2 ** bw - 1)
            delta = ((mx - mn) / steps).astype(np.float32)
            offset = np.rint(mn / delta).astype(np.float32)
            pc[f"x{k}"] = x
            pc[f"geom{k}"] = np.array([c, per, bw])
            pc[f"params{k}"] = np.stack([mn, mx, delta, offset])
            pc[f"out{k}"] = ref.qdq_per_channel(x, c, per, mn, mx, delta, offset)
            k += 1
    pc["count"] = np.array(k)
    np.savez_compressed(os.path.join(HERE, "per_channel.npz"), **pc)

    # ---- analyzers: batches -> histogram / encodings ----
    an = {}
    k = 0
    for kind in ("normal", "shifted", "relu", "positive", "tiny", "wide", "special"):
        for nbatch in (1, 3):
            batches = [make_tensor(rng, int(rng.integers(500, 5000)), kind) * np.float32(rng.uniform(0.5, 2))
                       for _ in range(nbatch)]
            tfe = RefAnalyzer(ref, QUANTIZATION_TF_ENHANCED)
            tf = RefAnalyzer(ref, QUANTIZATION_TF)
            for b in batches:
                tfe.update(b)
                tf.update(b)
            an[f"nbatch{k}"] = np.array(nbatch)
            for i, b in enumerate(batches):
                an[f"batch{k}_{i}"] = b
            xl, pdf = tfe.histogram()
            an[f"xleft{k}"], an[f"pdf{k}"] = xl, pdf
            encs_tfe, encs_tf = [], []
            for bw in (4, 8, 16):
                for (s, st, u) in VARIANTS:
                    encs_tfe.append(tfe.compute(bw, s, st, u))
                    encs_tf.append(tf.compute(bw, s, st, u))
            an[f"tfe{k}"] = np.array(encs_tfe, dtype=np.float64)
            an[f"tf{k}"] = np.array(encs_tf, dtype=np.float64)
            k += 1
    # zeros first, then data (the PDF must initialise from the first NON-zero batch); and zeros only
    tfe = RefAnalyzer(ref, QUANTIZATION_TF_ENHANCED)
    z = np.zeros(100, np.float32)
    d = make_tensor(rng, 2000, "shifted")
    tfe.update(z)
    tfe.update(d)
    an["zero_then_data_batch"] = d
    an["zero_then_data_pdf"] = tfe.histogram()[1]
    an["zero_then_data_enc"] = np.array(tfe.compute(8, 0, 0, 0))
    tfe = RefAnalyzer(ref, QUANTIZATION_TF_ENHANCED)
    tfe.update(z)
    an["zeros_only_enc"] = np.array([tfe.compute(bw, 0, 0, 0) for bw in (4, 8, 16)])
    an["count"] = np.array(k)
    np.savez_compressed(os.path.join(HERE, "analyzers.npz"), **an)

    # ---- partial encodings ----
    pe_in, pe_out = [], []
    for bw in (4, 8, 16):
        for (s, st, u) in VARIANTS:
            for (mn, mx, delta, offset) in ((-1.0, 2.0, 0, 0), (0.0, 3.5, 0, 0), (-2.5, 0.0, 0, 0), (-3.0, 3.0, 0, 0),
                                            (0, 0, 0.01, -128), (0, 0, 0.02, 0), (0, 0, 0.5, -7), (0, 0, 1e-9, -3),
                                            (-1.0, 1.0, 0.1, -10)):
                rc, enc = ref.partial_encoding(bw, (mn, mx, delta, offset, bw), s, u, st)
                pe_in.append([bw, s, u, st, mn, mx, delta, offset])
                pe_out.append([rc, *enc])
    np.savez_compressed(os.path.join(HERE, "partial.npz"), inputs=np.array(pe_in, dtype=np.float64),
                        outputs=np.array(pe_out, dtype=np.float64))
    # ---- the reference fixture's own data4 (TestTensorQuantizer.cpp:92-103) and what the reference computes from it ----
    import ctypes as C
    data4 = np.empty(6000, np.float32)
    ref.L.ref_kat_normal.argtypes = [C.c_uint, C.c_float, C.c_float, C.c_uint, C.POINTER(C.c_float)]
    ref.L.ref_kat_normal(1, 2.0, 2.0, 6000, data4.ctypes.data_as(C.POINTER(C.c_float)))
    tfe = RefAnalyzer(ref, QUANTIZATION_TF_ENHANCED)
    tfe.update(data4)
    enc = tfe.compute(8, 0, 0, 0)
    five = np.full(16, 5.0, np.float32)
    np.savez_compressed(os.path.join(HERE, "kat_n22.npz"), data4=data4, enc=np.array(enc),
                        qdq5=ref.qdq(five, enc[0], enc[1], 8))
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main()
