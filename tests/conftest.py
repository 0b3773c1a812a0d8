import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _has_cuda():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:   # pragma: no cover
        return False


def pytest_collection_modifyitems(config, items):
    if _has_cuda():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


_REGENERATED = {}


def sim_golden(filename, generator, generator_args, fingerprint):
    """A whole-model golden (tests/golden/*.json, produced by the reference's unmodified Python on a CPU forward).

    torch's CPU convolutions are not bit-reproducible between hosts (oneDNN picks kernels by ISA and thread count), and
    these goldens contain encodings derived from the forward's last bit. Each golden therefore carries the sha256 of the
    plain fp32 forward on the host that generated it. Same fingerprint here -> the committed file is the expectation.
    Different fingerprint -> the reference's Python is run again on THIS host (it exists only in the build container,
    under /root/reference) into a scratch directory; without the reference checkout the comparison is skipped."""
    import json
    path = os.path.join(GOLDEN, filename)
    gold = json.load(open(path))
    if gold.get("forward_fingerprint") == fingerprint:
        return gold
    if not os.path.isdir("/root/reference"):
        pytest.skip(f"{filename} was generated on a host whose torch CPU forward differs in the last bit, and the "
                    "reference checkout needed to regenerate it is not here")
    key = (generator, tuple(generator_args))
    if key not in _REGENERATED:
        import tempfile
        from oracle import bindings
        bindings.build(with_ref=True)
        out = tempfile.mkdtemp(prefix="aimet_b200_golden_")
        subprocess.run([sys.executable, os.path.join(GOLDEN, generator)] + list(generator_args), check=True,
                       env=dict(os.environ, GOLDEN_OUT=out), stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        _REGENERATED[key] = out
    gold = json.load(open(os.path.join(_REGENERATED[key], filename)))
    assert gold["forward_fingerprint"] == fingerprint, "the CPU forward is not reproducible even on one host"
    return gold


def forward_fingerprint(model, x):
    import hashlib

    import torch
    with torch.no_grad():
        return hashlib.sha256(model(x).numpy().tobytes()).hexdigest()


@pytest.fixture(scope="session")
def oracle():
    """The plain-C restatement (oracle/qsim_oracle.c), built on demand with gcc."""
    from oracle import bindings
    bindings.build(with_ref=os.path.isdir("/root/reference"))
    return bindings.Oracle()


@pytest.fixture(scope="session")
def reference():
    """The reference's own C++ (oracle/_ref/libaimet_ref.so); tests skip where it was not built."""
    from oracle import bindings
    if not os.path.exists(bindings.REF_SO):
        if os.path.isdir("/root/reference"):
            bindings.build(with_ref=True)
        else:
            pytest.skip("oracle/_ref/libaimet_ref.so not built (needs the reference checkout)")
    return bindings.Reference()


@pytest.fixture(scope="session")
def hostmath():
    """tests/native/hostmath_test.cpp: the product's host+device math headers compiled for the host."""
    import ctypes as C
    src = os.path.join(ROOT, "tests", "native", "hostmath_test.cpp")
    so = os.path.join(ROOT, "tests", "native", "libhostmath_test.so")
    deps = [src] + [os.path.join(ROOT, "aimet_b200", "csrc", h) for h in ("encoding_math.h", "tfe_math.h", "percentile_math.h", "mse_math.h")]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.run(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fPIC", "-shared", src, "-o", so],
                       check=True)
    lib = C.CDLL(so)
    dp = C.POINTER(C.c_double)
    fp = C.POINTER(C.c_float)
    lib.ht_tfe_compute.argtypes = [dp, C.c_double, C.c_double, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, dp]
    lib.ht_percentile_compute.argtypes = [dp, C.c_double, C.c_double, C.c_float, C.c_int, C.c_int, C.c_int, C.c_int, dp]
    lib.ht_mse_compute.argtypes = [dp, C.c_double, C.c_double, C.c_int, C.c_int, C.c_int, C.c_int, dp]
    lib.ht_init_pdf_range.argtypes = [C.c_float, C.c_float, dp, dp, fp, fp]
    lib.ht_x_left.restype = C.c_double
    lib.ht_x_left.argtypes = [C.c_double, C.c_double, C.c_int]
    lib.ht_fill_encoding_info.argtypes = [C.c_int, C.c_double, C.c_double, dp]
    return lib
