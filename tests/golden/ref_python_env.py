"""Import the reference's unmodified Python (aimet_common / aimet_torch) in this container, with empty stubs for the
modules the snapshot lacks and a stand-in for the two native modules backed by the reference's own C++ (oracle/_ref)."""
import importlib.machinery
import sys
import types

import numpy as np
import torch
import torchvision  # noqa: F401 (before the stubs go in)

REF = "/root/reference/TrainingExtensions"
sys.path.insert(0, REF + "/common/src/python")
sys.path.insert(0, REF + "/torch/src/python")
sys.path.insert(0, "/root/repo")

from oracle.bindings import Reference, RefAnalyzer  # noqa: E402


class _Dummy(torch.nn.Module):
    pass


def _stub(name):
    m = types.ModuleType(name)
    m.__spec__ = importlib.machinery.ModuleSpec(name, None)
    m.__path__ = []
    cache = {}

    def _ga(attr):
        if attr.startswith("__"):
            raise AttributeError(attr)
        if attr not in cache:
            cache[attr] = type(attr, (_Dummy,), {})   # a distinct placeholder class per name
        return cache[attr]
    m.__getattr__ = _ga
    sys.modules[name] = m
    return m


for n in ["bokeh", "bokeh.server", "bokeh.server.server", "bokeh.application", "spconv", "spconv.pytorch", "onnx",
          "onnxsim", "torch.onnx.symbolic_caffe2", "aimet_torch.v1.nn", "aimet_torch.v1.nn.modules",
          "aimet_torch.v1.nn.modules.custom", "aimet_torch.v2.experimental", "aimet_torch.v2.nn",
          "aimet_torch.v2.nn.fake_quant", "aimet_torch.v2.quantization", "aimet_torch.v2.quantsim",
          "aimet_torch.v2.visualization_tools"]:
    _stub(n)

for full in list(sys.modules):
    if "." in full:
        parent, child = full.rsplit(".", 1)
        if parent in sys.modules and isinstance(sys.modules[parent], types.ModuleType) and \
                getattr(sys.modules[full], "__getattr__", None) is not None and child not in ("symbolic_caffe2",):
            try:
                setattr(sys.modules[parent], child, sys.modules[full])
            except Exception:
                pass

# ---- native stand-ins over the REAL reference C++ ----
import enum  # noqa: E402

_REF = Reference()


class QuantizationMode(enum.IntEnum):
    QUANTIZATION_TF = 0
    QUANTIZATION_TF_ENHANCED = 1
    QUANTIZATION_RANGE_LEARNING = 2
    QUANTIZATION_PERCENTILE = 3
    QUANTIZATION_MSE = 4
    QUANTIZATION_ENTROPY = 5


class RoundingMode(enum.IntEnum):
    ROUND_NEAREST = 0
    ROUND_STOCHASTIC = 1


class TfEncoding:
    """Stand-in for the pybind11 struct (PyModelOptimizations.cpp:172-178): double min/max/delta/offset, int bw. Like
    pybind11's def_readwrite, assignment converts (a 0-dim tensor becomes a Python float)."""

    def __init__(self):
        self.__dict__.update(min=0.0, max=0.0, delta=0.0, offset=0.0, bw=0)

    def __setattr__(self, key, value):
        if key not in ("min", "max", "delta", "offset", "bw"):
            raise AttributeError(key)
        self.__dict__[key] = int(value) if key == "bw" else float(value)


import aimet_common.py_libpymo as _py  # noqa: E402  (the reference's own pure-python enum definitions)

pymo = types.ModuleType("aimet_common._libpymo")
for _k in dir(_py):
    if not _k.startswith("_"):
        setattr(pymo, _k, getattr(_py, _k))
pymo.TfEncoding = TfEncoding
sys.modules["aimet_common._libpymo"] = pymo

CALLS = {"updateStats": 0, "getEncoding": 0, "quantizeDequantize": 0, "quantizeDequantizePerChannel": 0, "objects": 0}


class AimetTensorQuantizer:
    def __init__(self, scheme):
        self.scheme = int(getattr(scheme, 'value', scheme))
        self.a = RefAnalyzer(_REF, self.scheme)
        self.valid = False
        CALLS["objects"] += 1

    def resetEncodingStats(self):
        self.valid = False
        self.a = RefAnalyzer(_REF, self.scheme)

    def updateStats(self, t, use_cuda):
        CALLS["updateStats"] += 1
        self.valid = True
        self.a.update(t.detach().contiguous().numpy().astype(np.float32).reshape(-1))

    def getEncoding(self, bw, sym, strict, unsigned):
        CALLS["getEncoding"] += 1
        e = TfEncoding()
        if self.valid:
            e.min, e.max, e.delta, e.offset, e.bw = self.a.compute(bw, sym, strict, unsigned)
        return e, self.valid

    def quantizeDequantize(self, t, enc, rm, use_cuda):
        CALLS["quantizeDequantize"] += 1
        x = t.detach().contiguous().numpy().astype(np.float32)
        return torch.from_numpy(_REF.qdq(x.reshape(-1), enc.min, enc.max, enc.bw).reshape(x.shape))

    def quantizeDequantizePerChannel(self, t, encs, c, n, per, rm, use_cuda):
        CALLS["quantizeDequantizePerChannel"] += 1
        # AimetTensorQuantizer.cpp:256-307 with torch CPU ops, then the reference kernel
        x = t.detach().contiguous()
        enc = torch.tensor([[e.min for e in encs], [e.max for e in encs]], dtype=torch.float64).to(torch.float32)
        mn, mx = enc[0], enc[1]
        steps = 2.0 ** encs[0].bw - 1
        if encs[0].min == -encs[0].max:
            steps -= 1
        zero = torch.zeros(1)
        mn = torch.minimum(mn, zero)
        mx = torch.maximum(mx, zero)
        mx = torch.maximum(mx, mn + 1e-5)
        delta = (mx - mn) / steps
        offset = torch.round(mn / delta)
        out = _REF.qdq_per_channel(x.numpy().reshape(-1), c, per, mn.numpy().copy(), mx.numpy().copy(),
                                   delta.numpy().copy(), offset.numpy().copy())
        return torch.from_numpy(out.reshape(x.shape))

    def getStatsHistogram(self):
        h = self.a.histogram()
        return list(zip(h[0].tolist(), h[1].tolist()))

    def makeDeltaOffsetTensor(self, device, encodings):
        # AimetTensorQuantizer.cpp:209-234: float32 [2, C] from the doubles, moved to `device`, rows returned
        t = torch.tensor([[e.delta for e in encodings], [e.offset for e in encodings]], dtype=torch.float64).to(torch.float32)
        t = t.to(device)
        return t[0], t[1]

    def setPercentileValue(self, p):
        if self.scheme == int(QuantizationMode.QUANTIZATION_PERCENTILE):   # AimetTensorQuantizer.cpp:200-207
            self.a.set_percentile(p)


atq = types.ModuleType("aimet_common.AimetTensorQuantizer")
atq.AimetTensorQuantizer = AimetTensorQuantizer
sys.modules["aimet_common.AimetTensorQuantizer"] = atq
