"""TEST / BASELINE INFRASTRUCTURE ONLY: `AimetTensorQuantizer`-shaped classes that run the path on the HOST CPU.

* ReferenceTensorQuantizer -- the reference's own C++ (oracle/_ref/libaimet_ref.so), kind "reference"
* PortTensorQuantizer      -- the plain-C restatement (oracle/qsim_oracle.c), kind "port"

Used by bench.py's `cpu_baseline` leg and `--impl reference` arm, and by tests/. The aimet_b200 package never imports
this module.
"""
import os

import numpy as np
import torch

from aimet_b200 import libpymo
from oracle import bindings

_REF = None
_PORT = None


def have_reference() -> bool:
    return os.path.exists(bindings.REF_SO)


def _ref():
    global _REF
    if _REF is None:
        _REF = bindings.Reference()
    return _REF


def _port():
    global _PORT
    if _PORT is None:
        _PORT = bindings.Oracle()
    return _PORT


def _np(t):
    return t.detach().to(torch.float32).contiguous().cpu().numpy().reshape(-1)


def _back(out, like):
    return torch.from_numpy(out).reshape(like.shape).to(like.device).to(like.dtype)


class ReferenceTensorQuantizer:
    """Same semantics as the reference's AimetTensorQuantizer.cpp with use_cuda=False."""
    KIND = "reference"

    def __init__(self, quantization_scheme):
        self._mode = int(quantization_scheme)
        self._valid = False
        self._a = bindings.RefAnalyzer(_ref(), self._mode)

    def resetEncodingStats(self):
        self._valid = False
        self._a = bindings.RefAnalyzer(_ref(), self._mode)

    def updateStats(self, t, use_cuda):
        self._valid = True
        self._a.update(_np(t))

    def getEncoding(self, bw, sym, strict, unsigned):
        if not self._valid:
            return libpymo.TfEncoding(), False
        mn, mx, delta, offset, b = self._a.compute(bw, sym, strict, unsigned)
        return libpymo.TfEncoding._from_values(mn, mx, delta, offset, b), True

    def quantizeDequantize(self, t, enc, round_mode, use_cuda):
        return _back(_ref().qdq(_np(t), enc.min, enc.max, enc.bw), t)

    def quantize(self, t, enc, round_mode, use_cuda, shift_to_signed):
        return _back(_ref().quantize(_np(t), enc.min, enc.max, enc.bw, shift_to_signed), t)

    def quantizeDequantizePerChannel(self, t, encs, num_channel, num_element, per_channel, round_mode, use_cuda):
        # host preparation with torch fp32 CPU ops exactly as AimetTensorQuantizer.cpp:256-299 does it
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
        out = _ref().qdq_per_channel(_np(t), num_channel, per_channel, mn.numpy().copy(), mx.numpy().copy(),
                                     delta.numpy().copy(), offset.numpy().copy())
        return _back(out, t)

    def getStatsHistogram(self):
        h = self._a.histogram()
        return [] if h is None else list(zip(h[0].tolist(), h[1].tolist()))

    def setPercentileValue(self, p):
        if self._mode == int(libpymo.QuantizationMode.QUANTIZATION_PERCENTILE):
            self._a.set_percentile(p)


class PortTensorQuantizer:
    KIND = "port"

    def __init__(self, quantization_scheme):
        self._tfe = int(quantization_scheme) == int(libpymo.QuantizationMode.QUANTIZATION_TF_ENHANCED)
        self._pct = int(quantization_scheme) == int(libpymo.QuantizationMode.QUANTIZATION_PERCENTILE)
        self._valid = False
        self._new()

    def _new(self):
        if self._pct:
            self._a = bindings.OraclePercentile(_port())
        else:
            self._a = bindings.OracleTfe(_port()) if self._tfe else bindings.OracleTf(_port())

    def resetEncodingStats(self):
        self._valid = False
        self._new()

    def updateStats(self, t, use_cuda):
        self._valid = True
        self._a.update(_np(t))

    def getEncoding(self, bw, sym, strict, unsigned):
        if not self._valid:
            return libpymo.TfEncoding(), False
        mn, mx, delta, offset, b = self._a.compute(bw, sym, strict, unsigned)
        return libpymo.TfEncoding._from_values(mn, mx, delta, offset, b), True

    def quantizeDequantize(self, t, enc, round_mode, use_cuda):
        return _back(_port().qdq(_np(t), enc.min, enc.max, enc.bw), t)

    def quantize(self, t, enc, round_mode, use_cuda, shift_to_signed):
        return _back(_port().quantize(_np(t), enc.min, enc.max, enc.bw, shift_to_signed), t)

    def quantizeDequantizePerChannel(self, t, encs, num_channel, num_element, per_channel, round_mode, use_cuda):
        o = _port()
        p = o.per_channel_prepare(np.array([e.min for e in encs]), np.array([e.max for e in encs]), encs[0].bw)
        return _back(o.qdq_per_channel(_np(t), num_channel, per_channel, *p), t)

    def getStatsHistogram(self):
        h = self._a.histogram()
        return [] if h is None else list(zip(h[0].tolist(), h[1].tolist()))

    def setPercentileValue(self, p):
        if self._pct:
            self._a.set_percentile(p)


def best_cpu_backend():
    """The reference itself where it was compiled, else the port."""
    return ReferenceTensorQuantizer if have_reference() else PortTensorQuantizer
