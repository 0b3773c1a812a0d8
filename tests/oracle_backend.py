"""TEST-ONLY: an `AimetTensorQuantizer`-shaped class backed by the CPU oracle (oracle/qsim_oracle.c).

It is injected into aimet_b200.quantsim with `set_default_op_factory` so that the HOST layer (wrappers, configurator,
QuantizationSimModel, export) can be driven on identical tensors by the oracle and by the CUDA ops, and so that it can
be compared with the reference's own Python on a machine without a GPU. Never imported by the aimet_b200 package.
"""
import numpy as np
import torch

from aimet_b200 import libpymo
from oracle.bindings import Oracle, OracleTf, OracleTfe

_ORACLE = None


def _oracle():
    global _ORACLE
    if _ORACLE is None:
        _ORACLE = Oracle()
    return _ORACLE


def _np(t):
    return t.detach().to(torch.float32).contiguous().cpu().numpy().reshape(-1)


class OracleTensorQuantizer:
    def __init__(self, quantization_scheme):
        self._tfe = int(quantization_scheme) == int(libpymo.QuantizationMode.QUANTIZATION_TF_ENHANCED)
        self._valid = False
        self._new()

    def _new(self):
        self._a = OracleTfe(_oracle()) if self._tfe else OracleTf(_oracle())

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
        out = _oracle().qdq(_np(t), enc.min, enc.max, enc.bw)
        return torch.from_numpy(out).reshape(t.shape).to(t.device).to(t.dtype)

    def quantize(self, t, enc, round_mode, use_cuda, shift_to_signed):
        out = _oracle().quantize(_np(t), enc.min, enc.max, enc.bw, shift_to_signed)
        return torch.from_numpy(out).reshape(t.shape).to(t.device).to(t.dtype)

    def quantizeDequantizePerChannel(self, t, encs, num_channel, num_element, per_channel, round_mode, use_cuda):
        o = _oracle()
        p = o.per_channel_prepare(np.array([e.min for e in encs]), np.array([e.max for e in encs]), encs[0].bw)
        out = o.qdq_per_channel(_np(t), num_channel, per_channel, *p)
        return torch.from_numpy(out).reshape(t.shape).to(t.device).to(t.dtype)

    def getStatsHistogram(self):
        h = self._a.histogram()
        return [] if h is None else list(zip(h[0].tolist(), h[1].tolist()))

    def setPercentileValue(self, p):
        pass
