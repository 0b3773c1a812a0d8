"""TEST INFRASTRUCTURE ONLY -- ctypes bindings for the oracle libraries (see oracle/__init__.py)."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "libqsim_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "libaimet_ref.so")

QUANTIZATION_TF, QUANTIZATION_TF_ENHANCED = 0, 1   # include/DlQuantization/Quantization.hpp:83-107
ROUND_NEAREST, ROUND_STOCHASTIC = 0, 1             # include/DlQuantization/Quantization.hpp:76-80
PDF_SIZE = 512


def build(with_ref=True):
    """Compile the C restatement and, if /root/reference is present, the reference itself."""
    targets = ["oracle"] + (["ref"] if with_ref else [])
    subprocess.run(["make", "-s", "-C", HERE, "-j8"] + targets, check=True)


class Encoding(C.Structure):
    _fields_ = [("min", C.c_double), ("max", C.c_double), ("delta", C.c_double), ("offset", C.c_double),
                ("bw", C.c_int)]

    def astuple(self):
        return (self.min, self.max, self.delta, self.offset, self.bw)


class TfeState(C.Structure):
    _fields_ = [("initialized", C.c_int), ("stats_updated", C.c_int), ("iterations", C.c_int),
                ("x_left", C.c_double * PDF_SIZE), ("pdf", C.c_double * PDF_SIZE)]


class TfState(C.Structure):
    _fields_ = [("stats_updated", C.c_int), ("min", C.c_double), ("max", C.c_double)]


class EntropyState(C.Structure):
    _fields_ = [("initialized", C.c_int), ("stats_updated", C.c_int), ("iterations", C.c_int), ("min", C.c_double),
                ("max", C.c_double), ("histogram", C.c_double * PDF_SIZE)]


_fp = C.POINTER(C.c_float)
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_u32p = C.POINTER(C.c_uint32)


def _f(a):
    assert a.dtype == np.float32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_fp)


def _d(a):
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_dp)


class Oracle:
    """The C restatement (oracle/qsim_oracle.c)."""

    def __init__(self):
        if not os.path.exists(ORACLE_SO):
            build(with_ref=False)
        L = self.L = C.CDLL(ORACLE_SO)
        L.qo_tf_encoding.restype = Encoding
        L.qo_tf_encoding.argtypes = [C.c_int, C.c_double, C.c_double, C.c_int, C.c_int, C.c_int]
        L.qo_fill_encoding_info.argtypes = [C.c_int, C.c_double, C.c_double, C.POINTER(Encoding)]
        L.qo_partial_encoding.argtypes = [C.c_int, C.POINTER(Encoding), C.c_int, C.c_int, C.c_int]
        L.qo_qdq_tensor.argtypes = [_fp, C.c_size_t, _fp, C.c_double, C.c_double, C.c_int]
        L.qo_quantize_tensor.argtypes = [_fp, C.c_size_t, _fp, C.c_double, C.c_double, C.c_int, C.c_int]
        L.qo_quantize_packed.restype = C.c_int64
        L.qo_quantize_packed.argtypes = [_fp, C.c_size_t, C.c_void_p, C.c_double, C.c_double, C.c_int, C.c_int]
        L.qo_per_channel_prepare.argtypes = [_dp, _dp, C.c_int, C.c_int, _fp, _fp, _fp, _fp]
        L.qo_qdq_per_channel.argtypes = [_fp, C.c_size_t, C.c_size_t, C.c_size_t, _fp, _fp, _fp, _fp, _fp]
        L.qo_qdq_broadcast.argtypes = [_fp, _fp, C.c_int64, C.c_int64, C.POINTER(C.c_int64), C.POINTER(C.c_int64), _fp, _fp,
                                       _fp, _fp]
        L.qo_ste_bwd.argtypes = [_fp, _fp, C.c_size_t, C.c_float, C.c_float, _fp]
        L.qo_ste_bwd_per_channel.argtypes = [_fp, _fp, C.c_size_t, C.c_size_t, C.c_size_t, _fp, _fp, _fp]
        L.qo_bf16_to_f32.restype = C.c_float
        L.qo_bf16_to_f32.argtypes = [C.c_uint16]
        L.qo_f32_to_bf16.restype = C.c_uint16
        L.qo_f32_to_bf16.argtypes = [C.c_float]
        L.qo_get_min.restype = C.c_float
        L.qo_get_max.restype = C.c_float
        L.qo_get_min.argtypes = [_fp, C.c_size_t]
        L.qo_get_max.argtypes = [_fp, C.c_size_t]
        L.qo_histogram.argtypes = [_fp, C.c_size_t, _u32p, C.c_float, C.c_float]
        L.qo_tfe_bucket_params.argtypes = [C.POINTER(TfeState), _fp, _fp]
        L.qo_tf_init.argtypes = [C.POINTER(TfState)]
        L.qo_tf_update.argtypes = [C.POINTER(TfState), _fp, C.c_size_t]
        L.qo_tf_compute.restype = Encoding
        L.qo_tf_compute.argtypes = [C.POINTER(TfState), C.c_int, C.c_int, C.c_int, C.c_int]
        L.qo_tfe_init.argtypes = [C.POINTER(TfeState)]
        L.qo_tfe_init_pdf.argtypes = [C.POINTER(TfeState), C.c_float, C.c_float]
        L.qo_tfe_update.argtypes = [C.POINTER(TfeState), _fp, C.c_size_t]
        L.qo_tfe_fold_histogram.argtypes = [C.POINTER(TfeState), _u32p, C.c_size_t]
        L.qo_tfe_compute.restype = Encoding
        L.qo_tfe_compute.argtypes = [C.POINTER(TfeState), C.c_int, C.c_int, C.c_int, C.c_int]
        L.qo_mse_compute.restype = Encoding
        L.qo_mse_compute.argtypes = [C.POINTER(TfeState), C.c_int, C.c_int, C.c_int, C.c_int]
        L.qo_percentile_compute.restype = Encoding
        L.qo_percentile_compute.argtypes = [C.POINTER(TfeState), C.c_float, C.c_int, C.c_int, C.c_int, C.c_int]
        L.qo_tfe_cost.restype = C.c_double
        L.qo_tfe_cost.argtypes = [C.POINTER(TfeState), C.c_int, C.c_float, C.c_int]
        L.qo_tfe_candidates.restype = C.c_int
        L.qo_tfe_candidates.argtypes = [C.POINTER(TfeState), C.c_int, C.c_int, C.c_int, C.c_int, _fp, _ip, _fp]
        L.qo_entropy_init.argtypes = [C.POINTER(EntropyState)]
        L.qo_entropy_update.argtypes = [C.POINTER(EntropyState), _fp, C.c_size_t]
        L.qo_entropy_compute.restype = Encoding
        L.qo_entropy_compute.argtypes = [C.POINTER(EntropyState), C.c_int, C.c_int, C.c_int, C.c_int]
        L.qo_rescale_histogram.argtypes = [_dp, C.c_double, C.c_double, C.c_double, C.c_double, _dp]

    # -- encodings ------------------------------------------------------------------------------
    def tf_encoding(self, bw, mn, mx, sym=False, strict=False, unsigned=False):
        return self.L.qo_tf_encoding(bw, mn, mx, int(sym), int(strict), int(unsigned)).astuple()

    def fill_encoding_info(self, bw, mn, mx):
        e = Encoding()
        self.L.qo_fill_encoding_info(bw, mn, mx, C.byref(e))
        return e.astuple()

    def partial_encoding(self, bw, enc, sym, unsigned, strict):
        e = Encoding(*enc[:4], int(enc[4]))
        rc = self.L.qo_partial_encoding(bw, C.byref(e), int(sym), int(unsigned), int(strict))
        return rc, e.astuple()

    # -- element-wise ---------------------------------------------------------------------------
    def qdq(self, x, mn, mx, bw):
        x = np.ascontiguousarray(x, dtype=np.float32)
        out = np.empty_like(x)
        self.L.qo_qdq_tensor(_f(x), x.size, _f(out), mn, mx, bw)
        return out

    def quantize(self, x, mn, mx, bw, shift_to_signed):
        x = np.ascontiguousarray(x, dtype=np.float32)
        out = np.empty_like(x)
        self.L.qo_quantize_tensor(_f(x), x.size, _f(out), mn, mx, bw, int(shift_to_signed))
        return out

    def quantize_packed(self, x, mn, mx, bw, shift_to_signed):
        """-> uint8 array of max(bw, 8) / 8 bytes per element (quantizeTensorPacked), or None for an unsupported bitwidth"""
        x = np.ascontiguousarray(x, dtype=np.float32)
        out = np.zeros(x.size * max(bw, 8) // 8, dtype=np.uint8)
        n = self.L.qo_quantize_packed(_f(x), x.size, out.ctypes.data_as(C.c_void_p), mn, mx, bw, int(shift_to_signed))
        return None if n < 0 else out[:n]

    def per_channel_prepare(self, mins, maxs, bw):
        mins = np.ascontiguousarray(mins, dtype=np.float64)
        maxs = np.ascontiguousarray(maxs, dtype=np.float64)
        c = mins.size
        outs = [np.empty(c, np.float32) for _ in range(4)]
        self.L.qo_per_channel_prepare(_d(mins), _d(maxs), c, bw, *[_f(o) for o in outs])
        return outs

    def qdq_per_channel(self, x, num_channel, num_per_channel, emin, emax, edelta, eoffset):
        x = np.ascontiguousarray(x, dtype=np.float32)
        out = np.empty_like(x)
        self.L.qo_qdq_per_channel(_f(x), num_channel, x.size, num_per_channel, _f(out), _f(emin), _f(emax),
                                  _f(edelta), _f(eoffset))
        return out

    def ste_bwd(self, x, grad, mn, mx):
        x = np.ascontiguousarray(x, dtype=np.float32)
        grad = np.ascontiguousarray(grad, dtype=np.float32)
        out = np.empty_like(x)
        self.L.qo_ste_bwd(_f(x), _f(grad), x.size, mn, mx, _f(out))
        return out

    def ste_bwd_per_channel(self, x, grad, num_channel, num_per_channel, emin, emax):
        x = np.ascontiguousarray(x, dtype=np.float32)
        grad = np.ascontiguousarray(grad, dtype=np.float32)
        out = np.empty_like(x)
        self.L.qo_ste_bwd_per_channel(_f(x), _f(grad), num_channel, x.size, num_per_channel, _f(emin), _f(emax),
                                      _f(out))
        return out

    def qdq_broadcast(self, x, enc_min, enc_max, enc_delta, enc_offset):
        """x: any-rank fp32 array; the four encoding arrays share a shape that broadcasts to x's."""
        x = np.ascontiguousarray(x, dtype=np.float32)
        encs = [np.ascontiguousarray(e, dtype=np.float32) for e in (enc_min, enc_max, enc_delta, enc_offset)]
        in_strides, enc_strides = broadcast_strides(x.shape, encs[0].shape)
        out = np.empty_like(x)
        i64 = C.c_int64 * len(in_strides)
        self.L.qo_qdq_broadcast(_f(x), _f(out), x.size, len(in_strides), i64(*in_strides), i64(*enc_strides),
                                *[_f(e) for e in encs])
        return out

    # -- statistics -----------------------------------------------------------------------------
    def get_min_max(self, x):
        x = np.ascontiguousarray(x, dtype=np.float32)
        return self.L.qo_get_min(_f(x), x.size), self.L.qo_get_max(_f(x), x.size)

    def histogram(self, x, bucket_size, pdf_offset):
        x = np.ascontiguousarray(x, dtype=np.float32)
        h = np.zeros(PDF_SIZE, np.uint32)
        self.L.qo_histogram(_f(x), x.size, h.ctypes.data_as(_u32p), bucket_size, pdf_offset)
        return h


def broadcast_strides(x_shape, enc_shape):
    """Element strides of a contiguous input and of the encoding tensor padded to the same rank (0 where it broadcasts)."""
    nd = max(len(x_shape), 1)
    xs = tuple(x_shape) if len(x_shape) else (1,)
    padded = (1,) * (nd - len(enc_shape)) + tuple(enc_shape)
    in_strides, enc_strides = [0] * nd, [0] * nd
    acc_in = acc_enc = 1
    for d in range(nd - 1, -1, -1):
        in_strides[d] = acc_in
        enc_strides[d] = acc_enc if padded[d] != 1 else 0
        acc_in *= xs[d]
        acc_enc *= padded[d]
    return in_strides, enc_strides


class OracleTf:
    def __init__(self, oracle):
        self.o, self.s = oracle, TfState()
        oracle.L.qo_tf_init(C.byref(self.s))

    def update(self, x):
        x = np.ascontiguousarray(x, dtype=np.float32)
        self.o.L.qo_tf_update(C.byref(self.s), _f(x), x.size)

    def compute(self, bw, sym=False, strict=False, unsigned=False):
        return self.o.L.qo_tf_compute(C.byref(self.s), bw, int(sym), int(strict), int(unsigned)).astuple()


class OracleEntropy:
    """EntropyEncodingAnalyzer<float> restated (qo_entropy_*)."""

    def __init__(self, oracle):
        self.o, self.s = oracle, EntropyState()
        oracle.L.qo_entropy_init(C.byref(self.s))

    def update(self, x):
        x = np.ascontiguousarray(x, dtype=np.float32).reshape(-1)
        self.o.L.qo_entropy_update(C.byref(self.s), _f(x), x.size)

    def compute(self, bw, sym=False, strict=False, unsigned=False):
        return self.o.L.qo_entropy_compute(C.byref(self.s), bw, int(sym), int(strict), int(unsigned)).astuple()

    def raw(self):
        """(histogram[512] or None, min, max, iterations)"""
        hist = np.array(self.s.histogram[:]) if self.s.initialized else None
        return hist, self.s.min, self.s.max, self.s.iterations


class OracleTfe:
    def __init__(self, oracle):
        self.o, self.s = oracle, TfeState()
        oracle.L.qo_tfe_init(C.byref(self.s))

    def update(self, x):
        x = np.ascontiguousarray(x, dtype=np.float32)
        self.o.L.qo_tfe_update(C.byref(self.s), _f(x), x.size)

    def init_pdf(self, mn, mx):
        self.o.L.qo_tfe_init_pdf(C.byref(self.s), mn, mx)
        self.s.stats_updated = 1

    def fold_histogram(self, hist, cnt):
        hist = np.ascontiguousarray(hist, dtype=np.uint32)
        self.s.stats_updated = 1
        self.o.L.qo_tfe_fold_histogram(C.byref(self.s), hist.ctypes.data_as(_u32p), cnt)

    def bucket_params(self):
        b, o = C.c_float(), C.c_float()
        self.o.L.qo_tfe_bucket_params(C.byref(self.s), C.byref(b), C.byref(o))
        return b.value, o.value

    def compute(self, bw, sym=False, strict=False, unsigned=False):
        return self.o.L.qo_tfe_compute(C.byref(self.s), bw, int(sym), int(strict), int(unsigned)).astuple()

    def cost(self, bw, delta, offset):
        return self.o.L.qo_tfe_cost(C.byref(self.s), bw, delta, offset)

    def candidates(self, bw, sym=False, strict=False, unsigned=False):
        d = np.empty(360, np.float32)
        o = np.empty(360, np.int32)
        ns = C.c_float()
        n = self.o.L.qo_tfe_candidates(C.byref(self.s), bw, int(sym), int(strict), int(unsigned), _f(d),
                                       o.ctypes.data_as(_ip), C.byref(ns))
        return d[:n].copy(), o[:n].copy(), ns.value

    def histogram(self):
        if not self.s.initialized:
            return None
        return np.array(self.s.x_left[:]), np.array(self.s.pdf[:])


class OraclePercentile(OracleTfe):
    """PercentileEncodingAnalyzer: the tf_enhanced statistics with a percentile-clipped range."""

    def __init__(self, oracle, percentile=100.0):
        super().__init__(oracle)
        self.percentile = float(percentile)

    def set_percentile(self, percentile):
        self.percentile = float(percentile)

    def compute(self, bw, sym=False, strict=False, unsigned=False):
        return self.o.L.qo_percentile_compute(C.byref(self.s), self.percentile, bw, int(sym), int(strict),
                                              int(unsigned)).astuple()


class OracleMse(OracleTfe):
    """MseEncodingAnalyzer: the tf_enhanced statistics, (min, max) chosen by least quantisation MSE on the bin centres."""

    def compute(self, bw, sym=False, strict=False, unsigned=False):
        return self.o.L.qo_mse_compute(C.byref(self.s), bw, int(sym), int(strict), int(unsigned)).astuple()


class Reference:
    """The reference's own C++ (CPU mode), compiled by oracle/Makefile into oracle/_ref/."""

    def __init__(self):
        if not os.path.exists(REF_SO):
            raise FileNotFoundError(REF_SO + " missing: run `make -C oracle ref` where /root/reference exists")
        L = self.L = C.CDLL(REF_SO)
        L.ref_analyzer_new.restype = C.c_void_p
        L.ref_analyzer_new.argtypes = [C.c_int]
        L.ref_analyzer_free.argtypes = [C.c_void_p]
        L.ref_analyzer_update.argtypes = [C.c_void_p, _fp, C.c_size_t]
        L.ref_analyzer_compute.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, _dp]
        L.ref_analyzer_set_percentile.argtypes = [C.c_void_p, C.c_float]
        L.ref_analyzer_histogram.restype = C.c_int
        L.ref_analyzer_histogram.argtypes = [C.c_void_p, _dp, _dp]
        L.ref_fill_encoding_info.argtypes = [C.c_int, C.c_double, C.c_double, _dp]
        L.ref_qdq.argtypes = [_fp, C.c_size_t, _fp, C.c_double, C.c_double, C.c_int, C.c_int]
        L.ref_quantize.argtypes = [_fp, C.c_size_t, _fp, C.c_double, C.c_double, C.c_int, C.c_int, C.c_int]
        L.ref_qdq_per_channel.argtypes = [_fp, C.c_size_t, C.c_size_t, C.c_size_t, _fp, _fp, _fp, _fp, _fp, C.c_int]
        if hasattr(L, "ref_quantize_packed"):      # (a library built before this entry point existed lacks it)
            L.ref_quantize_packed.restype = C.c_int64
            L.ref_quantize_packed.argtypes = [_fp, C.c_size_t, C.c_void_p, C.c_double, C.c_double, C.c_int, C.c_int]
        L.ref_qdq_broadcast.argtypes = [_fp, _fp, C.c_int64, C.c_int64, C.POINTER(C.c_int64), C.POINTER(C.c_int64), _fp, _fp,
                                        _fp, _fp]
        L.ref_tq_new.restype = C.c_void_p
        L.ref_tq_new.argtypes = [C.c_int, C.c_int]
        L.ref_tq_free.argtypes = [C.c_void_p]
        L.ref_tq_set_flags.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.ref_tq_update.argtypes = [C.c_void_p, _fp, C.c_size_t]
        L.ref_tq_compute.argtypes = [C.c_void_p, C.c_int, C.c_int, _dp]
        L.ref_tq_is_valid.argtypes = [C.c_void_p]
        L.ref_tq_reset.argtypes = [C.c_void_p]
        L.ref_tq_partial.argtypes = [C.c_void_p, C.c_int, _dp, C.c_int, C.c_int, C.c_int]
        if hasattr(L, "ref_tpp_new"):
            L.ref_tpp_new.restype = C.c_void_p
            L.ref_tpp_free.argtypes = [C.c_void_p]
            L.ref_tpp_update.argtypes = [C.c_void_p, _fp, C.c_int]
            L.ref_tpp_get.restype = C.c_int
            L.ref_tpp_get.argtypes = [C.c_void_p, _dp, _dp, _ip]
            L.ref_rescale_histogram.argtypes = [_dp, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, _dp]

    @staticmethod
    def _enc(out5):
        return (out5[0], out5[1], out5[2], out5[3], int(out5[4]))

    def fill_encoding_info(self, bw, mn, mx):
        out = np.zeros(5)
        self.L.ref_fill_encoding_info(bw, mn, mx, _d(out))
        return self._enc(out)

    def qdq(self, x, mn, mx, bw):
        x = np.ascontiguousarray(x, dtype=np.float32)
        out = np.empty_like(x)
        self.L.ref_qdq(_f(x), x.size, _f(out), mn, mx, bw, ROUND_NEAREST)
        return out

    def quantize(self, x, mn, mx, bw, shift_to_signed):
        x = np.ascontiguousarray(x, dtype=np.float32)
        out = np.empty_like(x)
        self.L.ref_quantize(_f(x), x.size, _f(out), mn, mx, bw, ROUND_NEAREST, int(shift_to_signed))
        return out

    def quantize_packed(self, x, mn, mx, bw, shift_to_signed):
        x = np.ascontiguousarray(x, dtype=np.float32)
        out = np.zeros(x.size * max(bw, 8) // 8 + 8, dtype=np.uint8)
        n = self.L.ref_quantize_packed(_f(x), x.size, out.ctypes.data_as(C.c_void_p), mn, mx, bw, int(shift_to_signed))
        return None if n < 0 else out[:n]

    def tq_qdq_per_channel_tensor(self, x4, axis, bw, strict, mode=QUANTIZATION_TF):
        """TensorQuantizer::quantizeDequantizePerChannelTensor on a 4-D tensor -> (output, [C][5] encodings) or None"""
        x4 = np.ascontiguousarray(x4, dtype=np.float32)
        shape = (C.c_uint32 * 4)(*x4.shape)
        out = np.empty_like(x4)
        enc = np.zeros((x4.shape[axis], 5), np.float64)
        n = self.L.ref_tq_qdq_per_channel_tensor(mode, _f(x4), shape, axis, bw, int(strict), _f(out), _d(enc))
        return None if n < 0 else (out, enc)

    def tq_packed_per_channel_tensor(self, x4, axis, bw, strict, mode=QUANTIZATION_TF):
        """TensorQuantizer::quantizePerChannelTensorPacked -> (bytes, encodings) or None"""
        x4 = np.ascontiguousarray(x4, dtype=np.float32)
        shape = (C.c_uint32 * 4)(*x4.shape)
        out = np.zeros(x4.size * max(bw, 8) // 8 + 8, dtype=np.uint8)
        enc = np.zeros((x4.shape[axis], 5), np.float64)
        self.L.ref_tq_packed_per_channel_tensor.restype = C.c_int64
        n = self.L.ref_tq_packed_per_channel_tensor(mode, _f(x4), shape, axis, bw, int(strict),
                                                    out.ctypes.data_as(C.c_void_p), _d(enc))
        return None if n < 0 else (out[:n], enc)

    def qdq_per_channel(self, x, num_channel, num_per_channel, emin, emax, edelta, eoffset):
        x = np.ascontiguousarray(x, dtype=np.float32)
        out = np.empty_like(x)
        self.L.ref_qdq_per_channel(_f(x), num_channel, x.size, num_per_channel, _f(out), _f(emin), _f(emax),
                                   _f(edelta), _f(eoffset), ROUND_NEAREST)
        return out

    def qdq_broadcast(self, x, enc_min, enc_max, enc_delta, enc_offset):
        x = np.ascontiguousarray(x, dtype=np.float32)
        encs = [np.ascontiguousarray(e, dtype=np.float32) for e in (enc_min, enc_max, enc_delta, enc_offset)]
        in_strides, enc_strides = broadcast_strides(x.shape, encs[0].shape)
        out = np.empty_like(x)
        i64 = C.c_int64 * len(in_strides)
        self.L.ref_qdq_broadcast(_f(x), _f(out), x.size, len(in_strides), i64(*in_strides), i64(*enc_strides),
                                 *[_f(e) for e in encs])
        return out

    def partial_encoding(self, bw, enc, sym, unsigned, strict):
        h = self.L.ref_tq_new(QUANTIZATION_TF, ROUND_NEAREST)
        buf = np.array([enc[0], enc[1], enc[2], enc[3], float(enc[4])])
        rc = self.L.ref_tq_partial(h, bw, _d(buf), int(sym), int(unsigned), int(strict))
        self.L.ref_tq_free(h)
        return rc, self._enc(buf)


class RefAnalyzer:
    def __init__(self, ref, mode):
        self.r = ref
        self.h = ref.L.ref_analyzer_new(mode)

    def __del__(self):
        if getattr(self, "h", None):
            self.r.L.ref_analyzer_free(self.h)
            self.h = None

    def update(self, x):
        x = np.ascontiguousarray(x, dtype=np.float32)
        self.r.L.ref_analyzer_update(self.h, _f(x), x.size)

    def set_percentile(self, percentile):
        self.r.L.ref_analyzer_set_percentile(self.h, float(percentile))

    def compute(self, bw, sym=False, strict=False, unsigned=False):
        out = np.zeros(5)
        self.r.L.ref_analyzer_compute(self.h, bw, int(sym), int(strict), int(unsigned), _d(out))
        return Reference._enc(out)

    def histogram(self):
        xl = np.zeros(PDF_SIZE)
        pdf = np.zeros(PDF_SIZE)
        n = self.r.L.ref_analyzer_histogram(self.h, _d(xl), _d(pdf))
        return (xl, pdf) if n == PDF_SIZE else None


class RefTensorHistogram:
    """The reference's updateTensorHistogram on a TensorProfilingParams of its own (the entropy scheme's raw statistics)."""

    def __init__(self, ref):
        self.r = ref
        self.h = ref.L.ref_tpp_new()

    def __del__(self):
        if getattr(self, "h", None):
            self.r.L.ref_tpp_free(self.h)
            self.h = None

    def update(self, x):
        x = np.ascontiguousarray(x, dtype=np.float32).reshape(-1)
        self.r.L.ref_tpp_update(self.h, _f(x), x.size)

    def raw(self):
        hist, mm, it = np.zeros(PDF_SIZE), np.zeros(2), C.c_int(0)
        n = self.r.L.ref_tpp_get(self.h, _d(hist), _d(mm), C.byref(it))
        return (hist if n == PDF_SIZE else None), mm[0], mm[1], it.value
