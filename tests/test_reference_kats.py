"""Further known answers of the reference's own C++ unit tests, values transcribed with their
tolerances; each case runs on the CPU oracle (not gpu) and on the CUDA path (gpu):

    DlQ/test/TestQuantizationFunctions.cpp:119-277   quantizeDequantizeBroadcast, three layouts, EXPECT_EQ
    DlQ/test/TestTensorQuantizationSim.cpp:265-446   per-channel packed / QDQ on an axis, 32-bit grid, fillEncodingInfo
    DlQ/test/TestTensorQuantizer.cpp:140-569         computeEncodingFromData, partial encodings, per-channel encodings,
                                                     packed per channel (asymmetric / strict symmetric), QDQ per channel,
                                                     packed + dequantize

The reference's per-channel 4-D helpers (TensorQuantizer.cpp:189-326: slice on `axis`, one TF encoding per slice, QDQ /
pack every slice, concat) are C++-only conveniences; here they are composed from the path's primitives -- move the axis
to the front, per-channel TF encodings, per-channel QDQ / packed grid -- which is what the Python hosts do.
"""
import numpy as np
import pytest

from oracle.bindings import OracleTf, OracleTfe


# ---- the two back ends ---------------------------------------------------------------------------------------------
class _Oracle:
    def __init__(self, oracle):
        self.o = oracle

    def tf_encoding(self, x, bw, sym, strict, unsigned):
        a = OracleTf(self.o)
        a.update(np.ascontiguousarray(x, np.float32).reshape(-1))
        return a.compute(bw, sym, strict, unsigned)

    def tfe_encoding(self, x, bw):
        a = OracleTfe(self.o)
        a.update(np.ascontiguousarray(x, np.float32).reshape(-1))
        return a.compute(bw)

    def fill(self, bw, mn, mx):
        return self.o.fill_encoding_info(bw, mn, mx)

    def partial(self, bw, enc, sym, unsigned, strict):
        return self.o.partial_encoding(bw, enc, sym, unsigned, strict)[1]

    def grid(self, x, mn, mx, bw, signed):
        return self.o.quantize(np.ascontiguousarray(x, np.float32), mn, mx, bw, signed)

    def packed(self, x, mn, mx, bw, signed):
        return self.o.quantize_packed(np.ascontiguousarray(x, np.float32), mn, mx, bw, signed)

    def qdq_rows(self, rows, encs, bw):
        return np.stack([self.o.qdq(np.ascontiguousarray(r), e[0], e[1], bw) for r, e in zip(rows, encs)])

    def broadcast(self, x, mn, mx, delta, offset):
        return self.o.qdq_broadcast(x, mn, mx, delta, offset)


class _Cuda(_Oracle):
    """The CUDA path through the drop-in classes / ops; every result is also compared with the oracle's, bit for bit."""

    def tf_encoding(self, x, bw, sym, strict, unsigned):
        from aimet_b200 import libpymo
        a = libpymo.EncodingAnalyzerForPython(libpymo.QuantizationMode.QUANTIZATION_TF)
        a.updateStats(np.ascontiguousarray(x, np.float32), True)
        e, valid = a.computeEncoding(bw, sym, strict, unsigned)
        got = (e.min, e.max, e.delta, e.offset, e.bw)
        assert valid and got == tuple(super().tf_encoding(x, bw, sym, strict, unsigned))
        return got

    def tfe_encoding(self, x, bw):
        from aimet_b200 import libpymo
        tq = libpymo.TensorQuantizer(libpymo.QuantizationMode.QUANTIZATION_TF_ENHANCED,
                                     libpymo.RoundingMode.ROUND_NEAREST)
        tq.updateStats(np.ascontiguousarray(x, np.float32), True)
        e = tq.computeEncoding(bw, False)
        got = (e.min, e.max, e.delta, e.offset, e.bw)
        assert got == tuple(super().tfe_encoding(x, bw))
        return got

    def fill(self, bw, mn, mx):
        from aimet_b200 import ops
        e = ops.fill_encoding_info(bw, mn, mx)
        got = (e.min, e.max, e.delta, e.offset, e.bw)
        assert got[:4] == tuple(super().fill(bw, mn, mx))[:4]
        return got

    def partial(self, bw, enc, sym, unsigned, strict):
        from aimet_b200 import libpymo
        tq = libpymo.TensorQuantizer(libpymo.QuantizationMode.QUANTIZATION_TF, libpymo.RoundingMode.ROUND_NEAREST)
        e = libpymo.TfEncoding()
        e.min, e.max, e.delta, e.offset, e.bw = enc
        tq.computePartialEncoding(bw, e, sym, unsigned, strict)
        got = (e.min, e.max, e.delta, e.offset, e.bw)
        assert got[:4] == tuple(super().partial(bw, enc, sym, unsigned, strict))[:4]
        return got

    def grid(self, x, mn, mx, bw, signed):
        import torch
        from aimet_b200 import ops
        got = ops.quantize_to_grid_impl(torch.from_numpy(np.ascontiguousarray(x, np.float32)).cuda(), mn, mx, bw, 0,
                                        signed).cpu().numpy()
        assert np.array_equal(got, super().grid(x, mn, mx, bw, signed))
        return got

    def packed(self, x, mn, mx, bw, signed):
        import torch
        from aimet_b200 import ops
        got = ops.quantize_to_packed_impl(torch.from_numpy(np.ascontiguousarray(x, np.float32)).cuda(), mn, mx, bw,
                                          signed).cpu().numpy()
        assert np.array_equal(got, super().packed(x, mn, mx, bw, signed))
        return got

    def qdq_rows(self, rows, encs, bw):
        import torch
        from aimet_b200 import AimetTensorQuantizer, libpymo
        q = AimetTensorQuantizer(libpymo.QuantizationMode.QUANTIZATION_TF)
        es = []
        for e in encs:
            t = libpymo.TfEncoding()
            t.min, t.max, t.delta, t.offset, t.bw = e
            es.append(t)
        x = torch.from_numpy(np.ascontiguousarray(rows, np.float32)).cuda()
        got = q.quantizeDequantizePerChannel(x, es, x.shape[0], x.numel(), x.shape[1],
                                             libpymo.RoundingMode.ROUND_NEAREST, True).cpu().numpy()
        assert np.array_equal(got.view(np.uint32), super().qdq_rows(rows, encs, bw).view(np.uint32))
        return got

    def broadcast(self, x, mn, mx, delta, offset):
        import torch
        from aimet_b200 import ops
        got = ops.qdq_broadcast_impl(torch.from_numpy(x).cuda(),
                                     *[torch.from_numpy(e).cuda() for e in (mn, mx, delta, offset)]).cpu().numpy()
        assert np.array_equal(got.view(np.uint32), super().broadcast(x, mn, mx, delta, offset).view(np.uint32))
        return got


@pytest.fixture(params=["oracle", pytest.param("cuda", marks=pytest.mark.gpu)])
def be(request, oracle):
    return _Oracle(oracle) if request.param == "oracle" else _Cuda(oracle)


# ---- the helpers of DlQ/test/test_quantization_lib.hpp:191-237 -----------------------------------------------------------
def get_tf_encoding(mn, mx, bw):
    steps = 2 ** bw - 1
    delta = (mx - mn) / steps
    offset = float(np.round(mn / delta))          # C round(): no exact .5 among the cases below
    return (delta * offset, delta * steps + delta * offset, delta, offset, bw)


def get_tf_symmetric_encoding(mx, bw):
    half_steps = 2 ** bw - 2
    positive = half_steps // 2
    delta = mx / positive
    offset = -float(-(-half_steps // 2))
    return (offset * delta, delta * positive, delta, offset, bw)


def double_eq(a, b):
    """gtest's EXPECT_DOUBLE_EQ: within 4 ULP"""
    a, b = np.float64(a), np.float64(b)
    if a == b:
        return True
    ia, ib = int(a.view(np.int64)), int(b.view(np.int64))
    ia = ia if ia >= 0 else -(ia & 0x7FFFFFFFFFFFFFFF)
    ib = ib if ib >= 0 else -(ib & 0x7FFFFFFFFFFFFFFF)
    return abs(ia - ib) <= 4


def float_eq(a, b):
    a, b = np.float32(a), np.float32(b)
    if a == b:
        return True
    ia, ib = int(a.view(np.int32)), int(b.view(np.int32))
    ia = ia if ia >= 0 else -(ia & 0x7FFFFFFF)
    ib = ib if ib >= 0 else -(ib & 0x7FFFFFFF)
    return abs(ia - ib) <= 4


def same_encoding(e, expected):
    return all(double_eq(a, b) for a, b in zip(e[:4], expected[:4])) and int(e[4]) == expected[4]


# the fixture data of TestTensorQuantizer.cpp:57-75
DATA1 = np.arange(24, dtype=np.float32).reshape(2, 3, 2, 2)
DATA2 = (np.arange(60, dtype=np.float32) * np.float32(0.5) - np.float32(15)).reshape(1, 4, 5, 3)


def per_channel(be, x, axis, bw, strict):
    """TensorQuantizer::generatePerChannelEncodings (TensorQuantizer.cpp:265-326): (rows [C, rest], one TF encoding each);
    symmetric <=> strict (`useSymmetricEncodings = _useStrictSymmetric || _useUnsignedSymmetric`)."""
    rows = np.moveaxis(x, axis, 0).reshape(x.shape[axis], -1)
    return rows, [be.tf_encoding(r, bw, strict, strict, False) for r in rows]


def concat(rows, shape, axis):
    moved = list(shape)
    moved.insert(0, moved.pop(axis))
    return np.moveaxis(np.asarray(rows).reshape(moved), 0, axis).reshape(-1)


# ---- TestQuantizationFunctions.cpp: quantizeDequantizeBroadcast ------------------------------------------------------
def test_broadcast_known_answers(be):
    f = np.float32
    # :119-171  input {2,2,2,2}, encodings {2,1,1,2}
    x = np.array([-125.1, -125.1, 48.3, 48.3, 68.3, 68.3, -3.1, -3.1] * 2, f).reshape(2, 2, 2, 2)
    enc = [np.array(v, f).reshape(2, 1, 1, 2) for v in ([-64.0, -128.0, -256.0, -512.0], [63.5, 127.0, 254.0, 508.0],
                                                        [0.5, 1.0, 2.0, 4.0], [-128] * 4)]
    exp = [-64.0, -125.0, 48.5, 48.0, 63.5, 68.0, -3.0, -3.0, -126.0, -124.0, 48.0, 48.0, 68.0, 68.0, -4.0, -4.0]
    assert be.broadcast(x, *enc).reshape(-1).tolist() == exp
    delta, offset = [0.25, 1.0, 0.5, 2.0, 0.25, 10.0], [0, 0, 0, -1, -10, 0]
    mn, mx = [0, 0, 0, -2, -2.5, 0], [255. * 0.25, 255.0, 127.5, 508., 245. * 0.25, 2550.]
    # :173-225  input {2,3,4}, encodings {2,3,1}
    x = np.array([0.126, 10.4, -12.3, 10000] * 6, f).reshape(2, 3, 4)
    enc = [np.array(v, f).reshape(2, 3, 1) for v in (mn, mx, delta, offset)]
    exp = [0.25, 10.5, 0, 63.75, 0., 10., 0., 255., 0., 10.5, 0., 127.5, 0., 10., -2., 508., 0.25, 10.5, -2.5, 61.25,
           0., 10., 0, 2550.]
    assert be.broadcast(x, *enc).reshape(-1).tolist() == exp
    # :228-277  input {4,2,3}, encodings {2,3}
    x = np.repeat(np.array([0.126, 10.4, -12.3, 10000], f), 6).reshape(4, 2, 3)
    enc = [np.array(v, f).reshape(2, 3) for v in (mn, mx, delta, offset)]
    exp = [0.25, 0., 0., 0., 0.25, 0., 10.5, 10., 10.5, 10., 10.5, 10., -0., 0., 0., -2, -2.5, 0., 63.75, 255., 127.5,
           508, 61.25, 2550]
    assert be.broadcast(x, *enc).reshape(-1).tolist() == exp       # EXPECT_EQ: -0. == 0.


# ---- TestTensorQuantizationSim.cpp ------------------------------------------------------------------------------------
SIX = np.array([-0.5, -0.25, 0, 0.25, 0.5, 0.75], np.float32)


def test_sim_packed_per_channel_unsigned(be):                                            # :265-307
    rows = SIX.reshape(2, 3)                                       # the two slices the test hands over, each {1,1,1,3}
    out = np.concatenate([be.packed(r, -0.46, 0.72, 8, False) for r in rows])
    assert out.tolist() == [0, 45, 99, 153, 207, 255]


def test_sim_qdq_per_channel(be):                                                        # :309-352
    rows = np.array([[-0.5, -0.25, 0, 0.25, 0.5, 0.75], [0, 0.25, 0.5, 0.75, -0.5, -0.25]], np.float32)
    got = be.qdq_rows(rows, [be.fill(8, -0.5, -0.1)] * 2, 8)       # max < 0 is gated to 0: the grid is [-0.5, 0]
    # two slices of shape {1,6} concatenated on axis 2 (past the slice's last dimension): element i of slice c lands at 2i + c
    out = np.stack(got, axis=1).reshape(-1)
    exp = np.array([-0.5, 0, -0.24902, 0, 0, 0, 0, 0, 0, -0.5, 0, -0.24902], np.float32)
    assert np.abs(out - exp).max() <= 1e-3


def test_sim_dequantize_per_channel(be):                                                 # :355-388 (inverse of :265-307)
    enc = be.fill(8, -0.46, 0.72)
    packed = np.array([0, 45, 99, 153, 207, 255], np.uint8)
    back = (enc[2] * (packed.astype(np.float64) + enc[3])).astype(np.float32)            # trim_functions.cpp:449-456
    assert np.abs(back - SIX).max() <= 0.06
    assert be.packed(back, -0.46, 0.72, 8, False).tolist() == packed.tolist()            # and the grid is a fixed point


def test_sim_32_bit_quantize_only_signed(be):                                            # :390-407
    out = be.grid(np.array([-1.0], np.float32), -1.0, 1.0, 32, True)
    assert float_eq(out[0], np.float32(-2147483648))


def test_sim_fill_encoding_info(be):                                                     # :409-446
    e = be.fill(3, -5.0, 10.0)
    assert double_eq(e[0], -4.2857142857142857142857142857142) and double_eq(e[1], 10.714285714285714285714285714286)
    e = be.fill(3, -5.0, 5.0)                                      # min == -max: one step fewer
    assert double_eq(e[0], -5.0) and double_eq(e[1], 5.0)


# ---- TestTensorQuantizer.cpp ------------------------------------------------------------------------------------------
def test_quantizer_compute_encoding_from_data(be):                                       # :140-172
    import os
    from tests.conftest import GOLDEN
    data4 = np.load(os.path.join(GOLDEN, "ref_unit_inputs.npz"))["n22_seed1"]
    e = be.tfe_encoding(data4, 8)
    assert abs(e[0] + 6.527) <= 0.001 and abs(e[1] - 8.884) <= 0.001
    e = be.tf_encoding(data4, 8, True, True, False)
    expected_max = max(abs(float(data4.min())), abs(float(data4.max())))
    assert abs(e[1] - expected_max) <= e[2] / 2 + 1e-4 and e[1] == -e[0]
    assert abs(e[0] + e[2] * (-e[3])) <= 1e-7 and float_eq(e[2], (e[1] - e[0]) / 254) and e[3] == -127 and e[4] == 8


def test_quantizer_partial_encodings(be):                                                # :174-247
    exp = get_tf_symmetric_encoding(15.0, 8)
    e = be.partial(8, (0.0, 0.0, exp[2], exp[3], 8), True, False, True)                  # delta / offset given
    assert abs(e[1] - exp[1]) <= 0.001 and abs(e[0] - exp[0]) <= 0.001
    assert float_eq(e[2], exp[2]) and float_eq(e[3], exp[3]) and e[4] == 8
    exp = get_tf_encoding(0.0, 24.0, 8)
    e = be.partial(8, (0.0, 0.0, exp[2], exp[3], 8), False, False, False)
    assert abs(e[1] - exp[1]) <= 0.001 and abs(e[0] - exp[0]) <= 0.001
    assert float_eq(e[2], exp[2]) and float_eq(e[3], exp[3]) and e[4] == 8
    expected_max = float(np.abs(DATA2).max())                                            # min / max given
    e = be.partial(8, (-expected_max, expected_max, 0.0, 0.0, 8), True, False, True)
    assert e[1] == expected_max and e[1] == -e[0] and abs(e[0] + e[2] * (-e[3])) <= 1e-7
    assert float_eq(e[2], (e[1] - e[0]) / 254) and e[3] == -127 and e[4] == 8


def test_quantizer_per_channel_encodings(be):                                            # :249-288
    rows, encs = per_channel(be, DATA1, 1, 8, False)
    assert rows.tolist() == [[0, 1, 2, 3, 12, 13, 14, 15], [4, 5, 6, 7, 16, 17, 18, 19], [8, 9, 10, 11, 20, 21, 22, 23]]
    for e, top in zip(encs, (15, 19, 23)):
        assert same_encoding(e, get_tf_encoding(0, top, 8))


def test_quantizer_packed_per_channel_asymmetric(be):                                    # :292-326
    rows, encs = per_channel(be, DATA2, 1, 8, False)
    for e, (lo, hi) in zip(encs, ((-15, 0), (-7.5, 0), (0, 7), (0, 14.5))):
        assert same_encoding(e, get_tf_encoding(lo, hi, 8))
    out = concat([be.packed(r, e[0], e[1], 8, False) for r, e in zip(rows, encs)], DATA2.shape, 1)
    assert out.tolist() == [0, 9, 17, 26, 34, 43, 51, 60, 68, 77, 85, 94, 102, 111, 119,
                            0, 17, 34, 51, 68, 85, 102, 119, 136, 153, 170, 187, 204, 221, 238,
                            0, 18, 36, 55, 73, 91, 109, 128, 146, 164, 182, 200, 219, 237, 255,
                            132, 141, 149, 158, 167, 176, 185, 193, 202, 211, 220, 229, 237, 246, 255]


def test_quantizer_packed_per_channel_symmetric(be):                                     # :329-360
    rows, encs = per_channel(be, DATA2, 1, 8, True)
    for e, top in zip(encs, (15, 7.5, 7, 14.5)):
        assert same_encoding(e, get_tf_symmetric_encoding(top, 8))
    out = concat([be.packed(r, e[0], e[1], 8, True) for r, e in zip(rows, encs)], DATA2.shape, 1)
    assert out.view(np.int8).tolist() == [
        -127, -123, -119, -114, -110, -106, -102, -97, -93, -89, -85, -80, -76, -72, -68, -127, -119, -110, -102, -93,
        -85, -76, -68, -59, -51, -42, -34, -25, -17, -8, 0, 9, 18, 27, 36, 45, 54, 64, 73, 82,
        91, 100, 109, 118, 127, 66, 70, 74, 79, 83, 88, 92, 96, 101, 105, 109, 114, 118, 123, 127]


def test_quantizer_qdq_per_channel_asymmetric(be):                                       # :363-402
    rows, encs = per_channel(be, DATA2, 3, 8, False)
    for e, (lo, hi) in zip(encs, ((-15, 13.5), (-14.5, 14), (-14, 14.5))):
        assert same_encoding(e, get_tf_encoding(lo, hi, 8))
    out = concat(be.qdq_rows(rows, encs, 8), DATA2.shape, 3)
    exp = [-14.9765, -14.5294, -13.9706, -13.5235, -12.9647, -12.5176, -11.9588, -11.5118, -10.9529, -10.5059,
           -9.94706, -9.5, -9.05294, -8.49412, -8.04706, -7.48824, -7.04118, -6.48235, -6.03529, -5.47647,
           -5.02941, -4.47059, -4.02353, -3.46471, -3.01765, -2.45882, -2.01176, -1.45294, -1.00588, -0.447059,
           0, 0.447059, 1.00588, 1.45294, 2.01176, 2.45882, 3.01765, 3.46471, 4.02353, 4.47059,
           5.02941, 5.47647, 6.03529, 6.48235, 7.04118, 7.48824, 8.04706, 8.49412, 9.05294, 9.5,
           9.94706, 10.5059, 10.9529, 11.5118, 11.9588, 12.5176, 12.9647, 13.5235, 13.9706, 14.5294]
    assert np.abs(out - np.array(exp)).max() <= 0.001 and np.abs(out - DATA2.reshape(-1)).max() <= 0.06


def test_quantizer_qdq_per_channel_symmetric(be):                                        # :405-443
    rows, encs = per_channel(be, DATA2, 2, 8, True)
    for e, top in zip(encs, (15, 13.5, 12, 13, 14.5)):
        assert same_encoding(e, get_tf_symmetric_encoding(top, 8))
    out = concat(be.qdq_rows(rows, encs, 8), DATA2.shape, 2)
    exp = [-15, -14.5276, -14.0551, -13.5, -12.9685, -12.5433, -12, -11.5276, -10.9606, -10.5433,
           -10.0315, -9.51968, -9.01968, -8.44882, -7.99213, -7.44094, -6.9685, -6.49606, -5.95276, -5.52756,
           -4.99606, -4.53543, -3.9685, -3.49606, -2.9685, -2.45669, -2.04724, -1.48425, -1.02756, -0.456693,
           0, 0.472441, 0.944882, 1.48819, 2.01969, 2.55118, 3.02362, 3.49606, 3.9685, 4.50394,
           5.01575, 5.52756, 6.05118, 6.50787, 6.96457, 7.55906, 8.0315, 8.50394, 9.03543, 9.46063,
           9.99213, 10.4882, 10.9606, 11.5276, 11.9764, 12.4882, 13, 13.4724, 14.0433, 14.5]
    assert np.abs(out - np.array(exp)).max() <= 0.0001 and np.abs(out - DATA2.reshape(-1)).max() <= 0.06


def test_quantizer_packed_and_dequantize(be):                                            # :448-518
    data = np.array([-40, -1, 0, 1, 2, -50, 80], np.float32)
    e = be.tf_encoding(data, 8, False, False, False)
    assert same_encoding(e, get_tf_encoding(-50, 80, 8))
    packed = be.packed(data, e[0], e[1], 8, False)
    assert packed.tolist() == [20, 96, 98, 100, 102, 0, 255]
    back = (e[2] * (packed.astype(np.float64) + e[3])).astype(np.float32)                # trim_functions.cpp:449-456
    assert np.abs(back - np.array([-39.7647, -1.01961, 0, 1.01961, 2.03922, -49.9608, 80.0392])).max() <= 1e-4


def test_quantizer_packed_and_dequantize_per_channel(be):                                # :525-569
    data = np.array([-40, -1, 0, 1, 2, -50, 80, 1, 30, 10, 25, 3, 2, -50, 70, 1], np.float32).reshape(1, 1, 2, 8)
    rows, encs = per_channel(be, data, 2, 8, False)
    packed = [be.packed(r, e[0], e[1], 8, False) for r, e in zip(rows, encs)]
    assert concat(packed, data.shape, 2).tolist() == [20, 96, 98, 100, 102, 0, 255, 100, 170, 127, 159, 112, 110, 0,
                                                      255, 108]
    back = concat([(e[2] * (p.astype(np.float64) + e[3])).astype(np.float32) for p, e in zip(packed, encs)],
                  data.shape, 2)
    exp = [-39.7647, -1.01961, 0, 1.01961, 2.03922, -49.9608, 80.0392, 1.01961, 30.1176, 9.88235, 24.9412, 2.82353,
           1.88235, -49.8824, 70.1176, 0.941176]
    assert np.abs(back - np.array(exp)).max() <= 1e-4


# ---- the composition above against the reference's per-channel 4-D helpers RUN LIVE (oracle/_ref) -----------------------
@pytest.mark.parametrize("strict", [False, True])
@pytest.mark.parametrize("axis", [0, 1, 2, 3])
def test_per_channel_composition_equals_the_reference_live(oracle, reference, axis, strict):
    """TensorQuantizer::quantizeDequantizePerChannelTensor / quantizePerChannelTensorPacked (TensorQuantizer.cpp:189-216)
    on random 4-D tensors: encodings, dequantized output and packed bytes of the composition the tests above use
    (axis to the front, one TF encoding per slice, per-slice QDQ / packed grid, back) are the reference's, bit for bit."""
    if not hasattr(reference.L, "ref_tq_qdq_per_channel_tensor"):
        pytest.skip("oracle/_ref built before this entry point existed")
    be = _Oracle(oracle)
    rng = np.random.default_rng(40 + axis)
    for shape in ((3, 5, 4, 6), (2, 7, 1, 9)):
        x = (rng.standard_normal(shape) * rng.uniform(0.5, 6.0)).astype(np.float32)
        x[(0,) * 4] = 0.0
        rows, encs = per_channel(be, x, axis, 8, strict)
        out_ref, enc_ref = reference.tq_qdq_per_channel_tensor(x, axis, 8, strict)
        assert [tuple(e[:4]) for e in encs] == [tuple(r[:4]) for r in enc_ref.tolist()]
        mine = concat(be.qdq_rows(rows, encs, 8), shape, axis)
        assert np.array_equal(mine.view(np.uint32), out_ref.reshape(-1).view(np.uint32))
        packed_ref, enc_ref2 = reference.tq_packed_per_channel_tensor(x, axis, 8, strict)
        assert np.array_equal(enc_ref2, enc_ref)
        mine_packed = concat([be.packed(r, e[0], e[1], 8, strict) for r, e in zip(rows, encs)], shape, axis)
        assert np.array_equal(mine_packed, packed_ref)
    assert reference.tq_qdq_per_channel_tensor(np.zeros((2, 2, 2, 2), np.float32), 1, 4, False) is None   # bw < 8: throws
