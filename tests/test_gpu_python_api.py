"""The drop-in classes (aimet_b200.AimetTensorQuantizer, aimet_b200.libpymo.*) used the way the reference's tests use
the originals (TrainingExtensions/torch/test/python/test_tensor_quantizer.py, DlQuantization/test/python/
test_tensor_quantizer.py), checked against the oracle."""
import pickle

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_aimet_tensor_quantizer_methods(oracle):
    from aimet_b200 import AimetTensorQuantizer, libpymo
    from oracle.bindings import OracleTfe
    q = AimetTensorQuantizer(libpymo.QuantizationMode.QUANTIZATION_TF_ENHANCED)
    enc, valid = q.getEncoding(8, False, False, False)
    assert valid is False                                                   # no stats yet (ATQ:180-192)
    rng = np.random.default_rng(0)
    x = (rng.standard_normal((4, 16, 8, 8)) * 2).astype(np.float32)
    q.updateStats(torch.from_numpy(x).cuda(), True)
    o = OracleTfe(oracle)
    o.update(x.reshape(-1))
    enc, valid = q.getEncoding(8, False, False, False)
    assert valid and (enc.min, enc.max, enc.delta, enc.offset, enc.bw) == o.compute(8)
    hist = q.getStatsHistogram()
    assert len(hist) == 512                                                 # test_tensor_quantizer.py:60-92
    assert np.array_equal(np.array(hist)[:, 1], o.histogram()[1])
    y = q.quantizeDequantize(torch.from_numpy(x).cuda(), enc, libpymo.RoundingMode.ROUND_NEAREST, True)
    assert y.shape == x.shape
    assert np.array_equal(y.cpu().numpy().reshape(-1), oracle.qdq(x.reshape(-1), enc.min, enc.max, 8))
    g = q.quantize(torch.from_numpy(x).cuda(), enc, libpymo.RoundingMode.ROUND_NEAREST, True, True)
    assert np.array_equal(g.cpu().numpy().reshape(-1), oracle.quantize(x.reshape(-1), enc.min, enc.max, 8, True))
    # a CPU tensor is staged to the GPU and comes back on the CPU (the reference's use_cuda=False call sites)
    y_cpu = q.quantizeDequantize(torch.from_numpy(x), enc, libpymo.RoundingMode.ROUND_NEAREST, False)
    assert y_cpu.device.type == "cpu" and torch.equal(y_cpu, y.cpu())
    delta, offset = q.makeDeltaOffsetTensor(torch.device("cuda"), [enc, enc])
    assert delta.tolist() == [np.float32(enc.delta)] * 2 and offset.tolist() == [enc.offset] * 2
    q.resetEncodingStats()
    assert q.getEncoding(8, False, False, False)[1] is False
    q.setPercentileValue(99.0)                                               # no-op outside the percentile scheme
    # channels_last input keeps its memory format (ATQ:140 suggest_memory_format)
    xc = torch.from_numpy(x).cuda().contiguous(memory_format=torch.channels_last)
    yc = q.quantizeDequantize(xc, enc, libpymo.RoundingMode.ROUND_NEAREST, True)
    assert yc.is_contiguous(memory_format=torch.channels_last) and torch.equal(yc, y)
    with pytest.raises(ValueError):
        q.quantizeDequantize(xc, enc, 7, True)                               # "Unknown rounding mode."


def test_libpymo_tensor_quantizer(oracle):
    from aimet_b200 import libpymo
    from oracle.bindings import OracleTf
    tq = libpymo.TensorQuantizer(libpymo.QuantizationMode.QUANTIZATION_TF, libpymo.RoundingMode.ROUND_NEAREST)
    assert tq.isEncodingValid is False
    rng = np.random.default_rng(1)
    x = (rng.standard_normal((2, 3, 5, 7)) * 3).astype(np.float32)          # nd-array shapes (test_tensor_quantizer.py)
    tq.updateStats(x, False)
    enc = tq.computeEncoding(8, False)
    assert tq.isEncodingValid is True
    o = OracleTf(oracle)
    o.update(x.reshape(-1))
    assert (enc.min, enc.max, enc.delta, enc.offset, enc.bw) == o.compute(8)
    out = np.zeros_like(x)
    tq.quantizeDequantize(x, out, enc.min, enc.max, 8, False)
    assert np.array_equal(out.reshape(-1), oracle.qdq(x.reshape(-1), enc.min, enc.max, 8))
    tq.setStrictSymmetric(True)                                              # setters reset the statistics
    assert tq.isEncodingValid is False and tq.getStrictSymmetric() is True
    tq.updateStats(x, False)
    enc = tq.computeEncoding(8, True)
    assert (enc.min, enc.max, enc.delta, enc.offset) == o.compute(8, True, True, False)[:4]
    with pytest.raises(AssertionError):
        tq.getStatsHistogram()                                               # TF analyzer keeps no histogram
    # partial encodings (DlQ/test/TestTensorQuantizer.cpp:173-247)
    e = libpymo.TfEncoding()
    e.min, e.max, e.bw = -1.0, 2.0, 8
    tq.computePartialEncoding(8, e, False, False, False)
    exp = oracle.partial_encoding(8, (-1.0, 2.0, 0.0, 0.0, 8), False, False, False)[1]
    assert (e.min, e.max, e.delta, e.offset) == exp[:4]
    bad = libpymo.TfEncoding()
    bad.min, bad.max, bad.delta, bad.offset, bad.bw = -1.0, 1.0, 0.1, -10.0, 8
    with pytest.raises(RuntimeError):
        tq.computePartialEncoding(8, bad, False, False, False)


def test_encoding_analyzer_and_sim_for_python(oracle):
    from aimet_b200 import libpymo
    from oracle.bindings import OracleTfe
    a = libpymo.EncodingAnalyzerForPython(libpymo.QuantizationMode.QUANTIZATION_TF_ENHANCED)
    assert a.computeEncoding(8, False, False, False)[1] is False
    rng = np.random.default_rng(2)
    x = rng.standard_normal(10000).astype(np.float32)
    a.updateStats(x, False)
    enc, valid = a.computeEncoding(8, False, False, False)
    o = OracleTfe(oracle)
    o.update(x)
    assert valid and (enc.min, enc.max, enc.delta, enc.offset) == o.compute(8)[:4]
    sim = libpymo.TensorQuantizationSimForPython()
    y = sim.quantizeDequantize(x.reshape(100, 100), enc, libpymo.RoundingMode.ROUND_NEAREST, 8, False)
    assert y.shape == (100, 100) and np.array_equal(y.reshape(-1), oracle.qdq(x, enc.min, enc.max, 8))
    y2 = sim.quantizeDequantize(x, enc, libpymo.RoundingMode.ROUND_NEAREST, False)
    assert np.array_equal(y2, y.reshape(-1))


def test_quantizer_pickle_round_trip():
    """Quantizers pickle without their native objects and re-create them (tensor_quantizer.py:128-220)."""
    from aimet_b200 import libpymo
    from aimet_b200.quantsim import QuantScheme, StaticGridPerChannelQuantizer, StaticGridPerTensorQuantizer
    q = StaticGridPerTensorQuantizer(8, libpymo.RoundingMode.ROUND_NEAREST, QuantScheme.post_training_tf_enhanced, False,
                                     True)
    x = torch.randn(1000, device="cuda")
    q.update_encoding_stats(x)
    q.compute_encoding()
    q2 = pickle.loads(pickle.dumps(q))
    assert (q2.encoding.min, q2.encoding.max) == (q.encoding.min, q.encoding.max) and q2.enabled
    assert torch.equal(q2.quantize_dequantize(x, libpymo.RoundingMode.ROUND_NEAREST),
                       q.quantize_dequantize(x, libpymo.RoundingMode.ROUND_NEAREST))
    pc = StaticGridPerChannelQuantizer(8, libpymo.RoundingMode.ROUND_NEAREST, QuantScheme.post_training_tf, True, 4, True)
    w = torch.randn(4, 3, 3, 3, device="cuda")
    pc.update_encoding_stats(w)
    pc.compute_encoding()
    pc2 = pickle.loads(pickle.dumps(pc))
    assert len(pc2.encoding) == 4 and len(pc2._cppOp) == 4
    assert torch.equal(pc2.quantize_dequantize(w, libpymo.RoundingMode.ROUND_NEAREST),
                       pc.quantize_dequantize(w, libpymo.RoundingMode.ROUND_NEAREST))


def test_device_prefetcher_yields_every_batch_in_order():
    """Double-buffered H2D feeding (aimet_b200.utils.DevicePrefetcher): values and order are those of the host batches,
    including a trailing batch of a different shape; works with depth 1 and 2."""
    import torch
    from aimet_b200.utils import DevicePrefetcher
    g = torch.Generator().manual_seed(0)
    host = [torch.randn(8, 3, 16, 16, generator=g).pin_memory() for _ in range(7)] + [torch.randn(3, 3, 16, 16, generator=g)]
    for depth in (1, 2):
        seen = []
        for x in DevicePrefetcher(host, torch.device("cuda", 0), depth=depth):
            assert x.is_cuda
            seen.append((x * 2.0).sum(dim=(1, 2, 3)).cpu())      # some work on the consumer's stream
        assert len(seen) == len(host)
        for got, h in zip(seen, host):
            assert torch.allclose(got, (h * 2.0).sum(dim=(1, 2, 3)), rtol=1e-5, atol=1e-4)
    assert list(DevicePrefetcher([], torch.device("cuda", 0))) == []


# ---- the deferring variant of the drop-in (what aimet_b200.install registers for the reference's own Python) -----------------
def _drive_per_channel(cls, weight, scheme, sym, rounds=2):
    """The reference's per-channel loops (v1/tensor_quantizer.py:296-305, 567-570; qc_quantize_op.py:231-243) on `cls`."""
    from aimet_b200 import libpymo
    ops_ = [cls(scheme) for _ in range(weight.shape[0])]
    result = []
    for r in range(rounds):
        for op in ops_:
            op.resetEncodingStats()
        for c, op in enumerate(ops_):
            op.updateStats(weight.select(0, c).contiguous(memory_format=torch.contiguous_format), True)
        if r == 1:                                        # a second batch on the same records, no reset in between
            for c, op in enumerate(ops_):
                op.updateStats(weight.select(0, c).contiguous() * 0.5, True)     # fresh tensors: not slices of one storage
        encs = [op.getEncoding(8, sym, False, False) for op in ops_]
        assert all(valid for _, valid in encs)
        result.append([(e.min, e.max, e.delta, e.offset, e.bw) for e, _ in encs])
    hist = ops_[3].getStatsHistogram() if scheme != libpymo.QuantizationMode.QUANTIZATION_TF else None
    return result, hist, ops_


@pytest.mark.parametrize("scheme_name", ["QUANTIZATION_TF_ENHANCED", "QUANTIZATION_TF"])
@pytest.mark.parametrize("shape", [(300, 147), (64, 1024)])
def test_deferred_dropin_equals_the_call_by_call_class(scheme_name, shape):
    from aimet_b200 import libpymo
    from aimet_b200 import tensor_quantizer_op as atq
    scheme = getattr(libpymo.QuantizationMode, scheme_name)
    w = (torch.randn(shape, generator=torch.Generator().manual_seed(5)) * 0.2).cuda()
    w[7] = 0.0                                             # an all-zero channel (no range on the first batch)
    for sym in (False, True):
        plain, hist_p, _ = _drive_per_channel(atq.AimetTensorQuantizer, w, scheme, sym)
        before = dict(atq.ops.LAUNCHES)
        lazy, hist_l, keep = _drive_per_channel(atq.DeferredAimetTensorQuantizer, w, scheme, sym)
        issued = sum(atq.ops.LAUNCHES[k] - before[k] for k in before)
        assert lazy == plain and hist_l == hist_p
        assert issued < 0.5 * 7 * shape[0], issued         # 7 calls per channel; only round 2's fresh tensors launch one by one
        del keep


def test_deferred_dropin_encoding_objects_and_hazards(oracle):
    from aimet_b200 import libpymo
    from aimet_b200 import tensor_quantizer_op as atq
    from oracle.bindings import OracleTfe
    import pickle
    mode = libpymo.QuantizationMode.QUANTIZATION_TF_ENHANCED
    x = torch.randn(4, 4096, generator=torch.Generator().manual_seed(1)).cuda()
    ops_ = [atq.DeferredAimetTensorQuantizer(mode) for _ in range(4)]
    assert ops_[0].getEncoding(8, False, False, False)[1] is False          # nothing queued for an op without statistics
    for c, op in enumerate(ops_):
        op.updateStats(x[c], True)
    encs = [op.getEncoding(8, False, False, False)[0] for op in ops_]
    assert all(e._lazy is not None for e in encs)                            # owed, not computed yet
    o = OracleTfe(oracle)
    o.update(x[2].cpu().numpy())
    assert (encs[2].min, encs[2].max, encs[2].delta, encs[2].offset, encs[2].bw) == o.compute(8)
    assert all(e._lazy is None for e in encs)                                # one read resolved them all
    # writing a field of an owed encoding resolves it first; an owed encoding pickles like any other
    e2 = ops_[1].getEncoding(8, False, False, False)[0]
    e2.max = 9.0
    assert e2.max == 9.0 and e2.min == encs[1].min and e2.bw == 8
    e3 = pickle.loads(pickle.dumps(ops_[1].getEncoding(8, False, False, False)[0]))
    assert (e3.min, e3.max) == (encs[1].min, encs[1].max)
    # a queued slice whose storage is written before the queue runs is an error, not a silently different histogram
    for op in ops_:
        op.resetEncodingStats()
    for c, op in enumerate(ops_):
        op.updateStats(x[c], True)
    owed = ops_[3].getEncoding(8, False, False, False)[0]
    x.mul_(2.0)
    with pytest.raises(RuntimeError, match="modified in place"):
        atq.flush_deferred_calls()
    with pytest.raises(RuntimeError, match="deferred native calls"):
        _ = owed.min                                                         # can never be computed: says so
    # a lone large tensor (an activation) is never queued: its statistics are issued by the call itself
    act = torch.randn(1 << 20, device="cuda")
    original = act.cpu().numpy()
    a = atq.DeferredAimetTensorQuantizer(mode)
    before = atq.ops.LAUNCHES["hist"]
    a.updateStats(act, True)
    assert atq.ops.LAUNCHES["hist"] == before + 1
    act.zero_()                                                              # the model may overwrite it right away
    o = OracleTfe(oracle)
    o.update(original)
    e = a.getEncoding(8, False, False, False)[0]
    assert (e.min, e.max, e.delta, e.offset, e.bw) == o.compute(8)
