"""ab_qc_quantize_op_compute -- the compute body of the reference's ONNX Runtime custom op (TrainingExtensions/onnx/src/
QcQuantizeOp.cpp:64-143, AimetOpUtils.h:98-322) -- called through the C ABI with ctypes, as an ORT kernel would call it,
against the CPU oracle: op modes, per tensor / per channel (any axis) / blockwise."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def api():
    from aimet_b200 import _lib
    return _lib, _lib.load()


def make_info(_lib, states, n, mode, per_channel=False, axis=0, block_axis=0, block_size=0, sym=False, bw=8, enabled=True):
    from aimet_b200 import ops
    enc = (_lib.Encoding * n)()
    for e in enc:
        e.bw = bw
    info = _lib.QcQuantizeInfo(states.arena.data_ptr() + states.first * ops.STATE_BYTES if states is not None else None,
                               enc, n, mode, ops.QUANTIZATION_TF_ENHANCED, int(sym), int(enabled), 1, int(per_channel), axis,
                               block_axis, block_size)
    info._keep = enc
    return info, enc


def run(api, info, x, n_enc):
    _lib, L = api
    out = torch.empty_like(x)
    ws = torch.empty(int(L.ab_qc_quantize_op_workspace_bytes(n_enc)), dtype=torch.uint8, device="cuda")
    shape = (C.c_int64 * x.dim())(*x.shape)
    stream = torch.cuda.current_stream().cuda_stream
    _lib.check(L.ab_qc_quantize_op_compute(C.byref(info), x.data_ptr(), out.data_ptr(), shape, x.dim(), _lib.AB_F32,
                                           ws.data_ptr(), stream))
    torch.cuda.synchronize()
    return out


def new_states(n):
    from aimet_b200.state import StateArena
    return StateArena.for_device(torch.device("cuda", 0)).allocate(n)


def test_per_tensor_modes(api, oracle):
    from oracle.bindings import OracleTfe
    _lib, _ = api
    rng = np.random.default_rng(0)
    x1 = (rng.standard_normal((3, 5, 70, 33)) * 2 + 1).astype(np.float32)
    x2 = (rng.standard_normal((3, 5, 70, 33)) * 3).astype(np.float32)
    st = new_states(1)
    info, enc = make_info(_lib, st, 1, _lib.OP_UPDATE_STATS)
    o = OracleTfe(oracle)
    for x in (x1, x2):
        out = run(api, info, torch.from_numpy(x).cuda(), 1)
        assert np.array_equal(out.cpu().numpy(), x)            # updateStats passes the tensor through
        o.update(x.reshape(-1))
    rec = st.read()[0]
    assert np.array_equal(rec["pdf"], o.histogram()[1]) and rec["iterations"] == 2
    # one shot: fresh statistics of THIS tensor, encoding on the host object, QDQ output, mode flips to quantizeDequantize
    info.op_mode = _lib.OP_ONE_SHOT_QDQ
    out = run(api, info, torch.from_numpy(x1).cuda(), 1)
    o1 = OracleTfe(oracle)
    o1.update(x1.reshape(-1))
    mn, mx, delta, offset, _ = o1.compute(8)
    assert (enc[0].min, enc[0].max, enc[0].delta, enc[0].offset, enc[0].bw) == (mn, mx, delta, offset, 8)
    assert info.op_mode == _lib.OP_QDQ
    assert np.array_equal(out.cpu().numpy().reshape(-1), oracle.qdq(x1.reshape(-1), mn, mx, 8))
    out2 = run(api, info, torch.from_numpy(x2).cuda(), 1)       # now plain QDQ with the kept encoding
    assert np.array_equal(out2.cpu().numpy().reshape(-1), oracle.qdq(x2.reshape(-1), mn, mx, 8))
    # disabled quantizer / passThrough
    info.enabled = 0
    assert np.array_equal(run(api, info, torch.from_numpy(x2).cuda(), 1).cpu().numpy(), x2)


def test_per_channel_one_shot_and_qdq_along_any_axis(api, oracle):
    from oracle.bindings import OracleTfe
    _lib, _ = api
    rng = np.random.default_rng(1)
    w = (rng.standard_normal((12, 7, 3, 3)) * np.linspace(0.05, 1.0, 12)[:, None, None, None]).astype(np.float32)
    st = new_states(12)
    info, enc = make_info(_lib, st, 12, _lib.OP_ONE_SHOT_QDQ, per_channel=True, axis=0, sym=True)
    out = run(api, info, torch.from_numpy(w).cuda(), 12)
    cols = [[], [], [], []]
    for c in range(12):
        o = OracleTfe(oracle)
        o.update(w[c].reshape(-1))
        mn, mx, delta, offset, _ = o.compute(8, True)
        assert (enc[c].min, enc[c].max, enc[c].delta, enc[c].offset) == (mn, mx, delta, offset), c
        for col, v in zip(cols, (mn, mx, delta, offset)):
            col.append(v)
    p = [np.array(col, dtype=np.float32) for col in cols]      # the op narrows the doubles, no gating (QuantizeDequantizeUtils.hpp:129-137)
    assert np.array_equal(out.cpu().numpy().reshape(-1), oracle.qdq_per_channel(w.reshape(-1), 12, 63, *p))
    # quantizeDequantize along axis 1 of an activation-shaped tensor
    x = (rng.standard_normal((4, 12, 9, 5)) * 0.5).astype(np.float32)
    info2, enc2 = make_info(_lib, None, 12, _lib.OP_QDQ, per_channel=True, axis=1)
    for c in range(12):
        enc2[c].min, enc2[c].max, enc2[c].delta, enc2[c].offset = enc[c].min, enc[c].max, enc[c].delta, enc[c].offset
    out2 = run(api, info2, torch.from_numpy(x).cuda(), 12)
    assert np.array_equal(out2.cpu().numpy().reshape(-1), oracle.qdq_per_channel(x.reshape(-1), 12, 45, *p))
    # statistics along an inner axis are not on this path (the reference copies slices)
    info2.op_mode = _lib.OP_UPDATE_STATS
    info2.states = st.arena.data_ptr()
    with pytest.raises(_lib.AbError):
        run(api, info2, torch.from_numpy(x).cuda(), 12)


def test_blockwise_qdq_and_one_shot(api, oracle):
    from oracle.bindings import OracleTfe, broadcast_strides
    _lib, _ = api
    rng = np.random.default_rng(2)
    w = (rng.standard_normal((6, 64)) * 0.3).astype(np.float32)
    blocks = 64 // 16
    st = new_states(6 * blocks)
    info, enc = make_info(_lib, st, 6 * blocks, _lib.OP_ONE_SHOT_QDQ, per_channel=True, axis=0, block_axis=1, block_size=16,
                          sym=True)
    out = run(api, info, torch.from_numpy(w).cuda(), 6 * blocks)
    cols = [[], [], [], []]
    for c in range(6):
        for b in range(blocks):
            o = OracleTfe(oracle)
            o.update(w[c, b * 16:(b + 1) * 16])
            vals = o.compute(8, True)[:4]
            e = enc[c * blocks + b]
            assert (e.min, e.max, e.delta, e.offset) == vals, (c, b)
            for col, v in zip(cols, vals):
                col.append(v)
    p = [np.array(col, dtype=np.float32).reshape(6, blocks, 1) for col in cols]
    exp = oracle.qdq_broadcast(w.reshape(6, blocks, 16), *p)
    assert np.array_equal(out.cpu().numpy().reshape(-1), exp.reshape(-1))
    assert info.op_mode == _lib.OP_QDQ
    # a wrong encoding count is rejected like the reference does
    bad, _ = make_info(_lib, None, 5, _lib.OP_QDQ, per_channel=True, axis=0, block_axis=1, block_size=16)
    with pytest.raises(ValueError):
        run(api, bad, torch.from_numpy(w).cuda(), 5)
