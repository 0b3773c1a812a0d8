"""Broadcast QDQ (blockwise / LPBQ): the C restatement against the reference's own quantizeDequantizeBroadcast (CPU), and
-- on a GPU -- ab_qdq_broadcast_fwd against the restatement, bit for bit."""
import numpy as np
import pytest
import torch

# (input shape, encoding shape)
CASES = [
    ((64, 8, 16), (64, 8, 1)),            # blockwise along the input-channel axis: block = 16
    ((32, 4, 16, 3, 3), (32, 4, 1, 1, 1)),  # LPBQ-style conv weight: [Cout, blocks, block, kh, kw]
    ((24, 50), (24, 1)),                  # per channel, axis 0
    ((24, 50), (1, 50)),                  # per channel, axis 1 (innermost dimension not broadcast)
    ((6, 5, 7, 9), (6, 1, 7, 1)),         # two non-adjacent axes
    ((1000,), (1,)),                      # per tensor
    ((5, 3, 4), (3, 1)),                  # fewer leading dimensions
    ((7, 3), (7, 3)),                     # nothing broadcast
    ((130, 2, 3), (130, 1, 3)),           # run length 3 < one 128-bit vector
    ((7, 2051), (7, 1)),                  # long runs of a ragged length: runs end inside 128-bit vectors
    ((3, 5, 1029), (3, 5, 1)),            # the same, blockwise layout
]


def make(idx, xshape, eshape):
    rng = np.random.default_rng(100 + idx)
    x = (rng.standard_normal(xshape) * 1.5).astype(np.float32)
    flat = x.reshape(-1)
    flat[::53] = 0.0
    flat[3::101] = 40.0
    flat[7::103] = -40.0
    span = rng.uniform(0.5, 3.0, eshape).astype(np.float32)
    mn = (-span * rng.uniform(0.2, 1.0, eshape)).astype(np.float32)
    mx = (span * rng.uniform(0.2, 1.0, eshape)).astype(np.float32)
    bw = 8 if idx % 2 == 0 else 4
    delta = ((mx - mn) / np.float32(2 ** bw - 1)).astype(np.float32)
    offset = np.round(mn / delta).astype(np.float32)
    return x, mn, mx, delta, offset


@pytest.mark.parametrize("idx", range(len(CASES)))
def test_port_matches_reference(oracle, reference, idx):
    x, mn, mx, delta, offset = make(idx, *CASES[idx])
    a = oracle.qdq_broadcast(x, mn, mx, delta, offset)
    b = reference.qdq_broadcast(x, mn, mx, delta, offset)
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    # and it is what element-wise QDQ with the broadcast encodings gives
    full = [np.broadcast_to(e.reshape((1,) * (x.ndim - e.ndim) + e.shape), x.shape).reshape(-1) for e in (mn, mx, delta, offset)]
    flat = x.reshape(-1)
    step = max(1, flat.size // 200)
    for i in range(0, flat.size, step):
        one = oracle.qdq_broadcast(flat[i:i + 1], *[f[i:i + 1] for f in full])
        assert one.view(np.uint32)[0] == a.reshape(-1).view(np.uint32)[i]


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("idx", range(len(CASES)))
def test_device_matches_port(oracle, idx, dtype):
    from aimet_b200 import ops
    x, mn, mx, delta, offset = make(idx, *CASES[idx])
    xt = torch.from_numpy(x).to(dtype)
    exp = oracle.qdq_broadcast(xt.float().numpy(), mn, mx, delta, offset)
    exp = torch.from_numpy(exp).to(dtype)
    enc = [torch.from_numpy(e).cuda() for e in (mn, mx, delta, offset)]
    out = ops.qdq_broadcast_impl(xt.cuda(), *enc)
    assert out.dtype == dtype and out.shape == xt.shape
    assert torch.equal(out.cpu().view(torch.int16 if dtype == torch.bfloat16 else torch.int32),
                       exp.view(torch.int16 if dtype == torch.bfloat16 else torch.int32))
    # an unaligned view takes the element-wise kernel and must agree
    if xt.numel() > 8:
        pad = torch.cat([torch.zeros(1, dtype=dtype), xt.reshape(-1)]).cuda()[1:].view(xt.shape)
        assert pad.data_ptr() % 16 != 0
        assert torch.equal(ops.qdq_broadcast_impl(pad, *enc), out)


@pytest.mark.gpu
def test_large_blockwise_and_errors(oracle):
    from aimet_b200 import ops
    rng = np.random.default_rng(9)
    x = rng.standard_normal((512, 32, 64)).astype(np.float32)           # 1 M elements, block 64
    mn = -rng.uniform(1, 3, (512, 32, 1)).astype(np.float32)
    mx = rng.uniform(1, 3, (512, 32, 1)).astype(np.float32)
    delta = ((mx - mn) / np.float32(15)).astype(np.float32)
    offset = np.round(mn / delta).astype(np.float32)
    out = ops.qdq_broadcast_impl(torch.from_numpy(x).cuda(), *[torch.from_numpy(e).cuda() for e in (mn, mx, delta, offset)])
    exp = oracle.qdq_broadcast(x, mn, mx, delta, offset)
    assert np.array_equal(out.cpu().numpy().view(np.uint32), exp.view(np.uint32))
    with pytest.raises(ValueError):
        ops.qdq_broadcast_impl(torch.zeros(4, 5, device="cuda"), *[torch.zeros(3, 1, device="cuda")] * 4)
    with pytest.raises(RuntimeError):
        ops.qdq_broadcast_impl(torch.zeros(4, 5), *[torch.zeros(4, 1)] * 4)
