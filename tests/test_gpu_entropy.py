"""The entropy scheme on the device (ab_entropy_update: min / max, range growth with redistribution, binning) and its host
closing (ab_entropy_compute_encoding) against the oracle -- itself pinned to the reference's C++ in test_entropy_oracle.py.
Bars: the raw histogram (doubles holding integers), its range and iteration count, and every encoding field are BIT-EXACT."""
import numpy as np
import pytest
import torch

from oracle.bindings import OracleEntropy
from tests.golden.make_entropy_cases import NUM_CASES, VARIANTS, batches

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    from aimet_b200 import ops as o
    return o


def new_record(ops):
    from aimet_b200.state import StateArena
    return StateArena.for_device(torch.device("cuda", 0)).allocate(1)


def check(ops, blk, a):
    hist, mn, mx, it = ops.entropy_histogram_impl(blk.arena, blk.first)
    want = a.raw()
    assert (hist is None) == (want[0] is None)
    if hist is not None:
        assert np.array_equal(hist, want[0])
        assert (mn, mx, it) == want[1:]
    for bw, s, st, u in VARIANTS:
        e = ops.entropy_compute_impl(blk.arena, blk.first, bw, s, st, u)
        assert (e.min, e.max, e.delta, e.offset, e.bw) == tuple(a.compute(bw, s, st, u)), (bw, s, st, u)


@pytest.mark.parametrize("case", range(NUM_CASES))
def test_device_statistics_and_encodings_equal_oracle(ops, oracle, case):
    blk, a = new_record(ops), OracleEntropy(oracle)
    check(ops, blk, a)                                   # nothing seen yet: the zero encoding
    for x in batches(case):
        ops.entropy_update_impl(torch.from_numpy(x).cuda(), blk.arena, blk.first)
        a.update(x)
        check(ops, blk, a)                               # after every batch, not only at the end


def test_bf16_large_and_unaligned(ops, oracle):
    g = torch.Generator(device="cuda").manual_seed(3)
    blk, a = new_record(ops), OracleEntropy(oracle)
    for scale, n in ((1.0, 5_000_011), (3.0, 20_000_000), (0.2, 777)):
        x = (torch.randn(n, device="cuda", generator=g) * scale + 0.3).to(torch.bfloat16)
        ops.entropy_update_impl(x, blk.arena, blk.first)
        a.update(x.float().cpu().numpy())
    check(ops, blk, a)
    blk, a = new_record(ops), OracleEntropy(oracle)
    base = torch.randn(100_003, device="cuda", generator=g) * 2
    for view in (base[1:], base[3:50_000], base[:7]):   # 4-byte aligned only, odd lengths
        assert view.data_ptr() % 16 != 0 or view.numel() % 4 != 0
        ops.entropy_update_impl(view, blk.arena, blk.first)
        a.update(view.cpu().numpy())
    check(ops, blk, a)
    blk, a = new_record(ops), OracleEntropy(oracle)     # a reset record starts over
    ops.entropy_update_impl(base, blk.arena, blk.first)
    ops.stats_reset_impl(blk.arena, blk.first, 1)
    ops.entropy_update_impl(base[:1000], blk.arena, blk.first)
    a.update(base[:1000].cpu().numpy())
    check(ops, blk, a)


def test_libpymo_and_torch_extension_surface(oracle):
    """QuantizationMode.QUANTIZATION_ENTROPY through the three Python entry points the reference exposes."""
    from aimet_b200 import libpymo
    from aimet_b200.tensor_quantizer_op import AimetTensorQuantizer
    xs = batches(3)
    a = OracleEntropy(oracle)
    tq = libpymo.TensorQuantizer(libpymo.QuantizationMode.QUANTIZATION_ENTROPY, libpymo.RoundingMode.ROUND_NEAREST)
    ea = libpymo.EncodingAnalyzerForPython(libpymo.QuantizationMode.QUANTIZATION_ENTROPY)
    op = AimetTensorQuantizer(libpymo.QuantizationMode.QUANTIZATION_ENTROPY)
    for x in xs:
        a.update(x)
        tq.updateStats(x, True)
        ea.updateStats(x, True)
        op.updateStats(torch.from_numpy(x).cuda(), True)
    want = a.compute(8, True, False, False)
    e = tq.computeEncoding(8, True)
    assert (e.min, e.max, e.delta, e.offset, e.bw) == want
    e, valid = ea.computeEncoding(8, True, False, False)
    assert valid and (e.min, e.max, e.delta, e.offset, e.bw) == want
    e, valid = op.getEncoding(8, True, False, False)
    assert valid and (e.min, e.max, e.delta, e.offset, e.bw) == want
    op.resetEncodingStats()
    e, valid = op.getEncoding(8, False, False, False)
    assert not valid


def test_throughput_is_reported(ops):
    """Not a bar, a record: two passes over the tensor (min / max, binning) at HBM speed is 2 x 4 B per fp32 element."""
    x = torch.randn(64 * 2**20, device="cuda")
    blk = new_record(ops)
    ops.entropy_update_impl(x, blk.arena, blk.first)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        ops.entropy_update_impl(x, blk.arena, blk.first)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 5
    print(f"entropy updateStats, 256 MB fp32: {ms * 1e3:.1f} us, {2 * x.numel() * 4 / ms / 1e6:.0f} GB/s of tensor reads")
