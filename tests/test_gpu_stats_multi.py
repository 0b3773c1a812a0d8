"""ab_stats_update_multi (many tensors, one histogram launch + one fold launch) against the CPU oracle and against the
one-launch-per-tensor path it replaces: records byte-identical, log entries identical, fold order respected.

Reference semantics: UpdatePdf / GetHistogram_cpu, DlQ/src/math_functions.cpp:243-288, 367-384 -- one call per tensor.
"""
import numpy as np
import pytest
import torch

from tests.test_gpu_parity import dev, make, new_state

pytestmark = pytest.mark.gpu

# sizes in elements: one vector, ragged tails, exactly one tile (8192 fp32), tile + 1, many tiles, > 148 * 4 tiles
SIZES = [4, 8, 13, 4099, 8192, 8193, 70001, 300000, 1_000_003, 5_000_011]


@pytest.fixture(scope="module")
def ops():
    from aimet_b200 import ops as o
    return o


def _first_batches(ops, oracle, blk, n_records, rng, dtype):
    """Fix every record's range with an ordinary per-tensor call; returns the oracle twins."""
    from oracle.bindings import OracleTfe
    twins = []
    for r in range(n_records):
        x = make(rng, 5000 + 37 * r, "shifted") * np.float32(1 + 0.3 * r)
        xd = dev(x, dtype)
        ops.stats_update_impl(xd, blk.arena, blk.first + r, ops.QUANTIZATION_TF_ENHANCED, None, 0)
        o = OracleTfe(oracle)
        o.update(xd.float().cpu().numpy())
        twins.append(o)
    return twins


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_multi_update_equals_oracle_and_single_tensor_path(ops, oracle, dtype):
    rng = np.random.default_rng(5)
    n_records = 6
    blk_multi, blk_single = new_state(n_records), new_state(n_records)
    twins = _first_batches(ops, oracle, blk_multi, n_records, np.random.default_rng(1), dtype)
    _first_batches(ops, oracle, blk_single, n_records, np.random.default_rng(1), dtype)
    # a table in which records repeat and interleave: 0 1 2 0 3 1 4 5 0 2
    order = [0, 1, 2, 0, 3, 1, 4, 5, 0, 2]
    kinds = ["normal", "relu", "shifted", "special"]
    tensors = []
    for k, r in enumerate(order):
        size = max(SIZES[k % len(SIZES)], 4 if dtype == torch.float32 else 8)      # at least one 128-bit vector
        x = make(rng, size, kinds[k % 4]) * np.float32(1 + 0.3 * r)
        tensors.append(dev(x, dtype))
    for rounds in range(2):              # second round: the scratch rows must have come back zeroed
        counts = torch.zeros((len(order), ops.LOG_WORDS), dtype=torch.int32, device="cuda") if rounds == 0 else counts
        ops.stats_update_multi_impl(tensors, order, blk_multi.arena, blk_multi.first, counts)
        for t, r in zip(tensors, order):
            ops.stats_update_impl(t, blk_single.arena, blk_single.first + r, ops.QUANTIZATION_TF_ENHANCED, None, 0)
            twins[r].update(t.float().cpu().numpy())
        assert not counts.any()
    multi, single = blk_multi.read(), blk_single.read()
    for r in range(n_records):
        h = twins[r].histogram()
        assert multi[r]["iterations"] == single[r]["iterations"] == twins[r].s.iterations
        assert np.array_equal(multi[r]["pdf"], h[1]), r                       # bit for bit against the oracle
        assert np.array_equal(multi[r]["pdf"], single[r]["pdf"]), r           # and against the path it replaces
        assert multi[r]["stats_updated"] == 1 and multi[r]["pending"] == 0 and not multi[r]["hist"].any()


def test_multi_update_large_table_many_ctas(ops, oracle):
    """More tensors than a CTA ever sees, sizes that put several segment changes into one CTA's chunk."""
    from oracle.bindings import OracleTfe
    rng = np.random.default_rng(9)
    n = 100
    blk = new_state(n)
    twins = _first_batches(ops, oracle, blk, n, np.random.default_rng(2), torch.float32)
    sizes = [int(rng.integers(4, 200_000)) for _ in range(n)]
    sizes[17], sizes[63] = 9_000_001, 3_333_333
    tensors = [dev(make(rng, s, "shifted") * np.float32(1 + 0.3 * r)) for r, s in enumerate(sizes)]
    counts = torch.zeros((n, ops.LOG_WORDS), dtype=torch.int32, device="cuda")
    ops.stats_update_multi_impl(tensors, list(range(n)), blk.arena, blk.first, counts)
    rec = blk.read()
    for r in range(n):
        twins[r].update(tensors[r].cpu().numpy())
        assert np.array_equal(rec[r]["pdf"], twins[r].histogram()[1]), (r, sizes[r])
        assert rec[r]["iterations"] == 2


def test_multi_update_log_only_rows_are_the_single_tensor_log_entries(ops, oracle):
    rng = np.random.default_rng(11)
    blk_a, blk_b = new_state(3), new_state(3)
    _first_batches(ops, oracle, blk_a, 3, np.random.default_rng(3), torch.float32)
    twins = _first_batches(ops, oracle, blk_b, 3, np.random.default_rng(3), torch.float32)
    before = blk_a.read()
    order = [2, 0, 1, 2]
    tensors = [dev(make(rng, s, "special") * np.float32(2.0)) for s in (12345, 8192 * 3, 77, 2_000_001)]
    log_a = torch.zeros((4, ops.LOG_WORDS), dtype=torch.int32, device="cuda")
    log_b = torch.zeros((4, ops.LOG_WORDS), dtype=torch.int32, device="cuda")
    ops.stats_update_multi_impl(tensors, order, blk_a.arena, blk_a.first, log_a, log_only=True)
    for k, (t, r) in enumerate(zip(tensors, order)):
        ops.stats_update_impl(t, blk_b.arena, blk_b.first + r, ops.QUANTIZATION_TF_ENHANCED, log_b, k)
    assert torch.equal(log_a, log_b)
    for k, (t, r) in enumerate(zip(tensors, order)):
        bucket, offset = twins[r].bucket_params()
        got = log_a[k].cpu().numpy().view(np.uint32)
        assert np.array_equal(got[:512], oracle.histogram(t.cpu().numpy(), bucket, offset))
        assert int(got[512]) == t.numel() and int(got[513]) == 0
    after = blk_a.read()
    assert np.array_equal(before["pdf"], after["pdf"]) and np.array_equal(before["iterations"], after["iterations"])


def test_multi_update_after_a_parked_single_tensor_batch_and_on_an_unset_range(ops, oracle):
    """The single-tensor kernel parks its counts in the record (lazy fold); a multi-tensor call that follows must fold
    them first. A record without a range is left alone."""
    from oracle.bindings import OracleTfe
    rng = np.random.default_rng(13)
    blk = new_state(3)
    twins = _first_batches(ops, oracle, blk, 2, np.random.default_rng(4), torch.float32)
    x1 = make(rng, 50_000, "normal") * np.float32(2.0)
    ops.stats_update_impl(dev(x1), blk.arena, blk.first, ops.QUANTIZATION_TF_ENHANCED, None, 0)     # parked
    twins[0].update(x1)
    assert blk.read_raw()[0]["pending"] == 1
    x2, x3, x4 = (make(rng, n, "shifted") for n in (30_001, 9, 4096))
    counts = torch.zeros((3, ops.LOG_WORDS), dtype=torch.int32, device="cuda")
    ops.stats_update_multi_impl([dev(x2), dev(x3), dev(x4)], [0, 2, 1], blk.arena, blk.first, counts)
    twins[0].update(x2)
    twins[1].update(x4)
    rec = blk.read_raw()
    assert rec[0]["pending"] == 0 and rec[0]["iterations"] == 3 and np.array_equal(rec[0]["pdf"], twins[0].histogram()[1])
    assert rec[1]["iterations"] == 2 and np.array_equal(rec[1]["pdf"], twins[1].histogram()[1])
    assert rec[2]["initialized"] == 0 and rec[2]["iterations"] == 0 and not rec[2]["pdf"].any()
    # and the single-tensor kernel after a multi-tensor call
    x5 = make(rng, 20_000, "relu")
    ops.stats_update_impl(dev(x5), blk.arena, blk.first, ops.QUANTIZATION_TF_ENHANCED, None, 0)
    twins[0].update(x5)
    assert np.array_equal(blk.read()[0]["pdf"], twins[0].histogram()[1])


def test_multi_update_rejects_what_it_cannot_do(ops):
    blk = new_state(1)
    counts = torch.zeros((2, ops.LOG_WORDS), dtype=torch.int32, device="cuda")
    base = torch.zeros(64, device="cuda")
    with pytest.raises(ValueError):
        ops.stats_update_multi_impl([base[1:33]], [0], blk.arena, blk.first, counts)          # not 16-byte aligned
    with pytest.raises(ValueError):
        ops.stats_update_multi_impl([base[:3]], [0], blk.arena, blk.first, counts)            # less than one vector
    with pytest.raises(TypeError):
        ops.stats_update_multi_impl([base, base.to(torch.bfloat16)], [0, 0], blk.arena, blk.first, counts)
    with pytest.raises(ValueError):
        ops.stats_update_multi_impl([base] * 129, [0] * 129, blk.arena, blk.first, counts)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("mode,sym", [(1, True), (1, False), (0, True)])
def test_multi_parameter_refresh_equals_one_call_per_tensor(ops, dtype, mode, sym):
    """ab_stats_refresh_encodings_multi (all weights of a model in a handful of launches) against ab_stats_refresh_encodings
    per weight: records, encodings, per-tensor and per-channel kernel parameters byte-identical. Shapes: per-channel conv
    weights (ragged channel lengths), per-tensor weights below and above the large-segment threshold, a 1-element tail."""
    torch.manual_seed(3)
    shapes = [(64, 147), (1, 4096), (256, 64), (1, 200_001), (96, 1, 9), (1000, 17), (1, 7), (1, 131072), (33, 577)]
    tensors = [(torch.randn(s, device="cuda") * (0.05 + 0.01 * k)).to(dtype).contiguous() for k, s in enumerate(shapes)]
    segs = [s[0] for s in shapes]
    total = sum(segs)
    blk_m, blk_s = new_state(total), new_state(total)
    enc_m = torch.zeros((total, 5), dtype=torch.float64, device="cuda")
    qdq4_m = torch.zeros((total, 4), dtype=torch.float32, device="cuda")
    par_m = torch.zeros(4 * total, dtype=torch.float32, device="cuda")
    for _ in range(2):          # twice: the second call starts from used records
        ops.stats_refresh_multi_impl(tensors, segs, blk_m.arena, blk_m.first, mode, 8, sym, False, False, enc_m, qdq4_m, par_m)
    at = 0
    for t, n in zip(tensors, segs):
        enc, qdq4, params = ops.stats_refresh_encodings_impl(t, blk_s.arena, blk_s.first + at, n, t.numel() // n, mode, 8, sym,
                                                             False, False)
        assert torch.equal(enc, enc_m[at:at + n]), (tuple(t.shape), "encodings")
        if n == 1:
            assert torch.equal(qdq4, qdq4_m[at:at + 1])
        else:
            assert torch.equal(params, par_m[4 * at:4 * (at + n)]), (tuple(t.shape), "per-channel parameters")
        at += n
    a, b = blk_m.read(), blk_s.read()
    for field in ("pdf", "iterations", "initialized", "stats_updated", "x_left0", "bucket_size_d", "bucket_size",
                  "pdf_offset", "run_min", "run_max"):
        assert np.array_equal(a[field], b[field]), field
    # a sub-range refresh (first_record > 0) touches only its own records
    keep = blk_m.read()
    ops.stats_refresh_multi_impl(tensors[2:4], segs[2:4], blk_m.arena, blk_m.first, mode, 8, sym, False, False, enc_m, qdq4_m,
                                 par_m, first_record=segs[0] + segs[1])
    again = blk_m.read()
    assert np.array_equal(keep["pdf"], again["pdf"]) and np.array_equal(keep["iterations"], again["iterations"])
