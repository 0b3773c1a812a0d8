"""GPU parity: the sm_100a kernels, called through the C ABI, against the CPU oracle on identical inputs.

Bars: integer grid values, histogram counts / PDF state and encodings BIT-EXACT; dequantized
outputs within 1 fp32 ULP (they are in fact bit-exact, and asserted so).
"""
import os

import numpy as np
import pytest
import torch

from tests import kat
from tests.conftest import GOLDEN
from tests.test_oracle_pin import same_f32, ulp_diff

pytestmark = pytest.mark.gpu
VARIANTS = [(0, 0, 0), (1, 0, 0), (1, 1, 0), (1, 0, 1)]


@pytest.fixture(scope="module")
def ops():
    from aimet_b200 import ops as o
    return o


def dev(a, dtype=torch.float32):
    return torch.from_numpy(np.ascontiguousarray(a)).to("cuda").to(dtype)


def host(t):
    return t.detach().float().cpu().numpy()


def make(rng, n, kind="normal"):
    x = rng.standard_normal(n).astype(np.float32)
    if kind == "relu":
        x = np.maximum(x, 0)
    elif kind == "shifted":
        x = x * 2 + 2
    elif kind == "special":
        x[::97] = np.nan
        x[1::193] = np.inf
        x[2::211] = -np.inf
        x[3::89] = 0.0
        x[4::101] = -0.0
    return x.astype(np.float32)


# ---------------------------------------------------------------------------------------------------------------------
# Job 1: QDQ / quantize / STE
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("mn,mx,bw,expected", kat.QDQ_KATS)
def test_qdq_reference_known_answers(ops, mn, mx, bw, expected):
    out = host(ops.qdq_per_tensor_impl(dev(kat.SIX), mn, mx, bw, 0, 0))
    assert ulp_diff(out, np.array(expected, np.float32)).max() <= 4


@pytest.mark.parametrize("mn,mx,bw,signed,expected", kat.GRID_KATS)
def test_grid_reference_known_answers(ops, mn, mx, bw, signed, expected):
    assert host(ops.quantize_to_grid_impl(dev(kat.SIX), mn, mx, bw, 0, signed, 0)).tolist() == expected


def test_qdq_golden_fixtures(ops):
    g = np.load(os.path.join(GOLDEN, "qdq.npz"))
    for k in range(int(g["count"])):
        mn, mx, bw = g[f"meta{k}"]
        x = dev(g[f"x{k}"])
        assert same_f32(host(ops.qdq_per_tensor_impl(x, mn, mx, int(bw), 0, 0)), g[f"qdq{k}"]), k
        assert same_f32(host(ops.quantize_to_grid_impl(x, mn, mx, int(bw), 0, False, 0)), g[f"grid_u{k}"]), k
        assert same_f32(host(ops.quantize_to_grid_impl(x, mn, mx, int(bw), 0, True, 0)), g[f"grid_s{k}"]), k


@pytest.mark.parametrize("n", [0, 1, 3, 4, 5, 1023, 1024, 4097, 262144 + 7, 3_000_001])
@pytest.mark.parametrize("bw", [4, 8, 16])
def test_qdq_fp32_vs_oracle(ops, oracle, n, bw):
    rng = np.random.default_rng(n + bw)
    x = make(rng, n, "special" if n > 1000 else "normal") * np.float32(rng.uniform(0.1, 5))
    for (mn, mx) in ((-1.7, 2.9), (-2.0, 2.0), (0.0, 3.0)):
        out = host(ops.qdq_per_tensor_impl(dev(x), mn, mx, bw, 0, 0))
        exp = oracle.qdq(x, mn, mx, bw)
        assert same_f32(out, exp)                       # bit-exact (bar: 1 ULP)
        grid = host(ops.quantize_to_grid_impl(dev(x), mn, mx, bw, 0, True, 0))
        assert same_f32(grid, oracle.quantize(x, mn, mx, bw, True))   # integer grid: bit-exact


def test_qdq_misaligned_views(ops, oracle):
    rng = np.random.default_rng(1)
    base = make(rng, 10007)
    t = dev(base)
    for off in (1, 2, 3, 5):
        v = t[off:]                                   # contiguous but not 16-byte aligned
        assert v.data_ptr() % 16 != 0
        assert same_f32(host(ops.qdq_per_tensor_impl(v, -1.2, 0.8, 8, 0, 0)), oracle.qdq(base[off:], -1.2, 0.8, 8))
    tb = dev(base, torch.bfloat16)
    vb = tb[1:]
    xb = host(vb)
    exp = dev(oracle.qdq(xb, -1.2, 0.8, 8)).to(torch.bfloat16)
    assert torch.equal(ops.qdq_per_tensor_impl(vb, -1.2, 0.8, 8, 0, 0), exp)


@pytest.mark.parametrize("n", [1, 7, 8, 9, 4096, 100003])
def test_qdq_bf16_vs_oracle(ops, oracle, n):
    """bf16 semantics = widen to fp32, QDQ, round to bf16 (tensor_quantizer.py:1129-1136)."""
    rng = np.random.default_rng(n)
    xb = dev(make(rng, n) * 3, torch.bfloat16)
    x32 = host(xb)
    for bw in (4, 8, 16):
        exp = dev(oracle.qdq(x32, -2.5, 4.0, bw)).to(torch.bfloat16)
        out = ops.qdq_per_tensor_impl(xb, -2.5, 4.0, bw, 0, 0)
        assert out.dtype == torch.bfloat16 and torch.equal(out, exp)


def test_qdq_extreme_values(ops, oracle):
    x = np.array([3.4e38, -3.4e38, 1e-45, -1e-45, 0.0, -0.0, 1.0, np.nan, np.inf, -np.inf], np.float32)
    for (mn, mx) in ((-3.4e38, 3.4e38), (-1e-30, 1e-30), (0.0, 3.4e38), (-1.0, 1.0)):
        for bw in (4, 8, 16):
            assert same_f32(host(ops.qdq_per_tensor_impl(dev(x), mn, mx, bw, 0, 0)), oracle.qdq(x, mn, mx, bw))


def test_qdq_device_resident_encoding(ops, oracle):
    rng = np.random.default_rng(4)
    x = make(rng, 50001)
    e = oracle.fill_encoding_info(8, -1.1, 2.3)
    enc4 = torch.tensor([e[0], e[1], e[2], e[3]], dtype=torch.float32, device="cuda")
    assert same_f32(host(ops.qdq_per_tensor_dev_impl(dev(x), enc4, 0, 0)), oracle.qdq(x, -1.1, 2.3, 8))


def test_stochastic_rounding_distribution(ops, oracle):
    """ROUND_STOCHASTIC is only defined distributionally (clock()/rand() seeded in the reference)."""
    x = np.full(400000, 0.3, np.float32)
    e = oracle.fill_encoding_info(8, 0.0, 25.5)          # delta = 0.1, grid points at multiples of 0.1
    out = host(ops.qdq_per_tensor_impl(dev(x), 0.0, 25.5, 8, 1, 1234))
    lo, hi = np.float32(e[2] * 3), np.float32(e[2] * 4)
    x2 = np.full(400000, 0.33, np.float32)
    out2 = host(ops.qdq_per_tensor_impl(dev(x2), 0.0, 25.5, 8, 1, 99))
    frac_up = np.mean(np.isclose(out2, hi))
    assert set(np.unique(out2)).issubset({lo, hi}) and abs(frac_up - 0.3) < 0.01
    assert np.abs(out.mean() - 0.3) < 1e-3
    assert not np.array_equal(out2, host(ops.qdq_per_tensor_impl(dev(x2), 0.0, 25.5, 8, 1, 100)))   # seed matters


def test_per_channel_golden(ops):
    g = np.load(os.path.join(GOLDEN, "per_channel.npz"))
    for k in range(int(g["count"])):
        c, per, _ = g[f"geom{k}"]
        params = dev(g[f"params{k}"].reshape(-1))
        out = ops.qdq_per_channel_impl(dev(g[f"x{k}"]), params, int(c), int(per), 0, 0)
        assert same_f32(host(out), g[f"out{k}"]), k


@pytest.mark.parametrize("c,per", [(64, 147), (256, 576), (2048, 1), (1000, 2048), (7, 4608), (3, 100001), (4100, 3)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_per_channel_vs_oracle(ops, oracle, c, per, dtype):
    rng = np.random.default_rng(c * 31 + per)
    x = (make(rng, c * per) * np.float32(2)).reshape(c, per)
    xd = dev(x.reshape(-1), dtype)
    x32 = host(xd)
    mins = -np.abs(x).max(axis=1).astype(np.float64) * rng.uniform(0.3, 1.0, c)
    maxs = np.abs(x).max(axis=1).astype(np.float64) * rng.uniform(0.3, 1.0, c)
    for bw in (4, 8):
        p = oracle.per_channel_prepare(mins, maxs, bw)
        host_params = ops.per_channel_params(list(mins), list(maxs), bw)
        assert same_f32(host_params.numpy(), np.concatenate(p))
        exp = oracle.qdq_per_channel(x32, c, per, *p)
        out = ops.qdq_per_channel_impl(xd, host_params.cuda(), c, per, 0, 0)
        if dtype == torch.bfloat16:
            assert torch.equal(out, dev(exp).to(torch.bfloat16))
        else:
            assert same_f32(host(out), exp)
    # channel index wraps: (i / per) % C with more elements than C*per (the reference's formula, trim_functions.cpp:703)
    if c * per < 200000 and dtype == torch.float32:
        x2 = np.concatenate([x32, x32])
        p = oracle.per_channel_prepare(mins, maxs, 8)
        out = ops.qdq_per_channel_impl(dev(x2), dev(np.concatenate(p)), c, per, 0, 0)
        assert same_f32(host(out), oracle.qdq_per_channel(x2, c, per, *p))


@pytest.mark.parametrize("n", [0, 1, 5, 4096, 1_000_003])
def test_ste_backward(ops, oracle, n):
    rng = np.random.default_rng(n)
    x, g = make(rng, n, "special" if n > 100 else "normal") * 2, make(rng, n)
    if n > 100:
        g[::17] = np.inf
    out = ops.ste_bwd_impl(dev(x), dev(g), -1.3, 0.77)
    assert same_f32(host(out), oracle.ste_bwd(x, g, -1.3, 0.77))
    xb, gb = dev(x, torch.bfloat16), dev(g, torch.bfloat16)
    outb = ops.ste_bwd_impl(xb, gb, -1.3, 0.77)
    exp = dev(oracle.ste_bwd(host(xb), host(gb), -1.3, 0.77)).to(torch.bfloat16)
    assert torch.equal(outb.view(torch.int16), exp.view(torch.int16)) or same_f32(host(outb), host(exp))


def test_ste_backward_per_channel(ops, oracle):
    rng = np.random.default_rng(8)
    for (c, per) in ((64, 147), (5, 1), (3, 10001)):
        x, g = make(rng, c * per) * 2, make(rng, c * per)
        mins = (-np.abs(rng.standard_normal(c))).astype(np.float32)
        maxs = np.abs(rng.standard_normal(c)).astype(np.float32)
        out = ops.ste_bwd_per_channel_impl(dev(x), dev(g), dev(mins), dev(maxs), c, per)
        assert same_f32(host(out), oracle.ste_bwd_per_channel(x, g, c, per, mins, maxs))


def test_autograd_matches_reference_ste(ops, oracle):
    x = torch.randn(4099, device="cuda", requires_grad=True)
    y = torch.ops.aimet_b200.qdq_per_tensor(x, -0.9, 1.1, 8)
    g = torch.randn_like(y)
    y.backward(g)
    assert same_f32(host(x.grad), oracle.ste_bwd(host(x), host(g), np.float32(-0.9), np.float32(1.1)))


# ---------------------------------------------------------------------------------------------------------------------
# Job 2: statistics
# ---------------------------------------------------------------------------------------------------------------------
def new_state(n=1):
    from aimet_b200.state import StateArena
    return StateArena.for_device(torch.device("cuda", torch.cuda.current_device())).allocate(n)


@pytest.mark.parametrize("n", [1, 2, 5, 4095, 4096, 4097, 70001, 2_500_003])
@pytest.mark.parametrize("kind", ["normal", "relu", "shifted", "special"])
def test_tfe_stats_state_is_bit_exact(ops, oracle, n, kind):
    from oracle.bindings import OracleTfe
    rng = np.random.default_rng(n)
    blk = new_state()
    o = OracleTfe(oracle)
    for b in range(3):
        x = make(rng, n, kind) * np.float32(rng.uniform(0.5, 1.5))
        ops.stats_update_impl(dev(x), blk.arena, blk.first, ops.QUANTIZATION_TF_ENHANCED, None, 0)
        o.update(x)
    rec = blk.read()[0]
    assert rec["stats_updated"] == 1 and rec["ticket"] == 0 and not rec["hist"].any()
    h = o.histogram()
    if h is None:
        assert rec["initialized"] == 0
        return
    assert rec["initialized"] == 1 and rec["iterations"] == o.s.iterations
    assert (rec["bucket_size"], rec["pdf_offset"]) == o.bucket_params()
    assert np.array_equal(np.array(blk.histogram(0))[:, 0], h[0])     # xLeft
    assert np.array_equal(rec["pdf"], h[1])                            # running PDF, bit for bit


def test_histogram_counts_are_bit_exact(ops, oracle):
    """Raw per-batch integer counts (via the batch log) against GetHistogram_cpu."""
    from oracle.bindings import OracleTfe
    rng = np.random.default_rng(77)
    blk = new_state()
    o = OracleTfe(oracle)
    log = torch.zeros((4, ops.LOG_WORDS), dtype=torch.int32, device="cuda")
    first = make(rng, 300001, "shifted")
    o.update(first)
    ops.stats_update_impl(dev(first), blk.arena, blk.first, ops.QUANTIZATION_TF_ENHANCED, log, 0)
    bucket, offset = o.bucket_params()
    for slot in (1, 2, 3):
        x = make(rng, 123457 * slot, "special") * np.float32(3.0)     # plenty of out-of-range + NaN/inf samples
        ops.stats_update_impl(dev(x), blk.arena, blk.first, ops.QUANTIZATION_TF_ENHANCED, log, slot)
        exp = oracle.histogram(x, bucket, offset)
        got = log[slot].cpu().numpy().view(np.uint32)
        assert np.array_equal(got[:512], exp)
        assert int(got[512]) == x.size and int(got[513]) == 0
    got0 = log[0].cpu().numpy().view(np.uint32)
    assert np.array_equal(got0[:512], oracle.histogram(first, bucket, offset))


def test_tfe_zero_batches_before_data(ops, oracle):
    from oracle.bindings import OracleTfe
    blk, o = new_state(), OracleTfe(oracle)
    z = np.zeros(1000, np.float32)
    rng = np.random.default_rng(5)
    d = make(rng, 20000, "shifted")
    for x in (z, z, d, z, d):
        ops.stats_update_impl(dev(x), blk.arena, blk.first, ops.QUANTIZATION_TF_ENHANCED, None, 0)
        o.update(x)
    rec = blk.read()[0]
    assert rec["iterations"] == o.s.iterations == 3
    assert np.array_equal(rec["pdf"], o.histogram()[1])


def test_tf_stats_running_min_max(ops, oracle):
    from oracle.bindings import OracleTf
    rng = np.random.default_rng(6)
    blk, o = new_state(), OracleTf(oracle)
    for n in (5, 100000, 33):
        x = make(rng, n, "special") * np.float32(rng.uniform(0.1, 9))
        ops.stats_update_impl(dev(x), blk.arena, blk.first, ops.QUANTIZATION_TF, None, 0)
        o.update(x)
    rec = blk.read()[0]
    assert (rec["run_min"], rec["run_max"]) == (o.s.min, o.s.max)
    for bw in (4, 8, 16):
        for (s, st, u) in VARIANTS:
            if s and st and u:
                continue
            enc, _ = ops.compute_encodings_impl(blk.arena, blk.first, 1, ops.QUANTIZATION_TF, bw, s, st, u)
            assert tuple(enc[0].tolist()[:4]) == o.compute(bw, s, st, u)[:4]


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("segs,seg_len", [(64, 147), (512, 4608), (1000, 2048), (96, 9), (3, 100003), (7, 1)])
def test_segmented_stats_and_encodings(ops, oracle, segs, seg_len, dtype):
    """One launch for all channels of a weight == the reference's per-channel Python loop (tensor_quantizer.py:567-570)."""
    from oracle.bindings import OracleTfe, OracleTf
    rng = np.random.default_rng(segs + seg_len)
    w = (make(rng, segs * seg_len) * np.float32(0.05)).reshape(segs, seg_len)
    w[min(2, segs - 1)] = 0.0                                   # an all-zero channel
    wd = dev(w.reshape(-1), dtype)
    w32 = host(wd).reshape(segs, seg_len)
    for mode, cls in ((ops.QUANTIZATION_TF_ENHANCED, OracleTfe), (ops.QUANTIZATION_TF, OracleTf)):
        blk = new_state(segs)
        ops.stats_update_segmented_impl(wd, blk.arena, blk.first, segs, seg_len, mode)
        enc, _ = ops.compute_encodings_impl(blk.arena, blk.first, segs, mode, 8, True, False, False)
        enc = enc.cpu().numpy()
        check = range(segs) if segs <= 100 else rng.choice(segs, 60, replace=False)
        for s in check:
            o = cls(oracle)
            o.update(w32[s])
            assert tuple(enc[s][:4]) == o.compute(8, 1, 0, 0)[:4], (mode, s)


# ---------------------------------------------------------------------------------------------------------------------
# Job 3: tf_enhanced grid search
# ---------------------------------------------------------------------------------------------------------------------
def test_tfe_reference_fixture_known_answer(ops):
    d = np.load(os.path.join(GOLDEN, "kat_n22.npz"))
    blk = new_state()
    ops.stats_update_impl(dev(d["data4"]), blk.arena, blk.first, ops.QUANTIZATION_TF_ENHANCED, None, 0)
    enc, qdq4 = ops.compute_encodings_impl(blk.arena, blk.first, 1, ops.QUANTIZATION_TF_ENHANCED, 8, 0, 0, 0, True)
    e = enc[0].tolist()
    assert tuple(e[:4]) == tuple(d["enc"][:4]) and e[4] == 8
    assert abs(e[0] + 6.52711) < 1e-3 and abs(e[1] - 8.88412) < 1e-3          # TestTensorQuantizer.cpp:126-127
    five = torch.full((16,), 5.0, device="cuda")
    out = ops.qdq_per_tensor_dev_impl(five, qdq4[0].contiguous(), 0, 0)         # encoding never left the device
    assert same_f32(host(out), d["qdq5"])


def test_tfe_all_zero_known_answer(ops):
    blk = new_state()
    ops.stats_update_impl(torch.zeros(100, device="cuda"), blk.arena, blk.first, ops.QUANTIZATION_TF_ENHANCED, None, 0)
    enc, _ = ops.compute_encodings_impl(blk.arena, blk.first, 1, ops.QUANTIZATION_TF_ENHANCED, 8, 0, 0, 0)
    mn, mx, delta, offset, bw = enc[0].tolist()
    k = kat.TFE_ALL_ZERO
    assert abs(mn - k["min"]) < k["tol"] and abs(mx - k["max"]) < k["tol"] and offset == k["offset"] and bw == 8


def test_no_stats_gives_zero_encoding(ops):
    blk = new_state()
    enc, _ = ops.compute_encodings_impl(blk.arena, blk.first, 1, ops.QUANTIZATION_TF_ENHANCED, 8, 0, 0, 0)
    assert enc[0].tolist() == [0, 0, 0, 0, 0]


def test_tfe_encodings_golden(ops):
    g = np.load(os.path.join(GOLDEN, "analyzers.npz"))
    for k in range(int(g["count"])):
        tfe, tf = new_state(), new_state()
        for i in range(int(g[f"nbatch{k}"])):
            x = dev(g[f"batch{k}_{i}"])
            ops.stats_update_impl(x, tfe.arena, tfe.first, ops.QUANTIZATION_TF_ENHANCED, None, 0)
            ops.stats_update_impl(x, tf.arena, tf.first, ops.QUANTIZATION_TF, None, 0)
        assert np.array_equal(tfe.read()[0]["pdf"], g[f"pdf{k}"]), k
        assert np.array_equal(np.array(tfe.histogram(0))[:, 0], g[f"xleft{k}"]), k
        j = 0
        for bw in (4, 8, 16):
            for (s, st, u) in VARIANTS:
                e1, _ = ops.compute_encodings_impl(tfe.arena, tfe.first, 1, ops.QUANTIZATION_TF_ENHANCED, bw, s, st, u)
                e2, _ = ops.compute_encodings_impl(tf.arena, tf.first, 1, ops.QUANTIZATION_TF, bw, s, st, u)
                assert tuple(e1[0].tolist()[:4]) == tuple(g[f"tfe{k}"][j][:4]), (k, bw, s, st, u)
                assert tuple(e2[0].tolist()[:4]) == tuple(g[f"tf{k}"][j][:4]), (k, bw, s, st, u)
                j += 1


def test_tfe_encodings_random_vs_oracle(ops, oracle):
    from oracle.bindings import OracleTfe
    rng = np.random.default_rng(21)
    n_q = 48
    blk = new_state(n_q)
    oracles = []
    for q in range(n_q):
        o = OracleTfe(oracle)
        for _ in range(int(rng.integers(1, 4))):
            x = (make(rng, int(rng.integers(10, 40000)), ["normal", "relu", "shifted"][q % 3]) *
                 np.float32(10 ** rng.uniform(-3, 3))).astype(np.float32)
            ops.stats_update_impl(dev(x), blk.arena, blk.first + q, ops.QUANTIZATION_TF_ENHANCED, None, 0)
            o.update(x)
        oracles.append(o)
    for bw in (4, 8, 16):
        for (s, st, u) in VARIANTS:
            enc, qdq4 = ops.compute_encodings_impl(blk.arena, blk.first, n_q, ops.QUANTIZATION_TF_ENHANCED, bw, s, st,
                                                   u, True)
            enc, qdq4 = enc.cpu().numpy(), qdq4.cpu().numpy()
            for q, o in enumerate(oracles):
                exp = o.compute(bw, s, st, u)
                assert tuple(enc[q][:4]) == exp[:4], (q, bw, s, st, u)
                full = oracle.fill_encoding_info(bw, exp[0], exp[1])
                assert same_f32(qdq4[q], np.array(full[:4], np.float32))


# ---------------------------------------------------------------------------------------------------------------------
# multi-GPU merge primitives (single device: the collective itself is covered by the gloo tests)
# ---------------------------------------------------------------------------------------------------------------------
def test_ordered_replay_reproduces_sequential_pdf(ops, oracle):
    from oracle.bindings import OracleTfe
    rng = np.random.default_rng(31)
    n_q, n_b = 5, 6
    batches = [[make(rng, int(rng.integers(1000, 30000)), "relu" if q % 2 else "shifted") for q in range(n_q)]
               for _ in range(n_b)]
    batches[0][1][:] = 0                                   # quantizer 1: all-zero first batch (skipped by the reference)
    seq = [OracleTfe(oracle) for _ in range(n_q)]
    for b in range(n_b):
        for q in range(n_q):
            seq[q].update(batches[b][q])
    # "two ranks": rank r logs batches b = r, r+2, ... into its own log; ranges come from the first non-zero batch
    blk = new_state(n_q)
    minmax = torch.zeros((n_q, 2), device="cuda")
    for q in range(n_q):
        first = next(b for b in range(n_b) if batches[b][q].any())
        minmax[q, 0], minmax[q, 1] = float(batches[first][q].min()), float(batches[first][q].max())
    ops.stats_init_range_impl(blk.arena, blk.first, n_q, minmax)
    log = torch.zeros((n_b, n_q, ops.LOG_WORDS), dtype=torch.int32, device="cuda")
    for b in range(n_b):
        for q in range(n_q):
            if b == 0 and q == 1:
                continue                                   # the skipped all-zero batch keeps count 0 in the log
            ops.stats_update_impl(dev(batches[b][q]), blk.arena, blk.first + q, ops.QUANTIZATION_TF_ENHANCED,
                                  log, b * n_q + q)
    offsets = torch.arange(n_b, device="cuda", dtype=torch.int64) * (n_q * ops.LOG_WORDS)
    ops.stats_fold_batches_impl(blk.arena, blk.first, n_q, log, offsets)
    rec = blk.read()
    for q in range(n_q):
        assert np.array_equal(rec[q]["pdf"], seq[q].histogram()[1]), q
        assert rec[q]["iterations"] == seq[q].s.iterations
        enc, _ = ops.compute_encodings_impl(blk.arena, blk.first + q, 1, ops.QUANTIZATION_TF_ENHANCED, 8, 0, 0, 0)
        assert tuple(enc[0].tolist()[:4]) == seq[q].compute(8)[:4]


# ---------------------------------------------------------------------------------------------------------------------
# full-size properties (BASELINE.json sizes; the oracle is too slow to run these element by element)
# ---------------------------------------------------------------------------------------------------------------------
def test_large_tensor_properties(ops, oracle):
    n = 32 * 64 * 112 * 112                                 # the largest ResNet activation: 25.7 M elements
    x = torch.randn(n, device="cuda") * 2 + 2
    y = ops.qdq_per_tensor_impl(x, -4.0, 8.0, 8, 0, 0)
    assert torch.equal(ops.qdq_per_tensor_impl(y, -4.0, 8.0, 8, 0, 0), y)              # idempotent
    e = oracle.fill_encoding_info(8, -4.0, 8.0)
    grid = ops.quantize_to_grid_impl(x, -4.0, 8.0, 8, 0, False, 0)
    assert grid.min() >= 0 and grid.max() <= 255 and torch.equal(grid, grid.round())
    assert torch.equal((grid + np.float32(e[3])) * np.float32(e[2]), y)                # dequantize(grid) == QDQ
    sl = slice(1_000_000, 1_050_000)
    assert same_f32(host(y[sl]), oracle.qdq(host(x[sl]), -4.0, 8.0, 8))                # spot parity
    blk = new_state()
    log = torch.zeros((1, ops.LOG_WORDS), dtype=torch.int32, device="cuda")
    ops.stats_update_impl(x, blk.arena, blk.first, ops.QUANTIZATION_TF_ENHANCED, log, 0)
    counts = log[0, :512].to(torch.int64)
    assert int(counts.sum()) == n                                                      # every sample lands in range
    rec = blk.read()[0]
    assert abs(rec["pdf"].sum() - 1.0) < 1e-12
    # linearity of the histogram: counts(a ++ b) == counts(a) + counts(b) under a frozen range
    half = n // 2
    log2 = torch.zeros((2, ops.LOG_WORDS), dtype=torch.int32, device="cuda")
    ops.stats_update_impl(x[:half], blk.arena, blk.first, ops.QUANTIZATION_TF_ENHANCED, log2, 0)
    ops.stats_update_impl(x[half:], blk.arena, blk.first, ops.QUANTIZATION_TF_ENHANCED, log2, 1)
    assert torch.equal(log2[0, :512].to(torch.int64) + log2[1, :512].to(torch.int64), counts)
    mn, mx = float(x.min()), float(x.max())
    tf = new_state()
    ops.stats_update_impl(x, tf.arena, tf.first, ops.QUANTIZATION_TF, None, 0)
    r = tf.read()[0]
    assert (r["run_min"], r["run_max"]) == (mn, mx)


@pytest.mark.parametrize("whole_vectors", [False, True])
def test_per_channel_more_than_2_31_elements(ops, oracle, whole_vectors):
    """The reference's `int` element counts overflow at 2^31 elements; the C ABI takes int64_t and the fast per-channel
    kernel keeps only the tile base in 64 bits (ragged channel length), the bf16 run kernel goes in slices of whole channels
    (channel length a whole number of vectors). Spot-checked against the oracle at the start, across 2^31 and at the end."""
    c = 2050
    per = (2**31 + 2**20) // c + 3            # ragged channel length, total just above 2^31
    if whole_vectors:
        per = (per + 7) // 8 * 8
    n = c * per
    assert n > 2**31
    g = torch.Generator(device="cuda").manual_seed(11)
    x = torch.empty(n, device="cuda", dtype=torch.bfloat16)
    chunk = 2**28
    for s in range(0, n, chunk):
        x[s:s + chunk] = torch.randn(min(chunk, n - s), device="cuda", generator=g).to(torch.bfloat16)
    rng = np.random.default_rng(3)
    mins = -rng.uniform(0.5, 3.0, c)
    maxs = rng.uniform(0.5, 3.0, c)
    params = ops.per_channel_params(list(mins), list(maxs), 8)
    out = ops.qdq_per_channel_impl(x, params.cuda(), c, per, 0, 0)
    p = oracle.per_channel_prepare(mins, maxs, 8)
    for start in (0, per - 100, 2**31 - 5000, 2**31 + 12345, n - 7000):
        stop = min(start + 6000, n)
        idx = np.arange(start, stop, dtype=np.int64)
        ch = (idx // per) % c
        xs = x[start:stop].float().cpu().numpy()
        exp = np.empty_like(xs)
        for cc in np.unique(ch):
            m = ch == cc
            exp[m] = oracle.qdq_per_channel(xs[m], 1, int(m.sum()), *[np.ascontiguousarray(a[cc:cc + 1]) for a in p])
        got = out[start:stop].float().cpu().numpy()
        assert np.array_equal(got, torch.from_numpy(exp).to(torch.bfloat16).float().numpy()), start


def test_tfe_stats_many_batches_of_mixed_sizes(ops, oracle):
    """40 updates of one record back to back, sizes from a handful of samples (16 CTAs) to several million (148 CTAs), fp32
    and bf16 alternating, with a read of the parked batch in the middle: every launch hands its bookkeeping over to the
    next one through the record (early ticket, parked fold), and the final PDF must be the reference's bit for bit."""
    from oracle.bindings import OracleTfe
    rng = np.random.default_rng(2024)
    blk = new_state()
    o = OracleTfe(oracle)
    sizes = [3, 700_001, 17, 4_000_003, 1, 65_536, 2_097_152, 33, 1_000_000, 5] * 4
    for b, n in enumerate(sizes):
        x = make(rng, n, "shifted" if b % 3 else "normal") * np.float32(rng.uniform(0.3, 1.2))
        if b % 2:
            xd = dev(x, torch.bfloat16)
            x = host(xd)
        else:
            xd = dev(x)
        ops.stats_update_impl(xd, blk.arena, blk.first, ops.QUANTIZATION_TF_ENHANCED, None, 0)
        o.update(x)
        if b == 17:
            assert np.array_equal(blk.read()[0]["pdf"], o.histogram()[1])
    rec = blk.read()[0]
    assert rec["ticket"] == 0 and rec["iterations"] == o.s.iterations
    assert np.array_equal(rec["pdf"], o.histogram()[1])


def test_hist_timer_hook(ops):
    """ab_debug_hist_timer: every histogram launch records {first CTA start, last CTA end, input bytes}; switched off again
    it leaves later launches alone."""
    from aimet_b200 import _lib
    L = _lib.load()
    blk = new_state()
    x = torch.randn(3_000_000, device="cuda")
    ops.stats_update_impl(x, blk.arena, blk.first, ops.QUANTIZATION_TF_ENHANCED, None, 0)      # fixes the range
    slots = torch.zeros((8, 3), dtype=torch.int64, device="cuda")
    slots[:, 0] = torch.iinfo(torch.int64).max
    L.ab_debug_hist_timer(slots.data_ptr(), 8)
    for n in (3_000_000, 1_000_000):
        ops.stats_update_impl(x[:n], blk.arena, blk.first, ops.QUANTIZATION_TF_ENHANCED, None, 0)
        ops.stats_update_impl(x[:n].bfloat16(), blk.arena, blk.first, ops.QUANTIZATION_TF_ENHANCED, None, 0)
    torch.cuda.synchronize()
    assert L.ab_debug_hist_timer(None, 0) == 4
    ops.stats_update_impl(x, blk.arena, blk.first, ops.QUANTIZATION_TF_ENHANCED, None, 0)
    torch.cuda.synchronize()
    rows = slots.cpu().tolist()
    assert [r[2] for r in rows[:4]] == [12_000_000, 6_000_000, 4_000_000, 2_000_000]
    for start, end, _ in rows[:4]:
        assert 0 < end - start < 5_000_000          # nanoseconds
    assert rows[4] == [torch.iinfo(torch.int64).max, 0, 0]


def test_histogram_more_than_2_31_elements(ops):
    """One statistics call on 2^31 + 2^20 bf16 samples (the reference's `int` counts stop at 2^31 - 1): the logged counts are
    the sum of the counts of four pieces binned separately (linearity under a frozen range), nothing is lost, and the
    element count comes back in two 32-bit words."""
    n = 2**31 + 2**20
    g = torch.Generator(device="cuda").manual_seed(23)
    x = torch.empty(n, device="cuda", dtype=torch.bfloat16)
    chunk = 2**28
    for s in range(0, n, chunk):
        x[s:s + chunk] = (torch.randn(min(chunk, n - s), device="cuda", generator=g) * 1.5 + 0.25).to(torch.bfloat16)
    blk = new_state()
    ops.stats_init_range_impl(blk.arena, blk.first, 1, torch.tensor([[-8.0, 8.0]], dtype=torch.float32, device="cuda"))
    log = torch.zeros((6, ops.LOG_WORDS), dtype=torch.int32, device="cuda")
    # pieces first: the first of them certifies the bf16 bin formula, the rest and the whole tensor use it
    bounds = [0, 2**29 + 8, 2**30 + 2**29, 2**31 - 16, n]
    for k in range(4):
        ops.stats_update_impl(x[bounds[k]:bounds[k + 1]], blk.arena, blk.first, ops.QUANTIZATION_TF_ENHANCED, log, k)
    ops.stats_update_impl(x, blk.arena, blk.first, ops.QUANTIZATION_TF_ENHANCED, log, 4)
    assert blk.read_raw()[0]["bf16_formula"] == 1
    got = log.cpu().numpy().view(np.uint32).astype(np.int64)
    assert np.array_equal(got[4, :512], got[:4, :512].sum(axis=0))
    assert int(got[4, :512].sum()) == n                                   # range (-24, 24): every sample is counted
    assert int(got[4, 512]) + (int(got[4, 513]) << 32) == n
    assert int(got[4, :512].max()) < 2**31
