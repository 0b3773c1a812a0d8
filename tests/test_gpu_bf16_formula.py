"""bf16 histogram statistics: the certified one-FFMA bin index (ab_stats_state.bf16_formula, csrc/stats.cu).

Once the range is frozen, the first statistics call on a large bf16 tensor checks nine (scale, shift) candidates against
the reference's  round(x / bucket - offset)  for all 65536 bf16 patterns and records the first exact one; later calls bin
with it. The tests drive every bf16 bit pattern -- NaN, +-inf, denormals, +-0 included -- through both forms and compare
the raw counts with the oracle's GetHistogram_cpu (bit-exact), for ranges of every shape: around zero, one-sided,
far from zero with buckets much finer than the bf16 spacing, tiny, huge.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

PATTERNS = np.arange(65536, dtype=np.uint16)
REPEAT = 32                                    # 2 Mi samples: above the certification threshold (2^20)


@pytest.fixture(scope="module")
def ops():
    from aimet_b200 import ops as o
    return o


def new_state(n=1):
    from aimet_b200.state import StateArena
    return StateArena.for_device(torch.device("cuda", torch.cuda.current_device())).allocate(n)


def bf16_tensor(patterns):
    return torch.from_numpy(np.ascontiguousarray(patterns).view(np.int16)).cuda().view(torch.bfloat16)


def widen(patterns):
    return (patterns.astype(np.uint32) << 16).view(np.float32)


def all_patterns(rng):
    p = np.tile(PATTERNS, REPEAT)
    rng.shuffle(p)
    return p


RANGES = [(-1.0, 1.0), (0.0, 6.25), (-0.37, 2.9), (100.0, 101.0), (-3e-3, 2e-3), (-70000.0, 5.0), (0.0, 0.013),
          (-1e-30, 1e-30), (3.0, 3.0), (-2.5e38, 2.5e38), (1e-3, 1e4), (-512.0, -511.0)]


def run_fixed_range(ops, oracle, mn, mx, rng, calls=3):
    """Freeze the range from (mn, mx), then bin every bf16 pattern `calls` times; returns the final record."""
    blk = new_state()
    ops.stats_init_range_impl(blk.arena, blk.first, 1, torch.tensor([[mn, mx]], dtype=torch.float32, device="cuda"))
    rec = blk.read_raw()[0]
    assert rec["initialized"] == 1 and rec["bf16_formula"] == 0
    bucket, offset = float(rec["bucket_size"]), float(rec["pdf_offset"])
    log = torch.zeros((calls, ops.LOG_WORDS), dtype=torch.int32, device="cuda")
    seen = []
    for slot in range(calls):
        p = all_patterns(rng)
        ops.stats_update_impl(bf16_tensor(p), blk.arena, blk.first, ops.QUANTIZATION_TF_ENHANCED, log, slot,
                              flags=ops.STATS_RANGE_FIXED if slot else 0)
        rec = blk.read_raw()[0]
        seen.append(int(rec["bf16_formula"]))
        got = log[slot].cpu().numpy().view(np.uint32)
        exp = oracle.histogram(widen(p), np.float32(bucket), np.float32(offset))
        assert np.array_equal(got[:512], exp), (mn, mx, slot, seen)
        assert int(got[512]) == p.size
        assert rec["bf16_fail_mask"] == 0 and rec["ticket"] == 0
    # the first call examines, every later call sees the verdict
    assert seen[0] in (1, -1) and seen[1:] == [seen[0]] * (calls - 1)
    return rec


@pytest.mark.parametrize("mn,mx", RANGES)
def test_every_bf16_pattern_lands_in_the_reference_bin(ops, oracle, mn, mx):
    rng = np.random.default_rng(int(abs(mn) * 1000 + abs(mx)) % 2**31)
    run_fixed_range(ops, oracle, mn, mx, rng)


def test_formula_is_certified_for_ordinary_ranges(ops, oracle):
    """The fast form must actually be in use where it matters (otherwise the test above only exercises the fallback)."""
    rng = np.random.default_rng(5)
    certified = 0
    cases = [(-1.0, 1.0), (0.0, 6.25), (-0.37, 2.9), (0.0, 0.013), (-4.2, 3.9), (0.0, 17.5), (-0.02, 0.05), (0.0, 1.0)]
    for mn, mx in cases:
        rec = run_fixed_range(ops, oracle, mn, mx, rng, calls=2)
        if rec["bf16_formula"] == 1:
            certified += 1
            c, b = np.float32(rec["bf16_scale"]), np.float32(rec["bf16_shift"])
            bucket, offset = np.float32(rec["bucket_size"]), np.float32(rec["pdf_offset"])
            # the recorded pair is one of the nine candidates around (1 / bucket, 0.5 - offset)
            assert abs(int(c.view(np.int32)) - int((np.float32(1) / bucket).view(np.int32))) <= 1
            assert abs(int(b.view(np.int32)) - int((np.float32(0.5) - offset).view(np.int32))) <= 1
    assert certified >= len(cases) - 1, certified


def test_random_ranges(ops, oracle):
    rng = np.random.default_rng(11)
    verdicts = []
    for it in range(24):
        kind = it % 4
        if kind == 0:
            mn, mx = 0.0, abs(rng.normal()) * 10 ** rng.uniform(-2, 2)
        elif kind == 1:
            a = abs(rng.normal()) * 10 ** rng.uniform(-2, 2)
            mn, mx = -a * rng.uniform(0.5, 1.5), a
        elif kind == 2:
            mn = rng.normal() * 5
            mx = mn + abs(rng.normal()) * 3 + 0.1
        else:
            mn, mx = -abs(rng.normal()) * 10 ** rng.uniform(-3, 1), abs(rng.normal()) * 10 ** rng.uniform(-3, 1)
        verdicts.append(int(run_fixed_range(ops, oracle, float(mn), float(mx), rng, calls=2)["bf16_formula"]))
    assert verdicts.count(1) >= 20, verdicts


@pytest.mark.parametrize("kind", ["normal", "relu", "shifted"])
def test_natural_flow_state_is_bit_exact(ops, oracle, kind):
    """No injected range, no batch log: min/max + first histogram, then the parked-fold steady state, on 3 Mi bf16 samples."""
    from oracle.bindings import OracleTfe
    rng = np.random.default_rng(3)
    blk = new_state()
    o = OracleTfe(oracle)
    n = 3 * 1024 * 1024 + 5
    for b in range(4):
        x = rng.standard_normal(n).astype(np.float32) * np.float32(rng.uniform(0.5, 1.5))
        if kind == "relu":
            x = np.maximum(x, 0)
        elif kind == "shifted":
            x = x * 2 + 2
        xb = torch.from_numpy(x).cuda().to(torch.bfloat16)
        ops.stats_update_impl(xb, blk.arena, blk.first, ops.QUANTIZATION_TF_ENHANCED, None, 0)
        o.update(xb.float().cpu().numpy())
    raw = blk.read_raw()[0]
    assert raw["bf16_formula"] == 1 and raw["bf16_fail_mask"] == 0
    rec = blk.read()[0]
    h = o.histogram()
    assert rec["initialized"] == 1 and rec["iterations"] == o.s.iterations
    assert (rec["bucket_size"], rec["pdf_offset"]) == o.bucket_params()
    assert np.array_equal(rec["pdf"], h[1])
    # unaligned base and a sub-vector tail go through the exact sequence inside the same launch
    y = torch.from_numpy(rng.standard_normal(n + 1).astype(np.float32)).cuda().to(torch.bfloat16)[1:]
    ops.stats_update_impl(y, blk.arena, blk.first, ops.QUANTIZATION_TF_ENHANCED, None, 0)
    o.update(y.float().cpu().numpy())
    assert np.array_equal(blk.read()[0]["pdf"], o.histogram()[1])


def test_small_tensors_and_fp32_never_examine(ops):
    for dtype, n in ((torch.bfloat16, 500_000), (torch.float32, 4_000_000)):
        blk = new_state()
        for _ in range(3):
            x = torch.randn(n, device="cuda").to(dtype)
            ops.stats_update_impl(x, blk.arena, blk.first, ops.QUANTIZATION_TF_ENHANCED, None, 0)
        assert blk.read_raw()[0]["bf16_formula"] == 0


def test_reset_forgets_the_formula(ops, oracle):
    rng = np.random.default_rng(2)
    blk = new_state()
    for scale in (1.0, 37.0):
        ops.stats_reset_impl(blk.arena, blk.first, 1)
        assert blk.read_raw()[0]["bf16_formula"] == 0
        from oracle.bindings import OracleTfe
        o = OracleTfe(oracle)
        for _ in range(3):
            x = (rng.standard_normal(2_000_000).astype(np.float32) * np.float32(scale))
            xb = torch.from_numpy(x).cuda().to(torch.bfloat16)
            ops.stats_update_impl(xb, blk.arena, blk.first, ops.QUANTIZATION_TF_ENHANCED, None, 0)
            o.update(xb.float().cpu().numpy())
        assert blk.read_raw()[0]["bf16_formula"] == 1
        assert np.array_equal(blk.read()[0]["pdf"], o.histogram()[1])
