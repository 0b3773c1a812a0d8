"""Out-of-bounds guard: every kernel that writes a tensor is run on views carved out of a larger, canary-filled buffer
(odd sizes, aligned and unaligned starts); the canaries on both sides of the output must be untouched and the result must
equal the one computed on a freshly allocated tensor. (compute-sanitizer is not available on the GPU pool.)"""
import pytest
import torch

pytestmark = pytest.mark.gpu
CANARY = 1234.5
PAD = 64   # elements on each side


@pytest.fixture(scope="module")
def ops():
    from aimet_b200 import ops as o
    return o


def carve(n, dtype, offset):
    """A contiguous view of n elements inside a canary-filled buffer, starting `offset` elements after the pad."""
    buf = torch.full((n + 2 * PAD + 8,), CANARY, dtype=dtype, device="cuda")
    return buf, buf[PAD + offset:PAD + offset + n]


def redirect_first_output(monkeypatch_ctx, view):
    """Make the op's FIRST torch.empty_like of the view's shape / dtype return `view` (its output tensor)."""
    real = torch.empty_like
    state = {"used": False}

    def fake(t, **kw):
        if not state["used"] and t.shape == view.shape and t.dtype == view.dtype:
            state["used"] = True
            return view
        return real(t, **kw)
    monkeypatch_ctx.setattr(torch, "empty_like", fake)


def canaries_intact(buf, n, offset):
    head = buf[:PAD + offset]
    tail = buf[PAD + offset + n:]
    return bool((head == CANARY).all()) and bool((tail == CANARY).all())


SIZES = [1, 7, 255, 4099, 65537, 1 << 20]


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("offset", [0, 1, 3])
def test_per_tensor_kernels_stay_in_bounds(ops, dtype, offset, monkeypatch):
    g = torch.Generator(device="cuda").manual_seed(0)
    for n in SIZES:
        x = (torch.randn(n, device="cuda", generator=g) * 2).to(dtype)
        grad = torch.randn(n, device="cuda", generator=g).to(dtype)
        ref_qdq = ops.qdq_per_tensor_impl(x, -1.5, 2.5, 8)
        ref_q = ops.quantize_to_grid_impl(x, -1.5, 2.5, 8, 0, True)
        ref_ste = ops.ste_bwd_impl(x, grad, -1.5, 2.5)
        mn, mx = torch.tensor([-1.5], device="cuda", dtype=dtype), torch.tensor([2.5], device="cuda", dtype=dtype)
        ref_lg = ops.lg_qdq_fwd_impl(x, mn.clone(), mx.clone(), 8, ops.LG_ASYMMETRIC)
        ref_lgb = ops.lg_qdq_bwd_impl(x, grad, mn, mx, 8, ops.LG_ASYMMETRIC)

        # route the op's output allocation into the canary buffer
        for name, call, ref in (
            ("qdq", lambda: ops.qdq_per_tensor_impl(x, -1.5, 2.5, 8), ref_qdq),
            ("quantize", lambda: ops.quantize_to_grid_impl(x, -1.5, 2.5, 8, 0, True), ref_q),
            ("ste", lambda: ops.ste_bwd_impl(x, grad, -1.5, 2.5), ref_ste),
            ("lg_fwd", lambda: ops.lg_qdq_fwd_impl(x, mn.clone(), mx.clone(), 8, ops.LG_ASYMMETRIC), ref_lg),
            ("lg_bwd", lambda: ops.lg_qdq_bwd_impl(x, grad, mn, mx, 8, ops.LG_ASYMMETRIC)[0], ref_lgb[0]),
        ):
            buf, view = carve(n, dtype, offset)
            if name == "ste" and view.data_ptr() % 16:
                # the STE entry point takes 16-byte aligned tensors only (include/aimet_b200.h) and says so
                with monkeypatch.context() as m:
                    redirect_first_output(m, view)
                    with pytest.raises(ValueError):
                        call()
                continue
            with monkeypatch.context() as m:
                redirect_first_output(m, view)
                out = call()
            assert out.data_ptr() == view.data_ptr(), name
            torch.cuda.synchronize()
            assert canaries_intact(buf, n, offset), (name, n, offset, dtype)
            assert torch.equal(out.float().nan_to_num(), ref.float().nan_to_num()), (name, n, offset, dtype)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_per_channel_kernels_stay_in_bounds(ops, dtype, monkeypatch):
    g = torch.Generator(device="cuda").manual_seed(1)
    for shape in [(3, 5), (16, 27), (64, 3, 7, 7), (7, 1031), (1000, 1), (130, 4, 2)]:
        c = shape[0]
        n = 1
        for d in shape:
            n *= d
        x = (torch.randn(shape, device="cuda", generator=g)).to(dtype)
        grad = torch.randn(shape, device="cuda", generator=g).to(dtype)
        params = ops.per_channel_params([-1.0 - 0.01 * i for i in range(c)], [1.0 + 0.01 * i for i in range(c)], 8).cuda()
        mn = torch.tensor([-1.0 - 0.01 * i for i in range(c)], device="cuda").to(dtype)
        mx = torch.tensor([1.0 + 0.01 * i for i in range(c)], device="cuda").to(dtype)
        per = n // c
        ref_pc = ops.qdq_per_channel_impl(x, params, c, per)
        ref_ste = ops.ste_bwd_per_channel_impl(x, grad, params[:c].contiguous(), params[c:2 * c].contiguous(), c, per)
        ref_lg = ops.lg_qdq_fwd_impl(x, mn.clone(), mx.clone(), 8, ops.LG_SIGNED_SYMMETRIC)
        ref_lgb = ops.lg_qdq_bwd_impl(x, grad, mn, mx, 8, ops.LG_SIGNED_SYMMETRIC)
        for name, call, ref in (
            ("pc", lambda: ops.qdq_per_channel_impl(x, params, c, per), ref_pc),
            ("ste_pc", lambda: ops.ste_bwd_per_channel_impl(x, grad, params[:c].contiguous(),
                                                            params[c:2 * c].contiguous(), c, per), ref_ste),
            ("lg_fwd_pc", lambda: ops.lg_qdq_fwd_impl(x, mn.clone(), mx.clone(), 8, ops.LG_SIGNED_SYMMETRIC), ref_lg),
            ("lg_bwd_pc", lambda: ops.lg_qdq_bwd_impl(x, grad, mn, mx, 8, ops.LG_SIGNED_SYMMETRIC)[0], ref_lgb[0]),
        ):
            buf = torch.full((n + 2 * PAD,), CANARY, dtype=dtype, device="cuda")
            view = buf[PAD:PAD + n].view(shape)
            with monkeypatch.context() as m:
                redirect_first_output(m, view)
                out = call()
            assert out.data_ptr() == view.data_ptr(), name
            torch.cuda.synchronize()
            assert bool((buf[:PAD] == CANARY).all()) and bool((buf[PAD + n:] == CANARY).all()), (name, shape, dtype)
            assert torch.equal(out.float(), ref.float()), (name, shape, dtype)
        # the encoding-gradient outputs are exactly num_channel long
        assert ref_lgb[1].numel() == c and ref_lgb[2].numel() == c


def test_statistics_do_not_write_outside_their_record(ops):
    """A statistics update and a grid search touch one ab_stats_state record and nothing next to it."""
    from aimet_b200.state import StateArena
    blk = StateArena.for_device(torch.device("cuda", 0)).allocate(3)
    torch.cuda.synchronize()
    before = blk.arena[blk.first * ops.STATE_BYTES:(blk.first + 3) * ops.STATE_BYTES].clone()
    x = torch.randn(100003, device="cuda") * 3 + 1
    for _ in range(3):
        ops.stats_update_impl(x, blk.arena, blk.first + 1, ops.QUANTIZATION_TF_ENHANCED, None, 0)
    ops.stats_update_segmented_impl(torch.randn(1, 4097, device="cuda"), blk.arena, blk.first + 1, 1, 4097,
                                    ops.QUANTIZATION_TF_ENHANCED)
    ops.compute_encodings_impl(blk.arena, blk.first + 1, 1, ops.QUANTIZATION_TF_ENHANCED, 8, False, False, False)
    torch.cuda.synchronize()
    after = blk.arena[blk.first * ops.STATE_BYTES:(blk.first + 3) * ops.STATE_BYTES]
    sb = ops.STATE_BYTES
    assert torch.equal(before[:sb], after[:sb]) and torch.equal(before[2 * sb:], after[2 * sb:])
    assert not torch.equal(before[sb:2 * sb], after[sb:2 * sb])
