"""The other BASELINE.json configs as parity cases (the headline config is bench.py's workload):
  [2] MobileNet-v2 quantization-aware training step, W8A8, bf16 weights and activations (QDQ forward + STE backward)
  [3] Llama-2-7B-shaped linear weights: W4 per-channel symmetric QDQ (bf16) + tf activation encodings (A16)
Each is run with the CUDA ops and with the CPU oracle injected under the same host layer, on the same device tensors.
"""
import numpy as np
import pytest
import torch
import torchvision

pytestmark = pytest.mark.gpu


def deterministic():
    torch.backends.cudnn.deterministic = True
    torch.backends.cudnn.benchmark = False
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False


def test_mobilenet_v2_bf16_qat_step(oracle):
    from aimet_b200 import AimetTensorQuantizer
    from aimet_b200.quantsim import QuantizationSimModel, tensor_quantizer
    from tests.oracle_backend import OracleTensorQuantizer

    def step(factory):
        prev = tensor_quantizer._set_op_class_for_testing(factory)
        try:
            deterministic()
            torch.manual_seed(0)
            model = torchvision.models.mobilenet_v2(num_classes=10).to(torch.bfloat16).cuda()
            torch.manual_seed(1)
            x = torch.randn(4, 3, 64, 64, device="cuda", dtype=torch.bfloat16)
            sim = QuantizationSimModel(model.eval(), dummy_input=x, quant_scheme="tf_enhanced", default_output_bw=8,
                                       default_param_bw=8)
            sim.compute_encodings(lambda m, _: m(x), None)
            sim.model.train()                        # weight encodings are re-derived on every forward in training
            out = sim.model(x)
            out.float().square().mean().backward()
            grads = {n: p.grad.clone() for n, p in sim.model.named_parameters() if p.grad is not None}
            act, par = sim.get_activation_param_encodings()
            return out.detach(), grads, (act, par)
        finally:
            tensor_quantizer._set_op_class_for_testing(prev)

    out_n, grads_n, enc_n = step(AimetTensorQuantizer)
    out_o, grads_o, enc_o = step(OracleTensorQuantizer)
    assert out_n.dtype == torch.bfloat16
    assert enc_n == enc_o
    assert torch.equal(out_n, out_o)
    assert grads_n.keys() == grads_o.keys() and len(grads_n) > 100
    for k in grads_n:
        assert torch.equal(grads_n[k], grads_o[k]), k


@pytest.mark.parametrize("shape", [(4096, 4096), (11008, 4096), (4096, 11008)])
def test_llama_shaped_w4_per_channel_weight_qdq(oracle, shape):
    """W4 per-channel (axis 0) symmetric weight QDQ in bf16 + the per-channel tf_enhanced and tf encodings behind it."""
    from aimet_b200 import ops
    from aimet_b200.state import StateArena
    from oracle.bindings import OracleTf, OracleTfe
    torch.manual_seed(shape[0] + shape[1])
    w = (torch.randn(shape, device="cuda") * 0.02).to(torch.bfloat16)
    w32 = w.float().cpu().numpy()
    c, per = shape
    rng = np.random.default_rng(0)
    sample = rng.choice(c, 24, replace=False)
    for mode, cls in ((ops.QUANTIZATION_TF, OracleTf), (ops.QUANTIZATION_TF_ENHANCED, OracleTfe)):
        blk = StateArena.for_device(w.device).allocate(c)
        ops.stats_update_segmented_impl(w, blk.arena, blk.first, c, per, mode)
        enc, _ = ops.compute_encodings_impl(blk.arena, blk.first, c, mode, 4, True, False, False)
        enc = enc.cpu().numpy()
        for ch in sample:
            o = cls(oracle)
            o.update(w32[ch])
            assert tuple(enc[ch][:4]) == o.compute(4, 1, 0, 0)[:4], (mode, ch)
        params = ops.per_channel_params(list(enc[:, 0]), list(enc[:, 1]), 4)
        out = ops.qdq_per_channel_impl(w, params.cuda(), c, per, 0, 0)
        p = oracle.per_channel_prepare(enc[:, 0].copy(), enc[:, 1].copy(), 4)
        rows = np.sort(sample)[:8]
        for ch in rows:           # the oracle is slow: whole channels, spot-checked
            exp = oracle.qdq_per_channel(w32[ch], 1, per, *[np.ascontiguousarray(a[ch:ch + 1]) for a in p])
            got = out[ch].float().cpu().numpy()
            exp_bf16 = torch.from_numpy(exp).to(torch.bfloat16).float().numpy()
            assert np.array_equal(got, exp_bf16), (mode, ch)
        grid = torch.unique((out[rows[0]].float() / float(p[2][rows[0]])).round())
        assert grid.numel() <= 16                      # 4 bits


def test_llama_shaped_a16_tf_activation_encodings(oracle):
    from aimet_b200 import AimetTensorQuantizer, libpymo
    from oracle.bindings import OracleTf
    torch.manual_seed(5)
    for last in (4096, 11008):
        q = AimetTensorQuantizer(libpymo.QuantizationMode.QUANTIZATION_TF)
        o = OracleTf(oracle)
        for _ in range(2):
            x = (torch.randn(8, 2048, last, device="cuda") * 3).to(torch.bfloat16)
            q.updateStats(x, True)
            o.update(x.float().cpu().numpy().reshape(-1))
        enc, valid = q.getEncoding(16, False, False, False)
        assert valid and (enc.min, enc.max, enc.delta, enc.offset) == o.compute(16)[:4]
        y = q.quantizeDequantize(x, enc, libpymo.RoundingMode.ROUND_NEAREST, True)
        sl = x.reshape(-1)[:200000]
        exp = torch.from_numpy(oracle.qdq(sl.float().cpu().numpy(), enc.min, enc.max, 16)).to(torch.bfloat16)
        assert torch.equal(y.reshape(-1)[:200000].cpu(), exp)
