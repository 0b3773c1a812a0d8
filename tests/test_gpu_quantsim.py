"""End to end on the GPU: QuantizationSimModel driven by the CUDA ops vs the same host layer driven by the CPU oracle on
the SAME device tensors (the model forward runs on the GPU in both, so every quantizer sees identical inputs).
Encodings JSON must be identical, byte for byte; the quantized forward output must be bit-identical.
"""
import hashlib
import json

import pytest
import torch
import torchvision

from tests.test_quantsim_host import CASES, build_and_calibrate

pytestmark = pytest.mark.gpu


def run(name, factory):
    from aimet_b200.quantsim import tensor_quantizer
    prev = tensor_quantizer._set_op_class_for_testing(factory)
    try:
        torch.backends.cudnn.deterministic = True
        torch.backends.cudnn.benchmark = False
        torch.backends.cuda.matmul.allow_tf32 = False
        torch.backends.cudnn.allow_tf32 = False
        sim, structure, out = build_and_calibrate(name, device="cuda")
        act, par = sim.get_activation_param_encodings()
        return json.dumps({"activation_encodings": act, "param_encodings": par}, sort_keys=True), structure, out
    finally:
        tensor_quantizer._set_op_class_for_testing(prev)


@pytest.mark.parametrize("name", ["resnet18_default_tfe", "resnet18_perchannel_tfe", "resnet18_default_tf",
                                  "mobilenet_v2_default_tfe", "resnet50_perchannel_tfe", "resnet18_percentile",
                                  "mobilenet_v2_perchannel_tfe", "resnet18_perchannel_tf", "vgg11_default_tfe"])
def test_cuda_ops_and_oracle_give_identical_encodings_json(oracle, name):
    from aimet_b200 import AimetTensorQuantizer
    from tests.oracle_backend import OracleTensorQuantizer
    json_native, struct_native, out_native = run(name, AimetTensorQuantizer)
    json_oracle, struct_oracle, out_oracle = run(name, OracleTensorQuantizer)
    assert struct_native == struct_oracle
    if json_native != json_oracle:
        a, b = json.loads(json_native), json.loads(json_oracle)
        for sect in a:
            for k in a[sect]:
                assert a[sect][k] == b[sect][k], (sect, k)
    assert hashlib.sha256(json_native.encode()).hexdigest() == hashlib.sha256(json_oracle.encode()).hexdigest()
    assert torch.equal(out_native, out_oracle)


def test_qat_step_forward_backward(oracle):
    """MobileNet-v2-style QAT step (BASELINE config 3, small): train mode re-derives weight encodings every forward,
    activations are QDQ'd, gradients pass the straight-through estimator. Checked against the oracle-backed host layer."""
    from aimet_b200 import AimetTensorQuantizer
    from aimet_b200.quantsim import QuantizationSimModel, tensor_quantizer
    from tests.oracle_backend import OracleTensorQuantizer

    def step(factory):
        prev = tensor_quantizer._set_op_class_for_testing(factory)
        try:
            torch.backends.cudnn.deterministic = True
            torch.manual_seed(0)
            model = torch.nn.Sequential(torch.nn.Conv2d(3, 16, 3, padding=1), torch.nn.BatchNorm2d(16), torch.nn.ReLU(),
                                        torch.nn.Conv2d(16, 8, 3, padding=1), torch.nn.ReLU(),
                                        torch.nn.AdaptiveAvgPool2d(1), torch.nn.Flatten(), torch.nn.Linear(8, 4)).cuda()
            x = torch.randn(8, 3, 16, 16, device="cuda")
            sim = QuantizationSimModel(model, dummy_input=x, quant_scheme="tf_enhanced")
            sim.compute_encodings(lambda m, _: m(x), None)
            sim.model.train()
            out = sim.model(x)
            out.square().mean().backward()
            grads = [p.grad.clone() for p in sim.model.parameters() if p.grad is not None]
            return out.detach(), grads
        finally:
            tensor_quantizer._set_op_class_for_testing(prev)

    out_n, grads_n = step(AimetTensorQuantizer)
    out_o, grads_o = step(OracleTensorQuantizer)
    assert torch.equal(out_n, out_o)
    assert len(grads_n) == len(grads_o) > 0
    for a, b in zip(grads_n, grads_o):
        assert torch.equal(a, b)


@pytest.mark.parametrize("name", ["resnet18_default_tfe", "resnet18_perchannel_tfe", "resnet18_default_tf"])
def test_cuda_graph_calibration_equals_eager(name):
    """compute_encodings_for_batches replays the steady-state step from a CUDA graph; its encodings must be identical
    to the eager compute_encodings over the same batches."""
    from aimet_b200.quantsim import QuantizationSimModel
    from aimet_b200.quantsim import config as qconfig
    torch.backends.cudnn.deterministic = True
    torch.backends.cudnn.benchmark = False
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    ctor, cfg, scheme, shape = CASES[name]
    g = torch.Generator().manual_seed(3)
    batches = [(torch.randn(*shape, generator=g) * (1 + 0.1 * i)).cuda() for i in range(6)]
    results = []
    for graphed in (False, True):
        torch.manual_seed(0)
        model = ctor().eval().cuda()
        sim = QuantizationSimModel(model, dummy_input=batches[0], quant_scheme=scheme,
                                   config_file=qconfig.DEFAULT_CONFIG_PER_CHANNEL if cfg == "per_channel" else None)
        if graphed:
            sim.compute_encodings_for_batches(batches, cuda_graph=True)
            from aimet_b200.quantsim import quantsim as qs
            assert qs.LAST_GRAPH_INFO["replays"] == 4 and qs.LAST_GRAPH_INFO["captured_launches"] > 0
        else:
            sim.compute_encodings(lambda m, _: [m(x) for x in batches], None)
        act, par = sim.get_activation_param_encodings()
        results.append(json.dumps({"a": act, "p": par}, sort_keys=True))
    assert results[0] == results[1]


def test_gating_and_clone_elision_changes_nothing():
    """The wrapper skips the parameter-gradient gating where no parameter quantizer is enabled and the input clone where the
    wrapped module cannot write to its input; the reference does both in every wrapper. Outputs and every gradient must be
    the same bits either way (MobileNet-v2 block shapes: in-place ReLU6, residual add, batch norm, dropout)."""
    import torchvision
    from aimet_b200.quantsim import QuantizationSimModel, qc_quantize_op

    def step(legacy):
        qc_quantize_op.ALWAYS_GATE_AND_CLONE = legacy
        try:
            torch.backends.cudnn.deterministic = True
            torch.backends.cudnn.benchmark = False
            torch.manual_seed(0)
            model = torchvision.models.mobilenet_v2(width_mult=0.35, num_classes=10).cuda()
            x = torch.randn(4, 3, 64, 64, device="cuda")
            sim = QuantizationSimModel(model, dummy_input=x, quant_scheme="tf_enhanced")
            with torch.no_grad():
                sim.compute_encodings(lambda m, _: m(x), None)
            sim.model.train()
            xin = x.clone().requires_grad_(True)
            out = sim.model(xin)
            out.square().mean().backward()
            grads = {n: p.grad.clone() for n, p in sim.model.named_parameters() if p.grad is not None}
            return out.detach(), xin.grad.clone(), grads
        finally:
            qc_quantize_op.ALWAYS_GATE_AND_CLONE = False

    out_a, gx_a, grads_a = step(False)
    out_b, gx_b, grads_b = step(True)
    assert torch.equal(out_a, out_b) and torch.equal(gx_a, gx_b)
    assert grads_a.keys() == grads_b.keys() and len(grads_a) > 50
    for n in grads_a:
        assert torch.equal(grads_a[n], grads_b[n]), n


@pytest.mark.parametrize("per_channel", [False, True])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_fused_parameter_refresh_equals_the_three_calls(per_channel, dtype):
    """refresh_encoding_from (one native call) against reset_encoding_stats + update_encoding_stats + compute_encoding:
    same device-resident encodings, same kernel parameters, same quantize-dequantized weight, call after call."""
    from aimet_b200 import libpymo
    from aimet_b200.quantsim import tensor_quantizer as tq
    from aimet_b200.quantsim.defs import QuantScheme

    def make():
        if per_channel:
            q = tq.StaticGridPerChannelQuantizer(8, libpymo.RoundingMode.ROUND_NEAREST, QuantScheme.post_training_tf_enhanced,
                                                 True, 24, True, ch_axis=0)
        else:
            q = tq.StaticGridPerTensorQuantizer(8, libpymo.RoundingMode.ROUND_NEAREST, QuantScheme.post_training_tf_enhanced,
                                                True, True)
        q._lazy_ok = True   # pylint: disable=protected-access
        return q

    fused, plain = make(), make()
    g = torch.Generator(device="cuda").manual_seed(5)
    for step in range(3):
        w = (torch.randn(24, 16, 3, 3, device="cuda", generator=g) * (0.05 + 0.02 * step)).to(dtype)
        assert fused.refresh_encoding_from(w)
        plain.reset_encoding_stats()
        plain.update_encoding_stats(w)
        plain.compute_encoding()
        assert torch.equal(fused._enc_dev, plain._enc_dev)   # pylint: disable=protected-access
        if per_channel:
            assert torch.equal(fused._params_dev, plain._params_dev)   # pylint: disable=protected-access
        else:
            assert torch.equal(fused._qdq4_dev, plain._qdq4_dev)       # pylint: disable=protected-access
        a = fused.quantize_dequantize(w, libpymo.RoundingMode.ROUND_NEAREST)
        b = plain.quantize_dequantize(w, libpymo.RoundingMode.ROUND_NEAREST)
        assert torch.equal(a, b) and not torch.equal(a, w)
        ea, eb = fused.encoding, plain.encoding
        ea, eb = (ea, eb) if isinstance(ea, list) else ([ea], [eb])
        assert [(e.min, e.max, e.delta, e.offset, e.bw) for e in ea] == [(e.min, e.max, e.delta, e.offset, e.bw) for e in eb]
    # quantizers that cannot take the fused route say so
    frozen = make()
    frozen.refresh_encoding_from(w)
    frozen.freeze_encoding()
    assert not frozen.refresh_encoding_from(w)
    unsigned = make()
    unsigned.use_unsigned_symmetric = True
    assert not unsigned.refresh_encoding_from(w)


def test_quantized_forward_is_cuda_graph_capturable():
    """Eval forward of a calibrated sim captured with torch.cuda.graph: replay == eager, bit for bit (the layers launch on
    the capturing stream, allocate through torch, and never synchronise once the encodings are on the device)."""
    import torchvision
    from aimet_b200.quantsim import QuantizationSimModel
    from aimet_b200.quantsim import config as qconfig
    torch.manual_seed(0)
    model = torchvision.models.resnet18().cuda().eval()
    x = torch.randn(4, 3, 64, 64, device="cuda")
    sim = QuantizationSimModel(model, dummy_input=x, quant_scheme="tf_enhanced", config_file=qconfig.DEFAULT_CONFIG_PER_CHANNEL)
    sim.compute_encodings(lambda m, _: m(x), None)
    with torch.no_grad():
        eager = sim.model(x)
        static_x = x.clone()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(2):
                sim.model(static_x)
        torch.cuda.current_stream().wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            static_y = sim.model(static_x)
        for scale in (1.0, 0.5):
            static_x.copy_(x * scale)
            graph.replay()
            torch.cuda.synchronize()
            assert torch.equal(static_y, sim.model(x * scale))
    assert torch.equal(eager, sim.model(x))


def test_capture_forward_helper():
    import torchvision
    from aimet_b200.quantsim import QuantizationSimModel
    torch.manual_seed(0)
    model = torchvision.models.resnet18().cuda().eval()
    x = torch.randn(2, 3, 64, 64, device="cuda")
    sim = QuantizationSimModel(model, dummy_input=x, quant_scheme="tf")
    sim.compute_encodings(lambda m, _: m(x), None)
    fwd = sim.capture_forward(x)
    with torch.no_grad():
        for scale in (1.0, -0.7, 2.5):
            assert torch.equal(fwd(x * scale), sim.model(x * scale))
    with pytest.raises(ValueError):
        fwd(x[:1])
    with pytest.raises(ValueError):
        sim.capture_forward(x.cpu())


def test_parameter_export_prefetch_serves_the_same_dictionaries():
    """During the calibration forwards the per-channel parameter encodings are copied out and turned into dictionaries
    (quantsim._ParamExportPrefetch); the export must hand out exactly what it would have built itself, serve the parked
    dictionaries once only, and never serve them for encodings that were recomputed afterwards."""
    import torchvision
    from aimet_b200.quantsim import QuantizationSimModel
    from aimet_b200.quantsim import config as qconfig
    torch.manual_seed(0)
    model = torchvision.models.resnet18().cuda().eval()
    xs = [torch.randn(4, 3, 64, 64, device="cuda") for _ in range(6)]
    sim = QuantizationSimModel(model, dummy_input=xs[0], quant_scheme="tf_enhanced", config_file=qconfig.DEFAULT_CONFIG_PER_CHANNEL)

    def calibrate(m, _):
        for x in xs:
            m(x)
            torch.cuda.synchronize()          # lets the stream-ordered copy land between two forwards

    sim.compute_encodings(calibrate, None)
    parked = [q for _, w in sim.quant_wrappers() for q in w.param_quantizers.values() if "_export_cache" in q.__dict__]
    assert len(parked) >= 20
    first = json.dumps(sim.get_activation_param_encodings(), sort_keys=True)
    assert not any("_export_cache" in q.__dict__ for q in parked)              # served once
    again = json.dumps(sim.get_activation_param_encodings(), sort_keys=True)   # rebuilt from the device tables
    assert first == again
    # a cache left over from an earlier calibration is not served for recomputed encodings
    sim.compute_encodings(calibrate, None)
    stale = {q: q.__dict__["_export_cache"] for q in parked}
    with torch.no_grad():
        for p in sim.model.parameters():
            p.mul_(1.5)
    sim.compute_encodings(lambda m, _: m(xs[0]), None)                         # one forward: no prefetch this time
    for q, c in stale.items():
        q.__dict__["_export_cache"] = c
    fresh = json.dumps(sim.get_activation_param_encodings(), sort_keys=True)
    assert fresh != first and fresh == json.dumps(sim.get_activation_param_encodings(), sort_keys=True)
