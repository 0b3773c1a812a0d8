"""The host layer (aimet_b200.quantsim) against the reference's own Python + C++.

The goldens in tests/golden/quantsim_*.json come from the reference's unmodified QuantizationSimModel running on the
reference's unmodified C++ (make_quantsim_golden.py). Here the mirror runs the same models on the same seeded inputs
with the CPU oracle injected as the native op, on CPU tensors, and must reproduce quantizer placement, every encoding
(bit for bit: compared through the canonical JSON's sha256) and the quantized model's output.
"""
import hashlib
import json
import os

import pytest
import torch
import torchvision

from tests.conftest import forward_fingerprint, sim_golden

CASES = {
    "resnet18_default_tfe": (torchvision.models.resnet18, "default", "tf_enhanced", (4, 3, 64, 64)),
    "resnet18_perchannel_tfe": (torchvision.models.resnet18, "per_channel", "tf_enhanced", (2, 3, 64, 64)),
    "resnet18_default_tf": (torchvision.models.resnet18, "default", "tf", (4, 3, 64, 64)),
    "mobilenet_v2_default_tfe": (torchvision.models.mobilenet_v2, "default", "tf_enhanced", (2, 3, 64, 64)),
    "resnet50_perchannel_tfe": (torchvision.models.resnet50, "per_channel", "tf_enhanced", (2, 3, 64, 64)),
    "mobilenet_v2_perchannel_tfe": (torchvision.models.mobilenet_v2, "per_channel", "tf_enhanced", (2, 3, 64, 64)),
    "resnet18_perchannel_tf": (torchvision.models.resnet18, "per_channel", "tf", (2, 3, 64, 64)),
    "vgg11_default_tfe": (torchvision.models.vgg11, "default", "tf_enhanced", (2, 3, 64, 64)),
}


@pytest.fixture()
def oracle_backend(oracle):
    from aimet_b200.quantsim import tensor_quantizer
    from tests.oracle_backend import OracleTensorQuantizer
    prev = tensor_quantizer._set_op_class_for_testing(OracleTensorQuantizer)
    yield
    tensor_quantizer._set_op_class_for_testing(prev)


PERCENTILE_CASE = ("resnet18_percentile", (torchvision.models.resnet18, "default", "percentile", (4, 3, 64, 64)), 99.9)


def build_and_calibrate(name, device="cpu"):
    from aimet_b200.quantsim import QuantizationSimModel
    from aimet_b200.quantsim import config as qconfig
    percentile = None
    if name == PERCENTILE_CASE[0]:
        ctor, cfg, scheme, shape = PERCENTILE_CASE[1]
        percentile = PERCENTILE_CASE[2]
    else:
        ctor, cfg, scheme, shape = CASES[name]
    torch.manual_seed(0)
    model = ctor().eval()
    torch.manual_seed(1)
    x = torch.randn(*shape)
    x2 = torch.randn(*shape) * 1.5
    fingerprint = forward_fingerprint(model, x)
    model, x, x2 = model.to(device), x.to(device), x2.to(device)
    sim = QuantizationSimModel(model, dummy_input=x, quant_scheme=scheme, default_output_bw=8, default_param_bw=8,
                               config_file=qconfig.DEFAULT_CONFIG_PER_CHANNEL if cfg == "per_channel" else None)
    if percentile is not None:
        sim.set_percentile_value(percentile)
    structure = {}
    for mname, w in sim.quant_wrappers():
        structure[mname] = {
            "type": type(w.get_original_module()).__name__,
            "inputs": [bool(q.enabled) for q in w.input_quantizers],
            "outputs": [bool(q.enabled) for q in w.output_quantizers],
            "params": {k: [bool(q.enabled), bool(q.use_symmetric_encodings), type(q).__name__]
                       for k, q in w.param_quantizers.items()},
        }

    def calib(m, _):
        m(x)
        m(x2)

    sim.compute_encodings(calib, None)
    with torch.no_grad():
        out = sim.model(x)
    sim.forward_fingerprint = fingerprint
    return sim, structure, out


@pytest.mark.parametrize("name", list(CASES))
def test_host_layer_reproduces_reference_python(oracle_backend, name):
    sim, structure, out = build_and_calibrate(name)
    golden = sim_golden(f"quantsim_{name}.json", "make_quantsim_golden.py", [name], sim.forward_fingerprint)
    assert structure == golden["structure"]
    act, par = sim.get_activation_param_encodings()
    enc = {"activation_encodings": act, "param_encodings": par}
    canonical = json.dumps(enc, sort_keys=True)
    if "encodings" in golden:
        mine = json.loads(canonical)
        assert mine["activation_encodings"] == golden["encodings"]["activation_encodings"]
        assert mine["param_encodings"] == golden["encodings"]["param_encodings"]
    else:
        assert json.loads(json.dumps(act, sort_keys=True)) == golden["activation_encodings"]
        for k, v in golden["param_encodings_sample"].items():
            assert json.loads(json.dumps(par[k][:3])) == v
    assert hashlib.sha256(canonical.encode()).hexdigest() == golden["sha256"]
    assert hashlib.sha256(out.numpy().tobytes()).hexdigest() == golden["output_sha256"]


def test_percentile_scheme_reproduces_reference_python(oracle_backend):
    """post_training_percentile with set_percentile_value(99.9): activations are clipped at the percentile, parameters
    keep the analyzer's default of 100 (the reference's wrapper only forwards the value to activation quantizers)."""
    sim, _, out = build_and_calibrate(PERCENTILE_CASE[0])
    golden = sim_golden("quantsim_resnet18_percentile.json", "make_percentile_golden.py", ["--sim-only"],
                        sim.forward_fingerprint)
    act, par = sim.get_activation_param_encodings()
    canonical = json.dumps({"activation_encodings": act, "param_encodings": par}, sort_keys=True)
    mine = json.loads(canonical)
    assert mine["activation_encodings"] == golden["encodings"]["activation_encodings"]
    assert mine["param_encodings"] == golden["encodings"]["param_encodings"]
    assert hashlib.sha256(canonical.encode()).hexdigest() == golden["sha256"]
    assert hashlib.sha256(out.numpy().tobytes()).hexdigest() == golden["output_sha256"]
    with pytest.raises(ValueError):
        sim.set_percentile_value(50)


def test_export_files(oracle_backend, tmp_path):
    sim, _, _ = build_and_calibrate("resnet18_default_tfe")
    sim.save_encodings_to_json(str(tmp_path), "enc")
    saved = json.load(open(tmp_path / "enc.json"))
    assert set(saved) == {"activation_encodings", "param_encodings"}
    entry = saved["param_encodings"]["conv1.weight"][0]
    # reference create_encoding_dict (aimet_torch/utils.py:1156-1185): exactly these keys, is_symmetric as a string
    assert set(entry) == {"min", "max", "scale", "offset", "bitwidth", "is_symmetric", "dtype"}
    assert entry["is_symmetric"] == "True" and entry["dtype"] == "int" and isinstance(entry["offset"], int)
    assert saved["activation_encodings"]["relu"]["output"]["0"]["is_symmetric"] == "False"
    sim.export(str(tmp_path), "model")
    exported = json.load(open(tmp_path / "model_torch.encodings"))
    assert exported["version"] == "0.6.1"
    assert exported["param_encodings"] == saved["param_encodings"]
    assert os.path.exists(tmp_path / "model.pth")
    # quantizer_args as extract_global_quantizer_args writes them (aimet_common/quantsim.py:280-311)
    assert exported["quantizer_args"] == {"quant_scheme": "post_training_tf_enhanced", "param_bitwidth": 8,
                                          "activation_bitwidth": 8, "dtype": "int", "is_symmetric": True,
                                          "per_channel_quantization": False}
    # <prefix>.pth is the pickled original model (reference v1/quantsim.py:548), wrappers removed
    restored = torch.load(tmp_path / "model.pth", weights_only=False)
    assert isinstance(restored, torchvision.models.ResNet) and isinstance(restored.conv1, torch.nn.Conv2d)
    # text form is json.dump(sort_keys=True, indent=4), as the reference writes it
    assert open(tmp_path / "enc.json").read() == json.dumps(saved, sort_keys=True, indent=4)


def test_load_encodings_and_checkpoint_round_trip(oracle_backend, tmp_path):
    """Encodings written by one sim load into a fresh one (reference load_encodings, v1/quantsim.py:1696-1757) and give
    the same quantized forward; a pickled checkpoint (save_checkpoint / load_checkpoint) does too."""
    from aimet_b200.quantsim import QuantizationSimModel, load_checkpoint, save_checkpoint
    sim, _, out = build_and_calibrate("resnet18_perchannel_tfe")
    sim.save_encodings_to_json(str(tmp_path), "enc")
    ctor, _, scheme, shape = CASES["resnet18_perchannel_tfe"]
    from aimet_b200.quantsim import config as qconfig
    torch.manual_seed(0)
    model = ctor().eval()
    torch.manual_seed(1)
    x = torch.randn(*shape)
    fresh = QuantizationSimModel(model, dummy_input=x, quant_scheme=scheme, config_file=qconfig.DEFAULT_CONFIG_PER_CHANNEL)
    fresh.load_encodings(str(tmp_path / "enc.json"))
    with torch.no_grad():
        assert torch.equal(fresh.model(x), out)
    a1, p1 = sim.get_activation_param_encodings()
    a2, p2 = fresh.get_activation_param_encodings()
    assert json.dumps(a1, sort_keys=True) == json.dumps(a2, sort_keys=True)
    assert json.dumps(p1, sort_keys=True) == json.dumps(p2, sort_keys=True)
    with pytest.raises(RuntimeError):
        fresh.load_encodings({"param_encodings": {"no.such.weight": p1["conv1.weight"]}, "activation_encodings": {}})
    save_checkpoint(sim, str(tmp_path / "sim.ckpt"))
    restored = load_checkpoint(str(tmp_path / "sim.ckpt"))
    with torch.no_grad():
        assert torch.equal(restored.model(x), out)


def test_scalar_in_dtype_rounds_like_torch():
    """The tensor-free narrowing of an encoding bound (float64 -> float32 -> bf16 / fp16, ties to even, overflow to inf)."""
    import struct

    import numpy as np
    from aimet_b200.quantsim.tensor_quantizer import scalar_in_dtype
    rng = np.random.default_rng(0)
    values = [0.0, -0.0, 1e39, -1e39, 3.4028235e38, 3.4028236e38, 65504.0, 65519.99, 65520.0, 5.96e-8, 2.98e-8, 1e-45,
              7e-46, float("inf"), float("-inf")]
    # float32 bit patterns that are exact bf16 ties (low half 0x8000) and their neighbours, even and odd upper halves
    for hi in (0x3f80, 0x3f81, 0x4000, 0x7f7f, 0x0001, 0xbf80, 0xbf81):
        for lo in (0x7fff, 0x8000, 0x8001):
            values.append(struct.unpack("<f", struct.pack("<I", (hi << 16) | lo))[0])
    values += list(rng.standard_normal(2000) * 10.0 ** rng.uniform(-10, 10, 2000))
    for v in values:
        for dt in (torch.float32, torch.bfloat16, torch.float16):
            ref = float(torch.tensor(float(v), dtype=torch.float32).to(dt))
            got = scalar_in_dtype(v, dt)
            assert struct.pack("<d", ref) == struct.pack("<d", got), (v, dt, ref, got)


def test_load_and_freeze_encodings(oracle_backend, tmp_path):
    """reference v1/quantsim.py:1759-1855: encodings loaded with allow_overwrite=False survive a later compute_encodings;
    set_and_freeze_param_encodings touches parameters only; exclude_param_from_quantization disables by parameter name."""
    from aimet_b200.quantsim import QuantizationSimModel
    sim_a, _, _ = build_and_calibrate("resnet18_default_tfe")
    sim_a.save_encodings_to_json(str(tmp_path), "enc")
    act_a, par_a = sim_a.get_activation_param_encodings()

    torch.manual_seed(0)
    model = torchvision.models.resnet18().eval()
    x = torch.randn(4, 3, 64, 64) * 3.0          # different calibration data: would give different encodings
    sim_b = QuantizationSimModel(model, dummy_input=x, quant_scheme="tf_enhanced")
    sim_b.load_and_freeze_encodings(str(tmp_path / "enc.json"))
    assert all(q.is_encoding_frozen for _, w in sim_b.named_qmodules() for q in w.param_quantizers.values() if q.enabled)
    sim_b.compute_encodings(lambda m, _: m(x), None)
    act_b, par_b = sim_b.get_activation_param_encodings()
    assert json.dumps(act_a, sort_keys=True) == json.dumps(act_b, sort_keys=True)
    assert json.dumps(par_a, sort_keys=True) == json.dumps(par_b, sort_keys=True)
    with torch.no_grad():
        torch.manual_seed(1)
        probe = torch.randn(4, 3, 64, 64)
        assert torch.equal(sim_a.model(probe), sim_b.model(probe))

    sim_c = QuantizationSimModel(model, dummy_input=x, quant_scheme="tf_enhanced")
    sim_c.set_and_freeze_param_encodings(str(tmp_path / "enc.json"))
    sim_c.compute_encodings(lambda m, _: m(x), None)
    act_c, par_c = sim_c.get_activation_param_encodings()
    assert json.dumps(par_a, sort_keys=True) == json.dumps(par_c, sort_keys=True)       # parameters: loaded and kept
    assert json.dumps(act_a, sort_keys=True) != json.dumps(act_c, sort_keys=True)       # activations: recalibrated on x
    # the older parameter-only file layout (reference :1726-1732)
    sim_d = QuantizationSimModel(model, dummy_input=x, quant_scheme="tf_enhanced")
    sim_d.load_encodings(par_a)
    assert json.dumps(sim_d.get_activation_param_encodings()[1], sort_keys=True) == json.dumps(par_a, sort_keys=True)

    assert any(w.param_quantizers["weight"].enabled for w in sim_c.qmodules() if "weight" in w.param_quantizers)
    sim_c.exclude_param_from_quantization("weight")
    assert not any(w.param_quantizers["weight"].enabled for w in sim_c.qmodules() if "weight" in w.param_quantizers)
