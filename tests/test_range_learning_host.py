"""The range-learning host layer (aimet_b200.quantsim.learned_grid + QuantizationSimModel) against the reference's own
Python: tests/golden/range_learning_sim_*.json were produced by the reference's unmodified QuantizationSimModel /
LearnedGridQuantWrapper / QuantizeDequantizeFunc (make_range_learning_sim_golden.py). Here the mirror runs the same
seeded flow on CPU tensors with the CPU oracles injected as the native op and as the quantize-dequantize function.

Exact: wrapper replacement, the initial `<name>_encoding_min/max` parameters, the encodings exported after calibration.
To rounding (the training step contains torch.sum reductions and convolutions whose summation order depends on the
machine): loss, outputs, encoding gradients, encodings after the optimizer step.
"""
import json
import os

import pytest
import torch

from tests.conftest import forward_fingerprint, sim_golden
from tests.golden.make_range_learning_cases import SIM_CASES, sim_inputs, sim_model


@pytest.fixture()
def oracle_backends(oracle):
    from aimet_b200.quantsim import learned_grid, tensor_quantizer
    from tests.oracle_backend import OracleTensorQuantizer, oracle_learned_grid_qdq
    prev_op = tensor_quantizer._set_op_class_for_testing(OracleTensorQuantizer)
    prev_fn = learned_grid.set_qdq_function(oracle_learned_grid_qdq)
    yield
    tensor_quantizer._set_op_class_for_testing(prev_op)
    learned_grid.set_qdq_function(prev_fn)


def encoding_params(model):
    return {n: p for n, p in model.named_parameters() if n.endswith("_encoding_min") or n.endswith("_encoding_max")}


def compact(act, par):
    return json.loads(json.dumps({"activation_encodings": act, "param_encodings": {k: v[:3] for k, v in par.items()},
                                  "param_channels": {k: len(v) for k, v in par.items()}}, sort_keys=True))


def assert_encodings_close(mine, gold, rel):
    assert mine["param_channels"] == gold["param_channels"]
    for section in ("activation_encodings", "param_encodings"):
        assert set(mine[section]) == set(gold[section])

    def walk(a, b, path):
        if isinstance(b, dict):
            assert set(a) == set(b), path
            for k in b:
                walk(a[k], b[k], path + (k,))
        elif isinstance(b, list):
            assert len(a) == len(b), path
            for i, (x, y) in enumerate(zip(a, b)):
                walk(x, y, path + (i,))
        elif isinstance(b, float):
            assert a == pytest.approx(b, rel=rel, abs=rel * 1e-3), path
        else:
            assert a == b, path
    walk(mine, gold, ())


def run_flow(name, device="cpu"):
    from aimet_b200.quantsim import QuantizationSimModel, QuantScheme
    from aimet_b200.quantsim import config as qconfig
    arch, cfg, scheme, shape = SIM_CASES[name]
    model = sim_model(arch).to(device)
    x, x2, target = sim_inputs(shape)
    fingerprint = forward_fingerprint(model.cpu(), x) if device == "cpu" else None
    x, x2, target = (t.to(device) for t in (x, x2, target))
    schemes = {"tf": QuantScheme.training_range_learning_with_tf_init,
               "tf_enhanced": QuantScheme.training_range_learning_with_tf_enhanced_init}
    sim = QuantizationSimModel(model, dummy_input=x, quant_scheme=schemes[scheme], default_output_bw=8,
                               default_param_bw=8,
                               config_file=qconfig.DEFAULT_CONFIG_PER_CHANNEL if cfg else None)

    def calib(m, _):
        m(x)
        m(x2)

    sim.compute_encodings(calib, None)
    res = {"wrapper_types": sorted({type(m).__name__ for m in sim.model.modules()
                                    if type(m).__name__.endswith("QuantWrapper")})}
    res["initial_params"] = {n: p.detach().cpu().tolist() for n, p in encoding_params(sim.model).items()}
    res["encodings_after_calibration"] = compact(*sim.get_activation_param_encodings())
    sim.model.eval()
    opt = torch.optim.SGD(sim.model.parameters(), lr=1e-3)
    out = sim.model(x)
    loss = torch.nn.functional.mse_loss(out, target)
    loss.backward()
    res["loss"] = float(loss.detach())
    res["output_head"] = out.detach().reshape(-1)[:8].cpu().tolist()
    res["grads"] = {n: (p.grad.cpu().tolist() if p.grad is not None else None)
                    for n, p in encoding_params(sim.model).items()}
    res["weight_grad_norms"] = {n: float(p.grad.norm()) for n, p in sim.model.named_parameters()
                                if p.grad is not None and not n.endswith(("_encoding_min", "_encoding_max"))}
    opt.step()
    with torch.no_grad():
        out2 = sim.model(x)
    res["output2_head"] = out2.reshape(-1)[:8].cpu().tolist()
    res["encodings_after_step"] = compact(*sim.get_activation_param_encodings())
    sim.forward_fingerprint = fingerprint
    return sim, res


def check_against_golden(res, gold, rel):
    assert res["wrapper_types"] == gold["wrapper_types"] == ["LearnedGridQuantWrapper"]
    assert res["initial_params"] == gold["initial_params"]
    assert res["encodings_after_calibration"] == gold["encodings_after_calibration"]
    assert res["loss"] == pytest.approx(gold["loss"], rel=rel)
    assert res["output_head"] == pytest.approx(gold["output_head"], rel=rel, abs=rel)
    assert set(res["grads"]) == set(gold["grads"])
    for n, g in gold["grads"].items():
        if g is None:
            assert res["grads"][n] is None, n
            continue
        mine = torch.tensor(res["grads"][n])
        ref = torch.tensor(g)
        scale = float(ref.abs().max()) + 1e-6
        assert float((mine - ref).abs().max()) <= rel * 20 * scale, n
    assert set(res["weight_grad_norms"]) == set(gold["weight_grad_norms"])
    for n, v in gold["weight_grad_norms"].items():
        assert res["weight_grad_norms"][n] == pytest.approx(v, rel=rel * 20, abs=1e-7), n
    assert res["output2_head"] == pytest.approx(gold["output2_head"], rel=rel * 10, abs=rel * 10)
    assert_encodings_close(res["encodings_after_step"], gold["encodings_after_step"], rel * 10)


@pytest.mark.parametrize("name", list(SIM_CASES))
def test_range_learning_sim_reproduces_reference_python(oracle_backends, name):
    sim, res = run_flow(name)
    gold = sim_golden(f"range_learning_sim_{name}.json", "make_range_learning_sim_golden.py", [], sim.forward_fingerprint)
    check_against_golden(res, gold, rel=1e-5)


def test_learned_grid_quantizer_interface(oracle_backends):
    from aimet_b200 import libpymo
    sim, _ = run_flow("resnet18_default_tf")
    wrapper = sim.model.conv1
    q = wrapper.param_quantizers["weight"]
    assert q.channel_axis == 0 and q.name == "weight" and q.wrapper_ref is wrapper
    enc = q.encoding
    assert isinstance(enc, libpymo.TfEncoding) and enc.bw == 8
    assert enc.min == pytest.approx(-enc.max)                      # symmetric weights are kept strictly symmetric inside
    eff = q.get_effective_encoding()
    assert eff.min == pytest.approx(enc.min - enc.delta)           # ... and exported with the extra bin below
    assert "LearnedGrid TensorQuantizer" in str(q)
    q.freeze_encoding()
    assert not wrapper.weight_encoding_min.requires_grad and q.is_encoding_frozen
    with pytest.raises(RuntimeError):
        q.encoding = enc
    out_q = wrapper.output_quantizers[0]
    bad = libpymo.TfEncoding()
    bad.bw = 4
    if out_q.enabled:
        with pytest.raises(RuntimeError):
            out_q.encoding = bad
    with pytest.raises(RuntimeError):
        sim.compute_encodings(lambda m, _: None, None)             # the encodings are parameters now


def test_load_encodings_into_range_learning_sim(oracle_backends, tmp_path):
    """reference qc_quantize_op.py:499-677: LearnedGridQuantWrapper inherits import_*_encodings, so load_encodings /
    load_and_freeze_encodings work on a sim whose wrappers have already been replaced (v1/quantsim.py:1696-1775)."""
    sim, _ = run_flow("resnet18_default_tf")
    act, par = sim.get_activation_param_encodings()
    params_before = {n: p.detach().clone() for n, p in encoding_params(sim.model).items()}
    # shift every encoding; the exported (effective) encoding of a symmetric weight has one extra bin below (min - delta)
    doc = json.loads(json.dumps({"activation_encodings": act, "param_encodings": par}))
    for sec in doc["activation_encodings"].values():
        for io in sec.values():
            for e in io.values():
                e["min"], e["max"] = e["min"] * 0.5, e["max"] * 0.5
                e["scale"] = (e["max"] - e["min"]) / 255
                e["offset"] = round(e["min"] / e["scale"])
    sim.load_encodings(doc, strict=True, partial=True, requires_grad=False, allow_overwrite=True)
    after = encoding_params(sim.model)
    changed = [n for n in params_before if n.startswith("relu") or ".relu" in n or "output0" in n]
    assert changed
    moved = 0
    for n, p in after.items():
        if "output" in n or "input" in n:
            if p is not None and n in params_before and not torch.equal(p.detach(), params_before[n]):
                moved += 1
                assert p.detach().abs().max() <= params_before[n].abs().max() * 0.5 + 1e-6
        assert p is None or not p.requires_grad               # requires_grad=False reached the wrapper parameters
    assert moved > 0
    # allow_overwrite=False freezes: a second load leaves the parameters alone
    sim.load_encodings(doc, strict=True, partial=True, requires_grad=None, allow_overwrite=False)
    frozen = {n: p.detach().clone() for n, p in encoding_params(sim.model).items() if p is not None}
    assert all(q.is_encoding_frozen for _, w in sim.quant_wrappers() for q in w.output_quantizers if q.enabled)
    sim.load_encodings({"activation_encodings": act, "param_encodings": par})
    for n, p in encoding_params(sim.model).items():
        if p is not None:
            assert torch.equal(p.detach(), frozen[n]), n
    # the file route, and a forward still runs
    with open(tmp_path / "e.json", "w") as f:
        json.dump(doc, f)
    sim2, _ = run_flow("resnet18_default_tf")
    sim2.load_and_freeze_encodings(str(tmp_path / "e.json"))
    x, _, _ = sim_inputs(SIM_CASES["resnet18_default_tf"][3])
    with torch.no_grad():
        assert torch.isfinite(sim2.model(x)).all()
