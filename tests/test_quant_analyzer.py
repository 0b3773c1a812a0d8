"""QuantAnalyzer (aimet_b200.quantsim.quant_analyzer) against the reference's own QuantAnalyzer.

CPU: tests/golden/quant_analyzer_*.json were produced by the reference's unmodified QuantAnalyzer / QuantizationSimModel /
C++ (make_quant_analyzer_golden.py); the mirror, driven by the CPU oracle as its native op, must return the same three
sensitivity scores, the same two per-layer sweeps (names in order of occurrence, scores bit for bit), the same encoding
ranges and the same per-layer MSE table, and write the same JSON files.
GPU: the mirror on the CUDA ops against the mirror on the oracle, same device tensors: identical tables.
"""
import json
import os
import sys

import pytest
import torch

from tests.conftest import GOLDEN, forward_fingerprint, sim_golden

sys.path.insert(0, GOLDEN)
from make_quant_analyzer_cases import CASES, callbacks, make_data, make_model  # noqa: E402


def analyze(name, results_dir, device="cpu"):
    from aimet_b200.quantsim import CallbackFunc, QuantAnalyzer
    from aimet_b200.quantsim import config as qconfig
    cfg, scheme, ignore = CASES[name]
    model = make_model()
    batches, target = make_data()
    fingerprint = forward_fingerprint(model, batches[0])
    model = model.to(device)
    batches = [b.to(device) for b in batches]
    fwd, ev = callbacks(batches, target)
    qa = QuantAnalyzer(model, batches[0], CallbackFunc(fwd, None), CallbackFunc(ev, None),
                       modules_to_ignore=[model.conv2] if ignore else None)
    qa.enable_per_layer_mse_loss(batches, 2)
    sim = qa._create_quantsim_and_encodings(scheme, 8, 8, qconfig.DEFAULT_CONFIG_PER_CHANNEL if cfg else None)   # pylint: disable=protected-access
    res = {"sensitivity": list(qa.check_model_sensitivity_to_quantization(sim))}
    res["enabled"] = qa.perform_per_layer_analysis_by_enabling_quant_wrappers(sim, results_dir)
    res["disabled"] = qa.perform_per_layer_analysis_by_disabling_quant_wrappers(sim, results_dir)
    res["weights"], res["activations"] = qa.export_per_layer_encoding_min_max_range(sim, results_dir)
    res["mse"] = qa.export_per_layer_mse_loss(sim, results_dir)
    if scheme == "tf_enhanced":
        qa.export_per_layer_stats_histogram(sim, results_dir)
    res["files"] = sorted(os.path.relpath(os.path.join(d, f), results_dir) for d, _, fs in os.walk(results_dir) for f in fs)
    return json.loads(json.dumps(res)), fingerprint, sim


@pytest.fixture()
def oracle_backend(oracle):
    from aimet_b200.quantsim import tensor_quantizer
    from tests.oracle_backend import OracleTensorQuantizer
    prev = tensor_quantizer._set_op_class_for_testing(OracleTensorQuantizer)
    yield
    tensor_quantizer._set_op_class_for_testing(prev)


@pytest.mark.parametrize("name", list(CASES))
def test_matches_the_reference_quant_analyzer(oracle_backend, tmp_path, name):
    res, fingerprint, sim = analyze(name, str(tmp_path))
    gold = sim_golden(f"quant_analyzer_{name}.json", "make_quant_analyzer_golden.py", [], fingerprint)
    assert res["sensitivity"] == gold["sensitivity"]
    for table in ("enabled", "disabled", "mse"):
        assert list(res[table]) == list(gold[table]), table          # same layers, in order of occurrence
        assert res[table] == gold[table], table                      # same scores, bit for bit
    assert res["weights"] == gold["weights"] and res["activations"] == gold["activations"]
    # the reference's JSON files are all there (plus the histogram numbers it only plots)
    assert set(gold["files"]) <= set(res["files"])
    for rel in gold["files"]:
        key = {"per_layer_quant_enabled.json": "enabled", "per_layer_quant_disabled.json": "disabled",
               "per_layer_mse_loss.json": "mse", "min_max_ranges/weights.json": "weights",
               "min_max_ranges/activations.json": "activations"}[rel]
        assert json.load(open(tmp_path / rel)) == gold[key], rel
    if CASES[name][1] == "tf_enhanced":
        hist = json.load(open(tmp_path / "activations_pdf" / "relu1_output_q0_0.json"))
        assert len(hist["histogram"]) == 512 and hist["encoding"]["bw"] == 8
        assert hist["encoding"]["max"] == gold["activations"]["relu1_output_0"][1]
        assert os.path.exists(tmp_path / "weights_pdf" / "conv1" / "conv1_weight_0.json")
    if CASES[name][2]:
        assert "conv2" not in res["enabled"] and not any(n == "conv2" for n, _ in sim.quant_wrappers())


def test_analyze_runs_everything_and_restores_the_quantizers(oracle_backend, tmp_path):
    from aimet_b200.quantsim import CallbackFunc, QuantAnalyzer
    model = make_model()
    batches, target = make_data()
    fwd, ev = callbacks(batches, target)
    qa = QuantAnalyzer(model, batches[0], CallbackFunc(fwd, None), CallbackFunc(ev, None))
    qa.enable_per_layer_mse_loss(batches, 3)
    sim = qa.analyze(results_dir=str(tmp_path))
    for rel in ("per_layer_quant_enabled.json", "per_layer_quant_disabled.json", "per_layer_mse_loss.json",
                "min_max_ranges/weights.json", "min_max_ranges/activations.json"):
        assert os.path.exists(tmp_path / rel), rel
    # every sweep leaves the quantizers as the config set them
    enabled = [q.enabled for _, w in sim.quant_wrappers() for q in list(w.param_quantizers.values()) +
               list(w.input_quantizers) + list(w.output_quantizers)]
    sim2 = qa._create_quantsim_and_encodings("tf_enhanced", 8, 8, None)   # pylint: disable=protected-access
    assert enabled == [q.enabled for _, w in sim2.quant_wrappers() for q in list(w.param_quantizers.values()) +
                       list(w.input_quantizers) + list(w.output_quantizers)]
    with pytest.raises(ValueError):
        QuantAnalyzer(model, batches[0], fwd, CallbackFunc(ev, None))
    with pytest.raises(ValueError):
        qa.enable_per_layer_mse_loss(batches, 4)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["net_default_tfe", "net_perchannel_tfe", "net_default_tf"])
def test_cuda_ops_and_oracle_give_identical_tables(oracle, tmp_path, name):
    from aimet_b200 import AimetTensorQuantizer
    from aimet_b200.quantsim import tensor_quantizer
    from tests.oracle_backend import OracleTensorQuantizer
    torch.backends.cudnn.deterministic = True
    torch.backends.cudnn.benchmark = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    results = []
    for sub, factory in (("native", AimetTensorQuantizer), ("oracle", OracleTensorQuantizer)):
        prev = tensor_quantizer._set_op_class_for_testing(factory)
        try:
            results.append(analyze(name, str(tmp_path / sub), device="cuda")[0])
        finally:
            tensor_quantizer._set_op_class_for_testing(prev)
    native, oracle_res = results
    assert native == oracle_res
