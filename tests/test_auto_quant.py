"""The AutoQuant mirror (aimet_b200.quantsim.auto_quant) against the reference's unmodified aimet_torch.v1.auto_quant:
tests/golden/auto_quant.json holds the decisions (chosen quant scheme, status and score of every stage, applied techniques),
the scores and the resulting weights / encodings of seeded runs of the reference's own code on its own C++
(make_auto_quant_golden.py). CPU: the mirror over the oracle backend reproduces them. GPU: the mirror over the CUDA ops takes
the same decisions and reaches the same scores as over the oracle on the same device."""
import json
import math
import os
import tempfile

import pytest
import torch

from tests.conftest import GOLDEN
from tests.golden.make_auto_quant_cases import ADAROUND_ITERATIONS, CASES, make_eval_callback, make_loader, make_model


def run_mirror(name, device="cpu", cle="raise"):
    from aimet_b200.quantsim import QuantScheme
    from aimet_b200.quantsim.adaround import AdaroundParameters
    from aimet_b200.quantsim.auto_quant import AutoQuant
    param_bw, output_bw, drop = CASES[name]
    model = make_model().to(device)
    loader = make_loader()
    eval_callback = make_eval_callback(model.cpu(), loader) if device == "cpu" else None
    if device != "cpu":
        cpu_model = make_model()
        eval_callback = make_eval_callback(cpu_model, loader)
        model = model.to(device)
    with tempfile.TemporaryDirectory() as tmp:
        aq = AutoQuant(model, next(iter(loader)).to(device), loader, eval_callback, param_bw=param_bw, output_bw=output_bw,
                       quant_scheme=QuantScheme.post_training_tf_enhanced, results_dir=tmp, strict_validation=False)
        aq.set_adaround_params(AdaroundParameters(loader, len(loader), default_num_iterations=ADAROUND_ITERATIONS))
        if cle == "raise":      # what the golden run did to the reference's CLE stage

            def no_cle(_model):
                raise RuntimeError("cross-layer equalization is not part of this comparison")
            aq.set_cross_layer_equalization_fn(no_cle)
        torch.manual_seed(1)
        sim, acc = aq.run_inference()
        act, par = sim.get_activation_param_encodings()
        torch.manual_seed(1)
        best_model, best_acc, enc_path = aq.optimize(allowed_accuracy_drop=drop)
        sessions = {s.title: {"status": s.result["status"], "accuracy": None if s.ptq_result is None else s.ptq_result.accuracy,
                              "techniques": None if s.ptq_result is None else s.ptq_result.applied_techniques}
                    for s in aq.eval_manager._all_sessions.values()}   # pylint: disable=protected-access
        if enc_path is not None:
            assert os.path.exists(enc_path)
        return {"run_inference": {"accuracy": acc, "encodings": {"activation_encodings": act, "param_encodings": par}},
                "accuracy": best_acc, "fp32_accuracy": aq._fp32_acc,   # pylint: disable=protected-access
                "quant_scheme": str(aq._quantsim_params["quant_scheme"]), "sessions": sessions,   # pylint: disable=protected-access
                "weights": None if best_model is None else {n: p.detach().cpu() for n, p in best_model.named_parameters()},
                "summary": aq.eval_manager.summary()}


def same_score(a, b, rel=1e-4):
    if a is None or b is None:
        return a is None and b is None
    return math.isclose(a, b, rel_tol=rel, abs_tol=1e-12)


def check_against(res, gold, weights_rtol=1e-6):
    assert res["quant_scheme"] == gold["quant_scheme"]
    assert same_score(res["fp32_accuracy"], gold["fp32_accuracy"])
    assert set(res["sessions"]) == set(gold["sessions"])
    for title, g in gold["sessions"].items():
        r = res["sessions"][title]
        assert r["status"] == g["status"], title
        assert r["techniques"] == g["techniques"], title
        assert same_score(r["accuracy"], g["accuracy"]), (title, r["accuracy"], g["accuracy"])
    assert same_score(res["accuracy"], gold["accuracy"])
    assert same_score(res["run_inference"]["accuracy"], gold["run_inference"]["accuracy"])
    if gold["weights"] is None:
        assert res["weights"] is None
    else:
        assert set(res["weights"]) == set(gold["weights"])
        for n, w in res["weights"].items():
            assert torch.allclose(w.double(), torch.tensor(gold["weights"][n], dtype=torch.float64), rtol=weights_rtol,
                                  atol=1e-7), n


@pytest.fixture()
def oracle_backend(oracle):
    from aimet_b200.quantsim import tensor_quantizer
    from tests.oracle_backend import OracleTensorQuantizer
    prev = tensor_quantizer._set_op_class_for_testing(OracleTensorQuantizer)
    yield
    tensor_quantizer._set_op_class_for_testing(prev)


@pytest.mark.parametrize("name", list(CASES))
def test_auto_quant_mirror_reproduces_reference(oracle_backend, name):
    gold = json.load(open(os.path.join(GOLDEN, "auto_quant.json")))[name]
    res = run_mirror(name)
    check_against(res, gold)
    # the calibrated sim of run_inference(): encodings of the batch-norm-folded model, exact
    assert json.loads(json.dumps(res["run_inference"]["encodings"])) == gold["run_inference"]["encodings"]


def test_batch_norm_folding_preserves_the_function_and_removes_the_norms():
    from aimet_b200.quantsim.batch_norm_fold import fold_all_batch_norms
    model = make_model()
    x = next(iter(make_loader()))
    want = model(x)
    import copy
    folded = copy.deepcopy(model)
    pairs = fold_all_batch_norms(folded, None, x)
    assert len(pairs) == 2 and isinstance(folded.bn1, torch.nn.Identity) and isinstance(folded.bn2, torch.nn.Identity)
    assert folded.conv2.bias is not None            # created by the fold
    assert torch.allclose(folded(x), want, rtol=1e-4, atol=1e-5)
    # a BatchNorm in front of a padded convolution must stay (the border would see the un-normalised zero)
    seq = torch.nn.Sequential(torch.nn.BatchNorm2d(3), torch.nn.Conv2d(3, 4, 3, padding=1)).eval()
    assert fold_all_batch_norms(seq) == [] and isinstance(seq[0], torch.nn.BatchNorm2d)
    # ... and folds forward into an un-padded one and into a Linear
    for net, inp in ((torch.nn.Sequential(torch.nn.BatchNorm2d(3), torch.nn.Conv2d(3, 4, 3)), torch.randn(2, 3, 8, 8)),
                     (torch.nn.Sequential(torch.nn.BatchNorm1d(6), torch.nn.Linear(6, 4)), torch.randn(5, 6))):
        net.eval()
        with torch.no_grad():
            net[0].running_mean.normal_()
            net[0].running_var.uniform_(0.5, 2.0)
            net[0].weight.uniform_(0.5, 2.0)
            net[0].bias.normal_()
            want = net(inp)
            assert len(fold_all_batch_norms(net)) == 1 and isinstance(net[0], torch.nn.Identity)
            assert torch.allclose(net(inp), want, rtol=1e-4, atol=1e-5)


def test_without_an_equalization_function_the_stage_is_skipped(oracle_backend):
    res = run_mirror("w4a8_tight", cle=None)
    assert res["sessions"]["Cross-Layer Equalization"]["status"] == "discarded"
    assert res["sessions"]["AdaRound"]["status"] == "success"
    assert res["summary"]["batchnorm_folding"]["applied_techniques"] == ["batchnorm_folding"]


def test_input_validation():
    from aimet_b200.quantsim.auto_quant import AutoQuant
    model, loader = make_model(), make_loader()
    cb = make_eval_callback(model, loader)
    x = next(iter(loader))
    with pytest.raises(ValueError):
        AutoQuant(model, x, loader, cb, param_bw=0)
    with pytest.raises(ValueError):
        AutoQuant(model, x, loader, "not callable")
    with pytest.raises(ValueError):
        AutoQuant(model, x, loader, cb, rounding_mode="up")
    aq = AutoQuant(model, x, loader, cb)
    with pytest.raises(ValueError):
        aq.optimize(allowed_accuracy_drop=-1.0)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["w8a8_loose", "w4a8_tight"])
def test_auto_quant_on_cuda_ops_equals_oracle_backend(oracle, name):
    from aimet_b200.quantsim import tensor_quantizer
    from tests.oracle_backend import OracleTensorQuantizer
    res = run_mirror(name, device="cuda")
    prev = tensor_quantizer._set_op_class_for_testing(OracleTensorQuantizer)
    try:
        want = run_mirror(name, device="cuda")
    finally:
        tensor_quantizer._set_op_class_for_testing(prev)
    assert res["quant_scheme"] == want["quant_scheme"]
    for title, g in want["sessions"].items():
        r = res["sessions"][title]
        assert r["status"] == g["status"] and r["techniques"] == g["techniques"], title
        assert same_score(r["accuracy"], g["accuracy"], rel=1e-3), (title, r["accuracy"], g["accuracy"])
    assert res["run_inference"]["encodings"] == want["run_inference"]["encodings"]     # same device tensors: exact
