"""The AdaRound mirror (aimet_b200.quantsim.adaround) against the reference's unmodified aimet_torch.v1.adaround:
tests/golden/adaround.json holds the adarounded weights and the exported parameter encodings of a seeded run of the
reference's own code on its own C++ (make_adaround_golden.py). CPU: the mirror over the oracle backend reproduces them.
GPU: the mirror over the CUDA ops equals the mirror over the oracle on the same device."""
import json
import os

import pytest
import torch

from tests.conftest import GOLDEN
from tests.golden.make_adaround_cases import CASES, make_batches, make_model


def run_mirror(name, device="cpu"):
    from aimet_b200.quantsim import QuantScheme
    from aimet_b200.quantsim import config as qconfig
    from aimet_b200.quantsim.adaround import Adaround, AdaroundParameters
    cfg, bw, iters = CASES[name]
    torch.manual_seed(0)
    model = make_model().eval().to(device)
    batches = [b.to(device) for b in make_batches()]
    params = AdaroundParameters(batches, num_batches=len(batches), default_num_iterations=iters)
    import tempfile
    with tempfile.TemporaryDirectory() as tmp:
        torch.manual_seed(1)
        rounded = Adaround.apply_adaround(model, batches[0], params, tmp, "ada", default_param_bw=bw,
                                          default_quant_scheme=QuantScheme.post_training_tf_enhanced,
                                          default_config_file=qconfig.DEFAULT_CONFIG_PER_CHANNEL if cfg == "per_channel" else None)
        enc = json.load(open(os.path.join(tmp, "ada.encodings")))
    weights = {n: p.detach().cpu() for n, p in rounded.named_parameters() if n.endswith("weight")}
    return enc, weights, model


@pytest.fixture()
def oracle_backend(oracle):
    from aimet_b200.quantsim import tensor_quantizer
    from tests.oracle_backend import OracleTensorQuantizer
    prev = tensor_quantizer._set_op_class_for_testing(OracleTensorQuantizer)
    yield
    tensor_quantizer._set_op_class_for_testing(prev)


@pytest.mark.parametrize("name", list(CASES))
def test_adaround_mirror_reproduces_reference(oracle_backend, name):
    gold = json.load(open(os.path.join(GOLDEN, "adaround.json")))[name]
    enc, weights, model = run_mirror(name)
    assert enc == gold["encodings"]                       # parameter encodings: exact
    assert set(weights) == set(gold["weights"])
    original = {n: p.detach() for n, p in model.named_parameters() if n.endswith("weight")}
    moved = 0
    for n, w in weights.items():
        g = torch.tensor(gold["weights"][n], dtype=torch.float64)
        # every weight sits on the same grid point the reference chose (the optimisation is torch arithmetic whose
        # reductions may differ in the last bit between hosts; a flipped rounding decision would be a whole grid step)
        assert torch.allclose(w.double(), g, rtol=1e-6, atol=1e-7), n
        moved += int((w != original[n]).sum())
    assert moved > 0
    # batch norm is not an AdaRound module: untouched
    assert torch.equal(weights["3.weight"], original["3.weight"])


def test_adaround_pieces():
    from aimet_b200.quantsim.adaround import (AdaroundHyperParameters, AdaroundLoss, AdaroundParameters,
                                              get_module_act_func_pair, get_ordered_list_of_modules)
    model = make_model()
    pairs = get_module_act_func_pair(model)
    assert pairs[model[0]] is model[1] and pairs[model[2]] is None and pairs[model[7]] is None    # conv->relu, conv->bn
    names = [n for n, _ in get_ordered_list_of_modules(model, make_batches()[0])]
    assert names == [str(i) for i in range(8)]
    with pytest.raises(ValueError):
        AdaroundParameters(make_batches(), num_batches=9)
    hp = AdaroundHyperParameters(100, 0.01, (20, 2), 0.2)
    assert AdaroundLoss.compute_round_loss(torch.zeros(4), hp, 5) == 0                           # warm start
    assert AdaroundLoss._compute_beta(100, 20, (20, 2), 0.2) == pytest.approx(20.0)
    assert AdaroundLoss._compute_beta(100, 99, (20, 2), 0.2) == pytest.approx(2.0, abs=0.01)


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(CASES))
def test_adaround_on_cuda_ops_equals_oracle_backend(oracle, name):
    from aimet_b200.quantsim import tensor_quantizer
    from tests.oracle_backend import OracleTensorQuantizer
    torch.backends.cudnn.deterministic = True
    torch.backends.cudnn.benchmark = False
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    enc_n, w_n, _ = run_mirror(name, "cuda")
    prev = tensor_quantizer._set_op_class_for_testing(OracleTensorQuantizer)
    try:
        enc_o, w_o, _ = run_mirror(name, "cuda")
    finally:
        tensor_quantizer._set_op_class_for_testing(prev)
    assert enc_n == enc_o
    for n in w_n:
        assert torch.equal(w_n[n], w_o[n]), n
