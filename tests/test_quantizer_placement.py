"""Quantizer placement of the mirror (torch.fx op graph, aimet_b200/quantsim/config.py) against the reference's unmodified
QuantizationSimModel (ConnectedGraph from a jit trace + QuantSimConfigurator) on 15 architectures beyond the ResNet /
MobileNet-v2 cases of test_quantsim_host.py: plain and batch-normed CNNs, concatenating nets (SqueezeNet, DenseNet,
GoogLeNet), grouped / depthwise convolutions, squeeze-excite blocks with Hardsigmoid / SiLU / Hardswish, a transformer
block (LayerNorm, GELU, softmax, matmul), default and per-channel configs. tests/golden/placement.json comes from
make_placement_golden.py. Placement decides which tensors reach the native statistics / QDQ entry points at all, so a
difference here is a parity bug however exact the kernels are."""
import json
import os

import pytest
import torch

from tests.conftest import GOLDEN
from tests.golden.make_placement_cases import CASES


def mirror_structure(name):
    from aimet_b200.quantsim import QuantizationSimModel
    from aimet_b200.quantsim import config as qconfig
    ctor, cfg, shape = CASES[name]
    torch.manual_seed(0)
    model = ctor().eval()
    x = torch.randn(*shape)
    sim = QuantizationSimModel(model, dummy_input=x, quant_scheme="tf_enhanced", default_output_bw=8, default_param_bw=8,
                               config_file=qconfig.DEFAULT_CONFIG_PER_CHANNEL if cfg == "per_channel" else None)
    structure = {}
    for mname, w in sim.quant_wrappers():
        structure[mname] = {
            "type": type(w.get_original_module()).__name__,
            "inputs": [bool(q.enabled) for q in w.input_quantizers],
            "outputs": [bool(q.enabled) for q in w.output_quantizers],
            "params": {k: [bool(q.enabled), bool(q.use_symmetric_encodings), type(q).__name__]
                       for k, q in w.param_quantizers.items()},
        }
    return structure


@pytest.fixture()
def oracle_backend(oracle):
    from aimet_b200.quantsim import tensor_quantizer
    from tests.oracle_backend import OracleTensorQuantizer
    prev = tensor_quantizer._set_op_class_for_testing(OracleTensorQuantizer)
    yield
    tensor_quantizer._set_op_class_for_testing(prev)


@pytest.mark.parametrize("name", list(CASES))
def test_placement_matches_reference(oracle_backend, name):
    gold = json.load(open(os.path.join(GOLDEN, "placement.json")))[name]
    mine = mirror_structure(name)
    assert set(mine) == set(gold), (sorted(set(gold) - set(mine)), sorted(set(mine) - set(gold)))
    diff = {m: (mine[m], gold[m]) for m in gold if mine[m] != gold[m]}
    assert not diff, f"{len(diff)} of {len(gold)} wrappers differ, e.g. {list(diff.items())[:4]}"
