"""Generates tests/golden/placement.json: the quantizer placement the REFERENCE's unmodified QuantizationSimModel
(ConnectedGraph from a jit trace + QuantSimConfigurator) decides for the architectures of make_placement_cases.py.
    python tests/golden/make_placement_golden.py [case ...]      (build container only)"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.environ.get("GOLDEN_OUT", HERE)
sys.path.insert(0, HERE)
import ref_python_env  # noqa: E402,F401
import torch  # noqa: E402
from aimet_common.defs import QuantScheme  # noqa: E402
from aimet_torch.v1.qc_quantize_op import QcQuantizeWrapper  # noqa: E402
from aimet_torch.v1.quantsim import QuantizationSimModel  # noqa: E402

from make_placement_cases import CASES  # noqa: E402

CFG = "/root/reference/TrainingExtensions/common/src/python/aimet_common/quantsim_config/"


def main():
    only = sys.argv[1:]
    path = os.path.join(OUT, "placement.json")
    out = json.load(open(path)) if os.path.exists(path) and only else {}
    for name, (ctor, cfg, shape) in CASES.items():
        if only and name not in only:
            continue
        torch.manual_seed(0)
        model = ctor().eval()
        x = torch.randn(*shape)
        try:
            sim = QuantizationSimModel(model, dummy_input=x, quant_scheme=QuantScheme.post_training_tf_enhanced,
                                       default_output_bw=8, default_param_bw=8,
                                       config_file=CFG + "default_config_per_channel.json" if cfg == "per_channel" else None)
        except Exception as exc:   # pylint: disable=broad-except
            print(name, "REFERENCE FAILED:", repr(exc)[:200])
            continue
        structure = {}
        for mname, w in sim.model.named_modules():
            if isinstance(w, QcQuantizeWrapper):
                structure[mname] = {
                    "type": type(w._module_to_wrap).__name__,
                    "inputs": [bool(q.enabled) for q in w.input_quantizers],
                    "outputs": [bool(q.enabled) for q in w.output_quantizers],
                    "params": {k: [bool(q.enabled), bool(q.use_symmetric_encodings), type(q).__name__]
                               for k, q in w.param_quantizers.items()},
                }
        out[name] = structure
        n_act = sum(sum(s["inputs"]) + sum(s["outputs"]) for s in structure.values())
        n_par = sum(sum(1 for p in s["params"].values() if p[0]) for s in structure.values())
        print(name, len(structure), "wrappers,", n_act, "activation /", n_par, "parameter quantizers enabled")
    json.dump(out, open(path, "w"), sort_keys=True)
    print(path)


if __name__ == "__main__":
    main()
