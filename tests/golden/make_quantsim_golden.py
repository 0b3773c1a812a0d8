"""Generates tests/golden/quantsim_*.json by running the REFERENCE's unmodified Python (aimet_torch.v1.quantsim) on top
of the reference's unmodified C++ (oracle/_ref) in this container -- see ref_python_env.py for the import shims.

For each model: which quantizers exist and are enabled / symmetric / per-channel after QuantizationSimModel(...), and
the encodings `sim.get_activation_param_encodings()` returns after `compute_encodings` on a seeded batch (CPU forward).
"""
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.environ.get("GOLDEN_OUT", HERE)   # tests regenerate into a scratch directory on a host with other CPU numerics
sys.path.insert(0, HERE)
import ref_python_env  # noqa: E402,F401
import torch  # noqa: E402
import torchvision  # noqa: E402
from aimet_common.defs import QuantScheme  # noqa: E402
from aimet_torch.v1.qc_quantize_op import QcQuantizeWrapper  # noqa: E402
from aimet_torch.v1.quantsim import QuantizationSimModel  # noqa: E402

CFG = "/root/reference/TrainingExtensions/common/src/python/aimet_common/quantsim_config/"
CASES = {
    # name: (constructor, config file or None, quant scheme, input shape, store full encodings)
    "resnet18_default_tfe": (torchvision.models.resnet18, None, QuantScheme.post_training_tf_enhanced, (4, 3, 64, 64), True),
    "resnet18_perchannel_tfe": (torchvision.models.resnet18, CFG + "default_config_per_channel.json",
                                QuantScheme.post_training_tf_enhanced, (2, 3, 64, 64), False),
    "resnet18_default_tf": (torchvision.models.resnet18, None, QuantScheme.post_training_tf, (4, 3, 64, 64), True),
    "mobilenet_v2_default_tfe": (torchvision.models.mobilenet_v2, None, QuantScheme.post_training_tf_enhanced,
                                 (2, 3, 64, 64), True),
    "resnet50_perchannel_tfe": (torchvision.models.resnet50, CFG + "default_config_per_channel.json",
                                QuantScheme.post_training_tf_enhanced, (2, 3, 64, 64), False),
    "mobilenet_v2_perchannel_tfe": (torchvision.models.mobilenet_v2, CFG + "default_config_per_channel.json",
                                    QuantScheme.post_training_tf_enhanced, (2, 3, 64, 64), False),
    "resnet18_perchannel_tf": (torchvision.models.resnet18, CFG + "default_config_per_channel.json",
                               QuantScheme.post_training_tf, (2, 3, 64, 64), False),
    "vgg11_default_tfe": (torchvision.models.vgg11, None, QuantScheme.post_training_tf_enhanced, (2, 3, 64, 64), True),
}


def canonical(enc):
    return json.dumps(enc, sort_keys=True)


def main():
    only = sys.argv[1:]
    for name, (ctor, cfg, scheme, shape, full) in CASES.items():
        if only and name not in only:
            continue
        torch.manual_seed(0)
        model = ctor().eval()
        torch.manual_seed(1)
        x = torch.randn(*shape)
        x2 = torch.randn(*shape) * 1.5
        with torch.no_grad():   # torch's CPU convolutions differ in the last bit between hosts: record which host this is
            fingerprint = hashlib.sha256(model(x).numpy().tobytes()).hexdigest()
        sim = QuantizationSimModel(model, dummy_input=x, quant_scheme=scheme, default_output_bw=8, default_param_bw=8,
                                   config_file=cfg)
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

        def calib(m, _):
            m(x)
            m(x2)

        sim.compute_encodings(calib, None)
        act, par = sim.get_activation_param_encodings()
        with torch.no_grad():
            out = sim.model(x)
        enc = {"activation_encodings": act, "param_encodings": par}
        golden = {"forward_fingerprint": fingerprint, "structure": structure, "sha256": hashlib.sha256(canonical(enc).encode()).hexdigest(),
                  "output_sha256": hashlib.sha256(out.numpy().tobytes()).hexdigest(),
                  "num_activation": sum(len(v.get("input", {})) + len(v.get("output", {})) for v in act.values()),
                  "num_param": len(par)}
        if full:
            golden["encodings"] = json.loads(canonical(enc))
        else:
            golden["activation_encodings"] = json.loads(canonical(act))
            first = {k: par[k][:3] for k in list(par)[:4]}
            golden["param_encodings_sample"] = json.loads(canonical(first))
        with open(os.path.join(OUT, f"quantsim_{name}.json"), "w") as f:
            json.dump(golden, f, sort_keys=True, indent=1)
        print(name, golden["num_activation"], golden["num_param"], golden["sha256"][:12])


if __name__ == "__main__":
    main()
