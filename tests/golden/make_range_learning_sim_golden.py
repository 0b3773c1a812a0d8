"""Generates tests/golden/range_learning_sim_*.json from the REFERENCE's unmodified Python (QuantizationSimModel with a
range-learning scheme -> LearnedGridQuantWrapper -> QuantizeDequantizeFunc) on the reference's unmodified C++.

For each case: calibrate, then one forward + backward + SGD step with the trainable encodings, recording
  * the `<name>_encoding_min/max` parameters right after calibration,
  * the exported encodings (get_activation_param_encodings) after calibration,
  * the loss, the output digest and every encoding-parameter gradient of the training step,
  * the exported encodings after the optimizer step and a second forward (which gates the updated parameters).
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
from aimet_torch.v1.quantsim import QuantizationSimModel  # noqa: E402

from make_range_learning_cases import SIM_CASES, sim_inputs, sim_model  # noqa: E402

CFG = "/root/reference/TrainingExtensions/common/src/python/aimet_common/quantsim_config/"
SCHEMES = {"tf": QuantScheme.training_range_learning_with_tf_init,
           "tf_enhanced": QuantScheme.training_range_learning_with_tf_enhanced_init}


def encoding_params(model):
    return {n: p for n, p in model.named_parameters() if n.endswith("_encoding_min") or n.endswith("_encoding_max")}


def compact(act, par):
    """All activation encodings; the first three channels of every parameter (per-channel lists are long)."""
    return json.loads(json.dumps({"activation_encodings": act, "param_encodings": {k: v[:3] for k, v in par.items()},
                                  "param_channels": {k: len(v) for k, v in par.items()}}, sort_keys=True))


def main():
    for name, (arch, cfg, scheme, shape) in SIM_CASES.items():
        model = sim_model(arch)
        x, x2, target = sim_inputs(shape)
        with torch.no_grad():   # torch's CPU convolutions differ in the last bit between hosts: record which host this is
            fingerprint = hashlib.sha256(model(x).numpy().tobytes()).hexdigest()
        sim = QuantizationSimModel(model, dummy_input=x, quant_scheme=SCHEMES[scheme], default_output_bw=8,
                                   default_param_bw=8, config_file=(CFG + cfg) if cfg else None)

        def calib(m, _):
            m(x)
            m(x2)

        sim.compute_encodings(calib, None)
        gold = {"forward_fingerprint": fingerprint, "wrapper_types": sorted({type(m).__name__ for m in sim.model.modules()
                                         if type(m).__name__.endswith("QuantWrapper")})}
        gold["initial_params"] = {n: p.detach().tolist() for n, p in encoding_params(sim.model).items()}
        act, par = sim.get_activation_param_encodings()
        gold["encodings_after_calibration"] = compact(act, par)
        sim.model.eval()
        opt = torch.optim.SGD(sim.model.parameters(), lr=1e-3)
        out = sim.model(x)
        loss = torch.nn.functional.mse_loss(out, target)
        loss.backward()
        gold["loss"] = float(loss)
        gold["output_sha256"] = hashlib.sha256(out.detach().numpy().tobytes()).hexdigest()
        gold["output_head"] = out.detach().reshape(-1)[:8].tolist()
        gold["grads"] = {n: (p.grad.tolist() if p.grad is not None else None)
                         for n, p in encoding_params(sim.model).items()}
        weight_grads = {n: p.grad for n, p in sim.model.named_parameters()
                        if p.grad is not None and not n.endswith(("_encoding_min", "_encoding_max"))}
        gold["weight_grad_sha256"] = {n: hashlib.sha256(g.numpy().tobytes()).hexdigest()
                                      for n, g in list(weight_grads.items())[:6]}
        gold["weight_grad_norms"] = {n: float(g.norm()) for n, g in weight_grads.items()}
        opt.step()
        with torch.no_grad():
            out2 = sim.model(x)
        gold["output2_head"] = out2.reshape(-1)[:8].tolist()
        act, par = sim.get_activation_param_encodings()
        gold["encodings_after_step"] = compact(act, par)
        with open(os.path.join(OUT, f"range_learning_sim_{name}.json"), "w") as f:
            json.dump(gold, f, sort_keys=True, indent=1)
        print(name, gold["wrapper_types"], "loss", gold["loss"], "params", len(gold["initial_params"]))


if __name__ == "__main__":
    main()
