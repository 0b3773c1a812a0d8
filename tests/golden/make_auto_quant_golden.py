"""Golden for the AutoQuant mirror, produced by the reference's UNMODIFIED aimet_torch.v1.auto_quant (its own batch-norm
folding, AdaRound and QuantizationSimModel) on the reference's own C++ (CPU).
    python tests/golden/make_auto_quant_golden.py      (build container only)

Two things are switched off in the reference run, because this repo does not build them: the ONNX export inside
_EvalSession._export (onnx is not installed; the model is pickled and the encodings saved as JSON instead) and the
cross-layer-equalization stage (made to raise, which AutoQuant tolerates under strict_validation=False: "best effort")."""
import json
import os
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_python_env  # noqa: F401,E402  (stubs + native stand-ins over oracle/_ref)

for name in ["aimet_torch.v2.nn.base", "bokeh.resources", "bokeh.plotting", "bokeh.models", "bokeh.transform", "bokeh.colors",
             "bokeh.model", "bokeh.embed", "jinja2"]:
    try:
        __import__(name)
    except Exception:   # pylint: disable=broad-except
        mod = ref_python_env._stub(name)
        parent, _, child = name.rpartition(".")
        if parent in sys.modules:
            setattr(sys.modules[parent], child, mod)
import torch  # noqa: E402
from aimet_common.defs import QuantScheme  # noqa: E402
from aimet_torch.v1 import auto_quant as ref_aq  # noqa: E402
from aimet_torch.v1.adaround.adaround_weight import AdaroundParameters  # noqa: E402

from make_auto_quant_cases import ADAROUND_ITERATIONS, CASES, make_eval_callback, make_loader, make_model  # noqa: E402


def _export(self, sim, export_kwargs):
    model_path = os.path.join(self._results_dir, f"{self.title_lowercase}.pth")
    torch.save(ref_aq.QuantizationSimModel.get_original_model(sim.model), model_path)
    sim.save_encodings_to_json(self._results_dir, self.title_lowercase)
    return model_path, os.path.join(self._results_dir, f"{self.title_lowercase}.json")


ref_aq._EvalSession._export = _export
ref_aq._EvalManager.export_diagnostics = lambda self: ""
_load = torch.load
torch.load = lambda *a, **k: _load(*a, **{**k, "weights_only": False})


def _no_cle(self, model):
    raise RuntimeError("cross-layer equalization is not part of this comparison")


ref_aq.AutoQuantBase._apply_cross_layer_equalization = _no_cle

out = {}
for name, (param_bw, output_bw, drop) in CASES.items():
    model = make_model()
    loader = make_loader()
    eval_callback = make_eval_callback(model, loader)
    with tempfile.TemporaryDirectory() as tmp:
        aq = ref_aq.AutoQuant(model, next(iter(loader)), loader, eval_callback, param_bw=param_bw, output_bw=output_bw,
                              quant_scheme=QuantScheme.post_training_tf_enhanced, results_dir=tmp, strict_validation=False,
                              model_prepare_required=False)
        aq.set_adaround_params(AdaroundParameters(loader, len(loader), default_num_iterations=ADAROUND_ITERATIONS))
        scores = {}
        orig = aq._evaluate_model_performance
        torch.manual_seed(1)
        sim, acc = aq.run_inference()
        run_inference = {"accuracy": acc, "encodings": json.loads(json.dumps(
            dict(zip(("activation_encodings", "param_encodings"), sim.get_activation_param_encodings()))))}
        torch.manual_seed(1)
        best_model, best_acc, enc_path = aq.optimize(allowed_accuracy_drop=drop)
        sessions = {t: {"status": s.result["status"], "accuracy": None if s.ptq_result is None else s.ptq_result.accuracy,
                        "techniques": None if s.ptq_result is None else s.ptq_result.applied_techniques}
                    for t, s in aq.eval_manager._all_sessions.items()}
        pair = aq._quantsim_params["quant_scheme"]
        out[name] = {"run_inference": run_inference, "accuracy": best_acc, "fp32_accuracy": aq._fp32_acc,
                     "quant_scheme": str(pair), "sessions": sessions,
                     "encoding_file": os.path.basename(enc_path) if enc_path else None,
                     "weights": None if best_model is None else
                     {n: p.detach().double().tolist() for n, p in best_model.named_parameters()}}
    print(name, out[name]["quant_scheme"], out[name]["accuracy"], {k: (v["status"], v["accuracy"]) for k, v in sessions.items()})
path = os.path.join(os.environ.get("GOLDEN_OUT", HERE), "auto_quant.json")
json.dump(out, open(path, "w"))
print(path)
