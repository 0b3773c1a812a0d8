"""Generates tests/golden/quant_analyzer_*.json by running the REFERENCE's unmodified QuantAnalyzer
(aimet_torch/v1/quant_analyzer.py) on the reference's unmodified QuantizationSimModel and C++ (see ref_python_env.py).

bokeh is not in this image: its modules are stubbed and the four plot writers of aimet_common.quant_analyzer are replaced
by functions that only create the output directory (which the real ones do as a side effect). Everything recorded here --
the three sensitivity scores, the two per-layer sweeps, the encoding ranges and the per-layer MSE table -- is what the
reference returns / writes with save_json.
"""
import hashlib
import json
import os
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.environ.get("GOLDEN_OUT", HERE)
sys.path.insert(0, HERE)
import ref_python_env  # noqa: E402

ref_python_env._stub("aimet_torch.v2.nn.base")   # imported (and unused for v1 wrappers) by utils.disable_all_quantizers
for _n in ("bokeh.models", "bokeh.plotting", "bokeh.layouts"):
    ref_python_env._stub(_n)
    setattr(sys.modules["bokeh"], _n.split(".")[1], sys.modules[_n])

import torch  # noqa: E402
import aimet_torch.v1.quant_analyzer as QA  # noqa: E402
from aimet_common.defs import QuantScheme  # noqa: E402
from aimet_common.utils import CallbackFunc  # noqa: E402

from make_quant_analyzer_cases import CASES, callbacks, make_data, make_model  # noqa: E402


def _only_make_directories(*args, **kwargs):
    for v in list(args) + list(kwargs.values()):
        if isinstance(v, str) and os.path.isabs(v):
            os.makedirs(v, exist_ok=True)


for _f in ("export_per_layer_sensitivity_analysis_plot", "create_and_export_min_max_ranges_plot",
           "export_per_layer_mse_plot", "export_stats_histogram_plot"):
    setattr(QA, _f, _only_make_directories)

CFG = "/root/reference/TrainingExtensions/common/src/python/aimet_common/quantsim_config/"
SCHEMES = {"tf": QuantScheme.post_training_tf, "tf_enhanced": QuantScheme.post_training_tf_enhanced}


def main():
    for name, (cfg, scheme, ignore) in CASES.items():
        model = make_model()
        batches, target = make_data()
        with torch.no_grad():
            fingerprint = hashlib.sha256(model(batches[0]).numpy().tobytes()).hexdigest()
        fwd, ev = callbacks(batches, target)
        qa = QA.QuantAnalyzer(model, batches[0], CallbackFunc(fwd, None), CallbackFunc(ev, None),
                              modules_to_ignore=[model.conv2] if ignore else None)
        qa.enable_per_layer_mse_loss(batches, 2)
        out = tempfile.mkdtemp(prefix="qa_ref_")
        sim = qa._create_quantsim_and_encodings(SCHEMES[scheme], 8, 8, (CFG + cfg) if cfg else None)   # pylint: disable=protected-access
        gold = {"forward_fingerprint": fingerprint}
        gold["sensitivity"] = list(qa.check_model_sensitivity_to_quantization(sim))
        gold["enabled"] = qa.perform_per_layer_analysis_by_enabling_quant_wrappers(sim, out)
        gold["disabled"] = qa.perform_per_layer_analysis_by_disabling_quant_wrappers(sim, out)
        weights, activations = qa.export_per_layer_encoding_min_max_range(sim, out)
        gold["weights"], gold["activations"] = weights, activations
        gold["mse"] = qa.export_per_layer_mse_loss(sim, out)
        gold["files"] = sorted(os.path.relpath(os.path.join(d, f), out) for d, _, fs in os.walk(out) for f in fs)
        gold = json.loads(json.dumps(gold))
        with open(os.path.join(OUT, f"quant_analyzer_{name}.json"), "w") as f:
            json.dump(gold, f, indent=1)   # insertion order kept: the sweeps are ordered by occurrence
        print(name, gold["sensitivity"], len(gold["enabled"]), len(gold["mse"]))


if __name__ == "__main__":
    main()
