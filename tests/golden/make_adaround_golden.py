"""Golden for the AdaRound mirror, produced by the reference's UNMODIFIED aimet_torch.v1.adaround on its own C++ (CPU).
    python tests/golden/make_adaround_golden.py      (build container only)"""
import json
import os
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_python_env  # noqa: F401,E402  (stubs + native stand-ins over oracle/_ref)
ref_python_env._stub("aimet_torch.v2.nn.base")       # utils.get_all_quantizers imports BaseQuantizationMixin from it
import numpy as np  # noqa: E402
import torch  # noqa: E402
from aimet_common.defs import QuantScheme  # noqa: E402
from aimet_torch.v1.adaround.adaround_weight import Adaround, AdaroundParameters  # noqa: E402

from make_adaround_cases import CASES, make_batches, make_model  # noqa: E402

out = {}
for name, (cfg, bw, iters) in CASES.items():
    torch.manual_seed(0)
    model = make_model().eval()
    batches = make_batches()
    params = AdaroundParameters(batches, num_batches=len(batches), default_num_iterations=iters)
    config_file = None
    if cfg == "per_channel":
        config_file = os.path.join(ref_python_env.REF, "common/src/python/aimet_common/quantsim_config/default_config_per_channel.json")
    with tempfile.TemporaryDirectory() as tmp:
        torch.manual_seed(1)
        rounded = Adaround.apply_adaround(model, batches[0], params, tmp, "ada", default_param_bw=bw,
                                          default_quant_scheme=QuantScheme.post_training_tf_enhanced,
                                          default_config_file=config_file)
        enc = json.load(open(os.path.join(tmp, "ada.encodings")))
    out[name] = {"encodings": enc,
                 "weights": {n: p.detach().numpy().astype(np.float64).tolist() for n, p in rounded.named_parameters()
                             if n.endswith("weight")}}
path = os.path.join(os.environ.get("GOLDEN_OUT", HERE), "adaround.json")
json.dump(out, open(path, "w"))
print(path, {k: list(v["weights"]) for k, v in out.items()})
