"""Generate tests/golden/percentile.npz from the reference's own PercentileEncodingAnalyzer (oracle/_ref) and
tests/golden/quantsim_resnet18_percentile.json from the reference's unmodified Python on top of it.

    python tests/golden/make_percentile_golden.py
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.environ.get("GOLDEN_OUT", HERE)   # tests regenerate the sim golden into a scratch directory (see conftest.py)
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from make_percentile_cases import ANALYZER_CASES, analyzer_batches  # noqa: E402
from oracle import bindings  # noqa: E402


def analyzer_goldens():
    ref = bindings.Reference()
    out = {}
    for name, spec in ANALYZER_CASES.items():
        batches = analyzer_batches(name)
        rows = []
        for pct in spec["percentiles"]:
            for (bw, sym, strict, unsigned) in spec["variants"]:
                a = bindings.RefAnalyzer(ref, 3)
                a.set_percentile(pct)
                for b in batches:
                    a.update(b)
                rows.append(list(a.compute(bw, sym, strict, unsigned)))
        out[name] = np.array(rows, dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, "percentile.npz"), **out)
    print("percentile.npz:", {k: v.shape for k, v in out.items()})
    # the MSE analyzer (QuantizationMode 4) on the same inputs
    out = {}
    for name, spec in ANALYZER_CASES.items():
        batches = analyzer_batches(name)
        rows = []
        for (bw, sym, strict, unsigned) in spec["variants"]:
            a = bindings.RefAnalyzer(ref, 4)
            for b in batches:
                a.update(b)
            rows.append(list(a.compute(bw, sym, strict, unsigned)))
        out[name] = np.array(rows, dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, "mse.npz"), **out)
    print("mse.npz:", {k: v.shape for k, v in out.items()})


def quantsim_golden():
    import ref_python_env  # noqa: F401
    import torch
    import torchvision
    from aimet_common.defs import QuantScheme
    from aimet_torch.v1.quantsim import QuantizationSimModel
    torch.manual_seed(0)
    model = torchvision.models.resnet18().eval()
    torch.manual_seed(1)
    x = torch.randn(4, 3, 64, 64)
    x2 = torch.randn(4, 3, 64, 64) * 1.5
    with torch.no_grad():   # torch's CPU convolutions differ in the last bit between hosts: record which host this is
        fingerprint = hashlib.sha256(model(x).numpy().tobytes()).hexdigest()
    sim = QuantizationSimModel(model, dummy_input=x, quant_scheme=QuantScheme.post_training_percentile,
                               default_output_bw=8, default_param_bw=8)
    sim.set_percentile_value(99.9)

    def calib(m, _):
        m(x)
        m(x2)

    sim.compute_encodings(calib, None)
    act, par = sim.get_activation_param_encodings()
    with torch.no_grad():
        out = sim.model(x)
    enc = json.loads(json.dumps({"activation_encodings": act, "param_encodings": par}, sort_keys=True))
    gold = {"forward_fingerprint": fingerprint, "encodings": enc, "sha256": hashlib.sha256(json.dumps(enc, sort_keys=True).encode()).hexdigest(),
            "output_sha256": hashlib.sha256(out.numpy().tobytes()).hexdigest(), "percentile": 99.9}
    with open(os.path.join(OUT, "quantsim_resnet18_percentile.json"), "w") as f:
        json.dump(gold, f, sort_keys=True, indent=1)
    print("quantsim_resnet18_percentile:", gold["sha256"][:12], len(act), len(par))


if __name__ == "__main__":
    if "--sim-only" not in sys.argv:
        analyzer_goldens()
    quantsim_golden()
