"""The reference's OWN, unmodified Python -- aimet_torch.v1.quantsim.QuantizationSimModel with its ConnectedGraph, wrappers
and per-channel loops, staged under baseline/_ref by tools/make_ref_python.py -- running on a B200 on top of aimet_b200's
drop-ins for the two native modules it imports (aimet_b200.install). Checked against the same reference Python over the CPU
oracle on the same device tensors: encodings JSON and quantized forward identical. (missing item 1.)
Each case runs in its own process: tests/ref_python_driver.py puts stub modules into sys.modules.
"""
import hashlib
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref")


def drive(*args, timeout=900):
    if not os.path.isdir(os.path.join(REF, "aimet_torch")):
        pytest.skip("baseline/_ref not staged (tools/make_ref_python.py needs the reference checkout)")
    res = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "ref_python_driver.py"), *args], cwd=ROOT,
                         stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=timeout)
    assert res.returncode == 0, res.stderr[-3000:]
    return json.loads(res.stdout.strip().splitlines()[-1])


def test_staged_reference_python_is_unmodified():
    if not os.path.exists(os.path.join(REF, "MANIFEST.json")):
        pytest.skip("baseline/_ref not staged")
    manifest = json.load(open(os.path.join(REF, "MANIFEST.json")))["files"]
    assert len(manifest) > 150
    for rel, sha in manifest.items():
        assert hashlib.sha256(open(os.path.join(REF, rel), "rb").read()).hexdigest() == sha, rel


@pytest.mark.parametrize("config,scheme", [("per_channel", "tf_enhanced"), ("default", "tf_enhanced"), ("default", "tf")])
def test_reference_quantsim_on_the_cuda_dropins_equals_reference_on_the_oracle(config, scheme):
    r = drive("--model", "resnet18", "--config", config, "--scheme", scheme, "--batch", "4", "--image", "64", "--steps", "2",
              "--backend", "both")
    assert r["quantsim_module"].startswith("baseline/_ref/aimet_torch")          # the reference's file, not the mirror
    assert r["native"]["wrappers"] == 52 and r["native"]["num_activation_encodings"] == 24
    assert r["native"]["num_param_encodings"] == (4801 if config == "per_channel" else 21)
    assert r["native"]["aimet_b200_launches"] > 100 and r["oracle"]["aimet_b200_launches"] == 0
    assert r["native"]["encodings_sha256"] == r["oracle"]["encodings_sha256"]
    assert r["native"]["output_sha256"] == r["oracle"]["output_sha256"]
    assert r["equal"]


def test_reference_quantsim_resnet18_on_the_baseline_shape():
    """BASELINE configs[0] through the reference's Python: one batch 32 x 3 x 224 x 224, default config."""
    r = drive("--model", "resnet18", "--config", "default", "--batch", "32", "--image", "224", "--steps", "1", "--backend", "both")
    assert r["equal"]
