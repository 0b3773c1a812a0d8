"""Stage the reference's own, unmodified Python layers (aimet_common + aimet_torch) under baseline/_ref/ so that they travel to
the GPU box with the repo snapshot (baseline/_ref is git-ignored: reference sources never enter this repository's history).

    python tools/make_ref_python.py          # needs the reference checkout at /root/reference (build container only)

What it is for: tests/test_gpu_reference_python.py and bench.py's `reference_python_api` leg run the reference's
QuantizationSimModel -- its ConnectedGraph, its wrappers, its per-channel Python loops -- on top of aimet_b200's drop-ins for
the two native modules (aimet_b200.install), on a B200. The files are copied byte for byte; MANIFEST.json records their
sha256 so a test can tell that nothing was edited.
"""
import hashlib
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/TrainingExtensions"
TREES = {"aimet_common": REF + "/common/src/python/aimet_common", "aimet_torch": REF + "/torch/src/python/aimet_torch"}
OUT = os.path.join(ROOT, "baseline", "_ref")


def main():
    if not os.path.isdir(REF):
        print("no reference checkout here: nothing staged", file=sys.stderr)
        return 1
    manifest = {}
    for name, src in TREES.items():
        dst = os.path.join(OUT, name)
        if os.path.isdir(dst):
            shutil.rmtree(dst)
        shutil.copytree(src, dst, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
        for base, _, files in os.walk(dst):
            for f in sorted(files):
                p = os.path.join(base, f)
                manifest[os.path.relpath(p, OUT)] = hashlib.sha256(open(p, "rb").read()).hexdigest()
    with open(os.path.join(OUT, "MANIFEST.json"), "w") as f:
        json.dump({"source": REF, "files": manifest}, f, indent=1, sort_keys=True)
    print(f"staged {len(manifest)} files under {OUT}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
