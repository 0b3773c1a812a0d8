"""Run the REFERENCE's own, unmodified Python (aimet_torch.v1.quantsim.QuantizationSimModel: its ConnectedGraph, its
wrappers, its per-channel Python loops) on a CUDA device on top of aimet_b200's drop-ins for the two native modules it
imports (aimet_common.AimetTensorQuantizer, aimet_common._libpymo -- aimet_common/aimet_tensor_quantizer.py:42-64,
aimet_common/libpymo.py:42-47), and print one JSON line.

    python tests/ref_python_driver.py --model resnet18 --config per_channel --batch 4 --image 64 --steps 2 --backend both

--backend native : the sm_100a drop-ins (aimet_b200.install)                      -> encodings hash, timing
--backend oracle : the same reference Python over the CPU oracle (test checker)   -> encodings hash
--backend both   : both, on the same device tensors, plus `equal`

The reference's Python is staged under baseline/_ref by tools/make_ref_python.py (git-ignored, travels with the snapshot).
Modules the image lacks are stubbed exactly as SURVEY.md section 8c lists them. Runs in its own process on purpose: the stubs go
into sys.modules.
"""
import argparse
import hashlib
import importlib.machinery
import json
import os
import sys
import time
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref")

STUBS = ["bokeh", "bokeh.server", "bokeh.server.server", "bokeh.application", "spconv", "spconv.pytorch", "onnx",
         "onnxsim", "torch.onnx.symbolic_caffe2", "aimet_torch.v1.nn", "aimet_torch.v1.nn.modules",
         "aimet_torch.v1.nn.modules.custom", "aimet_torch.v2.experimental", "aimet_torch.v2.nn",
         "aimet_torch.v2.nn.fake_quant", "aimet_torch.v2.quantization", "aimet_torch.v2.quantsim",
         "aimet_torch.v2.visualization_tools"]


def setup():
    """sys.path, stubs, drop-ins. Returns the reference's aimet_torch.v1.quantsim module."""
    import torch
    import torchvision  # noqa: F401  (before the stubs go in)
    if not os.path.isdir(os.path.join(REF, "aimet_torch")):
        raise SystemExit("baseline/_ref is empty: run tools/make_ref_python.py in the build container")
    sys.path.insert(0, REF)
    sys.path.insert(0, ROOT)

    class _Dummy(torch.nn.Module):
        pass

    def stub(name):
        m = types.ModuleType(name)
        m.__spec__ = importlib.machinery.ModuleSpec(name, None)
        m.__path__ = []
        cache = {}

        def _ga(attr):
            if attr.startswith("__"):
                raise AttributeError(attr)
            if attr not in cache:
                cache[attr] = type(attr, (_Dummy,), {})
            return cache[attr]
        m.__getattr__ = _ga
        sys.modules[name] = m

    for n in STUBS:
        stub(n)
    for full in list(sys.modules):
        if "." in full:
            parent, child = full.rsplit(".", 1)
            if parent in sys.modules and getattr(sys.modules[full], "__getattr__", None) is not None and \
                    child != "symbolic_caffe2":
                try:
                    setattr(sys.modules[parent], child, sys.modules[full])
                except Exception:   # pylint: disable=broad-except
                    pass
    import aimet_common.py_libpymo as py_pymo       # the reference's own pure-python stand-ins for out-of-scope bindings
    import aimet_b200.install as inst
    extra = {k: getattr(py_pymo, k) for k in dir(py_pymo) if not k.startswith("_")}
    inst.install(extra_pymo_names=extra)
    from aimet_torch.v1 import quantsim
    return quantsim


def use_backend(name):
    """Swap the class the reference's tensor quantizers instantiate. They bind it by name at import time
    (aimet_torch/v1/tensor_quantizer.py:47 `from aimet_common.aimet_tensor_quantizer import AimetTensorQuantizer`) and look
    that module global up at construction time, so the name is re-pointed in every loaded reference module that has it."""
    if name == "native":
        cls = sys.modules["aimet_common.AimetTensorQuantizer"].AimetTensorQuantizer     # what aimet_b200.install registered
    else:
        from oracle import cpu_backend          # TEST CHECKER ONLY
        cls = cpu_backend.best_cpu_backend()
    for mod_name, mod in list(sys.modules.items()):
        if mod is not None and mod_name.startswith(("aimet_common", "aimet_torch")) and \
                isinstance(mod.__dict__.get("AimetTensorQuantizer"), type):
            mod.AimetTensorQuantizer = cls


def run(quantsim, args, backend):
    import torch
    import torchvision
    from aimet_common.defs import QuantScheme
    use_backend(backend)
    torch.backends.cudnn.deterministic = True
    torch.backends.cudnn.benchmark = False
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    device = torch.device(args.device)
    torch.manual_seed(0)
    model = getattr(torchvision.models, args.model)().eval().to(device)
    batches = [torch.randn(args.batch, 3, args.image, args.image, generator=torch.Generator().manual_seed(1000 + b)).to(device)
               for b in range(args.steps)]
    cfg = None
    if args.config == "per_channel":
        cfg = os.path.join(REF, "aimet_common", "quantsim_config", "default_config_per_channel.json")
    scheme = {"tf_enhanced": QuantScheme.post_training_tf_enhanced, "tf": QuantScheme.post_training_tf}[args.scheme]
    t0 = time.perf_counter()
    sim = quantsim.QuantizationSimModel(model, dummy_input=batches[0][:1], quant_scheme=scheme, default_output_bw=8,
                                        default_param_bw=8, config_file=cfg)
    sync = torch.cuda.synchronize if device.type == "cuda" else (lambda: None)
    sync()
    t_build = time.perf_counter() - t0

    def calibrate(m, _):
        with torch.no_grad():
            for x in batches:
                m(x)

    def job():
        sim.compute_encodings(calibrate, None)
        return sim.get_activation_param_encodings()

    if args.warmup:
        job()
    sync()
    t0 = time.perf_counter()
    act, par = job()
    sync()
    dt = time.perf_counter() - t0
    doc = json.dumps({"activation_encodings": act, "param_encodings": par}, sort_keys=True)
    with torch.no_grad():
        out = sim.model(batches[0])
    return {"backend": backend, "seconds": round(dt, 4), "img_s": round(args.batch * args.steps / dt, 2),
            "build_seconds": round(t_build, 3), "num_activation_encodings": len(act), "num_param_tensors": len(par),
            "num_param_encodings": sum(len(v) for v in par.values()),
            "wrappers": sum(1 for m in sim.model.modules() if type(m).__name__ == "StaticGridQuantWrapper"),
            "encodings_sha256": hashlib.sha256(doc.encode()).hexdigest(),
            "output_sha256": hashlib.sha256(out.detach().cpu().numpy().tobytes()).hexdigest()}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="resnet18")
    ap.add_argument("--config", default="per_channel", choices=["default", "per_channel"])
    ap.add_argument("--scheme", default="tf_enhanced", choices=["tf", "tf_enhanced"])
    ap.add_argument("--batch", type=int, default=4)
    ap.add_argument("--image", type=int, default=64)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=0)
    ap.add_argument("--backend", default="native", choices=["native", "oracle", "both"])
    ap.add_argument("--device", default="cuda:0", help="cpu only makes sense with --backend oracle (a wiring check)")
    args = ap.parse_args()
    real_stdout = os.dup(1)
    os.dup2(2, 1)                     # the reference logs to stdout
    quantsim = setup()
    import aimet_b200
    from aimet_b200 import ops
    result = {"reference_python": os.path.relpath(REF, ROOT), "model": args.model, "config": args.config,
              "scheme": args.scheme, "batch": args.batch, "image": args.image, "steps": args.steps,
              "quantsim_module": quantsim.__file__.replace(ROOT + os.sep, "")}
    backends = ["native", "oracle"] if args.backend == "both" else [args.backend]
    for b in backends:
        before = ops.launches_total()
        result[b] = run(quantsim, args, b)
        result[b]["aimet_b200_launches"] = ops.launches_total() - before
    if args.backend == "both":
        result["equal"] = result["native"]["encodings_sha256"] == result["oracle"]["encodings_sha256"] and \
            result["native"]["output_sha256"] == result["oracle"]["output_sha256"]
    del aimet_b200
    os.write(real_stdout, (json.dumps(result) + "\n").encode())


if __name__ == "__main__":
    main()
