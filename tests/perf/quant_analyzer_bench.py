"""QuantAnalyzer per-layer sweeps on the GPU path, with the same host layer driven by the
reference's own C++ on the host cores beside it (bounded sample; lives under tests/ because it drives the oracle). ResNet-18, W8A8 tf_enhanced, default config; batch norms
are left unfolded in both arms (folding is outside this package), so both analyse the same 52-wrapper model.

    python tests/perf/quant_analyzer_bench.py [--batch 32] [--eval-batches 2] [--cpu-batch 4] [--out file.json]
"""
import argparse
import json
import os
import sys
import tempfile
import time

import torch
import torchvision

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def run(device, batch, eval_batches, factory=None):
    from aimet_b200.quantsim import CallbackFunc, QuantAnalyzer, tensor_quantizer
    prev = tensor_quantizer._set_op_class_for_testing(factory) if factory is not None else None
    try:
        torch.manual_seed(0)
        model = torchvision.models.resnet18().eval().to(device)
        data = [torch.randn(batch, 3, 224, 224, device=device) for _ in range(eval_batches)]

        def fwd(m, _):
            with torch.no_grad():
                for x in data:
                    m(x)

        def ev(m, _):
            with torch.no_grad():
                return float(sum(m(x).float().square().mean() for x in data))

        qa = QuantAnalyzer(model, data[0][:1], CallbackFunc(fwd, None), CallbackFunc(ev, None))
        sync = torch.cuda.synchronize if device == "cuda" else (lambda: None)
        out = tempfile.mkdtemp(prefix="qa_bench_")
        sync()
        t0 = time.perf_counter()
        sim = qa._create_quantsim_and_encodings("tf_enhanced", 8, 8, None)   # pylint: disable=protected-access
        sync()
        t1 = time.perf_counter()
        qa.check_model_sensitivity_to_quantization(sim)
        enabled = qa.perform_per_layer_analysis_by_enabling_quant_wrappers(sim, out)
        disabled = qa.perform_per_layer_analysis_by_disabling_quant_wrappers(sim, out)
        sync()
        t2 = time.perf_counter()
        qa.export_per_layer_encoding_min_max_range(sim, out)
        qa.export_per_layer_stats_histogram(sim, out)
        sync()
        t3 = time.perf_counter()
        evals = 3 + len(enabled) + len(disabled)
        images = evals * batch * eval_batches
        return {"device": device, "batch": batch, "eval_batches": eval_batches, "calibrate_s": round(t1 - t0, 4),
                "sweeps_s": round(t2 - t1, 4), "exports_s": round(t3 - t2, 4), "evaluations": evals,
                "evaluated_images_per_s": round(images / (t2 - t1), 1)}
    finally:
        if factory is not None:
            tensor_quantizer._set_op_class_for_testing(prev)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--eval-batches", type=int, default=2)
    ap.add_argument("--cpu-batch", type=int, default=4)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    torch.backends.cudnn.benchmark = True
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    run("cuda", args.batch, args.eval_batches)                       # warm-up (cuDNN autotune, arenas)
    gpu = run("cuda", args.batch, args.eval_batches)
    from oracle.cpu_backend import ReferenceTensorQuantizer          # the reference's own C++ (oracle/_ref), test infra
    cpu = run("cpu", args.cpu_batch, 1, factory=ReferenceTensorQuantizer)
    cpu["kind"] = "reference C++ (oracle/_ref) under the same host layer, torch CPU forward on %d threads" % torch.get_num_threads()
    res = {"workload": "QuantAnalyzer: sensitivity + two per-layer sweeps, ResNet-18 W8A8 tf_enhanced", "gpu": gpu,
           "cpu_reference": cpu,
           "speedup_images_per_s": round(gpu["evaluated_images_per_s"] / cpu["evaluated_images_per_s"], 1)}
    print(json.dumps(res))
    if args.out:
        json.dump(res, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
