"""Quantized eval forward of the ResNet-50 quantsim model (every activation and weight quantize-dequantized): ms per
batch with CUDA events. Used to A/B kernel variants in their real surroundings (producer and consumer kernels around each
QDQ), e.g. AB_QDQ_REVERSE=0/1.   python tools/eval_forward.py [--batch 32] [--iters 20] [--dtype fp32|bf16]"""
import argparse
import json
import os
import sys

import torch
import torchvision

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aimet_b200.quantsim import QuantizationSimModel  # noqa: E402
from aimet_b200.quantsim import config as qconfig  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--dtype", default="fp32")
    args = ap.parse_args()
    torch.backends.cudnn.benchmark = True
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    dt = torch.bfloat16 if args.dtype == "bf16" else torch.float32
    torch.manual_seed(0)
    model = torchvision.models.resnet50().cuda().eval().to(dt)
    x = torch.randn(args.batch, 3, 224, 224, device="cuda", dtype=dt)
    sim = QuantizationSimModel(model, dummy_input=x[:1], quant_scheme="tf_enhanced",
                               config_file=qconfig.DEFAULT_CONFIG_PER_CHANNEL)
    sim.compute_encodings(lambda m, _: m(x), None)
    with torch.no_grad():
        for _ in range(5):
            sim.model(x)
            model(x)
        torch.cuda.synchronize()
        res = {}
        for name, m in (("quantsim", sim.model), ("plain", model)):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(args.iters):
                m(x)
            b.record()
            torch.cuda.synchronize()
            res[name + "_ms"] = round(a.elapsed_time(b) / args.iters, 3)
        # the same forward replayed from a CUDA graph: the quantsim layers launch through the C ABI on the capturing stream
        # and keep their encodings on the device, so torch.cuda.graph captures them like any other op
        static_x = x.clone()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3):
                sim.model(static_x)
        torch.cuda.current_stream().wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            static_y = sim.model(static_x)
        eager_y = sim.model(x)
        graph.replay()
        torch.cuda.synchronize()
        res["graph_equals_eager"] = bool(torch.equal(static_y, eager_y))
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(args.iters):
            graph.replay()
        b.record()
        torch.cuda.synchronize()
        res["quantsim_graph_ms"] = round(a.elapsed_time(b) / args.iters, 3)
    res.update(batch=args.batch, dtype=args.dtype, qdq_overhead_ms=round(res["quantsim_ms"] - res["plain_ms"], 3),
               reverse=os.environ.get("AB_QDQ_REVERSE", "1"))
    print(json.dumps(res))


if __name__ == "__main__":
    main()
