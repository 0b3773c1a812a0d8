"""Range-learning QDQ: the fused forward / backward kernels against the torch-op sequence the reference launches on a GPU.

Algorithmic bytes: forward 2 s per element (read x, write y), backward 3 s (read x and grad, write grad_x). The torch-op
baseline is oracle/range_learning.py run on CUDA tensors -- the same operations, in the same order, as the reference's
quantsim_straight_through_grad.py (its only difference from the reference on a GPU is the module it is imported from).
Times with CUDA events on the launching stream; inputs rotate through a pool larger than L2.
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from aimet_b200 import ops  # noqa: E402
from oracle import range_learning as rl  # noqa: E402  (baseline leg only; this script lives under tests/ because it uses the oracle)

L2_BYTES = 126 * 2**20


def peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p))["hbm_gbs"], "measured"
    return 6650.0, "fallback"


def time_ms(fn, pool, iters=20, warmup=4):
    for i in range(warmup):
        fn(pool[i % len(pool)])
    torch.cuda.synchronize()
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    for i in range(iters):
        fn(pool[i % len(pool)])
    stop.record()
    torch.cuda.synchronize()
    return start.elapsed_time(stop) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes-mb", type=float, nargs="*", default=[16, 64, 256, 1024])
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    peak, which = peak_gbs()
    dev = torch.device("cuda", 0)
    rows = []
    for dtype, es, dname in ((torch.float32, 4, "fp32"), (torch.bfloat16, 2, "bf16")):
        for mb in args.sizes_mb:
            nbytes = int(mb * 2**20)
            c = 2048
            per = nbytes // es // c
            n = c * per
            npool = max(2, min(6, int(2 * L2_BYTES // nbytes) + 1))
            pool = [(torch.randn(c, per, device=dev) * 1.5 + 0.3).to(dtype) for _ in range(npool)]
            grad = torch.randn(c, per, device=dev).to(dtype)
            configs = {
                "per_tensor_asym8": (torch.tensor([-3.0], device=dev, dtype=dtype), torch.tensor([4.0], device=dev, dtype=dtype),
                                     rl.ASYMMETRIC),
                "per_channel_sym8": (torch.full((c,), -4.0, device=dev, dtype=dtype), torch.full((c,), 4.0, device=dev, dtype=dtype),
                                     rl.SIGNED_SYMMETRIC),
            }
            for cname, (mn, mx, mode) in configs.items():
                def fused_fwd(x):
                    return ops.lg_qdq_fwd_impl(x, mn, mx, 8, mode, False, 0, gate=True)

                def fused_bwd(x):
                    return ops.lg_qdq_bwd_impl(x, grad, mn, mx, 8, mode, False, 0)

                def torch_fwd(x):
                    rl.gate(mn, mx)
                    return rl.forward(x, mn, mx, 8, mode, False, 0)

                saved_holder = {}

                def torch_bwd(x):
                    # the reference keeps x_quant / mask / delta / offset from the forward; give the baseline that for free
                    if saved_holder.get("x") is not x:
                        saved_holder["x"] = x
                        saved_holder["saved"] = rl.forward(x, mn, mx, 8, mode, False, 0)[1]
                    return rl.backward(grad, saved_holder["saved"])

                for kname, fn, alg in (("fwd", fused_fwd, 2 * es * n), ("bwd", fused_bwd, 3 * es * n)):
                    torch.cuda.reset_peak_memory_stats()
                    base_mem = torch.cuda.memory_allocated()
                    ms = time_ms(fn, pool)
                    fused_peak = torch.cuda.max_memory_allocated() - base_mem
                    small = nbytes <= 256 * 2**20
                    ref_ms = ref_peak = None
                    if small or kname == "fwd":
                        torch.cuda.reset_peak_memory_stats()
                        ref_pool = pool[:1] if kname == "bwd" else pool
                        ref_ms = time_ms(torch_fwd if kname == "fwd" else torch_bwd, ref_pool, iters=5, warmup=2)
                        ref_peak = torch.cuda.max_memory_allocated() - base_mem
                        saved_holder.clear()
                    gbs = alg / ms / 1e6
                    row = dict(kernel=f"lg_{kname}", config=cname, dtype=dname, mb=mb, ms=round(ms, 4), gbs=round(gbs, 1),
                               frac=round(gbs / peak, 3), peak=which,
                               torch_ops_ms=None if ref_ms is None else round(ref_ms, 4),
                               speedup=None if ref_ms is None else round(ref_ms / ms, 1),
                               fused_extra_mb=round(fused_peak / 2**20, 1),
                               torch_extra_mb=None if ref_peak is None else round(ref_peak / 2**20, 1))
                    rows.append(row)
                    print(json.dumps(row), flush=True)
            del pool, grad
            torch.cuda.empty_cache()
    if args.out:
        json.dump(rows, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
