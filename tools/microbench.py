"""Kernel-level bandwidth sweep (BASELINE.json config 5): achieved GB/s of each hot-path kernel vs the measured HBM peak.
Times with CUDA events on the launching stream; inputs rotate through a pool larger than L2 so every pass reads HBM."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from aimet_b200 import ops  # noqa: E402
from aimet_b200.state import StateArena  # noqa: E402

L2_BYTES = 126 * 2**20


def peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p))["hbm_gbs"], "measured"
    return 6650.0, "fallback"


def time_ms(fn, pool, iters=20, warmup=5):
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
    ap.add_argument("--sizes-mb", type=float, nargs="*", default=[1, 16, 64, 256, 1024, 4096])
    ap.add_argument("--out", default=None)
    ap.add_argument("--only", nargs="*", default=None, help="substrings of kernel names to run")
    args = ap.parse_args()
    peak, which = peak_gbs()
    dev = torch.device("cuda", 0)
    arena = StateArena.for_device(dev)
    rows = []
    for dtype, es, name in ((torch.float32, 4, "fp32"), (torch.bfloat16, 2, "bf16")):
        for mb in args.sizes_mb:
            nbytes = int(mb * 2**20)
            n = nbytes // es
            npool = max(2, min(8, int(2 * L2_BYTES // nbytes) + 1)) if nbytes < 2 * L2_BYTES else 2
            if nbytes * npool > 40 * 2**30:
                npool = 2
            pool = [(torch.randn(n, device=dev) * 2 + 2).to(dtype) for _ in range(npool)]
            grads = [torch.randn(n, device=dev).to(dtype) for _ in range(2)]
            blk = arena.allocate(2)
            ops.stats_update_impl(pool[0], blk.arena, blk.first, ops.QUANTIZATION_TF_ENHANCED, None, 0)   # fixes the range
            c, per = 2048, n // 2048
            params = ops.per_channel_params([-4.0] * c, [8.0] * c, 8).cuda()
            npc = c * per
            # blockwise (LPBQ-style) encodings: one {min, max, delta, offset} per run of 64 / 16 elements
            nb64, nb16 = n // 64, n // 16
            enc64 = [torch.full((nb64, 1), v, device=dev) for v in (-4.0, 8.0, 12.0 / 255, -85.0)]
            enc16 = [torch.full((nb16, 1), v, device=dev) for v in (-4.0, 8.0, 12.0 / 255, -85.0)]
            cases = {
                "qdq_broadcast_block64": (lambda x: ops.qdq_broadcast_impl(x[:nb64 * 64].view(nb64, 64), *enc64),
                                          2 * es * nb64 * 64),
                "qdq_broadcast_block16": (lambda x: ops.qdq_broadcast_impl(x[:nb16 * 16].view(nb16, 16), *enc16),
                                          2 * es * nb16 * 16),
                "qdq_per_tensor_bw8": (lambda x: ops.qdq_per_tensor_impl(x, -4.0, 8.0, 8, 0, 0), 2 * es * n),
                "qdq_per_tensor_bw4": (lambda x: ops.qdq_per_tensor_impl(x, -4.0, 8.0, 4, 0, 0), 2 * es * n),
                "qdq_per_tensor_bw16": (lambda x: ops.qdq_per_tensor_impl(x, -4.0, 8.0, 16, 0, 0), 2 * es * n),
                "quantize_to_grid_bw8": (lambda x: ops.quantize_to_grid_impl(x, -4.0, 8.0, 8, 0, True, 0), 2 * es * n),
                "qdq_per_channel_c2048": (lambda x: ops.qdq_per_channel_impl(x[:npc], params, c, per, 0, 0),
                                          2 * es * npc),
                "ste_bwd": (lambda x: ops.ste_bwd_impl(x, grads[0], -4.0, 8.0), 3 * es * n),
                "stats_tf_minmax": (lambda x: ops.stats_update_impl(x, blk.arena, blk.first + 1, ops.QUANTIZATION_TF,
                                                                    None, 0), es * n),
                "stats_tfe_hist_steady": (lambda x: ops.stats_update_impl(x, blk.arena, blk.first,
                                                                          ops.QUANTIZATION_TF_ENHANCED, None, 0),
                                          es * n),
            }
            for kname, (fn, alg_bytes) in cases.items():
                if args.only and not any(o in kname for o in args.only):
                    continue
                ms = time_ms(fn, pool)
                gbs = alg_bytes / ms / 1e6
                row = dict(kernel=kname, dtype=name, mb=mb, ms=round(ms, 4), gbs=round(gbs, 1),
                           frac=round(gbs / peak, 3), peak=which)
                rows.append(row)
                print(json.dumps(row), flush=True)
            del pool, grads
            torch.cuda.empty_cache()
    if args.out:
        json.dump(rows, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
