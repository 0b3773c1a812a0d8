"""Device time of ONE tf_enhanced statistics call (range already fixed -> a single hist_kernel launch) versus tensor
size, with the host taken out of the picture: 20 calls are captured in a CUDA graph and the graph is timed."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aimet_b200 import ops  # noqa: E402
from aimet_b200.state import StateArena  # noqa: E402

dev = torch.device("cuda", 0)
peak = 6552.3
rows = []
for dtype in (torch.float32, torch.bfloat16):
    for kb in (64, 256, 1024, 4096, 12544, 25088, 51200, 102400, 262144, 1048576):
        n = kb * 1024 // (4 if dtype == torch.float32 else 2)
        reps = 20
        pool = [(torch.randn(n, device=dev) * 2 + 2).to(dtype) for _ in range(max(2, min(reps, (256 * 2**20) // (kb * 1024) + 1)))]
        blk = StateArena.for_device(dev).allocate(1)
        ops.stats_update_impl(pool[0], blk.arena, blk.first, ops.QUANTIZATION_TF_ENHANCED, None, 0)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for i in range(reps):
                ops.stats_update_impl(pool[i % len(pool)], blk.arena, blk.first, ops.QUANTIZATION_TF_ENHANCED, None, 0,
                                      ops.STATS_RANGE_FIXED)
        for _ in range(3):
            g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1000 / (5 * reps)
        gbs = kb * 1024 / us / 1e3
        row = dict(dtype=str(dtype).split(".")[-1], kb=kb, us_per_call=round(us, 2), gbs=round(gbs, 1), frac=round(gbs / peak, 3))
        rows.append(row)
        print(json.dumps(row), flush=True)
        del pool, g
        torch.cuda.empty_cache()
json.dump(rows, open(os.path.join("gpurun_out", "stats_latency.json"), "w"), indent=1)
