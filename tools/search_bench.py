"""The tf_enhanced grid search alone: 26 560 per-channel records (ResNet-50's weight channels, symmetric, 101 candidates) and
41 per-tensor activation records (asymmetric, 358 candidates), microseconds per quantizer.  python tools/search_bench.py"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aimet_b200 import ops  # noqa: E402
from aimet_b200.state import StateArena  # noqa: E402

dev = torch.device("cuda", 0)
arena = StateArena.for_device(dev)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
rows = []
for label, c, per, scale in (("weights_26560x576", 26560, 576, 0.05), ("weights_26560x4608", 26560, 4608, 0.05),
                             ("llama_channels_65536x4096", 65536, 4096, 0.02), ("activations_41x1M", 41, 1 << 20, 1.0)):
    w = torch.randn(c, per, device=dev) * scale
    blk = arena.allocate(c)
    ops.stats_update_segmented_impl(w, blk.arena, blk.first, c, per, ops.QUANTIZATION_TF_ENHANCED)
    out = torch.empty((c, 5), dtype=torch.float64, device=dev)
    for sym in (True, False):
        ts = []
        for _ in range(4):
            a.record()
            ops.compute_encodings_into(blk.arena, blk.first, c, ops.QUANTIZATION_TF_ENHANCED, 8, sym, False, False, out)
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        ms = sorted(ts[1:])[1]
        rows.append({"case": label, "symmetric": sym, "quantizers": c, "ms": round(ms, 4), "us_per_quantizer": round(ms * 1e3 / c, 4),
                     "sha": hash(out.cpu().numpy().tobytes()) & 0xffffffff})
    del w, blk, out
print(json.dumps(rows, indent=1))
