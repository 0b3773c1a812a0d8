"""A/B of the deferred multi-tensor statistics inside the ResNet-50 calibration job: ms per step for AB_DEFER_STATS off and
for several flush thresholds (bytes of activations kept alive before a flush is forced).   python tools/defer_ab.py"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from aimet_b200.quantsim import stats_batcher  # noqa: E402

torch.backends.cudnn.benchmark = True
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
device = torch.device("cuda", 0)
sim = bench.build_sim(device)
steps = 32
batches = [bench.synthetic_batch(i, 32, device) for i in range(8)]


def job():
    sim.compute_encodings(lambda m, _: [m(batches[i % 8]) for i in range(steps)], None)
    return sim.get_activation_param_encodings()


def timed():
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    job()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / steps


rows = []
for label, enabled, limit in (("off", False, 0), ("64MB", True, 64 << 20), ("128MB", True, 128 << 20),
                              ("256MB", True, 256 << 20), ("512MB", True, 512 << 20), ("1GB", True, 1 << 30),
                              ("8GB", True, 8 << 30), ("off", False, 0)):
    stats_batcher.ENABLED = enabled
    stats_batcher.FLUSH_BYTES = limit
    job()
    ms = min(timed() for _ in range(3))
    rows.append({"defer": label, "ms_per_step": round(ms, 3), "img_s": round(32 / ms * 1e3, 1),
                 "peak_mem_gb": round(torch.cuda.max_memory_allocated() / 2**30, 2)})
    torch.cuda.reset_peak_memory_stats()
    print(json.dumps(rows[-1]), flush=True)
json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "defer_ab.json"), "w"), indent=1)
