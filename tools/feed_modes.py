"""Compare ways of feeding host batches to the calibration job (ms/step, complete job, wall clock)."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from aimet_b200.utils import DevicePrefetcher  # noqa: E402

torch.backends.cudnn.benchmark = True
torch.backends.cudnn.allow_tf32 = False
dev = torch.device("cuda", 0)
steps = 32
sim = bench.build_sim(dev)
dev_batches = [bench.synthetic_batch(i, bench.BATCH, dev) for i in range(steps)]
host_batches = [bench.synthetic_batch(i, bench.BATCH, "cpu", pin=True) for i in range(steps)]


def run(name, it):
    def cb(m, _):
        for x in it():
            m(x)
    res = []
    for _ in range(3):
        torch.cuda.synchronize()
        t = time.perf_counter()
        sim.compute_encodings(cb, None)
        sim.get_activation_param_encodings()
        torch.cuda.synchronize()
        res.append((time.perf_counter() - t) * 1e3 / steps)
    print(f"{name:28s} " + " ".join(f"{ms:7.2f}" for ms in res) + " ms/step", flush=True)


import gc  # noqa: E402
if os.environ.get("FREEZE", "0") == "1":
    gc.collect()
    gc.freeze()
    print("gc frozen")
run("resident", lambda: iter(dev_batches))
run("resident", lambda: iter(dev_batches))
run("same-stream copy", lambda: (b.to(dev, non_blocking=True) for b in host_batches))
run("prefetch depth 1", lambda: DevicePrefetcher(host_batches, dev, depth=1))
run("prefetch depth 2", lambda: DevicePrefetcher(host_batches, dev, depth=2))
run("resident", lambda: iter(dev_batches))
