"""cProfile of the host side of the START of a calibration job (prepare + the first forward, which is host-bound: the GPU
idles behind it), over several short jobs of bench.py's own workload.   python tools/host_profile_job.py [jobs]"""
import cProfile
import os
import pstats
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

jobs = int(sys.argv[1]) if len(sys.argv) > 1 else 6
dev = torch.device("cuda", 0)
torch.backends.cudnn.benchmark = True
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
sim = bench.build_sim(dev)
xs = [bench.synthetic_batch(i, bench.BATCH, dev) for i in range(2)]
first = []


def cb(model, _):
    t = time.perf_counter()
    model(xs[0])
    first.append(time.perf_counter() - t)
    model(xs[1])


for _ in range(2):
    sim.compute_encodings(cb, None)
    sim.get_activation_param_encodings()
torch.cuda.synchronize()
del first[:]
pr = cProfile.Profile()
t0 = time.perf_counter()
pr.enable()
for _ in range(jobs):
    sim.compute_encodings(cb, None)
    sim.get_activation_param_encodings()
pr.disable()
torch.cuda.synchronize()
print("jobs", jobs, "wall ms/job", round((time.perf_counter() - t0) / jobs * 1e3, 2), "first forward host ms (under cProfile)",
      [round(f * 1e3, 2) for f in first])
pstats.Stats(pr).sort_stats("tottime").print_stats(45)
