"""Host cost of the fixed phases of a calibration job (prepare, closing grid search, export), bench.py's own workload:
wall time per phase without a profiler, then cProfile restricted to each phase.   python tools/phase_profile.py [steps]"""
import cProfile
import os
import pstats
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from aimet_b200.quantsim.quantsim import QuantizationSimModel  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 8
dev = torch.device("cuda", 0)
torch.backends.cudnn.benchmark = True
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
sim = bench.build_sim(dev)
xs = [bench.synthetic_batch(i, bench.BATCH, dev) for i in range(steps)]
_prepare = QuantizationSimModel.prepare_sim_for_compute_encodings
_finish = QuantizationSimModel.compute_layer_encodings_for_sim
walls = {"prepare": [], "finish": [], "export": []}
profiles = {k: cProfile.Profile() for k in walls}
profiling = False


def timed(name, fn):
    def inner(*a):
        if profiling:
            profiles[name].enable()
        t = time.perf_counter()
        out = fn(*a)
        walls[name].append(time.perf_counter() - t)
        if profiling:
            profiles[name].disable()
        return out
    return inner


QuantizationSimModel.prepare_sim_for_compute_encodings = staticmethod(timed("prepare", _prepare))
QuantizationSimModel.compute_layer_encodings_for_sim = staticmethod(timed("finish", _finish))
export = timed("export", sim.get_activation_param_encodings)


def cb(model, _):
    for x in xs:
        model(x)


def job():
    sim.compute_encodings(cb, None)
    export()
    torch.cuda.synchronize()


for _ in range(3):
    job()
for v in walls.values():
    del v[:]
for _ in range(5):
    job()
print({k: [round(t * 1e3, 2) for t in v] for k, v in walls.items()}, "host ms per phase, no profiler")
profiling = True
for _ in range(5):
    job()
for k, p in profiles.items():
    print("=" * 30, k)
    pstats.Stats(p).sort_stats("tottime").print_stats(14)
