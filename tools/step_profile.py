"""Kernel-level timeline of steady-state ResNet-50 calibration steps (torch.profiler / CUPTI): GPU busy time per step by
kernel family, idle time between kernels.   python tools/step_profile.py [defer:0|1]"""
import collections
import json
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from aimet_b200.quantsim import stats_batcher  # noqa: E402

stats_batcher.ENABLED = (sys.argv[1] if len(sys.argv) > 1 else "1") != "0"
torch.backends.cudnn.benchmark = True
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
device = torch.device("cuda", 0)
sim = bench.build_sim(device)
batches = [bench.synthetic_batch(i, 32, device) for i in range(8)]
STEPS = 12


def job(n):
    sim.compute_encodings(lambda m, _: [m(batches[i % 8]) for i in range(n)], None)
    return sim.get_activation_param_encodings()


job(STEPS)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    job(STEPS)
    torch.cuda.synchronize()
events = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
events.sort(key=lambda e: e.time_range.start)
t0, t1 = events[0].time_range.start, events[-1].time_range.end
busy = sum(e.time_range.end - e.time_range.start for e in events)
fam = collections.Counter()
cnt = collections.Counter()
for e in events:
    n = e.name
    key = ("ab::" + n.split("ab::")[1].split("(")[0].split("<")[0]) if "ab::" in n else \
        ("memcpy/memset" if n.startswith("Mem") else ("cudnn/cublas conv+gemm" if any(k in n for k in (
            "cudnn", "gemm", "conv", "cutlass", "implicit", "sm80", "sm90", "sm100", "xmma", "winograd", "nchw", "nhwc"))
            else "torch elementwise/other"))
    fam[key] += e.time_range.end - e.time_range.start
    cnt[key] += 1
out = {"defer": stats_batcher.ENABLED, "steps": STEPS, "wall_ms_per_step": round((t1 - t0) / 1e3 / STEPS, 3),
       "gpu_busy_ms_per_step": round(busy / 1e3 / STEPS, 3), "gpu_idle_ms_per_step": round((t1 - t0 - busy) / 1e3 / STEPS, 3),
       "by_family_ms_per_step": {k: [round(v / 1e3 / STEPS, 4), round(cnt[k] / STEPS, 1)] for k, v in fam.most_common()}}
print(json.dumps(out, indent=1))
json.dump(out, open(os.path.join(ROOT, "gpurun_out", f"step_profile_defer{int(stats_batcher.ENABLED)}.json"), "w"), indent=1)
