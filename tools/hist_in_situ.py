"""Per-launch device-clock durations of the histogram kernel inside a ResNet-50 calibration step (ab_debug_hist_timer):
bytes, microseconds, GB/s of every launch of the last of 4 steps, sorted by size.   python tools/hist_in_situ.py"""
import os
import sys

import torch
import torchvision

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aimet_b200 import _lib  # noqa: E402
from aimet_b200.quantsim import QuantizationSimModel  # noqa: E402
from aimet_b200.quantsim import config as qconfig  # noqa: E402

torch.backends.cudnn.benchmark = True
torch.backends.cudnn.allow_tf32 = False
batch = int(sys.argv[1]) if len(sys.argv) > 1 else 32
model = torchvision.models.resnet50().cuda().eval()
xs = [torch.randn(batch, 3, 224, 224, device="cuda") for _ in range(4)]
sim = QuantizationSimModel(model, dummy_input=xs[0][:1], quant_scheme="tf_enhanced", config_file=qconfig.DEFAULT_CONFIG_PER_CHANNEL)
sim.compute_encodings(lambda m, _: [m(x) for x in xs], None)   # warm-up job
cap = 1024
slots = torch.zeros((cap, 3), dtype=torch.int64, device="cuda")
slots[:, 0] = torch.iinfo(torch.int64).max
L = _lib.load()
L.ab_debug_hist_timer(slots.data_ptr(), cap)
sim.compute_encodings(lambda m, _: [m(x) for x in xs], None)
torch.cuda.synchronize()
used = int(L.ab_debug_hist_timer(None, 0))
rows = [r for r in slots[:used].cpu().tolist() if r[2] >= 64 * 1024]
last = rows[-(len(rows) // 4):]
print(f"{len(last)} launches in the last step")
for b, ns in sorted(((r[2], r[1] - r[0]) for r in last)):
    print(f"{b / 2**20:9.2f} MB {ns / 1000:8.2f} us {b / ns:8.1f} GB/s")
tot_b, tot_ns = sum(r[2] for r in last), sum(r[1] - r[0] for r in last)
print(f"total {tot_b / 2**20:.1f} MB in {tot_ns / 1000:.1f} us = {tot_b / tot_ns:.1f} GB/s")
