"""cProfile of the host side of one quantsim eval forward (bf16 ResNet-50, batch 32: GPU time 4 ms, so the wall time is the
Python / ctypes path). python tools/host_profile.py"""
import cProfile
import os
import pstats
import sys
import time

import torch
import torchvision

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aimet_b200.quantsim import QuantizationSimModel  # noqa: E402
from aimet_b200.quantsim import config as qconfig  # noqa: E402

torch.backends.cudnn.benchmark = True
model = torchvision.models.resnet50().cuda().eval().to(torch.bfloat16)
x = torch.randn(32, 3, 224, 224, device="cuda", dtype=torch.bfloat16)
sim = QuantizationSimModel(model, dummy_input=x[:1], quant_scheme="tf_enhanced", config_file=qconfig.DEFAULT_CONFIG_PER_CHANNEL)
sim.compute_encodings(lambda m, _: m(x), None)
with torch.no_grad():
    for _ in range(5):
        sim.model(x)
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(20):
        sim.model(x)
    host = (time.perf_counter() - t) / 20
    torch.cuda.synchronize()
    total = (time.perf_counter() - t) / 20
    print("host issue ms/forward", round(host * 1e3, 3), "wall ms/forward", round(total * 1e3, 3))
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(20):
        sim.model(x)
    pr.disable()
    torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("tottime").print_stats(28)
