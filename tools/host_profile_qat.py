"""cProfile of the host side of the MobileNet-v2 QAT step (BASELINE config 3 shape, one GPU): the step is host-bound, so this
is where its time goes. python tools/host_profile_qat.py [N lines]"""
import cProfile
import os
import pstats
import sys
import time

import torch
import torchvision

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aimet_b200.quantsim import QuantizationSimModel  # noqa: E402

torch.backends.cudnn.benchmark = True
dtype = torch.bfloat16
x = torch.randn(32, 3, 224, 224, device="cuda", dtype=dtype)
y = torch.randint(0, 1000, (32,), device="cuda")
model = torchvision.models.mobilenet_v2().cuda().to(dtype)
sim = QuantizationSimModel(model, dummy_input=x, quant_scheme="tf_enhanced")
with torch.no_grad():
    sim.compute_encodings(lambda m, _: m(x), None)
opt = torch.optim.SGD(sim.model.parameters(), lr=1e-3, momentum=0.9)
sim.model.train()


def step():
    opt.zero_grad(set_to_none=True)
    loss = torch.nn.functional.cross_entropy(sim.model(x).float(), y)
    loss.backward()
    opt.step()


for _ in range(5):
    step()
torch.cuda.synchronize()
t = time.perf_counter()
for _ in range(10):
    step()
host = (time.perf_counter() - t) / 10
torch.cuda.synchronize()
print("host issue ms/step", round(host * 1e3, 2), "wall ms/step", round((time.perf_counter() - t) / 10 * 1e3, 2))
pr = cProfile.Profile()
pr.enable()
for _ in range(10):
    step()
pr.disable()
torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("tottime").print_stats(int(sys.argv[1]) if len(sys.argv) > 1 else 40)
