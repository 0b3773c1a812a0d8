"""Where does a calibration step spend its host time? cProfile over a few steady-state steps + plain-forward timing."""
import cProfile
import os
import pstats
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

torch.backends.cudnn.benchmark = True
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device("cuda", 0)
sim = bench.build_sim(dev)
xs = [bench.synthetic_batch(i, bench.BATCH, dev) for i in range(4)]
import torchvision
plain = torchvision.models.resnet50().eval().to(dev)
with torch.no_grad():
    for _ in range(3):
        plain(xs[0])
    torch.cuda.synchronize()
    t = time.perf_counter()
    for i in range(8):
        plain(xs[i % 4])
    torch.cuda.synchronize()
    print("plain fp32 forward ms/step", (time.perf_counter() - t) / 8 * 1e3)
sim.compute_encodings(lambda m, _: [m(x) for x in xs], None)
torch.cuda.synchronize()
t = time.perf_counter()
sim.compute_encodings(lambda m, _: [m(xs[i % 4]) for i in range(16)], None)
torch.cuda.synchronize()
print("calibration job ms/step (16 steps, complete job)", (time.perf_counter() - t) / 16 * 1e3)
pr = cProfile.Profile()
pr.enable()
sim.compute_encodings(lambda m, _: [m(xs[i % 4]) for i in range(16)], None)
torch.cuda.synchronize()
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("cumulative").print_stats(35)
st.sort_stats("tottime").print_stats(25)
