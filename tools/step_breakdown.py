"""What does a calibration step cost beyond the plain forward? ms/step (CUDA events over 20 steps) of
  plain model | wrappers with every quantizer disabled | + parameter QDQ | + activation statistics (the real step)."""
import os
import sys

import torch
import torchvision

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

torch.backends.cudnn.benchmark = True
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device("cuda", 0)
xs = [bench.synthetic_batch(i, bench.BATCH, dev) for i in range(4)]


def timed(fn, n=20):
    with torch.no_grad():
        for i in range(3):
            fn(xs[i % 4])
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(n):
            fn(xs[i % 4])
        b.record()
        torch.cuda.synchronize()
    return a.elapsed_time(b) / n


plain = torchvision.models.resnet50().eval().to(dev)
print(f"plain fp32 forward                      {timed(plain):7.2f} ms")
sim = bench.build_sim(dev)
sim.compute_encodings(lambda m, _: [m(x) for x in xs[:2]], None)
from aimet_b200.quantsim.quantsim import QuantizationSimModel  # noqa: E402
QuantizationSimModel.prepare_sim_for_compute_encodings(sim)
sim.model.eval()
wrappers = [w for _, w in sim.quant_wrappers()]
state = [(q, q.enabled) for w in wrappers for q in w.input_quantizers + w.output_quantizers +
         list(w.param_quantizers.values())]
for q, _ in state:
    q.enabled = False
print(f"wrappers, every quantizer disabled      {timed(sim.model):7.2f} ms")
for w in wrappers:
    for q in w.param_quantizers.values():
        q.enabled = dict((id(a), b) for a, b in state)[id(q)]
timed(sim.model, 2)
print(f"+ parameter QDQ (per-channel weights)   {timed(sim.model):7.2f} ms")
for q, e in state:
    q.enabled = e
timed(sim.model, 2)
print(f"+ activation statistics (real step)     {timed(sim.model):7.2f} ms")
