"""BASELINE config 3 on one GPU: MobileNet-v2 quantization-aware training step (W8A8, QDQ forward + STE backward), bf16 and
fp32, against the plain training step of the same model. ms/step by CUDA events; kernel launches of ours per step."""
import os
import sys

import torch
import torchvision

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aimet_b200 import ops  # noqa: E402
from aimet_b200.quantsim import QuantizationSimModel  # noqa: E402

torch.backends.cudnn.benchmark = True
dev = torch.device("cuda", 0)
BATCH = 32


def train_steps(model, x, y, n=10, warm=3):
    opt = torch.optim.SGD(model.parameters(), lr=1e-3, momentum=0.9)
    model.train()

    def step():
        opt.zero_grad(set_to_none=True)
        loss = torch.nn.functional.cross_entropy(model(x).float(), y)
        loss.backward()
        opt.step()

    for _ in range(warm):
        step()
    torch.cuda.synchronize()
    before = ops.launches_total()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        step()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n, (ops.launches_total() - before) / n


for dtype in (torch.float32, torch.bfloat16):
    torch.manual_seed(0)
    x = torch.randn(BATCH, 3, 224, 224, device=dev, dtype=dtype)
    y = torch.randint(0, 1000, (BATCH,), device=dev)
    plain = torchvision.models.mobilenet_v2().to(dev).to(dtype)
    ms_plain, _ = train_steps(plain, x, y)
    model = torchvision.models.mobilenet_v2().to(dev).to(dtype)
    sim = QuantizationSimModel(model, dummy_input=x, quant_scheme="tf_enhanced", default_output_bw=8, default_param_bw=8)
    with torch.no_grad():
        sim.compute_encodings(lambda m, _: m(x), None)
    ms_sim, launches = train_steps(sim.model, x, y)
    by = {k: v for k, v in ops.LAUNCHES.items() if v}
    print(f"{str(dtype):16s} plain {ms_plain:7.2f} ms/step | quantsim QAT {ms_sim:7.2f} ms/step | our launches/step {launches:.0f}")
    print("   cumulative launches by kernel:", by)
