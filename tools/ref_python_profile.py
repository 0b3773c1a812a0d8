"""cProfile of one calibration job through the reference's own Python (baseline/_ref) on the registered drop-ins.
python tools/ref_python_profile.py [steps]"""
import cProfile
import os
import pstats
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import ref_python_driver as drv  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
real_stdout = os.dup(1)
os.dup2(2, 1)
quantsim = drv.setup()
import torch  # noqa: E402
import torchvision  # noqa: E402
from aimet_common.defs import QuantScheme  # noqa: E402

drv.use_backend("native")
dev = torch.device("cuda:0")
torch.manual_seed(0)
model = torchvision.models.resnet50().eval().to(dev)
xs = [torch.randn(32, 3, 224, 224, generator=torch.Generator().manual_seed(1000 + b)).to(dev) for b in range(steps)]
cfg = os.path.join(drv.REF, "aimet_common", "quantsim_config", "default_config_per_channel.json")
sim = quantsim.QuantizationSimModel(model, dummy_input=xs[0][:1], quant_scheme=QuantScheme.post_training_tf_enhanced,
                                    default_output_bw=8, default_param_bw=8, config_file=cfg)


def cal(m, _):
    with torch.no_grad():
        for x in xs:
            m(x)


def job():
    sim.compute_encodings(cal, None)
    return sim.get_activation_param_encodings()


job()
torch.cuda.synchronize()
t = time.perf_counter()
job()
torch.cuda.synchronize()
wall = time.perf_counter() - t
pr = cProfile.Profile()
pr.enable()
job()
torch.cuda.synchronize()
pr.disable()
os.dup2(real_stdout, 1)
print("job seconds without profiler", round(wall, 4))
pstats.Stats(pr).sort_stats("tottime").print_stats(30)
