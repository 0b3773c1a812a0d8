"""Host and device timeline of ONE real calibration job (the very calls bench.py times: compute_encodings +
get_activation_param_encodings), without any inserted synchronisation: host timestamps and CUDA events are taken at the
phase boundaries; the events are read after the job.   python tools/job_timeline.py [steps]   (works under torchrun)"""
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local_rank = int(os.environ.get("LOCAL_RANK", "0"))
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 8
torch.cuda.set_device(local_rank)
dev = torch.device("cuda", local_rank)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
torch.backends.cudnn.benchmark = True
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False

from aimet_b200.distributed import ShardedCalibrator  # noqa: E402
from aimet_b200.quantsim.quantsim import QuantizationSimModel  # noqa: E402

sim = bench.build_sim(dev)
xs = [bench.synthetic_batch(i * world + rank, bench.BATCH, dev) for i in range(steps)]
marks = []


def mark(label):
    e = torch.cuda.Event(enable_timing=True)
    e.record()
    marks.append((label, time.perf_counter(), e))


_prepare = QuantizationSimModel.prepare_sim_for_compute_encodings
_finish = QuantizationSimModel.compute_layer_encodings_for_sim


def prepare(s):
    _prepare(s)
    mark("prepared (reset + all parameter encodings enqueued)")


def finish(s):
    mark("callback + flush/merge done")
    _finish(s)
    mark("encodings computed")


QuantizationSimModel.prepare_sim_for_compute_encodings = staticmethod(prepare)
QuantizationSimModel.compute_layer_encodings_for_sim = staticmethod(finish)


def cb(model, _):
    for i, x in enumerate(xs):
        model(x)
        if i in (0, 1, steps - 1):
            mark(f"forward[{i}] issued")


def job():
    del marks[:]
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    mark("start")
    if world > 1:
        ShardedCalibrator(sim).compute_encodings(cb, None)
    else:
        sim.compute_encodings(cb, None)
    sim.get_activation_param_encodings()
    mark("exported to host dictionaries")
    torch.cuda.synchronize()
    mark("device idle")


job()
job()
if rank == 0:
    t0, e0 = marks[0][1], marks[0][2]
    print(f"{'phase':58s} {'host ms':>9s} {'device ms':>10s}   (world {world}, {steps} steps)")
    for label, t, e in marks:
        print(f"{label:58s} {(t - t0) * 1e3:9.2f} {e0.elapsed_time(e):10.2f}")
if world > 1:
    dist.destroy_process_group()
