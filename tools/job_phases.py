"""Where does one calibration JOB spend its time? Phases of bench.py's timed job, each closed with a device synchronise:
prepare, the K forward passes, (N>1) merge, encodings on the device, export to host dictionaries. Works under torchrun.
Never a source of bench numbers (the synchronisation points are not in the real job)."""
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
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 32
torch.cuda.set_device(local_rank)
dev = torch.device("cuda", local_rank)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
torch.backends.cudnn.benchmark = True
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False

from aimet_b200.distributed import ShardedCalibrator  # noqa: E402
from aimet_b200.quantsim.quantsim import QuantizationSimModel, in_eval_mode  # noqa: E402

sim = bench.build_sim(dev)
xs = [bench.synthetic_batch(i * world + rank, bench.BATCH, dev) for i in range(steps)]


def tick(label, t0, out):
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    out.append((label, (t1 - t0) * 1e3))
    return t1


def job():
    out = []
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t = time.perf_counter()
    QuantizationSimModel.prepare_sim_for_compute_encodings(sim)
    cal = batcher = None
    if world > 1:
        cal = ShardedCalibrator(sim)
        cal._attach(staging=False)
    else:
        from aimet_b200.quantsim.stats_batcher import StatsBatcher
        batcher = StatsBatcher.attach(sim)
    t = tick("prepare (reset + all parameter encodings)", t, out)
    from aimet_b200.quantsim.qc_quantize_op import CalibrationJob
    with in_eval_mode(sim.model), torch.no_grad(), CalibrationJob(sim):
        sim.model(xs[0])
        t = tick("forward[0]", t, out)
        if steps > 1:
            sim.model(xs[1])
            t = tick("forward[1] (+ range exchange when sharded)", t, out)
        for i in range(2, steps):
            sim.model(xs[i])
        t = tick(f"forward[2..{steps - 1}]", t, out)
    if cal is not None:
        cal._merge()
        cal._detach()
        t = tick("merge", t, out)
    elif batcher is not None:
        batcher.flush()
        batcher.detach()
    QuantizationSimModel.compute_layer_encodings_for_sim(sim)
    t = tick("encodings on device", t, out)
    sim.get_activation_param_encodings()
    t = tick("export to host dicts", t, out)
    return out


job()
res = job()
if rank == 0:
    total = sum(ms for _, ms in res)
    for label, ms in res:
        print(f"{label:50s} {ms:9.2f} ms")
    print(f"{'total':50s} {total:9.2f} ms  ({total / steps:.2f} ms/step over {steps} steps, world {world})")
if world > 1:
    dist.destroy_process_group()
