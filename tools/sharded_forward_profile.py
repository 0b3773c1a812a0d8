"""Host time of every forward of a SHARDED calibration job next to the unsharded job on the same rank, and a cProfile of
the sharded job's second forward.   torchrun --nproc-per-node 2 tools/sharded_forward_profile.py [steps]"""
import cProfile
import os
import pstats
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local_rank = int(os.environ.get("LOCAL_RANK", "0"))
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 6
torch.cuda.set_device(local_rank)
dev = torch.device("cuda", local_rank)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
dist.init_process_group("nccl", device_id=dev)
torch.backends.cudnn.benchmark = True
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False

from aimet_b200.distributed import ShardedCalibrator  # noqa: E402

sim = bench.build_sim(dev)
xs = [bench.synthetic_batch(i * world + rank, bench.BATCH, dev) for i in range(steps)]
host, profile = [], None


def cb(model, _):
    for i, x in enumerate(xs):
        if profile is not None and i == 1:
            profile.enable()
        t = time.perf_counter()
        model(x)
        host.append(round((time.perf_counter() - t) * 1e3, 2))
        if profile is not None and i == 1:
            profile.disable()


def job(sharded):
    del host[:]
    dist.barrier()
    torch.cuda.synchronize()
    t = time.perf_counter()
    if sharded:
        ShardedCalibrator(sim).compute_encodings(cb, None)
    else:
        sim.compute_encodings(cb, None)
    sim.get_activation_param_encodings()
    torch.cuda.synchronize()
    return round((time.perf_counter() - t) * 1e3, 2), list(host)


for sharded in (False, True, False, True, False, True):
    out = job(sharded)
    if rank == 0:
        print("sharded" if sharded else "single ", "job ms", out[0], "forward host ms", out[1])
profile = cProfile.Profile()
job(True)
if rank == 0:
    pstats.Stats(profile).sort_stats("tottime").print_stats(25)
dist.destroy_process_group()
