"""BASELINE config 3: MobileNet-v2 quantization-aware training, W8A8 (QDQ forward + straight-through backward), bf16, one
process per GPU under DistributedDataParallel (NCCL), batch 32 per GPU. Reports, per dtype: ms per step and img/s (max over
ranks, CUDA events) for the plain model, the quantsim model eager, and -- on one GPU -- the quantsim step replayed from a
CUDA graph (QuantizationSimModel.capture_train_step); own kernel launches per step.

    python tools/qat_ddp.py                                                     # 1 GPU
    python -m torch.distributed.run --nproc-per-node 8 tools/qat_ddp.py        # 8 GPUs
"""
import json
import os
import sys

import torch
import torch.distributed as dist
import torchvision

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aimet_b200 import ops  # noqa: E402
from aimet_b200.quantsim import QuantizationSimModel  # noqa: E402

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local_rank = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local_rank)
dev = torch.device("cuda", local_rank)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ["TORCH_NCCL_ASYNC_ERROR_HANDLING"] = "0"   # torch's rule for capturing DDP's all-reduces in a CUDA graph
    os.environ["NCCL_ASYNC_ERROR_HANDLING"] = "0"
    dist.init_process_group("nccl", device_id=dev)
torch.backends.cudnn.benchmark = True
BATCH, STEPS, WARM = 32, 20, 5


def loss_fn(out, y):
    return torch.nn.functional.cross_entropy(out.float(), y)


def timed(step):
    for _ in range(WARM):
        step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    before = ops.launches_total()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(STEPS):
        step()
    b.record()
    torch.cuda.synchronize()
    ms = torch.tensor([a.elapsed_time(b) / STEPS], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms), (ops.launches_total() - before) / STEPS


def eager_step(model, opt, x, y):
    def step():
        opt.zero_grad(set_to_none=True)
        loss_fn(model(x), y).backward()
        opt.step()
    return step


rows = []
for dtype in (torch.bfloat16, torch.float32):
    torch.manual_seed(0)
    x = torch.randn(BATCH, 3, 224, 224, device=dev, dtype=dtype)
    y = torch.randint(0, 1000, (BATCH,), device=dev)
    plain = torchvision.models.mobilenet_v2().to(dev).to(dtype).train()
    wrapped = torch.nn.parallel.DistributedDataParallel(plain, device_ids=[local_rank]) if world > 1 else plain
    ms_plain, _ = timed(eager_step(wrapped, torch.optim.SGD(plain.parameters(), lr=1e-3, momentum=0.9), x, y))
    del wrapped, plain

    torch.manual_seed(0)
    model = torchvision.models.mobilenet_v2().to(dev).to(dtype)
    sim = QuantizationSimModel(model, dummy_input=x, quant_scheme="tf_enhanced", default_output_bw=8, default_param_bw=8)
    sim.compute_encodings(lambda m, _: m(x), None)
    sim.model.train()
    if world > 1:
        side = torch.cuda.Stream()          # torch's rule: a DDP wrapper that will be captured is built on a side stream
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            wrapped = torch.nn.parallel.DistributedDataParallel(sim.model, device_ids=[local_rank])
        torch.cuda.current_stream().wait_stream(side)
    else:
        wrapped = sim.model
    opt = torch.optim.SGD(sim.model.parameters(), lr=1e-3, momentum=0.9)
    ms_sim, launches = timed(eager_step(wrapped, opt, x, y))
    row = {"dtype": str(dtype).split(".")[-1], "gpus": world, "batch_per_gpu": BATCH, "plain_ms": round(ms_plain, 3),
           "quantsim_ms": round(ms_sim, 3), "quantsim_img_s": round(BATCH * world / ms_sim * 1e3, 1),
           "own_launches_per_step": launches}
    try:
        graphed = sim.capture_train_step(loss_fn, opt, (x,), y, warmup=3 if world == 1 else 11, model=wrapped)
        ms_graph, _ = timed(lambda: graphed(x, target=y))
        row.update(quantsim_cuda_graph_ms=round(ms_graph, 3),
                   quantsim_cuda_graph_img_s=round(BATCH * world / ms_graph * 1e3, 1))
    except Exception as exc:   # pylint: disable=broad-except
        row["quantsim_cuda_graph_error"] = repr(exc)[:300]
    rows.append(row)
    if rank == 0:
        print(json.dumps(row), flush=True)
    graphed = None
    del wrapped, sim, model, opt
    torch.cuda.empty_cache()
if rank == 0:
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", f"qat_ddp_n{world}.json")
    json.dump(rows, open(out, "w"), indent=1)
if world > 1:
    # a process group whose collectives were captured into CUDA graphs does not always tear down cleanly: synchronise, meet
    # at a barrier, and leave without running the destructors
    torch.cuda.synchronize()
    dist.barrier()
    sys.stdout.flush()
    os._exit(0)
