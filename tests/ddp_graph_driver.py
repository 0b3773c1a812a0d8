"""Helper process of test_gpu_train_graph.py: a QAT step under DistributedDataParallel (NCCL, world size 1 -- the reducer and
its all-reduce run exactly as with more ranks) eager and replayed from a CUDA graph; prints both loss sequences as JSON."""
import json
import os
import sys

os.environ["TORCH_NCCL_ASYNC_ERROR_HANDLING"] = "0"     # torch's rule for capturing DDP's collectives
os.environ["NCCL_ASYNC_ERROR_HANDLING"] = "0"
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
os.environ.setdefault("MASTER_PORT", sys.argv[1] if len(sys.argv) > 1 else "29533")
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.test_gpu_train_graph import _loss, _setup  # noqa: E402

torch.cuda.set_device(0)
dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", 0))


def wrap(model):
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        ddp = torch.nn.parallel.DistributedDataParallel(model, device_ids=[0])
    torch.cuda.current_stream().wait_stream(side)
    return ddp


sim_e, opt_e, xs, ys = _setup()
ddp_e = wrap(sim_e.model)
eager = []
for x, y in zip(xs, ys):
    opt_e.zero_grad(set_to_none=True)
    loss = _loss(ddp_e(x), y)
    loss.backward()
    opt_e.step()
    eager.append(float(loss))
sim_g, opt_g, xs, ys = _setup()
ddp_g = wrap(sim_g.model)
step = sim_g.capture_train_step(_loss, opt_g, (xs[0],), ys[0], warmup=11, model=ddp_g)
graphed = [float(step(x, target=y)) for x, y in zip(xs, ys)]
close = all(torch.allclose(a, b, rtol=1e-4, atol=1e-6)
            for (_, a), (_, b) in zip(sim_e.model.named_parameters(), sim_g.model.named_parameters()))
print("RESULT " + json.dumps({"eager": eager, "graphed": graphed, "parameters_close": close}), flush=True)
torch.cuda.synchronize()
os._exit(0)
