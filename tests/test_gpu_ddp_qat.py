"""Quantization-aware training under torch.nn.parallel.DistributedDataParallel (BASELINE configs[2] is QAT "on 8 x B200",
SURVEY section 8e: plain DDP, the QDQ / statistics kernels are rank-local): two ranks, each with half the batch, must produce
the gradients one process computes on the whole batch -- including the straight-through gate on the parameter gradients
(reference: SteGatingFuncForParameters, aimet_torch/v1/qc_quantize_op.py:1314-1366), which a DDP reducer silently drops
when it is applied to `param.grad` after the fact. Also the reference's own multi-GPU form, torch.nn.DataParallel replicas
(`_is_replica`, qc_quantize_op.py:257-268, 785-788; its test: test_quantizer.py:1087-1138), when two GPUs are visible.
"""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
WORLD = 2


def build():
    from aimet_b200.quantsim import QuantizationSimModel
    torch.backends.cudnn.deterministic = True
    torch.backends.cudnn.benchmark = False
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Conv2d(3, 16, 3, padding=1), torch.nn.ReLU(), torch.nn.Conv2d(16, 8, 3, padding=1),
                                torch.nn.ReLU(), torch.nn.AdaptiveAvgPool2d(1), torch.nn.Flatten(), torch.nn.Linear(8, 4)).cuda()
    x = torch.randn(8, 3, 16, 16, generator=torch.Generator().manual_seed(1)).cuda()
    sim = QuantizationSimModel(model, dummy_input=x, quant_scheme="tf_enhanced")
    sim.compute_encodings(lambda m, _: m(x), None)
    # eval mode keeps the calibrated parameter encodings (training mode re-derives them from the weights on every
    # forward); one weight is then moved far outside its encoding range: its gradient must be gated to zero
    sim.model.eval()
    with torch.no_grad():
        sim.model[0]._module_to_wrap.weight[0, 0, 0, 0] = 40.0      # (sim.model is a copy of `model`)
        sim.model[2]._module_to_wrap.weight[0, 0, 0, 0] = 40.0
    return sim, x


def grads_of(module):
    return {n: p.grad.detach().cpu().clone() for n, p in module.named_parameters() if p.grad is not None}


def worker(rank, port, queue):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=WORLD)
    try:
        sim, x = build()
        ddp = torch.nn.parallel.DistributedDataParallel(sim.model, device_ids=[0])
        local = x[rank * 4:(rank + 1) * 4]
        out = ddp(local)
        out.square().mean().backward()
        # numpy, not tensors: a CPU tensor travels through a multiprocessing queue as a shared-memory handle that dies with
        # this process
        queue.put((rank, out.detach().cpu().numpy(), {k: v.numpy() for k, v in grads_of(sim.model).items()}))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_ddp_qat_step_equals_single_process_on_the_whole_batch():
    ctx = mp.get_context("spawn")
    queue = ctx.Queue()
    port = 30700 + (os.getpid() % 2000)
    procs = [ctx.Process(target=worker, args=(r, port, queue)) for r in range(WORLD)]
    for p in procs:
        p.start()
    got = {}
    for _ in range(WORLD):
        rank, out, grads = queue.get(timeout=300)
        got[rank] = (torch.from_numpy(out), {k: torch.from_numpy(v) for k, v in grads.items()})
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    from aimet_b200.quantsim import qc_quantize_op
    results = {}
    for legacy in (False, True):          # the hook form, and the reference's gating-function form
        qc_quantize_op.ALWAYS_GATE_AND_CLONE = legacy
        try:
            sim, x = build()
            out = sim.model(x)
            out.square().mean().backward()
            results[legacy] = (out.detach().cpu(), grads_of(sim.model))
        finally:
            qc_quantize_op.ALWAYS_GATE_AND_CLONE = False
    out_ref, grads_ref = results[True]
    assert torch.equal(results[False][0], out_ref)
    for k, g in grads_ref.items():
        assert torch.equal(results[False][1][k], g), k
    w_key = "2._module_to_wrap.weight"
    assert grads_ref[w_key][0, 0, 0, 0] == 0 and grads_ref[w_key].abs().sum() > 0       # the gate is not trivial
    # the reference's quirk, reproduced: the first layer's input does not require grad, so its gating function never runs
    assert grads_ref["0._module_to_wrap.weight"][0, 0, 0, 0] != 0
    assert torch.allclose(torch.cat([got[0][0], got[1][0]]), out_ref, rtol=1e-5, atol=1e-6)
    assert got[0][1].keys() == grads_ref.keys()
    for k, g in grads_ref.items():
        assert torch.equal(got[0][1][k], got[1][1][k]), k                               # all-reduced: ranks agree
        assert torch.allclose(got[0][1][k], g, rtol=1e-4, atol=1e-7), k
    assert got[0][1][w_key][0, 0, 0, 0] == 0                                            # the gate survived the reducer


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="torch.nn.DataParallel replicas need two GPUs")
def test_data_parallel_replicas_match_single_gpu():
    """reference test_quantizer.py:1087-1138"""
    sim, x = build()
    out_single = sim.model(x)
    out_single.flatten().sum().backward()
    grads_single = grads_of(sim.model)
    weights_before = {n: p.detach().clone() for n, p in sim.model.named_parameters()}
    sim.model.zero_grad(set_to_none=True)
    dp = torch.nn.DataParallel(sim.model, device_ids=[0, 1])
    out_multi = dp(x)
    assert torch.allclose(out_multi, out_single, rtol=1e-5, atol=1e-6)
    for n, p in sim.model.named_parameters():
        assert torch.equal(p.detach(), weights_before[n]), n          # the originals are not left quantized
    out_multi.flatten().sum().backward()
    for n, g in grads_of(sim.model).items():
        assert torch.allclose(g, grads_single[n], rtol=1e-4, atol=1e-6), n
