"""ShardedCalibrator end to end: two processes (ranks) sharing cuda:0 over the gloo backend calibrate a quantsim model
on interleaved batches; the merged encodings must equal, byte for byte, a single-process run over all batches in order.
(Real multi-GPU NCCL runs are exercised by bench.py --gpus N.)"""
import json
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
WORLD = 2
N_BATCH = 6


def make_model_and_batches(scheme, n_batch=N_BATCH, zeros=(0,)):
    import torchvision
    torch.manual_seed(0)
    model = torchvision.models.resnet18().eval().cuda()
    g = torch.Generator().manual_seed(7)
    batches = [(torch.randn(2, 3, 64, 64, generator=g) * (1 + 0.2 * b)).cuda() for b in range(n_batch)]
    for z in zeros:
        batches[z][:] = 0.0  # an all-zero first global batch: tf_enhanced ranges must come from batch 1 (rank 1)
    return model, batches


def encodings_json(sim):
    act, par = sim.get_activation_param_encodings()
    return json.dumps({"activation_encodings": act, "param_encodings": par}, sort_keys=True)


def worker(rank, port, scheme, queue, n_batch=N_BATCH, zeros=(0,)):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=WORLD)
    try:
        torch.backends.cudnn.deterministic = True
        torch.backends.cudnn.benchmark = False
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False
        from aimet_b200.distributed import ShardedCalibrator
        from aimet_b200.quantsim import QuantizationSimModel
        model, batches = make_model_and_batches(scheme, n_batch, zeros)
        sim = QuantizationSimModel(model, dummy_input=torch.randn(2, 3, 64, 64, device="cuda"), quant_scheme=scheme)
        mine = batches[rank::WORLD]
        ShardedCalibrator(sim).compute_encodings(lambda m, _: [m(x) for x in mine], None)
        queue.put((rank, encodings_json(sim)))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("scheme,n_batch,zeros", [
    ("tf_enhanced", N_BATCH, (0,)),
    ("tf", N_BATCH, (0,)),
    ("tf_enhanced", 5, ()),            # uneven shares: rank 0 runs three batches, rank 1 two
    ("tf_enhanced", 7, (0, 1, 2)),     # the first W (and more) global batches are all zeros: a second range round is needed
    ("tf_enhanced", 1, ()),            # fewer batches than ranks: rank 1 has nothing to run and still joins every collective
])
def test_sharded_calibration_equals_single_process(scheme, n_batch, zeros):
    ctx = mp.get_context("spawn")
    queue = ctx.Queue()
    port = 29700 + (os.getpid() % 2000)
    procs = [ctx.Process(target=worker, args=(r, port, scheme, queue, n_batch, zeros)) for r in range(WORLD)]
    for p in procs:
        p.start()
    results = dict(queue.get(timeout=300) for _ in range(WORLD))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    torch.backends.cudnn.deterministic = True
    torch.backends.cudnn.benchmark = False
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    from aimet_b200.quantsim import QuantizationSimModel
    model, batches = make_model_and_batches(scheme, n_batch, zeros)
    sim = QuantizationSimModel(model, dummy_input=torch.randn(2, 3, 64, 64, device="cuda"), quant_scheme=scheme)
    sim.compute_encodings(lambda m, _: [m(x) for x in batches], None)
    single = encodings_json(sim)
    assert results[0] == results[1]
    if results[0] != single:
        a, b = json.loads(results[0]), json.loads(single)
        for sect in b:
            assert set(a[sect]) == set(b[sect]), sect
            for k in b[sect]:
                assert a[sect][k] == b[sect][k], (sect, k)
    assert results[0] == single


def graph_worker(rank, port, queue):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=WORLD)
    try:
        torch.backends.cudnn.deterministic = True
        torch.backends.cudnn.benchmark = False
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False
        from aimet_b200.distributed import ShardedCalibrator
        from aimet_b200.quantsim import QuantizationSimModel
        model, batches = make_model_and_batches("tf_enhanced")
        batches = batches + [b * 0.9 for b in batches[1:]] + [batches[2] * 1.1]     # 12 global batches, 6 per rank
        sim = QuantizationSimModel(model, dummy_input=batches[1], quant_scheme="tf_enhanced")
        ShardedCalibrator(sim).compute_encodings_for_batches(batches[rank::WORLD], cuda_graph=True)
        queue.put((rank, encodings_json(sim)))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_sharded_cuda_graph_calibration_equals_single_process():
    ctx = mp.get_context("spawn")
    queue = ctx.Queue()
    port = 29900 + (os.getpid() % 2000)
    procs = [ctx.Process(target=graph_worker, args=(r, port, queue)) for r in range(WORLD)]
    for p in procs:
        p.start()
    results = dict(queue.get(timeout=300) for _ in range(WORLD))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    torch.backends.cudnn.deterministic = True
    torch.backends.cudnn.benchmark = False
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    from aimet_b200.quantsim import QuantizationSimModel
    model, batches = make_model_and_batches("tf_enhanced")
    batches = batches + [b * 0.9 for b in batches[1:]] + [batches[2] * 1.1]
    sim = QuantizationSimModel(model, dummy_input=batches[1], quant_scheme="tf_enhanced")
    sim.compute_encodings(lambda m, _: [m(x) for x in batches], None)
    assert results[0] == results[1] == encodings_json(sim)
