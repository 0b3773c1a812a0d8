"""BASELINE.json configs[0], [3] and [4] as short, driver-runnable measurements (bench.py puts their results into its JSON line
as `other_configs`; each also runs alone: python tools/other_configs.py [resnet18|qat|llama]).

  configs[0]  ResNet-18 W8A8 tf_enhanced, default config, one calibration batch of 32 x 3 x 224 x 224 + one quantized eval
              forward: this repo on the GPU, and the same host layer on the reference's own C++ on the host cores
              (`cpu_reference`; oracle/ is executed here only as that baseline). Encodings of the two are compared.
  configs[2]  (index 2 in BASELINE.json's list, "config 3" in SURVEY section 8d) MobileNet-v2 QAT W8A8 (QDQ forward + STE
              backward), bf16, batch 32 per GPU, DistributedDataParallel when launched under torchrun: plain step, quantsim
              step eager, and -- on one GPU -- the quantsim step replayed from a CUDA graph.
  configs[3]  Llama-2-7B-shaped bf16 weights (224 matrices, 6.48 G weights): per-channel tf / tf_enhanced weight encodings,
              W4 per-channel QDQ over all of them, tf statistics + 16-bit QDQ on [8, 2048, 4096] / [8, 2048, 11008] activations.
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _events(fn, reps, warm):
    import torch
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def resnet18(device, with_cpu_reference=True):
    """configs[0]"""
    import torch
    import torchvision

    import bench
    from aimet_b200.quantsim import QuantizationSimModel, tensor_quantizer

    def build(dev, factory=None):
        prev = tensor_quantizer._set_op_class_for_testing(factory) if factory is not None else None
        try:
            torch.manual_seed(0)
            model = torchvision.models.resnet18().eval().to(dev)
            return QuantizationSimModel(model, dummy_input=torch.zeros(2, 3, 224, 224, device=dev), quant_scheme="tf_enhanced",
                                        default_output_bw=8, default_param_bw=8, in_place=True)
        finally:
            if prev is not None:
                tensor_quantizer._set_op_class_for_testing(prev)

    x_cpu = bench.synthetic_batch(0, 32)
    x = x_cpu.to(device)
    sim = build(device)

    def calibrate():
        sim.compute_encodings(lambda m, _: m(x), None)
        return sim.get_activation_param_encodings()

    ms_cal = _events(calibrate, 5, 2)
    with torch.no_grad():
        ms_fwd = _events(lambda: sim.model(x), 10, 3)
        plain = QuantizationSimModel.get_original_model(sim.model)
        ms_plain = _events(lambda: plain(x), 10, 3)
    act, par = calibrate()
    out = {"workload": "ResNet-18 W8A8 tf_enhanced, default config, one batch of 32 x 3 x 224 x 224 (BASELINE configs[0])",
           "compute_encodings_ms": round(ms_cal, 3), "quantized_eval_forward_ms": round(ms_fwd, 3),
           "plain_eval_forward_ms": round(ms_plain, 3), "num_activation_encodings": len(act), "num_param_encodings": len(par)}
    if with_cpu_reference:
        from oracle import cpu_backend
        factory = cpu_backend.best_cpu_backend()
        threads = bench.host_threads()
        csim = build("cpu", factory)
        prev = tensor_quantizer._set_op_class_for_testing(factory)
        try:
            t0 = time.perf_counter()
            csim.compute_encodings(lambda m, _: m(x_cpu), None)
            cact, cpar = csim.get_activation_param_encodings()
            t_cal = time.perf_counter() - t0
            with torch.no_grad():
                csim.model(x_cpu)
                t0 = time.perf_counter()
                csim.model(x_cpu)
                t_fwd = time.perf_counter() - t0
        finally:
            tensor_quantizer._set_op_class_for_testing(prev)
        out["cpu_reference"] = {"kind": factory.KIND, "cores": threads, "compute_encodings_ms": round(t_cal * 1e3, 1),
                                "quantized_eval_forward_ms": round(t_fwd * 1e3, 1),
                                "note": "torch CPU forward on all threads, the reference's C++ statistics / QDQ single-threaded"}
        # the GPU forward (cuDNN) and the CPU forward (oneDNN) differ in the last bits, so activation encodings can differ in
        # the last digits; parameter encodings depend on the weights alone and must be identical
        out["param_encodings_equal_cpu_reference"] = json.dumps(par, sort_keys=True) == json.dumps(cpar, sort_keys=True)
        out["speedup_compute_encodings"] = round(t_cal * 1e3 / ms_cal, 1)
    return out


def qat(device, world=1, local_rank=0, graph=True):
    """configs[2]: MobileNet-v2 QAT step. Under torchrun (world > 1) the process group must exist already."""
    import torch
    import torch.distributed as dist
    import torchvision

    from aimet_b200 import ops
    from aimet_b200.quantsim import QuantizationSimModel
    batch, steps, warm = 32, 10, 4
    dtype = torch.bfloat16

    def loss_fn(out, y):
        return torch.nn.functional.cross_entropy(out.float(), y)

    def timed(step):
        for _ in range(warm):
            step()
        if world > 1:
            dist.barrier()
        ms = torch.tensor([_events(step, steps, 0)], device=device, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)

    def eager(model, opt):
        def step():
            opt.zero_grad(set_to_none=True)
            loss_fn(model(x), y).backward()
            opt.step()
        return step

    torch.manual_seed(0)
    x = torch.randn(batch, 3, 224, 224, device=device, dtype=dtype)
    y = torch.randint(0, 1000, (batch,), device=device)
    ddp = (lambda m: torch.nn.parallel.DistributedDataParallel(m, device_ids=[local_rank])) if world > 1 else (lambda m: m)
    plain = torchvision.models.mobilenet_v2().to(device).to(dtype).train()
    ms_plain = timed(eager(ddp(plain), torch.optim.SGD(plain.parameters(), lr=1e-3, momentum=0.9)))
    del plain
    torch.manual_seed(0)
    model = torchvision.models.mobilenet_v2().to(device).to(dtype)
    sim = QuantizationSimModel(model, dummy_input=x, quant_scheme="tf_enhanced", default_output_bw=8, default_param_bw=8)
    sim.compute_encodings(lambda m, _: m(x), None)
    sim.model.train()
    opt = torch.optim.SGD(sim.model.parameters(), lr=1e-3, momentum=0.9)
    before = ops.launches_total()
    ms_sim = timed(eager(ddp(sim.model), opt))
    launches = (ops.launches_total() - before) / (steps + warm)
    out = {"workload": "MobileNet-v2 QAT W8A8 (QDQ forward + STE backward), bf16, batch 32 per GPU (BASELINE configs[2])",
           "gpus": world, "plain_step_ms": round(ms_plain, 3), "quantsim_step_ms": round(ms_sim, 3),
           "quantsim_img_s": round(batch * world / ms_sim * 1e3, 1), "own_launches_per_step": round(launches, 1),
           "parallelism": "DistributedDataParallel" if world > 1 else "single process"}
    if graph and world == 1:
        graphed = sim.capture_train_step(loss_fn, opt, (x,), y)
        ms_graph = timed(lambda: graphed(x, target=y))
        out.update(quantsim_cuda_graph_step_ms=round(ms_graph, 3), quantsim_cuda_graph_img_s=round(batch / ms_graph * 1e3, 1))
    elif world > 1:
        out["note"] = "the step with DDP's all-reduces captured in one CUDA graph: tools/qat_ddp.py (profiles/r2_qat_ddp_n8.json)"
    del sim, model, opt
    torch.cuda.empty_cache()
    return out


def llama(device, layers=32):
    """configs[3]: the kernel-level half of tools/llama_w4.py (weights created on the device)."""
    import torch

    import bench
    from aimet_b200 import ops
    from aimet_b200.state import StateArena
    H, F = 4096, 11008
    peak, _ = bench.peak_hbm()
    g = torch.Generator(device=device).manual_seed(0)
    shapes = [(H, H)] * 4 + [(F, H)] * 2 + [(H, F)]
    weights = [(torch.randn(s, device=device, generator=g) * 0.02).to(torch.bfloat16) for _ in range(layers) for s in shapes]
    n_weights = sum(w.numel() for w in weights)
    channels = sum(w.shape[0] for w in weights)
    wbytes = 2 * n_weights
    arena = StateArena.for_device(device)
    per_block = len(shapes)
    blk_channels = sum(s[0] for s in shapes)
    blk = arena.allocate(blk_channels)
    enc = torch.empty((blk_channels, 5), dtype=torch.float64, device=device)
    params = torch.empty(4 * blk_channels, dtype=torch.float32, device=device)

    def refresh(mode):
        for b in range(layers):
            ws = weights[b * per_block:(b + 1) * per_block]
            ops.stats_refresh_multi_impl(ws, [w.shape[0] for w in ws], blk.arena, blk.first, mode, 4, True, False, False, enc,
                                         None, params)

    out = {"workload": "Llama-2-7B-shaped bf16 weights, W4 per-channel symmetric + tf activation encodings (BASELINE configs[3])",
           "matrices": len(weights), "weights": n_weights, "channels": channels}
    for name, mode in (("tf", ops.QUANTIZATION_TF), ("tf_enhanced", ops.QUANTIZATION_TF_ENHANCED)):
        ms = _events(lambda mode=mode: refresh(mode), 2, 1)
        row = {"ms": round(ms, 2), "channels_per_s": round(channels / ms * 1e3)}
        if name == "tf":      # one pass over the weights: HBM-bound; the tf_enhanced grid search is FP64-bound instead
            row.update(read_gbs=round(wbytes / ms / 1e6, 1), frac=round(wbytes / ms / 1e6 / peak, 3))
        else:
            row["bound"] = "FP64 grid search (101 candidates x 512 bins per channel), not HBM"
        out[f"weight_encodings_{name}"] = row
    blocks = []
    for b in range(layers):                      # the parameter blocks of every matrix, then QDQ over all of them
        ws = weights[b * per_block:(b + 1) * per_block]
        p = torch.empty(4 * blk_channels, dtype=torch.float32, device=device)
        ops.stats_refresh_multi_impl(ws, [w.shape[0] for w in ws], blk.arena, blk.first, ops.QUANTIZATION_TF, 4, True, False,
                                     False, enc, None, p)
        at = 0
        for w in ws:
            blocks.append(p[4 * at:4 * (at + w.shape[0])])
            at += w.shape[0]

    def qdq_all():
        for w, p in zip(weights, blocks):
            ops.qdq_per_channel_impl(w, p, w.shape[0], w.shape[1], 0, 0)

    ms = _events(qdq_all, 2, 1)
    out["weight_qdq_w4_per_channel"] = {"ms": round(ms, 2), "gbs": round(2 * wbytes / ms / 1e6, 1),
                                        "frac": round(2 * wbytes / ms / 1e6 / peak, 3), "launches": len(weights)}
    st = arena.allocate(2)
    for k, (name, cols) in enumerate((("[8,2048,4096]", H), ("[8,2048,11008]", F))):
        a = (torch.randn(8, 2048, cols, device=device, generator=g) * 0.5).to(torch.bfloat16)
        nbytes = a.numel() * 2
        ms_s = _events(lambda a=a, k=k: ops.stats_update_impl(a, st.arena, st.first + k, ops.QUANTIZATION_TF, None, 0), 5, 2)
        ms_q = _events(lambda a=a: ops.qdq_per_tensor_impl(a, -3.0, 3.0, 16, 0, 0), 5, 2)
        out[f"activation_{name}"] = {"tf_stats_frac": round(nbytes / ms_s / 1e6 / peak, 3),
                                     "qdq_a16_frac": round(2 * nbytes / ms_q / 1e6 / peak, 3)}
        del a
    del weights, blocks
    torch.cuda.empty_cache()
    return out


def main():
    import torch
    which = sys.argv[1:] or ["resnet18", "qat", "llama"]
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    out = {}
    if "resnet18" in which:
        out["resnet18_w8a8"] = resnet18(dev)
    if "qat" in which:
        out["mobilenet_v2_qat"] = qat(dev)
    if "llama" in which:
        out["llama7b_w4a16"] = llama(dev)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
