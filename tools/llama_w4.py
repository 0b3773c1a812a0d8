"""BASELINE config 4: a Llama-2-7B-shaped, randomly initialised stack of Linear layers (32 blocks x {4 x [4096, 4096],
2 x [11008, 4096], 1 x [4096, 11008]} = 6.48 G weights, bf16; no attention arithmetic -- the metric is about the quantization
path), W4 per-channel (axis 0) symmetric weight quantize-dequantize + tf activation encodings at 16 bit on activations of
shape [8, 2048, 4096] / [8, 2048, 11008].

Reports (one JSON): the whole job through the public API (QuantizationSimModel with a per-channel config that includes
Linear layers, quant_scheme tf: compute_encodings over one batch + one quantized forward), and the path's own kernels on the
same tensors: per-channel statistics (tf and tf_enhanced) and the grid search for all 1.36 M channels, per-channel W4 QDQ over
all 224 matrices, tf statistics and 16-bit QDQ on the activations -- seconds and GB/s against the measured HBM peak.

    python tools/llama_w4.py [--layers 32]
"""
import argparse
import json
import os
import sys
import time

import torch
from torch import nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from aimet_b200 import ops  # noqa: E402
from aimet_b200.quantsim import QuantizationSimModel  # noqa: E402
from aimet_b200.quantsim import config as qconfig  # noqa: E402
from aimet_b200.state import StateArena  # noqa: E402

H, F = 4096, 11008


class Block(nn.Module):
    def __init__(self):
        super().__init__()
        self.q, self.k, self.v, self.o = (nn.Linear(H, H, bias=False) for _ in range(4))
        self.gate, self.up, self.down = nn.Linear(H, F, bias=False), nn.Linear(H, F, bias=False), nn.Linear(F, H, bias=False)

    def forward(self, x):
        h = self.o(self.v(self.k(self.q(x))))
        return self.down(self.gate(h) * self.up(h))


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e30
    for _ in range(reps):
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best / 1e3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--layers", type=int, default=32)
    ap.add_argument("--no-sim", action="store_true")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    peak, _ = bench.peak_hbm()
    torch.manual_seed(0)
    model = nn.Sequential(*[Block() for _ in range(args.layers)]).to(torch.bfloat16)
    for p in model.parameters():
        nn.init.normal_(p, std=0.02)
    model = model.to(dev).eval()
    weights = [p.data for p in model.parameters()]
    n_weights = sum(w.numel() for w in weights)
    channels = sum(w.shape[0] for w in weights)
    x = (torch.randn(8, 2048, H, device=dev) * 0.5).to(torch.bfloat16)
    out = {"layers": args.layers, "matrices": len(weights), "weights": n_weights, "channels": channels, "peak_gbs": peak,
           "dtype": "bf16"}

    # ---- the path's kernels on these tensors -----------------------------------------------------------------------
    arena = StateArena.for_device(dev)
    per_block = 7
    blk_channels = sum(w.shape[0] for w in weights[:per_block])
    blk = arena.allocate(blk_channels)
    enc = torch.empty((blk_channels, 5), dtype=torch.float64, device=dev)
    params = torch.empty(4 * blk_channels, dtype=torch.float32, device=dev)
    wbytes = 2 * n_weights

    def refresh(mode):
        for b in range(args.layers):
            ws = weights[b * per_block:(b + 1) * per_block]
            ops.stats_refresh_multi_impl(ws, [w.shape[0] for w in ws], blk.arena, blk.first, mode, 4, True, False, False,
                                         enc, None, params)

    for name, mode in (("tf", ops.QUANTIZATION_TF), ("tf_enhanced", ops.QUANTIZATION_TF_ENHANCED)):
        s = timed(lambda mode=mode: refresh(mode))
        out[f"weight_encodings_{name}"] = {"seconds": round(s, 4), "channels_per_s": round(channels / s),
                                           "read_gbs": round(wbytes / s / 1e9, 1),
                                           "note": "reset + per-channel statistics + encoding (grid search for tf_enhanced) + "
                                                   "per-channel parameter block, one native call per block of 7 matrices"}
    all_params = []
    for w in weights:                                       # per-matrix parameter blocks for the QDQ sweep (tf encodings)
        c = w.shape[0]
        b1 = arena.allocate(c)
        e1 = torch.empty((c, 5), dtype=torch.float64, device=dev)
        p1 = torch.empty(4 * c, dtype=torch.float32, device=dev)
        ops.stats_refresh_multi_impl([w], [c], b1.arena, b1.first, ops.QUANTIZATION_TF, 4, True, False, False, e1, None, p1)
        all_params.append(p1)
        del b1
    outs = [torch.empty_like(w) for w in weights[:per_block]]

    def qdq_all():
        for i, (w, p) in enumerate(zip(weights, all_params)):
            ops.qdq_per_channel_impl(w, p, w.shape[0], w.shape[1], 0, 0)

    s = timed(qdq_all)
    out["weight_qdq_w4_per_channel"] = {"seconds": round(s, 4), "gbs": round(2 * wbytes / s / 1e9, 1),
                                        "frac": round(2 * wbytes / s / 1e9 / peak, 3), "launches": len(weights)}
    del outs
    # a quantized weight really sits on a 16-level grid
    w0, p0 = weights[0], all_params[0]
    y0 = ops.qdq_per_channel_impl(w0, p0, w0.shape[0], w0.shape[1], 0, 0)
    c0 = w0.shape[0]
    levels = torch.unique(torch.round(y0[5].float() / p0[2 * c0 + 5]))
    out["w4_levels_in_one_channel"] = int(levels.numel())

    acts = {"[8,2048,4096]": x, "[8,2048,11008]": (torch.randn(8, 2048, F, device=dev) * 0.5).to(torch.bfloat16)}
    st = arena.allocate(2)
    for k, (name, a) in enumerate(acts.items()):
        nbytes = a.numel() * 2
        s_stat = timed(lambda a=a, k=k: ops.stats_update_impl(a, st.arena, st.first + k, ops.QUANTIZATION_TF, None, 0), 5)
        s_qdq = timed(lambda a=a: ops.qdq_per_tensor_impl(a, -3.0, 3.0, 16, 0, 0), 5)
        out[f"activation_{name}"] = {"tf_stats_gbs": round(nbytes / s_stat / 1e9, 1), "tf_stats_frac": round(nbytes / s_stat / 1e9 / peak, 3),
                                     "qdq_a16_gbs": round(2 * nbytes / s_qdq / 1e9, 1), "qdq_a16_frac": round(2 * nbytes / s_qdq / 1e9 / peak, 3)}
    del all_params, acts
    torch.cuda.empty_cache()

    # ---- the whole job through the public API ------------------------------------------------------------------------
    if not args.no_sim:
        cfg = json.loads(json.dumps(qconfig.DEFAULT_CONFIG_PER_CHANNEL))
        cfg["op_type"].pop("Gemm", None)                 # Linear layers per channel too (the stock file excludes Gemm)
        t0 = time.perf_counter()
        sim = QuantizationSimModel(model, dummy_input=x[:1, :8], quant_scheme="tf", default_output_bw=16, default_param_bw=4,
                                   config_file=cfg, in_place=True)
        torch.cuda.synchronize()
        t_build = time.perf_counter() - t0
        launches0 = ops.launches_total()

        def job():
            sim.compute_encodings(lambda m, _: m(x), None)

        job()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        job()
        torch.cuda.synchronize()
        t_job = time.perf_counter() - t0
        launches = ops.launches_total() - launches0
        with torch.no_grad():
            t_fwd = timed(lambda: sim.model(x), 3)
            plain = QuantizationSimModel.get_original_model(sim.model)
            t_plain = timed(lambda: plain(x), 3)
        n_param_q = sum(1 for _, w in sim.quant_wrappers() for q in w.param_quantizers.values() if q.enabled)
        n_act_q = len(sim.activation_quantizers())
        out["quantsim_api"] = {"build_seconds": round(t_build, 2), "compute_encodings_seconds": round(t_job, 4),
                               "own_launches_two_jobs": launches, "param_quantizers": n_param_q,
                               "activation_quantizers": n_act_q, "quantized_forward_seconds": round(t_fwd, 4),
                               "plain_forward_seconds": round(t_plain, 4), "tokens": 8 * 2048}
    print(json.dumps(out, indent=1))
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "llama_w4.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
