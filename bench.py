#!/usr/bin/env python
"""Headline benchmark (BASELINE.json): ResNet-50 QuantSim calibration img/s, plus the HBM roofline of the dominant kernel.

  python bench.py --gpus N --steps K --warmup W            # this repo: sm_100a kernels behind the reference's API
  python bench.py --impl reference --steps K --warmup W    # the reference's own CPU implementation on the host cores

Workload (BASELINE.json configs[1]): torchvision ResNet-50, random init (seed 0), W8A8, per-channel weights
(default_config_per_channel), quant scheme tf_enhanced; one STEP = one calibration batch of 32 synthetic 3x224x224
images through the quantsim model (fp32 forward with every wrapper collecting statistics). The timed region is a
COMPLETE calibration job of K steps: reset -> K batches (the first one also derives the 26 560 per-channel weight
encodings) -> (N > 1: NCCL merge of the per-rank statistics) -> grid search for every quantizer -> encodings on the host.
value = images calibrated by all ranks / that time. N > 1 is weak scaling: every rank runs K batches.

One JSON line on stdout (rank 0). See README / DESIGN.md for the key meanings.
"""
import argparse
import gc
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

BATCH = 32
IMAGE = (3, 224, 224)
CPU_BATCH = 4          # bounded sample for the CPU arms: images per step ...
CPU_IMAGE_BUDGET = 256  # ... shrunk when K steps of it would exceed this many images (~2 minutes of host time)


def cpu_batch_for(steps):
    return max(1, min(CPU_BATCH, CPU_IMAGE_BUDGET // max(steps, 1)))
METRIC = "resnet50_quantsim_calibration_throughput"
UNIT = "img/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=32)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--eager", action="store_true", help="do not replay the steady state from a CUDA graph")
    ap.add_argument("--cpu-baseline-steps", type=int, default=2)
    ap.add_argument("--images-per-step", type=int, default=BATCH,
                    help="calibration batch (SURVEY section 8d fixes 32 for the headline; larger batches mean larger "
                         "tensors per statistics launch -- a sensitivity knob, not the headline config)")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------------------------------
# shared pieces
# ---------------------------------------------------------------------------------------------------------------------
def build_sim(device, op_factory=None):
    import torch
    import torchvision

    from aimet_b200.quantsim import QuantizationSimModel, tensor_quantizer
    from aimet_b200.quantsim import config as qconfig
    prev = tensor_quantizer._set_op_class_for_testing(op_factory) if op_factory is not None else None
    try:
        torch.manual_seed(0)
        model = torchvision.models.resnet50().eval().to(device)
        dummy = torch.zeros((2,) + IMAGE, device=device)
        sim = QuantizationSimModel(model, dummy_input=dummy, quant_scheme="tf_enhanced", default_output_bw=8,
                                   default_param_bw=8, config_file=qconfig.DEFAULT_CONFIG_PER_CHANNEL, in_place=True)
    finally:
        if prev is not None:
            tensor_quantizer._set_op_class_for_testing(prev)
    return sim


def synthetic_batch(global_index, batch, device="cpu", pin=False):
    import torch
    g = torch.Generator().manual_seed(1000 + global_index)
    x = torch.randn((batch,) + IMAGE, generator=g)
    if pin:
        x = x.pin_memory()
    return x.to(device) if device != "cpu" else x


class ClockSampler:
    """nvidia-smi in the background during the timed region (B200_PROFILING.md clocks line)."""
    FIELDS = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu_index = gpu_index
        self.proc = None
        self.path = None
        self.window = None      # (t0, t1) wall-clock bounds of the timed region: only samples inside it count

    @staticmethod
    def _stamp(text):
        """'2026/10/18 21:04:05.123' -> seconds since the epoch (local time, like time.time())"""
        import datetime
        try:
            return datetime.datetime.strptime(text, "%Y/%m/%d %H:%M:%S.%f").timestamp()
        except ValueError:
            return None          # unknown format: the sample is kept (see stop)

    def start(self):
        try:
            f = tempfile.NamedTemporaryFile("w", suffix=".csv", delete=False)
            self.path = f.name
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu_index)], stdout=f, stderr=subprocess.DEVNULL)
        except Exception:   # pylint: disable=broad-except
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:   # pylint: disable=broad-except
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        try:
            rows = []
            for line in open(self.path):
                parts = [p.strip() for p in line.split(",")]
                if len(parts) < 9:
                    continue
                try:
                    rows.append((self._stamp(parts[0]), float(parts[1]), float(parts[2]), parts))
                except ValueError:
                    continue
            inside = [r for r in rows if r[0] is None or
                      (self.window and self.window[0] - 0.05 <= r[0] <= self.window[1] + 0.05)]
            for _, a, b, parts in (inside or rows):       # a region shorter than the sampling period: all samples
                sm.append(a)
                mx.append(b)
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"),
                                     parts[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:   # pylint: disable=broad-except
            pass
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


def peak_hbm():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:   # pylint: disable=broad-except
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic_ratio():
    """dram bytes / algorithmic bytes of the histogram kernel from the committed ncu --set full capture."""
    p = os.path.join(ROOT, "profiles", "hist_kernel_traffic.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return float(d["dram_bytes_per_launch"]) / float(d["algorithmic_bytes_per_launch"])
        except Exception:   # pylint: disable=broad-except
            pass
    return None


# ---------------------------------------------------------------------------------------------------------------------
# CPU arms (the reference's own implementation on the host cores)
# ---------------------------------------------------------------------------------------------------------------------
def cpu_job(steps, warmup):
    """A complete calibration job of `steps` batches of cpu_batch_for(steps) images on the host: torch CPU forward (all
    threads) and the reference's C++ statistics / encodings (single-threaded, as the reference is). Returns (img/s, info)."""
    import torch
    batch = cpu_batch_for(steps)

    from aimet_b200.quantsim import tensor_quantizer
    from oracle import cpu_backend
    factory = cpu_backend.best_cpu_backend()
    sim = build_sim("cpu", factory)
    prev = tensor_quantizer._set_op_class_for_testing(factory)
    try:
        batches = [synthetic_batch(b, batch) for b in range(max(steps, 1))]
        if warmup > 0:
            # warm-up: forward passes only touch the allocator / thread pool; one tiny complete job primes everything
            sim.compute_encodings(lambda m, _: [m(batches[i % len(batches)][:1]) for i in range(1)], None)
        t0 = time.perf_counter()
        sim.compute_encodings(lambda m, _: [m(batches[i]) for i in range(steps)], None)
        act, par = sim.get_activation_param_encodings()
        dt = time.perf_counter() - t0
    finally:
        tensor_quantizer._set_op_class_for_testing(prev)
    info = {"kind": factory.KIND, "cores": torch.get_num_threads(),
            "images_per_step": batch,
            "sample": f"complete calibration job of {steps} steps x {batch} images (ResNet-50 per-channel "
                      f"tf_enhanced; includes the 26 560 weight-channel encodings and the final grid search); "
                      f"torch CPU forward on {torch.get_num_threads()} threads, reference statistics single-threaded",
            "num_activation_encodings": len(act), "num_param_encodings": len(par), "seconds": round(dt, 3)}
    return steps * batch / dt, info


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    value, info = cpu_job(args.steps, args.warmup)
    line = {"impl": "reference", "metric": METRIC, "value": round(value, 3), "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(1000.0 * info["images_per_step"] / value, 3),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "ResNet-50 W8A8 per-channel weights, tf_enhanced calibration (BASELINE configs[1])",
                       "images_per_step": info["images_per_step"],
                       "note": "host CPU only; bounded sample of the same workload"},
            "cpu_baseline": dict(info, value=round(value, 3), unit=UNIT),
            "e2e": {"value": round(value, 3), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


# ---------------------------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: aimet_b200 has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)
    torch.backends.cudnn.benchmark = os.environ.get("BENCH_CUDNN_BENCHMARK", "1") == "1"
    torch.backends.cudnn.allow_tf32 = False          # the reference's forward is plain fp32
    torch.backends.cuda.matmul.allow_tf32 = False

    import aimet_b200  # noqa: F401
    from aimet_b200 import ops
    from aimet_b200.distributed import ShardedCalibrator
    from aimet_b200.utils import DevicePrefetcher

    sim = build_sim(device)
    steps, warmup = args.steps, args.warmup
    # global batch index of this rank's i-th step: i * world + rank
    dev_batches = [synthetic_batch(i * world + rank, BATCH, device) for i in range(steps)]
    host_batches = [synthetic_batch(i * world + rank, BATCH, "cpu", pin=True) for i in range(steps)]

    # Replaying the steady-state step from a CUDA graph is implemented and verified (tests/test_gpu_quantsim.py), but for
    # ONE job of a few dozen steps capturing the ~700-kernel step costs more than the host time it saves, so the default
    # measures the plain eager path; BENCH_CUDA_GRAPH=1 switches it on.
    use_graph = os.environ.get("BENCH_CUDA_GRAPH", "0") == "1" and not args.eager

    def job(batches_of, n, graph=None):
        """One complete calibration job over n batches through the public API. `batches_of(n)` returns the iterable of
        this rank's n batches."""
        graph = use_graph if graph is None else graph
        if not graph:
            def cb(model, _):
                for x in batches_of(n):
                    model(x)
            if world > 1:
                ShardedCalibrator(sim).compute_encodings(cb, None)
            else:
                sim.compute_encodings(cb, None)
        else:
            if world > 1:
                ShardedCalibrator(sim).compute_encodings_for_batches(batches_of(n), cuda_graph=True)
            else:
                sim.compute_encodings_for_batches(batches_of(n), cuda_graph=True)
        return sim.get_activation_param_encodings()

    def resident(n):
        return (dev_batches[i % steps] for i in range(n))

    def from_host(n):
        # pinned host batches, copied to the device step by step INSIDE the timed region; the copy of the next batch is
        # enqueued on a second stream while the current one is being processed (aimet_b200.utils.DevicePrefetcher)
        return DevicePrefetcher(host_batches[:n], device)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up: W steps of a complete job (cuDNN autotune, allocator, lazy module loading) ----
    if warmup > 0:
        job(resident, warmup)
        job(from_host, warmup)      # the host-fed path too: copy stream, staging buffers, first async H2D of the process
    barrier()
    # Everything built so far (model, wrappers, torch internals: ~1e6 Python objects) is long-lived: move it out of the
    # collector's sight so that a generation-2 pass triggered by the 26 560 encodings a job exports does not walk it in
    # the middle of a timed job (measured: +77 ms on every other 32-step job without this).
    gc.collect()
    gc.freeze()

    # ---- timed: value (inputs resident in HBM) ----
    gc.collect()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        # nvidia-smi takes a while to start and takes driver locks while it does (measured: +8 ms on an 8-step job when it
        # starts together with the timed region): start it first, let it settle, and count the samples inside the region
        sampler.start()
        time.sleep(0.5)
    launches0 = dict(ops.LAUNCHES)
    barrier()
    t_region0 = time.time()
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.nvtx.range_push("timed")          # lets `ncu --nvtx --nvtx-include "timed/"` see only the timed region
    start.record()
    act, par = job(resident, steps)
    stop.record()
    torch.cuda.nvtx.range_pop()
    barrier()
    ms = start.elapsed_time(stop)
    from aimet_b200.quantsim import quantsim as _qs
    graph_info = dict(_qs.LAST_GRAPH_INFO)
    eager_batches = min(steps, 2) if use_graph else steps
    launched = {k: ops.LAUNCHES[k] - launches0[k] for k in ops.LAUNCHES}
    sampler.window = (t_region0, t_region0 + ms / 1000.0 + 0.01)
    clocks = sampler.stop() if rank == 0 else None

    # ---- timed: e2e (host buffers; H2D of every batch and D2H of the result inside the region) ----
    gc.collect()
    barrier()
    t0 = time.perf_counter()
    act, par = job(from_host, steps)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    barrier()

    # ---- the same job once more with CUDA events around every statistics call (roofline of the dominant kernel) ----
    # The eager step is host-bound, so an event pair around a launch would also time the GPU waiting for the host to
    # enqueue it. This repeat therefore replays the steady-state step from a CUDA graph in which the event pairs are
    # external event-record nodes: what is read afterwards are DEVICE-side durations of the statistics launches of the
    # last replayed step -- same kernels, same tensors, same order as the timed job.
    ops.reserve_timing_events(2 * 100 * steps + 64)
    ops.STATS_TIMING = []
    per_step0 = dict(ops.LAUNCHES)
    job(resident, steps, graph=steps > 2)
    barrier()
    timing, ops.STATS_TIMING = ops.STATS_TIMING, None
    # launches of one captured (replayed) step = launches issued while capturing = total of this job minus the eager ones
    job_launches = {k: ops.LAUNCHES[k] - per_step0[k] for k in ops.LAUNCHES}

    # ---- and a short eager pass with the kernel's own clock (ab_debug_hist_timer): first CTA start to last CTA end of every
    # histogram launch on the GPU's global timer. No launch latency, no event records between the kernels: what the
    # histogram itself takes in its real surroundings (input just written by the producing layer, L2 in whatever state).
    from aimet_b200 import _lib as ab_lib
    dev_timer = None
    if True:   # every rank runs it (the sharded job has collectives); only rank 0's numbers are printed
        probe_steps = min(steps, 4)
        cap = 128 * probe_steps + 64
        slots = torch.zeros((cap, 3), dtype=torch.int64, device=device)
        slots[:, 0] = torch.iinfo(torch.int64).max
        ab_lib.load().ab_debug_hist_timer(slots.data_ptr(), cap)
        job(resident, probe_steps)
        torch.cuda.synchronize()
        used = int(ab_lib.load().ab_debug_hist_timer(None, 0))
        if 0 < used <= cap:
            rows = slots[:used].cpu().tolist()
            # the activation statistics of the LAST step: the trailing launches, as many as one steady-state step makes
            per_step = n_act_hist = sum(1 for r in rows if r[2] >= 64 * 1024) // probe_steps
            last = [r for r in rows if r[2] >= 64 * 1024][-per_step:] if per_step else []
            if last:
                b, t_ns = sum(r[2] for r in last), sum(r[1] - r[0] for r in last)
                big = [r for r in last if r[2] >= 32 * 2**20]
                dev_timer = {"launches": len(last), "avg_launch_us": round(t_ns / 1000.0 / len(last), 2),
                             "achieved": round(b / t_ns, 1), "algorithmic_bytes_per_launch": round(b / len(last), 1),
                             "achieved_large_tensors":
                                 round(sum(r[2] for r in big) / sum(r[1] - r[0] for r in big), 1) if big else None}
    barrier()

    # max over ranks
    t = torch.tensor([ms, e2e_s * 1000.0], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, e2e_ms = t.tolist()

    # ---- roofline of the dominant kernel: tf_enhanced statistics (histogram) launches of the timed region ----
    tot_bytes = tot_ms = 0.0
    big_bytes = big_ms = 0.0
    n_l = 0
    steady = [t for t in timing if t[4]] if any(t[4] for t in timing) else timing
    for nbytes, e0, e1, mode, _captured in steady:
        if mode != ops.QUANTIZATION_TF_ENHANCED:
            continue
        d = e0.elapsed_time(e1)
        tot_bytes += nbytes
        tot_ms += d
        n_l += 1
        if nbytes >= 32 * 2**20:
            big_bytes += nbytes
            big_ms += d
    peak, peak_src = peak_hbm()
    achieved = tot_bytes / tot_ms / 1e6 if tot_ms > 0 else 0.0
    ratio = ncu_traffic_ratio()
    roofline = {"bound": "hbm", "kernel": "hist_kernel<float>",
                "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4),
                "traffic": round(ratio * tot_bytes / max(n_l, 1), 1) if ratio else None,
                "algorithmic_bytes_per_launch": round(tot_bytes / max(n_l, 1), 1), "launches": n_l,
                "avg_launch_us": round(1000.0 * tot_ms / max(n_l, 1), 2), "peak_source": peak_src,
                "achieved_large_tensors": round(big_bytes / big_ms / 1e6, 1) if big_ms > 0 else None,
                "device_timer": (dict(dev_timer, frac=round(dev_timer["achieved"] / peak, 4), unit="GB/s",
                                      how="first CTA start to last CTA end on %globaltimer, last step of a short "
                                          "eager repeat (no launch latency, no event records between kernels)")
                                 if dev_timer else None),
                "note": "4 B/element x elements of every activation tensor handed to updateStats, divided by the "
                        "CUDA-event time of those launches (events on the launching stream; in CUDA-graph mode: "
                        "external event nodes inside the replayed step, read for the last step of an instrumented "
                        "repeat of the timed job); achieved_large_tensors restricts to tensors >= 32 MB"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # kernels of ours that ran in the timed region: launches issued from Python (a captured launch counts once, for the
    # replay that follows the capture) + the captured step's launches for every further replay of the graph
    gpu_launches = sum(launched.values()) + graph_info["captured_launches"] * max(graph_info["replays"] - 1, 0)
    images = BATCH * steps * world
    value = images / (ms / 1000.0)
    e2e = images / (e2e_ms / 1000.0)
    enc_bytes = (sum(len(v.get("input", {})) + len(v.get("output", {})) for v in act.values()) +
                 sum(len(v) for v in par.values())) * 5 * 8
    line = {"metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
            "ms_per_step": round(ms / steps, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "ResNet-50 W8A8 per-channel weights, tf_enhanced calibration (BASELINE configs[1])",
                       "images_per_step": BATCH, "image": list(IMAGE), "global_images": images,
                       "timed_region": "complete job: reset, K batches (incl. per-channel weight encodings), "
                                       "merge (N>1), grid search, encodings on host",
                       "l2": "activation working set per step (2.2 GB) exceeds the 126 MB L2; no flush needed",
                       "num_activation_encodings": len(act), "num_param_tensors": len(par),
                       "parallelism": f"batch-sharded x{world}"},
            "e2e": {"value": round(e2e, 2), "unit": UNIT, "h2d_bytes_per_step": BATCH * 3 * 224 * 224 * 4,
                    "d2h_bytes_per_step": int(enc_bytes / steps),
                    "h2d": "every step's batch is copied from pinned host memory inside the timed region, on a copy "
                           "stream one batch ahead of the compute stream (aimet_b200.utils.DevicePrefetcher)"},
            "gpu_launches": gpu_launches, "launches_by_kernel": launched,
            "cuda_graph": dict(graph_info, enabled=bool(use_graph)),
            "roofline": roofline, "clocks": clocks}

    if world == 1 and not args.no_cpu_baseline:
        try:
            v, info = cpu_job(args.cpu_baseline_steps, 1)
            line["cpu_baseline"] = dict(info, value=round(v, 3), unit=UNIT)
        except Exception as exc:   # pylint: disable=broad-except
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "error": str(exc)[:200]}
    emit(line)
    if world > 1:
        dist.destroy_process_group()


_REAL_STDOUT = None


def emit(line: dict):
    """The ONE JSON line goes to the real stdout; everything else this process (or NCCL, cuDNN ...) prints went to stderr."""
    data = (json.dumps(line) + "\n").encode()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, data)


def main():
    global _REAL_STDOUT, BATCH
    args = parse_args()
    BATCH = args.images_per_step
    # C libraries print to fd 1 too (e.g. "NCCL version ..." on the first collective): keep stdout clean for the JSON line
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
