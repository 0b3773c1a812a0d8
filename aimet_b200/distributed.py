"""Calibration sharded by batch across GPUs, with encodings bit-identical to a single-process run.

Net-new relative to the reference (it has no collective anywhere: SURVEY.md sections 2.3, 8e). One process per GPU;
rank r of W runs global batches r, r+W, r+2W, ...; parameter encodings are computed redundantly (identical weights on
every rank); activation statistics are merged with collectives on `torch.distributed` (NCCL over NVLink on a B200 box,
gloo in the CPU tests):

  tf           the running (min, max) of every quantizer: all_reduce(MIN) / all_reduce(MAX) -- exact.
  tf_enhanced  the reference's result depends on (i) the histogram range fixed by the first non-zero batch in GLOBAL
               order and (ii) a sequential running mean over batches (DlQuantization/src/math_functions.cpp:248-287).
               (1) during its first local batch a rank only records each call's (min, max) and keeps a copy of the tensor;
                   one all_gather of the [Q, calls, 2] table lets every rank pick, per quantizer, the range the globally
                   first non-zero call defines, fix it on the device (ab_stats_init_range) and only then bin the kept
                   tensors -- nobody waits for rank 0's batch to finish. (Quantizers that have seen only zeros so far
                   repeat this round at the end of the next forward; ranks that have run out of batches join the rounds
                   from their merge, so every rank issues the same sequence of collectives.)
               (2) from then on the calls go through the deferred multi-tensor statistics of quantsim.stats_batcher in
                   LOG mode: the raw integer counts of every call that happens land in one row of a device log -- one
                   histogram launch per forward, no per-batch staging copies, no padding for calls that do not happen;
               (3) one all_reduce(MAX) of the row count, ONE all_gather of the logs (their (batch, quantizer) tags ride in
                   the same buffer), then ONE kernel replays pdf = (pdf*k + hist/cnt)/(k+1) in global batch order for
                   all quantizers (ab_stats_fold_log). Integer counts make the merge exact.

The pure-tensor helpers (`choose_first_ranges`, `replay_plan`, ...) run on any device and are what the gloo tests exercise;
the device work goes through aimet_b200.ops.
"""
from typing import List, Optional

import torch
import torch.distributed as dist

from . import ops
from .quantsim.defs import QuantScheme
from .state import StateArena, field_index

LOG_WORDS = ops.LOG_WORDS


# ---------------------------------------------------------------------------------------------------------------------
# pure helpers
# ---------------------------------------------------------------------------------------------------------------------
def choose_first_ranges(gathered: torch.Tensor) -> torch.Tensor:
    """gathered: [W, Q, C, 2] -- (min, max) of call c of each rank's FIRST local batch (global batch r), +inf / -inf
    where the call did not happen. Returns [Q, 2]: for each quantizer the (min, max) of the first call, in global order
    (rank-major, then call), that is not all-zero -- the one the reference's UpdatePdf would initialise from
    (math_functions.cpp:248-262) -- or (0, 0) if there is none."""
    w, q, c, _ = gathered.shape
    table = gathered.permute(1, 0, 2, 3).reshape(q, w * c, 2)
    happened = torch.isfinite(table[..., 0]) & torch.isfinite(table[..., 1])
    nonzero = happened & ((table[..., 0] != 0) | (table[..., 1] != 0))
    has = nonzero.any(dim=1)
    first = torch.argmax(nonzero.to(torch.int8), dim=1)
    picked = table[torch.arange(q, device=table.device), first]
    return torch.where(has[:, None], picked, torch.zeros_like(picked)).contiguous()


def first_call_positions(gathered: torch.Tensor) -> torch.Tensor:
    """[Q] index (rank * C + call) of the call `choose_first_ranges` picked, or W*C where none qualifies."""
    w, q, c, _ = gathered.shape
    table = gathered.permute(1, 0, 2, 3).reshape(q, w * c, 2)
    happened = torch.isfinite(table[..., 0]) & torch.isfinite(table[..., 1])
    nonzero = happened & ((table[..., 0] != 0) | (table[..., 1] != 0))
    first = torch.argmax(nonzero.to(torch.int8), dim=1)
    return torch.where(nonzero.any(dim=1), first, torch.full_like(first, w * c))


def global_replay_offsets(world: int, local_batches: int, calls: int, num_quantizers: int) -> torch.Tensor:
    """Word offsets into the gathered log [W, local_batches * calls, Q, LOG_WORDS] in the order a single process would
    have produced the updateStats calls: global batch b = i * W + r (i-th local batch of rank r), then call index."""
    i = torch.arange(local_batches).view(-1, 1, 1)
    r = torch.arange(world).view(1, -1, 1)
    c = torch.arange(calls).view(1, 1, -1)
    slot = r * (local_batches * calls) + i * calls + c
    return (slot.reshape(-1) * (num_quantizers * LOG_WORDS)).to(torch.int64)


def replay_plan(meta: torch.Tensor, num_records: int):
    """meta: int [W, R, 2] -- (local batch, record) of row r of rank w's log, record -1 for padding rows. Returns
    (entry_rows int64 [E], record_begin int64 [num_records + 1]): for every record the rows (w * R + r) of the gathered log
    it replays, in the order a single process would have made the calls: global batch = local batch * W + w, then call
    order inside the batch (= row order)."""
    w, r, _ = meta.shape
    meta = meta.to(torch.int64).cpu()
    rank = torch.arange(w).view(w, 1).expand(w, r)
    row = torch.arange(r).view(1, r).expand(w, r)
    record = meta[..., 1]
    keep = record >= 0
    global_batch = meta[..., 0] * w + rank
    # lexicographic (record, global batch, row): rows of one rank and batch are already in call order
    key = (record[keep] * (int(global_batch.max()) + 2 if keep.any() else 1) + global_batch[keep]) * r + row[keep]
    order = torch.argsort(key)
    entry_rows = (rank[keep] * r + row[keep])[order].contiguous()
    counts = torch.bincount(record[keep], minlength=num_records)
    record_begin = torch.zeros(num_records + 1, dtype=torch.int64)
    record_begin[1:] = torch.cumsum(counts, 0)
    return entry_rows, record_begin


def _all_gather(t: torch.Tensor, group) -> torch.Tensor:
    """[W, *t.shape]. NCCL gathers device tensors in place over NVLink; with the gloo backend (CPU tests, or several
    ranks sharing one GPU) the payload is staged through host memory."""
    world = dist.get_world_size(group)
    if t.is_cuda and dist.get_backend(group) == "nccl":
        out = torch.empty((world,) + tuple(t.shape), dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(out, t.contiguous(), group=group)
        return out
    host = t.detach().cpu().contiguous()
    out = torch.empty((world,) + tuple(host.shape), dtype=host.dtype)
    dist.all_gather(list(out.unbind(0)), host, group=group)
    return out.to(t.device)


_WARMED = set()


def warm_collectives(group, device):
    """Once per (process group, device): run the gather this module uses at a small, a medium and a large payload, so
    that NCCL's lazily established protocols / channels (LL, LL128, Simple pick themselves by message size) exist before
    a job needs them. Without it the first job whose log is larger than anything gathered before pays tens of
    milliseconds of connection setup inside `_merge`."""
    key = (id(group) if group is not None else None, str(device))
    if key in _WARMED or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    _WARMED.add(key)
    if device.type != "cuda" or dist.get_backend(group) != "nccl":
        return
    for words in (8 * 1024, 512 * 1024, 16 * 1024 * 1024):
        _all_gather(torch.zeros(words, dtype=torch.int32, device=device), group)
    t = torch.zeros(4, dtype=torch.float64, device=device)
    _all_reduce(t, dist.ReduceOp.MAX, group)
    torch.cuda.current_stream(device).synchronize()


def _all_reduce(t: torch.Tensor, op, group):
    if t.is_cuda and dist.get_backend(group) != "nccl":
        host = t.detach().cpu()
        dist.all_reduce(host, op=op, group=group)
        t.copy_(host)
    else:
        dist.all_reduce(t, op=op, group=group)


# ---------------------------------------------------------------------------------------------------------------------
# the calibrator
# ---------------------------------------------------------------------------------------------------------------------
class ShardedCalibrator:
    """`ShardedCalibrator(sim).compute_encodings(cb, args)` is QuantizationSimModel.compute_encodings for a job whose
    batches are dealt round-robin to the ranks of `group`; `cb(model, args)` runs THIS rank's batches, one model forward
    per batch. Every rank ends up with the same encodings as a single process calibrating on all batches in order."""

    def __init__(self, sim, group=None, max_calls_per_batch: int = 4):
        self.sim = sim
        self.group = group
        self.max_calls = max_calls_per_batch
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0

    # -- set-up ----------------------------------------------------------------------------------------------------
    def _attach(self, staging: bool):
        from .quantsim.stats_batcher import LogSink, StatsBatcher
        sim = self.sim
        self.quantizers = list(sim._act_block_quantizers)   # pylint: disable=protected-access
        self.block = sim._act_block                         # pylint: disable=protected-access
        self.device = self.block.device
        warm_collectives(self.group, self.device)
        q_count = len(self.quantizers)
        self.tfe = sim._quant_scheme != QuantScheme.post_training_tf   # pylint: disable=protected-access
        self.batcher = None
        if not self.tfe:
            return            # tf: the ordinary per-call min/max path, merged with two all_reduces
        self.sink = LogSink(self.device, capacity=max(4096, 64 * q_count),
                            staging_rows=2 * q_count * self.max_calls if staging else 0)
        self.batcher = StatsBatcher.attach(sim, sink=self.sink)
        if self.batcher is None:
            raise RuntimeError("sharded tf_enhanced calibration needs the activation statistics on a CUDA device")
        self.batcher.collect = self._collect
        self.batcher.range_exchange = self._forward_end
        self.range_block = StateArena.for_device(self.device).allocate(q_count * self.max_calls)
        self.calls = [0] * q_count
        self.held = []                      # (record, call, private copy) of this forward's calls on records without a range
        self.called_anywhere = [False] * q_count
        self.need_round = True              # round 0 always happens; later ones while called records have no range yet
        self.rounds = 0
        self.local_fallback = False

    def _detach(self):
        if self.batcher is not None:
            self.batcher.detach()
            self.batcher = None
        self.held = []
        self.range_block = None

    # -- exchange 1: ranges ----------------------------------------------------------------------------------------
    def _collect(self, i: int, tensor: torch.Tensor):
        """A call on a record whose range is not known yet: note its (min, max), keep a private copy for binning later."""
        if not self.need_round:
            # first call ever on this record, after the range rounds are over (the forward's control flow differs between
            # batches): no collective can be started unilaterally -- the range is fixed from this rank's data
            self.local_fallback = True
            self.batcher.update_now(i, tensor)
            self.batcher.fixed[i] = True
            return
        call = self.calls[i]
        self.calls[i] += 1
        if call >= self.max_calls:
            raise RuntimeError(f"a quantizer was updated more than max_calls_per_batch={self.max_calls} times in one "
                               "forward pass; raise the limit")
        ops.stats_update_impl(tensor, self.range_block.arena, self.range_block.first + i * self.max_calls + call,
                              ops.QUANTIZATION_TF, None, 0)
        self.held.append((i, call, tensor.detach().clone()))

    def _forward_end(self, batcher):
        if self.need_round:
            self._range_round(done=False)

    def _range_round(self, done: bool):
        """One round of exchange 1 (identical on every rank): gather the (min, max) tables of the forward that just ended,
        fix the ranges their globally first non-zero calls define, bin the kept tensors."""
        q_count, c = len(self.quantizers), self.max_calls
        rec = self.range_block.bytes_view().view(torch.float64).view(q_count * c, -1)
        base = field_index("run_min", 8)          # run_min, run_max are adjacent doubles (include/aimet_b200.h)
        # an un-updated record holds (+DBL_MAX, -DBL_MAX) -> (+inf, -inf) in float32: "call did not happen"
        table = rec[:, base:base + 2].to(torch.float32).reshape(-1)
        payload = torch.cat([table, torch.full((1,), 1.0 if done else 0.0, device=self.device)])
        gathered = _all_gather(payload, self.group) if self.world > 1 else payload.unsqueeze(0)
        tables = gathered[:, :-1].reshape(self.world, q_count, c, 2)
        chosen = choose_first_ranges(tables)
        ops.stats_init_range_impl(self.block.arena, self.block.first, q_count, chosen)
        happened = torch.isfinite(tables[..., 0]).any(dim=2).any(dim=0)                        # [Q] called on some rank
        host = torch.cat([((chosen[:, 0] != 0) | (chosen[:, 1] != 0)).to(torch.float32), happened.to(torch.float32),
                          first_call_positions(tables).to(torch.float32), gathered[:, -1]]).tolist()   # one read-back
        newly, called = host[:q_count], host[q_count:2 * q_count]
        first_pos, all_done = host[2 * q_count:3 * q_count], all(f != 0 for f in host[3 * q_count:])
        batcher = self.batcher
        for i in range(q_count):
            if newly[i] and not batcher.fixed[i]:
                batcher.fixed[i] = True
                batcher.native[i]._range_fixed = True   # pylint: disable=protected-access
            self.called_anywhere[i] = self.called_anywhere[i] or bool(called[i])
        for i, call, kept in self.held:
            # calls that precede the chosen first call in global order saw only zeros: the reference skips them
            if batcher.fixed[i] and self.rank * c + call >= first_pos[i]:
                if kept.data_ptr() % 16 == 0 and kept.numel() * kept.element_size() >= 16:
                    batcher.enqueue(i, kept, owned=True)
                else:
                    batcher.update_now(i, kept)
        self.held = []
        self.calls = [0] * q_count
        self.rounds += 1
        waiting = any(self.called_anywhere[i] and not batcher.fixed[i] for i in range(q_count))
        self.need_round = waiting and not all_done
        if self.need_round:
            self.range_block.reset()

    # -- exchange 2: statistics ------------------------------------------------------------------------------------
    def _merge_tf(self):
        q_count = len(self.quantizers)
        rec = self.block.bytes_view().view(torch.float64).view(q_count, -1)
        base = field_index("run_min", 8)
        mins, maxs = rec[:, base].clone(), rec[:, base + 1].clone()
        updated = torch.tensor([float(q._cppOp[0]._is_encoding_valid) for q in self.quantizers],   # pylint: disable=protected-access
                               device=self.device, dtype=torch.float64)
        if self.world > 1:
            _all_reduce(mins, dist.ReduceOp.MIN, self.group)
            _all_reduce(maxs, dist.ReduceOp.MAX, self.group)
            _all_reduce(updated, dist.ReduceOp.MAX, self.group)
        rec[:, base] = mins
        rec[:, base + 1] = maxs
        flags = self.block.bytes_view().view(torch.int32).view(q_count, -1)
        flags[:, field_index("stats_updated", 4)] = updated.to(torch.int32)
        for q, u in zip(self.quantizers, updated.tolist()):
            q._cppOp[0]._is_encoding_valid = bool(u)   # pylint: disable=protected-access

    def _merge(self):
        if not self.tfe:
            self._merge_tf()
            return
        q_count = len(self.quantizers)
        if self.batcher.forward < 0:
            # a rank that was dealt no batch: its wrappers never ran, so nothing has derived its parameter encodings
            from .quantsim.qc_quantize_op import StaticGridQuantWrapper
            with torch.no_grad():
                for _, w in self.sim.quant_wrappers():
                    if isinstance(w, StaticGridQuantWrapper):
                        w.ensure_param_encodings()
        # ranks that are out of batches join the range rounds the others may still run (same collective sequence everywhere)
        while self.need_round:
            self._range_round(done=True)
        self.batcher.flush()
        sink = self.sink
        # Everything the replay needs besides the counts themselves -- how many rows each rank logged, their (local batch,
        # record) tags, which records were called -- is known on the HOST the moment the last forward has been issued, while
        # the device is still tens of milliseconds behind. That exchange, its read-back and the replay plan therefore run on
        # a side stream that does not wait for the forwards; the main stream gets one gather of the rows and the fold launch
        # queued behind the last forward, with no host synchronisation in between.
        main = torch.cuda.current_stream(self.device)
        side = getattr(self, "_side", None)
        if side is None:
            side = self._side = torch.cuda.Stream(self.device)
        with torch.cuda.stream(side):
            used = torch.tensor([sink.used], dtype=torch.int64).to(self.device)
            if self.world > 1:
                _all_reduce(used, dist.ReduceOp.MAX, self.group)
            rows = int(used.item())
            if rows == 0:
                return
            tags = torch.full((rows, 2), -1, dtype=torch.int32)
            if sink.used:
                tags[:sink.used] = torch.tensor(sink.meta, dtype=torch.int32).view(-1, 2)
            called = torch.tensor([int(q._cppOp[0]._is_encoding_valid) for q in self.quantizers], dtype=torch.int32)   # pylint: disable=protected-access
            mine = torch.cat([tags.view(-1), called]).to(self.device)
            tail = (_all_gather(mine, self.group) if self.world > 1 else mine.unsqueeze(0)).cpu()   # read-back, side stream only
            meta = tail[:, :2 * rows].reshape(self.world, rows, 2)
            called_anywhere = tail[:, 2 * rows:].max(dim=0).values.tolist()
            entry_rows, record_begin = replay_plan(meta, q_count)    # rows of rank w sit at w * rows + r in the gathered log
            entry_rows, record_begin = entry_rows.to(self.device), record_begin.to(self.device)
            updated = torch.tensor(called_anywhere, dtype=torch.int32).to(self.device)
            for t in (entry_rows, record_begin, updated):
                t.record_stream(main)
            ready = side.record_event()
        main.wait_event(ready)
        payload = torch.zeros((rows, LOG_WORDS), dtype=torch.int32, device=self.device)
        payload[:sink.used] = sink.rows[:sink.used]
        gathered = _all_gather(payload, self.group) if self.world > 1 else payload.unsqueeze(0)
        ops.stats_fold_log_impl(self.block.arena, self.block.first, q_count, gathered.view(-1, LOG_WORDS), entry_rows,
                                record_begin)
        flags = self.block.bytes_view().view(torch.int32).view(q_count, -1)
        flags[:, field_index("stats_updated", 4)] = updated
        for q, u in zip(self.quantizers, called_anywhere):
            q._cppOp[0]._is_encoding_valid = bool(u)   # pylint: disable=protected-access
            q._stats_dirty = True                      # pylint: disable=protected-access

    # -- public ----------------------------------------------------------------------------------------------------
    def compute_encodings_for_batches(self, batches, cuda_graph: bool = True):
        """Like compute_encodings with the callback `for x in batches: model(x)` over THIS rank's batches, with the
        steady state replayed from a CUDA graph (see QuantizationSimModel.compute_encodings_for_batches)."""
        from .quantsim.qc_quantize_op import CalibrationJob
        from .quantsim.quantsim import QuantizationSimModel, in_eval_mode, run_batches
        sim = self.sim
        QuantizationSimModel.prepare_sim_for_compute_encodings(sim)
        if not getattr(sim, "_act_block_quantizers", None):
            raise RuntimeError("sharded calibration needs the model on a CUDA device")
        self._attach(staging=cuda_graph)
        try:
            after_each = (lambda n: self.sink.commit(n)) if (self.tfe and cuda_graph) else None
            with in_eval_mode(sim.model), torch.no_grad(), CalibrationJob(sim):
                run_batches(sim.model, batches, cuda_graph, after_each=after_each,
                            after_first=None if self.tfe else sim._learn_fixed_ranges)   # pylint: disable=protected-access
            self._merge()
        finally:
            self._detach()
        QuantizationSimModel.compute_layer_encodings_for_sim(sim)

    def compute_encodings(self, forward_pass_callback, forward_pass_callback_args):
        from .quantsim.qc_quantize_op import CalibrationJob
        from .quantsim.quantsim import QuantizationSimModel, _ParamExportPrefetch, in_eval_mode
        sim = self.sim
        QuantizationSimModel.prepare_sim_for_compute_encodings(sim)
        if not getattr(sim, "_act_block_quantizers", None):
            raise RuntimeError("sharded calibration needs the model on a CUDA device")
        self._attach(staging=False)
        prefetch = _ParamExportPrefetch(sim)     # parameter encodings are rank-local and final after the first forward
        try:
            with in_eval_mode(sim.model), torch.no_grad(), CalibrationJob(sim):
                forward_pass_callback(sim.model, forward_pass_callback_args)
            self._merge()
        finally:
            prefetch.close()
            self._detach()
        QuantizationSimModel.compute_layer_encodings_for_sim(sim)
