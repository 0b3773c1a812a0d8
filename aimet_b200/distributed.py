"""Calibration sharded by batch across GPUs, with encodings bit-identical to a single-process run.

Net-new relative to the reference (it has no collective anywhere: SURVEY.md sections 2.3, 8e). One process per GPU;
rank r of W runs global batches r, r+W, r+2W, ...; parameter encodings are computed redundantly (identical weights on
every rank); activation statistics are merged with two collectives on `torch.distributed` (NCCL over NVLink on a B200
box, gloo in the CPU tests):

  tf           the running (min, max) of every quantizer: all_reduce(MIN) / all_reduce(MAX) -- exact.
  tf_enhanced  the reference's result depends on (i) the histogram range fixed by the first non-zero batch in GLOBAL
               order and (ii) a sequential running mean over batches (DlQuantization/src/math_functions.cpp:248-287).
               (1) during its first local batch a rank only records each activation's (min, max) and keeps a copy of the
                   tensor; one all_gather of the [Q, calls, 2] table lets every rank pick, per quantizer, the range the
                   globally first non-zero call defines, fix it on the device (ab_stats_init_range) and only then bin
                   the kept tensors -- nobody waits for rank 0's batch to finish;
               (2) every batch's raw integer counts are logged on the device ([slot, Q, 514] uint32);
               (3) one all_gather of the logs, then ONE kernel replays pdf = (pdf*k + hist/cnt)/(k+1) in global batch
                   order for all quantizers (ab_stats_fold_batches). Integer counts make the merge exact.

The pure-tensor helpers (`choose_first_ranges`, `global_replay_offsets`) run on any device and are what the gloo tests
exercise; the device work goes through aimet_b200.ops.
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

    # -- hooks ---------------------------------------------------------------------------------------------------
    def _install(self):
        sim = self.sim
        self.quantizers = list(sim._act_block_quantizers)   # pylint: disable=protected-access
        self.block = sim._act_block                         # pylint: disable=protected-access
        self.device = self.block.device
        warm_collectives(self.group, self.device)
        q_count = len(self.quantizers)
        self.tfe = sim._quant_scheme == QuantScheme.post_training_tf_enhanced   # pylint: disable=protected-access
        self.local_batch = -1
        self.calls = [0] * q_count
        self.deferred: List[List[torch.Tensor]] = [[] for _ in range(q_count)]
        self.first_block = StateArena.for_device(self.device).allocate(q_count * self.max_calls) if self.tfe else None
        self.log = None
        self.log_slots = 0
        # every call of the current batch logs into a fixed staging area (constant addresses, so the step can live in a
        # CUDA graph); `_end_batch` files it under the batch's slot
        self.stage = torch.zeros((self.max_calls, q_count, LOG_WORDS), dtype=torch.int32, device=self.device) \
            if self.tfe else None
        self.staged_batch = -1
        self.fixed = [False] * q_count
        self.manual_batches = False
        self.ranges_fixed = not self.tfe
        self._pre = sim.model.register_forward_pre_hook(lambda m, a: self._begin_batch())
        for i, q in enumerate(self.quantizers):
            q._calib_hook = (lambda t, i=i: self._on_update(i, t))   # pylint: disable=protected-access

    def _uninstall(self):
        self._pre.remove()
        for q in self.quantizers:
            q._calib_hook = None   # pylint: disable=protected-access
        self.deferred = []
        self.first_block = None

    def _begin_batch(self):
        if self.manual_batches:
            return
        self._end_batch()
        if self.local_batch == 0 and not self.ranges_fixed:
            self._fix_ranges()
        self.local_batch += 1
        self.calls = [0] * len(self.quantizers)
        if self.tfe and self.ranges_fixed:
            self.stage.zero_()     # the statistics launches ADD their counts to their log entry (one memset per batch)

    def _end_batch(self):
        """File the staged log of the batch that just ran under its slot."""
        if self.tfe and self.ranges_fixed and self.local_batch >= 0 and self.staged_batch != self.local_batch:
            first = self.local_batch * self.max_calls
            self._ensure_log(first + self.max_calls)
            self.log[first:first + self.max_calls].copy_(self.stage)
            self.staged_batch = self.local_batch

    def _ensure_log(self, slots):
        if self.log is None or slots > self.log_slots:
            new_slots = max(slots, 2 * self.log_slots, 8 * self.max_calls)
            new = torch.zeros((new_slots, len(self.quantizers), LOG_WORDS), dtype=torch.int32, device=self.device)
            if self.log is not None:
                new[:self.log_slots] = self.log
            self.log, self.log_slots = new, new_slots

    def _on_update(self, i: int, tensor: torch.Tensor):
        q = self.quantizers[i]
        op = q._cppOp[0]   # pylint: disable=protected-access
        call = self.calls[i]
        self.calls[i] += 1
        if call >= self.max_calls:
            raise RuntimeError(f"a quantizer was updated more than max_calls_per_batch={self.max_calls} times in one "
                               "forward pass; raise the limit")
        if tensor.dtype not in (torch.float32, torch.bfloat16):
            tensor = tensor.to(torch.float32)
        op._is_encoding_valid = True   # pylint: disable=protected-access
        q._stats_dirty = True          # pylint: disable=protected-access
        if not self.tfe:
            ops.stats_update_impl(tensor, self.block.arena, self.block.first + i, ops.QUANTIZATION_TF, None, 0)
            return
        if not self.ranges_fixed:
            # first local batch: record (min, max) of this call, keep the tensor for binning once the range is known
            ops.stats_update_impl(tensor, self.first_block.arena, self.first_block.first + i * self.max_calls + call,
                                  ops.QUANTIZATION_TF, None, 0)
            self.deferred[i].append(tensor.detach().clone())
            return
        ops.stats_update_impl(tensor, self.block.arena, self.block.first + i, ops.QUANTIZATION_TF_ENHANCED, self.stage,
                              call * len(self.quantizers) + i, ops.STATS_RANGE_FIXED if self.fixed[i] else 0)

    # -- exchange 1: ranges --------------------------------------------------------------------------------------
    def _fix_ranges(self):
        q_count = len(self.quantizers)
        rec = self.first_block.bytes_view().view(torch.float64).view(q_count * self.max_calls, -1)
        base = field_index("run_min", 8)          # run_min, run_max are adjacent doubles (include/aimet_b200.h)
        table = rec[:, base:base + 2].to(torch.float32).view(q_count, self.max_calls, 2)
        # an un-updated record holds (+DBL_MAX, -DBL_MAX) -> (+inf, -inf) in float32: "call did not happen"
        gathered = _all_gather(table, self.group) if self.world > 1 else table.unsqueeze(0)
        chosen = choose_first_ranges(gathered)
        self.first_positions = first_call_positions(gathered)
        ops.stats_init_range_impl(self.block.arena, self.block.first, q_count, chosen)
        self.ranges_fixed = True
        self.fixed = ((chosen[:, 0] != 0) | (chosen[:, 1] != 0)).tolist()     # the one host read-back of the job
        # now bin the tensors kept from local batch 0 (into the staging log, filed by _end_batch)
        self.stage.zero_()
        for i, kept in enumerate(self.deferred):
            for call, tensor in enumerate(kept):
                ops.stats_update_impl(tensor, self.block.arena, self.block.first + i, ops.QUANTIZATION_TF_ENHANCED,
                                      self.stage, call * q_count + i, ops.STATS_RANGE_FIXED if self.fixed[i] else 0)
        self.deferred = [[] for _ in range(q_count)]
        self._end_batch()

    # -- exchange 2: statistics ------------------------------------------------------------------------------------
    def _merge(self):
        q_count = len(self.quantizers)
        if not self.tfe:
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
            return
        if not self.ranges_fixed:          # the callback ran a single batch (or none)
            if self.local_batch >= 0:
                self._fix_ranges()
        self._end_batch()
        local_batches = torch.tensor([self.local_batch + 1], device=self.device, dtype=torch.int64)
        if self.world > 1:
            _all_reduce(local_batches, dist.ReduceOp.MAX, self.group)
        n_local = int(local_batches.item())
        if n_local == 0:
            return
        slots = n_local * self.max_calls
        self._ensure_log(slots)
        local_log = self.log[:slots].contiguous()
        # the reference SKIPS all-zero batches seen before the range was fixed: entries that precede the chosen first
        # call were binned here with a range they would not have had yet -> void them (count 0 == skipped in the replay)
        if hasattr(self, "first_positions"):
            pos = self.rank * self.max_calls + torch.arange(self.max_calls, device=self.device)      # batch 0 of this rank
            void = pos[:, None] < self.first_positions[None, :].to(self.device)                        # [calls, Q]
            local_log[:self.max_calls][void] = 0
        gathered = _all_gather(local_log, self.group) if self.world > 1 else local_log.unsqueeze(0)
        offsets = global_replay_offsets(self.world, n_local, self.max_calls, q_count).to(self.device)
        ops.stats_fold_batches_impl(self.block.arena, self.block.first, q_count, gathered, offsets)
        counts = gathered.view(self.world * slots, q_count, LOG_WORDS)[:, :, 512:].to(torch.int64).sum(dim=(0, 2))
        for q, n in zip(self.quantizers, counts.tolist()):
            q._cppOp[0]._is_encoding_valid = q._cppOp[0]._is_encoding_valid or n > 0   # pylint: disable=protected-access

    # -- public ----------------------------------------------------------------------------------------------------
    def compute_encodings_for_batches(self, batches, cuda_graph: bool = True):
        """Like compute_encodings with the callback `for x in batches: model(x)` over THIS rank's batches, with the
        steady state replayed from a CUDA graph (see QuantizationSimModel.compute_encodings_for_batches)."""
        from .quantsim.quantsim import QuantizationSimModel, in_eval_mode, run_batches
        sim = self.sim
        QuantizationSimModel.prepare_sim_for_compute_encodings(sim)
        if not getattr(sim, "_act_block_quantizers", None):
            raise RuntimeError("sharded calibration needs the model on a CUDA device")
        self._install()
        self.manual_batches = True
        try:
            def start_batch():
                self.local_batch += 1
                self.calls = [0] * len(self.quantizers)
                if self.tfe and self.ranges_fixed:
                    self.stage.zero_()          # before the forward, eager or replayed: the launches add to their entries

            def after_each(n):
                if n == 0:
                    if not self.ranges_fixed:
                        self._fix_ranges()      # exchange 1, then bins the kept tensors of batch 0 and files them
                else:
                    self._end_batch()
                start_batch()                   # the next forward (eager or replayed) belongs to the next batch

            with in_eval_mode(sim.model), torch.no_grad():
                start_batch()
                n = run_batches(sim.model, batches, cuda_graph, after_each=after_each)
                self.local_batch = n - 1        # after_each advanced one past the last batch
            self._merge()
        finally:
            self._uninstall()
        QuantizationSimModel.compute_layer_encodings_for_sim(sim)

    def compute_encodings(self, forward_pass_callback, forward_pass_callback_args):
        from .quantsim.quantsim import QuantizationSimModel, _ParamExportPrefetch, in_eval_mode
        sim = self.sim
        QuantizationSimModel.prepare_sim_for_compute_encodings(sim)
        if not getattr(sim, "_act_block_quantizers", None):
            raise RuntimeError("sharded calibration needs the model on a CUDA device")
        self._install()
        prefetch = _ParamExportPrefetch(sim)     # parameter encodings are rank-local and final after the first forward
        try:
            with in_eval_mode(sim.model), torch.no_grad():
                forward_pass_callback(sim.model, forward_pass_callback_args)
            self._merge()
        finally:
            prefetch.close()
            self._uninstall()
        QuantizationSimModel.compute_layer_encodings_for_sim(sim)
