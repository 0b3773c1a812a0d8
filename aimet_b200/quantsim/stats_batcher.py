"""Deferred, multi-tensor statistics for the activation quantizers of one calibration job.

The reference makes one native `updateStats` call per activation tensor, inside the wrapper's forward
(aimet_torch/v1/qc_quantize_op.py:837-897 -> v1/tensor_quantizer.py:452-480 -> AimetTensorQuantizer.cpp:94-126). One
launch per tensor pays a pipeline ramp and a launch gap for every one of the ~70 tensors of a ResNet-50 forward, which is
what kept the histogram kernel at a third of the HBM roofline inside the workload. In ANALYSIS mode the wrapper hands the
tensor on untouched and nothing downstream waits for the statistics, so the call can be DEFERRED to the end of the forward
and all deferred calls issued as one launch (`ab_stats_update_multi`) -- provided the tensor still holds the same values
then. What can break that is an in-place write after the call (torchvision's `out += identity` writes into the output of
bn3, `nn.ReLU(inplace=True)` into the output of a batch norm). The batcher therefore

  * runs the first forward of a job exactly as before (one launch per call, which also fixes the histogram ranges) and
    records each call's tensor and its autograd version counter; at the end of that forward a quantizer is marked
    deferrable iff none of its tensors was written to after its call (version unchanged) and its range is fixed;
  * from the second forward on, queues the calls of deferrable quantizers -- (tensor, record index, version) -- and flushes
    the queue in ONE launch at the end of the forward (or earlier: 128 entries, FLUSH_BYTES of tensors kept alive, an
    out-of-order call on the same record, somebody reading the statistics);
  * re-checks the version of every queued tensor at flush time and raises if one changed: a forward whose in-place
    behaviour differs from batch to batch cannot be deferred safely (AB_DEFER_STATS=0 switches the batcher off).

Results are bit-identical to the per-call path (tests/test_gpu_stats_multi.py, tests/test_gpu_quantsim.py): counts are
integers, and the fold launch replays the calls of every record in call order.
"""
import os

import torch

from .. import ops
from ..state import field_index

ENABLED = os.environ.get("AB_DEFER_STATS", "1") != "0"
FLUSH_BYTES = int(os.environ.get("AB_DEFER_STATS_BYTES", str(8 << 30)))   # bound on the activations kept alive
_INITIALIZED_WORD = field_index("initialized", 4)


class LogSink:
    """Where the raw counts of every call go when the job keeps a per-call log instead of folding locally (multi-GPU exact
    merge, aimet_b200.distributed): rows of LOG_WORDS int32 words, handed out in call order; `meta[k]` = (local batch,
    record index) of row k. With `staging_rows`, rows are first handed out from a fixed staging area -- constant
    addresses, so the step can live in a CUDA graph -- and `commit` files them under the batch that just ran."""

    def __init__(self, device, capacity=2048, staging_rows=0):
        self.device = device
        self.rows = torch.zeros((capacity, ops.LOG_WORDS), dtype=torch.int32, device=device)
        self.used = 0
        self.meta = []
        self.staging = torch.zeros((staging_rows, ops.LOG_WORDS), dtype=torch.int32, device=device) \
            if staging_rows else None
        self.staged = 0
        self.staged_records = []
        self.last_step_records = []

    def _reserve(self, n):
        if self.used + n > self.rows.shape[0]:
            grown = torch.zeros((max(2 * self.rows.shape[0], self.used + n), ops.LOG_WORDS), dtype=torch.int32,
                                device=self.device)
            grown[:self.used] = self.rows[:self.used]
            self.rows = grown

    def take(self, n, metas):
        """-> int32 view [n, LOG_WORDS] of zeroed rows for the next n calls."""
        if self.staging is not None:
            if self.staged + n > self.staging.shape[0]:
                raise RuntimeError("more statistics calls in one forward pass than the staging log holds")
            first = self.staged
            self.staged += n
            self.staged_records.extend(r for _, r in metas)
            return self.staging[first:first + n]
        self._reserve(n)
        first = self.used
        self.used += n
        self.meta.extend(metas)
        return self.rows[first:first + n]

    def commit(self, local_batch):
        """Staging mode, after every forward (eager or replayed): file the staged rows under `local_batch`. A replayed
        forward ran no Python, so it staged the same calls as the captured one."""
        if self.staging is None:
            return
        if self.staged:
            self.last_step_records, self.staged_records = self.staged_records, []
        n = len(self.last_step_records)
        self.staged = 0
        if n == 0:
            return
        self._reserve(n)
        self.rows[self.used:self.used + n].copy_(self.staging[:n])
        self.staging[:n].zero_()
        self.meta.extend((local_batch, r) for r in self.last_step_records)
        self.used += n


class StatsBatcher:
    def __init__(self, sim, sink: LogSink = None):
        self.sim = sim
        self.block = sim._act_block                          # pylint: disable=protected-access
        self.quantizers = list(sim._act_block_quantizers)    # pylint: disable=protected-access
        self.native = [q._cppOp[0] for q in self.quantizers]  # pylint: disable=protected-access
        n = len(self.quantizers)
        self.sink = sink
        self.forward = -1                 # index of the forward in progress (or last finished)
        self.in_forward = False
        self.probing = True
        self.probe = []                   # (record, tensor, version) of the first forward's calls
        self.fixed = [False] * n          # host knowledge: the record's histogram range is fixed
        self.defer = [False] * n
        self.pending = []                 # (tensor, record, version, owned)
        self.pending_records = set()
        self.pending_bytes = 0
        self.scratch = None if sink is not None else \
            torch.zeros((ops.MULTI_MAX_SEGMENTS, ops.LOG_WORDS), dtype=torch.int32, device=self.block.device)
        self._handles = []
        # sharded calibration (aimet_b200.distributed): histogram ranges come from a global exchange, not from this rank's
        # first batch. `collect(i, tensor)` takes the calls on records whose range is not known yet; `range_exchange(self)`
        # runs at the end of every forward, before the flush: it may fix ranges (`fixed`) and enqueue tensors it kept.
        self.collect = None
        self.range_exchange = None

    # ---- wiring ------------------------------------------------------------------------------------------------
    @classmethod
    def attach(cls, sim, sink=None):
        """A batcher hooked into `sim` for one calibration job, or None where deferral does not apply (no CUDA records,
        a scheme that keeps no histogram, switched off). With a `sink` (sharded calibration) the batcher is the logging
        mechanism itself and is attached even when deferral is switched off: every call is then issued immediately."""
        quantizers = getattr(sim, "_act_block_quantizers", None)
        if (not ENABLED and sink is None) or not quantizers or getattr(sim, "_act_block", None) is None:
            return None
        if not all(ops.keeps_histogram(q._cppOp[0]._code) for q in quantizers):   # pylint: disable=protected-access
            return None
        self = cls(sim, sink)
        self._handles = [sim.model.register_forward_pre_hook(lambda m, a: self._begin_forward()),
                         sim.model.register_forward_hook(lambda m, a, o: self._end_forward())]
        for i, q in enumerate(self.quantizers):
            q._calib_hook = (lambda t, i=i: self._on_update(i, t))   # pylint: disable=protected-access
        return self

    def detach(self):
        for h in self._handles:
            h.remove()
        self._handles = []
        for q in self.quantizers:
            q._calib_hook = None   # pylint: disable=protected-access
        self.pending, self.probe = [], []
        self.pending_records.clear()

    # ---- forward boundaries ------------------------------------------------------------------------------------
    def _begin_forward(self):
        self.forward += 1
        self.in_forward = True

    def _end_forward(self):
        self.in_forward = False
        if self.range_exchange is not None:
            self.range_exchange(self)
        if self.probing and self.forward == 0:
            self._finish_probe()
        self.flush()

    def _finish_probe(self):
        """Classify the quantizers after the first forward: deferrable = range fixed and no tensor written after its call."""
        n = len(self.quantizers)
        called, unstable = [False] * n, [False] * n
        for i, t, version in self.probe:
            called[i] = True
            if t._version != version:   # pylint: disable=protected-access
                unstable[i] = True
        self.probe = []
        if self.range_exchange is None:
            words = self.block.bytes_view().view(torch.int32).view(n, -1)
            flags = words[:, _INITIALIZED_WORD].cpu().tolist()      # the one host read-back of the job
            self.fixed = [bool(f) for f in flags]
        for i in range(n):
            self.native[i]._range_fixed = self.fixed[i]          # pylint: disable=protected-access
            self.defer[i] = ENABLED and self.fixed[i] and called[i] and not unstable[i]
        self.probing = False

    # ---- the call ------------------------------------------------------------------------------------------------
    def _on_update(self, i, tensor):
        q, op = self.quantizers[i], self.native[i]
        owned = False
        if tensor.dtype not in (torch.float32, torch.bfloat16):
            tensor, owned = tensor.to(torch.float32), True
        if not (tensor.is_contiguous() or (tensor.dim() == 4 and tensor.is_contiguous(memory_format=torch.channels_last))):
            tensor, owned = tensor.contiguous(), True
        op._is_encoding_valid = True      # pylint: disable=protected-access
        op._updates += 1                  # pylint: disable=protected-access
        q._stats_dirty = True             # pylint: disable=protected-access
        es = tensor.element_size()
        if self.collect is not None and not self.fixed[i]:
            if self.probing and not owned:
                self.probe.append((i, tensor, tensor._version))   # pylint: disable=protected-access
            self.collect(i, tensor)
            return
        if (self.defer[i] or (owned and self.fixed[i])) and tensor.data_ptr() % 16 == 0 and tensor.numel() * es >= 16:
            self.enqueue(i, tensor, owned)
            return
        self.update_now(i, tensor)        # (flushes first when the record has queued calls: call order is fold order)
        if self.probing and not owned:
            self.probe.append((i, tensor, tensor._version))   # pylint: disable=protected-access

    def enqueue(self, i, tensor, owned):
        """Queue one call for the next flush. `owned`: the tensor is a private copy nobody else can write to."""
        self.pending.append((tensor, i, tensor._version, owned))   # pylint: disable=protected-access
        self.pending_records.add(i)
        self.pending_bytes += tensor.numel() * tensor.element_size()
        if len(self.pending) >= ops.MULTI_MAX_SEGMENTS or self.pending_bytes >= FLUSH_BYTES:
            self.flush()

    def update_now(self, i, tensor):
        """One call issued immediately through the single-tensor kernels (a range must be fixed in sharded mode)."""
        if i in self.pending_records:
            self.flush()
        row = self.sink.take(1, [(self.forward, i)]) if self.sink is not None else None
        ops.stats_update_impl(tensor, self.block.arena, self.block.first + i, ops.QUANTIZATION_TF_ENHANCED, row, 0,
                              ops.STATS_RANGE_FIXED if self.fixed[i] else 0)

    def flush(self):
        """Issue every queued call: one histogram launch + one fold launch per run of up to 128 same-dtype tensors."""
        if not self.pending:
            return
        pending, self.pending = self.pending, []
        self.pending_records.clear()
        self.pending_bytes = 0
        for t, i, version, owned in pending:
            if not owned and t._version != version:   # pylint: disable=protected-access
                raise RuntimeError(
                    "aimet_b200: an activation tensor was modified in place between its updateStats call and the end of "
                    "the forward pass, although the first batch did not do that (data-dependent in-place operation?). "
                    "Deferred statistics cannot be used with this model: set AB_DEFER_STATS=0.")
        at = 0
        while at < len(pending):
            dtype = pending[at][0].dtype
            end = at
            while end < len(pending) and end - at < ops.MULTI_MAX_SEGMENTS and pending[end][0].dtype == dtype:
                end += 1
            chunk = pending[at:end]
            tensors = [c[0] for c in chunk]
            records = [c[1] for c in chunk]
            if self.sink is not None:
                rows = self.sink.take(len(chunk), [(self.forward, r) for r in records])
                ops.stats_update_multi_impl(tensors, records, self.block.arena, self.block.first, rows, log_only=True)
            else:
                ops.stats_update_multi_impl(tensors, records, self.block.arena, self.block.first, self.scratch)
            at = end
