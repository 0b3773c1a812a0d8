"""Drop-in for the reference's torch extension module ``aimet_common.AimetTensorQuantizer``.

Same class name, same nine methods, same argument meaning as the pybind11 class in
TrainingExtensions/torch/src/AimetTensorQuantizer.cpp:75-331 -- what aimet_torch/v1/tensor_quantizer.py calls as
``self._cppOp[i].updateStats / getEncoding / quantizeDequantize / quantizeDequantizePerChannel / quantize``.

Differences a caller can observe, all deliberate:
  * computation always runs on a CUDA device through the sm_100a kernels. A CPU tensor (the reference's
    `use_cuda=False` path, e.g. the two-element tensor built for encoding_min_max_fixed_vals) is staged to the current
    device and the result copied back; without a CUDA device every method raises. There is no CPU fallback.
  * bfloat16 tensors are accepted natively (the reference reads `input.data<float>()` only).
  * `updateStats` never synchronises: statistics are device resident; the only host round trip is in `getEncoding`
    / `getStatsHistogram`.
"""
import os
import threading

import torch

from . import libpymo
from . import ops
from .state import StateArena, field_index


def _to_device_tensor(t: torch.Tensor):
    """-> (fp32/bf16 CUDA tensor, original device)"""
    orig = t.device
    if not t.is_cuda:
        if not torch.cuda.is_available():
            raise RuntimeError("aimet_b200 needs a CUDA device: there is no CPU path")
        t = t.to(torch.device("cuda", torch.cuda.current_device()))
    if t.dtype not in (torch.float32, torch.bfloat16):
        t = t.to(torch.float32)   # the reference upcasts half tensors before the native call (tensor_quantizer.py:477)
    return t, orig


class ValidityGroup:
    """`isEncodingValid` shared by the ops of one per-channel quantizer: the host layer updates and resets all channels
    together (one launch), so it flips ONE flag instead of looping over up to thousands of ops. An op that is written
    individually leaves the group (`detached` counts those), and the host layer falls back to looking at every op."""
    __slots__ = ("valid", "detached")

    def __init__(self):
        self.valid = False
        self.detached = 0


class AimetTensorQuantizer:
    """One encoding analyzer + one quantize-dequantize simulator (AimetTensorQuantizer.cpp:75-83)."""

    @property
    def _is_encoding_valid(self):
        group = self._group
        return group.valid if group is not None else self._valid

    @_is_encoding_valid.setter
    def _is_encoding_valid(self, value):
        if self._group is not None:
            self._group.detached += 1
            self._group = None
        self._valid = bool(value)

    def __init__(self, quantization_scheme):
        self._group = None
        self._valid = False
        self._scheme = libpymo.QuantizationMode(int(quantization_scheme))
        self._code = libpymo.scheme_code(self._scheme)
        self._percentile = 100.0 if libpymo.is_percentile(self._scheme) else None   # None: not the percentile scheme
        self._is_encoding_valid = False
        self._block = None            # StateBlock with one record, allocated on first use
        self._index = 0
        self._pc_cache = None         # (key, device tensor) for the per-channel parameter block
        self._range_fixed = False     # host knowledge that the device record's histogram range is fixed
        self._probe = None            # (pinned int32 tensor, event) of an in-flight read-back of `initialized`
        self._updates = 0             # updateStats calls since the last reset / bind

    # ---- wiring used by the batched host layer (aimet_b200.quantsim): share one contiguous block per weight -------
    def _bind(self, block, index, group=None):
        self._block, self._index = block, index
        self._range_fixed, self._probe, self._updates = False, None, 0
        if group is not None:
            self._group = group

    def _ensure_state(self, device):
        if self._block is None or self._block.device != device:
            self._block, self._index = StateArena.for_device(device).allocate(1), 0

    # ---- reference API ---------------------------------------------------------------------------------------
    def _reset_host_state(self):
        """The host half of resetEncodingStats; the caller has reset (or is about to reset) the record itself, together with
        its neighbours, in one launch."""
        self._is_encoding_valid = False
        if self._percentile is not None:
            self._percentile = 100.0
        self._range_fixed, self._probe, self._updates = False, None, 0

    def resetEncodingStats(self):
        """AimetTensorQuantizer.cpp:85-92 (a new analyzer: the percentile value falls back to its default as well)"""
        self._reset_host_state()
        if self._block is not None:
            ops.stats_reset_impl(self._block.arena, self._block.first + self._index, 1)

    def updateStats(self, input, use_cuda):   # pylint: disable=redefined-builtin
        """AimetTensorQuantizer.cpp:94-126"""
        t, _ = _to_device_tensor(input)
        self._is_encoding_valid = True
        self._ensure_state(t.device)
        if self._code == ops.QUANTIZATION_ENTROPY:
            ops.entropy_update_impl(t, self._block.arena, self._block.first + self._index)
            return
        ops.stats_update_impl(t, self._block.arena, self._block.first + self._index, self._code, None, 0,
                              ops.STATS_RANGE_FIXED if self._range_fixed else 0)
        self._updates += 1
        # The probe only pays off when the same record is updated again and again (calibration). A record that is reset
        # before every update -- a weight quantizer in training mode -- never gets there, and a probe costs a pinned
        # buffer, a copy and an event (2.4 ms per MobileNet-v2 QAT step): start with the second update after a reset.
        if ops.keeps_histogram(self._code) and not self._range_fixed and self._updates >= 2:
            self._poll_range_fixed()

    _INITIALIZED_WORD = field_index("initialized", 4)      # ab_stats_state.initialized as an int32 index

    def _poll_range_fixed(self):
        """Learn, without ever blocking, that the record's range got fixed: a 4-byte asynchronous read-back of the
        `initialized` flag into pinned memory, looked at on a later call once its event has completed."""
        if self._probe is not None:
            flag, event = self._probe
            if not event.query():
                return
            self._probe = None
            if int(flag[0]) == 1:
                self._range_fixed = True
                return
        if torch.cuda.is_current_stream_capturing():
            return
        rec = self._block.arena[(self._block.first + self._index) * ops.STATE_BYTES:
                                (self._block.first + self._index + 1) * ops.STATE_BYTES].view(torch.int32)
        flag = torch.empty(1, dtype=torch.int32, pin_memory=True)
        flag.copy_(rec[self._INITIALIZED_WORD:self._INITIALIZED_WORD + 1], non_blocking=True)
        event = torch.cuda.Event()
        event.record(torch.cuda.current_stream(self._block.device))
        self._probe = (flag, event)

    def getEncoding(self, bitwidth, use_symmetric_encodings, use_strict_symmetric, use_unsigned_symmetric):
        """AimetTensorQuantizer.cpp:180-192 -> (TfEncoding, is_valid)"""
        if not self._is_encoding_valid or self._block is None:
            return libpymo.TfEncoding(), self._is_encoding_valid
        if self._code == ops.QUANTIZATION_ENTROPY:
            return libpymo.TfEncoding._from_c(ops.entropy_compute_impl(
                self._block.arena, self._block.first + self._index, bitwidth, use_symmetric_encodings, use_strict_symmetric,
                use_unsigned_symmetric)), True
        if use_symmetric_encodings and self._code == ops.QUANTIZATION_TF:
            assert not (use_strict_symmetric and use_unsigned_symmetric)   # TfEncodingAnalyzer.cpp:85-86
        enc, _ = ops.compute_encodings_impl(self._block.arena, self._block.first + self._index, 1, self._code,
                                            bitwidth, use_symmetric_encodings, use_strict_symmetric,
                                            use_unsigned_symmetric, percentile=self._percentile)
        v = enc[0].tolist()
        return libpymo.TfEncoding._from_values(v[0], v[1], v[2], v[3], int(v[4])), True

    def quantizeDequantize(self, input, encoding, rounding_mode, use_cuda):   # pylint: disable=redefined-builtin
        """AimetTensorQuantizer.cpp:129-153: uses only encoding.min / max / bw."""
        t, orig = _to_device_tensor(input)
        fn = torch.ops.aimet_b200.qdq_per_tensor if (t.requires_grad and torch.is_grad_enabled()) else \
            ops.qdq_per_tensor_impl
        out = fn(t, float(encoding.min), float(encoding.max), int(encoding.bw), int(rounding_mode), _next_seed())
        return _restore(out, input, orig)

    def quantize(self, input, encoding, rounding_mode, use_cuda, shift_to_signed):   # pylint: disable=redefined-builtin
        """AimetTensorQuantizer.cpp:155-178"""
        t, orig = _to_device_tensor(input)
        out = torch.ops.aimet_b200.quantize_to_grid(t, float(encoding.min), float(encoding.max), int(encoding.bw),
                                                    int(rounding_mode), bool(shift_to_signed), _next_seed())
        return _restore(out, input, orig)

    def quantizeDequantizePerChannel(self, input, encodings, num_channel, num_element, num_element_per_channel,   # pylint: disable=redefined-builtin
                                     rounding_mode, use_cuda):
        """AimetTensorQuantizer.cpp:256-307"""
        t, orig = _to_device_tensor(input)
        params = self._per_channel_params(encodings, t.device)
        fn = torch.ops.aimet_b200.qdq_per_channel if (t.requires_grad and torch.is_grad_enabled()) else \
            ops.qdq_per_channel_impl
        out = fn(t, params, int(num_channel), int(num_element_per_channel), int(rounding_mode), _next_seed())
        return _restore(out, input, orig)

    def makeDeltaOffsetTensor(self, device, encodings):
        """AimetTensorQuantizer.cpp:209-234 -> (delta, offset) fp32 tensors on `device`"""
        n = len(encodings)
        host = torch.tensor([[e.delta for e in encodings], [e.offset for e in encodings]], dtype=torch.float32)
        host = host.reshape(2, n).to(device)
        return host[0], host[1]

    def getStatsHistogram(self):
        """AimetTensorQuantizer.cpp:194-198"""
        if not ops.keeps_histogram(self._code):
            raise AssertionError("No real histogram data is kept for TF Encoding analyzer")
        if self._block is None:
            return []
        return self._block.histogram(self._index)

    def setPercentileValue(self, percentile):
        """AimetTensorQuantizer.cpp:200-207: a no-op unless the scheme is percentile."""
        if self._percentile is not None:
            self._percentile = float(percentile)

    # ---- helpers ---------------------------------------------------------------------------------------------
    def _per_channel_params(self, encodings, device):
        # valid while the very same list object is passed and no TfEncoding anywhere has been written since
        key = (device, id(encodings), len(encodings), libpymo.encoding_epoch())
        if self._pc_cache is not None and self._pc_cache[0] == key and self._pc_cache[2] is encodings:
            return self._pc_cache[1]
        if torch.cuda.is_current_stream_capturing():
            raise RuntimeError("per-channel parameters changed while a CUDA graph was being captured")
        host = ops.per_channel_params([e.min for e in encodings], [e.max for e in encodings], encodings[0].bw)
        dev = host.pin_memory().to(device, non_blocking=True) if torch.cuda.is_available() else host.to(device)
        self._pc_cache = (key, dev, encodings)
        return dev


# ---------------------------------------------------------------------------------------------------------------------
# the same class with its native calls DEFERRED and batched -- for callers that drive it one channel at a time
# ---------------------------------------------------------------------------------------------------------------------
class _FailedCalls:
    """Stands in for the queue of an encoding whose deferred calls failed."""

    def __init__(self, exc):
        self.exc = exc

    def resolve(self):
        raise RuntimeError("aimet_b200: the deferred native calls this encoding depends on failed") from self.exc


class _DeferredCalls:
    """FIFO of native calls that have been asked for but not issued yet (one queue per thread).

    The reference's Python drives a per-channel weight quantizer one channel at a time (aimet_torch/v1/tensor_quantizer.py:
    296-305, 567-570; qc_quantize_op.py:231-243): C x resetEncodingStats, C x updateStats(slice c), C x getEncoding -- for
    ResNet-50 26 560 channels, twice per calibration job, each call a launch or two and, for getEncoding, a blocking
    read-back. Queued instead, consecutive calls of one kind on consecutive records become ONE launch (a block-wide reset,
    a segmented statistics launch over the slices of one weight, one grid-search launch) and all encodings owed come back
    in ONE copy; `flush` runs when somebody reads a field of an encoding that is still owed, before any native call that
    is not queued, and when the queue is long. Results are those of the one-by-one calls: same kernels' arithmetic, calls
    on a record issued in their original order. Every run keeps the objects that asked for it: an object with work in the
    queue stays alive, so its record cannot go back to the pool (and to somebody else) before that work has run.

    The queue is a list of RUNS, extended as the calls come in:
        ["R", arena, first record, n, objects]
        ["U", arena, first record, n, objects, scheme code, (pointer, bytes, storage pointer, dtype) of slice 0, tensors, versions]
        ["S", arena, first record, n, objects, (scheme code, bw, sym, strict, unsigned, percentile), encodings]"""
    LIMIT = 1 << 16

    def __init__(self):
        self.lock = threading.RLock()
        self.runs = []
        self.calls = 0
        self.stream = None
        self.last_update = None   # (arena, record, code, where, tensor) of the latest updateStats, issued or queued

    def resolve(self):
        self.flush()

    def _tail(self, kind, arena, rec, device_index):
        """The run a call of `kind` on record `rec` may extend, or None. Lock held."""
        stream = _current_raw_stream(device_index)
        if stream != self.stream:
            if self.runs:
                self.flush()          # queued work belongs to the stream that was current when it was asked for
            self.stream = stream
        if self.calls >= self.LIMIT:
            self.flush()
        self.calls += 1
        if self.runs:
            run = self.runs[-1]
            if run[0] == kind and run[1] is arena and run[2] + run[3] == rec:
                return run
        return None

    def push_reset(self, arena, rec, obj):
        with self.lock:
            run = self._tail("R", arena, rec, arena.device.index)
            if run is None:
                self.runs.append(["R", arena, rec, 1, [obj]])
            else:
                run[3] += 1
                run[4].append(obj)

    def push_update(self, arena, rec, obj, code, where, tensor):
        with self.lock:
            run = self._tail("U", arena, rec, arena.device.index)
            if run is not None and run[5] == code and \
                    where == (run[6][0] + run[3] * run[6][1], run[6][1], run[6][2], run[6][3]):
                run[3] += 1
                run[4].append(obj)
                run[7].append(tensor)
                run[8].append(tensor._version)   # pylint: disable=protected-access
            else:
                self.runs.append(["U", arena, rec, 1, [obj], code, where, [tensor], [tensor._version]])   # pylint: disable=protected-access

    def push_search(self, arena, rec, obj, key, enc):
        with self.lock:
            run = self._tail("S", arena, rec, arena.device.index)
            if run is not None and run[5] == key:
                run[3] += 1
                run[4].append(obj)
                run[6].append(enc)
            else:
                self.runs.append(["S", arena, rec, 1, [obj], key, [enc]])

    def flush(self):
        with self.lock:
            runs, self.runs, self.calls = self.runs, [], 0
            if not runs:
                return
            try:
                self._issue(runs)
            except BaseException as exc:
                # the encodings this queue still owes can never be computed now: reading one must say so
                failed = _FailedCalls(exc)
                for run in runs:
                    if run[0] == "S":
                        for enc in run[6]:
                            if getattr(enc, "_lazy", None) is not None:
                                object.__setattr__(enc, "_lazy", failed)
                raise

    @staticmethod
    def _issue(runs):
        """Issue the runs in order, one launch each; one read-back per device for every encoding owed."""
        owed = {}                                     # device -> number of encodings owed
        for run in runs:
            if run[0] == "S":
                owed[run[1].device] = owed.get(run[1].device, 0) + run[3]
        outs = {d: [torch.empty((c, 5), dtype=torch.float64, device=d), 0, []] for d, c in owed.items()}
        for run in runs:
            kind, arena, first, n = run[0], run[1], run[2], run[3]
            if kind == "R":
                ops.stats_reset_impl(arena, first, n)
            elif kind == "U":
                code, tensors, versions = run[5], run[7], run[8]
                for t, version in zip(tensors, versions):
                    if t._version != version:   # pylint: disable=protected-access
                        raise RuntimeError("aimet_b200: a tensor handed to updateStats was modified in place before its "
                                           "deferred statistics ran; set AB_DEFER_DROPIN=0")
                t0 = tensors[0]
                if n == 1:
                    ops.stats_update_impl(t0, arena, first, code, None, 0, 0)
                else:
                    whole = t0.new_empty(0).set_(t0.untyped_storage(), t0.storage_offset(), (n * t0.numel(),), (1,))
                    ops.stats_update_segmented_impl(whole, arena, first, n, t0.numel(), code)
            else:
                code, bw, sym, strict, unsigned, percentile = run[5]
                slot = outs[arena.device]
                ops.compute_encodings_into(arena, first, n, code, bw, sym, strict, unsigned, slot[0][slot[1]:slot[1] + n],
                                           percentile=percentile)
                slot[2].extend(run[6])
                slot[1] += n
        for out, _, fills in outs.values():
            for enc, row in zip(fills, out.cpu().tolist()):       # the one read-back
                enc._fill(row)                                    # pylint: disable=protected-access


_TLS = threading.local()
_current_raw_stream = torch._C._cuda_getCurrentRawStream   # pylint: disable=protected-access  (0.1 us; no Stream object)


def _calls() -> _DeferredCalls:
    q = getattr(_TLS, "queue", None)
    if q is None:
        q = _TLS.queue = _DeferredCalls()
    return q


def flush_deferred_calls():
    """Issue whatever the calling thread's DeferredAimetTensorQuantizer objects have queued."""
    _calls().flush()


class _RecordPool:
    """Statistics records for objects that ask for them one at a time: carved 256 at a time (one reset launch per 256, not
    one per object), so that objects created together -- the channels of one weight -- sit on consecutive records."""
    BLOCK = 256
    _pools = {}
    _lock = threading.Lock()

    def __init__(self, device):
        self.device = device
        self.block, self.used = None, 0
        self.free = []

    @classmethod
    def take(cls, device):
        with cls._lock:
            pool = cls._pools.get(device)
            if pool is None:
                pool = cls._pools[device] = _RecordPool(device)
            if pool.free:
                block, index = pool.free.pop()
                recycled = True
            else:
                if pool.block is None or pool.used == cls.BLOCK:
                    pool.block, pool.used = StateArena.for_device(device).allocate(cls.BLOCK), 0
                block, index, recycled = pool.block, pool.used, False
                pool.used += 1
        if recycled:
            ops.stats_reset_impl(block.arena, block.first + index, 1)
        return block, index

    @classmethod
    def give_back(cls, block, index):
        pool = cls._pools.get(block.device)
        if pool is not None:
            pool.free.append((block, index))     # list.append is atomic; the record is reset when it is handed out again


class DeferredAimetTensorQuantizer(AimetTensorQuantizer):
    """AimetTensorQuantizer for callers that make one native call per channel -- the class `aimet_b200.install` registers
    for the reference's own Python. Same methods, same results; resetEncodingStats, updateStats on the slices of one
    tensor, and getEncoding are queued and issued in batches (see _DeferredCalls). The encodings getEncoding returns are
    complete TfEncoding objects from the caller's point of view: the first read of a field has the queue executed."""

    def _ensure_state(self, device):
        if self._block is None or self._block.device != device:
            self._release_record()
            self._block, self._index = _RecordPool.take(device)
            self._pooled = True
            self._clean = True        # handed out reset

    def _release_record(self):
        if getattr(self, "_pooled", False) and self._block is not None:
            _RecordPool.give_back(self._block, self._index)
            self._pooled = False

    def __del__(self):
        try:
            self._release_record()
        except Exception:   # pylint: disable=broad-except  (interpreter shutdown)
            pass

    def _bind(self, block, index, group=None):
        _calls().flush()
        self._release_record()
        super()._bind(block, index, group)

    def resetEncodingStats(self):
        self._reset_host_state()
        # a record that has not been touched since its last reset (or since the pool handed it out) is not reset again:
        # the reference's Python resets every channel up to four times per calibration job
        if self._block is not None and not getattr(self, "_clean", False):
            self._clean = True
            _calls().push_reset(self._block.arena, self._block.first + self._index, self)

    def updateStats(self, input, use_cuda):   # pylint: disable=redefined-builtin
        t, _ = _to_device_tensor(input)
        queue = _calls()
        deferrable = self._code != ops.QUANTIZATION_ENTROPY and t.is_contiguous() and t.numel() > 0 and \
            not torch.cuda.is_current_stream_capturing()
        if not deferrable:
            queue.flush()
            queue.last_update = None
            super().updateStats(t, use_cuda)
            self._clean = False
            return
        self._is_encoding_valid = True
        self._ensure_state(t.device)
        self._clean = False
        arena, rec = self._block.arena, self._block.first + self._index
        self._updates += 1
        nbytes = t.numel() * t.element_size()
        where = (t.data_ptr(), nbytes, t.untyped_storage().data_ptr(), t.dtype)
        last = queue.last_update
        queue.last_update = (arena, rec, self._code, where, t)      # (t: keeps the storage, hence the pointers, alive)
        # A call is queued only when it CONTINUES the previous one -- the next record, the next slice of the same storage:
        # the caller is walking over the channels of one tensor and control has not gone back to the model in between, so
        # nothing can have written to that tensor. A call that starts something new (an activation: the model may overwrite
        # it in place as soon as this method returns) is issued at once.
        if last is not None and last[0] is arena and last[1] + 1 == rec and last[2] == self._code and \
                where == (last[3][0] + nbytes, last[3][1], last[3][2], last[3][3]):
            queue.push_update(arena, rec, self, self._code, where, t)
            return
        queue.flush()
        ops.stats_update_impl(t, arena, rec, self._code, None, 0, ops.STATS_RANGE_FIXED if self._range_fixed else 0)
        if ops.keeps_histogram(self._code) and not self._range_fixed and self._updates >= 2:
            self._poll_range_fixed()

    def getEncoding(self, bitwidth, use_symmetric_encodings, use_strict_symmetric, use_unsigned_symmetric):
        if not self._is_encoding_valid or self._block is None:
            return libpymo.TfEncoding(), self._is_encoding_valid
        queue = _calls()
        if self._code == ops.QUANTIZATION_ENTROPY or torch.cuda.is_current_stream_capturing():
            queue.flush()
            return super().getEncoding(bitwidth, use_symmetric_encodings, use_strict_symmetric, use_unsigned_symmetric)
        if use_symmetric_encodings and self._code == ops.QUANTIZATION_TF:
            assert not (use_strict_symmetric and use_unsigned_symmetric)   # TfEncodingAnalyzer.cpp:85-86
        enc = libpymo.TfEncoding._deferred(queue)   # pylint: disable=protected-access
        key = (self._code, int(bitwidth), bool(use_symmetric_encodings), bool(use_strict_symmetric),
               bool(use_unsigned_symmetric), self._percentile)
        queue.push_search(self._block.arena, self._block.first + self._index, self, key, enc)
        return enc, True

    # everything else reads encodings (which flushes by itself) or the record: issue what is queued first
    def quantizeDequantize(self, input, encoding, rounding_mode, use_cuda):   # pylint: disable=redefined-builtin
        _calls().flush()
        return super().quantizeDequantize(input, encoding, rounding_mode, use_cuda)

    def quantize(self, input, encoding, rounding_mode, use_cuda, shift_to_signed):   # pylint: disable=redefined-builtin
        _calls().flush()
        return super().quantize(input, encoding, rounding_mode, use_cuda, shift_to_signed)

    def quantizeDequantizePerChannel(self, input, encodings, num_channel, num_element, num_element_per_channel,   # pylint: disable=redefined-builtin
                                     rounding_mode, use_cuda):
        _calls().flush()
        return super().quantizeDequantizePerChannel(input, encodings, num_channel, num_element, num_element_per_channel,
                                                    rounding_mode, use_cuda)

    def getStatsHistogram(self):
        _calls().flush()
        return super().getStatsHistogram()


DEFER_DROPIN = os.environ.get("AB_DEFER_DROPIN", "1") != "0"

_SEED = [0]


def _next_seed():
    _SEED[0] = (_SEED[0] + 0x9E3779B97F4A7C15) & (2**63 - 1)
    return _SEED[0]


def _restore(out, original, orig_device):
    """Give the result the dtype/device the reference would: same dtype as the (fp32) input, same device."""
    if out.dtype != original.dtype and original.dtype in (torch.float16, torch.float64):
        out = out.to(original.dtype)
    if out.device != orig_device:
        out = out.to(orig_device)
    return out
