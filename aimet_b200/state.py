"""Device-resident statistics records (ab_stats_state) and the arena that hands them out.

The reference keeps one encoding-analyzer object per quantizer -- and one per *channel* for per-channel weights
(aimet_torch/v1/tensor_quantizer.py:424,506: 26 560 of them for ResNet-50) -- in host memory. Here the records live
in HBM, packed in large uint8 tensors, so that statistics kernels update them in place with no host round trip and a
per-channel quantizer owns one contiguous block that a single segmented launch can update.
"""
import collections
import threading

import numpy as np
import torch

from . import ops

PDF_SIZE = 512

# mirrors `ab_stats_state` in include/aimet_b200.h
STATE_DTYPE = np.dtype([
    ("pdf", "<f8", PDF_SIZE), ("hist", "<u4", (2, PDF_SIZE)), ("x_left0", "<f8"), ("bucket_size_d", "<f8"),
    ("run_min", "<f8"), ("run_max", "<f8"), ("pending_count", "<f8"), ("bucket_size", "<f4"), ("pdf_offset", "<f4"),
    ("batch_min_bits", "<i4"), ("batch_max_bits", "<i4"), ("initialized", "<i4"), ("stats_updated", "<i4"),
    ("iterations", "<i4"), ("ticket", "<u4"), ("pending", "<i4"), ("write_parity", "<i4"), ("bf16_scale", "<f4"), ("bf16_shift", "<f4"), ("bf16_formula", "<i4"), ("bf16_fail_mask", "<u4")])
assert STATE_DTYPE.itemsize == ops.STATE_BYTES, (STATE_DTYPE.itemsize, ops.STATE_BYTES)


def field_index(name: str, word_bytes: int) -> int:
    """Index of a scalar field inside one record viewed as an array of `word_bytes`-wide words."""
    offset = STATE_DTYPE.fields[name][1]
    assert offset % word_bytes == 0
    return offset // word_bytes


def fold_pending(records: np.ndarray) -> np.ndarray:
    """Host twin of the device-side lazy fold (ab_stats_state.pending): returns records in which a parked batch has been
    folded into the PDF with the reference's formula pdf = (pdf*k + hist/cnt)/(k+1) (math_functions.cpp:279-287)."""
    out = records.copy()
    for r in out:
        if r["pending"]:
            pp = int(r["write_parity"]) ^ 1
            k = int(r["iterations"])
            prob = r["hist"][pp].astype(np.float64) / np.float64(r["pending_count"])
            r["pdf"] = (r["pdf"] * np.float64(k) + prob) / np.float64(k + 1)
            r["hist"][pp] = 0
            r["iterations"] = k + 1
            r["pending"] = 0
    return out


class StateBlock:
    """`count` consecutive records inside one arena chunk."""

    def __init__(self, owner, arena, first, count):
        self._owner = owner
        self.arena = arena          # uint8 CUDA tensor (the whole chunk)
        self.first = first
        self.count = count

    @property
    def device(self):
        return self.arena.device

    def reset(self):
        ops.stats_reset_impl(self.arena, self.first, self.count)

    def bytes_view(self) -> torch.Tensor:
        return self.arena[self.first * ops.STATE_BYTES:(self.first + self.count) * ops.STATE_BYTES]

    def read_raw(self) -> np.ndarray:
        """Synchronising copy of the records to the host, as a structured numpy array, exactly as they sit in HBM."""
        return self.bytes_view().cpu().numpy().view(STATE_DTYPE)

    def read(self) -> np.ndarray:
        """Like read_raw, with any batch still parked in `hist` folded into the PDF (what the reference's state holds)."""
        return fold_pending(self.read_raw())

    def write(self, records: np.ndarray):
        self.bytes_view().copy_(torch.from_numpy(records.view(np.uint8).reshape(-1)))

    def histogram(self, i=0):
        """[(xLeft, pdf)] * 512 as IQuantizationEncodingAnalyzer::getStatsHistogram returns it; [] if no range yet."""
        rec = self.read()[i]
        if not rec["initialized"]:
            return []
        x_left = rec["x_left0"] + np.arange(PDF_SIZE, dtype=np.float64) * rec["bucket_size_d"]
        return list(zip(x_left.tolist(), rec["pdf"].tolist()))

    def __del__(self):
        owner = getattr(self, "_owner", None)
        if owner is not None:
            owner._release(self)


class StateArena:
    """Per-device pool of statistics records."""
    CHUNK_STATES = 1024
    _arenas = {}
    # Re-entrant: the cyclic garbage collector may run inside a locked region and finalise StateBlocks on this very thread.
    # (StateBlock.__del__ does not take the lock at all -- see `_released` -- this is the second line of defence.)
    _lock = threading.RLock()

    def __init__(self, device):
        self.device = torch.device(device)
        self._chunk = None
        self._used = 0
        self._free = {}        # count -> [(chunk, first)]
        self._released = collections.deque()   # blocks given back by finalisers: appended lock-free, drained by allocate()

    @classmethod
    def for_device(cls, device) -> "StateArena":
        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError("aimet_b200 statistics live in CUDA memory: there is no CPU path")
        if device.index is None:
            device = torch.device("cuda", torch.cuda.current_device())
        with cls._lock:
            arena = cls._arenas.get(device)
            if arena is None:
                arena = cls._arenas[device] = StateArena(device)
            return arena

    def _carve(self, count: int):
        """(chunk, first) from the free lists or the current chunk, or None when a new chunk is needed. Lock held."""
        while self._released:
            try:
                n, arena, first = self._released.popleft()
            except IndexError:
                break
            self._free.setdefault(n, []).append((arena, first))
        free = self._free.get(count)
        if free:
            return free.pop()
        cap = 0 if self._chunk is None else self._chunk.numel() // ops.STATE_BYTES
        if self._chunk is not None and self._used + count <= cap:
            first = self._used
            self._used += count
            return self._chunk, first
        return None

    def allocate(self, count: int) -> StateBlock:
        with self._lock:
            got = self._carve(count)
        if got is None:
            # the device allocation runs OUTSIDE the lock (it can trigger the garbage collector, synchronise, or raise)
            fresh = torch.empty(max(count, self.CHUNK_STATES) * ops.STATE_BYTES, dtype=torch.uint8, device=self.device)
            with self._lock:
                got = self._carve(count)           # somebody else may have installed a chunk meanwhile
                if got is None:
                    self._chunk, self._used = fresh, count
                    got = (fresh, 0)
        chunk, first = got
        block = StateBlock(self, chunk, first, count)
        block.reset()
        return block

    def _release(self, block):
        # called from StateBlock.__del__, possibly by the garbage collector in the middle of allocate(): no lock here.
        # deque.append is atomic; allocate() files the entry under the lock.
        try:
            self._released.append((block.count, block.arena, block.first))
        except Exception:   # interpreter shutdown
            pass
