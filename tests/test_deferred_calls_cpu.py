"""Host logic of the deferred drop-in's call queue (aimet_b200.tensor_quantizer_op._DeferredCalls) without a GPU: the
native entry points are replaced by a recorder, the queue is fed the entries DeferredAimetTensorQuantizer would push, and
what must hold is checked -- calls on a record are issued in their original order, consecutive calls of one kind on
consecutive records become one launch, slices are merged only when they are adjacent in one storage, every owed encoding is
filled from its own row, a failure marks what is still owed. (The arithmetic itself: tests/test_gpu_python_api.py.)"""
import pytest
import torch

from aimet_b200 import libpymo
from aimet_b200 import tensor_quantizer_op as atq


class _Recorder:
    LOG_WORDS = 514

    def __init__(self):
        self.calls = []

    def stats_reset_impl(self, arena, first, count):
        self.calls.append(("reset", first, count))

    def stats_update_impl(self, t, arena, rec, code, log, entry, flags=0):
        self.calls.append(("update", rec, t.numel()))

    def stats_update_segmented_impl(self, whole, arena, first, n, seg_len, code):
        assert whole.numel() == n * seg_len
        self.calls.append(("segmented", first, n, seg_len, whole.clone()))

    def compute_encodings_into(self, arena, first, count, code, bw, sym, strict, unsigned, out, percentile=None):
        self.calls.append(("search", first, count, bw, sym))
        for i in range(count):
            out[i] = torch.tensor([-(first + i), first + i, 0.5, -7.0, bw], dtype=torch.float64)


@pytest.fixture()
def queue(monkeypatch):
    rec = _Recorder()
    monkeypatch.setattr(atq, "ops", rec)
    monkeypatch.setattr(atq, "_current_raw_stream", lambda device_index: 0)
    return atq._DeferredCalls(), rec          # pylint: disable=protected-access


def _where(t):
    return (t.data_ptr(), t.numel() * t.element_size(), t.untyped_storage().data_ptr(), t.dtype)


def _update(q, arena, rec, t, code=1, where=None):
    q.push_update(arena, rec, None, code, where or _where(t), t)


def test_runs_are_coalesced_and_order_is_kept(queue):
    q, rec = queue
    arena, other = torch.zeros(1), torch.zeros(1)
    w = torch.arange(40, dtype=torch.float32).view(5, 8)
    key = (1, 8, True, False, False, None)
    encs = [libpymo.TfEncoding._deferred(q) for _ in range(6)]       # pylint: disable=protected-access
    for r in (10, 11, 12, 20):
        q.push_reset(arena, r, None)
    q.push_reset(other, 21, None)                                     # another arena: its own launch
    _update(q, arena, 11, w[1])
    _update(q, arena, 12, w[2])
    _update(q, arena, 13, w[3])
    _update(q, arena, 14, w[0])                                       # next record, but not the next slice
    q.push_search(arena, 10, None, key, encs[0])
    q.push_search(arena, 11, None, key, encs[1])
    q.push_search(arena, 12, None, (1, 4, True, False, False, None), encs[2])   # other arguments: its own launch
    q.push_reset(arena, 10, None)
    q.push_search(arena, 10, None, key, encs[3])
    q.push_search(arena, 30, None, key, encs[4])
    q.push_search(arena, 31, None, key, encs[5])
    assert len(q.runs) == 10 and rec.calls == []
    q.flush()
    kinds = [c[:3] if c[0] != "segmented" else c[:4] for c in rec.calls]
    assert kinds == [("reset", 10, 3), ("reset", 20, 1), ("reset", 21, 1), ("segmented", 11, 3, 8), ("update", 14, 8),
                     ("search", 10, 2), ("search", 12, 1), ("reset", 10, 1), ("search", 10, 1), ("search", 30, 2)]
    assert torch.equal(rec.calls[3][4], w[1:4].reshape(-1))          # the merged view covers exactly the three slices
    # every encoding got its own row, whichever launch computed it
    assert [(e.min, e.max, e.bw) for e in encs] == [(-10, 10, 8), (-11, 11, 8), (-12, 12, 4), (-10, 10, 8), (-30, 30, 8),
                                                   (-31, 31, 8)]
    assert all(e._lazy is None for e in encs) and not q.runs         # pylint: disable=protected-access


def test_reading_an_owed_encoding_runs_the_queue(queue):
    q, rec = queue
    arena = torch.zeros(1)
    e = libpymo.TfEncoding._deferred(q)                              # pylint: disable=protected-access
    q.push_reset(arena, 3, None)
    q.push_search(arena, 3, None, (1, 8, False, False, False, None), e)
    assert rec.calls == []
    assert e.offset == -7.0                                          # first read of any field
    assert [c[0] for c in rec.calls] == ["reset", "search"] and e.delta == 0.5


def test_adjacent_addresses_in_different_storages_are_not_merged(queue):
    q, rec = queue
    arena = torch.zeros(1)
    a, b = torch.zeros(8), torch.zeros(8)
    _update(q, arena, 5, a)
    _update(q, arena, 6, b, where=(_where(a)[0] + 32, 32, b.untyped_storage().data_ptr(), b.dtype))   # adjacent address
    q.flush()
    assert [c[0] for c in rec.calls] == ["update", "update"]


def test_a_stream_change_or_a_long_queue_flushes(queue, monkeypatch):
    q, rec = queue
    arena = torch.zeros(1)
    q.push_reset(arena, 0, None)
    monkeypatch.setattr(atq, "_current_raw_stream", lambda device_index: 7)
    q.push_reset(arena, 1, None)                                     # asked for on another stream: what was queued goes first
    assert rec.calls == [("reset", 0, 1)] and len(q.runs) == 1
    monkeypatch.setattr(q, "LIMIT", 4)
    for r in range(2, 8):
        q.push_reset(arena, r, None)
    assert rec.calls[1] == ("reset", 1, 4) and q.runs[0][2:4] == [5, 3]     # flushed when the 5th call came in


def test_a_tensor_written_before_the_flush_is_an_error_and_marks_what_is_owed(queue):
    q, rec = queue
    arena = torch.zeros(1)
    w = torch.zeros(3, 8)
    e = libpymo.TfEncoding._deferred(q)                              # pylint: disable=protected-access
    _update(q, arena, 0, w[0])
    _update(q, arena, 1, w[1])
    q.push_search(arena, 0, None, (1, 8, False, False, False, None), e)
    w.add_(1.0)
    with pytest.raises(RuntimeError, match="modified in place"):
        q.flush()
    assert rec.calls == [] and not q.runs
    with pytest.raises(RuntimeError, match="deferred native calls"):
        _ = e.min
