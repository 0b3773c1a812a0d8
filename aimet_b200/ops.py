"""torch custom ops (namespace ``aimet_b200``) over the C ABI in include/aimet_b200.h.

Every op hands raw device pointers and the current CUDA stream to libaimet_b200.so; torch is used only for device
memory, streams and autograd plumbing. There is no CPU implementation registered for any of them: calling an op with
a CPU tensor raises (the class-level API in tensor_quantizer_op.py stages host tensors onto the GPU first).

Functional surface (schema strings below); autograd is registered for the two QDQ forwards with the reference's
straight-through estimator (aimet_torch/v1/quantsim_straight_through_grad.py:91-118) as their backward.
"""
import ctypes as C

import torch

from . import _lib

ROUND_NEAREST, ROUND_STOCHASTIC = 0, 1
QUANTIZATION_TF, QUANTIZATION_TF_ENHANCED, QUANTIZATION_PERCENTILE, QUANTIZATION_MSE = 0, 1, 3, 4
QUANTIZATION_ENTROPY = 5   # its own entry points (ab_entropy_*): a histogram whose range grows, KL search on the host


def keeps_histogram(quant_mode) -> bool:
    """tf_enhanced, percentile and mse all keep the 512-bin running PDF (UpdatePdf); tf keeps a running min/max."""
    return int(quant_mode) in (QUANTIZATION_TF_ENHANCED, QUANTIZATION_PERCENTILE, QUANTIZATION_MSE)

STATE_BYTES = None          # filled at import from ab_stats_state_bytes()
LOG_WORDS = _lib.PDF_SIZE + 2

_L = _lib.load()
STATE_BYTES = int(_L.ab_stats_state_bytes())

# kernel launches issued through this module, by kernel family (bench.py reports their sum as gpu_launches)
LAUNCHES = {"qdq": 0, "quantize": 0, "qdq_per_channel": 0, "ste_bwd": 0, "minmax": 0, "hist": 0, "segmented": 0,
            "search": 0, "reset": 0, "init_range": 0, "fold": 0, "lg_fwd": 0, "lg_bwd": 0, "qdq_broadcast": 0,
            "hist_multi": 0, "fold_segments": 0, "entropy": 0}
# when a list, stats_update_impl brackets its launches with CUDA events and appends (bytes, start, stop, quant_mode)
STATS_TIMING = None
_EVENT_POOL = []


def reserve_timing_events(n: int):
    """Pre-create CUDA events so that timing a launch costs two cudaEventRecord calls and nothing else."""
    while len(_EVENT_POOL) < n:
        _EVENT_POOL.append(torch.cuda.Event(enable_timing=True))


def _timing_event():
    if torch.cuda.is_current_stream_capturing():
        # inside a CUDA-graph capture: an EXTERNAL event becomes an event-record node, re-recorded by every replay
        return torch.cuda.Event(enable_timing=True, external=True)
    return _EVENT_POOL.pop() if _EVENT_POOL else torch.cuda.Event(enable_timing=True)


def launches_total() -> int:
    return sum(LAUNCHES.values())


def _dtype_code(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return _lib.AB_F32
    if t.dtype == torch.bfloat16:
        return _lib.AB_BF16
    raise TypeError(f"aimet_b200 kernels take float32 or bfloat16 tensors, got {t.dtype}")


def _require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("aimet_b200 has no CPU path: tensor must live on a CUDA device "
                               "(reference analogue: 'Not compiled for GPU mode.' in the other direction)")


_raw_stream = torch._C._cuda_getCurrentRawStream   # pylint: disable=protected-access
_current_device = torch._C._cuda_getDevice         # pylint: disable=protected-access


def _stream(t: torch.Tensor) -> int:
    """The raw cudaStream_t torch currently uses on `t`'s device. torch.cuda.current_stream() builds a Python Stream object on
    every call (8.6 us measured): the QAT step makes 444 such calls (tools/host_profile_qat.py)."""
    return _raw_stream(t.device.index)


class _on_device:
    """Make `t`'s device current for the duration of a C-ABI call (cheap no-op when it already is)."""

    def __init__(self, t):
        self.idx = t.device.index
        self.prev = None

    def __enter__(self):
        cur = _current_device()
        if cur != self.idx:
            self.prev = cur
            torch.cuda.set_device(self.idx)

    def __exit__(self, *exc):
        if self.prev is not None:
            torch.cuda.set_device(self.prev)


def _contig16(t: torch.Tensor) -> torch.Tensor:
    """Contiguous and 16-byte aligned (the per-channel / STE kernels require it)."""
    t = t.contiguous()
    if t.data_ptr() % 16:
        t = t.clone(memory_format=torch.contiguous_format)
    return t


# ---------------------------------------------------------------------------------------------------------------------
# implementations (CUDA)
# ---------------------------------------------------------------------------------------------------------------------
def qdq_per_tensor_impl(x, enc_min, enc_max, bw, round_mode=0, seed=0):
    _require_cuda(x)
    x = x.contiguous(memory_format=torch.contiguous_format) if not (
        x.is_contiguous() or x.is_contiguous(memory_format=torch.channels_last)) else x
    out = torch.empty_like(x)
    with _on_device(x):
        _lib.check(_L.ab_qdq_per_tensor_fwd(x.data_ptr(), out.data_ptr(), x.numel(), _dtype_code(x), float(enc_min),
                                            float(enc_max), int(bw), int(round_mode), int(seed) & (2**64 - 1),
                                            _stream(x)))
    LAUNCHES["qdq"] += 1
    return out


def qdq_per_tensor_dev_impl(x, enc4, round_mode=0, seed=0):
    _require_cuda(x, enc4)
    if enc4.dtype != torch.float32 or enc4.numel() < 4:
        raise ValueError("enc4 must be a float32 CUDA tensor {min, max, delta, offset}")
    x = x if (x.is_contiguous() or x.is_contiguous(memory_format=torch.channels_last)) else x.contiguous()
    out = torch.empty_like(x)
    with _on_device(x):
        _lib.check(_L.ab_qdq_per_tensor_fwd_dev(x.data_ptr(), out.data_ptr(), x.numel(), _dtype_code(x),
                                                enc4.data_ptr(), int(round_mode), int(seed) & (2**64 - 1),
                                                _stream(x)))
    LAUNCHES["qdq"] += 1
    return out


def quantize_to_grid_impl(x, enc_min, enc_max, bw, round_mode=0, shift_to_signed=False, seed=0):
    _require_cuda(x)
    x = x if (x.is_contiguous() or x.is_contiguous(memory_format=torch.channels_last)) else x.contiguous()
    out = torch.empty_like(x)
    with _on_device(x):
        _lib.check(_L.ab_quantize_to_grid(x.data_ptr(), out.data_ptr(), x.numel(), _dtype_code(x), float(enc_min),
                                          float(enc_max), int(bw), int(round_mode), int(bool(shift_to_signed)),
                                          int(seed) & (2**64 - 1), _stream(x)))
    LAUNCHES["quantize"] += 1
    return out


def quantize_to_packed_impl(x, enc_min, enc_max, bw, shift_to_signed=False):
    """quantizeTensorPacked: the integer grid values as a flat uint8 CUDA tensor of max(bw, 8) / 8 bytes per element (view it
    as int8 / uint16 / int16 / int32 ... with `.view(dtype)`), as the reference fills its std::vector<uint8_t>."""
    _require_cuda(x)
    x = x.contiguous()
    out = torch.empty(x.numel() * max(int(bw), 8) // 8, dtype=torch.uint8, device=x.device)
    with _on_device(x):
        _lib.check(_L.ab_quantize_to_packed(x.data_ptr(), out.data_ptr(), x.numel(), _dtype_code(x), float(enc_min),
                                            float(enc_max), int(bw), int(bool(shift_to_signed)), _stream(x)))
    LAUNCHES["quantize"] += 1
    return out


def qdq_per_channel_impl(x, params, num_channel, num_element_per_channel, round_mode=0, seed=0):
    _require_cuda(x, params)
    if params.dtype != torch.float32 or params.numel() != 4 * num_channel or not params.is_contiguous():
        raise ValueError("params must be a contiguous float32 CUDA tensor of 4*num_channel values")
    x = _contig16(x)
    out = torch.empty_like(x, memory_format=torch.contiguous_format)
    with _on_device(x):
        _lib.check(_L.ab_qdq_per_channel_fwd(x.data_ptr(), out.data_ptr(), int(num_channel), x.numel(),
                                             int(num_element_per_channel), _dtype_code(x), params.data_ptr(),
                                             int(round_mode), int(seed) & (2**64 - 1), _stream(x)))
    LAUNCHES["qdq_per_channel"] += 1
    return out


def qdq_broadcast_impl(x, enc_min, enc_max, enc_delta, enc_offset):
    """QDQ of `x` with an encoding tensor that broadcasts against it (torch broadcasting rules: the four encoding tensors
    share one shape, which is x's shape with some dimensions of size 1, possibly with fewer leading dimensions). The
    encodings are used as they are -- the reference's quantizeDequantizeBroadcast does no gating -- rounding to nearest."""
    _require_cuda(x, enc_min, enc_max, enc_delta, enc_offset)
    shape = tuple(enc_min.shape)
    for e in (enc_min, enc_max, enc_delta, enc_offset):
        if e.dtype != torch.float32 or tuple(e.shape) != shape or not e.is_contiguous():
            raise ValueError("the four encoding tensors must be contiguous float32 tensors of one shape")
    if len(shape) > x.dim():
        raise ValueError("the encoding tensor has more dimensions than the input")
    x = x.contiguous()
    nd = max(x.dim(), 1)
    xs = tuple(x.shape) if x.dim() else (1,)
    padded = (1,) * (nd - len(shape)) + shape
    for a, b in zip(padded, xs):
        if a not in (1, b):
            raise ValueError(f"encoding shape {shape} does not broadcast to input shape {tuple(x.shape)}")
    in_strides, enc_strides = [0] * nd, [0] * nd
    acc_in = acc_enc = 1
    for d in range(nd - 1, -1, -1):
        in_strides[d] = acc_in
        enc_strides[d] = acc_enc if padded[d] != 1 else 0
        acc_in *= xs[d]
        acc_enc *= padded[d]
    out = torch.empty_like(x)
    arr = C.c_int64 * nd
    with _on_device(x):
        _lib.check(_L.ab_qdq_broadcast_fwd(x.data_ptr(), out.data_ptr(), x.numel(), nd, arr(*in_strides), arr(*enc_strides),
                                           enc_min.data_ptr(), enc_max.data_ptr(), enc_delta.data_ptr(),
                                           enc_offset.data_ptr(), _dtype_code(x), _stream(x)))
    LAUNCHES["qdq_broadcast"] += 1
    return out


def per_channel_params_dev(enc5: torch.Tensor, bw: int) -> torch.Tensor:
    """[4*C] fp32 parameter block on the device from a device [C,5] float64 encoding table (no host round trip)."""
    _require_cuda(enc5)
    if enc5.dtype != torch.float64 or enc5.dim() != 2 or enc5.shape[1] != 5 or not enc5.is_contiguous():
        raise ValueError("enc5 must be a contiguous float64 CUDA tensor [C, 5]")
    c = enc5.shape[0]
    params = torch.empty(4 * c, dtype=torch.float32, device=enc5.device)
    with _on_device(enc5):
        _lib.check(_L.ab_per_channel_params_dev(enc5.data_ptr(), c, int(bw), params.data_ptr(), _stream(enc5)))
    LAUNCHES["search"] += 0
    return params


def ste_bwd_impl(x, grad, enc_min, enc_max):
    _require_cuda(x, grad)
    if x.dtype != grad.dtype or x.shape != grad.shape:
        raise ValueError("x and grad must share dtype and shape")
    if x.is_contiguous(memory_format=torch.channels_last) and grad.is_contiguous(memory_format=torch.channels_last) \
            and not x.is_contiguous():
        pass   # same physical layout on both: element-wise is layout agnostic
    else:
        x, grad = x.contiguous(), grad.contiguous()
    if x.data_ptr() % 16:
        x = x.clone()
    if grad.data_ptr() % 16:
        grad = grad.clone()
    out = torch.empty_like(grad)
    with _on_device(x):
        _lib.check(_L.ab_qdq_ste_bwd(x.data_ptr(), grad.data_ptr(), out.data_ptr(), x.numel(), _dtype_code(x),
                                     float(enc_min), float(enc_max), _stream(x)))
    LAUNCHES["ste_bwd"] += 1
    return out


def ste_bwd_enc5_impl(x, grad, enc5, num_channel, num_element_per_channel, range_in_bf16=False):
    """STE backward with the range read from device-resident encoding rows ([num_channel, 5] float64, as
    compute_encodings_impl returns them): no slicing / casting kernels on the host side."""
    _require_cuda(x, grad, enc5)
    if x.dtype != grad.dtype or x.shape != grad.shape:
        raise ValueError("x and grad must share dtype and shape")
    if enc5.dtype != torch.float64 or enc5.numel() != 5 * num_channel or not enc5.is_contiguous():
        raise ValueError("enc5 must be a contiguous float64 CUDA tensor [num_channel, 5]")
    x, grad = _contig16(x), _contig16(grad)
    out = torch.empty_like(grad, memory_format=torch.contiguous_format)
    with _on_device(x):
        _lib.check(_L.ab_qdq_ste_bwd_enc5(x.data_ptr(), grad.data_ptr(), out.data_ptr(), int(num_channel), x.numel(),
                                          int(num_element_per_channel), _dtype_code(x), enc5.data_ptr(),
                                          1 if range_in_bf16 else 0, _stream(x)))
    LAUNCHES["ste_bwd"] += 1
    return out


def ste_bwd_per_channel_impl(x, grad, enc_min, enc_max, num_channel, num_element_per_channel):
    _require_cuda(x, grad, enc_min, enc_max)
    if x.dtype != grad.dtype or x.shape != grad.shape:
        raise ValueError("x and grad must share dtype and shape")
    for e in (enc_min, enc_max):
        if e.dtype != torch.float32 or e.numel() != num_channel or not e.is_contiguous():
            raise ValueError("per-channel min/max must be contiguous float32 CUDA tensors of num_channel values")
    x, grad = _contig16(x), _contig16(grad)
    out = torch.empty_like(grad, memory_format=torch.contiguous_format)
    with _on_device(x):
        _lib.check(_L.ab_qdq_ste_bwd_per_channel(x.data_ptr(), grad.data_ptr(), out.data_ptr(), int(num_channel),
                                                 x.numel(), int(num_element_per_channel), _dtype_code(x),
                                                 enc_min.data_ptr(), enc_max.data_ptr(), _stream(x)))
    LAUNCHES["ste_bwd"] += 1
    return out


def _state_ptr(states: torch.Tensor, index: int) -> int:
    if states.dtype != torch.uint8 or not states.is_contiguous():
        raise ValueError("state arena must be a contiguous uint8 CUDA tensor")
    return states.data_ptr() + index * STATE_BYTES


def stats_reset_impl(states, first, count):
    _require_cuda(states)
    with _on_device(states):
        _lib.check(_L.ab_stats_reset(_state_ptr(states, first), int(count), _stream(states)))
    LAUNCHES["reset"] += 1


STATS_RANGE_FIXED = 1


def stats_update_impl(x, states, index, quant_mode, batch_log, log_entry, flags=0):
    """One updateStats call for the quantizer whose record is `states[index]`. flags=STATS_RANGE_FIXED: the caller has
    read back that the histogram range is fixed, so the (immediately exiting) min/max launch is skipped."""
    _require_cuda(x, states)
    x = x if (x.is_contiguous() or x.is_contiguous(memory_format=torch.channels_last)) else x.contiguous()
    log_ptr = None
    if batch_log is not None:
        _require_cuda(batch_log)
        if batch_log.dtype != torch.int32 and batch_log.dtype != torch.uint32:
            raise ValueError("batch_log must be a 32-bit integer CUDA tensor")
        log_ptr = batch_log.data_ptr() + int(log_entry) * LOG_WORDS * 4
    timing = STATS_TIMING
    with _on_device(x):
        if timing is not None:
            start, stop = _timing_event(), _timing_event()
            start.record()
        _lib.check(_L.ab_stats_update(x.data_ptr(), x.numel(), _dtype_code(x), int(quant_mode),
                                      _state_ptr(states, index), log_ptr, int(flags), _stream(x)))
        if timing is not None:
            stop.record()
            timing.append((x.numel() * x.element_size(), start, stop, int(quant_mode),
                           torch.cuda.is_current_stream_capturing()))
    if keeps_histogram(quant_mode):
        LAUNCHES["hist"] += 1
        if not flags & STATS_RANGE_FIXED:
            LAUNCHES["minmax"] += 1
    else:
        LAUNCHES["minmax"] += 1


def stats_update_segmented_impl(x, states, first, num_segments, segment_len, quant_mode):
    _require_cuda(x, states)
    x = x.contiguous()
    if x.numel() != num_segments * segment_len:
        raise ValueError("tensor size does not match num_segments * segment_len")
    with _on_device(x):
        _lib.check(_L.ab_stats_update_segmented(x.data_ptr(), int(num_segments), int(segment_len), _dtype_code(x),
                                                int(quant_mode), _state_ptr(states, first), _stream(x)))
    LAUNCHES["segmented"] += 1


MULTI_MAX_SEGMENTS = _lib.STATS_MULTI_MAX_SEGMENTS
# when a list, stats_update_multi_impl brackets its histogram + fold launches with CUDA events and appends
# (bytes, start, stop, num_segments, captured)
MULTI_TIMING = None


def stats_update_multi_impl(tensors, state_indices, states, first, seg_counts, log_only=False):
    """tf_enhanced updateStats for up to MULTI_MAX_SEGMENTS (tensor, record) pairs in one histogram launch + one fold launch
    (ab_stats_update_multi). `tensors`: dense CUDA tensors of ONE dtype (float32 or bfloat16), 16-byte aligned; tensor k
    updates record `first + state_indices[k]` of the arena `states`; `seg_counts`: zeroed int32 CUDA tensor with at least
    len(tensors) rows of LOG_WORDS words (scratch, or -- `log_only` -- the log entries of the calls)."""
    n = len(tensors)
    if n == 0:
        return
    if n > MULTI_MAX_SEGMENTS:
        raise ValueError(f"at most {MULTI_MAX_SEGMENTS} tensors per multi-tensor statistics call")
    _require_cuda(states, seg_counts, *tensors)
    if seg_counts.dtype not in (torch.int32, torch.uint32) or seg_counts.numel() < n * LOG_WORDS or \
            not seg_counts.is_contiguous():
        raise ValueError("seg_counts must be a contiguous 32-bit integer CUDA tensor of len(tensors) x LOG_WORDS words")
    code = _dtype_code(tensors[0])
    segs = (_lib.StatsSegment * n)()
    nbytes = 0
    for k, (t, idx) in enumerate(zip(tensors, state_indices)):
        if t.dtype != tensors[0].dtype:
            raise TypeError("all tensors of one multi-tensor statistics call must share a dtype")
        seg = segs[k]
        seg.data, seg.count, seg.state_index = t.data_ptr(), t.numel(), idx
        nbytes += t.numel() * t.element_size()
    timing = MULTI_TIMING
    flags = _lib.STATS_MULTI_LOG_ONLY if log_only else 0
    args = (segs, n, code, _state_ptr(states, first), seg_counts.data_ptr())
    with _on_device(states):
        if timing is None:
            _lib.check(_L.ab_stats_update_multi(*args, flags, _stream(states)))
        else:
            # the two launches enqueued separately, with a CUDA-event pair around the histogram launch alone
            start, stop = _timing_event(), _timing_event()
            start.record()
            _lib.check(_L.ab_stats_update_multi(*args, flags | 2, _stream(states)))
            stop.record()
            _lib.check(_L.ab_stats_update_multi(*args, flags | 4, _stream(states)))
            timing.append((nbytes, start, stop, n, torch.cuda.is_current_stream_capturing()))
    LAUNCHES["hist_multi"] += 1
    LAUNCHES["fold_segments"] += 1


def _search(states, first, count, quant_mode, bw, sym, strict, unsigned_sym, enc_ptr, qdq4_ptr, percentile):
    """ab_compute_encodings, or its percentile twin when `percentile` is given (the percentile scheme's statistics are
    the tf_enhanced ones; only the closing computation differs)."""
    flags = (int(bool(sym)), int(bool(strict)), int(bool(unsigned_sym)))
    if percentile is None:
        _lib.check(_L.ab_compute_encodings(_state_ptr(states, first), int(count), int(quant_mode), int(bw), *flags,
                                           enc_ptr, qdq4_ptr, _stream(states)))
    else:
        _lib.check(_L.ab_compute_encodings_percentile(_state_ptr(states, first), int(count), float(percentile), int(bw),
                                                      *flags, enc_ptr, qdq4_ptr, _stream(states)))
    LAUNCHES["search"] += 1


def entropy_update_impl(x, states, index):
    """EntropyEncodingAnalyzer::updateStats on record `index` (min / max, range growth, binning: three launches, no sync)."""
    _require_cuda(x)
    x = _contig16(x) if x.dtype in (torch.float32, torch.bfloat16) else x.float().contiguous()
    with _on_device(x):
        _lib.check(_L.ab_entropy_update(x.data_ptr(), x.numel(), _dtype_code(x), _state_ptr(states, index), _stream(x)))
    LAUNCHES["entropy"] += 3 if x.numel() else 1


def entropy_compute_impl(states, index, bw, sym, strict, unsigned_sym):
    """EntropyEncodingAnalyzer::computeEncoding: reads the histogram back (synchronises the stream), KL search on the host.
    Returns an _lib.Encoding."""
    e = _lib.Encoding()
    with _on_device(states):
        _lib.check(_L.ab_entropy_compute_encoding(_state_ptr(states, index), int(bw), int(bool(sym)), int(bool(strict)),
                                                  int(bool(unsigned_sym)), C.byref(e), _stream(states)))
    return e


def entropy_histogram_impl(states, index):
    """(histogram[512] float64 or None before the first non-zero batch, min, max, iterations) -- TensorProfilingParams."""
    import numpy as np
    hist, mm, info = np.zeros(_lib.PDF_SIZE), np.zeros(2), (C.c_int * 3)()
    with _on_device(states):
        _lib.check(_L.ab_entropy_histogram(_state_ptr(states, index), hist.ctypes.data_as(C.POINTER(C.c_double)),
                                           mm.ctypes.data_as(C.POINTER(C.c_double)), info, _stream(states)))
    return (hist if info[0] else None), float(mm[0]), float(mm[1]), int(info[2])


def compute_encodings_impl(states, first, count, quant_mode, bw, sym, strict, unsigned_sym, want_qdq4=False,
                           percentile=None):
    """Returns (enc[count,5] float64 CUDA tensor, qdq4[count,4] float32 CUDA tensor or None)."""
    _require_cuda(states)
    enc = torch.empty((count, 5), dtype=torch.float64, device=states.device)
    qdq4 = torch.empty((count, 4), dtype=torch.float32, device=states.device) if want_qdq4 else None
    with _on_device(states):
        _search(states, first, count, quant_mode, bw, sym, strict, unsigned_sym, enc.data_ptr(),
                qdq4.data_ptr() if want_qdq4 else None, percentile)
    return enc, qdq4


def stats_refresh_encodings_impl(x, states, first, num_segments, segment_len, quant_mode, bw, sym, strict, unsigned_sym):
    """reset + updateStats + computeEncoding for one tensor in one native call (see ab_stats_refresh_encodings).
    Returns (enc [n, 5] float64, qdq4 [1, 4] float32 or None, params [4 * n] float32 or None)."""
    _require_cuda(x, states)
    if x.numel() != num_segments * segment_len:
        raise ValueError("tensor size does not match num_segments * segment_len")
    per_tensor = num_segments == 1
    enc = torch.empty((num_segments, 5), dtype=torch.float64, device=states.device)
    qdq4 = torch.empty((1, 4), dtype=torch.float32, device=states.device) if per_tensor else None
    params = None if per_tensor else torch.empty(4 * num_segments, dtype=torch.float32, device=states.device)
    with _on_device(x):
        _lib.check(_L.ab_stats_refresh_encodings(x.data_ptr(), int(num_segments), int(segment_len), _dtype_code(x),
                                                 int(quant_mode), _state_ptr(states, first), int(bw), int(bool(sym)),
                                                 int(bool(strict)), int(bool(unsigned_sym)), enc.data_ptr(),
                                                 qdq4.data_ptr() if per_tensor else None,
                                                 None if per_tensor else params.data_ptr(), _stream(x)))
    LAUNCHES["reset"] += 1
    LAUNCHES["search"] += 1
    if per_tensor:
        LAUNCHES["minmax"] += 1
        if keeps_histogram(quant_mode):
            LAUNCHES["hist"] += 1
    else:
        LAUNCHES["segmented"] += 1
    return enc, qdq4, params


REFRESH_MAX_ITEMS = _lib.REFRESH_MULTI_MAX_ITEMS


def stats_refresh_multi_impl(tensors, num_segments, states, first, quant_mode, bw, sym, strict, unsigned_sym,
                             enc, qdq4=None, params=None, first_record=0):
    """reset + updateStats + computeEncoding (+ per-channel parameter blocks) for up to REFRESH_MAX_ITEMS parameter tensors
    in a handful of launches (ab_stats_refresh_encodings_multi). tensors[k]: contiguous CUDA tensor whose num_segments[k]
    equal segments update consecutive records, starting at `first + first_record` for k = 0 and following on from there.
    enc / qdq4 / params: the output tables for ALL records of the block ([N, 5] float64, [N, 4] float32, [4 N] float32);
    this call fills their rows from `first_record` on."""
    n = len(tensors)
    if n == 0:
        return
    if n > REFRESH_MAX_ITEMS:
        raise ValueError(f"at most {REFRESH_MAX_ITEMS} tensors per call")
    _require_cuda(states, enc, qdq4, params, *tensors)
    items = (_lib.RefreshItem * n)()
    at = 0
    for k, (t, segs) in enumerate(zip(tensors, num_segments)):
        if t.dtype != tensors[0].dtype or not t.is_contiguous() or t.numel() % segs or t.numel() == 0:
            raise ValueError("tensors must be contiguous, non-empty, of one dtype and divisible into their segments")
        it = items[k]
        it.data, it.num_segments, it.segment_len, it.first_record = t.data_ptr(), segs, t.numel() // segs, at
        at += segs
    with _on_device(states):
        _lib.check(_L.ab_stats_refresh_encodings_multi(
            items, n, _dtype_code(tensors[0]), int(quant_mode), _state_ptr(states, first + first_record), int(bw),
            int(bool(sym)), int(bool(strict)), int(bool(unsigned_sym)), enc.data_ptr() + first_record * 40,
            None if qdq4 is None else qdq4.data_ptr() + first_record * 16,
            None if params is None else params.data_ptr() + first_record * 16, _stream(states)))
    LAUNCHES["reset"] += 1
    LAUNCHES["segmented"] += 1
    LAUNCHES["search"] += 1
    return at


def compute_encodings_into(states, first, count, quant_mode, bw, sym, strict, unsigned_sym, out, percentile=None):
    """Same as compute_encodings_impl, but writes into `out` (float64 CUDA, [count, 5], contiguous): lets a caller
    enqueue many searches and read them back with one copy."""
    _require_cuda(states, out)
    if out.dtype != torch.float64 or out.numel() != 5 * count or not out.is_contiguous():
        raise ValueError("out must be a contiguous float64 CUDA tensor [count, 5]")
    with _on_device(states):
        _search(states, first, count, quant_mode, bw, sym, strict, unsigned_sym, out.data_ptr(), None, percentile)


def stats_init_range_impl(states, first, count, minmax):
    _require_cuda(states, minmax)
    if minmax.dtype != torch.float32 or minmax.numel() != 2 * count or not minmax.is_contiguous():
        raise ValueError("minmax must be a contiguous float32 CUDA tensor [count, 2]")
    with _on_device(states):
        _lib.check(_L.ab_stats_init_range(_state_ptr(states, first), int(count), minmax.data_ptr(),
                                          _stream(states)))
    LAUNCHES["init_range"] += 1


def stats_fold_batches_impl(states, first, count, batch_log, batch_offsets):
    _require_cuda(states, batch_log, batch_offsets)
    if batch_offsets.dtype != torch.int64 or not batch_offsets.is_contiguous():
        raise ValueError("batch_offsets must be a contiguous int64 CUDA tensor")
    with _on_device(states):
        _lib.check(_L.ab_stats_fold_batches(_state_ptr(states, first), int(count), batch_log.data_ptr(),
                                            batch_offsets.data_ptr(), batch_offsets.numel(), _stream(states)))
    LAUNCHES["fold"] += 1


def stats_fold_log_impl(states, first, count, log, entry_rows, record_begin):
    """Rebuild `count` records from per-call log rows: record s replays rows entry_rows[record_begin[s]:record_begin[s+1]]
    of `log` in that order (ab_stats_fold_log)."""
    _require_cuda(states, log, entry_rows, record_begin)
    for t in (entry_rows, record_begin):
        if t.dtype != torch.int64 or not t.is_contiguous():
            raise ValueError("entry_rows / record_begin must be contiguous int64 CUDA tensors")
    if record_begin.numel() != count + 1 or not log.is_contiguous():
        raise ValueError("record_begin must have count + 1 entries and the log must be contiguous")
    with _on_device(states):
        _lib.check(_L.ab_stats_fold_log(_state_ptr(states, first), int(count), log.data_ptr(), entry_rows.data_ptr(),
                                        record_begin.data_ptr(), _stream(states)))
    LAUNCHES["fold"] += 1


# ---------------------------------------------------------------------------------------------------------------------
# range learning (learned grid): fused forward / backward of the reference's QuantizeDequantizeFunc
# ---------------------------------------------------------------------------------------------------------------------
LG_ASYMMETRIC, LG_SIGNED_SYMMETRIC, LG_UNSIGNED_SYMMETRIC = 0, 1, 2
LG_GATE = 1
_LG_WORKSPACES = {}


def lg_symmetry_mode(use_symmetric_encodings: bool, is_unsigned_symmetric: bool) -> int:
    """The branch get_computed_encodings takes (reference v1/quantsim_straight_through_grad.py:145-158)."""
    if use_symmetric_encodings and not is_unsigned_symmetric:
        return LG_SIGNED_SYMMETRIC
    return LG_UNSIGNED_SYMMETRIC if use_symmetric_encodings else LG_ASYMMETRIC


def _lg_workspace(device, num_channel):
    """Zero-filled scratch, one per (device, stream): the kernels leave it zeroed, so it is reused launch after launch."""
    key = (device.index, _raw_stream(device.index))
    need = int(_L.ab_lg_workspace_bytes(int(num_channel)))
    ws = _LG_WORKSPACES.get(key)
    if ws is None or ws.numel() < need:
        ws = torch.zeros(max(need, 4096), dtype=torch.uint8, device=device)
        _LG_WORKSPACES[key] = ws
    return ws


def _lg_geometry(x, enc_min, ch_axis):
    c = enc_min.numel()
    if c == 1:
        return 1, 1, x.numel()
    if x.shape[ch_axis] != c:
        raise ValueError(f"encoding has {c} channels, tensor has {x.shape[ch_axis]} along axis {ch_axis}")
    outer = 1
    for d in x.shape[:ch_axis]:
        outer *= d
    inner = x.numel() // (outer * c) if outer * c else 0
    return outer, c, inner


def _lg_check(x, enc_min, enc_max, bw):
    _require_cuda(x, enc_min, enc_max)
    if int(bw) >= 32:
        raise RuntimeError(f"Invalid bitwidth: {bw}")   # calculate_forward_pass (:207-208)
    if not (x.dtype == enc_min.dtype == enc_max.dtype):
        # same message as calculate_forward_pass (:199-202)
        raise RuntimeError("Data type mismatch. Expected the input and encoding min & max to be of same dtype."
                           f"Got {x.dtype} input, {enc_min.dtype} encoding_min, and {enc_max.dtype} encoding_max")
    if enc_min.numel() != enc_max.numel() or not enc_min.is_contiguous() or not enc_max.is_contiguous():
        raise ValueError("encoding min / max must be contiguous tensors of equal length")


def lg_qdq_fwd_impl(x, enc_min, enc_max, bw, sym_mode, strict=False, ch_axis=0, gate=False):
    """Fused learned-grid forward. With gate=True, enc_min / enc_max are first clamped IN PLACE (apply_gating_logic)."""
    _lg_check(x, enc_min, enc_max, bw)
    per_tensor = enc_min.numel() == 1
    if not (x.is_contiguous() or (per_tensor and x.is_contiguous(memory_format=torch.channels_last))):
        x = x.contiguous()
    outer, c, inner = _lg_geometry(x, enc_min, ch_axis)
    out = torch.empty_like(x)
    with _on_device(x):
        ws = None if per_tensor else _lg_workspace(x.device, c).data_ptr()
        _lib.check(_L.ab_lg_qdq_fwd(x.data_ptr(), out.data_ptr(), outer, c, inner, _dtype_code(x), enc_min.data_ptr(),
                                    enc_max.data_ptr(), int(bw), int(sym_mode), int(bool(strict)),
                                    LG_GATE if gate else 0, ws, _stream(x)))
    LAUNCHES["lg_fwd"] += 1 if per_tensor else 2
    return out


def lg_qdq_bwd_impl(x, grad, enc_min, enc_max, bw, sym_mode, strict=False, ch_axis=0, need_grad_x=True,
                    need_grad_enc=True):
    """Fused learned-grid backward: (grad_x or None, grad_min or None, grad_max or None)."""
    _lg_check(x, enc_min, enc_max, bw)
    _require_cuda(grad)
    if grad.dtype != x.dtype or grad.shape != x.shape:
        raise ValueError("x and grad must share dtype and shape")
    per_tensor = enc_min.numel() == 1
    if per_tensor and x.is_contiguous(memory_format=torch.channels_last) and \
            grad.is_contiguous(memory_format=torch.channels_last) and not x.is_contiguous():
        pass
    else:
        x, grad = x.contiguous(), grad.contiguous()
    outer, c, inner = _lg_geometry(x, enc_min, ch_axis)
    gx = torch.empty_like(grad) if need_grad_x else None
    gmin = torch.empty_like(enc_min) if need_grad_enc else None
    gmax = torch.empty_like(enc_max) if need_grad_enc else None
    with _on_device(x):
        ws = _lg_workspace(x.device, c)
        _lib.check(_L.ab_lg_qdq_bwd(x.data_ptr(), grad.data_ptr(), gx.data_ptr() if need_grad_x else None, outer, c,
                                    inner, _dtype_code(x), enc_min.data_ptr(), enc_max.data_ptr(), int(bw),
                                    int(sym_mode), int(bool(strict)), gmin.data_ptr() if need_grad_enc else None,
                                    gmax.data_ptr() if need_grad_enc else None, ws.data_ptr(), _stream(x)))
    LAUNCHES["lg_bwd"] += 1 if per_tensor else 2
    return gx, gmin, gmax


class LearnedGridQdq(torch.autograd.Function):
    """Drop-in for the reference's QuantizeDequantizeFunc (v1/tensor_quantizer.py:854-963): same inputs, same three
    gradients, one kernel each way; only x and a copy of (min, max) are kept for the backward."""

    @staticmethod
    def forward(ctx, x, enc_min, enc_max, bw, sym_mode, strict, ch_axis, gate=False):
        y = lg_qdq_fwd_impl(x, enc_min, enc_max, bw, sym_mode, strict, ch_axis, gate=gate)
        ctx.config = (bw, sym_mode, strict, ch_axis)
        # the reference saves clones too: a reused module clamps the parameters again before the backward (:905-911)
        ctx.save_for_backward(x, enc_min.detach().clone(), enc_max.detach().clone())
        return y

    @staticmethod
    def backward(ctx, grad):
        x, enc_min, enc_max = ctx.saved_tensors
        bw, sym_mode, strict, ch_axis = ctx.config
        need_x = ctx.needs_input_grad[0]
        need_enc = ctx.needs_input_grad[1] or ctx.needs_input_grad[2]
        if not (need_x or need_enc):
            return (None,) * 8
        gx, gmin, gmax = lg_qdq_bwd_impl(x, grad, enc_min, enc_max, bw, sym_mode, strict, ch_axis, need_x, need_enc)
        return gx, gmin, gmax, None, None, None, None, None


# ---------------------------------------------------------------------------------------------------------------------
# host helpers (pure C, no device)
# ---------------------------------------------------------------------------------------------------------------------
def fill_encoding_info(bw, enc_min, enc_max):
    e = _lib.Encoding()
    _lib.check(_L.ab_fill_encoding_info(int(bw), float(enc_min), float(enc_max), C.byref(e)))
    return e


def per_channel_params(mins, maxs, bw):
    """Host fp32 parameter block [min | max | delta | offset] for a list of per-channel (min, max) doubles."""
    n = len(mins)
    a = (C.c_double * n)(*mins)
    b = (C.c_double * n)(*maxs)
    out = torch.empty(4 * n, dtype=torch.float32)
    _lib.check(_L.ab_per_channel_params(a, b, n, int(bw), C.cast(out.data_ptr(), C.POINTER(C.c_float))))
    return out


# ---------------------------------------------------------------------------------------------------------------------
# registration as torch custom ops
# ---------------------------------------------------------------------------------------------------------------------
_LIBRARY = torch.library.Library("aimet_b200", "DEF")
_LIBRARY.define("qdq_per_tensor(Tensor x, float enc_min, float enc_max, int bw, int round_mode=0, int seed=0) -> Tensor")
_LIBRARY.define("qdq_per_tensor_dev(Tensor x, Tensor enc4, int round_mode=0, int seed=0) -> Tensor")
_LIBRARY.define("quantize_to_grid(Tensor x, float enc_min, float enc_max, int bw, int round_mode=0, "
                "bool shift_to_signed=False, int seed=0) -> Tensor")
_LIBRARY.define("qdq_per_channel(Tensor x, Tensor params, int num_channel, int num_element_per_channel, "
                "int round_mode=0, int seed=0) -> Tensor")
_LIBRARY.define("ste_bwd(Tensor x, Tensor grad, float enc_min, float enc_max) -> Tensor")
_LIBRARY.define("ste_bwd_per_channel(Tensor x, Tensor grad, Tensor enc_min, Tensor enc_max, int num_channel, "
                "int num_element_per_channel) -> Tensor")
_LIBRARY.define("stats_update(Tensor x, Tensor(a!) states, int index, int quant_mode) -> ()")
_LIBRARY.define("stats_update_segmented(Tensor x, Tensor(a!) states, int first, int num_segments, int segment_len, "
                "int quant_mode) -> ()")

_LIBRARY.impl("qdq_per_tensor", qdq_per_tensor_impl, "CUDA")
_LIBRARY.impl("qdq_per_tensor_dev", qdq_per_tensor_dev_impl, "CUDA")
_LIBRARY.impl("quantize_to_grid", quantize_to_grid_impl, "CUDA")
_LIBRARY.impl("qdq_per_channel", qdq_per_channel_impl, "CUDA")
_LIBRARY.impl("ste_bwd", ste_bwd_impl, "CUDA")
_LIBRARY.impl("ste_bwd_per_channel", ste_bwd_per_channel_impl, "CUDA")
_LIBRARY.impl("stats_update", lambda x, states, index, quant_mode: stats_update_impl(x, states, index, quant_mode,
                                                                                      None, 0), "CUDA")
_LIBRARY.impl("stats_update_segmented", stats_update_segmented_impl, "CUDA")

for _name in ("qdq_per_tensor", "qdq_per_tensor_dev", "quantize_to_grid", "qdq_per_channel"):
    torch.library.register_fake(f"aimet_b200::{_name}", lambda x, *a, **k: torch.empty_like(x))
for _name in ("ste_bwd", "ste_bwd_per_channel"):
    torch.library.register_fake(f"aimet_b200::{_name}", lambda x, grad, *a, **k: torch.empty_like(grad))


def _qdq_per_tensor_setup(ctx, inputs, output):
    x, enc_min, enc_max = inputs[0], inputs[1], inputs[2]
    ctx.save_for_backward(x)
    # torch.tensor(python float) is float32: the reference compares against the fp32 image of the encoding range
    ctx.range = (float(torch.tensor(enc_min, dtype=torch.float32)), float(torch.tensor(enc_max, dtype=torch.float32)))


def _qdq_per_tensor_backward(ctx, grad):
    (x,) = ctx.saved_tensors
    return (torch.ops.aimet_b200.ste_bwd(x, grad, ctx.range[0], ctx.range[1]), None, None, None, None, None)


def _qdq_per_channel_setup(ctx, inputs, output):
    x, params, num_channel, per_channel = inputs[0], inputs[1], inputs[2], inputs[3]
    ctx.save_for_backward(x, params)
    ctx.geometry = (num_channel, per_channel)


def _qdq_per_channel_backward(ctx, grad):
    x, params = ctx.saved_tensors
    c, per = ctx.geometry
    # op-level autograd masks with the gated per-channel range the forward clamps to
    g = torch.ops.aimet_b200.ste_bwd_per_channel(x, grad, params[:c].contiguous(), params[c:2 * c].contiguous(), c, per)
    return (g, None, None, None, None, None)


torch.library.register_autograd("aimet_b200::qdq_per_tensor", _qdq_per_tensor_backward,
                                setup_context=_qdq_per_tensor_setup)
torch.library.register_autograd("aimet_b200::qdq_per_channel", _qdq_per_channel_backward,
                                setup_context=_qdq_per_channel_setup)
