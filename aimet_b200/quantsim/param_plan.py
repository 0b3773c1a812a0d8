"""All parameter quantizers of a sim refreshed together.

The reference derives the encodings of a wrapper's parameters inside that wrapper's forward -- reset, updateStats,
computeEncoding, one native call per channel (aimet_torch/v1/qc_quantize_op.py:753-798, v1/tensor_quantizer.py:567-570)
-- before every training-mode forward and once per calibration job. The parameters are all known before the forward
starts, so here the eligible parameter quantizers of a model share ONE contiguous block of statistics records and are
refreshed by ONE native call (ab_stats_refresh_encodings_multi: one reset, one statistics launch over all channels of all
weights, one grid search, one parameter-block launch) instead of four launches per weight; the wrappers then find their
encodings ready. Same kernels' arithmetic, same results (tests/test_gpu_quantsim.py).
"""
import weakref

import torch

from .. import ops
from ..state import StateArena, StateBlock
from ..tensor_quantizer_op import AimetTensorQuantizer, ValidityGroup
from . import tensor_quantizer as tq
from .qc_quantize_op import ForwardToken, StaticGridQuantWrapper  # noqa: F401  (ForwardToken re-exported)


class _Entry:
    __slots__ = ("wrapper", "name", "q", "first", "count", "shape", "axis")


class _Group:
    """Entries that share scheme, bitwidth, encoding flags and dtype -- one native call refreshes them all."""

    def __init__(self, key, device):
        self.key = key
        self.device = device
        self.entries = []
        self.total = 0
        self.block = None
        self.enc = self.qdq4 = self.params = None


def _eligible(wrapper, name, param, q):
    if not (tq.FUSED_REFRESH and q.enabled and q._lazy_ok and not q.is_encoding_frozen and q.bitwidth != 32):   # pylint: disable=protected-access
        return False
    if getattr(q, "_calib_hook", None) is not None or q.encoding_min_max_fixed_vals is not None:
        return False
    if q.use_symmetric_encodings and q.use_unsigned_symmetric:
        return False        # is_unsigned_symmetric needs the values on the host
    op0 = q._cppOp[0]       # pylint: disable=protected-access
    if not isinstance(op0, AimetTensorQuantizer) or op0._percentile is not None:   # pylint: disable=protected-access
        return False
    if not param.is_cuda or param.dtype not in (torch.float32, torch.bfloat16) or param.numel() == 0:
        return False
    if q.channel_axis is not None and len(q._cppOp) != param.shape[q.channel_axis]:   # pylint: disable=protected-access
        return False
    return True


def _key(param, q):
    return (q._cppOp[0]._code, q.bitwidth, bool(q.use_symmetric_encodings), bool(q.use_strict_symmetric),   # pylint: disable=protected-access
            bool(q.use_unsigned_symmetric), param.dtype, param.device)


class ParamPlan:
    def __init__(self, sim):
        self._sim = weakref.ref(sim)
        self.groups = []
        self._signature = None

    # ---- construction / validation ---------------------------------------------------------------------------
    def _scan(self, wrappers=None):
        found = []
        for wrapper in (wrappers if wrappers is not None else [w for _, w in self._sim().quant_wrappers()]):
            if not isinstance(wrapper, StaticGridQuantWrapper):
                continue
            for name, param in wrapper.get_named_parameters():
                q = wrapper.param_quantizers.get(name)
                if q is not None and _eligible(wrapper, name, param, q):
                    found.append((wrapper, name, param, q))
        return found

    def ensure(self, wrappers=None):
        """(Re)build the plan when the set of eligible quantizers, their flags or their parameters' shapes changed.
        `wrappers`: the sim's wrappers, when the caller has just walked the model anyway."""
        found = self._scan(wrappers)
        signature = tuple((id(q), id(q._cppOp[0]), _key(p, q), tuple(p.shape)) for _, _, p, q in found)   # pylint: disable=protected-access
        if signature == self._signature:
            return
        self._signature = signature
        groups = {}
        for wrapper, name, param, q in found:
            g = groups.setdefault(_key(param, q), _Group(_key(param, q), param.device))
            e = _Entry()
            e.wrapper, e.name, e.q, e.shape = wrapper, name, q, tuple(param.shape)
            e.axis = q.channel_axis
            e.count = 1 if q.channel_axis is None else len(q._cppOp)   # pylint: disable=protected-access
            e.first = g.total
            g.total += e.count
            g.entries.append(e)
        self.groups = list(groups.values())
        for g in self.groups:
            g.block = StateArena.for_device(g.device).allocate(g.total)
            g.enc = torch.empty((g.total, 5), dtype=torch.float64, device=g.device)
            g.qdq4 = torch.empty((g.total, 4), dtype=torch.float32, device=g.device)
            g.params = torch.empty(4 * g.total, dtype=torch.float32, device=g.device)
            for e in g.entries:
                ops_ = e.q._cppOp   # pylint: disable=protected-access
                if e.axis is None:
                    ops_[0]._bind(g.block, e.first)   # pylint: disable=protected-access
                else:
                    sub = StateBlock(None, g.block.arena, g.block.first + e.first, e.count)
                    group = ValidityGroup()
                    e.q._block, e.q._group = sub, group   # pylint: disable=protected-access
                    for i, op in enumerate(ops_):
                        op._bind(sub, i, group)           # pylint: disable=protected-access

    # ---- the refresh ---------------------------------------------------------------------------------------------
    def mark_reset_pending(self):
        """prepare_sim_for_compute_encodings is about to reset every quantizer: the planned ones need no launch of their
        own for that, the refresh that follows resets their whole block at once."""
        for g in self.groups:
            for e in g.entries:
                e.q._reset_is_pending = True   # pylint: disable=protected-access

    def refresh(self, only=None, stamp=None):
        """Refresh the encodings of the planned quantizers (`only(entry) -> bool` selects a subset) from the current
        parameter values. Returns the number of quantizers refreshed."""
        launched = self.launch(only)
        self.stamp(launched, stamp)
        return sum(len(run) for _, run in launched)

    def launch(self, only=None):
        """The device half of `refresh`: the native calls are enqueued, the quantizers are not touched yet. The caller may
        do unrelated host work (which may reset the very quantizers: prepare_sim_for_compute_encodings) while the device
        derives the encodings, and must then hand the result to `stamp`."""
        launched = []
        for g in self.groups:
            runs, run = [], []
            for e in g.entries:                       # consecutive records by construction
                if only is not None and not only(e):
                    if run:
                        runs.append(run)
                        run = []
                    continue
                if len(run) >= ops.REFRESH_MAX_ITEMS:
                    runs.append(run)
                    run = []
                run.append(e)
            if run:
                runs.append(run)
            for run in runs:
                self._launch_run(g, run, run[0].first)
                launched.append((g, run))
        return launched

    def stamp(self, launched, stamp=None):
        """The host half of `refresh`: the quantizers of the launched runs now own the device-resident encodings."""
        for g, run in launched:
            self._stamp_run(g, run, stamp)

    @staticmethod
    def _launch_run(g, run, start):
        code, bw, sym, strict, unsigned_sym = g.key[:5]
        tensors, segs = [], []
        for e in run:
            data = dict(e.wrapper.get_named_parameters())[e.name].data
            if e.axis not in (None, 0):
                data = data.movedim(e.axis, 0)
            tensors.append(data.contiguous(memory_format=torch.contiguous_format))
            segs.append(e.count)
        ops.stats_refresh_multi_impl(tensors, segs, g.block.arena, g.block.first, code, bw, sym, strict, unsigned_sym,
                                     g.enc, g.qdq4, g.params, first_record=start)

    @staticmethod
    def _stamp_run(g, run, stamp):
        for e in run:
            q = e.q
            q._enc_dev = g.enc[e.first:e.first + e.count]                      # pylint: disable=protected-access
            if e.axis is None:
                q._qdq4_dev, q._params_dev = g.qdq4[e.first:e.first + 1], None   # pylint: disable=protected-access
                op0 = q._cppOp[0]                                               # pylint: disable=protected-access
                op0._is_encoding_valid = True                                   # pylint: disable=protected-access
                op0._range_fixed, op0._probe, op0._updates = False, None, 1     # pylint: disable=protected-access
            else:
                q._qdq4_dev, q._params_dev = None, g.params[4 * e.first:4 * (e.first + e.count)]   # pylint: disable=protected-access
                q._group.valid = True                                           # pylint: disable=protected-access
            q._encoding = tq._LAZY                                              # pylint: disable=protected-access
            q.is_unsigned_symmetric = False
            q._stats_dirty = False                                              # pylint: disable=protected-access
            q.__dict__.pop("_reset_is_pending", None)
            q._fresh_token = stamp                                              # pylint: disable=protected-access
