"""Host mirror of the reference's QuantizationSimModel on the static-grid (tf / tf_enhanced) path.

Reference: TrainingExtensions/torch/src/python/aimet_torch/v1/quantsim.py -- __init__ :225-307,
prepare_sim_for_compute_encodings :381-400, compute_layer_encodings_for_sim :403-423, compute_encodings :425-448,
save_encodings_to_json / get_activation_param_encodings :680-729, export :486-575 (the `_torch.encodings` file; the
ONNX-keyed twin needs the `onnx` package and an ONNX export, which are outside this path).
"""
import copy
import json
import os
from collections import OrderedDict, defaultdict
from contextlib import contextmanager
from typing import Callable, Dict, List, Optional, Tuple

import torch
from torch import nn

from . import config as qconfig
from .defs import MAP_ROUND_MODE_TO_PYMO, QuantizationDataType, QuantScheme
from .learned_grid import LearnedGridQuantWrapper, construct_and_initialize_trainable_wrapper
from .qc_quantize_op import CalibrationJob, QcQuantizeOpMode, StaticGridQuantWrapper

ENCODING_VERSION = "0.6.1"   # reference aimet_common/quantsim.py:55-57
_RANGE_LEARNING_SCHEMES = (QuantScheme.training_range_learning_with_tf_init,
                           QuantScheme.training_range_learning_with_tf_enhanced_init)
_SUPPORTED_SCHEMES = (QuantScheme.post_training_tf, QuantScheme.post_training_tf_enhanced,
                      QuantScheme.post_training_percentile) + _RANGE_LEARNING_SCHEMES
_WRAPPER_TYPES = (StaticGridQuantWrapper, LearnedGridQuantWrapper)
unquantizable_modules = (nn.Identity,)


@contextmanager
def in_eval_mode(model: nn.Module):
    was_training = {m: m.training for m in model.modules()}
    model.eval()
    try:
        yield
    finally:
        for m, t in was_training.items():
            m.training = t


def _count_inout(model: nn.Module, dummy_input) -> Dict[nn.Module, Tuple[int, int]]:
    """One forward with hooks: number of input / output tensors of each leaf (reference utils.get_inout_tensor_shape_per_module)."""
    counts = {}
    hooks = []

    def hook(mod, inp, out):
        n_in = len(inp) if isinstance(inp, (tuple, list)) else 1
        n_out = len(out) if isinstance(out, (tuple, list)) else 1
        counts[mod] = (n_in, n_out)

    for m in model.modules():
        if len(list(m.children())) == 0:
            hooks.append(m.register_forward_hook(hook))
    with in_eval_mode(model), torch.no_grad():
        if isinstance(dummy_input, (tuple, list)):
            model(*dummy_input)
        else:
            model(dummy_input)
    for h in hooks:
        h.remove()
    return counts


EAGER_BATCHES = 2   # batch 0 derives encodings / fixes ranges; batch 1 refreshes every host-side cache; then capture
LAST_GRAPH_INFO = {"captured_launches": 0, "replays": 0}   # bookkeeping of the most recent run_batches (for bench.py)


def run_batches(model, batches, cuda_graph: bool, after_first=None, after_each=None, device=None):
    """Run `model` over `batches`; with `cuda_graph`, every batch from the third on that has the first one's shape is a
    replay of one captured forward. Returns the number of batches run."""
    if device is None:
        try:
            device = next(model.parameters()).device
        except StopIteration:
            device = None
    from .. import ops
    graph = static_in = None
    shape = dtype = None
    n = 0
    LAST_GRAPH_INFO.update(captured_launches=0, replays=0)
    for b in batches:
        on_device = b if (device is None or b.device == device) else None
        if n == 0:
            x = on_device if on_device is not None else b.to(device, non_blocking=True)
            model(x)
            shape, dtype = tuple(x.shape), x.dtype
            if after_each is not None:
                after_each(0)
            if after_first is not None:
                after_first()
        elif cuda_graph and n >= EAGER_BATCHES and device is not None and device.type == "cuda" and \
                tuple(b.shape) == shape and b.dtype == dtype:
            if graph is None:
                static_in = torch.empty(shape, dtype=dtype, device=device)
                static_in.copy_(b, non_blocking=True)
                torch.cuda.current_stream(device).synchronize()
                graph = torch.cuda.CUDAGraph()
                before = ops.launches_total()
                with torch.cuda.graph(graph):
                    model(static_in)
                LAST_GRAPH_INFO["captured_launches"] = ops.launches_total() - before
            else:
                static_in.copy_(b, non_blocking=True)
            graph.replay()
            LAST_GRAPH_INFO["replays"] += 1
            if after_each is not None:
                after_each(n)
        else:
            model(on_device if on_device is not None else b.to(device, non_blocking=True))
            if after_each is not None:
                after_each(n)
        n += 1
    return n


class _ModelForwardHooks:
    """Forward pre / post hooks of `sim.model`: in training mode every wrapper re-derives its parameter encodings before
    every forward (reference qc_quantize_op.py:753-798); the pre hook does that for ALL parameters of the model in one
    native call (quantsim.param_plan) and stamps them as fresh for this forward. Holds the sim weakly; copies of the model
    (deepcopy, unpickling) keep an inert instance."""

    def __init__(self, sim):
        import weakref
        self._sim = weakref.ref(sim)

    def __deepcopy__(self, memo):
        return self

    def __getstate__(self):
        return {}

    def __setstate__(self, state):
        self._sim = lambda: None

    def pre(self, module, args):   # pylint: disable=unused-argument
        sim = self._sim()
        if sim is not None and module is sim.model:
            sim._begin_model_forward()   # pylint: disable=protected-access

    def post(self, module, args, output):   # pylint: disable=unused-argument
        from .param_plan import ForwardToken
        sim = self._sim()
        if sim is not None and module is sim.model:
            ForwardToken.active = False


class QuantizationSimModel:
    """Adds quantization-simulation wrappers to a model, calibrates them and exports the encodings."""

    def __init__(self, model: nn.Module, dummy_input, quant_scheme="tf_enhanced", rounding_mode: str = "nearest",
                 default_output_bw: int = 8, default_param_bw: int = 8, in_place: bool = False, config_file=None,
                 default_data_type: QuantizationDataType = QuantizationDataType.int):
        if isinstance(quant_scheme, str):
            quant_scheme = QuantScheme.from_str(quant_scheme)
        if quant_scheme not in _SUPPORTED_SCHEMES:
            raise NotImplementedError(f"{quant_scheme} is outside the aimet_b200 hot path (tf / tf_enhanced / percentile "
                                      "and the range-learning schemes initialised from tf / tf_enhanced)")
        if default_data_type != QuantizationDataType.int:
            raise NotImplementedError("only integer quantization simulation is on the aimet_b200 hot path")
        self.model = model if in_place else copy.deepcopy(model)
        self._quant_scheme = quant_scheme
        self._rounding_mode = rounding_mode
        self._default_output_bw = default_output_bw
        self._default_param_bw = default_param_bw
        self._percentile_value = 100   # reference :278
        self._config = qconfig.load_config(config_file)

        try:
            ops = qconfig.build_op_graph(self.model)          # call-site graph, before the wrappers go in
        except Exception as exc:   # pylint: disable=broad-except
            raise RuntimeError(f"torch.fx could not trace the model ({exc}); the supergroup / model-input rules of the "
                               "quantsim config need an op graph") from exc
        inout = _count_inout(self.model, dummy_input)
        self._wrappers: Dict[nn.Module, StaticGridQuantWrapper] = {}
        self._add_quantization_wrappers(self.model, inout)
        # bias parameters are never quantized (reference :290 exclude_param_from_quantization("bias"))
        qconfig.configure(self.model, self._wrappers, self._config, ops)
        for w in self._wrappers.values():
            if "bias" in w.param_quantizers:
                w.param_quantizers["bias"].enabled = False
        self._register_model_hooks()

    def __getstate__(self):
        state = self.__dict__.copy()
        state.pop("_act_block", None)               # device-resident statistics are not part of a checkpoint
        state.pop("_act_block_quantizers", None)
        state.pop("_param_plan", None)
        state.pop("_model_hooks", None)
        return state

    def __setstate__(self, state):
        self.__dict__.update(state)
        self._register_model_hooks()

    def _register_model_hooks(self):
        hooks = self._model_hooks = _ModelForwardHooks(self)
        self.model.register_forward_pre_hook(hooks.pre)
        self.model.register_forward_hook(hooks.post, always_call=True)

    # ---- all parameter encodings at once -------------------------------------------------------------------------
    def _plan(self):
        """The sim's ParamPlan, or None when the model's parameters are not on a CUDA device."""
        from .param_plan import ParamPlan
        device = next((p.device for p in self.model.parameters()), None)
        if device is None or device.type != "cuda":
            return None
        plan = self.__dict__.get("_param_plan")
        if plan is None:
            plan = self._param_plan = ParamPlan(self)
        return plan

    def _begin_model_forward(self):
        """Pre hook of every forward of self.model. Training mode: refresh, in one native call, the encodings of every
        planned parameter whose wrapped module is in training mode, and stamp them fresh for this forward."""
        from .param_plan import ForwardToken
        ForwardToken.current += 1
        ForwardToken.active = False
        if not any(w._module_to_wrap.training for w in self._wrappers.values()   # pylint: disable=protected-access
                   if isinstance(w, StaticGridQuantWrapper) and w.param_quantizers):
            return
        plan = self._plan()
        if plan is None:
            return
        plan.ensure()
        with torch.no_grad():
            plan.refresh(only=lambda e: e.wrapper._module_to_wrap.training or not e.q._has_encoding(),   # pylint: disable=protected-access
                         stamp=ForwardToken.current)
        ForwardToken.active = True

    # ---- model surgery -----------------------------------------------------------------------------------------
    @staticmethod
    def _is_quantizable_module(module: nn.Module) -> bool:
        return type(module) != nn.Module and not isinstance(module, unquantizable_modules) and \
            not isinstance(module, _WRAPPER_TYPES)   # pylint: disable=unidiomatic-typecheck

    def _add_quantization_wrappers(self, module: nn.Module, inout):
        for name, child in list(module.named_children()):
            if isinstance(child, _WRAPPER_TYPES):
                continue
            if len(list(child.children())) == 0:
                if self._is_quantizable_module(child):
                    # a module that the forward pass never reaches is wrapped too, with one input and one output
                    # (reference :1409-1411)
                    n_in, n_out = inout.get(child, (1, 1))
                    w = StaticGridQuantWrapper(child, self._default_param_bw, self._default_output_bw,
                                               MAP_ROUND_MODE_TO_PYMO[self._rounding_mode], self._quant_scheme,
                                               num_inputs=n_in, num_outputs=n_out)
                    self._wrappers[child] = w
                    setattr(module, name, w)
            else:
                self._add_quantization_wrappers(child, inout)

    def quant_wrappers(self):
        """(name, wrapper) for every wrapper in the model (reference quant_wrappers())."""
        for name, m in self.model.named_modules():
            if isinstance(m, _WRAPPER_TYPES):
                yield name, m

    _get_qc_quantized_layers = lambda self, model=None: list(self.quant_wrappers())   # noqa: E731

    def exclude_layers_from_quantization(self, layers_to_exclude):
        """Take the wrappers inside the given layers (modules of `self.model`) out again (reference :731-751)."""
        doomed = {m for layer in layers_to_exclude for m in layer.modules() if isinstance(m, _WRAPPER_TYPES)}
        if not doomed:
            return

        def strip(parent):
            for name, child in list(parent.named_children()):
                if child in doomed:
                    setattr(parent, name, child.get_original_module())
                else:
                    strip(child)

        strip(self.model)
        self._wrappers = {orig: w for orig, w in self._wrappers.items() if w not in doomed}
        self.__dict__.pop("_act_block", None)               # the activation statistics block is rebuilt on the next bind
        self.__dict__.pop("_act_block_quantizers", None)
        self.__dict__.pop("_param_plan", None)

    # ---- calibration -------------------------------------------------------------------------------------------
    @staticmethod
    def prepare_sim_for_compute_encodings(sim: "QuantizationSimModel"):
        wrappers = [layer for _, layer in sim.quant_wrappers()]      # ONE walk over the model for everything below
        if any(isinstance(layer, LearnedGridQuantWrapper) for layer in wrappers):
            raise RuntimeError("the wrappers have already been replaced by range-learning wrappers; their encodings are "
                               "trainable parameters now and are not re-calibrated")
        plan = sim._plan()              # pylint: disable=protected-access
        launched = None
        if plan is not None:
            # The parameters are known now: derive all their encodings in one native call instead of four launches per
            # weight inside the first forward (the wrappers find them ready; same values either way). The call is enqueued
            # FIRST -- it resets the planned records block-wide itself -- so the device works through the 26 561 grid
            # searches while the host does the per-wrapper bookkeeping below; the quantizers are stamped at the end.
            plan.ensure(wrappers)
            plan.mark_reset_pending()   # their records are reset block-wide by that call, not one launch each
            with torch.no_grad():
                launched = plan.launch()
        sim._bind_activation_states(wrappers)   # pylint: disable=protected-access
        if getattr(sim, "_act_block", None) is not None:
            sim._act_block.reset()      # all activation records in one launch  # pylint: disable=protected-access
            for q in sim._act_block_quantizers:   # pylint: disable=protected-access
                q._reset_done_blockwide = True    # pylint: disable=protected-access
        for layer in wrappers:
            layer.reset_encodings()
            layer.set_mode(QcQuantizeOpMode.ANALYSIS)
        if launched is not None:
            plan.stamp(launched)
        if sim._quant_scheme == QuantScheme.post_training_percentile:   # pylint: disable=protected-access
            for layer in wrappers:                                      # reference :397-400
                layer.set_percentile_value(sim._percentile_value)       # pylint: disable=protected-access

    def set_percentile_value(self, percentile_value: float):
        """reference :478-484"""
        if percentile_value < 90 or percentile_value > 100:
            raise ValueError("Percentile value must be in range [90, 100]")
        self._percentile_value = percentile_value

    def activation_quantizers(self, wrappers=None):
        """Enabled per-tensor activation quantizers in a deterministic (module, input/output, index) order."""
        out = []
        for layer in (wrappers if wrappers is not None else [w for _, w in self.quant_wrappers()]):
            if not isinstance(layer, StaticGridQuantWrapper):
                continue
            for q in layer.input_quantizers + layer.output_quantizers:
                if q.enabled and q.bitwidth != 32 and not q.is_encoding_frozen:
                    out.append(q)
        return out

    def _bind_activation_states(self, wrappers=None):
        """Give all activation quantizers ONE contiguous block of device statistics records, so that range injection,
        the ordered replay and the final grid search are single launches over the whole model."""
        from ..state import StateArena
        from ..tensor_quantizer_op import AimetTensorQuantizer
        try:
            device = next(self.model.parameters()).device
        except StopIteration:
            return
        if device.type != "cuda":
            return
        quantizers = [q for q in self.activation_quantizers(wrappers) if isinstance(q._cppOp[0], AimetTensorQuantizer)]   # pylint: disable=protected-access
        if not quantizers:
            return
        block = getattr(self, "_act_block", None)
        if block is None or block.count != len(quantizers) or block.device != device:
            block = self._act_block = StateArena.for_device(device).allocate(len(quantizers))
        for i, q in enumerate(quantizers):
            q._cppOp[0]._bind(block, i)   # pylint: disable=protected-access
        self._act_block_quantizers = quantizers

    @staticmethod
    def compute_layer_encodings_for_sim(sim: "QuantizationSimModel"):
        wrappers = [layer for _, layer in sim.quant_wrappers()]
        # The grid searches of all activation quantizers are enqueued and their read-back started; the per-wrapper host work
        # that does not need those results (every other quantizer's compute_encoding -- an early exit for parameters whose
        # encodings exist --, the mode switch) runs while the device is still working through the forwards; only then does
        # the host wait for the copy.
        finish, handled = sim._compute_activation_encodings_batched(wrappers)   # pylint: disable=protected-access
        for layer in wrappers:
            if handled:
                for q in layer.input_quantizers + list(layer.param_quantizers.values()) + layer.output_quantizers:
                    if id(q) not in handled:
                        q.compute_encoding()
            else:
                layer.compute_encoding()
            layer.set_mode(QcQuantizeOpMode.ACTIVE)
        finish()
        sim.replace_wrappers_for_quantize_dequantize()

    def replace_wrappers_for_quantize_dequantize(self):
        """Range-learning schemes: every static-grid wrapper becomes a LearnedGridQuantWrapper whose (min, max) parameters
        start from the encodings just computed (reference :764-845)."""
        if self._quant_scheme not in _RANGE_LEARNING_SCHEMES:
            return
        try:
            device = next(self.model.parameters()).device
        except StopIteration:
            device = torch.device("cpu")

        def replace(parent):
            for name, child in list(parent.named_children()):
                if isinstance(child, StaticGridQuantWrapper):
                    new = construct_and_initialize_trainable_wrapper(
                        child, device, self._default_param_bw, self._default_output_bw, self._rounding_mode,
                        self._quant_scheme)
                    self._wrappers[child.get_original_module()] = new
                    setattr(parent, name, new)
                elif not isinstance(child, LearnedGridQuantWrapper):
                    replace(child)

        replace(self.model)

    def compute_encodings(self, forward_pass_callback: Callable, forward_pass_callback_args):
        """Runs the user's calibration callback with every wrapper collecting statistics, then computes the encodings."""
        from .stats_batcher import StatsBatcher
        QuantizationSimModel.prepare_sim_for_compute_encodings(self)
        prefetch = _ParamExportPrefetch(self)
        batcher = StatsBatcher.attach(self)     # activation statistics: deferred, one launch per forward (or None)
        try:
            with in_eval_mode(self.model), torch.no_grad(), CalibrationJob(self):
                _ = forward_pass_callback(self.model, forward_pass_callback_args)
            if batcher is not None:
                batcher.flush()
        finally:
            prefetch.close()
            if batcher is not None:
                batcher.detach()
        QuantizationSimModel.compute_layer_encodings_for_sim(self)

    def compute_encodings_for_batches(self, batches, cuda_graph: bool = True):
        """compute_encodings for the common case where the calibration callback is `for x in batches: model(x)`.

        Knowing the loop lets the steady state be captured in a CUDA graph: the first batch runs eagerly (it fixes the
        histogram ranges and derives the weight encodings, which needs host decisions); from the second batch on one
        forward -- wrapped modules, statistics kernels, weight QDQ -- is recorded once and replayed per batch, so the
        host cost of ~130 Python wrappers and ~250 launches per step disappears. Statistics are device-resident and all
        their control flow is on the device, which is what makes the step capturable. Results are identical to
        compute_encodings(). `batches`: iterable of CUDA tensors, or pinned host tensors (copied per step)."""
        from .stats_batcher import StatsBatcher
        QuantizationSimModel.prepare_sim_for_compute_encodings(self)
        batcher = StatsBatcher.attach(self)
        try:
            with in_eval_mode(self.model), torch.no_grad(), CalibrationJob(self):
                # the batcher learns the fixed ranges itself at the end of the first forward
                run_batches(self.model, batches, cuda_graph,
                            after_first=None if batcher is not None else self._learn_fixed_ranges)
            if batcher is not None:
                batcher.flush()
        finally:
            if batcher is not None:
                batcher.detach()
        QuantizationSimModel.compute_layer_encodings_for_sim(self)

    def _learn_fixed_ranges(self):
        """One read-back after the first batch: which activation records have their histogram range fixed (so that the
        min/max launch is skipped from now on, and is not baked into a captured graph)."""
        block = getattr(self, "_act_block", None)
        if block is None:
            return
        rec = block.read()
        for q, initialized in zip(self._act_block_quantizers, rec["initialized"].tolist()):
            q._cppOp[0]._range_fixed = bool(initialized)   # pylint: disable=protected-access

    def _compute_activation_encodings_batched(self, wrappers=None):
        """All per-tensor grid searches are enqueued first and read back with ONE device->host copy per device (the
        reference does one blocking native call per quantizer). Returns (finish, handled): `handled` holds the ids of the
        quantizers whose searches are in flight, `finish()` waits for the copy and gives them their encodings. Quantizers
        that are not backed by the native op are left to the generic path (compute_encoding())."""
        from .. import libpymo, ops
        from ..tensor_quantizer_op import AimetTensorQuantizer
        from .tensor_quantizer import StaticGridPerTensorQuantizer
        pending = []
        for layer in (wrappers if wrappers is not None else [w for _, w in self.quant_wrappers()]):
            for q in layer.input_quantizers + list(layer.param_quantizers.values()) + layer.output_quantizers:
                if not isinstance(q, StaticGridPerTensorQuantizer) or not q.enabled or q.is_encoding_frozen or \
                        q.bitwidth == 32 or (q._has_encoding() and not q._stats_dirty):   # pylint: disable=protected-access
                    continue
                op = q._cppOp[0]   # pylint: disable=protected-access
                if isinstance(op, AimetTensorQuantizer) and op._is_encoding_valid and op._block is not None:   # pylint: disable=protected-access
                    pending.append((q, op))
        if not pending:
            return (lambda: None), set()
        by_device = defaultdict(list)
        for q, op in pending:
            by_device[op._block.device].append((q, op))   # pylint: disable=protected-access
        in_flight = []
        for device, items in by_device.items():
            out = torch.empty((len(items), 5), dtype=torch.float64, device=device)
            first_q, first_op = items[0]
            key = lambda q, op: (op._code, op._percentile, q.bitwidth, q.use_symmetric_encodings,   # noqa: E731
                                 q.use_strict_symmetric, q.use_unsigned_symmetric)
            contiguous = all(op._block is first_op._block and op._index == first_op._index + i and   # pylint: disable=protected-access
                             key(q, op) == key(first_q, first_op) for i, (q, op) in enumerate(items))
            if contiguous:
                ops.compute_encodings_into(first_op._block.arena, first_op._block.first + first_op._index, len(items),   # pylint: disable=protected-access
                                           first_op._code, first_q.bitwidth, first_q.use_symmetric_encodings,   # pylint: disable=protected-access
                                           first_q.use_strict_symmetric, first_q.use_unsigned_symmetric, out,
                                           percentile=first_op._percentile)   # pylint: disable=protected-access
            for row, (q, op) in enumerate(items if not contiguous else []):
                ops.compute_encodings_into(op._block.arena, op._block.first + op._index, 1, op._code, q.bitwidth,   # pylint: disable=protected-access
                                           q.use_symmetric_encodings, q.use_strict_symmetric,
                                           q.use_unsigned_symmetric, out[row:row + 1], percentile=op._percentile)   # pylint: disable=protected-access
            host = torch.empty(out.shape, dtype=out.dtype, pin_memory=True)
            with torch.cuda.device(device):
                host.copy_(out, non_blocking=True)
                event = torch.cuda.Event()
                event.record()
            in_flight.append((items, host, event, out))
        # a quantizer whose search is in flight must not be recomputed by the loop that runs meanwhile
        for q, _ in pending:
            q._stats_dirty = False   # pylint: disable=protected-access

        def finish():
            for items, host, event, _ in in_flight:
                event.synchronize()
                for (q, _), r in zip(items, host.tolist()):
                    q._encoding = [libpymo.TfEncoding._from_values(r[0], r[1], r[2], r[3], int(r[4]))]   # pylint: disable=protected-access
                    q.is_unsigned_symmetric = q.use_symmetric_encodings and q.use_unsigned_symmetric and \
                        r[0] >= 0 and r[1] >= 0

        return finish, {id(q) for q, _ in pending}

    # ---- export ------------------------------------------------------------------------------------------------
    def get_activation_param_encodings(self):
        """reference :692-729"""
        activation_encodings = OrderedDict()
        param_encodings = OrderedDict()
        for module_name, module in self.quant_wrappers():
            activation_encodings[module_name] = defaultdict(OrderedDict)
            for i, encoding in enumerate(module.export_input_encodings()):
                if not encoding:
                    continue
                activation_encodings[module_name]["input"][i] = encoding[0] if len(encoding) == 1 else encoding
            for i, encoding in enumerate(module.export_output_encodings()):
                if not encoding:
                    continue
                activation_encodings[module_name]["output"][i] = encoding[0] if len(encoding) == 1 else encoding
            if not activation_encodings[module_name]:
                del activation_encodings[module_name]
            for param_name, encoding in module.export_param_encodings().items():
                if not encoding:
                    continue
                param_encodings[f"{module_name}.{param_name}"] = encoding
        return activation_encodings, param_encodings

    def save_encodings_to_json(self, path: str, filename_prefix: str):
        """reference :680-690"""
        activation_encodings, param_encodings = self.get_activation_param_encodings()
        encodings_dict = {"activation_encodings": activation_encodings, "param_encodings": param_encodings}
        with open(os.path.join(path, filename_prefix + ".json"), "w") as f:
            json.dump(encodings_dict, f, sort_keys=True, indent=4)

    def load_encodings(self, encodings, strict: bool = True, partial: bool = True, requires_grad=None,
                       allow_overwrite: bool = True):
        """reference :1696-1757 -- `encodings`: the dictionary (or the path of the JSON file) that save_encodings_to_json /
        export wrote: {'activation_encodings': {module: {'input'|'output': {idx: enc}}}, 'param_encodings': {name: [enc]}}.
        After loading, every wrapper is in ACTIVE mode and quantize-dequantizes with the loaded encodings."""
        if isinstance(encodings, (str, os.PathLike)):
            with open(encodings) as f:
                encodings = json.load(f)
        if "param_encodings" not in encodings:     # the older AdaRound export: parameter encodings only (reference :1726-1732)
            param_encodings, activation_encodings = encodings, {}
        else:
            param_encodings = encodings.get("param_encodings", {})
            activation_encodings = encodings.get("activation_encodings", {})
        if not param_encodings and not activation_encodings:
            raise RuntimeError("no encodings to load")
        wrappers = dict(self.quant_wrappers())
        if strict:
            known = set(wrappers) | {f"{m}.{p}" for m, w in wrappers.items() for p in w.param_quantizers}
            missing = (set(param_encodings) | set(activation_encodings)) - known
            if missing:
                raise RuntimeError("Encoding dictionary contains modules/parameters that doesn't exist in the model: " +
                                   ", ".join(sorted(missing)))
        for name, wrapper in wrappers.items():
            own = {p: param_encodings.get(f"{name}.{p}") for p in wrapper.param_quantizers}
            wrapper.import_param_encodings({k: v for k, v in own.items() if v}, strict, partial, requires_grad,
                                           allow_overwrite)
            act = activation_encodings.get(name, {})
            wrapper.import_input_encodings(act.get("input", {}), strict, partial, requires_grad, allow_overwrite)
            wrapper.import_output_encodings(act.get("output", {}), strict, partial, requires_grad, allow_overwrite)
            wrapper.set_mode(QcQuantizeOpMode.ACTIVE)

    def load_and_freeze_encodings(self, encoding_path: str, ignore_when_quantizer_disabled: bool = False):
        """Set activation and parameter encodings from a `<prefix>_torch.encodings` / save_encodings_to_json file and
        freeze them: later compute_encodings / load_encodings calls leave them alone (reference :1759-1775)."""
        self.load_encodings(encoding_path, strict=not ignore_when_quantizer_disabled, partial=True, requires_grad=False,
                            allow_overwrite=False)

    def set_and_freeze_param_encodings(self, encoding_path: str):
        """Parameter encodings only, frozen (reference :1838-1855; deprecated there in favour of load_encodings)."""
        with open(encoding_path) as f:
            encodings = json.load(f)
        encodings.pop("activation_encodings", None)
        self.load_encodings(encodings, strict=True, partial=True, requires_grad=False, allow_overwrite=False)

    def exclude_param_from_quantization(self, param_name_to_exclude: str):
        """Disable the quantizer of every parameter with this name, e.g. "bias" (reference :753-762)."""
        for _, wrapper in self.quant_wrappers():
            if param_name_to_exclude in wrapper.param_quantizers:
                wrapper.param_quantizers[param_name_to_exclude].enabled = False

    def named_qmodules(self):
        """(name, quantized module) pairs (reference :1857-1862)."""
        yield from self.quant_wrappers()

    def qmodules(self):
        yield from (m for _, m in self.named_qmodules())

    def export(self, path: str, filename_prefix: str, dummy_input=None, **_unused):
        """Writes `<prefix>_torch.encodings` (torch-module-name keyed, reference :1000-1042 layout) and `<prefix>.pth`, the
        pickled original model as the reference saves it. The ONNX-tensor-name keyed `<prefix>.encodings` needs an ONNX
        export and is not produced here."""
        os.makedirs(path, exist_ok=True)
        activation_encodings, param_encodings = self.get_activation_param_encodings()
        torch_encodings = {"version": ENCODING_VERSION,
                           "activation_encodings": activation_encodings,
                           "param_encodings": param_encodings,
                           "excluded_layers": [],
                           "quantizer_args": self._quantizer_args()}
        with open(os.path.join(path, filename_prefix + "_torch.encodings"), "w") as f:
            json.dump(torch_encodings, f, sort_keys=True, indent=4)
        # `<prefix>.pth` is the pickled ORIGINAL model (wrappers removed, weights untouched), as the reference writes it
        # (v1/quantsim.py:529-531 torch.save(model_to_export, ...)); the quantize-dequantized weights additionally go to
        # `<prefix>_qdq_state_dict.pth` (not a reference artefact; a plain state_dict).
        torch.save(self.get_original_model(self.model), os.path.join(path, filename_prefix + ".pth"))
        torch.save(self.get_original_model(self.model, qdq_weights=True).state_dict(),
                   os.path.join(path, filename_prefix + "_qdq_state_dict.pth"))

    def _quantizer_args(self) -> Dict:
        """reference aimet_common/quantsim.py:280-311 (extract_global_quantizer_args): range-learning schemes are reported
        as the post-training scheme they were initialised from; is_symmetric is the parameter default of the config,
        falling back to the per-channel flag."""
        defaults = self._config.get("defaults", {})
        per_channel = qconfig._truthy(defaults.get("per_channel_quantization", False))   # pylint: disable=protected-access
        params = defaults.get("params", {})
        scheme = {QuantScheme.training_range_learning_with_tf_init: QuantScheme.post_training_tf,
                  QuantScheme.training_range_learning_with_tf_enhanced_init: QuantScheme.post_training_tf_enhanced
                  }.get(self._quant_scheme, self._quant_scheme)
        return {"quant_scheme": scheme.name,
                "param_bitwidth": self._default_param_bw,
                "activation_bitwidth": self._default_output_bw,
                "dtype": "int",
                "is_symmetric": qconfig._truthy(params["is_symmetric"]) if "is_symmetric" in params else per_channel,   # pylint: disable=protected-access
                "per_channel_quantization": per_channel}

    def capture_forward(self, *sample_inputs, warmup: int = 3) -> "GraphedForward":
        """The calibrated model's inference forward as a CUDA graph (no counterpart in the reference; B200 idiom for a
        launch-bound loop). A quantsim forward is ~2x the kernel launches of the plain model and, at small batch or in
        bf16, host-bound: ResNet-50 bf16 batch 32 takes 6.6 ms eager and 3.7 ms replayed (the plain model eager: 3.7 ms).
        The layers launch on the capturing stream, allocate through torch and keep their encodings on the device, so the
        capture needs nothing special; it must be repeated after anything that changes encodings, enabled flags or weights'
        storage. The returned callable copies its arguments into the captured input buffers, replays, and returns the
        captured output tensors (overwritten by the next call)."""
        return GraphedForward(self.model, sample_inputs, warmup)

    def capture_train_step(self, loss_fn: Callable, optimizer: torch.optim.Optimizer, sample_inputs, sample_target,
                           warmup: int = 3, model: Optional[nn.Module] = None) -> "GraphedTrainStep":
        """One quantization-aware training step -- forward through the wrappers, loss, backward with the straight-through
        gates, optimizer step -- captured in a CUDA graph (no counterpart in the reference; the B200 idiom for a
        launch-bound loop). A QAT step issues ~2x the kernels of the plain step from ~140 Python wrappers and ~120 autograd
        functions: MobileNet-v2, batch 32, eager 18 ms per step against 13 ms for the plain model; replayed from the graph
        the host cost is gone. Everything the step does is device-side and allocation-free where torch is not the allocator
        (the parameter encodings of the whole model are re-derived by one native call into preallocated tables), which is
        what makes it capturable. Returns a callable `step(*inputs, target) -> loss tensor` (static, overwritten by the
        next call). Capture again after anything that changes quantizer configuration or enabled flags.

        `model`: the module the step calls, when that is a wrapper around `self.model` -- a DistributedDataParallel
        instance, so that its reducer's all-reduces are captured with the backward. torch's rules for that apply: build the
        DDP wrapper on a side stream, run at least 11 eager DDP iterations before capture (`warmup >= 11`), and switch off
        NCCL's asynchronous error handling (TORCH_NCCL_ASYNC_ERROR_HANDLING=0) before the process group is created."""
        return GraphedTrainStep(model if model is not None else self.model, loss_fn, optimizer, sample_inputs, sample_target,
                                warmup)

    @staticmethod
    def get_original_model(model: nn.Module, qdq_weights: bool = False) -> nn.Module:
        """A copy of the model with the wrappers removed (reference :1502-1526); optionally with QDQ'd weights."""
        original = copy.deepcopy(model)

        def strip(parent):
            for name, child in list(parent.named_children()):
                if isinstance(child, LearnedGridQuantWrapper):
                    if qdq_weights:
                        with torch.no_grad():
                            for pname, param in child.get_named_parameters():
                                q = child.param_quantizers[pname]
                                if q.enabled and q.bitwidth != 32:
                                    param.data = q.quantize_dequantize(param.data,
                                                                       getattr(child, pname + "_encoding_min"),
                                                                       getattr(child, pname + "_encoding_max"))
                    setattr(parent, name, child.get_original_module())
                elif isinstance(child, StaticGridQuantWrapper):
                    if qdq_weights:
                        from .. import libpymo
                        for pname, param in child.get_named_parameters():
                            q = child.param_quantizers[pname]
                            if q.enabled and q.bitwidth != 32 and q._has_encoding():   # pylint: disable=protected-access
                                param.data = q.quantize_dequantize(param.data, libpymo.RoundingMode.ROUND_NEAREST)
                    setattr(parent, name, child.get_original_module())
                else:
                    strip(child)

        strip(original)
        return original


class _ParamExportPrefetch:
    """The parameter encodings of a calibration job exist after its FIRST forward (26 560 of them for per-channel ResNet-50),
    but are only asked for -- as Python dictionaries -- when the job is over, where building them is 9 ms of pure host time
    with the GPU idle. During the forward passes it is the host that waits (it issues a ResNet-50 step in 7 ms, the GPU
    needs 10), so the work is moved there: at the second forward the tables are gathered and copied to pinned memory with
    a stream-ordered, non-blocking copy; at a later forward, once that copy has landed, the dictionaries are built and
    parked on the quantizers, keyed by the identity of the device table they were made from (a recomputed encoding is a
    new tensor: a stale cache can never be served). `export_quantizer_encoding` hands them out once."""

    def __init__(self, sim):
        self._sim = sim
        self._forwards = 0
        self._pending = None
        self._done = False
        self._handle = None
        device = next((p.device for p in sim.model.parameters()), None)
        if device is not None and device.type == "cuda":          # host tensors (the oracle-backed test runs): nothing to do
            self._handle = sim.model.register_forward_pre_hook(self._on_forward)

    def close(self):
        if self._handle is not None:
            self._handle.remove()
        self._pending = None

    BUDGET_S = 2.0e-3      # host time spent on dictionaries per forward

    def _on_forward(self, _module, _inputs):
        import time
        self._forwards += 1
        if self._done or torch.cuda.is_current_stream_capturing():
            return
        if self._pending is None:
            # (the parameter encodings of planned quantizers exist before the first forward: quantsim.param_plan)
            from .tensor_quantizer import _LAZY
            qs = [q for _, w in self._sim.quant_wrappers() if isinstance(w, StaticGridQuantWrapper)
                  for q in w.param_quantizers.values()
                  if q.enabled and getattr(q, "_encoding", None) is _LAZY and q._enc_dev is not None]   # pylint: disable=protected-access
            if len(qs) < 2 or len({q._enc_dev.device for q in qs}) != 1:                                 # pylint: disable=protected-access
                self._done = self._forwards >= 2
                return
            tables = [q._enc_dev for q in qs]                                                            # pylint: disable=protected-access
            gathered = torch.cat(tables)
            host = torch.empty(gathered.shape, dtype=gathered.dtype, pin_memory=True)
            host.copy_(gathered, non_blocking=True)
            event = torch.cuda.Event()
            event.record()
            self._pending = [qs, tables, host, event, 0, 0]      # ..., next quantizer, its first row
            return
        # The first forwards of a job are issued with the device waiting for the host (the statistics batcher reads the
        # range flags back at the end of forward 0): building dictionaries there would stall the GPU. From the third
        # forward on the host runs ahead of the device again; a slice of the work per forward uses that lead up.
        if self._forwards < 3:
            return
        qs, tables, host, event, nxt, at = self._pending
        if not event.query():
            return
        rows = host.numpy()
        deadline = time.perf_counter() + self.BUDGET_S
        while nxt < len(qs):
            q, table = qs[nxt], tables[nxt]
            n = table.shape[0]
            if q._enc_dev is table:              # pylint: disable=protected-access  (else: recomputed meanwhile, leave it)
                a = rows[at:at + n]
                sym = str(q.use_symmetric_encodings)
                q._export_cache = (table, sym, [                                                         # pylint: disable=protected-access
                    {"min": mn, "max": mx, "scale": sc, "offset": off, "bitwidth": bw, "is_symmetric": sym, "dtype": "int"}
                    for mn, mx, sc, off, bw in zip(a[:, 0].tolist(), a[:, 1].tolist(), a[:, 2].tolist(),
                                                   a[:, 3].astype("int64").tolist(), a[:, 4].astype("int64").tolist())])
            at += n
            nxt += 1
            if time.perf_counter() > deadline:
                break
        self._pending[4], self._pending[5] = nxt, at
        if nxt >= len(qs):
            self._pending = None
            self._done = True


class GraphedForward:
    """See QuantizationSimModel.capture_forward."""

    def __init__(self, model: nn.Module, sample_inputs, warmup: int = 3):
        if not sample_inputs or not all(isinstance(t, torch.Tensor) and t.is_cuda for t in sample_inputs):
            raise ValueError("capture_forward needs CUDA tensors as sample inputs")
        self._model = model
        self._inputs = [t.detach().clone() for t in sample_inputs]
        was_training = model.training
        model.eval()
        try:
            with torch.no_grad():
                side = torch.cuda.Stream(device=self._inputs[0].device)
                side.wait_stream(torch.cuda.current_stream(self._inputs[0].device))
                with torch.cuda.stream(side):
                    for _ in range(max(1, warmup)):
                        model(*self._inputs)
                torch.cuda.current_stream(self._inputs[0].device).wait_stream(side)
                self._graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(self._graph):
                    self._outputs = model(*self._inputs)
        finally:
            model.train(was_training)

    def __call__(self, *inputs):
        if len(inputs) != len(self._inputs):
            raise ValueError(f"captured with {len(self._inputs)} inputs, called with {len(inputs)}")
        for static, new in zip(self._inputs, inputs):
            if static.shape != new.shape or static.dtype != new.dtype:
                raise ValueError("inputs must have the shapes and dtypes the forward was captured with")
            static.copy_(new, non_blocking=True)
        self._graph.replay()
        return self._outputs


class GraphedTrainStep:
    """See QuantizationSimModel.capture_train_step."""

    def __init__(self, model: nn.Module, loss_fn, optimizer, sample_inputs, sample_target, warmup: int = 3):
        if isinstance(sample_inputs, torch.Tensor):
            sample_inputs = (sample_inputs,)
        if not all(isinstance(t, torch.Tensor) and t.is_cuda for t in sample_inputs) or not sample_target.is_cuda:
            raise ValueError("capture_train_step needs CUDA tensors as sample inputs and target")
        self._model, self._optimizer = model, optimizer
        self._inputs = [t.detach().clone() for t in sample_inputs]
        self._target = sample_target.detach().clone()
        device = self._inputs[0].device
        # the warm-up steps below are real optimizer steps on the sample batch: remember everything they touch ...
        saved = [(t, t.detach().clone()) for t in list(model.parameters()) + list(model.buffers())]
        state_before = {id(p): {k: (v.detach().clone() if isinstance(v, torch.Tensor) else v) for k, v in st.items()}
                        for p, st in optimizer.state.items()}

        def one_step():
            optimizer.zero_grad(set_to_none=True)
            loss = loss_fn(model(*self._inputs), self._target)
            loss.backward()
            optimizer.step()
            return loss

        side = torch.cuda.Stream(device=device)
        side.wait_stream(torch.cuda.current_stream(device))
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):
                one_step()
        torch.cuda.current_stream(device).wait_stream(side)
        self._graph = torch.cuda.CUDAGraph()
        optimizer.zero_grad(set_to_none=True)
        with torch.cuda.graph(self._graph):
            self._loss = one_step()
        # ... and put it back IN PLACE (the graph holds the addresses): parameters, buffers, optimizer state. State the
        # optimizer created during the warm-up (momentum buffers) starts from zero.
        with torch.no_grad():
            for t, before in saved:
                t.copy_(before)
            for p, st in optimizer.state.items():
                old = state_before.get(id(p), {})
                for k, v in st.items():
                    if isinstance(v, torch.Tensor):
                        if isinstance(old.get(k), torch.Tensor):
                            v.copy_(old[k])
                        else:
                            v.zero_()

    def __call__(self, *inputs, target):
        if len(inputs) != len(self._inputs):
            raise ValueError(f"captured with {len(self._inputs)} inputs, called with {len(inputs)}")
        for static, new in zip(self._inputs, inputs):
            static.copy_(new, non_blocking=True)
        self._target.copy_(target, non_blocking=True)
        self._graph.replay()
        return self._loss


def save_checkpoint(quant_sim_model: QuantizationSimModel, file_path: str):
    """Pickle the whole sim (reference :2216-2228); quantizers drop their native objects and statistics, keep encodings."""
    import pickle
    with open(file_path, "wb") as f:
        pickle.dump(quant_sim_model, f)


def load_checkpoint(file_path: str) -> QuantizationSimModel:
    """reference :2231-2240"""
    import pickle
    with open(file_path, "rb") as f:
        return pickle.load(f)
