"""QuantAnalyzer on the B200 quantsim path (a caller that re-runs the hot path many times).

Host mirror of the reference's `aimet_torch.v1.quant_analyzer.QuantAnalyzer` (TEt/src/python/aimet_torch/v1/
quant_analyzer.py:65-763): same constructor, same public methods, same result dictionaries and the same JSON files under
`results_dir` (`per_layer_quant_enabled.json`, `per_layer_quant_disabled.json`, `min_max_ranges/{weights,activations}.json`,
`per_layer_mse_loss.json`). Every forward it triggers -- one calibration, then one evaluation per quant wrapper for each of
the two sweeps, then two truncated forwards per layer and batch for the MSE table -- runs the sm_100a QDQ / statistics
kernels through `QuantizationSimModel`.

Differences, both because the dependency is outside the hot path and absent from this image:
  * no bokeh: where the reference writes `.html` plots, the numbers behind them are written as JSON instead (the
    sensitivity, range and MSE tables are the reference's own JSON files; the per-quantizer histograms, which the
    reference only plots, go to `activations_pdf/<name>.json` / `weights_pdf/<module>/<name>.json`);
  * no batch-norm folding: the reference folds batch norms first (`fold_all_batch_norms`, a libpymo / ConnectedGraph
    feature, quant_analyzer.py:187-190). Fold them before handing the model over; a model that still has BatchNorm
    layers is analysed as it is, with a warning.
"""
import contextlib
import json
import logging
import os
from collections import OrderedDict
from typing import Callable, Dict, List, Tuple

import torch

from .defs import QuantScheme
from .qc_quantize_op import StaticGridQuantWrapper
from .quantsim import QuantizationSimModel

_logger = logging.getLogger("aimet_b200.QuantAnalyzer")


class CallbackFunc:
    """A callback and its argument, as `aimet_common.utils.CallbackFunc` (TEc/.../aimet_common/utils.py)."""

    def __init__(self, func: Callable, func_callback_args=None):
        self.func = func
        self.args = func_callback_args


class _StopForward(Exception):
    pass


def save_json(dictionary: Dict, results_dir: str, title: str):
    """aimet_common/quant_analyzer.py:82-91"""
    with open(os.path.join(results_dir, title), "w") as f:
        json.dump(dictionary, f, indent=4)


@contextlib.contextmanager
def _in_eval_mode(model: torch.nn.Module):
    was_training = {m: m.training for m in model.modules()}
    model.eval()
    try:
        yield
    finally:
        for m, t in was_training.items():
            m.training = t


def _leaf_modules(model: torch.nn.Module, wrapper_types=()):
    """Modules a forward hook is attached to: leaves, with a quant wrapper counting as one (its wrapped module is not
    visited) -- run_hook_for_layers_with_given_input, aimet_torch/utils.py:300-364."""
    skip = set()
    for m in model.modules():
        if m in skip:
            continue
        if wrapper_types and isinstance(m, wrapper_types):
            skip.update(m.modules())
            skip.discard(m)
            yield m
        elif len(list(m.children())) == 0:
            yield m


def _run_with_hooks(model, dummy_input, modules, hook):
    handles = [m.register_forward_hook(hook) for m in modules]
    try:
        with _in_eval_mode(model), torch.no_grad():
            if isinstance(dummy_input, (list, tuple)):
                model(*dummy_input)
            else:
                model(dummy_input)
    finally:
        for h in handles:
            h.remove()


def _output_of(model, module, model_inputs):
    """The output `module` produces inside `model` for these inputs; the forward stops there (utils.ModuleData)."""
    captured = []

    def hook(_, __, out):
        captured.append(out)
        raise _StopForward

    handle = module.register_forward_hook(hook)
    device = next((p.device for p in model.parameters()), None)

    def place(t):
        return t.to(device) if isinstance(t, torch.Tensor) and device is not None else t
    try:
        with _in_eval_mode(model), torch.no_grad():
            if isinstance(model_inputs, (list, tuple)):
                model(*[place(t) for t in model_inputs])
            else:
                model(place(model_inputs))
    except _StopForward:
        pass
    finally:
        handle.remove()
    out = captured[0] if captured else None
    return out.detach() if isinstance(out, torch.Tensor) else None


class QuantAnalyzer:
    """1) model sensitivity to weight / activation quantization, 2) per-layer sensitivity by enabling and by disabling quant
    wrappers, 3) per-layer encoding ranges, 4) per-layer statistics histograms, 5) per-layer MSE (reference :65-73)."""

    def __init__(self, model: torch.nn.Module, dummy_input, forward_pass_callback: CallbackFunc,
                 eval_callback: CallbackFunc, modules_to_ignore: List[torch.nn.Module] = None):
        if not isinstance(forward_pass_callback, CallbackFunc):
            raise ValueError('forward_pass_callback and its argument(s) are not encapsulated by CallbackFunc class.')
        if not isinstance(eval_callback, CallbackFunc):
            raise ValueError('eval_callback and its argument(s) are not encapsulated by CallbackFunc class.')
        self._model = model
        self._dummy_input = dummy_input
        self._forward_pass_callback = forward_pass_callback
        self._eval_callback = eval_callback
        self._unlabeled_dataset_iterable = None
        self._num_batches = None
        self._modules_to_ignore = modules_to_ignore

    # ---- top level (reference :105-171) ----------------------------------------------------------------------------
    def analyze(self, quant_scheme=QuantScheme.post_training_tf_enhanced, default_param_bw: int = 8,
                default_output_bw: int = 8, config_file: str = None, results_dir: str = "./tmp/"):
        if isinstance(quant_scheme, str):
            quant_scheme = QuantScheme.from_str(quant_scheme)
        sim = self._create_quantsim_and_encodings(quant_scheme, default_param_bw, default_output_bw, config_file)
        results_dir = os.path.abspath(results_dir)
        os.makedirs(results_dir, exist_ok=True)
        self.check_model_sensitivity_to_quantization(sim)
        self.perform_per_layer_analysis_by_enabling_quant_wrappers(sim, results_dir)
        self.perform_per_layer_analysis_by_disabling_quant_wrappers(sim, results_dir)
        self.export_per_layer_encoding_min_max_range(sim, results_dir)
        if quant_scheme == QuantScheme.post_training_tf_enhanced:
            self.export_per_layer_stats_histogram(sim, results_dir)
        if self._unlabeled_dataset_iterable:
            self.export_per_layer_mse_loss(sim, results_dir)
        return sim

    def enable_per_layer_mse_loss(self, unlabeled_dataset_iterable, num_batches: int):
        if len(unlabeled_dataset_iterable) < num_batches:
            raise ValueError(f'Can not fetch {num_batches} batches from '
                             f'a data loader of length {len(unlabeled_dataset_iterable)}.')
        self._unlabeled_dataset_iterable = unlabeled_dataset_iterable
        self._num_batches = num_batches

    def _create_quantsim_and_encodings(self, quant_scheme, default_param_bw, default_output_bw, config_file):
        if any(isinstance(m, torch.nn.modules.batchnorm._BatchNorm) for m in self._model.modules()):   # pylint: disable=protected-access
            _logger.warning("the model still has BatchNorm layers: the reference folds them before analysing "
                            "(fold_all_batch_norms); batch-norm folding is outside this package, fold them first")
        sim = QuantizationSimModel(self._model, self._dummy_input, quant_scheme=quant_scheme,
                                   default_output_bw=default_output_bw, default_param_bw=default_param_bw,
                                   config_file=config_file)
        if self._modules_to_ignore:
            self._exclude_modules_from_quantization(self._model, sim, self._modules_to_ignore)
        sim.compute_encodings(self._forward_pass_callback.func, self._forward_pass_callback.args)
        return sim

    # ---- evaluation helpers (reference :205-239) -------------------------------------------------------------------
    def _eval_model(self, model: torch.nn.Module) -> float:
        with _in_eval_mode(model), torch.no_grad():
            return self._eval_callback.func(model, self._eval_callback.args)

    def _eval_weight_quantized_model(self, sim) -> float:
        with self._disable_activation_quantizers(sim):
            return self._eval_model(sim.model)

    def _eval_activation_quantized_model(self, sim) -> float:
        with self._disable_param_quantizers(sim):
            return self._eval_model(sim.model)

    def check_model_sensitivity_to_quantization(self, sim) -> Tuple[float, float, float]:
        """FP32, weight-quantized and activation-quantized eval scores (reference :409-431)."""
        fp32_eval_score = self._eval_model(self._model)
        _logger.info("FP32 eval score (W32A32): %f", fp32_eval_score)
        weight_quantized_eval_score = self._eval_weight_quantized_model(sim)
        _logger.info("Weight-quantized eval score (W%dA32): %f", sim._default_param_bw, weight_quantized_eval_score)   # pylint: disable=protected-access
        act_quantized_eval_score = self._eval_activation_quantized_model(sim)
        _logger.info("Activation-quantized eval score (W32A%d): %f", sim._default_output_bw, act_quantized_eval_score)   # pylint: disable=protected-access
        return fp32_eval_score, weight_quantized_eval_score, act_quantized_eval_score

    # ---- per-layer sweeps (reference :241-385, 433-504) ------------------------------------------------------------
    def _sort_quant_wrappers_based_on_occurrence(self, sim) -> "OrderedDict[str, StaticGridQuantWrapper]":
        names = {m: n for n, m in sim.model.named_modules()}
        ordered = OrderedDict()

        def hook(wrapper, *_):
            ordered[names[wrapper]] = wrapper

        wrappers = [m for m in sim.model.modules() if isinstance(m, self._get_quant_wrapper_type())]
        _run_with_hooks(sim.model, self._dummy_input, wrappers, hook)
        return ordered

    @classmethod
    def _get_enabled_quantizers(cls, sorted_quant_wrappers: Dict) -> Dict:
        enabled = OrderedDict()
        for wrapper in sorted_quant_wrappers.values():
            qs = [q for q in wrapper.param_quantizers.values() if cls._is_quantizer_enabled(q)]
            qs += [q for q in wrapper.output_quantizers if cls._is_quantizer_enabled(q)]
            qs += [q for q in wrapper.input_quantizers if cls._is_quantizer_enabled(q)]
            if qs:
                enabled[wrapper] = qs
        return enabled

    @classmethod
    def _get_enabled_param_quantizers(cls, sim) -> List:
        return [q for w in cls._get_quantized_modules(sim) for q in w.param_quantizers.values()
                if cls._is_quantizer_enabled(q)]

    @classmethod
    def _get_enabled_activation_quantizers(cls, sim) -> List:
        return [q for w in cls._get_quantized_modules(sim) for q in list(w.input_quantizers) + list(w.output_quantizers)
                if cls._is_quantizer_enabled(q)]

    @staticmethod
    def _enable_disable_quantizers(quantizers: List, enabled: bool):
        for q in quantizers:
            q.enabled = enabled

    def _perform_per_layer_analysis(self, sim, disable_all_quantizers: bool, enabled_before: bool,
                                    enabled_after: bool) -> Dict:
        assert (disable_all_quantizers, enabled_before, enabled_after) in ((True, True, False), (False, False, True))
        sorted_quant_wrappers = self._sort_quant_wrappers_based_on_occurrence(sim)
        enabled_quant_wrappers = self._get_enabled_quantizers(sorted_quant_wrappers)
        eval_score_dict = {}
        for name, wrapper in sorted_quant_wrappers.items():
            if wrapper not in enabled_quant_wrappers:
                continue
            with contextlib.ExitStack() as stack:
                if disable_all_quantizers and enabled_before:
                    for other in enabled_quant_wrappers:          # everything but this wrapper off
                        if other is not wrapper:
                            stack.enter_context(self._disable_quant_wrapper(other))
                else:
                    stack.enter_context(self._disable_quant_wrapper(wrapper))   # only this wrapper off
                eval_score_dict[name] = self._eval_model(sim.model)
        return eval_score_dict

    def perform_per_layer_analysis_by_enabling_quant_wrappers(self, sim, results_dir: str) -> Dict:
        """Option 1: everything disabled, one wrapper's quantizers enabled at a time (reference :433-467)."""
        results_dir = os.path.abspath(results_dir)
        os.makedirs(results_dir, exist_ok=True)
        scores = self._perform_per_layer_analysis(sim, disable_all_quantizers=True, enabled_before=True,
                                                  enabled_after=False)
        save_json(scores, results_dir, title="per_layer_quant_enabled.json")
        return scores

    def perform_per_layer_analysis_by_disabling_quant_wrappers(self, sim, results_dir: str) -> Dict:
        """Option 2: everything enabled, one wrapper's quantizers disabled at a time (reference :469-504)."""
        results_dir = os.path.abspath(results_dir)
        os.makedirs(results_dir, exist_ok=True)
        scores = self._perform_per_layer_analysis(sim, disable_all_quantizers=False, enabled_before=False,
                                                  enabled_after=True)
        save_json(scores, results_dir, title="per_layer_quant_disabled.json")
        return scores

    # ---- exports (reference :506-651) ------------------------------------------------------------------------------
    def export_per_layer_encoding_min_max_range(self, sim, results_dir: str) -> Tuple[Dict, Dict]:
        min_max_ranges_dir = os.path.join(results_dir, "min_max_ranges")
        os.makedirs(min_max_ranges_dir, exist_ok=True)
        names = {m: n for n, m in sim.model.named_modules()}
        activations, weights = {}, {}
        for wrapper in self._get_quantized_modules(sim):
            wname = names[wrapper]
            for kind, quantizers in (("input", wrapper.input_quantizers), ("output", wrapper.output_quantizers)):
                for index, q in enumerate(quantizers):
                    if self._is_quantizer_enabled(q):
                        enc = self._get_quantizer_encodings(q)[0]
                        activations[f"{wname}_{kind}_{index}"] = (enc.min, enc.max)
            for pname, q in wrapper.param_quantizers.items():
                if self._is_quantizer_enabled(q):
                    name = f"{wname}_{pname}"
                    encs = self._get_quantizer_encodings(q)
                    if len(encs) > 1:
                        weights[name] = {f"{name}_{i}": (e.min, e.max) for i, e in enumerate(encs)}
                    else:
                        weights[name] = (encs[0].min, encs[0].max)
        save_json(weights, min_max_ranges_dir, title="weights.json")
        save_json(activations, min_max_ranges_dir, title="activations.json")
        return weights, activations

    def export_per_layer_stats_histogram(self, sim, results_dir: str):
        """tf_enhanced only. The reference plots every quantizer's 512-bin PDF with its encoding (:572-617); here the
        same numbers are written as JSON: {"histogram": [[xLeft, pdf] * 512], "encoding": {min, max, delta, offset, bw}}."""
        weights_pdf_dir = os.path.join(results_dir, "weights_pdf")
        activations_pdf_dir = os.path.join(results_dir, "activations_pdf")
        names = {m: n for n, m in sim.model.named_modules()}
        for wrapper in self._get_quantized_modules(sim):
            wname = names[wrapper]
            for index, q in enumerate(wrapper.input_quantizers):
                if q is not None and self._get_quantizer_encodings(q):
                    self._export_stats_histogram(q, activations_pdf_dir, f"{wname}_input_q{index}")
            for index, q in enumerate(wrapper.output_quantizers):
                if q is not None and self._get_quantizer_encodings(q):
                    self._export_stats_histogram(q, activations_pdf_dir, f"{wname}_output_q{index}")
            for pname, q in wrapper.param_quantizers.items():
                if q is not None and self._get_quantizer_encodings(q):
                    self._export_stats_histogram(q, os.path.join(weights_pdf_dir, wname), f"{wname}_{pname}")

    def _export_stats_histogram(self, quantizer, results_dir: str, title: str):
        os.makedirs(results_dir, exist_ok=True)
        histograms = quantizer.get_stats_histogram()
        encodings = self._get_quantizer_encodings(quantizer)
        for index, (histogram, enc) in enumerate(zip(histograms, encodings)):
            save_json({"histogram": [list(b) for b in histogram],
                       "encoding": {"min": enc.min, "max": enc.max, "delta": enc.delta, "offset": enc.offset, "bw": enc.bw}},
                      results_dir, title=f"{title}_{index}.json")

    def export_per_layer_mse_loss(self, sim, results_dir: str) -> Dict:
        """MSE between the fp32 model's and the quantsim model's output of every layer (reference :619-688)."""
        results_dir = os.path.abspath(results_dir)
        os.makedirs(results_dir, exist_ok=True)
        sim_modules = dict(sim.model.named_modules())
        fp32_names = {m: n for n, m in self._model.named_modules()}
        ordered = []
        _run_with_hooks(self._model, self._dummy_input, list(_leaf_modules(self._model)),
                        lambda m, *_: ordered.append((fp32_names[m], m)))
        mse_loss_dict = {}
        for name, module in ordered:
            mse_loss_dict[name] = self._compute_mse_loss(module, sim_modules[name], self._model, sim)
        save_json(mse_loss_dict, results_dir, title="per_layer_mse_loss.json")
        return mse_loss_dict

    def _compute_mse_loss(self, module, quant_wrapper, fp32_model, sim) -> float:
        total, loss, batch_index = 0, 0.0, 0
        for model_inputs in self._unlabeled_dataset_iterable:
            assert isinstance(model_inputs, (torch.Tensor, tuple, list))
            quantized_out = _output_of(sim.model, quant_wrapper, model_inputs)
            fp32_out = _output_of(fp32_model, module, model_inputs)
            loss += torch.nn.functional.mse_loss(fp32_out, quantized_out).item()
            total += fp32_out.size(0)
            batch_index += 1
            if batch_index == self._num_batches:
                break
        return loss / total

    # ---- plumbing (reference :690-763) -----------------------------------------------------------------------------
    @staticmethod
    def _exclude_modules_from_quantization(model, sim, modules_to_ignore: List[torch.nn.Module]):
        sim_modules = dict(sim.model.named_modules())
        names = {m: n for n, m in model.named_modules()}
        sim.exclude_layers_from_quantization([sim_modules[names[m]] for m in modules_to_ignore])

    @staticmethod
    def _get_quantsim_cls():
        return QuantizationSimModel

    @staticmethod
    def _get_quant_wrapper_type():
        return (StaticGridQuantWrapper,)

    @staticmethod
    def _is_quantizer_enabled(quantizer) -> bool:
        return quantizer.enabled

    @staticmethod
    def _get_quantizer_encodings(quantizer):
        enc = quantizer.encoding
        if enc and not isinstance(enc, list):
            return [enc]
        return enc

    @classmethod
    @contextlib.contextmanager
    def _disable_param_quantizers(cls, sim):
        quantizers = cls._get_enabled_param_quantizers(sim)
        cls._enable_disable_quantizers(quantizers, enabled=False)
        yield
        cls._enable_disable_quantizers(quantizers, enabled=True)

    @classmethod
    @contextlib.contextmanager
    def _disable_activation_quantizers(cls, sim):
        quantizers = cls._get_enabled_activation_quantizers(sim)
        cls._enable_disable_quantizers(quantizers, enabled=False)
        yield
        cls._enable_disable_quantizers(quantizers, enabled=True)

    @staticmethod
    @contextlib.contextmanager
    def _disable_quant_wrapper(module):
        """utils.disable_all_quantizers (aimet_torch/utils.py:1001-1028) for one wrapper."""
        active = [q for q in list(module.param_quantizers.values()) + list(module.input_quantizers) +
                  list(module.output_quantizers) if q.enabled]
        for q in active:
            q.enabled = False
        try:
            yield
        finally:
            for q in active:
                q.enabled = True

    @classmethod
    def _get_quantized_modules(cls, sim):
        for module in sim.model.modules():
            if isinstance(module, cls._get_quant_wrapper_type()):
                yield module
