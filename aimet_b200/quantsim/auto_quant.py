"""AutoQuant on the CUDA hot path (reference aimet_torch/v1/auto_quant.py:204-816, 1386-1466).

AutoQuant is a CALLER of the quantization-simulation path: it builds a QuantizationSimModel and runs
compute_encodings once per quant-scheme candidate (five by default), once for the W32 check and once per post-training
technique, which is why it belongs next to QuantAnalyzer and AdaRound. This mirror keeps the reference's public surface
and decision logic -- `AutoQuant(model, dummy_input, data_loader, eval_callback, ...)`, `run_inference()`,
`optimize(allowed_accuracy_drop)`, `set_adaround_params`, `get/set_quant_scheme_candidates`, the "best result so far"
bookkeeping of the evaluation sessions and their error handling (`strict_validation`) -- on `aimet_b200.quantsim`.

Stages, in the reference's order (`_optimize_main`, :714-816):
  1. quant-scheme selection: every candidate pair (parameter scheme, activation scheme[, percentile]) is calibrated and
     evaluated, the best one becomes the default (:670-712);
  2. W32 evaluation: parameters in fp32, activations quantized; below the target -> give up early (:733-757);
  3. batch-norm folding (`aimet_b200.quantsim.batch_norm_fold`);
  4. cross-layer equalization -- NOT built here: CLE is a weight-rewriting subsystem of its own (out of
     scope). The stage takes a user-supplied callable (`set_cross_layer_equalization_fn`); without one it is skipped the
     way the reference skips a stage that raises under `strict_validation=False` (status "discarded");
  5. AdaRound on the best model so far (`aimet_b200.quantsim.adaround`).
Not carried over: the HTML diagnostics (jinja2 / bokeh), the on-disk cache of stage results, model preparer / validator
(torch.fx rewriting of the user's model), ONNX export arguments, AutoQuantWithAutoMixedPrecision.
"""
import contextlib
import copy
import functools
import itertools
import math
import os
import traceback
from collections import OrderedDict
from dataclasses import dataclass
from typing import Any, Callable, Dict, List, Optional, Tuple, Union

import torch

from .adaround import Adaround, AdaroundParameters
from .batch_norm_fold import fold_all_batch_norms
from .defs import QuantScheme
from .quantsim import QuantizationSimModel

NUM_SAMPLES_FOR_PERFORMANCE_EVALUATION = None   # reference :94: the whole evaluation set


class _StageSkipped(Exception):
    pass


@dataclass(frozen=True)
class _QuantSchemePair:
    param_quant_scheme: QuantScheme
    output_quant_scheme: QuantScheme
    param_percentile: Optional[float] = None
    output_percentile: Optional[float] = None

    def __str__(self):
        def scheme_to_str(quant_scheme, percentile):
            if quant_scheme == QuantScheme.post_training_percentile:
                return f"{percentile}%ile"
            if quant_scheme in (QuantScheme.post_training_tf, QuantScheme.training_range_learning_with_tf_init):
                return "tf"
            if quant_scheme in (QuantScheme.post_training_tf_enhanced,
                                QuantScheme.training_range_learning_with_tf_enhanced_init):
                return "tf-enhanced"
            raise ValueError
        return (f"W@{scheme_to_str(self.param_quant_scheme, self.param_percentile)} / "
                f"A@{scheme_to_str(self.output_quant_scheme, self.output_percentile)}")


# reference :123-152
_QUANT_SCHEME_CANDIDATES = (
    _QuantSchemePair(QuantScheme.post_training_tf, QuantScheme.post_training_tf),
    _QuantSchemePair(QuantScheme.post_training_tf_enhanced, QuantScheme.post_training_tf),
    _QuantSchemePair(QuantScheme.post_training_tf_enhanced, QuantScheme.post_training_tf_enhanced),
    _QuantSchemePair(QuantScheme.post_training_tf_enhanced, QuantScheme.post_training_percentile, output_percentile=99.9),
    _QuantSchemePair(QuantScheme.post_training_tf_enhanced, QuantScheme.post_training_percentile, output_percentile=99.99),
)


@contextlib.contextmanager
def in_eval_mode(model: torch.nn.Module):
    modes = {m: m.training for m in model.modules()}
    model.eval()
    try:
        yield
    finally:
        for m, training in modes.items():
            m.training = training


def _device_of(model: torch.nn.Module) -> torch.device:
    return next(model.parameters()).device


def _to_device(data, device):
    if isinstance(data, torch.Tensor):
        return data.to(device)
    if isinstance(data, (tuple, list)):
        return type(data)(_to_device(d, device) for d in data)
    return data


def get_all_quantizers(model: torch.nn.Module):
    """(param, input, output) quantizers of every wrapper (reference aimet_torch/utils.py get_all_quantizers)."""
    from .qc_quantize_op import QcQuantizeWrapper
    from .learned_grid import LearnedGridQuantWrapper
    params, inputs, outputs = [], [], []
    for m in model.modules():
        if isinstance(m, (QcQuantizeWrapper, LearnedGridQuantWrapper)):
            params.extend(m.param_quantizers.values())
            inputs.extend(m.input_quantizers)
            outputs.extend(m.output_quantizers)
    return params, inputs, outputs


def _validate_inputs(model, data_loader, eval_callback, dummy_input, results_dir, strict_validation, quant_scheme, param_bw,
                     output_bw, rounding_mode):
    """reference :155-201"""
    if not isinstance(model, torch.nn.Module):
        raise ValueError('Model must be of type torch.nn.Module, not ' + str(type(model).__name__))
    if not hasattr(data_loader, "__iter__") or not hasattr(data_loader, "__len__"):
        raise ValueError('data_loader must be a sized iterable (torch DataLoader), not ' + str(type(data_loader).__name__))
    if not callable(eval_callback):
        raise ValueError('eval_callback must be of type Callable, not ' + str(type(eval_callback).__name__))
    if not isinstance(dummy_input, (torch.Tensor, tuple, list)):
        raise ValueError('dummy_input must be of type torch.Tensor or Tuple, not ' + str(type(dummy_input).__name__))
    if not isinstance(results_dir, str):
        raise ValueError('results_dir must be of type str, not ' + str(type(results_dir).__name__))
    if not isinstance(strict_validation, bool):
        raise ValueError('strict_validation must be of type bool, not ' + str(type(strict_validation).__name__))
    if param_bw <= 0 or param_bw > 32:
        raise ValueError('param_bw must be an integer in [1, 32], not ' + str(param_bw))
    if output_bw <= 0 or output_bw > 32:
        raise ValueError('output_bw must be an integer in [1, 32], not ' + str(output_bw))
    if not isinstance(quant_scheme, QuantScheme):
        raise ValueError('quant_scheme must be of type QuantScheme, not ' + str(type(quant_scheme).__name__))
    if rounding_mode not in ("nearest", "stochastic"):
        raise ValueError('rounding_mode must be "nearest" or "stochastic", not ' + str(rounding_mode))


@dataclass
class PtqResult:
    """One evaluated post-training result (reference :818-845). The model is kept on disk, as in the reference."""
    model_path: str
    device: torch.device
    encoding_path: str
    accuracy: float
    applied_techniques: List[str]

    def load_model(self) -> torch.nn.Module:
        return torch.load(self.model_path, weights_only=False).to(self.device)

    def as_dict(self):
        return dict(model=self.load_model(), accuracy=self.accuracy, encoding_path=self.encoding_path,
                    applied_techniques=self.applied_techniques)


class _EvalSession:
    """One named stage (reference :962-1229): swallows `_StageSkipped` always and other exceptions unless
    strict_validation; keeps at most one PtqResult."""

    def __init__(self, title, quantsim_factory, eval_func, results_dir, strict_validation, ptq):
        self.title = title
        self._quantsim_factory = quantsim_factory
        self._eval_func = eval_func
        self._results_dir = results_dir
        self._strict_validation = strict_validation
        self._ptq = ptq
        self.result = {"status": None, "error": None, "target_satisfied": False, "effective": True}
        self.title_lowercase = "_".join(self.title.lower().replace("-", " ").split())
        self._ptq_result = None
        self._cached_result = None
        os.makedirs(self._results_dir, exist_ok=True)

    def is_ptq_session(self):
        return self._ptq

    def reset_status(self):
        self.result = {"status": None, "error": None, "target_satisfied": False, "effective": True}

    def wrap(self, fn):
        """The stage's function, evaluated at most once per AutoQuant object (reference :1042-1072 pickles the value)."""
        @functools.wraps(fn)
        def wrapper(*args, **kwargs):
            if self._cached_result is not None:
                return copy.deepcopy(self._cached_result[0])
            ret = fn(*args, **kwargs)
            self._cached_result = (copy.deepcopy(ret),)
            return ret
        return wrapper

    def eval(self, model: torch.nn.Module, **kwargs):
        sim = self._quantsim_factory(model, **kwargs)
        return self._eval_func(sim.model)

    def __enter__(self):
        return self

    def __exit__(self, exc_type, exc_val, exc_tb):
        if exc_val:
            if exc_type == _StageSkipped:
                print(exc_val.args[0])
            else:
                text = "".join(traceback.format_exception(exc_type, exc_val, exc_tb))
                print(text if self._strict_validation else
                      "WARNING: The following exception was raised but ignored:\n\n" + text)
        self.result["error"] = exc_val
        if not exc_val:
            self.result["status"] = "success"
        elif exc_type == _StageSkipped:
            self.result["status"] = "discarded"
            return True
        elif self._strict_validation:
            self.result["status"] = "error-failed"
        else:
            self.result["status"] = "error-ignored"
        if exc_val and not self._strict_validation:
            return True
        return None

    @property
    def ptq_result(self) -> Optional[PtqResult]:
        return self._ptq_result

    def set_ptq_result(self, applied_techniques: List[str], model: torch.nn.Module = None, sim: QuantizationSimModel = None,
                       acc: float = None, export_kwargs=None, **kwargs) -> None:
        """Exactly one of `model` and (`sim`, `acc`) (reference :1143-1178)."""
        del export_kwargs
        if sim is None:
            assert acc is None
            assert model is not None
            sim = self._quantsim_factory(model, **kwargs)
            acc = self._eval_func(sim.model)
        else:
            assert acc is not None
            assert model is None
        if self._ptq_result is not None:
            raise RuntimeError("sess.eval() can be called only once per each _EvalSession instance.")
        device = _device_of(sim.model)
        sim.export(path=self._results_dir, filename_prefix=self.title_lowercase)
        model_path = os.path.join(self._results_dir, f"{self.title_lowercase}.pth")
        # the module-name keyed encodings file of this repo's export (the ONNX-name keyed one needs an ONNX export)
        encoding_path = os.path.join(self._results_dir, f"{self.title_lowercase}_torch.encodings")
        self._ptq_result = PtqResult(model_path=model_path, device=device, encoding_path=encoding_path, accuracy=acc,
                                     applied_techniques=applied_techniques)


class _EvalManager:
    """reference :848-959 without the HTML report"""

    def __init__(self, quantsim_factory, eval_func, results_dir, strict_validation):
        self._quantsim_factory = quantsim_factory
        self._eval_func = eval_func
        self._results_dir = results_dir
        self._strict_validation = strict_validation
        os.makedirs(self._results_dir, exist_ok=True)
        self._all_sessions = OrderedDict()

    def clear(self):
        for sess in self._all_sessions.values():
            sess.reset_status()

    def get_best_ptq_result(self) -> Optional[PtqResult]:
        results = [s.ptq_result for s in self._all_sessions.values() if s.ptq_result is not None]
        if not results:
            return None
        return max(results, key=lambda r: r.accuracy)

    def session(self, title: str, ptq: bool = False) -> _EvalSession:
        if title not in self._all_sessions:
            self._all_sessions[title] = _EvalSession(title, self._quantsim_factory, self._eval_func,
                                                     os.path.join(self._results_dir, ".trace"), self._strict_validation, ptq)
        return self._all_sessions[title]

    def summary(self) -> Dict[str, Dict]:
        """What the reference renders into diagnostics.html, as a dictionary."""
        out = OrderedDict()
        for sess in self._all_sessions.values():
            entry = dict(sess.result)
            entry["error"] = None if entry["error"] is None else repr(entry["error"])
            if sess.ptq_result is not None:
                entry["accuracy"] = sess.ptq_result.accuracy
                entry["applied_techniques"] = list(sess.ptq_result.applied_techniques)
            out[sess.title_lowercase] = entry
        return out


class AutoQuant:   # pylint: disable=too-many-instance-attributes
    """Integrate and apply post-training quantization techniques: 1) batch-norm folding, 2) cross-layer equalization (if a
    callable for it is supplied), 3) AdaRound, applied in a best-effort manner until the model meets the evaluation goal
    given as allowed_accuracy_drop (reference :204-212)."""

    def __init__(self, model: torch.nn.Module, dummy_input: Union[torch.Tensor, Tuple], data_loader,   # pylint: disable=too-many-arguments
                 eval_callback: Callable[[torch.nn.Module], float], param_bw: int = 8, output_bw: int = 8,
                 quant_scheme: QuantScheme = QuantScheme.post_training_tf_enhanced, rounding_mode: str = 'nearest',
                 config_file: str = None, results_dir: str = "/tmp", cache_id: str = None, strict_validation: bool = True,
                 model_prepare_required: bool = False) -> None:
        _validate_inputs(model, data_loader, eval_callback, dummy_input, results_dir, strict_validation, quant_scheme,
                         param_bw, output_bw, rounding_mode)
        del cache_id   # the on-disk stage cache is not carried over
        if model_prepare_required:
            raise NotImplementedError("the model preparer is not part of this path: pass an fx-traceable model and "
                                      "model_prepare_required=False")
        self.fp32_model = model
        self.dummy_input = dummy_input
        self.data_loader = data_loader
        self._quantsim_params = dict(param_bw=param_bw, output_bw=output_bw,
                                     quant_scheme=_QuantSchemePair(quant_scheme, quant_scheme),
                                     rounding_mode=rounding_mode, config_file=config_file)
        self.results_dir = results_dir

        def forward_pass_callback(model, _: Any = None):
            device = _device_of(model)
            with in_eval_mode(model), torch.no_grad():
                for input_data in data_loader:
                    input_data = _to_device(input_data, device)
                    if isinstance(input_data, torch.Tensor):
                        model(input_data)
                    else:
                        assert isinstance(input_data, (tuple, list))
                        model(*input_data)

        self.forward_pass_callback = forward_pass_callback

        @functools.wraps(eval_callback)
        def eval_callback_wrapper(model: torch.nn.Module, *args, **kwargs) -> float:
            with in_eval_mode(model), torch.no_grad():
                return eval_callback(model, *args, **kwargs)

        self.eval_callback = eval_callback_wrapper

        # at most 2000 samples for AdaRound (reference :282-287)
        dataset = getattr(self.data_loader, "dataset", None)
        batch_size = getattr(self.data_loader, "batch_size", None) or 1
        num_samples = min(len(dataset) if dataset is not None else len(self.data_loader) * batch_size, 2000)
        num_batches = min(math.ceil(num_samples / batch_size), len(self.data_loader))
        self.adaround_params = AdaroundParameters(self.data_loader, num_batches)
        self._export_kwargs = {}
        self._cle_fn = None
        self.eval_manager = _EvalManager(quantsim_factory=self._create_quantsim_and_encodings,
                                         eval_func=self._evaluate_model_performance, results_dir=self.results_dir,
                                         strict_validation=strict_validation)
        self._quant_scheme_candidates = _QUANT_SCHEME_CANDIDATES
        self._fp32_acc = None

    # ---- knobs ---------------------------------------------------------------------------------------------------------
    def set_adaround_params(self, adaround_params: AdaroundParameters) -> None:
        self.adaround_params = adaround_params

    def set_export_params(self, onnx_export_args=-1, propagate_encodings: bool = None) -> None:
        """Accepted for API compatibility; this repo's export writes no ONNX file (reference :388-405)."""
        if onnx_export_args != -1:
            self._export_kwargs["onnx_export_args"] = onnx_export_args
        if propagate_encodings is not None:
            self._export_kwargs["propagate_encodings"] = propagate_encodings

    def set_cross_layer_equalization_fn(self, fn: Optional[Callable[[torch.nn.Module], torch.nn.Module]]) -> None:
        """`fn(model_copy) -> equalized model` for the CLE stage; None (default) skips the stage."""
        self._cle_fn = fn

    def get_quant_scheme_candidates(self) -> Tuple[_QuantSchemePair, ...]:
        return self._quant_scheme_candidates

    def set_quant_scheme_candidates(self, candidates: Tuple[_QuantSchemePair, ...]):
        self._quant_scheme_candidates = copy.copy(candidates)

    # ---- the public flow -----------------------------------------------------------------------------------------------
    def _evaluate_model_performance(self, model) -> float:
        return self.eval_callback(model, NUM_SAMPLES_FOR_PERFORMANCE_EVALUATION)

    def run_inference(self) -> Tuple[QuantizationSimModel, float]:
        """Batch-norm folding + a calibrated sim and its score (reference :335-364)."""
        model = self.fp32_model
        with self.eval_manager.session("Batchnorm Folding", ptq=True) as sess:
            model, _ = sess.wrap(self._apply_batchnorm_folding)(model)
            if sess.ptq_result is None:
                sess.set_ptq_result(model=model, applied_techniques=["batchnorm_folding"], export_kwargs=self._export_kwargs)
        sim = self._create_quantsim_and_encodings(model)
        if sess.ptq_result is None:
            acc = self._evaluate_model_performance(sim.model)   # folding failed: measure
        else:
            acc = sess.ptq_result.accuracy
        return sim, acc

    def optimize(self, allowed_accuracy_drop: float = 0.0) -> Tuple[torch.nn.Module, float, str]:
        """(best model, eval score, encoding path) (reference :366-376)."""
        result = self._optimize_helper(self._optimize_main, allowed_accuracy_drop)
        return result["model"], result["accuracy"], result["encoding_path"]

    # ---- sims ----------------------------------------------------------------------------------------------------------
    def _create_quantsim_and_encodings(self, model: torch.nn.Module, rounding_mode: str = None, output_bw: int = None,   # pylint: disable=too-many-arguments
                                       output_quant_scheme: QuantScheme = None, output_percentile: float = None,
                                       param_bw: int = None, param_quant_scheme: QuantScheme = None,
                                       param_percentile: float = None, config_file: str = None,
                                       encoding_path: str = None) -> QuantizationSimModel:
        """reference :427-502: explicit arguments override the defaults chosen so far; an `encoding_path` freezes the
        parameter encodings found there."""
        if output_bw is not None:
            assert output_bw <= 32
        if param_bw is not None:
            assert param_bw <= 32
        if output_quant_scheme is None or param_quant_scheme is None:
            assert self._quantsim_params["quant_scheme"] is not None
        kwargs = dict(rounding_mode=(rounding_mode or self._quantsim_params["rounding_mode"]),
                      default_output_bw=(output_bw or self._quantsim_params["output_bw"]),
                      default_param_bw=(param_bw or self._quantsim_params["param_bw"]),
                      config_file=(config_file or self._quantsim_params["config_file"]))
        sim = QuantizationSimModel(model, self.dummy_input, **kwargs)
        default = self._quantsim_params.get("quant_scheme")
        if default is not None:
            output_quant_scheme = output_quant_scheme or default.output_quant_scheme
            output_percentile = output_percentile or default.output_percentile
            param_quant_scheme = param_quant_scheme or default.param_quant_scheme
            param_percentile = param_percentile or default.param_percentile
        self._configure_quantsim(sim, output_bw, output_quant_scheme, output_percentile, param_bw, param_quant_scheme,
                                 param_percentile, encoding_path)
        if self._has_enabled_quantizers(sim):
            sim.compute_encodings(self.forward_pass_callback, None)
        return sim

    @staticmethod
    def _configure_quantsim(sim, output_bw, output_quant_scheme, output_percentile, param_bw, param_quant_scheme,   # pylint: disable=too-many-arguments
                            param_percentile, encoding_path):
        """reference :1409-1448"""
        param_quantizers, input_quantizers, output_quantizers = get_all_quantizers(sim.model)
        for quantizer in itertools.chain(input_quantizers, output_quantizers):
            quantizer.quant_scheme = output_quant_scheme
            if quantizer.quant_scheme == QuantScheme.post_training_percentile and output_percentile is not None:
                quantizer.set_percentile_value(output_percentile)
        for quantizer in param_quantizers:
            quantizer.quant_scheme = param_quant_scheme
            if quantizer.quant_scheme == QuantScheme.post_training_percentile and param_percentile is not None:
                quantizer.set_percentile_value(param_percentile)
        if encoding_path:
            sim.set_and_freeze_param_encodings(encoding_path)
        param_quantizers, input_quantizers, output_quantizers = get_all_quantizers(sim.model)
        if output_bw == 32:      # fp32 stands in for int32
            for quantizer in input_quantizers + output_quantizers:
                quantizer.enabled = False
        if param_bw == 32:
            for quantizer in param_quantizers:
                quantizer.enabled = False

    @staticmethod
    def _has_enabled_quantizers(sim):
        return any(q.enabled for q in itertools.chain(*get_all_quantizers(sim.model)))

    @staticmethod
    def _disable_activation_quantizers(sim):
        _, input_quantizers, output_quantizers = get_all_quantizers(sim.model)
        for quantizer in itertools.chain(input_quantizers, output_quantizers):
            quantizer.enabled = False

    # ---- techniques (the input model is never mutated) ---------------------------------------------------------------
    def _apply_batchnorm_folding(self, model: torch.nn.Module) -> Tuple[torch.nn.Module, List[Tuple]]:
        model = copy.deepcopy(model)
        folded_pairs = fold_all_batch_norms(model, None, self.dummy_input)
        return model, folded_pairs

    def _apply_cross_layer_equalization(self, model: torch.nn.Module) -> torch.nn.Module:
        if self._cle_fn is None:
            raise _StageSkipped("Skipping Cross-Layer Equalization (no equalization function was supplied; see "
                                "set_cross_layer_equalization_fn)")
        return self._cle_fn(copy.deepcopy(model))

    def _apply_adaround(self, model: torch.nn.Module) -> Tuple[torch.nn.Module, str]:
        """reference :576-600; AdaRound itself deep-copies the model"""
        filename_prefix = "adaround"
        adaround_encoding_path = os.path.join(self.results_dir, f"{filename_prefix}.encodings")
        sim = self._create_quantsim_and_encodings(model)
        self._disable_activation_quantizers(sim)
        model = Adaround._apply_adaround(sim, model, self.dummy_input, self.adaround_params, path=self.results_dir,   # pylint: disable=protected-access
                                         filename_prefix=filename_prefix)
        return model, adaround_encoding_path

    # ---- the search ----------------------------------------------------------------------------------------------------
    def _optimize_helper(self, optimize_fn: Callable, allowed_accuracy_drop: float) -> Dict[str, Any]:
        allowed_accuracy_drop = float(allowed_accuracy_drop)
        if allowed_accuracy_drop < 0:
            raise ValueError("`allowed_accuracy_drop` must be a positive value. Got {:.2f}".format(allowed_accuracy_drop))
        self.eval_manager.clear()
        with in_eval_mode(self.fp32_model):
            self._fp32_acc = self._evaluate_model_performance(self.fp32_model)
            target_acc = self._fp32_acc - allowed_accuracy_drop
            return optimize_fn(self.fp32_model, target_acc)

    def _choose_default_quant_scheme(self) -> _QuantSchemePair:
        """reference :670-712"""
        def eval_fn(pair: _QuantSchemePair):
            sim = self._create_quantsim_and_encodings(self.fp32_model, param_quant_scheme=pair.param_quant_scheme,
                                                      param_percentile=pair.param_percentile,
                                                      output_quant_scheme=pair.output_quant_scheme,
                                                      output_percentile=pair.output_percentile)
            score = self._evaluate_model_performance(sim.model)
            self.quant_scheme_scores[str(pair)] = score
            return score

        self.quant_scheme_scores = OrderedDict()
        candidates = self.get_quant_scheme_candidates()
        if self._quantsim_params["param_bw"] >= 16:     # enough precision: always tf
            candidates = [c for c in candidates if c.param_quant_scheme == QuantScheme.post_training_tf]
        if self._quantsim_params["output_bw"] >= 16:
            candidates = [c for c in candidates if c.output_quant_scheme == QuantScheme.post_training_tf]
        if len(candidates) == 1:
            return candidates[0]
        assert candidates
        return max(candidates, key=eval_fn)    # the first of equal scores, as max() does in the reference

    def _optimize_main(self, fp32_model: torch.nn.Module, target_acc: float) -> Dict[str, Any]:   # pylint: disable=too-many-branches
        """reference :714-816"""
        fp32_model = self.fp32_model
        with self.eval_manager.session("Prepare Model") as sess:
            raise _StageSkipped("Skipping Model Preparer")

        with self.eval_manager.session("QuantScheme Selection") as sess:
            self._quantsim_params["quant_scheme"] = sess.wrap(self._choose_default_quant_scheme)()

        with self.eval_manager.session("W32 Evaluation") as sess:
            w32_eval_score = sess.wrap(sess.eval)(model=fp32_model, param_bw=32)
            self.w32_eval_score = w32_eval_score
            if w32_eval_score < target_acc:
                # unlikely that post-training techniques reach the target: the reference returns all-None here
                return {"model": None, "accuracy": None, "encoding_path": None, "applied_techniques": None}
            sess.result["target_satisfied"] = True

        with self.eval_manager.session("Batchnorm Folding", ptq=True) as sess:
            model, _ = sess.wrap(self._apply_batchnorm_folding)(fp32_model)
            if sess.ptq_result is None:
                sess.set_ptq_result(model=model, applied_techniques=["batchnorm_folding"], export_kwargs=self._export_kwargs)

        best_result = self.eval_manager.get_best_ptq_result()
        if best_result and best_result.accuracy >= target_acc:
            sess.result["target_satisfied"] = True
            return best_result.as_dict()

        with self.eval_manager.session("Cross-Layer Equalization", ptq=True) as sess:
            model = sess.wrap(self._apply_cross_layer_equalization)(fp32_model)
            if sess.ptq_result is None:
                sess.set_ptq_result(model=model, applied_techniques=["cross_layer_equalization"],
                                    export_kwargs=self._export_kwargs)

        best_result = self.eval_manager.get_best_ptq_result()
        if best_result and best_result.accuracy >= target_acc:
            sess.result["target_satisfied"] = True
            return best_result.as_dict()

        if best_result is None:
            model = fp32_model
            applied_techniques = []
        else:
            if "cross_layer_equalization" not in best_result.applied_techniques:
                sess.result["effective"] = False
            model = best_result.load_model()
            applied_techniques = best_result.applied_techniques

        with self.eval_manager.session("AdaRound", ptq=True) as sess:
            model, encoding_path = self._apply_adaround(model)
            if sess.ptq_result is None:
                sess.set_ptq_result(model=model, encoding_path=encoding_path,
                                    applied_techniques=[*applied_techniques, "adaround"], export_kwargs=self._export_kwargs)

        best_result = self.eval_manager.get_best_ptq_result()
        if best_result:
            if "adaround" not in best_result.applied_techniques:
                sess.result["effective"] = False
            if best_result.accuracy >= target_acc:
                sess.result["target_satisfied"] = True
            return best_result.as_dict()

        raise RuntimeError("None of batchnorm folding, CLE, or Adaround has been finished successfully.")
