"""Host mirror of the reference's AdaRound (adaptive rounding of weights for post-training quantization), a caller that
drives the QuantSim hot path: it builds a QuantizationSimModel, derives the parameter
encodings through the native boundary (one call for all weights here), then learns, layer by layer, whether each weight
rounds up or down by minimising the layer's reconstruction error.

Reference: TrainingExtensions/torch/src/python/aimet_torch/v1/adaround/
  adaround_weight.py     AdaroundParameters :73-110, Adaround :113-642 (apply_adaround :118-167, _apply_adaround :170-205,
                         _adaround_model :208-330 [the path without a checkpoints config], _run_adaround_model :333-383,
                         _compute_param_encodings :386-412, _export_encodings_to_json :494-514)
  adaround_wrapper.py    AdaroundWrapper :93-224 (delta / offset through AimetTensorQuantizer.makeDeltaOffsetTensor, :186)
  adaround_optimizer.py  AdaroundOptimizer :63-368
  adaround_loss.py       AdaroundHyperParameters :46-62, AdaroundLoss :65-135
  activation_sampler.py  ActivationSampler :176-256 (+ aimet_torch/utils.py ModuleData :106-186)
Same class and method names, same arithmetic in the same order (the optimisation is plain torch: Adam on `alpha`), so that a
seeded run reproduces the reference's adarounded weights (tests/test_adaround.py, against a golden produced by the
reference's unmodified code). Not mirrored: the checkpoints-config variants that split very large models into cached blocks.
"""
import contextlib
import json
import os
from typing import Any, Callable, Dict, List, Optional, Tuple, Union

import numpy as np
import torch
import torch.distributed as dist
from torch.nn import functional

from .defs import MAP_QUANT_SCHEME_TO_PYMO, QuantizationDataType, QuantScheme
from .qc_quantize_op import QcQuantizeOpMode, StaticGridQuantWrapper
from .quantsim import QuantizationSimModel
from .tensor_quantizer import StaticGridPerChannelQuantizer

AdaroundSupportedModules = (torch.nn.Conv2d, torch.nn.ConvTranspose2d, torch.nn.Linear)
BATCH_SIZE = 32
# aimet_torch/meta/connectedgraph_utils.py:52-58
ActivationTypes = (torch.nn.ReLU6, torch.nn.ReLU, torch.nn.PReLU, torch.nn.RReLU, torch.nn.LeakyReLU, torch.nn.Sigmoid,
                   torch.nn.LogSigmoid, torch.nn.Softmin, torch.nn.Softmax, torch.nn.LogSoftmax, torch.nn.Tanh,
                   torch.nn.Hardtanh, torch.nn.ELU, torch.nn.Hardshrink, torch.nn.Hardsigmoid, torch.nn.Hardswish,
                   torch.nn.MultiheadAttention, torch.nn.SELU, torch.nn.CELU, torch.nn.GELU, torch.nn.SiLU, torch.nn.Mish,
                   torch.nn.Softplus, torch.nn.Softshrink, torch.nn.Softsign, torch.nn.Tanhshrink, torch.nn.Threshold,
                   torch.nn.GLU, torch.nn.Softmax2d, torch.nn.AdaptiveLogSoftmaxWithLoss)


class AdaroundConstants:
    """aimet_common/defs.py:302-306"""
    GAMMA = -0.1
    ZETA = 1.1


class AdaroundHyperParameters:
    def __init__(self, num_iterations: int, reg_param: float, beta_range: Tuple, warm_start: float):
        self.num_iterations = num_iterations
        self.reg_param = reg_param
        self.beta_range = beta_range
        self.warm_start = warm_start


class AdaroundParameters:
    """Configuration parameters for Adaround (reference adaround_weight.py:73-110)."""

    def __init__(self, data_loader, num_batches: int, default_num_iterations: int = None, default_reg_param: float = 0.01,
                 default_beta_range: Tuple = (20, 2), default_warm_start: float = 0.2,
                 forward_fn: Callable[[torch.nn.Module, Any], Any] = None):
        if len(data_loader) < num_batches:
            raise ValueError(f'Can not fetch {num_batches} batches from '
                             f'a data loader of length {len(data_loader)}.')
        self.data_loader = data_loader
        self.num_batches = num_batches
        self.num_iterations = default_num_iterations
        self.reg_param = default_reg_param
        self.beta_range = default_beta_range
        self.warm_start = default_warm_start
        self.forward_fn = forward_fn


class AdaroundLoss:
    """reference adaround_loss.py:65-135"""

    @staticmethod
    def compute_recon_loss(ada_quantized_output: torch.Tensor, orig_output: torch.Tensor) -> torch.Tensor:
        return (torch.norm(ada_quantized_output - orig_output, p="fro", dim=1) ** 2).mean()

    @classmethod
    def compute_round_loss(cls, alpha: torch.Tensor, opt_params: AdaroundHyperParameters, cur_iter: int):
        if cur_iter < opt_params.num_iterations * opt_params.warm_start:
            return 0
        h_alpha = torch.clamp(torch.sigmoid(alpha) * (AdaroundConstants.ZETA - AdaroundConstants.GAMMA) +
                              AdaroundConstants.GAMMA, 0, 1)
        beta = cls._compute_beta(opt_params.num_iterations, cur_iter, opt_params.beta_range, opt_params.warm_start)
        reg_term = torch.add(1, -(torch.add(2 * h_alpha, -1).abs()).pow(beta)).sum()
        return opt_params.reg_param * reg_term

    @staticmethod
    def _compute_beta(max_iter: int, cur_iter: int, beta_range: Tuple, warm_start: float) -> float:
        assert cur_iter < max_iter, 'Current iteration should be less than total maximum number of iterations.'
        start_beta, end_beta = beta_range
        warm_start_end_iter = warm_start * max_iter
        rel_iter = (cur_iter - warm_start_end_iter) / (max_iter - warm_start_end_iter)
        return end_beta + 0.5 * (start_beta - end_beta) * (1 + np.cos(rel_iter * np.pi))


def broadcast_to_tensor(tensor, encoding, ch_axis):
    """reference quantsim_straight_through_grad.py:70-88: an encoding value (scalar or one per channel) shaped to broadcast
    against `tensor` along `ch_axis`."""
    if not isinstance(encoding, torch.Tensor):
        encoding = torch.Tensor([encoding]).to(tensor.device)
    if encoding.numel() > 1:
        shape = [1] * tensor.dim()
        shape[ch_axis] = -1
        return encoding.to(tensor.device).view(*shape)
    return encoding.to(tensor.device)


class AdaroundWrapper(torch.nn.Module):
    """Wraps one StaticGridQuantWrapper while its weight rounding is being learned (reference adaround_wrapper.py:93-224)."""
    weight_name = "weight"

    def __init__(self, module: StaticGridQuantWrapper):
        super().__init__()
        assert self.weight_name in module.param_quantizers
        self.module_to_wrap = module
        self._init_param()

    @property
    def weight(self) -> torch.Tensor:
        return getattr(self.get_original_module(), self.weight_name)

    def get_original_module(self) -> torch.nn.Module:
        return self.module_to_wrap._module_to_wrap   # pylint: disable=protected-access

    def forward(self, *args, **kwargs):
        original = self.get_original_module()
        weight = self.weight
        if self._quantizer().enabled:
            weight = self.apply_adaround(weight)
        with self._disable_weight_quantizer(), _patch_attr(original, self.weight_name, weight):
            return self.module_to_wrap.forward(*args, **kwargs)

    def apply_adaround(self, tensor: torch.Tensor) -> torch.Tensor:
        input_dtype = tensor.dtype
        alpha = self.alpha.to(device=tensor.device, dtype=tensor.dtype)
        tensor = torch.floor(tensor / self.broadcasted_delta)
        if self.use_soft_rounding:
            h_alpha = torch.clamp(torch.sigmoid(alpha) * (AdaroundConstants.ZETA - AdaroundConstants.GAMMA) +
                                  AdaroundConstants.GAMMA, 0, 1)
        else:
            h_alpha = (alpha >= 0).to(tensor.dtype)
        tensor = tensor + h_alpha
        tensor_quant = torch.clamp(tensor - self.broadcasted_offset, self.clip_min, self.clip_max)
        tensor_dequant = (tensor_quant + self.broadcasted_offset) * self.broadcasted_delta
        return tensor_dequant.to(input_dtype)

    def _quantizer(self):
        return self.module_to_wrap.param_quantizers[self.weight_name]

    @contextlib.contextmanager
    def _disable_weight_quantizer(self):
        quantizer = self._quantizer()
        is_enabled = quantizer.enabled
        quantizer.enabled = False
        try:
            yield
        finally:
            quantizer.enabled = is_enabled

    def _get_weight_quantizer_delta_and_offset(self):
        quantizer = self._quantizer()
        encoding = quantizer.encoding
        if isinstance(encoding, list):
            # the native boundary call of this wrapper (reference :186)
            cpp_op = quantizer._op_factory(MAP_QUANT_SCHEME_TO_PYMO[quantizer.quant_scheme])   # pylint: disable=protected-access
            if hasattr(cpp_op, "makeDeltaOffsetTensor"):
                delta, offset = cpp_op.makeDeltaOffsetTensor(self.weight.device, encoding)
            else:       # (a test backend without it)
                delta = torch.tensor([e.delta for e in encoding], dtype=torch.float32, device=self.weight.device)
                offset = torch.tensor([e.offset for e in encoding], dtype=torch.float32, device=self.weight.device)
        else:
            delta, offset = encoding.delta, encoding.offset
        ch_axis = quantizer._ch_axis if isinstance(quantizer, StaticGridPerChannelQuantizer) else 0   # pylint: disable=protected-access
        return broadcast_to_tensor(self.weight, delta, ch_axis), broadcast_to_tensor(self.weight, offset, ch_axis)

    def _init_param(self):
        self.broadcasted_delta, self.broadcasted_offset = self._get_weight_quantizer_delta_and_offset()
        self.alpha = self._generate_alpha_parameter(self.weight, self.broadcasted_delta)
        self.bitwidth = self._quantizer().bitwidth
        self.use_soft_rounding = True
        self.clip_max = 2 ** self.bitwidth - 1
        self.clip_min = 0

    @staticmethod
    def _generate_alpha_parameter(tensor: torch.Tensor, delta: torch.Tensor) -> torch.nn.Parameter:
        tensor_floor = torch.floor(tensor / delta)
        tensor = (tensor / delta) - tensor_floor
        alpha = - torch.log((AdaroundConstants.ZETA - AdaroundConstants.GAMMA) / (tensor - AdaroundConstants.GAMMA) - 1)
        return torch.nn.Parameter(alpha.float(), requires_grad=True)


@contextlib.contextmanager
def _patch_attr(obj, name, value):
    """reference aimet_torch/utils.py patch_attr: temporarily shadow an attribute (a Parameter by a plain tensor)."""
    had = name in obj.__dict__
    old = obj.__dict__.get(name)
    param = obj._parameters.pop(name, None) if hasattr(obj, "_parameters") else None   # pylint: disable=protected-access
    obj.__dict__[name] = value
    try:
        yield
    finally:
        if had:
            obj.__dict__[name] = old
        else:
            obj.__dict__.pop(name, None)
        if param is not None:
            obj._parameters[name] = param   # pylint: disable=protected-access


class _StopForward(Exception):
    pass


class ActivationSampler:
    """Input of the quantized module (all preceding weights quantized) and output of the original module for one batch of
    model inputs (reference activation_sampler.py:176-256, utils.ModuleData :106-186)."""

    def __init__(self, orig_module, quant_module, orig_model, quant_model, forward_fn):
        self._orig = (orig_model, orig_module)
        self._quant = (quant_model, quant_module)
        self._forward_fn = forward_fn or (lambda model, inputs: model(*inputs) if isinstance(inputs, (list, tuple))
                                          else model(inputs))

    def _collect(self, model, module, model_input, collect_input, collect_output):
        got = {}

        def hook(_, inp, out):
            if collect_input:
                got["inp"] = inp[0]
            if collect_output:
                got["out"] = out
            raise _StopForward

        handle = module.register_forward_hook(hook)
        device = next(model.parameters()).device
        model_input = _to_device(model_input, device)
        was_training = {m: m.training for m in model.modules()}
        model.eval()
        try:
            with torch.no_grad():
                self._forward_fn(model, model_input)
        except _StopForward:
            pass
        finally:
            handle.remove()
            for m, t in was_training.items():
                m.training = t
        inp, out = got.get("inp"), got.get("out")
        return (inp.detach() if isinstance(inp, torch.Tensor) else None,
                out.detach() if isinstance(out, torch.Tensor) else None)

    def sample_acts(self, model_inputs, collect_input=True, collect_output=True):
        inp_data = out_data = None
        if collect_input:
            inp_data, _ = self._collect(*self._quant, model_inputs, True, False)
        if collect_output:
            _, out_data = self._collect(*self._orig, model_inputs, False, True)
        return inp_data, out_data

    def sample_and_place_all_acts_on_cpu(self, cached_dataset):
        all_inp, all_out = [], []
        for model_inputs in cached_dataset:
            inp, out = self.sample_acts(model_inputs)
            all_inp.append(inp.cpu())
            all_out.append(out.cpu())
        return torch.cat(all_inp, dim=0), torch.cat(all_out, dim=0)


def _to_device(x, device):
    if isinstance(x, torch.Tensor):
        return x.to(device)
    if isinstance(x, (list, tuple)):
        return type(x)(_to_device(t, device) for t in x)
    return x


class AdaroundOptimizer:
    """Optimizes the weight rounding of one quantized wrapper module (reference adaround_optimizer.py:63-368)."""

    @classmethod
    def adaround_module(cls, module, quant_module: AdaroundWrapper, orig_model, quant_model, act_func, cached_dataset,
                        forward_fn, opt_params: AdaroundHyperParameters):
        assert isinstance(quant_module, AdaroundWrapper), f'{quant_module} is not adaround wrapper module.'
        act_sampler = ActivationSampler(module, quant_module, orig_model, quant_model, forward_fn)
        inp_data, out_data = act_sampler.sample_acts(cached_dataset[0])
        before = cls._compute_recons_metrics(quant_module, act_func, inp_data, out_data)
        cls._optimize_rounding(module, quant_module, orig_model, quant_model, act_func, cached_dataset, forward_fn,
                               opt_params)
        after = cls._compute_recons_metrics(quant_module, act_func, inp_data, out_data)
        quant_module.use_soft_rounding = False       # hard rounding from here on
        return before, after

    @classmethod
    def _optimize_rounding(cls, module, quant_module: AdaroundWrapper, orig_model, quant_model, act_func, cached_dataset,
                           forward_fn, opt_params: AdaroundHyperParameters):
        rank, world_size = (dist.get_rank(), dist.get_world_size()) if dist.is_initialized() else (0, 1)
        cached_dataset = [cached_dataset[i] for i in range(rank, len(cached_dataset), world_size)]   # shard the batches
        assert quant_module.use_soft_rounding, 'optimization should use soft rounding only.'
        assert quant_module.alpha is not None, 'alpha parameter should be initialized.'
        optimizer = torch.optim.Adam([quant_module.alpha])
        for group in optimizer.param_groups:
            group['lr'] *= world_size
        act_sampler = ActivationSampler(module, quant_module, orig_model, quant_model, forward_fn)
        device = next(module.parameters()).device
        # all intermediate activations are sampled once and kept next to the layer (the reference stages them through CPU
        # memory and moves the two models off the GPU meanwhile; on a 180 GB part they simply stay)
        all_inp_data, all_orig_out_data = act_sampler.sample_and_place_all_acts_on_cpu(cached_dataset)
        all_inp_data, all_orig_out_data = all_inp_data.to(device), all_orig_out_data.to(device)
        for iteration in range(opt_params.num_iterations // world_size):
            indices = torch.randperm(all_inp_data.size(0))[:BATCH_SIZE]
            inp_data = all_inp_data[indices.to(all_inp_data.device)].to(device)
            orig_out_data = all_orig_out_data[indices.to(all_inp_data.device)].to(device)
            optimizer.zero_grad()
            quant_out_data = cls._compute_output_with_adarounded_weights(quant_module, inp_data)
            if act_func is not None:
                orig_out_data = act_func(orig_out_data)
                quant_out_data = act_func(quant_out_data)
            recon_loss = AdaroundLoss.compute_recon_loss(quant_out_data, orig_out_data)
            round_loss = AdaroundLoss.compute_round_loss(quant_module.alpha, opt_params, iteration)
            total_loss = recon_loss + round_loss
            total_loss.backward()
            if dist.is_initialized():
                dist.all_reduce(quant_module.alpha.grad)
            quant_module.alpha.grad /= world_size
            optimizer.step()

    @classmethod
    def _compute_recons_metrics(cls, quant_module: AdaroundWrapper, act_func, inp_data, out_data):
        quant_module.use_soft_rounding = False
        out_hard = cls._compute_output_with_adarounded_weights(quant_module, inp_data)
        quant_module.use_soft_rounding = True
        out_soft = cls._compute_output_with_adarounded_weights(quant_module, inp_data)
        if act_func is not None:
            out_data, out_soft, out_hard = act_func(out_data), act_func(out_soft), act_func(out_hard)
        return float(functional.mse_loss(out_hard, out_data).detach()), float(functional.mse_loss(out_soft, out_data).detach())

    @staticmethod
    def _compute_output_with_adarounded_weights(quant_module: AdaroundWrapper, inp_data: torch.Tensor):
        module = quant_module.get_original_module()
        quant_module.to(inp_data.device)
        w = quant_module.apply_adaround(quant_module.weight)
        if isinstance(module, torch.nn.Conv2d):
            return functional.conv2d(inp_data, w, bias=module.bias, stride=module.stride, dilation=module.dilation,
                                     padding=module.padding, groups=module.groups)
        if isinstance(module, torch.nn.ConvTranspose2d):
            return functional.conv_transpose2d(inp_data, w, bias=module.bias, stride=module.stride, padding=module.padding,
                                               output_padding=module.output_padding, groups=module.groups,
                                               dilation=module.dilation)
        if isinstance(module, torch.nn.Linear):
            return functional.linear(inp_data, w, bias=module.bias)
        raise ValueError('AdaRound is not supported for the module: ', module)


def get_ordered_list_of_modules(model: torch.nn.Module, dummy_input) -> List[Tuple[str, torch.nn.Module]]:
    """(name, leaf module) in order of execution (reference aimet_torch/utils.py get_ordered_list_of_modules)."""
    names = {m: n for n, m in model.named_modules()}
    order, hooks = [], []

    def hook(mod, _inp, _out):
        order.append((names[mod], mod))

    for m in model.modules():
        if not list(m.children()):
            hooks.append(m.register_forward_hook(hook))
    was_training = model.training
    model.eval()
    try:
        with torch.no_grad():
            model(*dummy_input) if isinstance(dummy_input, (list, tuple)) else model(dummy_input)
    finally:
        for h in hooks:
            h.remove()
        model.train(was_training)
    return order


def get_module_act_func_pair(model: torch.nn.Module) -> Dict[torch.nn.Module, Optional[torch.nn.Module]]:
    """module -> the activation module that immediately consumes its output, else None (reference
    connectedgraph_utils.py:61-106; the op graph here comes from torch.fx, see quantsim/config.py)."""
    from . import config as qconfig
    ops_ = qconfig.build_op_graph(model)
    first_consumer = {}
    for op in ops_:
        for producer in op.inputs:
            if producer is not None and id(producer) not in first_consumer:
                first_consumer[id(producer)] = op
    pairs = {}
    for op in ops_:
        if op.module is None:
            continue
        pairs[op.module] = None
        nxt = first_consumer.get(id(op))
        if nxt is not None and isinstance(nxt.module, ActivationTypes):
            pairs[op.module] = nxt.module
    return pairs


class Adaround:
    """Weight-rounding mechanism for post-training quantization (reference adaround_weight.py:113-642)."""

    @classmethod
    def apply_adaround(cls, model: torch.nn.Module, dummy_input, params: AdaroundParameters, path: str,
                       filename_prefix: str, default_param_bw: int = 4,
                       param_bw_override_list: List[Tuple[torch.nn.Module, int]] = None,
                       ignore_quant_ops_list: List[torch.nn.Module] = None,
                       default_quant_scheme: QuantScheme = QuantScheme.post_training_tf_enhanced,
                       default_config_file=None) -> torch.nn.Module:
        """Returns a copy of `model` with the weight of every Conv / Linear module rounded adaptively onto its quantization
        grid, and writes `<path>/<filename_prefix>.encodings` with the matching parameter encodings (load them into a
        QuantizationSimModel with set_and_freeze_param_encodings)."""
        quant_sim = QuantizationSimModel(model, dummy_input=dummy_input, quant_scheme=default_quant_scheme,
                                         default_param_bw=default_param_bw, config_file=default_config_file)
        if param_bw_override_list:
            cls._override_param_bitwidth(model, quant_sim, param_bw_override_list)
        if ignore_quant_ops_list:
            cls._exclude_modules(model, quant_sim, ignore_quant_ops_list)
        cls._compute_param_encodings(quant_sim)
        return cls._apply_adaround(quant_sim, model, dummy_input, params, path, filename_prefix)

    @classmethod
    def _apply_adaround(cls, quant_sim, model, dummy_input, params, path, filename_prefix):
        for _, w in quant_sim.quant_wrappers():
            for q in list(w.input_quantizers) + list(w.output_quantizers):
                assert not q.enabled                      # all activation quantizers must be off
        module_act_func_pair = get_module_act_func_pair(model)
        cls._adaround_model(model, quant_sim, module_act_func_pair, params, dummy_input)
        cls._export_encodings_to_json(path, filename_prefix, quant_sim)
        return QuantizationSimModel.get_original_model(quant_sim.model)

    @classmethod
    def _adaround_model(cls, model, quant_sim, module_act_func_pair, params: AdaroundParameters, dummy_input):
        num_iterations = params.num_iterations
        if num_iterations is None:
            lowest = min(q.bitwidth for _, w in quant_sim.quant_wrappers() for q in w.param_quantizers.values()
                         if q.enabled and q.data_type == QuantizationDataType.int)
            num_iterations = 15000 if lowest < 8 else 10000
        cached_dataset = []
        for batch in params.data_loader:                   # reference utils.CachedDataset: the first num_batches batches
            if len(cached_dataset) >= params.num_batches:
                break
            cached_dataset.append(batch)
        opt_params = AdaroundHyperParameters(num_iterations, params.reg_param, params.beta_range, params.warm_start)
        modules = get_ordered_list_of_modules(model, dummy_input)
        cls._run_adaround_model(modules, model, quant_sim.model, module_act_func_pair, opt_params, params.forward_fn,
                                cached_dataset)

    @classmethod
    def _run_adaround_model(cls, modules, model, quant_sim_model, module_act_func_pair, opt_params, forward_fn,
                            cached_dataset):
        for name, module in modules:
            if not isinstance(module, AdaroundSupportedModules):
                continue
            quant_wrapper = cls._get_quant_wrapper(quant_sim_model, name)
            if not quant_wrapper:
                continue
            with cls._replace_quantization_layer(quant_sim_model, name) as adaround_wrapper:
                act_func = module_act_func_pair.get(module)
                AdaroundOptimizer.adaround_module(module, adaround_wrapper, model, quant_sim_model, act_func, cached_dataset,
                                                  forward_fn, opt_params)
                weight = adaround_wrapper.weight
                with torch.no_grad():                       # fold the trained alpha into the weight
                    adaround_wrapper.use_soft_rounding = True
                    weight.copy_(adaround_wrapper.apply_adaround(weight))

    @staticmethod
    def _compute_param_encodings(quant_sim: QuantizationSimModel):
        """Parameter encodings only, activation quantizers off, wrappers ACTIVE (reference :386-412). All planned parameter
        quantizers are derived by one native call (quantsim.param_plan); the rest one by one."""
        for _, wrapper in quant_sim.quant_wrappers():
            for q in list(wrapper.input_quantizers) + list(wrapper.output_quantizers):
                q.enabled = False
        plan = quant_sim._plan()   # pylint: disable=protected-access
        if plan is not None:
            plan.ensure()
            with torch.no_grad():
                plan.refresh()
        for _, wrapper in quant_sim.quant_wrappers():
            if isinstance(wrapper, StaticGridQuantWrapper):
                for name, param in wrapper.get_named_parameters():
                    q = wrapper.param_quantizers[name]
                    if not (q.enabled and q.bitwidth != 32) or q._has_encoding():   # pylint: disable=protected-access
                        continue
                    q.reset_encoding_stats()
                    q.update_encoding_stats(param.data)
                    q.compute_encoding()
                wrapper.set_mode(QcQuantizeOpMode.ACTIVE)

    @staticmethod
    def _get_quant_wrapper(quant_sim_model, module_name: str) -> Union[StaticGridQuantWrapper, None]:
        for name, module in quant_sim_model.named_modules():
            if name == module_name and isinstance(module, StaticGridQuantWrapper):
                return module
        return None

    @classmethod
    @contextlib.contextmanager
    def _replace_quantization_layer(cls, quant_sim_model, module_name: str):
        quant_module = dict(quant_sim_model.named_modules())[module_name]
        assert quant_module.param_quantizers['weight'], f'{quant_module} does not have weight parameter.'
        assert quant_module.param_quantizers['weight'].encoding, f'{quant_module} encoding needs to be set.'
        adaround_layer = AdaroundWrapper(quant_module)
        upper_name, _, target = module_name.rpartition('.')
        upper = dict(quant_sim_model.named_modules())[upper_name] if upper_name else quant_sim_model
        original = getattr(upper, target)
        setattr(upper, target, adaround_layer)
        try:
            yield adaround_layer
        finally:
            setattr(upper, target, original)

    @classmethod
    def _export_encodings_to_json(cls, path: str, filename_prefix: str, quant_sim: QuantizationSimModel):
        param_encodings = {}
        for name, wrapper in quant_sim.quant_wrappers():
            if isinstance(wrapper.get_original_module(), AdaroundSupportedModules) and 'weight' in wrapper.param_quantizers:
                encodings = wrapper.export_param_encodings().get('weight')
                if encodings:
                    param_encodings[name + '.weight'] = encodings
        os.makedirs(os.path.abspath(path), exist_ok=True)
        with open(os.path.join(path, filename_prefix + '.encodings'), 'w') as f:
            json.dump({'param_encodings': param_encodings}, f, sort_keys=True, indent=4)

    @staticmethod
    def _override_param_bitwidth(model, quant_sim, param_bw_override_list):
        module_to_name = {m: n for n, m in model.named_modules() if isinstance(m, AdaroundSupportedModules)}
        wrappers = dict(quant_sim.quant_wrappers())
        for module, bw in param_bw_override_list:
            wrappers[module_to_name[module]].param_quantizers['weight'].bitwidth = bw

    @classmethod
    def _exclude_modules(cls, model, quant_sim, ignore_quant_ops_list):
        names = {m: n for n, m in model.named_modules()}
        sim_modules = dict(quant_sim.model.named_modules())
        doomed = []
        for module in ignore_quant_ops_list:
            for m in module.modules():
                w = sim_modules.get(names.get(m))
                if isinstance(w, StaticGridQuantWrapper):
                    doomed.append(w)
        quant_sim.exclude_layers_from_quantization(doomed)
