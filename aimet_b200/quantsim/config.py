"""Quantsim configuration: which quantizers exist / are enabled / are symmetric / are per-channel.

Mirrors the effect of the reference's QuantSimConfigurator
(TrainingExtensions/torch/src/python/aimet_torch/quantsim_config/quantsim_config.py:113-700 on top of
aimet_common/quantsim_config/quantsim_config.py) for the JSON schema of aimet_common/quantsim_config/default_config.json
and default_config_per_channel.json, in the reference's order of specificity:
  defaults -> params -> op_type -> supergroups (+ conv/linear -> batchnorm fusing) -> model_input -> model_output.
The reference walks its own ConnectedGraph (a torch.jit trace); here the op graph comes from torch.fx, which yields the
same producer/consumer relations for the feed-forward models this path targets.
"""
import json
import operator
from typing import Dict, List, Optional

import torch
import torch.fx
from torch import nn

# ---- the two stock configurations (same content as the reference's JSON files) -----------------------------------------
_OP_TYPE_COMMON = {
    "BatchNormalization": {"params": {"running_mean": {"is_quantized": "False"},
                                      "running_var": {"is_quantized": "False"}}},
    "Cast": {"is_output_quantized": "False"},
    "Dropout": {"is_output_quantized": "False"},
    "Expand": {"is_output_quantized": "False"},
    "Squeeze": {"is_output_quantized": "False"},
    "Pad": {"is_output_quantized": "False"},
    "Mean": {"is_output_quantized": "False"},
    "Gather": {"is_output_quantized": "False"},
}
_SUPERGROUPS = [{"op_list": ["Conv", "Relu"]}, {"op_list": ["ConvTranspose", "Relu"]}, {"op_list": ["Conv", "Clip"]},
                {"op_list": ["Add", "Relu"]}, {"op_list": ["Gemm", "Relu"]}]

DEFAULT_CONFIG = {
    "defaults": {"ops": {"is_output_quantized": "True"},
                 "params": {"is_quantized": "True", "is_symmetric": "True"},
                 "strict_symmetric": "False", "per_channel_quantization": "False"},
    "params": {"bias": {"is_quantized": "False"}},
    "op_type": dict(_OP_TYPE_COMMON),
    "supergroups": _SUPERGROUPS,
    "model_input": {"is_input_quantized": "True"},
    "model_output": {},
}
DEFAULT_CONFIG_PER_CHANNEL = {
    "defaults": {"ops": {"is_output_quantized": "True"},
                 "params": {"is_quantized": "True", "is_symmetric": "True"},
                 "strict_symmetric": "False", "per_channel_quantization": "True"},
    "params": {"bias": {"is_quantized": "False"}},
    "op_type": dict(_OP_TYPE_COMMON, Gemm={"per_channel_quantization": "False"},
                    MatMul={"per_channel_quantization": "False"}, LayerNorm={"per_channel_quantization": "False"}),
    "supergroups": _SUPERGROUPS,
    "model_input": {"is_input_quantized": "True"},
    "model_output": {},
}

# torch module type -> ONNX-style op types (reference aimet_torch/onnx_utils.py:85-145, the entries that matter here)
MODULE_OP_TYPES = {
    nn.Conv1d: ["Conv"], nn.Conv2d: ["Conv"], nn.Conv3d: ["Conv"],
    nn.ConvTranspose1d: ["ConvTranspose"], nn.ConvTranspose2d: ["ConvTranspose"], nn.ConvTranspose3d: ["ConvTranspose"],
    nn.Linear: ["Gemm", "MatMul"],
    nn.BatchNorm1d: ["BatchNormalization"], nn.BatchNorm2d: ["BatchNormalization"], nn.BatchNorm3d: ["BatchNormalization"],
    nn.ReLU: ["Relu"], nn.ReLU6: ["Clip"], nn.Hardtanh: ["Clip"], nn.LeakyReLU: ["LeakyRelu"], nn.PReLU: ["PRelu"],
    nn.Sigmoid: ["Sigmoid"], nn.Tanh: ["Tanh"], nn.Softmax: ["Softmax"], nn.GELU: ["GELU"], nn.Hardswish: ["HardSwish"],
    nn.SiLU: ["SiLU"],
    nn.MaxPool2d: ["MaxPool"], nn.AvgPool2d: ["AveragePool"], nn.AdaptiveAvgPool2d: ["GlobalAveragePool", "AveragePool"],
    nn.Dropout: ["Dropout"], nn.Dropout2d: ["Dropout"], nn.Flatten: ["Flatten"], nn.Embedding: ["Gather"],
    nn.LayerNorm: ["LayerNorm"], nn.GroupNorm: ["GroupNorm"], nn.Upsample: ["Upsample"],
}
_FUNCTIONAL_OP_TYPES = {   # the four element-wise functionals the reference maps (onnx_utils.py:156-161)
    operator.add: "Add", operator.iadd: "Add", torch.add: "Add", "add": "Add", "add_": "Add",
    operator.mul: "Mul", operator.imul: "Mul", torch.mul: "Mul", "mul": "Mul", "mul_": "Mul",
    operator.truediv: "Div", torch.div: "Div", "div": "Div",
    torch.cat: "Concat", torch.concat: "Concat",
}


def _truthy(v) -> bool:
    return v is True or (isinstance(v, str) and v == "True")


def load_config(config) -> Dict:
    if config is None:
        return DEFAULT_CONFIG
    if isinstance(config, dict):
        return config
    with open(config) as f:
        return json.load(f)


class Op:
    """One node of the op graph: a wrapped leaf module call site, an element-wise functional, or anything else."""

    def __init__(self, name, op_type, module=None, elementwise=False):
        self.name, self.type, self.module, self.elementwise = name, op_type, module, elementwise
        self.inputs: List[Optional["Op"]] = []      # producers (None == model input / constant)
        self.input_is_model_input: List[bool] = []
        self.consumers: List["Op"] = []

    def __repr__(self):
        return f"Op({self.name}:{self.type})"


def _is_torch_nn_module(m: nn.Module) -> bool:
    """reference aimet_torch/utils.py:919-927"""
    return type(m) in torch.nn.__dict__.values()


def _interior_op_count(m: nn.Module) -> int:
    """Number of operations in the forward of a childless module, or 1 when it cannot be traced on its own."""
    try:
        graph = torch.fx.Tracer().trace(m)
    except Exception:   # pylint: disable=broad-except
        return 1
    return sum(1 for n in graph.nodes if n.op in ("call_function", "call_method"))


def _is_pass_through(m: nn.Module) -> bool:
    """Modules whose eval-mode forward hands its input on untouched leave no operation in the reference's jit trace: the
    ConnectedGraph connects their producer straight to their consumer (and keeps an isolated op for the module)."""
    if type(m) is nn.Identity:   # pylint: disable=unidiomatic-typecheck
        return True
    return type(m).__name__ == "StochasticDepth"   # the reference traces the model in eval mode, whatever mode it is in


class _LeafTracer(torch.fx.Tracer):
    """Modules without children are leaves (the reference wraps exactly those: v1/quantsim.py:1440-1454) and one op of the
    graph -- except a childless module that is not a torch.nn class and runs more than one operation in its forward: the
    reference's ConnectedGraph parses into those (meta/connectedgraph.py:502-512, 1315-1329), so their interior shows up
    as module-less functional ops and the module itself (still wrapped) has no op of its own. torchvision's LayerNorm2d
    (permute, layer_norm, permute) is the common case."""

    def __init__(self):
        super().__init__()
        self._verdict = {}

    def is_leaf_module(self, m, module_qualified_name):
        if len(list(m.children())) != 0:
            return super().is_leaf_module(m, module_qualified_name)
        if _is_torch_nn_module(m) or _is_pass_through(m):
            return True
        if type(m) not in self._verdict:
            self._verdict[type(m)] = _interior_op_count(m) <= 1
        return self._verdict[type(m)]


def build_op_graph(model: nn.Module) -> List[Op]:
    graph = _LeafTracer().trace(model)
    modules = dict(model.named_modules())
    node_to_op: Dict[torch.fx.Node, Optional[Op]] = {}
    ops: List[Op] = []
    placeholders = set()

    def producer_ops(arg, out_ops, out_flags):
        if isinstance(arg, torch.fx.Node):
            if arg in placeholders:
                out_ops.append(None)
                out_flags.append(True)
            else:
                out_ops.append(node_to_op.get(arg))
                out_flags.append(False)
        elif isinstance(arg, (list, tuple)):
            for a in arg:
                producer_ops(a, out_ops, out_flags)

    for node in graph.nodes:
        if node.op == "placeholder":
            placeholders.add(node)
            continue
        if node.op in ("output", "get_attr"):
            node_to_op[node] = None
            continue
        if node.op == "call_module" and _is_pass_through(modules[node.target]) and node.args and \
                isinstance(node.args[0], torch.fx.Node):
            # transparent, as in the reference's ConnectedGraph (an Identity leaves no op in the jit trace): what follows a
            # folded-away batch norm still forms a supergroup with what precedes it
            src = node.args[0]
            if src in placeholders:
                placeholders.add(node)
            else:
                node_to_op[node] = node_to_op.get(src)
            continue
        if node.op == "call_module":
            mod = modules[node.target]
            types = MODULE_OP_TYPES.get(type(mod), [type(mod).__name__])
            op = Op(node.target, types[0], module=mod)
            op.all_types = types
        else:
            key = node.target
            t = _FUNCTIONAL_OP_TYPES.get(key)
            op = Op(node.name, t if t else f"fn:{getattr(key, '__name__', key)}", elementwise=t is not None)
            op.all_types = [op.type]
        for a in list(node.args) + list(node.kwargs.values()):
            producer_ops(a, op.inputs, op.input_is_model_input)
        for p in op.inputs:
            if p is not None:
                p.consumers.append(op)
        node_to_op[node] = op
        ops.append(op)
    return ops


def _match(op: Op, pattern: List[str], ignored) -> List[List[Op]]:
    """GraphSearcher._match_pattern (aimet_common/graph_searcher.py:88-124): a pattern matches along ANY consumer
    edge; ops in `ignored` are skipped through."""
    if op in ignored:
        out = None
        for c in op.consumers:
            m = _match(c, pattern, ignored)
            if m is not None:
                out = (out or []) + m
        return out
    if pattern[0] not in op.all_types:
        return None
    if len(pattern) == 1:
        return [[op]]
    out = None
    for c in op.consumers:
        m = _match(c, pattern[1:], ignored)
        if m:
            out = (out or []) + [[op] + lst for lst in m]
    return out


def configure(model: nn.Module, wrappers: Dict[nn.Module, "object"], config, ops: Optional[List[Op]]):
    """Apply `config` to the wrappers (module -> StaticGridQuantWrapper). `ops` is the op graph or None."""
    cfg = load_config(config)
    defaults = cfg["defaults"]
    # ---- 1. defaults (reference _set_default_configs :291-339) ----
    out_default = _truthy(defaults["ops"].get("is_output_quantized", "False"))
    in_default = _truthy(defaults["ops"].get("is_input_quantized", "False"))
    sym_acts = _truthy(defaults["ops"].get("is_symmetric", "False"))
    p_defaults = defaults.get("params", {})
    strict = _truthy(defaults.get("strict_symmetric", "False"))
    unsigned = _truthy(defaults.get("unsigned_symmetric", "False"))
    per_channel_default = _truthy(defaults.get("per_channel_quantization", "False"))

    def op_types_of(module):
        return MODULE_OP_TYPES.get(type(module), [type(module).__name__])

    # per-channel first: it replaces the param quantizer objects (reference quantsim.py:1469-1478 via config generator)
    for module, w in wrappers.items():
        pcq = per_channel_default
        for t in op_types_of(module):
            if t in cfg.get("op_type", {}) and "per_channel_quantization" in cfg["op_type"][t]:
                pcq = _truthy(cfg["op_type"][t]["per_channel_quantization"])
        if pcq and w.param_quantizers:
            w.enable_per_channel_quantization()

    for w in wrappers.values():
        for q in w.input_quantizers:
            q.enabled = in_default
            q.use_symmetric_encodings = sym_acts
        for q in w.output_quantizers:
            q.enabled = out_default
            q.use_symmetric_encodings = sym_acts
        for q in w.param_quantizers.values():
            if "is_quantized" in p_defaults:
                q.enabled = _truthy(p_defaults["is_quantized"])
            if "is_symmetric" in p_defaults:
                q.use_symmetric_encodings = _truthy(p_defaults["is_symmetric"])
        for q in w._all_quantizers():   # pylint: disable=protected-access
            q.use_strict_symmetric = strict
            q.use_unsigned_symmetric = unsigned
    # Element-wise functionals (add / mul / div / cat) have no wrapper of their own: "their output is quantized" is
    # realised by the input quantizers of whoever consumes it (reference _get_tensor_quantizers_for_output_true_setting
    # :212-237 applied by _set_default_configs_for_ops :316-328)
    if ops is not None and out_default:
        for op in ops:
            if op.elementwise:
                for c in op.consumers:
                    if c.module is not None and c.module in wrappers:
                        for q in wrappers[c.module].input_quantizers:
                            q.enabled = True
    # ---- 2. params (reference _set_param_configs :368-378) ----
    for pname, pcfg in cfg.get("params", {}).items():
        for w in wrappers.values():
            for name, q in w.param_quantizers.items():
                if name == pname:
                    _set_param(q, pcfg)
    # ---- 3. op_type (reference _set_op_type_configs :401-478) ----
    for module, w in wrappers.items():
        for t in op_types_of(module):
            ocfg = cfg.get("op_type", {}).get(t)
            if not ocfg:
                continue
            if "is_input_quantized" in ocfg:
                for q in w.input_quantizers:
                    q.enabled = _truthy(ocfg["is_input_quantized"])
            if "is_output_quantized" in ocfg:
                on = _truthy(ocfg["is_output_quantized"])
                for q in w.output_quantizers:
                    q.enabled = on
                if not on and ops is not None:
                    # turning a tensor's quantization off also turns off the consumers' quantizers on that same tensor
                    # (reference _get_tensor_quantizers_for_output_false_setting :259-277)
                    for op in ops:
                        if op.module is module:
                            for c in op.consumers:
                                if c.module is not None and c.module in wrappers:
                                    for q in wrappers[c.module].input_quantizers:
                                        q.enabled = False
            if "is_symmetric" in ocfg:
                for q in w.input_quantizers + w.output_quantizers:
                    q.use_symmetric_encodings = _truthy(ocfg["is_symmetric"])
            for pname, pcfg in ocfg.get("params", {}).items():
                if pname in w.param_quantizers:
                    _set_param(w.param_quantizers[pname], pcfg)
    if ops is None:
        return
    # ---- 4. supergroups + conv/linear->batchnorm fusing (reference _set_supergroup_configs :480-553) ----
    conv_bn_pairs = []
    for op in ops:
        for first in ("Conv", "ConvTranspose", "Gemm"):
            for m in _match(op, [first, "BatchNormalization"], []) or []:
                conv, bn = m
                if isinstance(conv.module, (nn.ConvTranspose2d,)) and conv.module.groups != 1:
                    continue
                if (conv, bn) not in conv_bn_pairs:
                    conv_bn_pairs.append((conv, bn))
    foldable_bns = [bn for _, bn in conv_bn_pairs]
    for sg in sorted(cfg.get("supergroups", []), key=lambda s: len(s["op_list"]), reverse=True):
        pattern = sg["op_list"]
        for op in ops:
            for matched in _match(op, pattern, foldable_bns) or []:
                for index, mop in enumerate(matched):
                    if mop.elementwise or mop.module is None or mop.module not in wrappers:
                        continue
                    w = wrappers[mop.module]
                    if index == 0:
                        for q in w.output_quantizers:
                            q.enabled = False
                    elif index == len(matched) - 1:
                        for q in w.input_quantizers:
                            q.enabled = False
                    else:
                        for q in w.input_quantizers + w.output_quantizers:
                            q.enabled = False
    for conv, bn in conv_bn_pairs:
        if conv.module not in wrappers or bn.module not in wrappers:
            continue
        cw, bw_ = wrappers[conv.module], wrappers[bn.module]
        for q in bw_.input_quantizers:
            q.enabled = False
        for q in bw_.param_quantizers.values():
            q.enabled = False
        for cq, bq in zip(cw.output_quantizers, bw_.output_quantizers):
            bq.enabled = cq.enabled
            cq.enabled = False
    # ---- 5. model input (reference _set_model_input_configs :555-574) ----
    if _truthy(cfg.get("model_input", {}).get("is_input_quantized", "False")):
        for op in ops:
            if op.module is not None and op.module in wrappers:
                w = wrappers[op.module]
                for idx, q in enumerate(w.input_quantizers):
                    if idx < len(op.input_is_model_input) and op.input_is_model_input[idx]:
                        q.enabled = True
    # ---- 6. model output (reference _set_model_output_configs :576-588) ----
    mo = cfg.get("model_output", {})
    if "is_output_quantized" in mo:
        for op in ops:
            if not op.consumers and op.module is not None and op.module in wrappers:
                for q in wrappers[op.module].output_quantizers:
                    q.enabled = _truthy(mo["is_output_quantized"])


def _set_param(q, pcfg):
    """reference _set_config_for_param :747-757"""
    if "is_quantized" in pcfg:
        q.enabled = _truthy(pcfg["is_quantized"])
    if "is_symmetric" in pcfg:
        q.use_symmetric_encodings = _truthy(pcfg["is_symmetric"])
