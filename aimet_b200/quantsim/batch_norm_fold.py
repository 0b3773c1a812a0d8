"""Batch-norm folding, as far as AutoQuant needs it (reference aimet_torch/batch_norm_fold.py:81-657,
aimet_common/batch_norm_fold.py:71-100).

Not part of the quantization-simulation hot path (SURVEY section 2 lists BN folding as its own subsystem); it is here
because AutoQuant's first stage (v1/auto_quant.py:543-556) folds the batch norms before it builds its sims. Same
arithmetic as the reference's Python implementation, in the same order and in numpy float32:

    BN after the layer  ("fold backward"):  W' = W * (gamma / sigma)[out],     b' = beta - (mu - b) * (gamma / sigma)
    BN before the layer ("fold forward"):   W' = W * (gamma / sigma)[in],      b' = W2d @ beta - W2d @ (mu * gamma / sigma) + b
    sigma = sqrt(running_var + eps)

Pairs are found on the torch.fx graph of the model (the reference walks its ConnectedGraph): a Conv / Linear whose only
consumer is a BatchNorm (and that BatchNorm's only producer), or a BatchNorm whose only consumer is an un-padded,
un-grouped Conv / Linear. A folded BatchNorm module is replaced by torch.nn.Identity.
"""
from typing import Iterable, List, Tuple

import numpy as np
import torch
import torch.fx
from torch import nn

_LAYERS = (nn.Conv1d, nn.Conv2d, nn.Conv3d, nn.ConvTranspose2d, nn.Linear)
_BNS = (nn.BatchNorm1d, nn.BatchNorm2d, nn.BatchNorm3d)


def _expand_shape_to_4d(shape) -> List[int]:
    shape = list(shape)
    if len(shape) > 5:
        raise RuntimeError
    if len(shape) < 4:
        return shape + [1] * (4 - len(shape))
    if len(shape) == 5:
        return shape[:3] + [int(np.prod(shape[3:]))]
    return shape


def _fold_arrays(weight, bias, gamma, beta, mu, sigma, fold_backward: bool):
    """aimet_common/batch_norm_fold.py:71-100"""
    assert weight.ndim == 4
    assert not np.any(sigma == 0)
    scale = gamma / sigma
    if fold_backward:
        return weight * scale[:, None, None, None], beta - (mu - bias) * scale
    w2d = weight.sum(3).sum(2)
    mu_hat = np.matmul(w2d, mu * scale)
    beta_hat = np.matmul(w2d, beta)
    return weight * scale[None, :, None, None], beta_hat - mu_hat + bias


def fold_to_weight(layer: nn.Module, bn: nn.Module, fold_backward: bool):
    """BatchNormFold._fold_to_weight (batch_norm_fold.py:253-283) with the Python arithmetic (:136-160)."""
    transposed = isinstance(layer, nn.ConvTranspose2d) and layer.groups == 1
    with torch.no_grad():
        if transposed:
            layer.weight.data = layer.weight.data.permute(1, 0, 2, 3)
        if layer.bias is None:
            out = layer.out_features if isinstance(layer, nn.Linear) else layer.out_channels
            layer.bias = nn.Parameter(torch.zeros(out, device=layer.weight.device, dtype=layer.weight.dtype))
        gamma = bn.weight.detach().cpu().numpy()
        beta = bn.bias.detach().cpu().numpy()
        mu = bn.running_mean.detach().cpu().numpy()
        sigma = torch.sqrt(bn.running_var + bn.eps).detach().cpu().numpy()
        w = layer.weight.detach().cpu().numpy()
        b = layer.bias.detach().cpu().numpy()
        w4, b2 = _fold_arrays(w.reshape(_expand_shape_to_4d(w.shape)), b, gamma, beta, mu, sigma, fold_backward)
        layer.bias.copy_(torch.from_numpy(np.ascontiguousarray(b2)).reshape_as(layer.bias))
        layer.weight.copy_(torch.from_numpy(np.ascontiguousarray(w4)).reshape_as(layer.weight))
        if transposed:
            layer.weight.data = layer.weight.data.permute(1, 0, 2, 3)


def _can_fold_forward(layer: nn.Module) -> bool:
    """A BatchNorm in FRONT of a layer folds only if the layer sees every input element with the same weight: no padding
    (the border would see the un-normalised zero), no groups (batch_norm_fold.py:440-470)."""
    if isinstance(layer, nn.Linear):
        return True
    if isinstance(layer, nn.ConvTranspose2d):
        return False
    padding = layer.padding if isinstance(layer.padding, tuple) else (layer.padding,)
    return layer.groups == 1 and all(p == 0 for p in padding) and not isinstance(layer.padding, str)


def find_all_batch_norms_to_fold(model: nn.Module) -> Tuple[List[Tuple[nn.Module, nn.Module]], List[Tuple[nn.Module, nn.Module]]]:
    """(conv/linear -> bn pairs, bn -> conv/linear pairs), each module in at most one pair, in graph order."""
    traced = torch.fx.symbolic_trace(model)
    modules = dict(model.named_modules())
    uses = {}
    for node in traced.graph.nodes:
        if node.op == "call_module":
            uses[node.target] = uses.get(node.target, 0) + 1

    def module_of(node):
        if node.op == "call_module" and uses.get(node.target) == 1:   # a reused module cannot be folded
            return modules.get(node.target)
        return None

    layer_bn, bn_layer, taken = [], [], set()
    for node in traced.graph.nodes:
        m = module_of(node)
        if isinstance(m, _LAYERS) and m not in taken and len(node.users) == 1:
            nxt = module_of(next(iter(node.users)))
            if isinstance(nxt, _BNS) and nxt not in taken and nxt.track_running_stats and nxt.affine:
                layer_bn.append((m, nxt))
                taken.update((m, nxt))
    for node in traced.graph.nodes:
        m = module_of(node)
        if isinstance(m, _BNS) and m not in taken and len(node.users) == 1 and m.track_running_stats and m.affine:
            nxt = module_of(next(iter(node.users)))
            if isinstance(nxt, _LAYERS) and nxt not in taken and _can_fold_forward(nxt):
                bn_layer.append((m, nxt))
                taken.update((m, nxt))
    return layer_bn, bn_layer


def _replace(model: nn.Module, targets: Iterable[nn.Module]):
    targets = set(targets)
    for parent in model.modules():
        for name, child in list(parent.named_children()):
            if child in targets:
                setattr(parent, name, nn.Identity())


def fold_all_batch_norms(model: nn.Module, input_shapes=None, dummy_input=None) -> List[Tuple[nn.Module, nn.Module]]:
    """Fold every foldable BatchNorm of `model` IN PLACE into its neighbour and replace it by Identity; returns the list of
    (layer, batch norm) pairs (reference fold_all_batch_norms_to_weight, batch_norm_fold.py:330-371)."""
    del input_shapes, dummy_input   # the fx graph needs neither
    was_training = model.training
    model.eval()
    try:
        layer_bn, bn_layer = find_all_batch_norms_to_fold(model)
    finally:
        model.train(was_training)
    for layer, bn in layer_bn:
        fold_to_weight(layer, bn, fold_backward=True)
    for bn, layer in bn_layer:
        fold_to_weight(layer, bn, fold_backward=False)
    _replace(model, [bn for _, bn in layer_bn] + [bn for bn, _ in bn_layer])
    return layer_bn + [(layer, bn) for bn, layer in bn_layer]
