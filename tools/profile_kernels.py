"""Launches each hot-path kernel a few times on a 256 MB tensor: the command ncu wraps (never a source of bench numbers)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aimet_b200 import ops  # noqa: E402
from aimet_b200.state import StateArena  # noqa: E402

dev = torch.device("cuda", 0)
n = 64 * 2**20
which = sys.argv[1:] or ["hist", "minmax", "qdq", "qdq_bf16", "pc", "ste"]
x = torch.randn(n, device=dev) * 2 + 2
xb = x.to(torch.bfloat16)
g = torch.randn(n, device=dev)
blk = StateArena.for_device(dev).allocate(2)
ops.stats_update_impl(x, blk.arena, blk.first, ops.QUANTIZATION_TF_ENHANCED, None, 0)
params = ops.per_channel_params([-4.0] * 2048, [8.0] * 2048, 8).cuda()
torch.cuda.synchronize()
for _ in range(3):
    if "hist" in which:
        ops.stats_update_impl(x, blk.arena, blk.first, ops.QUANTIZATION_TF_ENHANCED, None, 0)
        ops.stats_update_impl(xb, blk.arena, blk.first, ops.QUANTIZATION_TF_ENHANCED, None, 0)
    if "minmax" in which:
        ops.stats_update_impl(x, blk.arena, blk.first + 1, ops.QUANTIZATION_TF, None, 0)
    if "qdq" in which:
        ops.qdq_per_tensor_impl(x, -4.0, 8.0, 8, 0, 0)
    if "qdq_bf16" in which:
        ops.qdq_per_tensor_impl(xb, -4.0, 8.0, 8, 0, 0)
    if "pc" in which:
        ops.qdq_per_channel_impl(x, params, 2048, n // 2048, 0, 0)
        ops.qdq_per_channel_impl(xb, params, 2048, n // 2048, 0, 0)
    if "ste" in which:
        ops.ste_bwd_impl(x, g, -4.0, 8.0)
    if "lg" in which:   # range learning: per-tensor asymmetric fp32 / bf16, per-channel symmetric fp32
        mn1, mx1 = torch.tensor([-3.0], device=dev), torch.tensor([4.0], device=dev)
        ops.lg_qdq_fwd_impl(x, mn1, mx1, 8, ops.LG_ASYMMETRIC, False, 0, gate=True)
        ops.lg_qdq_bwd_impl(x, g, mn1, mx1, 8, ops.LG_ASYMMETRIC)
        ops.lg_qdq_fwd_impl(xb, mn1.bfloat16(), mx1.bfloat16(), 8, ops.LG_ASYMMETRIC, False, 0, gate=True)
        ops.lg_qdq_bwd_impl(xb, g.bfloat16(), mn1.bfloat16(), mx1.bfloat16(), 8, ops.LG_ASYMMETRIC)
        mnc, mxc = torch.full((2048,), -4.0, device=dev), torch.full((2048,), 4.0, device=dev)
        xc, gc = x.view(2048, -1), g.view(2048, -1)
        ops.lg_qdq_fwd_impl(xc, mnc, mxc, 8, ops.LG_SIGNED_SYMMETRIC, False, 0, gate=True)
        ops.lg_qdq_bwd_impl(xc, gc, mnc, mxc, 8, ops.LG_SIGNED_SYMMETRIC)
torch.cuda.synchronize()
print("ok")
