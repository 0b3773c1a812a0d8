"""Range-learning backward at 64 MB (bf16 per tensor, bf16 / fp32 per channel): the command ncu wraps."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aimet_b200 import ops  # noqa: E402

dev = torch.device("cuda", 0)
n = 32 * 2**20
xb = (torch.randn(n, device=dev) * 2).to(torch.bfloat16)
gb = torch.randn(n, device=dev).to(torch.bfloat16)
xf, gf = torch.randn(n // 2, device=dev) * 2, torch.randn(n // 2, device=dev)
mn1, mx1 = torch.tensor([-3.0], device=dev).bfloat16(), torch.tensor([4.0], device=dev).bfloat16()
mnc, mxc = torch.full((2048,), -4.0, device=dev), torch.full((2048,), 4.0, device=dev)
torch.cuda.synchronize()
for _ in range(3):
    ops.lg_qdq_bwd_impl(xb, gb, mn1, mx1, 8, ops.LG_ASYMMETRIC)
    ops.lg_qdq_bwd_impl(xb.view(2048, -1), gb.view(2048, -1), mnc.bfloat16(), mxc.bfloat16(), 8, ops.LG_SIGNED_SYMMETRIC)
    ops.lg_qdq_bwd_impl(xf.view(2048, -1), gf.view(2048, -1), mnc, mxc, 8, ops.LG_SIGNED_SYMMETRIC)
torch.cuda.synchronize()
print("ok")
