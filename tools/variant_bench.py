"""Tuning helper: QDQ / STE bandwidth at 1 GiB for whatever library AIMET_B200_LIB points at."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aimet_b200 import ops  # noqa: E402
import tools.microbench as mb  # noqa: E402

dev = torch.device("cuda", 0)
for dtype, es in ((torch.float32, 4), (torch.bfloat16, 2)):
    n = 1024 * 2**20 // es
    pool = [(torch.randn(n, device=dev) * 2 + 2).to(dtype) for _ in range(2)]
    ms = mb.time_ms(lambda x: ops.qdq_per_tensor_impl(x, -4.0, 8.0, 8, 0, 0), pool)
    g = torch.randn(n, device=dev).to(dtype)
    ms2 = mb.time_ms(lambda x: ops.ste_bwd_impl(x, g, -4.0, 8.0), pool)
    print(os.environ.get("AIMET_B200_LIB", "default").split("/")[-1], str(dtype), "qdq GB/s", round(2 * es * n / ms / 1e6),
          "ste GB/s", round(3 * es * n / ms2 / 1e6), flush=True)
    del pool, g
