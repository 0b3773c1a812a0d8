"""Generate tests/golden/range_learning.npz from the reference's UNMODIFIED range-learning code
(aimet_torch/v1/tensor_quantizer.py QuantizeDequantizeFunc + set_encoding_min_max_gating_threshold, which call
v1/quantsim_straight_through_grad.py). Run in the build container (needs /root/reference):

    python tests/golden/make_range_learning_golden.py
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import ref_python_env  # noqa: F401,E402  (stubs + sys.path for the reference)

from aimet_torch.v1.tensor_quantizer import QuantizeDequantizeFunc, set_encoding_min_max_gating_threshold  # noqa: E402
from aimet_torch.v1.tensor_quantizer import LearnedGridTensorQuantizer  # noqa: E402
from aimet_common.defs import QuantScheme, QuantizationDataType  # noqa: E402
from aimet_common import libpymo  # noqa: E402

from make_range_learning_cases import CASES, make_inputs  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "range_learning.npz")

def main():
    out = {}
    for idx, (name, shape, dtype, bw, sym, strict, unsigned, axis, dist, enc) in enumerate(CASES):
        x, grad, mn, mx = make_inputs(idx, shape, dtype, axis, dist, enc)
        q = LearnedGridTensorQuantizer(bw, libpymo.RoundingMode.ROUND_NEAREST,
                                       QuantScheme.training_range_learning_with_tf_init, sym, True,
                                       QuantizationDataType.int)
        q.use_strict_symmetric = strict
        q.use_unsigned_symmetric = unsigned
        q.is_unsigned_symmetric = unsigned
        q._ch_axis = 0 if axis is None else axis
        out[name + ".x"] = x.float().numpy()
        out[name + ".grad"] = grad.float().numpy()
        out[name + ".min_in"] = mn.float().numpy()
        out[name + ".max_in"] = mx.float().numpy()
        p_min = torch.nn.Parameter(mn.clone())
        p_max = torch.nn.Parameter(mx.clone())
        set_encoding_min_max_gating_threshold(p_min, p_max)
        xin = x.clone().requires_grad_(True)
        y = QuantizeDequantizeFunc.apply(xin, p_min, p_max, q)
        y.backward(grad)
        out[name + ".min_gated"] = p_min.detach().float().numpy()
        out[name + ".max_gated"] = p_max.detach().float().numpy()
        out[name + ".y"] = y.detach().float().numpy()
        out[name + ".grad_x"] = xin.grad.float().numpy()
        out[name + ".grad_min"] = p_min.grad.float().numpy()
        out[name + ".grad_max"] = p_max.grad.float().numpy()
        print(name, "y[:4]", out[name + ".y"].ravel()[:4], "gmin", out[name + ".grad_min"][:2], "gmax",
              out[name + ".grad_max"][:2])
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
