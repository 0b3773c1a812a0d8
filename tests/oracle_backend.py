"""TEST-ONLY: the CPU-oracle-backed native op used to drive the host layer (see oracle/cpu_backend.py)."""
from oracle.cpu_backend import PortTensorQuantizer as OracleTensorQuantizer  # noqa: F401
from oracle.cpu_backend import ReferenceTensorQuantizer  # noqa: F401


# ---- range learning: the oracle as the quantize-dequantize function of aimet_b200.quantsim.learned_grid ---------------
import torch  # noqa: E402

from oracle import range_learning as _rl  # noqa: E402


class _OracleLearnedGridQdq(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, enc_min, enc_max, bw, mode, strict, ch_axis, gate):   # pylint: disable=arguments-differ
        if gate:
            _rl.gate(enc_min, enc_max)
        y, saved = _rl.forward(x.detach(), enc_min.detach().clone(), enc_max.detach().clone(), bw, mode, strict, ch_axis)
        ctx.saved = saved
        return y

    @staticmethod
    def backward(ctx, grad):   # pylint: disable=arguments-differ
        gx, gmin, gmax = _rl.backward(grad, ctx.saved)
        need = ctx.needs_input_grad
        return (gx if need[0] else None, gmin if need[1] else None, gmax if need[2] else None, None, None, None, None,
                None)


def oracle_learned_grid_qdq(tensor, encoding_min, encoding_max, quantizer, gate):
    mode = _rl.symmetry_mode(quantizer.use_symmetric_encodings, quantizer.is_unsigned_symmetric)
    return _OracleLearnedGridQdq.apply(tensor, encoding_min, encoding_max, quantizer.bitwidth, mode,
                                       quantizer.use_strict_symmetric, quantizer.channel_axis, gate)
