"""Register aimet_b200's drop-ins under the reference's module names.

The reference's Python reaches its native code through exactly two imports
(TrainingExtensions/common/src/python/aimet_common/aimet_tensor_quantizer.py:42 `from aimet_common.AimetTensorQuantizer
import *` and aimet_common/libpymo.py:43 `from aimet_common._libpymo import *`). After `install()` those imports resolve
to the sm_100a implementation, so an unmodified `aimet_torch.v1.quantsim.QuantizationSimModel` runs on it:

    import aimet_b200.install; aimet_b200.install.install()
    from aimet_torch.v1.quantsim import QuantizationSimModel      # the reference's own Python

Call it before anything imports aimet_common.libpymo.
"""
import sys
import types

from . import libpymo as _pymo
from . import tensor_quantizer_op as _atq

_IN_SCOPE = ["ComputationMode", "QuantizationMode", "RoundingMode", "TensorQuantizerOpMode", "TfEncoding",
             "TensorQuantizer", "EncodingAnalyzerForPython", "TensorQuantizationSimForPython", "PtrToInt64",
             "COMP_MODE_CPU", "COMP_MODE_GPU", "QUANTIZATION_TF", "QUANTIZATION_TF_ENHANCED",
             "QUANTIZATION_RANGE_LEARNING", "QUANTIZATION_PERCENTILE", "QUANTIZATION_MSE", "QUANTIZATION_ENTROPY",
             "ROUND_NEAREST", "ROUND_STOCHASTIC"]


def install(extra_pymo_names=None):
    """Put `aimet_common.AimetTensorQuantizer` and `aimet_common._libpymo` into sys.modules.

    `extra_pymo_names`: optional dict of additional attributes for the `_libpymo` stand-in (the reference's
    out-of-scope bindings -- SVD, BN fold, QnnDatatype ... -- that some of its modules import at load time; take them
    from the reference's own pure-python `aimet_common.py_libpymo`)."""
    atq = types.ModuleType("aimet_common.AimetTensorQuantizer")
    # The reference's Python makes one native call per weight channel: it gets the class that queues and batches them
    # (same methods, same results; AB_DEFER_DROPIN=0 registers the plain one-call-one-launch class instead).
    atq.AimetTensorQuantizer = _atq.DeferredAimetTensorQuantizer if _atq.DEFER_DROPIN else _atq.AimetTensorQuantizer
    atq.__all__ = ["AimetTensorQuantizer"]
    sys.modules["aimet_common.AimetTensorQuantizer"] = atq

    pymo = types.ModuleType("aimet_common._libpymo")
    names = list(_IN_SCOPE)
    for n in _IN_SCOPE:
        setattr(pymo, n, getattr(_pymo, n))
    for n, v in (extra_pymo_names or {}).items():
        if not hasattr(pymo, n):
            setattr(pymo, n, v)
            names.append(n)
    pymo.__all__ = names
    sys.modules["aimet_common._libpymo"] = pymo
    return atq, pymo
