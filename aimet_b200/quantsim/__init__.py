"""Host mirror of the reference's Python layers on the quantization-simulation hot path."""
from .defs import MAP_QUANT_SCHEME_TO_PYMO, MAP_ROUND_MODE_TO_PYMO, QuantizationDataType, QuantScheme  # noqa: F401
from .qc_quantize_op import (QcPostTrainingWrapper, QcQuantizeOpMode, QcQuantizeWrapper,  # noqa: F401
                             StaticGridQuantWrapper, SteGatingFuncForParameters)
from .quantsim import QuantizationSimModel, load_checkpoint, save_checkpoint  # noqa: F401
from .tensor_quantizer import (Quantize, QuantizeDequantize, StaticGridPerChannelQuantizer,  # noqa: F401
                               StaticGridPerTensorQuantizer, StaticGridTensorQuantizer, compute_dloss_by_dx)
from .learned_grid import (LearnedGridQuantWrapper, LearnedGridTensorQuantizer,  # noqa: F401
                           set_encoding_min_max_gating_threshold)
from .quant_analyzer import CallbackFunc, QuantAnalyzer  # noqa: F401,E402
from .adaround import Adaround, AdaroundParameters  # noqa: F401,E402
from .auto_quant import AutoQuant  # noqa: F401,E402
from .batch_norm_fold import fold_all_batch_norms  # noqa: F401,E402
