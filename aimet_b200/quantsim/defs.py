"""Enums and maps of the reference's Python layer that the hot path needs (same names and values).

Reference: TrainingExtensions/common/src/python/aimet_common/defs.py (QuantScheme :49-76, QuantizationDataType,
MAP_QUANT_SCHEME_TO_PYMO :79-96, MAP_ROUND_MODE_TO_PYMO :97-98).
"""
import enum

from .. import libpymo


class QuantScheme(enum.Enum):
    post_training_tf = 1
    post_training_tf_enhanced = 2
    training_range_learning_with_tf_init = 3
    training_range_learning_with_tf_enhanced_init = 4
    training_range_learning = 5
    post_training_percentile = 6

    @classmethod
    def from_str(cls, alias: str) -> "QuantScheme":
        try:
            return _ALIASES[alias]
        except KeyError:
            raise ValueError(f"Invalid string literal {alias}. Expected one of {list(_ALIASES)}") from None


_ALIASES = {
    "tf": QuantScheme.post_training_tf,
    "tf_enhanced": QuantScheme.post_training_tf_enhanced,
    "percentile": QuantScheme.post_training_percentile,
    "min_max": QuantScheme.post_training_tf,
    "post_training_tf": QuantScheme.post_training_tf,
    "post_training_tf_enhanced": QuantScheme.post_training_tf_enhanced,
}


class QuantizationDataType(enum.Enum):
    undefined = 0
    int = 1
    float = 2


MAP_QUANT_SCHEME_TO_PYMO = {
    QuantScheme.post_training_tf: libpymo.QuantizationMode.QUANTIZATION_TF,
    QuantScheme.post_training_tf_enhanced: libpymo.QuantizationMode.QUANTIZATION_TF_ENHANCED,
    QuantScheme.post_training_percentile: libpymo.QuantizationMode.QUANTIZATION_PERCENTILE,
    # range learning initialises from tf / tf_enhanced statistics (reference defs.py:84-91)
    QuantScheme.training_range_learning_with_tf_init: libpymo.QuantizationMode.QUANTIZATION_TF,
    QuantScheme.training_range_learning_with_tf_enhanced_init: libpymo.QuantizationMode.QUANTIZATION_TF_ENHANCED,
}
MAP_ROUND_MODE_TO_PYMO = {"nearest": libpymo.RoundingMode.ROUND_NEAREST,
                          "stochastic": libpymo.RoundingMode.ROUND_STOCHASTIC}
