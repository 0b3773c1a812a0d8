"""Known-answer vectors transcribed from the reference's own unit tests (values only; citations per entry).
DlQ = /root/reference/ModelOptimizations/DlQuantization ; TEt = /root/reference/TrainingExtensions/torch
"""
import numpy as np

SIX = np.array([-0.5, -0.25, 0, 0.25, 0.5, 0.75], dtype=np.float32)

# DlQ/test/TestTensorQuantizationSim.cpp:51-159 -- (min, max, bw, expected); EXPECT_FLOAT_EQ == within 4 ULP
QDQ_KATS = [
    (-0.46, 0.72, 8, [-0.45811754, -0.2498823, 0.0, 0.2498823, 0.49976459, 0.72188222]),   # :51-76
    (0.5, 1.0, 8, [0.0, 0.0, 0.0, 0.25098041, 0.49803925, 0.74901962]),                      # :78-105 gated min
    (0.5, 0.5, 8, [0.0, 0.0, 0.0, 0.24901962, 0.5, 0.5]),                                    # :107-134 min == max
    (-0.5, -0.1, 8, [-0.5, -0.24901962, 0.0, 0.0, 0.0, 0.0]),                                # :136-159 gated max
]
# DlQ/test/TestTensorQuantizationSim.cpp:161-185 and :212-236 -- exact integer grids
GRID_KATS = [
    (-0.46, 0.72, 8, False, [0, 45, 99, 153, 207, 255]),
    (-0.46, 0.72, 8, True, [-128, -83, -29, 25, 79, 127]),
]
# DlQ/test/TestTfEnhancedEncodingAnalyzer.cpp:176-197 -- all-zero input, bw 8
TFE_ALL_ZERO = dict(min=-1.00392, max=0.996078, offset=-128, tol=1e-5)
# DlQ/test/TestTensorQuantizer.cpp:126-133 as reproduced with g++ 13.3: N(2,2), mt19937(1), 6000
TFE_N22 = dict(min=-6.52710772, max=8.88411903, delta=0.0604361817, offset=-108, qdq5=5.01620293)
