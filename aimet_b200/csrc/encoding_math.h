// Encoding arithmetic shared by the host entry points and the device kernels (compiled for both).
//
// Everything here must give bit-identical results on the host (g++ via nvcc, -ffp-contract=off) and on the
// device (--fmad=false, IEEE double/float). The expressions are typed exactly as the reference types them
// (which operand is float, which is double, where int converts), because that typing decides the last bit.
// Citations: DlQ = /root/reference/ModelOptimizations/DlQuantization.
#pragma once

#include <float.h>
#include <limits.h>
#include <math.h>
#include <stdint.h>

#include "../../include/aimet_b200.h"

#if defined(__CUDACC__)
#define AB_HD __host__ __device__ __forceinline__
#else
#define AB_HD inline
#endif

namespace ab
{
namespace em
{
constexpr double kEpsilon  = 1e-5;   // DlQ/src/quantization_utils.hpp:51
constexpr double kMinRange = 0.01;   // DlQ/src/TfEncodingAnalyzer.h:79, TfEnhancedEncodingAnalyzer.h:105

// libstdc++'s std::min / std::max selection rule (matters for NaN and signed zeros)
template <typename T>
AB_HD T smin(T a, T b)
{
    return (b < a) ? b : a;
}
template <typename T>
AB_HD T smax(T a, T b)
{
    return (a < b) ? b : a;
}

// float/double -> int as x86-64's cvtt* instructions do it: NaN and out-of-range give INT_MIN. The reference's
// compiled code relies on that for out-of-range histogram samples and candidate indices.
AB_HD int f2i_x86(float v)
{
    if (!(v > -2147483904.0f && v < 2147483648.0f))
        return INT_MIN;
    return (int) v;
}
AB_HD int d2i_x86(double v)
{
    if (!(v > -2147483649.0 && v < 2147483648.0))
        return INT_MIN;
    return (int) v;
}
AB_HD int iadd_wrap(int a, int b)
{
    return (int) ((unsigned) a + (unsigned) b);
}

AB_HD double pow2(int bw)
{
    return ldexp(1.0, bw);   // == pow(2, bw) exactly
}

// round / roundf with identical semantics on both sides (half away from zero)
AB_HD float roundf_away(float v)
{
    return ::roundf(v);
}
AB_HD double round_away(double v)
{
    return ::round(v);
}

// gateMinMax -- DlQ/src/quantization_utils.cpp:145-156
AB_HD void gate_min_max(double& mn, double& mx)
{
    mn = smin(mn, 0.0);
    mx = smax(mx, 0.0);
    mx = smax(mx, mn + kEpsilon);
}

// fillEncodingInfo + generateScaleOffset -- DlQ/src/TensorQuantizationSim.cpp:63-92, trim_functions.cpp:61-73
AB_HD void fill_encoding_info(int bw, double mn, double mx, ab_encoding& e)
{
    bw   = (uint8_t) bw;
    e.bw = bw;
    gate_min_max(mn, mx);
    double steps = pow2(bw) - 1;
    if (mn == -mx)
        steps -= 1;
    e.delta  = (mx - mn) / steps;
    e.offset = round_away(mn / e.delta);
    e.min    = e.offset * e.delta;
    e.max    = e.delta * steps + e.min;
}

// getComputedEncodings -- DlQ/src/quantization_utils.cpp:58-143
AB_HD void tf_encoding(int bw, double mn, double mx, bool sym, bool strict, bool unsigned_sym, ab_encoding& e)
{
    bw           = (uint8_t) bw;
    double steps = pow2(bw) - 1;
    if (sym && strict)
        steps -= 1;
    e.bw = bw;
    if (isinf(mn))
        mn = -(double) FLT_MAX;
    if (isinf(mx))
        mx = (double) FLT_MAX;
    if (sym && ((mn < 0.0) || !unsigned_sym))
    {
        mx                     = smax(fabs(mx), fabs(mn));
        unsigned int pos_steps = (unsigned int) floor(steps / 2);
        e.delta                = mx / pos_steps;
        e.offset               = -ceil(steps / 2);
        e.min                  = smax(e.offset * e.delta, -(double) FLT_MAX);
        e.max                  = smin(e.delta * pos_steps, (double) FLT_MAX);
        return;
    }
    e.delta = (mx - mn) / steps;
    if (mn < 0 && mx > 0)
    {
        double b_zero = round_away(-mn / e.delta);
        b_zero        = smin(steps, smax(0.0, b_zero));
        e.offset      = -b_zero;
    }
    else
    {
        e.offset = round_away(mn / e.delta);
        e.min    = mn;
        e.max    = mx;
        return;
    }
    const double lo = e.delta * e.offset;
    e.min           = (lo >= -(double) FLT_MAX && lo <= (double) FLT_MAX) ? lo : -(double) FLT_MAX;
    e.max           = mx - mn + e.min;
    if (e.max > (double) FLT_MAX)
        e.max = (double) FLT_MAX;
}

// TfEncodingAnalyzer::computeEncoding -- DlQ/src/TfEncodingAnalyzer.cpp:81-101
AB_HD void tf_analyzer_encoding(int bw, double run_min, double run_max, bool sym, bool strict, bool unsigned_sym,
                                ab_encoding& e)
{
    double new_min = smin(0.0, run_min);
    double new_max = smax(0.0, run_max);
    new_max        = smax(new_max, new_min + kMinRange);
    tf_encoding(bw, new_min, new_max, sym, strict, unsigned_sym, e);
}

// One channel of the per-channel preparation AimetTensorQuantizer::quantizeDequantizePerChannel performs with torch
// fp32 CPU ops (TrainingExtensions/torch/src/AimetTensorQuantizer.cpp:236-254, 272-299).
AB_HD void per_channel_param(double enc_min, double enc_max, float steps_f, float& o_min, float& o_max,
                             float& o_delta, float& o_offset)
{
    float mn       = (float) enc_min;
    float mx       = (float) enc_max;
    mn             = (mn < 0.0f) ? mn : 0.0f;      // torch.minimum(min, 0)
    mx             = (mx > 0.0f) ? mx : 0.0f;      // torch.maximum(max, 0)
    const float lo = mn + (float) 1e-5;            // tensor + Scalar: the scalar narrows to fp32
    mx             = (mx > lo) ? mx : lo;
    const float d  = (mx - mn) / steps_f;          // tensor / Scalar on the CPU: a true fp32 division
    o_min          = mn;
    o_max          = mx;
    o_delta        = d;
    o_offset       = ::nearbyintf(mn / d);         // at::round: half to even
}

// InitializePdf (signed) -- DlQ/src/math_functions.cpp:207-241. Produces xLeft[0] and the double bucket size
// (xLeft[i] = x_left0 + i * bucket_d), and the two floats UpdatePdf derives from them (:264-268).
AB_HD void init_pdf_range(float min_val, float max_val, double& x_left0, double& bucket_d, float& bucket_f,
                          float& pdf_offset_f)
{
    if (min_val == max_val)
        max_val = smax(max_val, min_val + (float) 0.01);
    const float center = (max_val + min_val) / 2;
    min_val            = smax(-FLT_MAX, center - 3 * (center - min_val));
    max_val            = smin(FLT_MAX, center + 3 * (max_val - center));
    bucket_d           = ((double) max_val - (double) min_val) / AB_PDF_SIZE;
    x_left0            = min_val + 0 * bucket_d;
    const double x1    = min_val + 1 * bucket_d;
    bucket_f           = (float) (x1 - x_left0);
    const float mn_f   = (float) x_left0;
    pdf_offset_f       = mn_f / bucket_f;
}

// xLeft[i] as InitializePdf stores it: float(min) + i * bucket (double multiply, then double add)
AB_HD double x_left_at(double x_left0, double bucket_d, int i)
{
    return x_left0 + i * bucket_d;
}

}   // namespace em
}   // namespace ab
