// Percentile calibration on the 512-bin PDF, host + device.
//
// Reference: PercentileEncodingAnalyzer<float> (DlQ/src/PercentileEncodingAnalyzer.cpp:77-196): statistics are the same
// UpdatePdf running-mean PDF the tf_enhanced analyzer keeps (:69-75); the encoding clips the observed range to the bins
// where the cumulative distribution crosses (1 - p/100) from the left and p/100 from the right, then goes through
// getComputedEncodings like the tf scheme. The mixed float / double typing of every expression is the reference's.
#pragma once
#include "encoding_math.h"
#include "tfe_math.h"

namespace ab
{
namespace pct
{

// computeEncoding's all-zero branch (:91-105): unlike the tf_enhanced analyzer it has already taken the strict-symmetric
// step off num_steps at this point (:84-89)
AB_HD void all_zero_encoding(int bw, bool symmetric_and_strict, ab_encoding& e)
{
    float num_steps = (float) (em::pow2((uint8_t) bw) - 1);
    if (symmetric_and_strict)
        num_steps -= 1;
    e.min    = -1;
    e.max    = 1;
    e.delta  = (e.max - e.min) / (int) num_steps;
    e.offset = floor(e.min / e.delta);
    e.min    = e.offset * e.delta;
    e.max    = e.min + (int) num_steps * e.delta;
    e.bw     = (uint8_t) bw;
}

// _computePercentileRange (:125-196). `cdf` is scratch for AB_PDF_SIZE doubles.
AB_HD void percentile_range(const tfe::PdfView& p, double* cdf, float percentile, float& out_min, float& out_max)
{
    float min_val, max_val;
    tfe::find_range(p, min_val, max_val);   // findOriginalRange (math_functions.cpp:404-436) == the tf_enhanced one
    if (percentile == 100.0f)
    {
        out_min = min_val, out_max = max_val;
        return;
    }
    const float bin_width = (float) (p.x_left(1) - p.x_left(0));
    const float hist_min  = (float) p.x_left(0);
    const float hist_max  = (float) (p.x_left(AB_PDF_SIZE - 1) + bin_width);
    float pmin = hist_min, pmax = hist_max;

    cdf[0] = p.pdf[0];
    for (int i = 1; i < AB_PDF_SIZE; ++i)
        cdf[i] = p.pdf[i] + cdf[i - 1];

    const float left = 1 - percentile / 100;
    for (int i = 0; i < AB_PDF_SIZE; ++i)
        if (cdf[i] >= left)
        {
            pmin = (float) p.x_left(i);
            break;
        }
    const float right = percentile / 100;
    for (int i = AB_PDF_SIZE - 1; i >= 0; --i)
        if (cdf[i] < right && p.x_left(i) < max_val)   // never beyond the largest value seen
        {
            pmax = (float) (p.x_left(i) + bin_width);
            break;
        }
    if (pmin == pmax)
        pmax += bin_width;
    out_min = pmin, out_max = pmax;
}

// computeEncoding for an initialised PDF (:107-122)
AB_HD void encoding(const tfe::PdfView& p, double* cdf, float percentile, int bw, bool sym, bool strict, bool unsigned_sym,
                    ab_encoding& e)
{
    float a_min, a_max;
    percentile_range(p, cdf, percentile, a_min, a_max);
    a_min = em::smin(a_min, 0.0f);
    a_max = em::smax(a_max, 0.0f);
    em::tf_encoding(bw, a_min, a_max, sym, strict, unsigned_sym, e);
}

}   // namespace pct
}   // namespace ab
