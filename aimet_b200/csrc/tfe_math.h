// tf_enhanced (TFE) encoding search: candidate generation, the quantisation+saturation cost, and the final
// encoding -- compiled for host and device from one source so the two cannot drift.
//
// Restates TfEnhancedEncodingAnalyzer<float> (DlQ/src/TfEnhancedEncodingAnalyzer.cpp:79-397) operation by
// operation: DTYPE is float there, so many intermediates are deliberately narrowed to float.
#pragma once

#include "encoding_math.h"

namespace ab
{
namespace tfe
{
using namespace ab::em;

constexpr float kGamma        = 3.0f;   // DlQ/src/TfEnhancedEncodingAnalyzer.h:102
constexpr int kAsymDeltas     = 17;     // f = 1/16 .. 17/16 (TfEnhancedEncodingAnalyzer.cpp:196)
constexpr int kAsymOffsets    = 21;     // i = 0 .. 20       (:200)
constexpr int kAsymCandidates = kAsymDeltas * kAsymOffsets + 1;
constexpr int kMaxSymDeltas   = 104;    // the float loop runs 101 times; leave head-room
constexpr int kMaxCandidates  = kAsymCandidates;   // 358

// read-only view of one quantizer's PDF
struct PdfView
{
    const double* pdf;   // 512 probabilities
    double x_left0;      // xLeft[0]
    double bucket_d;     // xLeft spacing used to build xLeft[i]
    AB_HD double x_left(int i) const { return x_left_at(x_left0, bucket_d, i); }
};

// _findRangeOfAggregateStats -- :256-291
AB_HD void find_range(const PdfView& p, float& min_val, float& max_val)
{
    min_val = (float) p.x_left(0);
    max_val = (float) p.x_left(AB_PDF_SIZE - 1);
    for (int i = 0; i < AB_PDF_SIZE; ++i)
        if (p.pdf[i] > 0)
        {
            min_val = (float) p.x_left(i);
            break;
        }
    for (int i = AB_PDF_SIZE - 1; i > 0; --i)   // stops at bin 1, never looks at bin 0
        if (p.pdf[i] > 0)
        {
            max_val = (float) p.x_left(i);
            break;
        }
    min_val = smin(min_val, 0.0f);
    max_val = smax(max_val, 0.0f);
    max_val = smax(max_val, min_val + (float) kMinRange);
}

// numSteps as getComputedEncodings computes it -- :363-373
AB_HD float num_steps_for(int bw, bool sym, bool strict)
{
    float steps = (float) (pow2(bw) - 1);
    if (sym && strict)
        steps -= 1;
    return steps;
}

// ---- asymmetric candidates -- _pickTestCandidatesAsymmetric :178-214, _clampToObservedMinMax :146-175 ----------
struct AsymSetup
{
    float obs_min, obs_max, obs_delta, num_steps;
    int obs_offset;
};

AB_HD AsymSetup asym_setup(float min_val, float max_val, float num_steps)
{
    AsymSetup s;
    s.num_steps  = num_steps;
    s.obs_delta  = (float) (((double) max_val - (double) min_val) / num_steps);
    s.obs_offset = f2i_x86(roundf_away(min_val / s.obs_delta));
    s.obs_min    = smax(s.obs_delta * s.obs_offset, -FLT_MAX);
    s.obs_max    = smin(s.obs_delta * (s.obs_offset + num_steps), FLT_MAX);
    return s;
}

// candidate k in the reference's push order; returns false where the reference `continue`s
AB_HD bool asym_candidate(const AsymSetup& s, int k, float& delta, int& offset)
{
    if (k == kAsymCandidates - 1)
    {
        delta  = s.obs_delta;
        offset = s.obs_offset;
        return true;
    }
    const int a   = k / kAsymOffsets;
    const int i   = k - a * kAsymOffsets;
    const float f = (float) (a + 1) * 0.0625f;   // the float loop f += 1/16 is exact
    delta         = f * s.obs_delta;
    offset        = d2i_x86(-s.num_steps + s.num_steps / 20.0 * i);
    float t_min   = smax(delta * offset, -FLT_MAX);
    float t_max   = smin(delta * (offset + s.num_steps), FLT_MAX);
    if ((t_min < s.obs_min) && (t_max > s.obs_max))
        return false;
    t_min = smax(s.obs_min, t_min);
    t_max = smin(s.obs_max, t_max);
    if (t_min == t_max)
        return false;
    delta  = (float) (((double) t_max - t_min) / s.num_steps);
    offset = f2i_x86(roundf_away(t_min / delta));
    return true;
}

// ---- symmetric candidates -- _pickTestCandidatesSymmetric :217-253 -----------------------------------------------
// Fills deltas[] (at most kMaxSymDeltas) and returns the count; all candidates share `offset`.
AB_HD int sym_candidates(float min_val, float max_val, float num_steps, bool unsigned_sym, float* deltas,
                         int& offset)
{
    float delta_max;
    if ((min_val == 0.0) && unsigned_sym)
    {
        delta_max = max_val / num_steps;
        offset    = 0;
    }
    else
    {
        const float abs_max = smax(fabsf(max_val), fabsf(min_val));
        delta_max           = (float) (abs_max / (num_steps / 2.0));
        offset              = f2i_x86(floorf(-num_steps / 2));
    }
    int n = 0;
    // `for (DTYPE f = 1.0/100; f <= 1 + 1.0/100; f += 1.0/100)`: the sum is formed in double and narrowed each time
    for (float f = (float) (1.0 / 100); f <= 1 + 1.0 / 100 && n < kMaxSymDeltas; f = (float) (f + 1.0 / 100))
        deltas[n++] = f * delta_max;
    return n;
}

// ---- cost -- _quantAndSatCost :294-355 ---------------------------------------------------------------------------
AB_HD double cost(const PdfView& p, int bw, float delta, int offset)
{
    const float min_val   = delta * offset;
    const float step_size = (float) (pow2(bw) - 1);
    const float max_val   = delta * (offset + step_size);
    const float pdf_start = (float) p.x_left(0);
    const double pdf_step = p.x_left(1) - p.x_left(0);
    int min_ind           = d2i_x86(floor((min_val - pdf_start) / pdf_step));
    min_ind               = smin(smax(0, min_ind), AB_PDF_SIZE - 1);
    int max_ind           = d2i_x86(floor((max_val - pdf_start) / pdf_step));
    max_ind               = smin(smax(0, max_ind), AB_PDF_SIZE - 1);
    const float min_mid   = (float) (pdf_start + (min_ind * pdf_step) + pdf_step / 2);
    const float max_mid   = (float) (pdf_start + (max_ind * pdf_step) + pdf_step / 2);
    const float offset_f  = (float) offset;

    // The reference runs three loops (bottom saturation, top saturation, quantisation), each visiting its bins in
    // ascending order into its own accumulator. One ascending pass with three accumulators adds the same terms in
    // the same order to each accumulator, so every sum is bit-identical.
    // An empty bin contributes pr * (d * d) = +0.0 * finite = +0.0 (with a finite histogram geometry and a finite, non-zero
    // delta every d is finite and its square cannot overflow a double), and adding +0.0 changes no accumulator, so empty
    // bins are skipped: the PDF spans three times the first batch's range (InitializePdf), i.e. most of its 512 bins are
    // empty. With a degenerate geometry (a range fixed from +-inf inputs) nothing is skipped: there 0 * NaN must poison
    // the sum exactly as it does in the reference.
    const bool skip_empty = (pdf_start - pdf_start == 0.0f) && (pdf_step - pdf_step == 0.0) && (delta - delta == 0.0f) &&
                            delta != 0.0f;
    double sat_bottom = 0, sat_top = 0, quant = 0;
    for (int i = 0; i < AB_PDF_SIZE; ++i)
    {
        const double pr  = p.pdf[i];
        if (skip_empty && pr == 0.0)
            continue;
        const double mid = pdf_start + i * pdf_step + pdf_step / 2;
        if (i < min_ind)
        {
            const double d = mid - min_mid;
            sat_bottom += pr * (d * d);
        }
        if (i >= max_ind)
        {
            const double d = mid - max_mid;
            sat_top += pr * (d * d);
        }
        if (i >= min_ind && i < max_ind)
        {
            const float float_val   = (float) mid;
            const int quantized     = f2i_x86(roundf_away(float_val / delta - offset_f));
            const float dequantized = delta * (float) iadd_wrap(quantized, offset);
            const double d          = (double) (float_val - dequantized);
            quant += pr * (d * d);
        }
    }
    const double sqnr = kGamma * (sat_bottom + sat_top) + quant;
    return smin(sqnr, DBL_MAX);
}

// encoding from the winning candidate -- getComputedEncodings :386-397
AB_HD void finish(float best_delta, int best_offset, float num_steps, int bw, ab_encoding& e)
{
    const float best_min = smax(best_delta * best_offset, -FLT_MAX);
    const float best_max = smin(best_delta * (best_offset + num_steps), FLT_MAX);
    e.delta              = best_delta;
    e.offset             = best_offset;
    e.bw                 = (uint8_t) bw;
    e.min                = best_min;
    e.max                = best_max;
}

// statistics were updated but only ever saw zeros -- computeEncoding :85-100
AB_HD void all_zero_encoding(int bw, ab_encoding& e)
{
    const float num_steps = (float) (pow2((uint8_t) bw) - 1);
    e.min                 = -1;
    e.max                 = 1;
    e.delta               = (e.max - e.min) / (int) num_steps;
    e.offset              = floor(e.min / e.delta);
    e.min                 = e.offset * e.delta;
    e.max                 = e.min + (int) num_steps * e.delta;
    e.bw                  = (uint8_t) bw;
}

}   // namespace tfe
}   // namespace ab
