// MSE calibration on the 512-bin PDF, host + device.
//
// Reference: MseEncodingAnalyzer<float> (DlQ/src/MseEncodingAnalyzer.cpp:77-285). Statistics are the tf_enhanced running-mean
// PDF (:70-76). The encoding picks, among all (min, max) pairs built from the histogram's bin edges on either side of zero,
// the pair whose quantizer has the least probability-weighted squared error on the bin centres (first strict minimum in
// the reference's push order), then goes through getComputedEncodings. DTYPE is float there: bin edges and centres are
// float ACCUMULATIONS (i += width), the cost accumulator is a float fed with double terms -- all reproduced literally.
#pragma once
#include "encoding_math.h"
#include "tfe_math.h"

namespace ab
{
namespace mse
{

// The reference's edge loop runs ~513 times; the cap only matters for a degenerate histogram whose width is below the
// resolution of its own edges (where the reference's `i += width` loop would not terminate at all).
constexpr int kMaxEdges = 1024;

struct Tables
{
    float edges[kMaxEdges + 2];     // binEdges: observed min, then every histogram edge inside [min, max + width]
    float centers[kMaxEdges + 2];   // binCentersPdf[i].first
    float cpdf[kMaxEdges + 2];      // binCentersPdf[i].second
    int n_edges, n_centers;
    int n_neg, first_pos, n_pos;    // edges < 0 are edges[0 .. n_neg); edges > 0 are edges[first_pos .. first_pos + n_pos)
    float min_val, max_val;         // the default answer: observed range, max extended by one bin (:151-154, :190)
};

// _minimizeMSE up to the candidate loop (:139-186)
AB_HD void build_tables(const tfe::PdfView& p, Tables& t)
{
    const float width    = (float) (p.x_left(1) - p.x_left(0));
    const float hist_min = (float) p.x_left(0);
    const float hist_max = (float) (p.x_left(AB_PDF_SIZE - 1) + width);
    float min_val, max_val;
    tfe::find_range(p, min_val, max_val);   // findOriginalRange
    max_val = max_val + width;
    t.min_val = min_val, t.max_val = max_val;

    int n        = 0;
    t.edges[n++] = min_val;
    int guard    = 0;
    for (float i = hist_min; i <= hist_max && n < kMaxEdges && guard < 4 * kMaxEdges; i += width, ++guard)
        if (i >= min_val && i <= max_val)
            t.edges[n++] = i;
    t.n_edges = n;
    // edges ascend (min_val first, then increasing values >= min_val): negatives form a prefix, positives a suffix
    int n_neg = 0;
    while (n_neg < n && t.edges[n_neg] < 0)
        ++n_neg;
    int first_pos = n_neg;
    while (first_pos < n && !(t.edges[first_pos] > 0))
        ++first_pos;
    t.n_neg = n_neg, t.first_pos = first_pos, t.n_pos = n - first_pos;

    const float pdf_start = (float) p.x_left(0);
    const float pdf_step  = (float) (p.x_left(1) - p.x_left(0));
    t.n_centers           = n - 1;
    float c               = min_val + width / 2;
    for (int i = 0; i < t.n_centers; ++i)
    {
        if (i > 0)
            c = c + width;
        int ind      = em::f2i_x86(floorf((c - pdf_start) / pdf_step));
        ind          = em::smin(em::smax(0, ind), AB_PDF_SIZE - 1);
        t.centers[i] = c;
        t.cpdf[i]    = (float) p.pdf[ind];
    }
}

AB_HD int num_candidates(const Tables& t)
{
    return (t.n_neg + 1) * (t.n_pos + 1) - 1;   // every (min, max) pair except the trailing {0, 0} (:204-238)
}

// candidate k in push order: min candidates outer (negative edges ascending, then 0), max candidates inner
AB_HD void candidate(const Tables& t, int k, float& cmin, float& cmax)
{
    const int n_max = t.n_pos + 1;
    const int i     = k / n_max;
    const int j     = k - i * n_max;
    cmin            = (i < t.n_neg) ? t.edges[i] : 0.0f;
    cmax            = (j < t.n_pos) ? t.edges[t.first_pos + j] : 0.0f;
}

// _computeMSECost (:241-264)
AB_HD float cost(const Tables& t, int bw, float cmin, float cmax, bool sym, bool strict, bool unsigned_sym)
{
    ab_encoding e;
    em::tf_encoding(bw, cmin, cmax, sym, strict, unsigned_sym, e);
    float acc = 0;
    for (int i = 0; i < t.n_centers; ++i)
    {
        const float fv      = t.centers[i];
        const float clamped = em::smax(cmin, em::smin(fv, cmax));
        const int quantized = em::d2i_x86(em::round_away(clamped / e.delta - e.offset));
        const float deq     = (float) (e.delta * (quantized + e.offset));
        const double diff   = (double) (fv - deq);
        acc                 = (float) ((double) acc + (double) t.cpdf[i] * (diff * diff));
    }
    return acc;
}

// computeEncoding once the best pair is known (:113-122)
AB_HD void finish(int bw, float a_min, float a_max, bool sym, bool strict, bool unsigned_sym, ab_encoding& e)
{
    a_min = em::smin(a_min, 0.0f);
    a_max = em::smax(a_max, 0.0f);
    em::tf_encoding(bw, a_min, a_max, sym, strict, unsigned_sym, e);
}

// whole computeEncoding for an initialised PDF, sequentially (host side; the device spreads the candidates over threads)
AB_HD void encoding(const tfe::PdfView& p, Tables& t, int bw, bool sym, bool strict, bool unsigned_sym, ab_encoding& e)
{
    build_tables(p, t);
    float best_min = t.min_val, best_max = t.max_val;
    float mse_min  = FLT_MAX;
    const int n    = num_candidates(t);
    for (int k = 0; k < n; ++k)
    {
        float cmin, cmax;
        candidate(t, k, cmin, cmax);
        const float c = cost(t, bw, cmin, cmax, sym, strict, unsigned_sym);
        if (c < mse_min)
        {
            mse_min  = c;
            best_min = cmin, best_max = cmax;
        }
    }
    finish(bw, best_min, best_max, sym, strict, unsigned_sym, e);
}

}   // namespace mse
}   // namespace ab
