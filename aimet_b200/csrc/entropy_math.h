// The entropy scheme's arithmetic on its 512-bin histogram: the range bookkeeping of updateTensorHistogram
// (DlQ/src/math_functions.cpp:476-560) -- host AND device, it runs in a one-thread kernel before the binning pass -- and the
// closing KL-divergence search of EntropyEncodingAnalyzer::computeEncoding (DlQ/src/EntropyEncodingAnalyzer.cpp:98-436,
// rescaleHistogram math_functions.cpp:562-640) -- host only.
//
// Why the search runs on the host: it is a sequential walk over at most ~130 shrinking windows of 512 doubles (about
// 100 us of work), every window costs `std::log` calls, and the result is defined by glibc's correctly rounded log: a
// device log (1 ulp) would make the arg-min tie-breaks platform dependent. The reference computes every scheme's closing
// step on the host; for this scheme the histogram is read back (4 KB) and the same arithmetic is run here. The pass over
// the TENSOR -- the part that is HBM-bound -- is on the device (entropy.cu).
//
// Typing traps kept literally: `std::accumulate(first, last, 0.f)` accumulates in FLOAT; getBin() takes its arguments as
// floats; _optimizeKL returns its thresholds as floats; absoluteMax is a float.
#pragma once

#include <math.h>
#include <stddef.h>
#include <stdint.h>

#include "encoding_math.h"

namespace ab
{
namespace ent
{
constexpr int kBins = 512;

// static_cast<size_t>(float) as x86-64 gcc compiles it (the histogram's getBin depends on it for negative and huge
// quotients): v < 2^63 -> cvttss2si (truncation; NaN and anything outside the int64 range give 0x8000000000000000);
// otherwise cvttss2si(v - 2^63) ^ 2^63. The comparison is `v >= 2^63` (false for NaN).
AB_HD uint64_t float_to_size_t(float v)
{
    const float two63 = 9223372036854775808.0f;
    if (v >= two63)
    {
        const float w = v - two63;
        const uint64_t r = (w >= two63 || w != w) ? 0x8000000000000000ull : (uint64_t) (int64_t) w;
        return r ^ 0x8000000000000000ull;
    }
    if (v != v || v < -two63)
        return 0x8000000000000000ull;
    return (uint64_t) (int64_t) v;   // truncation toward zero; negative values wrap to huge numbers
}

// std::min(static_cast<size_t>(q), nBins - 1) for the quotient q of getBin
AB_HD int bin_of_quotient(float q)
{
    const uint64_t b = float_to_size_t(q);
    return (int) (b < (uint64_t) (kBins - 1) ? b : (uint64_t) (kBins - 1));
}

// getBin (math_functions.cpp:466-470)
AB_HD int get_bin(float bin_width, float min_value, float value)
{
    if (bin_width == 0)
        return 0;
#ifdef __CUDA_ARCH__
    const float q = __fdiv_rn(__fsub_rn(value, min_value), bin_width);
#else
    const float q = (value - min_value) / bin_width;
#endif
    return bin_of_quotient(q);
}

// Range bookkeeping of one updateTensorHistogram call BEFORE its binning loop (math_functions.cpp:478-548).
// hist/mn/mx/initialized: the TensorProfilingParams; (batch_min, batch_max) from GetMin / GetMax of the tensor.
// Returns false when the call is a no-op (all-zero tensor); otherwise *bin_width_f / *min_f are what the binning uses.
AB_HD bool prepare_update(double* hist, double& mn, double& mx, int& initialized, float batch_min, float batch_max,
                          float* bin_width_f, float* min_f)
{
    double min_input = (double) batch_min;
    double max_input = (double) batch_max;
    if (min_input == 0 && max_input == 0)
        return false;
    if (min_input == max_input)
    {
        const double cand = min_input + (double) 0.01f;
        max_input         = max_input < cand ? cand : max_input;   // std::max(maxInput, minInput + 0.01f)
    }
    if (!initialized)
    {
        for (int i = 0; i < kBins; ++i)
            hist[i] = 0.0;
        mn = min_input, mx = max_input;
        initialized = 1;
    }
    if (min_input < mn || max_input > mx)
    {
        const double new_min = min_input < mn ? min_input : mn;
        const double new_max = mx < max_input ? max_input : mx;
        const double dest_w  = (new_max - new_min) / kBins;
        const double src_w   = (mx - mn) / kBins;
        double scaled[kBins];
        for (int i = 0; i < kBins; ++i)
            scaled[i] = 0.0;
        for (int i = 0; i < kBins; ++i)
        {
            if (hist[i] == 0)
                continue;
            const double src_begin = mn + src_w * i;
            // size_t destBin = (srcBinBegin - newMin) / destBinWidth: a double -> size_t cast of a non-negative quotient
            const uint64_t dest_bin = (uint64_t) ((src_begin - new_min) / dest_w);
            const double dest_end   = new_min + dest_w * (double) (dest_bin + 1);
            double cnt              = round((dest_end - src_begin) / src_w * hist[i]);
            cnt                     = hist[i] < cnt ? hist[i] : cnt;   // std::min(round(..), hist[i])
            scaled[get_bin((float) dest_w, (float) new_min, (float) src_begin)] += cnt;
            if (cnt < hist[i])
                scaled[get_bin((float) dest_w, (float) new_min, (float) (src_begin + dest_w))] += hist[i] - cnt;
        }
        for (int i = 0; i < kBins; ++i)
            hist[i] = scaled[i];
        mn = new_min, mx = new_max;
    }
    *bin_width_f = (float) ((mx - mn) / kBins);
    *min_f       = (float) mn;
    return true;
}

// ---- host only (plain inline functions, never called from a kernel): the closing search -----------------------------------------------------------------------------------
inline double accumulate_in_float(const double* p, size_t n)   // std::accumulate(p, p + n, 0.f)
{
    float acc = 0.f;
    for (size_t i = 0; i < n; ++i)
        acc = (float) ((double) acc + p[i]);
    return (double) acc;
}

// rescaleHistogram (math_functions.cpp:562-640); dst must not alias src
inline void rescale_histogram(const double* src, double src_min, double src_max, double dst_min, double dst_max, double* dst)
{
    if (src_min == dst_min && src_max == dst_max)
    {
        for (int i = 0; i < kBins; ++i)
            dst[i] = src[i];
        return;
    }
    const size_t n      = kBins;
    const double src_w  = (src_max - src_min) / n;
    const double dest_w = (dst_max - dst_min) / n;
    for (int i = 0; i < kBins; ++i)
        dst[i] = 0.0;
    for (size_t s = 0; s < n; ++s)
    {
        const double val = src[s];
        if (val == 0)
            continue;
        const double s_start = src_min + s * src_w;
        const double s_stop  = src_min + (s + 1) * src_w;
        const double f0      = floor((s_start - dst_min) / dest_w);
        const double f1      = ceil((s_stop - dst_min) / dest_w);
        size_t d0            = (size_t) (f0 > 0.0 ? f0 : 0.0);
        size_t d1            = (size_t) (f1 > 0.0 ? f1 : 0.0);
        if (d0 >= n)
            d0 = n - 1;
        if (d1 >= n)
            d1 = n - 1;
        double rem = val;
        for (size_t d = d0; d <= d1; ++d)
        {
            const double d_start = dst_min + d * dest_w;
            const double d_stop  = dst_min + (d + 1) * dest_w;
            const double o_start = s_start < d_start ? d_start : s_start;   // std::max(srcBinStart, destBinStart)
            const double o_stop  = d_stop < s_stop ? d_stop : s_stop;       // std::min(srcBinStop, destBinStop)
            double ratio         = (o_stop - o_start) / src_w;
            ratio                = ratio >= 0.0f ? ratio : 0.0f;
            ratio                = ratio <= 1.0f ? ratio : 1.0f;
            double dist          = round(ratio * val);
            dist                 = dist <= rem ? dist : rem;
            dst[d] += dist;
            rem -= dist;
        }
    }
}

inline void condition_histogram(double* hist, size_t length)   // EntropyEncodingAnalyzer.cpp:151-194
{
    const double eps_zero = 0.0001;
    if (length == 0)
        return;
    size_t zeros = 0;
    for (size_t i = 0; i < length; ++i)
        zeros += (hist[i] == 0.f);
    if (zeros == length)
        return;
    const size_t non_zeros = length - zeros;
    const double eps_non   = eps_zero * (double) zeros / (double) non_zeros;
    if (eps_non >= 1.0)
        return;
    for (size_t i = 0; i < length; ++i)
    {
        const int is_zero = hist[i] == 0.f;   // decided on the ORIGINAL value, as the reference's isZero vector is
        hist[i] += eps_zero * is_zero;
        hist[i] -= eps_non * (1 - is_zero);
    }
}

inline double compute_kl(double* P, double* Q, size_t length)   // :196-219
{
    const double sum_p = accumulate_in_float(P, length);
    const double sum_q = accumulate_in_float(Q, length);
    double divergence  = 0;
    for (size_t i = 0; i < length; ++i)
    {
        P[i] /= sum_p;
        Q[i] /= sum_q;
        if (P[i] > 0 && Q[i] > 0)
            divergence += P[i] * log(P[i] / Q[i]);
    }
    return divergence;
}

// _optimizeKL (:221-428) for DTYPE = float
inline void optimize_kl(const double* histogram, double hist_min, double hist_max, int bw, bool sym, bool strict,
                        bool unsigned_sym, float* o_min, float* o_max)
{
    double hist[kBins];
    if (sym && (hist_min < 0.0 || !unsigned_sym))
    {
        const float abs_max = (float) fmax(fabs(hist_max), fabs(hist_min));   // std::max(std::abs(..), std::abs(..)) -> DTYPE
        const float abs_min = -abs_max;
        rescale_histogram(histogram, hist_min, hist_max, (double) abs_min, (double) abs_max, hist);
        hist_min = abs_min, hist_max = abs_max;
    }
    else
        for (int i = 0; i < kBins; ++i)
            hist[i] = histogram[i];
    const size_t num_bins = kBins, num_q = 255;
    if (bw != 8)
    {
        *o_min = (float) hist_min, *o_max = (float) hist_max;
        return;
    }
    const double bin_w = (hist_max - hist_min) / (double) num_bins;
    double best        = INFINITY;
    double t_min = hist_min, t_max = hist_max;
    size_t start = 0, stop = num_bins - 1;
    double P[kBins], Q[kBins];
    while ((stop - start + 1) >= num_q)
    {
        const size_t win  = stop - start + 1;
        const double* wp  = hist + start;
        for (size_t i = 0; i < win; ++i)
            P[i] = 0.0, Q[i] = 0.0;
        double left = 0;
        for (size_t i = 0; i <= start; ++i)
            left += hist[i];
        P[0] += left;
        for (size_t i = start + 1; i < stop; ++i)
            P[i - start] = hist[i];
        double right = 0;
        for (size_t i = stop; i < num_bins; ++i)
            right += hist[i];
        P[win - 1] += right;
        const double merged = (double) win / (double) num_q;
        for (size_t q = 0; q < num_q; ++q)
        {
            const size_t i0 = (size_t) ceil(q * merged);
            const size_t i1 = (q < num_q - 1) ? (size_t) ceil((q + 1) * merged) : win;
            double sum = 0, norm = 0;
            for (size_t i = i0; i < i1; ++i)
            {
                sum += wp[i];
                norm += (wp[i] != 0);
            }
            if (norm != 0)
                for (size_t i = i0; i < i1; ++i)
                    if (wp[i])
                        Q[i] = sum / norm;
        }
        const double sum_p = accumulate_in_float(P, win);
        const double sum_q = accumulate_in_float(Q, win);
        if (sum_p == 0 || sum_q == 0)
            break;
        condition_histogram(P, win);
        condition_histogram(Q, win);
        const double divergence = compute_kl(P, Q, win);
        if (divergence < best)
        {
            best  = divergence;
            t_min = hist_min + start * bin_w;
            t_max = hist_min + (stop + 1) * bin_w;
        }
        if (sym || strict)
        {
            start++;
            stop--;
        }
        else
        {
            const double loss[3] = {hist[start] + hist[stop], hist[start] + hist[start + 1], hist[stop] + hist[stop - 1]};
            int k = 0;   // std::min_element: the first minimum
            for (int j = 1; j < 3; ++j)
                if (loss[j] < loss[k])
                    k = j;
            if ((k == 0 && (hist_min + (start + 1) * bin_w) > 0) || (k == 1 && (hist_min + (start + 2) * bin_w) > 0))
                k = 2;
            else if ((k == 0 && (hist_min + stop * bin_w) < 0) || (k == 2 && (hist_min + (stop - 1) * bin_w) < 0))
                k = 1;
            if (k == 0)
                start++, stop--;
            else if (k == 1)
                start += 2;
            else
                stop -= 2;
        }
    }
    *o_min = (float) t_min, *o_max = (float) t_max;
}

// computeEncoding (:98-143). `initialized`: the histogram exists; `stats_updated`: updateStats was called at all.
inline void compute_encoding(const double* histogram, double hist_min, double hist_max, int initialized, int stats_updated,
                             int bw, bool sym, bool strict, bool unsigned_sym, ab_encoding& e)
{
    e.min = e.max = e.delta = e.offset = 0.0;
    e.bw = 0;
    float num_steps = (float) (pow(2, bw) - 1);
    if (sym && strict)
        num_steps -= 1;
    if (!initialized)
    {
        if (stats_updated)   // only all-zero tensors so far: a valid encoding that covers 0
        {
            e.min    = -1;
            e.max    = 1;
            e.delta  = (e.max - e.min) / (int) num_steps;
            e.offset = floor(e.min / e.delta);
            e.min    = e.offset * e.delta;
            e.max    = e.min + (int) num_steps * e.delta;
            e.bw     = bw;
        }
        return;
    }
    float a_min, a_max;
    optimize_kl(histogram, hist_min, hist_max, bw, sym, strict, unsigned_sym, &a_min, &a_max);
    a_min = (0.f < a_min) ? 0.f : a_min;   // std::min(aMin, 0.f)
    a_max = (a_max < 0.f) ? 0.f : a_max;   // std::max(aMax, 0.f)
    em::tf_encoding(bw, (double) a_min, (double) a_max, sym, strict, unsigned_sym, e);
}
}   // namespace ent
}   // namespace ab
