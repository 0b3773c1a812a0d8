// TEST-ONLY: compiles the product's host+device math headers (aimet_b200/csrc/{encoding_math,tfe_math}.h) for the
// HOST, so that the exact code the CUDA kernels run can be checked against the oracle on a machine with no GPU.
// This library is never loaded by the aimet_b200 package.
#include <cstring>

#include "../../aimet_b200/csrc/mse_math.h"
#include "../../aimet_b200/csrc/percentile_math.h"
#include "../../aimet_b200/csrc/tfe_math.h"

using namespace ab;

extern "C"
{
// mirrors compute_encodings_kernel (tfe_search.cu) with the candidates visited serially
void ht_tfe_compute(const double* pdf, double x_left0, double bucket_d, int initialized, int stats_updated, int bw,
                    int sym, int strict, int unsigned_sym, double* out5)
{
    ab_encoding e = {0, 0, 0, 0, 0};
    if (!initialized)
    {
        if (stats_updated)
            tfe::all_zero_encoding(bw, e);
    }
    else
    {
        tfe::PdfView view {pdf, x_left0, bucket_d};
        float mn, mx;
        tfe::find_range(view, mn, mx);
        const float steps = tfe::num_steps_for(bw, sym, strict);
        double best_cost  = INFINITY;
        int best_idx      = INT_MAX;
        float best_delta = -1.0f;
        int best_offset  = -1;
        float sym_deltas[tfe::kMaxSymDeltas];
        int sym_offset = 0, n = tfe::kAsymCandidates;
        tfe::AsymSetup as {};
        if (sym)
            n = tfe::sym_candidates(mn, mx, steps, unsigned_sym, sym_deltas, sym_offset);
        else
            as = tfe::asym_setup(mn, mx, steps);
        for (int k = 0; k < n; ++k)
        {
            float d;
            int o;
            bool valid = true;
            if (sym)
                d = sym_deltas[k], o = sym_offset;
            else
                valid = tfe::asym_candidate(as, k, d, o);
            if (!valid)
                continue;
            const double c = tfe::cost(view, bw, d, o);
            if (c < DBL_MAX && (c < best_cost || (c == best_cost && k < best_idx)))
                best_cost = c, best_idx = k, best_delta = d, best_offset = o;
        }
        tfe::finish(best_delta, best_offset, steps, bw, e);
    }
    out5[0] = e.min, out5[1] = e.max, out5[2] = e.delta, out5[3] = e.offset, out5[4] = e.bw;
}

// the percentile and MSE modes of compute_encodings_kernel on an initialised PDF
void ht_percentile_compute(const double* pdf, double x_left0, double bucket_d, float percentile, int bw, int sym,
                           int strict, int unsigned_sym, double* out5)
{
    tfe::PdfView view {pdf, x_left0, bucket_d};
    double cdf[AB_PDF_SIZE];
    ab_encoding e;
    pct::encoding(view, cdf, percentile, bw, sym, strict, unsigned_sym, e);
    out5[0] = e.min, out5[1] = e.max, out5[2] = e.delta, out5[3] = e.offset, out5[4] = e.bw;
}

void ht_mse_compute(const double* pdf, double x_left0, double bucket_d, int bw, int sym, int strict, int unsigned_sym,
                    double* out5)
{
    tfe::PdfView view {pdf, x_left0, bucket_d};
    static mse::Tables tables;
    ab_encoding e;
    mse::encoding(view, tables, bw, sym, strict, unsigned_sym, e);
    out5[0] = e.min, out5[1] = e.max, out5[2] = e.delta, out5[3] = e.offset, out5[4] = e.bw;
}

void ht_init_pdf_range(float mn, float mx, double* x_left0, double* bucket_d, float* bucket_f, float* offset_f)
{
    em::init_pdf_range(mn, mx, *x_left0, *bucket_d, *bucket_f, *offset_f);
}

double ht_x_left(double x_left0, double bucket_d, int i)
{
    return em::x_left_at(x_left0, bucket_d, i);
}

void ht_fill_encoding_info(int bw, double mn, double mx, double* out5)
{
    ab_encoding e;
    em::fill_encoding_info(bw, mn, mx, e);
    out5[0] = e.min, out5[1] = e.max, out5[2] = e.delta, out5[3] = e.offset, out5[4] = e.bw;
}
}
