// Job 3 -- the tf_enhanced SQNR grid search on the device, batched over quantizers.
//
// Reference: TfEnhancedEncodingAnalyzer<float>::computeEncoding (DlQ/src/TfEnhancedEncodingAnalyzer.cpp:79-397): a
// serial loop over <= 358 (asymmetric) or ~101 (symmetric) (delta, offset) candidates, each costing a pass over the
// 512-bin PDF; ~0.6 ms per quantizer on a CPU core, x 26 560 weight channels for per-channel ResNet-50.
//
// Here: one CTA per quantizer, one thread per candidate. The PDF (4 KB of doubles) is staged in shared memory; every
// thread evaluates its candidate's cost with the reference's exact mixed float/double arithmetic and its exact bin
// order (tfe_math.h), so each cost is bit-identical to the CPU value. The winner is the lowest-index strict minimum
// (`cost < bestCost`, :138), found with a (cost, index) block reduction -- NOT a reduction over bins, which would
// change rounding. The epilogue also emits the fp32 kernel parameters fillEncodingInfo would derive, so a per-tensor
// QDQ can follow on the same stream without a host round trip.
#include "common.cuh"
#include "mse_math.h"
#include "percentile_math.h"
#include "tfe_math.h"

namespace ab
{
namespace
{

constexpr int kSearchThreads = 384;   // >= 358 candidates

struct SearchArgs
{
    int quant_mode, bw, sym, strict, unsigned_sym;
    float percentile;   // AB_QUANTIZATION_PERCENTILE only
};

__device__ __forceinline__ void write_encoding(double* enc_out, float* qdq4_out, int64_t s, const ab_encoding& e,
                                               bool valid)
{
    double* o = enc_out + s * 5;
    o[0] = e.min, o[1] = e.max, o[2] = e.delta, o[3] = e.offset, o[4] = (double) e.bw;
    if (qdq4_out)
    {
        float4 p = make_float4(0.f, 0.f, 1.f, 0.f);
        if (valid)
        {
            ab_encoding full;
            em::fill_encoding_info(e.bw, e.min, e.max, full);
            p = make_float4((float) full.min, (float) full.max, (float) full.delta, (float) full.offset);
        }
        reinterpret_cast<float4*>(qdq4_out)[s] = p;
    }
}

// ---- the tf_enhanced cost on a compacted bin list ---------------------------------------------------------------------
// tfe::cost (tfe_math.h) walks all 512 bins per candidate and re-derives every bin's mid-point, its float image and the
// three range tests for each of the <= 358 candidates of the CTA. Nothing of that depends on the candidate: the CTA builds,
// once per quantizer, the ascending list of the bins tfe::cost would not skip (probability, mid-point as double and as
// float) and the rank of every bin in that list; a candidate then runs three branch-free loops over list ranges -- bottom
// saturation [0, rank(min_ind)), quantisation [rank(min_ind), rank(max_ind)), top saturation [rank(max_ind), n) -- each
// adding the same terms in the same (ascending) order to its own accumulator as the one-pass form does, so every sum is
// bit-identical. In the quantisation loop the division by the candidate's delta is the hoisted-reciprocal exact division of
// common.cuh, C round() is round_half_away_small, and (float) (int(q) + offset) is q + offset in fp32 (both integers below
// 2^22 in magnitude, checked per candidate; otherwise the candidate runs the literal form on the same list). No F2I, I2F,
// MUFU or FRND in the loop. ResNet-50's 26 560 weight channels (symmetric, 101 candidates): 1.58 -> 0.63 ms; asymmetric (358): 4.55 -> 1.86 ms.
struct BinList
{
    double mid[AB_PDF_SIZE];          // pdf_start + i * pdf_step + pdf_step / 2 of listed bin j
    double pr[AB_PDF_SIZE];           // its probability
    float fv[AB_PDF_SIZE];            // (float) mid
    uint16_t rank[AB_PDF_SIZE + 2];   // rank[i] = number of listed bins with index < i; rank[512] = n
};

// Called by all threads of the CTA. Lists every bin with pr != 0.0 (what tfe::cost visits when its skip_empty holds).
__device__ __forceinline__ void build_bin_list(const double* __restrict__ s_pdf, float pdf_start, double pdf_step,
                                               BinList& L, int* s_wcount, int& s_base)
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, n_warps = (int) blockDim.x >> 5;
    if (tid == 0)
        s_base = 0;
    __syncthreads();
    for (int c0 = 0; c0 < AB_PDF_SIZE; c0 += (int) blockDim.x)
    {
        const int i       = c0 + tid;
        const bool in     = i < AB_PDF_SIZE;
        const double pr   = in ? s_pdf[i] : 0.0;
        const bool keep   = in && !(pr == 0.0);
        const unsigned b  = __ballot_sync(0xffffffffu, keep);
        if (lane == 0)
            s_wcount[warp] = __popc(b);
        __syncthreads();
        int pos = s_base;
        for (int w = 0; w < warp; ++w)
            pos += s_wcount[w];
        pos += __popc(b & ((1u << lane) - 1u));
        if (in)
        {
            L.rank[i] = (uint16_t) pos;
            if (keep)
            {
                const double mid = pdf_start + i * pdf_step + pdf_step / 2;   // the expression of tfe::cost
                L.mid[pos]       = mid;
                L.pr[pos]        = pr;
                L.fv[pos]        = (float) mid;
            }
        }
        __syncthreads();
        if (tid == 0)
        {
            int t = s_base;
            for (int w = 0; w < n_warps; ++w)
                t += s_wcount[w];
            s_base = t;
        }
        __syncthreads();
    }
    if (tid == 0)
        L.rank[AB_PDF_SIZE] = (uint16_t) s_base;
    __syncthreads();
}

// tfe::cost for a finite geometry and a finite, non-zero delta (its skip_empty case), on the list
__device__ __forceinline__ double cost_on_list(const BinList& L, float pdf_start, double pdf_step, int bw, float delta,
                                               int offset)
{
    using namespace em;
    const float min_val   = delta * offset;
    const float step_size = (float) (pow2(bw) - 1);
    const float max_val   = delta * (offset + step_size);
    int min_ind           = d2i_x86(floor((min_val - pdf_start) / pdf_step));
    min_ind               = smin(smax(0, min_ind), AB_PDF_SIZE - 1);
    int max_ind           = d2i_x86(floor((max_val - pdf_start) / pdf_step));
    max_ind               = smin(smax(0, max_ind), AB_PDF_SIZE - 1);
    const double min_mid  = (double) (float) (pdf_start + (min_ind * pdf_step) + pdf_step / 2);
    const double max_mid  = (double) (float) (pdf_start + (max_ind * pdf_step) + pdf_step / 2);
    const float offset_f  = (float) offset;
    const int n = L.rank[AB_PDF_SIZE], lo = L.rank[min_ind], hi = L.rank[max_ind];

    double sat_bottom = 0, sat_top = 0, quant = 0;
    for (int j = 0; j < lo; ++j)
    {
        const double d = L.mid[j] - min_mid;
        sat_bottom += L.pr[j] * (d * d);
    }
    for (int j = hi; j < n; ++j)
    {
        const double d = L.mid[j] - max_mid;
        sat_top += L.pr[j] * (d * d);
    }
    if (lo < hi)
    {
        // |fv / delta - offset| over the range is bounded by its two ends (fv ascends); the hoisted reciprocal is within a
        // few ulp of 1 / delta, the margin below 2^22 covers it
        const Divisor dv  = make_divisor(delta);
        const float bound = __fadd_rn(__fmul_rn(fmaxf(fabsf(L.fv[lo]), fabsf(L.fv[hi - 1])), fabsf(dv.y)), fabsf(offset_f));
        if (dv.fast && bound < 4194000.0f)
        {
            for (int j = lo; j < hi; ++j)
            {
                const float fv  = L.fv[j];
                const float q   = round_half_away_small(__fsub_rn(div_fast(fv, dv), offset_f));
                const float deq = __fmul_rn(delta, __fadd_rn(q, offset_f));   // exact integer sum below 2^23
                const double d  = (double) __fsub_rn(fv, deq);
                quant += L.pr[j] * (d * d);
            }
        }
        else
        {
            for (int j = lo; j < hi; ++j)
            {
                const float fv          = L.fv[j];
                const int quantized     = f2i_x86(roundf_away(fv / delta - offset_f));
                const float dequantized = delta * (float) iadd_wrap(quantized, offset);
                const double d          = (double) (fv - dequantized);
                quant += L.pr[j] * (d * d);
            }
        }
    }
    const double sqnr = tfe::kGamma * (sat_bottom + sat_top) + quant;
    return smin(sqnr, DBL_MAX);
}

__global__ void __launch_bounds__(kSearchThreads)
    compute_encodings_kernel(const ab_stats_state* __restrict__ states, int64_t count, SearchArgs a,
                             double* __restrict__ enc_out, float* __restrict__ qdq4_out)
{
    __shared__ double s_pdf[AB_PDF_SIZE];
    // one scheme per launch: the percentile scan, the MSE tables and the tf_enhanced bin list share their storage
    union ModeScratch
    {
        double cdf[AB_PDF_SIZE];
        mse::Tables mse;
        BinList bins;
    };
    __shared__ __align__(16) unsigned char s_scratch_raw[sizeof(ModeScratch)];
    ModeScratch& s_scratch = *reinterpret_cast<ModeScratch*>(s_scratch_raw);
    double* const s_cdf    = s_scratch.cdf;
    mse::Tables& s_mse     = s_scratch.mse;
    BinList& s_bins        = s_scratch.bins;
    __shared__ int s_wcount[kSearchThreads / 32];
    __shared__ int s_list_base;
    __shared__ float s_sym_deltas[tfe::kMaxSymDeltas];
    __shared__ tfe::AsymSetup s_asym;
    __shared__ float s_num_steps;
    __shared__ int s_sym_offset, s_num_cand;
    __shared__ double s_best_cost[kSearchThreads / 32];
    __shared__ int s_best_idx[kSearchThreads / 32];

    const int tid = threadIdx.x;
    for (int64_t s = blockIdx.x; s < count; s += gridDim.x)
    {
        const ab_stats_state* st = states + s;
        __syncthreads();   // previous quantizer is completely done with shared memory

        if (a.quant_mode == AB_QUANTIZATION_TF)
        {
            if (tid == 0)
            {
                ab_encoding e = {0, 0, 0, 0, 0};
                const bool ok = st->stats_updated != 0;
                if (ok)
                    em::tf_analyzer_encoding(a.bw, st->run_min, st->run_max, a.sym != 0, a.strict != 0,
                                             a.unsigned_sym != 0, e);
                write_encoding(enc_out, qdq4_out, s, e, ok);
            }
            continue;
        }
        if (!st->initialized)
        {
            if (tid == 0)
            {
                ab_encoding e = {0, 0, 0, 0, 0};
                const bool ok = st->stats_updated != 0;
                if (ok)
                {
                    // only zeros seen so far (TfEnhancedEncodingAnalyzer.cpp:85-100, PercentileEncodingAnalyzer.cpp:91-105)
                    if (a.quant_mode == AB_QUANTIZATION_PERCENTILE || a.quant_mode == AB_QUANTIZATION_MSE)
                        pct::all_zero_encoding(a.bw, a.sym != 0 && a.strict != 0, e);   // MseEncodingAnalyzer.cpp:91-105: same
                    else
                        tfe::all_zero_encoding(a.bw, e);
                }
                write_encoding(enc_out, qdq4_out, s, e, ok);
            }
            continue;
        }

        {
            // a batch still parked in hist[write_parity ^ 1] (ab_stats_state.pending) is folded on the fly:
            // pdf = (pdf * k + hist / cnt) / (k + 1), the reference's running mean (DlQ/src/math_functions.cpp:279-287)
            const bool pending = st->pending != 0;
            const int pp       = st->write_parity ^ 1;
            const int k        = st->iterations;
            const double cnt   = st->pending_count;
            for (int i = tid; i < AB_PDF_SIZE; i += blockDim.x)
            {
                double p = st->pdf[i];
                if (pending)
                    p = __ddiv_rn(__dadd_rn(__dmul_rn(p, (double) k), (double) st->hist[pp][i] / cnt), (double) (k + 1));
                s_pdf[i] = p;
            }
        }
        __syncthreads();
        const tfe::PdfView view {s_pdf, st->x_left0, st->bucket_size_d};

        if (a.quant_mode == AB_QUANTIZATION_PERCENTILE)
        {
            // The cumulative sum is sequential in the reference (cdf[i] += cdf[i - 1]) and so it is here: one thread,
            // 512 dependent double additions, ~2 us -- with one CTA per quantizer that is still thousands of
            // quantizers per millisecond.
            if (tid == 0)
            {
                ab_encoding e;
                pct::encoding(view, s_cdf, a.percentile, a.bw, a.sym != 0, a.strict != 0, a.unsigned_sym != 0, e);
                write_encoding(enc_out, qdq4_out, s, e, true);
            }
            continue;
        }

        if (a.quant_mode == AB_QUANTIZATION_MSE)
        {
            // thread 0 lays out bin edges / centres (float accumulations, inherently sequential), then the (min, max)
            // candidates -- up to 257 x 257 of them -- are spread over the CTA; (cost, index) arg-min = the reference's
            // first strict minimum
            if (tid == 0)
                mse::build_tables(view, s_mse);
            __syncthreads();
            const int n_cand = mse::num_candidates(s_mse);
            double bc = INFINITY;
            int bi    = INT_MAX;
            for (int k = tid; k < n_cand; k += blockDim.x)
            {
                float cmin, cmax;
                mse::candidate(s_mse, k, cmin, cmax);
                const float c = mse::cost(s_mse, a.bw, cmin, cmax, a.sym != 0, a.strict != 0, a.unsigned_sym != 0);
                if (c < FLT_MAX && (double) c < bc)   // `mse < mseMin` with mseMin starting at FLT_MAX; ascending k per thread
                    bc = c, bi = k;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1)
            {
                const double oc = __shfl_xor_sync(0xffffffffu, bc, o);
                const int oi    = __shfl_xor_sync(0xffffffffu, bi, o);
                if (oc < bc || (oc == bc && oi < bi))
                    bc = oc, bi = oi;
            }
            if ((tid & 31) == 0)
                s_best_cost[tid >> 5] = bc, s_best_idx[tid >> 5] = bi;
            __syncthreads();
            if (tid == 0)
            {
                for (int w = 1; w < (int) blockDim.x / 32; ++w)
                    if (s_best_cost[w] < bc || (s_best_cost[w] == bc && s_best_idx[w] < bi))
                        bc = s_best_cost[w], bi = s_best_idx[w];
                float best_min = s_mse.min_val, best_max = s_mse.max_val;
                if (bi != INT_MAX)
                    mse::candidate(s_mse, bi, best_min, best_max);
                ab_encoding e;
                mse::finish(a.bw, best_min, best_max, a.sym != 0, a.strict != 0, a.unsigned_sym != 0, e);
                write_encoding(enc_out, qdq4_out, s, e, true);
            }
            continue;
        }

        if (tid == 0)
        {
            float min_val, max_val;
            tfe::find_range(view, min_val, max_val);
            const float steps = tfe::num_steps_for(a.bw, a.sym != 0, a.strict != 0);
            s_num_steps       = steps;
            if (a.sym)
            {
                int off;
                s_num_cand   = tfe::sym_candidates(min_val, max_val, steps, a.unsigned_sym != 0, s_sym_deltas, off);
                s_sym_offset = off;
            }
            else
            {
                s_asym     = tfe::asym_setup(min_val, max_val, steps);
                s_num_cand = tfe::kAsymCandidates;
            }
        }
        // finite histogram geometry (always, unless the range was fixed from +-inf inputs): candidates with a finite,
        // non-zero delta run on the compacted list; everything else takes the literal one-pass form
        const float pdf_start = (float) view.x_left(0);
        const double pdf_step = view.x_left(1) - view.x_left(0);
        const bool geom_ok    = (pdf_start - pdf_start == 0.0f) && (pdf_step - pdf_step == 0.0);
        if (geom_ok)
            build_bin_list(s_pdf, pdf_start, pdf_step, s_bins, s_wcount, s_list_base);
        __syncthreads();

        double my_cost = INFINITY;
        int my_idx     = INT_MAX;
        float my_delta = -1.0f;
        int my_offset  = -1;
        if (tid < s_num_cand)
        {
            bool valid;
            if (a.sym)
            {
                my_delta  = s_sym_deltas[tid];
                my_offset = s_sym_offset;
                valid     = true;
            }
            else
                valid = tfe::asym_candidate(s_asym, tid, my_delta, my_offset);
            if (valid)
            {
                const bool on_list = geom_ok && (my_delta - my_delta == 0.0f) && my_delta != 0.0f;
                const double c     = on_list ? cost_on_list(s_bins, pdf_start, pdf_step, a.bw, my_delta, my_offset)
                                             : tfe::cost(view, a.bw, my_delta, my_offset);
                if (c < DBL_MAX)   // `cost < bestCost` with bestCost starting at DBL_MAX (:126,138); NaN never wins
                {
                    my_cost = c;
                    my_idx  = tid;
                }
            }
        }
        // argmin over (cost, index): lowest cost, ties -> lowest index == the first strict minimum in push order
        double bc = my_cost;
        int bi    = my_idx;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
        {
            const double oc = __shfl_xor_sync(0xffffffffu, bc, o);
            const int oi    = __shfl_xor_sync(0xffffffffu, bi, o);
            if (oc < bc || (oc == bc && oi < bi))
                bc = oc, bi = oi;
        }
        if ((tid & 31) == 0)
            s_best_cost[tid >> 5] = bc, s_best_idx[tid >> 5] = bi;
        __syncthreads();
        if (tid == 0)
        {
            for (int w = 1; w < (int) blockDim.x / 32; ++w)
                if (s_best_cost[w] < bc || (s_best_cost[w] == bc && s_best_idx[w] < bi))
                    bc = s_best_cost[w], bi = s_best_idx[w];
            s_best_idx[0] = bi;
        }
        __syncthreads();
        const int winner = s_best_idx[0];
        if (winner == INT_MAX)
        {
            if (tid == 0)   // no candidate beat DBL_MAX: the reference keeps bestDelta = bestOffset = -1 (:121-122)
            {
                ab_encoding e;
                tfe::finish(-1.0f, -1, s_num_steps, a.bw, e);
                write_encoding(enc_out, qdq4_out, s, e, true);
            }
        }
        else if (tid == winner)
        {
            ab_encoding e;
            tfe::finish(my_delta, my_offset, s_num_steps, a.bw, e);
            write_encoding(enc_out, qdq4_out, s, e, true);
        }
    }
}

}   // namespace
}   // namespace ab

using namespace ab;

namespace
{
int launch_search(const ab_stats_state* states, int64_t count, int quant_mode, float percentile, int bw, int use_symmetric,
                  int use_strict_symmetric, int use_unsigned_symmetric, double* enc_out, float* qdq4_out, void* stream)
{
    if (count < 0 || (count > 0 && (states == nullptr || enc_out == nullptr)))
    {
        set_error("null pointer or negative count");
        return AB_ERR_INVALID;
    }
    if (quant_mode != AB_QUANTIZATION_TF && quant_mode != AB_QUANTIZATION_TF_ENHANCED &&
        quant_mode != AB_QUANTIZATION_PERCENTILE && quant_mode != AB_QUANTIZATION_MSE)
    {
        set_error("unsupported quantization mode %d", quant_mode);
        return AB_ERR_INVALID;
    }
    if (bw < 1 || bw > 32)
    {
        set_error("bitwidth %d out of range", bw);
        return AB_ERR_INVALID;
    }
    if (qdq4_out != nullptr && (reinterpret_cast<uintptr_t>(qdq4_out) & 15u) != 0)
    {
        set_error("qdq4_out must be 16-byte aligned");
        return AB_ERR_INVALID;
    }
    if (count == 0)
        return AB_OK;
    SearchArgs a {quant_mode, bw, use_symmetric, use_strict_symmetric, use_unsigned_symmetric, percentile};
    // one thread per candidate: 358 asymmetric, ~101 symmetric (tfe_math.h) -- the symmetric search (all weights) runs
    // with a third of the threads, i.e. three times as many quantizers resident per SM
    const int threads = (quant_mode == AB_QUANTIZATION_TF_ENHANCED && use_symmetric) ? 128 : kSearchThreads;
    int per_sm        = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, compute_encodings_kernel, threads, 0) != cudaSuccess ||
        per_sm <= 0)
        per_sm = 1;
    int64_t grid = (int64_t) per_sm * num_sms();
    if (count < grid)
        grid = count;
    compute_encodings_kernel<<<(unsigned) grid, threads, 0, (cudaStream_t) stream>>>(states, count, a, enc_out, qdq4_out);
    AB_CUDA_CHECK(cudaGetLastError());
    return AB_OK;
}
}   // namespace

extern "C" int ab_compute_encodings(const ab_stats_state* states, int64_t count, int quant_mode, int bw,
                                    int use_symmetric, int use_strict_symmetric, int use_unsigned_symmetric,
                                    double* enc_out, float* qdq4_out, void* stream)
{
    if (quant_mode == AB_QUANTIZATION_PERCENTILE)
    {
        set_error("the percentile scheme needs its percentile: call ab_compute_encodings_percentile");
        return AB_ERR_INVALID;
    }
    return launch_search(states, count, quant_mode, 100.0f, bw, use_symmetric, use_strict_symmetric,
                         use_unsigned_symmetric, enc_out, qdq4_out, stream);
}

extern "C" int ab_compute_encodings_percentile(const ab_stats_state* states, int64_t count, float percentile, int bw,
                                               int use_symmetric, int use_strict_symmetric, int use_unsigned_symmetric,
                                               double* enc_out, float* qdq4_out, void* stream)
{
    return launch_search(states, count, AB_QUANTIZATION_PERCENTILE, percentile, bw, use_symmetric, use_strict_symmetric,
                         use_unsigned_symmetric, enc_out, qdq4_out, stream);
}

// reset -> updateStats -> computeEncoding (-> per-channel parameter block) for one tensor, enqueued by ONE host call.
// A parameter quantizer in training mode does exactly this sequence before every forward (TEt/.../v1/qc_quantize_op.py:
// 753-798); issuing it as four separate calls from Python costs more host time than the kernels take on the device.
extern "C" int ab_stats_refresh_encodings(const void* in, int64_t num_segments, int64_t segment_len, int dtype,
                                          int quant_mode, ab_stats_state* states, int bw, int use_symmetric,
                                          int use_strict_symmetric, int use_unsigned_symmetric, double* enc_out,
                                          float* qdq4_out, float* params_out, void* stream)
{
    if (num_segments < 1 || segment_len < 0 || enc_out == nullptr)
    {
        ab::set_error("invalid refresh arguments");
        return AB_ERR_INVALID;
    }
    int rc = ab_stats_reset(states, num_segments, stream);
    if (rc != AB_OK)
        return rc;
    rc = num_segments == 1 ? ab_stats_update(in, segment_len, dtype, quant_mode, states, nullptr, 0, stream)
                           : ab_stats_update_segmented(in, num_segments, segment_len, dtype, quant_mode, states, stream);
    if (rc != AB_OK)
        return rc;
    rc = ab_compute_encodings(states, num_segments, quant_mode, bw, use_symmetric, use_strict_symmetric,
                              use_unsigned_symmetric, enc_out, qdq4_out, stream);
    if (rc != AB_OK || params_out == nullptr)
        return rc;
    return ab_per_channel_params_dev(enc_out, num_segments, bw, params_out, stream);
}
