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

__global__ void __launch_bounds__(kSearchThreads)
    compute_encodings_kernel(const ab_stats_state* __restrict__ states, int64_t count, SearchArgs a,
                             double* __restrict__ enc_out, float* __restrict__ qdq4_out)
{
    __shared__ double s_pdf[AB_PDF_SIZE];
    __shared__ double s_cdf[AB_PDF_SIZE];
    __shared__ mse::Tables s_mse;
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
                const double c = tfe::cost(view, a.bw, my_delta, my_offset);
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
