// Job 2 -- statistics kernels for sm_100a: min/max and the 512-bin histogram, accumulated into device-resident
// per-quantizer state (ab_stats_state) with no host synchronisation.
//
// Reference semantics (CPU path = parity target): DlQ/src/math_functions.cpp:207-288 (InitializePdf, UpdatePdf),
// :327-347 (GetMin/GetMax), :367-384 (GetHistogram_cpu); DlQ/src/TfEncodingAnalyzer.cpp:60-71.
// The reference's own GPU path (math_functions.cu:52-64,125-211) does two thrust reductions with host round trips,
// per-thread global-memory histograms on at most 32 blocks, a single-block reduce and a blocking copy per call.
//
// Design here
//   minmax_kernel   : one streaming pass, 128-bit loads, per-thread fmin/fmax, warp-shuffle + block reduce, one
//                     atomicMin/atomicMax per CTA on an order-preserving integer image of the float; the last CTA
//                     (ticket) folds the batch result into the state. For tf_enhanced it exits immediately once the
//                     histogram range is fixed, so the steady state is ONE pass over the data.
//   hist_kernel     : persistent, one CTA per SM. Tiles of 32 KB are staged into a 4-deep shared-memory ring by the
//                     TMA engine (cp.async.bulk + mbarrier complete_tx), walking the tensor from its end backwards (what
//                     the producing kernel wrote last is what the L2 still holds), so no registers or LSU issue slots
//                     are spent on global loads. Bins are counted in a shared-memory histogram privatised PER LANE
//                     (bin b of lane l lives at word b*32+l: bank == lane, so a warp's 32 atomics never conflict,
//                     whatever the data distribution -- ReLU outputs put half the samples in one bin). The CTA then
//                     reduces its 32 copies and flushes with at most 512 global atomics. The batch's counts stay parked
//                     in the record; the NEXT launch's keeper warps fold them into the running PDF in double precision,
//                     exactly as UpdatePdf does, while that launch streams. Nobody waits at the end of a launch: the
//                     record is read once per CTA into a shared-memory snapshot, announced with an early ticket, and the
//                     CTA that drew the last ticket writes the bookkeeping (see `fast_tail`). bf16 tensors bin with a
//                     one-FFMA index certified against the reference sequence over all 65536 bit patterns.
//   segmented kernel: one CTA per segment (= per output channel of a weight): min/max, range, histogram and PDF fold
//                     for thousands of quantizers in ONE launch, replacing a Python loop of per-channel calls.
#include "common.cuh"
#include "encoding_math.h"

namespace ab
{
namespace
{

constexpr int kBins = AB_PDF_SIZE;

// ---- small helpers --------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_min(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
        v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_max(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
        v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// GetHistogram_cpu's bin index (DlQ/src/math_functions.cpp:375-376): round(x / bucket - offset) in float, then the
// x86 float->int conversion (NaN / out of range -> INT_MIN, i.e. dropped). Returns -1 for "not counted".
__device__ __forceinline__ int bin_index(float x, float bucket, float offset)
{
    const float r = round_half_away(__fsub_rn(__fdiv_rn(x, bucket), offset));
    // r is an integer-valued float (or NaN / inf). In-range test done in float so NaN and inf fall out.
    return (r >= 0.0f && r < (float) kBins) ? (int) r : -1;
}

struct Range
{
    float bucket, offset;
    bool valid;   // false: uninitialised PDF and an all-zero batch -> nothing to count
};

// Range for this call: the frozen one, or (first non-zero batch) the one InitializePdf derives from the batch min/max.
template <typename S>
__device__ __forceinline__ Range resolve_range(const S* st, double* x_left0, double* bucket_d)
{
    Range r;
    if (st->initialized)
    {
        r.bucket = st->bucket_size;
        r.offset = st->pdf_offset;
        r.valid  = true;
        return r;
    }
    const float mn = ordered_to_float(st->batch_min_bits);
    const float mx = ordered_to_float(st->batch_max_bits);
    if (mn == 0 && mx == 0)   // DlQ/src/math_functions.cpp:254-259
    {
        r.bucket = 1.0f, r.offset = 0.0f, r.valid = false;
        return r;
    }
    double x0, bd;
    em::init_pdf_range(mn, mx, x0, bd, r.bucket, r.offset);
    if (x_left0)
        *x_left0 = x0, *bucket_d = bd;
    r.valid = true;
    return r;
}

// Fold one batch's counts into the running PDF (DlQ/src/math_functions.cpp:279-287). One thread per bin.
__device__ __forceinline__ void fold_bin(double* pdf, uint32_t h, double cnt, int iterations)
{
    const double prob = (double) h / cnt;
    *pdf              = __ddiv_rn(__dadd_rn(__dmul_rn(*pdf, (double) iterations), prob), (double) (iterations + 1));
}

// Fold the batch parked in hist[write_parity ^ 1] (see ab_stats_state.pending). `lane` / `lanes`: the calling threads
// split the 512 bins among themselves; the caller orders this before any later write of `pending` / `iterations`.
__device__ __forceinline__ void fold_pending(ab_stats_state* st, int lane, int lanes)
{
    const int pp         = st->write_parity ^ 1;
    const int iterations = st->iterations;
    const double cnt     = st->pending_count;
    for (int b = lane; b < kBins; b += lanes)
    {
        fold_bin(&st->pdf[b], st->hist[pp][b], cnt, iterations);
        st->hist[pp][b] = 0;
    }
}

constexpr int32_t kPosInfBits = 0x7f800000;                     // ordered image of +inf
constexpr int32_t kNegInfBits = (int32_t) 0xff800000 ^ 0x7fffffff;   // ordered image of -inf

// ---------------------------------------------------------------------------------------------------------------
// reset
// ---------------------------------------------------------------------------------------------------------------
__global__ void reset_kernel(ab_stats_state* states, int64_t count)
{
    const int64_t s = blockIdx.x;
    if (s >= count)
        return;
    ab_stats_state* st = states + s;
    for (int i = threadIdx.x; i < kBins; i += blockDim.x)
    {
        st->pdf[i]     = 0.0;
        st->hist[0][i] = 0;
        st->hist[1][i] = 0;
    }
    if (threadIdx.x == 0)
    {
        st->pending_count  = 0.0;
        st->pending        = 0;
        st->write_parity   = 0;
        st->x_left0        = 0.0;
        st->bucket_size_d  = 0.0;
        st->run_min        = DBL_MAX;    // DlQ/src/TfEncodingAnalyzer.h:88-91
        st->run_max        = -DBL_MAX;
        st->bucket_size    = 0.0f;
        st->pdf_offset     = 0.0f;
        st->batch_min_bits = kPosInfBits;
        st->batch_max_bits = kNegInfBits;
        st->initialized    = 0;
        st->stats_updated  = 0;
        st->iterations     = 0;
        st->ticket         = 0;
        st->bf16_scale     = 0.0f;
        st->bf16_shift     = 0.0f;
        st->bf16_formula   = 0;
        st->bf16_fail_mask = 0;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// min / max pass
// ---------------------------------------------------------------------------------------------------------------
constexpr int kMmThreads = 256;
constexpr int kMmUnroll  = 4;

template <typename T>
__device__ __forceinline__ void thread_minmax(const T* __restrict__ in, int64_t count, int64_t cta, int64_t num_cta,
                                              float& lo, float& hi)
{
    constexpr int kV = Elem<T>::kPerVec;
    lo               = INFINITY;    // float(+DBL_MAX) == +inf (DlQ/src/math_functions.cpp:341)
    hi               = -INFINITY;
    if ((reinterpret_cast<uintptr_t>(in) & 15u) == 0)
    {
        const int64_t num_vec   = count / kV;
        const int64_t per_tile  = (int64_t) kMmThreads * kMmUnroll;
        const int64_t num_tiles = (num_vec + per_tile - 1) / per_tile;
        for (int64_t tile = cta; tile < num_tiles; tile += num_cta)
        {
            const int64_t v0 = tile * per_tile + threadIdx.x;
            uint4 raw[kMmUnroll];
#pragma unroll
            for (int u = 0; u < kMmUnroll; ++u)
            {
                const int64_t v = v0 + (int64_t) u * kMmThreads;
                if (v < num_vec)
                    raw[u] = ldg_stream(reinterpret_cast<const uint4*>(in) + v);
            }
#pragma unroll
            for (int u = 0; u < kMmUnroll; ++u)
            {
                const int64_t v = v0 + (int64_t) u * kMmThreads;
                if (v < num_vec)
                {
                    float f[kV];
                    Elem<T>::unpack(raw[u], f);
#pragma unroll
                    for (int k = 0; k < kV; ++k)
                    {
                        lo = fminf(lo, f[k]);   // fminf / fmaxf drop NaN, like std::min / std::max with a NaN 2nd arg
                        hi = fmaxf(hi, f[k]);
                    }
                }
            }
        }
        if (cta == 0)
        {
            const int64_t i = num_vec * kV + threadIdx.x;
            if (i < count)
            {
                const float x = Elem<T>::load(in + i);
                lo = fminf(lo, x), hi = fmaxf(hi, x);
            }
        }
    }
    else
    {
        for (int64_t i = cta * kMmThreads + threadIdx.x; i < count; i += num_cta * kMmThreads)
        {
            const float x = Elem<T>::load(in + i);
            lo = fminf(lo, x), hi = fmaxf(hi, x);
        }
    }
}

__device__ __forceinline__ void block_minmax(float& lo, float& hi)
{
    __shared__ float s_lo[32], s_hi[32];
    lo = warp_min(lo), hi = warp_max(hi);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0)
        s_lo[warp] = lo, s_hi[warp] = hi;
    __syncthreads();
    if (warp == 0)
    {
        const int nw = (blockDim.x + 31) >> 5;
        lo           = lane < nw ? s_lo[lane] : INFINITY;
        hi           = lane < nw ? s_hi[lane] : -INFINITY;
        lo = warp_min(lo), hi = warp_max(hi);
    }
}

// quant_mode TF: fold into run_min / run_max.  TF_ENHANCED: leave the batch min/max in the scratch fields for
// hist_kernel (which runs next on the same stream) -- and do nothing at all once the range is fixed.
template <typename T>
__global__ void __launch_bounds__(kMmThreads)
    minmax_kernel(const T* __restrict__ in, int64_t count, int quant_mode, ab_stats_state* st)
{
    if (quant_mode == AB_QUANTIZATION_TF_ENHANCED && st->initialized)
        return;
    float lo, hi;
    thread_minmax(in, count, blockIdx.x, gridDim.x, lo, hi);
    block_minmax(lo, hi);
    if (threadIdx.x == 0)
    {
        atomicMin(&st->batch_min_bits, float_to_ordered(lo));
        atomicMax(&st->batch_max_bits, float_to_ordered(hi));
        if (quant_mode == AB_QUANTIZATION_TF)
        {
            __threadfence();
            const uint32_t t = atomicAdd(&st->ticket, 1u);
            if (t == gridDim.x - 1)
            {
                __threadfence();
                const double cur_min = (double) ordered_to_float(*(volatile int32_t*) &st->batch_min_bits);
                const double cur_max = (double) ordered_to_float(*(volatile int32_t*) &st->batch_max_bits);
                st->run_min          = em::smin(st->run_min, cur_min);   // DlQ/src/TfEncodingAnalyzer.cpp:69-70
                st->run_max          = em::smax(st->run_max, cur_max);
                st->stats_updated    = 1;
                st->batch_min_bits   = kPosInfBits;
                st->batch_max_bits   = kNegInfBits;
                st->ticket           = 0;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// histogram pass (tf_enhanced): TMA-staged tiles, warp-specialised, per-lane privatised shared-memory bins
// ---------------------------------------------------------------------------------------------------------------
constexpr int kConsumerWarps = 16;
constexpr int kHistThreads   = (kConsumerWarps + 2) * 32;   // + a producer warp (drives the TMA engine) + a housekeeping warp
#ifndef AB_HIST_TILE_BYTES
#define AB_HIST_TILE_BYTES 32768
#endif
#ifndef AB_HIST_STAGES
#define AB_HIST_STAGES 4
#endif
constexpr int kTileBytes     = AB_HIST_TILE_BYTES;   // tuning hooks: -DAB_HIST_TILE_BYTES=... -DAB_HIST_STAGES=... (AB_NVCC_EXTRA)
constexpr int kStages        = AB_HIST_STAGES;
constexpr int kLaneCopies    = 32;
constexpr int kVecPerThread  = kTileBytes / 16 / (kConsumerWarps * 32);
constexpr int kHistWords     = (kBins + 1) * kLaneCopies;   // + the dump row for dropped samples
constexpr size_t kHistSmem   = (size_t) kStages * kTileBytes + (size_t) kHistWords * 4 + 2 * kStages * 8 + 64;
static_assert(kVecPerThread * kConsumerWarps * 32 * 16 == kTileBytes, "tile must divide evenly over the consumers");

__device__ __forceinline__ uint32_t smem_u32(const void* p)
{
    return (uint32_t) __cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// 1-D bulk copy global -> shared through the TMA engine, completion signalled on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// The per-sample work of the histogram: 12 instructions, none on the XU pipe, no branch. With M = 1.5 * 2^23:
//   v = x / bucket - offset            exact IEEE quotient via the hoisted reciprocal, then an exact subtraction
//   w = RZ(v + 0.5);  t = RD(w + M)    t == M + floor(v + 0.5) == M + round_half_away(v) for every v > -0.5 (see
//                                      round_half_away_small); the ulp of t is 1, so bits(t) - bits(M) IS the bin
//   u = bits(t) - bits(M), unsigned    every v outside [-0.5, 511.5) -- too small, too large, +-inf, NaN -- gives
//                                      u >= 512: one unsigned min implements the reference's
//                                      `index >= 0 && index < 512` and its NaN / out-of-range drop (x86 cvttss2si
//                                      yields INT_MIN for those)
//   v == -0.5 exactly                  C round() gives -1 (dropped), floor(v + 0.5) gives 0: sent to the dump row too
// Dropped samples increment row 512 of the privatised histogram (never read), which keeps the atomic unconditional.
constexpr int kDumpBin = kBins;

struct Binner
{
    Divisor dv;
    float offset;
    bool fast;
    __device__ __forceinline__ void count(uint32_t* s_hist_lane, float x) const
    {
        uint32_t u;
        if (fast)
        {
            constexpr float kMagic = 12582912.0f;
            float v                = __fsub_rn(div_fast(x, dv), offset);
            v                      = (v == -0.5f) ? -1.0f : v;   // round(-0.5) = -1: dropped, like every v in [-1.5, -0.5)
            const float t          = __fadd_rd(__fadd_rz(v, 0.5f), kMagic);
            u                      = min(__float_as_uint(t) - __float_as_uint(kMagic), (uint32_t) kDumpBin);
        }
        else
        {
            const int b = bin_index(x, dv.d, offset);
            u           = b >= 0 ? (uint32_t) b : (uint32_t) kDumpBin;
        }
        atomicAdd(s_hist_lane + u * kLaneCopies, 1u);
    }
};

template <typename T, bool kFast>
__device__ __forceinline__ void consume_tile(const uint4* __restrict__ src, int nb, int ctid, uint32_t* s_hist_lane,
                                             const Binner& binner)
{
    constexpr int kV = Elem<T>::kPerVec;
    uint4 raw[kVecPerThread];
#pragma unroll
    for (int u = 0; u < kVecPerThread; ++u)
    {
        const int v = ctid + u * (kConsumerWarps * 32);
        if (v * 16 < nb)
            raw[u] = src[v];
    }
#pragma unroll
    for (int u = 0; u < kVecPerThread; ++u)
    {
        const int v = ctid + u * (kConsumerWarps * 32);
        if (v * 16 < nb)
        {
            float f[kV];
            Elem<T>::unpack(raw[u], f);
#pragma unroll
            for (int e = 0; e < kV; ++e)
            {
                Binner b = binner;
                b.fast   = kFast;
                b.count(s_hist_lane, f[e]);
            }
        }
    }
}

// ---- bf16 only: the bin index in one FFMA, certified against the reference sequence over the whole bf16 domain -------
// A bf16 sample has 2^16 bit patterns, so "is floor(fma(x, c, b)) the reference's round(x / bucket - offset) for every
// input?" can be answered by trying them all. The first statistics call on a large bf16 tensor after the range is known
// does that (each CTA's keeper warp takes a slice of the patterns while the consumers stream) for nine (c, b) candidates
// around (1 / bucket, 0.5 - offset); the last CTA records the first candidate that reproduced every pattern -- counted
// bins, samples below and above the range, NaN, +-inf, denormals, the round(-0.5) = -1 corner. Later calls bin with
//   t = RD(fma(x, c, b) + M);  u = min(bits(t) - bits(M), 512)
// 6 instructions per sample with the unpack and the atomic, against 14 for the exact sequence, which made the bf16
// histogram issue-bound at 0.66 of the HBM roofline. If no candidate is exact the record keeps the exact sequence.
// The comparison uses bf16_formula_bin itself, so the claim is about the very instructions the hot loop executes.
constexpr int kBf16Candidates          = 9;
constexpr int64_t kBf16FormulaMinCount = 1 << 20;   // below this the call is latency-bound either way: no certification

__device__ __forceinline__ uint32_t bf16_formula_bin(float x, float c, float b)
{
    constexpr float kMagic = 12582912.0f;   // 1.5 * 2^23: RD(v + M) = M + floor(v), and the ulp of the sum is 1
    const float t          = __fadd_rd(__fmaf_rn(x, c, b), kMagic);
    return min(__float_as_uint(t) - __float_as_uint(kMagic), (uint32_t) kDumpBin);
}
// candidate j: (1 / bucket, 0.5 - offset) moved by -1 / 0 / +1 units in the last place each; j = 0 is the unmoved pair
__device__ __forceinline__ void bf16_candidate(float bucket, float offset, int j, float& c, float& b)
{
    const int i = (j + 4) % kBf16Candidates;
    c           = __int_as_float(__float_as_int(__frcp_rn(bucket)) + (i % 3 - 1));
    b           = __int_as_float(__float_as_int(__fsub_rn(0.5f, offset)) + (i / 3 - 1));
}
// One warp checks the patterns [first, last) against the reference sequence; returns the candidates that failed (bit j).
__device__ __forceinline__ uint32_t bf16_certify_slice(uint32_t first, uint32_t last, int lane, float bucket, float offset)
{
    float c[kBf16Candidates], b[kBf16Candidates];
#pragma unroll
    for (int j = 0; j < kBf16Candidates; ++j)
        bf16_candidate(bucket, offset, j, c[j], b[j]);
    uint32_t failed = 0;
    for (uint32_t p = first + lane; p < last; p += 32)
    {
        const float x      = __uint_as_float(p << 16);
        const int ref      = bin_index(x, bucket, offset);
        const uint32_t exp = ref >= 0 ? (uint32_t) ref : (uint32_t) kDumpBin;
#pragma unroll
        for (int j = 0; j < kBf16Candidates; ++j)
            failed |= (bf16_formula_bin(x, c[j], b[j]) != exp ? 1u : 0u) << j;
    }
    return __reduce_or_sync(0xffffffffu, failed);
}

// Last CTA, one thread, after every CTA has taken its ticket (each published its failures before that): keep the first
// candidate nobody saw fail.
__device__ __forceinline__ void record_bf16_formula(ab_stats_state* st, float bucket, float offset)
{
    __threadfence();
    const uint32_t failed = atomicExch(&st->bf16_fail_mask, 0u);
    const uint32_t good   = ~failed & ((1u << kBf16Candidates) - 1u);
    if (good)
    {
        float c, b;
        bf16_candidate(bucket, offset, __ffs((int) good) - 1, c, b);
        st->bf16_scale   = c;
        st->bf16_shift   = b;
        st->bf16_formula = 1;
    }
    else
        st->bf16_formula = -1;
}

// consumer loop of the certified form: tile words hold two samples each
__device__ __forceinline__ void consume_tile_bf16_formula(const uint4* __restrict__ src, int nb, int ctid,
                                                          uint32_t* s_hist_lane, float c, float b)
{
    uint4 raw[kVecPerThread];
#pragma unroll
    for (int u = 0; u < kVecPerThread; ++u)
    {
        const int v = ctid + u * (kConsumerWarps * 32);
        if (v * 16 < nb)
            raw[u] = src[v];
    }
#pragma unroll
    for (int u = 0; u < kVecPerThread; ++u)
    {
        const int v = ctid + u * (kConsumerWarps * 32);
        if (v * 16 < nb)
        {
            const uint32_t w[4] = {raw[u].x, raw[u].y, raw[u].z, raw[u].w};
#pragma unroll
            for (int e = 0; e < 4; ++e)
            {
                atomicAdd(s_hist_lane + bf16_formula_bin(__uint_as_float(w[e] << 16), c, b) * kLaneCopies, 1u);
                atomicAdd(s_hist_lane + bf16_formula_bin(__uint_as_float(w[e] & 0xffff0000u), c, b) * kLaneCopies, 1u);
            }
        }
    }
}

// What a histogram launch needs to know about its record, read from global memory by ONE thread and handed to the rest of
// the CTA through shared memory: after the barrier that publishes it, no thread of this CTA has a load of these fields in
// flight any more, which is what lets the CTA announce "I have read the record" with a single atomic (see fast_tail).
struct StateSnap
{
    float bucket_size, pdf_offset;
    int32_t batch_min_bits, batch_max_bits, initialized, iterations, pending, write_parity, bf16_formula;
    float bf16_scale, bf16_shift;
    double pending_count;
};

__device__ __forceinline__ unsigned long long global_timer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

template <typename T, bool kPdl>
__global__ void __launch_bounds__(kHistThreads, 1)
    hist_kernel(const T* __restrict__ in, int64_t count, ab_stats_state* st, uint32_t* batch_log, int reverse,
                unsigned long long* timer_slot)
{
    // measurement hook (ab_debug_hist_timer): earliest start / latest end of any CTA of this launch on the GPU's global
    // timer, i.e. the launch's execution time without launch latency or event overhead. nullptr in normal operation.
    auto stamp_start = [&]() {
        if (timer_slot != nullptr && threadIdx.x == 0)
        {
            atomicMin(timer_slot, global_timer_ns());
            if (blockIdx.x == 0)
                timer_slot[2] = (unsigned long long) count * sizeof(T);
        }
    };
    if (!kPdl)
        stamp_start();   // with PDL the CTA may be resident while the previous kernel still runs: stamped after the wait
    struct TimerEnd
    {
        unsigned long long* slot;
        __device__ ~TimerEnd()
        {
            if (slot != nullptr && threadIdx.x == 0)
                atomicMax(slot + 1, global_timer_ns());
        }
    } timer_end {timer_slot};
    constexpr int kV = Elem<T>::kPerVec;
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* s_tiles  = smem;
    uint32_t* s_hist  = reinterpret_cast<uint32_t*>(smem + (size_t) kStages * kTileBytes);
    uint64_t* s_full  = reinterpret_cast<uint64_t*>(smem + (size_t) kStages * kTileBytes + (size_t) kHistWords * 4);
    uint64_t* s_empty = s_full + kStages;
    __shared__ double s_x_left0, s_bucket_d;
    __shared__ uint32_t s_is_last;

    const int tid  = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    constexpr int kProducerWarp = kConsumerWarps;
    constexpr int kKeeperWarp   = kConsumerWarps + 1;

    // ---- prologue, ordered for latency: the first tiles are requested from HBM before anything else happens (a TMA load
    // does not depend on the histogram range), then the state is read and the bins are zeroed while they are in flight.
    const bool aligned      = (reinterpret_cast<uintptr_t>(in) & 15u) == 0;
    const int64_t bytes     = aligned ? (count / kV) * 16 : 0;   // the 16-byte-granular body goes through TMA
    const int64_t num_tiles = (bytes + kTileBytes - 1) / kTileBytes;
    // tiles owned by this CTA: blockIdx.x, blockIdx.x + gridDim.x, ...
    const int64_t my_tiles = (num_tiles > blockIdx.x) ? (num_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    // Tile order: a statistics call usually follows the kernel that PRODUCED the tensor, whose most recently written part
    // (its end) is what the L2 still holds. Walking the tiles from the end backwards consumes that part from L2 before it
    // is evicted; walking forwards would evict it to make room for the start. Bin counts do not depend on the order.
    auto tile_offset = [&](int64_t k) {
        const int64_t t = blockIdx.x + k * gridDim.x;
        return (reverse ? num_tiles - 1 - t : t) * (int64_t) kTileBytes;
    };
    auto issue_tile = [&](int64_t k) {
        const int s       = (int) (k % kStages);
        const int64_t off = tile_offset(k);
        const uint32_t nb = (uint32_t) min((int64_t) kTileBytes, bytes - off);
        mbar_expect_tx(s_full + s, nb);
        tma_load_1d(s_tiles + (size_t) s * kTileBytes, reinterpret_cast<const uint8_t*>(in) + off, nb, s_full + s);
    };
    if (warp == kProducerWarp && lane == 0)
    {
        for (int s = 0; s < kStages; ++s)
        {
            mbar_init(s_full + s, 1);
            mbar_init(s_empty + s, kConsumerWarps);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (kPdl)
    {
        // Launched with programmatic stream serialization: this CTA may be resident while the previous kernel on the stream
        // is still draining. Everything above and the zeroing of the bins touch shared memory only; nothing in global
        // memory (the tensor the previous kernel produced, the state record the previous statistics call updated) is
        // looked at before the dependency is resolved.
        for (int i = tid; i < kHistWords / 4; i += kHistThreads)
            reinterpret_cast<uint4*>(s_hist)[i] = make_uint4(0, 0, 0, 0);
        asm volatile("griddepcontrol.wait;" ::: "memory");
        stamp_start();
    }
    if (warp == kProducerWarp && lane == 0)
        for (int64_t k = 0; k < my_tiles && k < kStages; ++k)
            issue_tile(k);

    __shared__ StateSnap s_snap;
    if (tid == 0)
    {
        s_snap.bucket_size    = st->bucket_size;
        s_snap.pdf_offset     = st->pdf_offset;
        s_snap.batch_min_bits = st->batch_min_bits;
        s_snap.batch_max_bits = st->batch_max_bits;
        s_snap.initialized    = st->initialized;
        s_snap.iterations     = st->iterations;
        s_snap.pending        = st->pending;
        s_snap.write_parity   = st->write_parity;
        s_snap.pending_count  = st->pending_count;
        s_snap.bf16_formula   = st->bf16_formula;
        s_snap.bf16_scale     = st->bf16_scale;
        s_snap.bf16_shift     = st->bf16_shift;
    }
    if (!kPdl)
        for (int i = tid; i < kHistWords / 4; i += kHistThreads)
            reinterpret_cast<uint4*>(s_hist)[i] = make_uint4(0, 0, 0, 0);
    __syncthreads();   // record snapshot published, barriers initialised, bins zeroed
    const StateSnap sn = s_snap;
    double x_left0 = 0, bucket_d = 0;
    const Range rg   = resolve_range(&sn, &x_left0, &bucket_d);   // every thread derives the same range
    // stable for the whole launch: whoever changes them does so after every CTA has taken its snapshot
    const int parity       = sn.write_parity;
    const bool had_pending = sn.pending != 0;
    const int iterations0  = sn.iterations;
    // The batch's counts stay parked in hist[parity]; the NEXT call on this record folds them while it streams its own
    // data. With a batch log (multi-GPU exact merge) every CTA also ADDS its counts to the log entry, which the caller has
    // zeroed. Only a certifying launch with a log still folds and logs at its end (`lazy` false: the old path).
    const bool lazy = batch_log == nullptr;
    // bf16: the certified one-FFMA bin index (see bf16_formula_bin). `formula` and the pair are stable for the launch.
    constexpr bool kIsBf16 = sizeof(T) == 2;
    const int formula      = kIsBf16 ? sn.bf16_formula : 0;
    const float formula_c  = kIsBf16 ? sn.bf16_scale : 0.0f;
    const float formula_b  = kIsBf16 ? sn.bf16_shift : 0.0f;
    const bool certify     = kIsBf16 && rg.valid && formula == 0 && count >= kBf16FormulaMinCount;
    // Tail of the launch. Somebody has to update the record's bookkeeping (pending / parity / iterations ...) once no CTA
    // needs the old values any more. Logging and certifying launches need the LAST CTA TO FINISH (it reads what the others
    // flushed): a ticket taken at the very end, whose round trip to L2 sits on every CTA's critical path. All other
    // launches only need "every CTA has read the record", which is known much earlier: the ticket is taken right here --
    // this CTA's only reads of those fields went into the snapshot above -- and its result is looked at when the CTA is
    // done, by which time it has long arrived. The CTA that drew the last ticket writes the bookkeeping; nobody waits, the
    // consumer warps synchronise among themselves only, and the keeper warp's fold is off the critical path.
    const bool fast_tail = !certify;
    uint32_t my_ticket   = 0xffffffffu;
    if (fast_tail && tid == 0)
        my_ticket = atomicAdd(&st->ticket, 1u);

    if (rg.valid)
    {
        Binner binner;
        binner.dv     = make_divisor(rg.bucket);
        binner.offset = rg.offset;
        binner.fast   = binner.dv.fast;
        uint32_t* s_hist_lane = s_hist + lane;

        if (warp == kProducerWarp)
        {
            // ---- producer warp: one lane keeps the ring full ----
            if (lane == 0)
                for (int64_t k = kStages; k < my_tiles; ++k)
                {
                    const int s = (int) (k % kStages);
                    mbar_wait(s_empty + s, (uint32_t) (((k / kStages) - 1) & 1));   // consumers released the slot
                    issue_tile(k);
                }
        }
        else if (warp == kKeeperWarp)
        {
            // ---- housekeeping warps: fold the previous batch of this record while the consumers stream. Every CTA's
            // keeper takes a 32-bin slice (one independent load / divide / store per lane), so the whole fold is one
            // memory round trip off the critical path; `iterations` / `pending` are updated by the elected CTA. ----
            if (had_pending)
            {
                const int pp     = parity ^ 1;
                const double cnt = sn.pending_count;
                for (int b = blockIdx.x * 32 + lane; b < kBins; b += gridDim.x * 32)
                {
                    fold_bin(&st->pdf[b], st->hist[pp][b], cnt, iterations0);
                    st->hist[pp][b] = 0;
                }
            }
            if (certify)
            {
                const uint32_t per    = (65536u + gridDim.x - 1) / gridDim.x;
                const uint32_t first  = min(65536u, blockIdx.x * per);
                const uint32_t failed = bf16_certify_slice(first, min(65536u, first + per), lane, rg.bucket, rg.offset);
                if (lane == 0 && failed)
                    atomicOr(&st->bf16_fail_mask, failed);
            }
            if (!fast_tail && (had_pending || certify))
                __threadfence();   // ordered before this CTA's end ticket, hence before what the last CTA reads
        }
        else
        {
            // ---- consumer warps ----
            for (int64_t k = 0; k < my_tiles; ++k)
            {
                const int s       = (int) (k % kStages);
                const int64_t off = tile_offset(k);
                const int nb      = (int) min((int64_t) kTileBytes, bytes - off);
                mbar_wait(s_full + s, (uint32_t) ((k / kStages) & 1));
                const uint4* src = reinterpret_cast<const uint4*>(s_tiles + (size_t) s * kTileBytes);
                if (kIsBf16 && formula == 1)
                    consume_tile_bf16_formula(src, nb, tid, s_hist_lane, formula_c, formula_b);
                else if (binner.fast)
                    consume_tile<T, true>(src, nb, tid, s_hist_lane, binner);
                else
                    consume_tile<T, false>(src, nb, tid, s_hist_lane, binner);
                __syncwarp();
                if (lane == 0)
                    mbar_arrive(s_empty + s);   // this warp is done reading stage s
            }
            // elements the TMA body did not cover: the sub-vector tail, or everything when the base is misaligned
            const int64_t body = aligned ? (count / kV) * kV : 0;
            for (int64_t i = body + (int64_t) blockIdx.x * (kConsumerWarps * 32) + tid; i < count;
                 i += (int64_t) gridDim.x * (kConsumerWarps * 32))
                binner.count(s_hist_lane, Elem<T>::load(in + i));
        }
        if (!fast_tail)
            __syncthreads();
        else if (warp < kConsumerWarps)
            asm volatile("bar.sync 1, %0;" ::"n"(kConsumerWarps * 32) : "memory");   // the consumers among themselves

        // reduce the 32 lane copies of bin `tid` (rotated start: conflict-free) and flush
        if (tid < kBins)
        {
            uint32_t sum = 0;
#pragma unroll
            for (int l = 0; l < kLaneCopies; ++l)
                sum += s_hist[tid * kLaneCopies + ((l + tid) & (kLaneCopies - 1))];
            if (sum)
            {
                atomicAdd(&st->hist[parity][tid], sum);
                if (fast_tail && batch_log != nullptr)
                    atomicAdd(&batch_log[tid], sum);
            }
        }
    }
    else if (warp == kProducerWarp && lane == 0)
    {
        // nothing to count (no range yet and an all-zero batch), but the tiles requested in the prologue must land before
        // this CTA's shared memory is released
        for (int64_t k = 0; k < my_tiles && k < kStages; ++k)
            mbar_wait(s_full + k, 0);
    }

    if (fast_tail)
    {
        // ---- the CTA that drew the last "I have read the record" ticket does the bookkeeping; everybody leaves ----------
        if (tid == 0 && my_ticket == gridDim.x - 1)
        {
            if (rg.valid)
            {
                if (!sn.initialized)
                {
                    st->x_left0       = x_left0;
                    st->bucket_size_d = bucket_d;
                    st->bucket_size   = rg.bucket;
                    st->pdf_offset    = rg.offset;
                    st->initialized   = 1;
                }
                if (had_pending)
                    st->iterations = iterations0 + 1;   // the keepers fold the previous batch during this launch
                st->pending_count = (double) count;
                st->pending       = 1;
                st->write_parity  = parity ^ 1;
            }
            st->stats_updated  = 1;
            st->batch_min_bits = kPosInfBits;
            st->batch_max_bits = kNegInfBits;
            st->ticket         = 0;
            if (batch_log != nullptr)
            {
                // element count (0 when the batch was skipped, as the reference skips all-zero batches before init)
                const uint64_t c     = rg.valid ? (uint64_t) count : 0;
                batch_log[kBins]     = (uint32_t) c;
                batch_log[kBins + 1] = (uint32_t) (c >> 32);
            }
        }
        return;
    }

    // ---- election of the last CTA to finish (logging / certifying launches) ---------------------------------------------
    if (!lazy)
        __threadfence();   // the last CTA will READ the flushed counts in this very launch
    __syncthreads();
    if (tid == 0)
    {
        const uint32_t t = atomicAdd(&st->ticket, 1u);
        s_is_last        = (t == gridDim.x - 1);
        if (s_is_last && rg.valid && !sn.initialized)
            s_x_left0 = x_left0, s_bucket_d = bucket_d;
    }
    __syncthreads();
    if (!s_is_last)
        return;

    if (lazy)
    {
        // Only bookkeeping here: a handful of scalar stores by one thread. The 512-bin fold is the next call's business.
        if (tid == 0)
        {
            if (rg.valid)
            {
                if (!st->initialized)
                {
                    st->x_left0       = s_x_left0;
                    st->bucket_size_d = s_bucket_d;
                    st->bucket_size   = rg.bucket;
                    st->pdf_offset    = rg.offset;
                    st->initialized   = 1;
                }
                if (had_pending)
                    st->iterations = iterations0 + 1;   // the keepers folded the previous batch during this launch
                st->pending_count = (double) count;
                st->pending       = 1;
                st->write_parity  = parity ^ 1;
                if (certify)
                    record_bf16_formula(st, rg.bucket, rg.offset);
            }
            st->stats_updated  = 1;
            st->batch_min_bits = kPosInfBits;
            st->batch_max_bits = kNegInfBits;
            st->ticket         = 0;
        }
        return;
    }

    // ---- logging mode: fold this batch now and write its raw counts to the log ------------------------------------------
    __threadfence();
    const int iterations = iterations0 + ((had_pending && rg.valid) ? 1 : 0);   // the keepers folded a parked batch first
    if (tid < kBins)
    {
        if (rg.valid)
        {
            const uint32_t h = *(volatile uint32_t*) &st->hist[parity][tid];
            fold_bin(&st->pdf[tid], h, (double) count, iterations);
            st->hist[parity][tid] = 0;
            batch_log[tid]        = h;
        }
        else
            batch_log[tid] = 0;
    }
    __syncthreads();   // every thread has read st->iterations / st->initialized before they change
    if (tid == 0)
    {
        // element count (0 when the batch was skipped, as the reference skips all-zero batches before init)
        const uint64_t c     = rg.valid ? (uint64_t) count : 0;
        batch_log[kBins]     = (uint32_t) c;
        batch_log[kBins + 1] = (uint32_t) (c >> 32);
        if (rg.valid)
        {
            if (!st->initialized)
            {
                st->x_left0       = s_x_left0;
                st->bucket_size_d = s_bucket_d;
                st->bucket_size   = rg.bucket;
                st->pdf_offset    = rg.offset;
                st->initialized   = 1;
            }
            st->iterations = iterations + 1;
            st->pending    = 0;
            if (certify)
                record_bf16_formula(st, rg.bucket, rg.offset);
        }
        st->stats_updated  = 1;
        st->batch_min_bits = kPosInfBits;
        st->batch_max_bits = kNegInfBits;
        st->ticket         = 0;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// segmented statistics: one CTA per segment
// ---------------------------------------------------------------------------------------------------------------
constexpr int kSegThreads = 256;

// updateStats of ONE segment x[0 .. seg_len) on its record, by one CTA of kSegThreads threads. s_hist: kBins words.
template <typename T>
__device__ __forceinline__ void segment_update(const T* __restrict__ x, int64_t seg_len, int quant_mode, ab_stats_state* st,
                                               uint32_t* s_hist, float* s_lo_hi)
{
    constexpr int kV   = Elem<T>::kPerVec;
    const bool vec     = (reinterpret_cast<uintptr_t>(x) & 15u) == 0;
    const int64_t nvec = vec ? seg_len / kV : 0;
    const bool need_mm = (quant_mode == AB_QUANTIZATION_TF) || !st->initialized;

    __syncthreads();   // previous segment's fold has finished with s_hist / s_lo_hi
    if (st->pending)   // a batch parked by hist_kernel: fold it before this call's own batch
    {
        fold_pending(st, threadIdx.x, kSegThreads);
        __syncthreads();
        if (threadIdx.x == 0)
        {
            st->iterations = st->iterations + 1;
            st->pending    = 0;
        }
        __syncthreads();
    }
    if (need_mm)
    {
        float lo = INFINITY, hi = -INFINITY;
        for (int64_t v = threadIdx.x; v < nvec; v += kSegThreads)
        {
            float f[kV];
            Elem<T>::unpack(__ldg(reinterpret_cast<const uint4*>(x) + v), f);
#pragma unroll
            for (int k = 0; k < kV; ++k)
                lo = fminf(lo, f[k]), hi = fmaxf(hi, f[k]);
        }
        for (int64_t i = nvec * kV + threadIdx.x; i < seg_len; i += kSegThreads)
        {
            const float xv = Elem<T>::load(x + i);
            lo = fminf(lo, xv), hi = fmaxf(hi, xv);
        }
        block_minmax(lo, hi);
        if (threadIdx.x == 0)
            s_lo_hi[0] = lo, s_lo_hi[1] = hi;
    }
    for (int i = threadIdx.x; i < kBins; i += kSegThreads)
        s_hist[i] = 0;
    __syncthreads();
    const float s_lo = s_lo_hi[0], s_hi = s_lo_hi[1];

    if (quant_mode == AB_QUANTIZATION_TF)
    {
        if (threadIdx.x == 0)
        {
            st->run_min       = em::smin(st->run_min, (double) s_lo);
            st->run_max       = em::smax(st->run_max, (double) s_hi);
            st->stats_updated = 1;
        }
        return;
    }

    // tf_enhanced
    Range rg;
    double x_left0 = 0, bucket_d = 0;
    const bool was_init = st->initialized != 0;
    if (was_init)
    {
        rg.bucket = st->bucket_size, rg.offset = st->pdf_offset, rg.valid = true;
    }
    else if (s_lo == 0 && s_hi == 0)
    {
        rg.bucket = 1.0f, rg.offset = 0.0f, rg.valid = false;
    }
    else
    {
        em::init_pdf_range(s_lo, s_hi, x_left0, bucket_d, rg.bucket, rg.offset);
        rg.valid = true;
    }
    if (rg.valid)
    {
        for (int64_t v = threadIdx.x; v < nvec; v += kSegThreads)
        {
            float f[kV];
            Elem<T>::unpack(__ldg(reinterpret_cast<const uint4*>(x) + v), f);
#pragma unroll
            for (int k = 0; k < kV; ++k)
            {
                const int b = bin_index(f[k], rg.bucket, rg.offset);
                if (b >= 0)
                    atomicAdd(s_hist + b, 1u);
            }
        }
        for (int64_t i = nvec * kV + threadIdx.x; i < seg_len; i += kSegThreads)
        {
            const int b = bin_index(Elem<T>::load(x + i), rg.bucket, rg.offset);
            if (b >= 0)
                atomicAdd(s_hist + b, 1u);
        }
        __syncthreads();
        const int iterations = st->iterations;
        for (int i = threadIdx.x; i < kBins; i += kSegThreads)
            fold_bin(&st->pdf[i], s_hist[i], (double) seg_len, iterations);
        __syncthreads();   // every thread has read st->iterations before it changes
        if (threadIdx.x == 0)
        {
            if (!was_init)
            {
                st->x_left0       = x_left0;
                st->bucket_size_d = bucket_d;
                st->bucket_size   = rg.bucket;
                st->pdf_offset    = rg.offset;
                st->initialized   = 1;
            }
            st->iterations = iterations + 1;
        }
    }
    if (threadIdx.x == 0)
        st->stats_updated = 1;
}

template <typename T>
__global__ void __launch_bounds__(kSegThreads)
    segmented_kernel(const T* __restrict__ in, int64_t num_segments, int64_t seg_len, int quant_mode,
                     ab_stats_state* states)
{
    __shared__ uint32_t s_hist[kBins];
    __shared__ float s_lo_hi[2];
    for (int64_t seg = blockIdx.x; seg < num_segments; seg += gridDim.x)
        segment_update(in + seg * seg_len, seg_len, quant_mode, states + seg, s_hist, s_lo_hi);
}

// The same for the segments of MANY tensors in one launch: record r belongs to the item whose [first, first + count) range
// holds it (a refresh of all parameter encodings of a model: ab_stats_refresh_encodings_multi).
constexpr int kMaxItems = AB_REFRESH_MULTI_MAX_ITEMS;
struct ItemTable
{
    const void* data[kMaxItems];
    int64_t segment_len[kMaxItems];
    int32_t first[kMaxItems + 1];   // first record of item i; [num_items] = total records
    int32_t skip[kMaxItems];        // 1: handled by the single-tensor kernels (one large per-tensor segment)
    int32_t num_items;
};
static_assert(sizeof(ItemTable) <= 3200, "must travel as a kernel parameter");

__device__ __forceinline__ int item_of(const ItemTable& t, int64_t record)
{
    int lo = 0, hi = t.num_items - 1;   // largest i with first[i] <= record
    while (lo < hi)
    {
        const int mid = (lo + hi + 1) >> 1;
        if (t.first[mid] <= record)
            lo = mid;
        else
            hi = mid - 1;
    }
    return lo;
}

template <typename T>
__global__ void __launch_bounds__(kSegThreads)
    segmented_multi_kernel(const __grid_constant__ ItemTable t, int quant_mode, ab_stats_state* states)
{
    __shared__ uint32_t s_hist[kBins];
    __shared__ float s_lo_hi[2];
    const int64_t total = t.first[t.num_items];
    for (int64_t rec = blockIdx.x; rec < total; rec += gridDim.x)
    {
        const int i = item_of(t, rec);
        if (t.skip[i])
            continue;
        const int64_t len = t.segment_len[i];
        segment_update(reinterpret_cast<const T*>(t.data[i]) + (rec - t.first[i]) * len, len, quant_mode, states + rec,
                       s_hist, s_lo_hi);
    }
}

// The tf scheme's refresh of many parameter tensors, fused: reset + updateStats + computeEncoding of a record need nothing but
// the segment's min / max, so ONE WARP per record reads its segment (128-bit loads), reduces with shuffles, and lane 0 writes
// the record's header as reset + update would leave it, the encoding row and the per-tensor kernel parameters. No 8 KB reset
// of a PDF the scheme never looks at, no CTA per record for a scalar computation: a Llama-2-7B-shaped W4 model (1.36 M
// weight channels, 13 GB of bf16 weights) went from 10.7 ms to the time of reading the weights once. The arrays of the
// record (pdf, hist) are left as they are: the tf scheme never reads them, and every tf_enhanced entry point that needs
// them zero starts with a full ab_stats_reset.
template <typename T>
__global__ void __launch_bounds__(256)
    tf_refresh_kernel(const __grid_constant__ ItemTable t, ab_stats_state* states, int bw, int sym, int strict, int unsigned_sym,
                      double* __restrict__ enc_out, float* __restrict__ qdq4_out)
{
    constexpr int kV     = Elem<T>::kPerVec;
    const int lane       = threadIdx.x & 31;
    const int64_t total  = t.first[t.num_items];
    const int64_t warps  = (int64_t) gridDim.x * (blockDim.x >> 5);
    for (int64_t rec = (int64_t) blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); rec < total; rec += warps)
    {
        const int i = item_of(t, rec);
        if (t.skip[i])
            continue;
        const int64_t len = t.segment_len[i];
        const T* x        = reinterpret_cast<const T*>(t.data[i]) + (rec - t.first[i]) * len;
        float lo = INFINITY, hi = -INFINITY;
        // scalar head up to the first 16-byte boundary, 128-bit body, scalar tail
        int64_t head = ((16 - (reinterpret_cast<uintptr_t>(x) & 15u)) & 15u) / sizeof(T);
        head         = head < len ? head : len;
        for (int64_t k = lane; k < head; k += 32)
        {
            const float v = Elem<T>::load(x + k);
            lo = fminf(lo, v), hi = fmaxf(hi, v);
        }
        const int64_t nvec = (len - head) / kV;
        const uint4* xv    = reinterpret_cast<const uint4*>(x + head);
        for (int64_t v = lane; v < nvec; v += 32)
        {
            float f[kV];
            Elem<T>::unpack(__ldg(xv + v), f);
#pragma unroll
            for (int k = 0; k < kV; ++k)
                lo = fminf(lo, f[k]), hi = fmaxf(hi, f[k]);
        }
        for (int64_t k = head + nvec * kV + lane; k < len; k += 32)
        {
            const float v = Elem<T>::load(x + k);
            lo = fminf(lo, v), hi = fmaxf(hi, v);
        }
        lo = warp_min(lo), hi = warp_max(hi);
        if (lane != 0)
            continue;
        ab_stats_state* st = states + rec;
        st->pending_count  = 0.0;            // reset_kernel's header ...
        st->pending        = 0;
        st->write_parity   = 0;
        st->x_left0        = 0.0;
        st->bucket_size_d  = 0.0;
        st->bucket_size    = 0.0f;
        st->pdf_offset     = 0.0f;
        st->batch_min_bits = kPosInfBits;
        st->batch_max_bits = kNegInfBits;
        st->initialized    = 0;
        st->iterations     = 0;
        st->ticket         = 0;
        st->bf16_scale     = 0.0f;
        st->bf16_shift     = 0.0f;
        st->bf16_formula   = 0;
        st->bf16_fail_mask = 0;
        const double run_min = em::smin(DBL_MAX, (double) lo);   // ... and the one update (TfEncodingAnalyzer.cpp:69-70)
        const double run_max = em::smax(-DBL_MAX, (double) hi);
        st->run_min          = run_min;
        st->run_max          = run_max;
        st->stats_updated    = 1;
        ab_encoding e;
        em::tf_analyzer_encoding(bw, run_min, run_max, sym != 0, strict != 0, unsigned_sym != 0, e);
        double* o = enc_out + rec * 5;
        o[0] = e.min, o[1] = e.max, o[2] = e.delta, o[3] = e.offset, o[4] = (double) e.bw;
        if (qdq4_out)
        {
            ab_encoding full;
            em::fill_encoding_info(e.bw, e.min, e.max, full);
            reinterpret_cast<float4*>(qdq4_out)[rec] =
                make_float4((float) full.min, (float) full.max, (float) full.delta, (float) full.offset);
        }
    }
}

// per-channel QDQ parameter blocks of many tensors from their encoding rows: item i's float[4][C_i] block starts at
// params + 4 * first[i]; the step count is decided from the item's channel 0 (ATQ:286-294), as per_channel_params_kernel does
__global__ void per_channel_params_multi_kernel(const __grid_constant__ ItemTable t, const double* __restrict__ enc5, int bw,
                                                float* __restrict__ params)
{
    const int64_t total = t.first[t.num_items];
    const int64_t rec   = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (rec >= total)
        return;
    const int i        = item_of(t, rec);
    const int64_t f    = t.first[i];
    const int64_t nch  = t.first[i + 1] - f;
    const int64_t c    = rec - f;
    double steps       = em::pow2(bw) - 1;
    if (enc5[f * 5] == -enc5[f * 5 + 1])
        steps -= 1;
    float* p = params + 4 * f;
    em::per_channel_param(enc5[rec * 5], enc5[rec * 5 + 1], (float) steps, p[c], p[nch + c], p[2 * nch + c], p[3 * nch + c]);
}

// ---------------------------------------------------------------------------------------------------------------
// range injection and ordered replay (multi-GPU exact merge)
// ---------------------------------------------------------------------------------------------------------------
// minmax: [count][2] floats = the batch min/max that fixes each quantizer's range; (0,0) leaves it uninitialised
__global__ void init_range_kernel(ab_stats_state* states, int64_t count, const float* __restrict__ minmax)
{
    const int64_t s = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= count)
        return;
    ab_stats_state* st = states + s;
    const float mn = minmax[2 * s], mx = minmax[2 * s + 1];
    if (st->initialized || (mn == 0 && mx == 0))
        return;
    double x0, bd;
    float bf, of;
    em::init_pdf_range(mn, mx, x0, bd, bf, of);
    st->x_left0       = x0;
    st->bucket_size_d = bd;
    st->bucket_size   = bf;
    st->pdf_offset    = of;
    st->initialized   = 1;
}

// One CTA per quantizer, one thread per bin. Replays pdf = (pdf*k + hist/cnt)/(k+1) over the batches in order.
__global__ void __launch_bounds__(kBins)
    fold_batches_kernel(ab_stats_state* states, int64_t count, const uint32_t* __restrict__ batch_log,
                        const int64_t* __restrict__ batch_offsets, int64_t num_batches)
{
    const int64_t s = blockIdx.x;
    if (s >= count)
        return;
    ab_stats_state* st = states + s;
    const int b        = threadIdx.x;
    double pdf         = 0.0;
    int iterations     = 0;
    bool seen          = false;
    for (int64_t k = 0; k < num_batches; ++k)
    {
        const uint32_t* entry = batch_log + batch_offsets[k] + s * (kBins + 2);
        const uint64_t cnt    = (uint64_t) entry[kBins] | ((uint64_t) entry[kBins + 1] << 32);
        if (cnt == 0)
            continue;
        seen              = true;
        const double prob = (double) entry[b] / (double) cnt;
        pdf               = __ddiv_rn(__dadd_rn(__dmul_rn(pdf, (double) iterations), prob), (double) (iterations + 1));
        ++iterations;
    }
    st->pdf[b]     = pdf;
    st->hist[0][b] = 0;
    st->hist[1][b] = 0;
    if (b == 0)
    {
        st->pending    = 0;
        st->iterations = iterations;
        if (seen)
            st->stats_updated = 1;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// multi-tensor histogram: many (tensor, record) pairs in ONE persistent launch
// ---------------------------------------------------------------------------------------------------------------
// A calibration forward hands ~70 activation tensors of 0.4 ... 100 MB to updateStats. One launch per tensor costs a
// pipeline ramp-up and drain (~2 us inside the kernel) plus a launch gap each, against 4.5 us of streaming for the
// average 30 MB tensor. Where the host layer can prove that a tensor is not written between the call and the end of the
// forward (aimet_b200/quantsim/stats_batcher.py), it defers the call, and the whole batch of deferred calls becomes ONE
// launch here: the 32 KB tiles of all tensors form one global tile list, every persistent CTA takes a contiguous chunk of
// it (so it changes tensor -- flushes and re-arms its privatised bins, re-reads a range -- once or twice per launch),
// and the TMA ring never drains between tensors. The raw counts go to one row per call (`seg_counts`); a second, tiny
// launch folds the rows into the running PDFs in call order, exactly as UpdatePdf would have (math_functions.cpp:279-287)
// -- or leaves them as log entries for the multi-GPU exact merge.
constexpr int kMaxSeg   = AB_STATS_MULTI_MAX_SEGMENTS;
constexpr int kLogWords = kBins + 2;
struct MultiParams
{
    const void* data[kMaxSeg];
    int64_t count[kMaxSeg];
    int32_t state_index[kMaxSeg];
    int32_t first_tile[kMaxSeg + 1];   // prefix sum of the tensors' tile counts; [num_segments] is the total
    int32_t num_segments;
};
static_assert(sizeof(MultiParams) <= 3200, "must travel as a kernel parameter (4 KB limit with the other arguments)");

template <typename T>
__global__ void __launch_bounds__(kHistThreads, 1)
    hist_multi_kernel(const __grid_constant__ MultiParams p, const ab_stats_state* __restrict__ states,
                      uint32_t* __restrict__ seg_counts, unsigned long long* timer_slot)
{
    constexpr int kV = Elem<T>::kPerVec;
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* s_tiles  = smem;
    uint32_t* s_hist  = reinterpret_cast<uint32_t*>(smem + (size_t) kStages * kTileBytes);
    uint64_t* s_full  = reinterpret_cast<uint64_t*>(smem + (size_t) kStages * kTileBytes + (size_t) kHistWords * 4);
    uint64_t* s_empty = s_full + kStages;

    const int tid  = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    constexpr int kProducerWarp = kConsumerWarps;

    const int total    = p.first_tile[p.num_segments];
    const int t0       = (int) ((int64_t) blockIdx.x * total / gridDim.x);
    const int t1       = (int) ((int64_t) (blockIdx.x + 1) * total / gridDim.x);
    const int my_tiles = t1 - t0;
    auto seg_of = [&](int tile, int hint) {
        while (tile >= p.first_tile[hint + 1])
            ++hint;
        return hint;
    };
    auto tile_bytes = [&](int seg, int tile, int64_t& off) {
        const int64_t body = (p.count[seg] / kV) * 16;
        off                = (int64_t) (tile - p.first_tile[seg]) * kTileBytes;
        return (uint32_t) min((int64_t) kTileBytes, body - off);
    };

    if (warp == kProducerWarp && lane == 0)
    {
        for (int s = 0; s < kStages; ++s)
        {
            mbar_init(s_full + s, 1);
            mbar_init(s_empty + s, kConsumerWarps);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < kHistWords / 4; i += kHistThreads)
        reinterpret_cast<uint4*>(s_hist)[i] = make_uint4(0, 0, 0, 0);
    // programmatic stream serialization: everything above touched shared memory only (a no-op in a plain launch)
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (timer_slot != nullptr && tid == 0)
    {
        atomicMin(timer_slot, global_timer_ns());
        if (blockIdx.x == 0)
        {
            unsigned long long bytes = 0;
            for (int s = 0; s < p.num_segments; ++s)
                bytes += (unsigned long long) p.count[s] * sizeof(T);
            timer_slot[2] = bytes;
        }
    }
    __syncthreads();   // barriers initialised, bins zeroed

    if (warp == kProducerWarp)
    {
        if (lane == 0 && my_tiles > 0)
        {
            int seg = seg_of(t0, 0);
            for (int k = 0; k < my_tiles; ++k)
            {
                const int tile = t0 + k;
                seg            = seg_of(tile, seg);
                const int s    = k % kStages;
                if (k >= kStages)
                    mbar_wait(s_empty + s, (uint32_t) (((k / kStages) - 1) & 1));   // consumers released the slot
                int64_t off;
                const uint32_t nb = tile_bytes(seg, tile, off);
                mbar_expect_tx(s_full + s, nb);
                tma_load_1d(s_tiles + (size_t) s * kTileBytes, reinterpret_cast<const uint8_t*>(p.data[seg]) + off, nb,
                            s_full + s);
            }
        }
    }
    else if (warp < kConsumerWarps)
    {
        uint32_t* s_hist_lane = s_hist + lane;
        // reduce the 32 lane copies of bin `tid`, re-arm them, add the sum to the call's row
        auto flush_bins = [&](int seg) {
            asm volatile("bar.sync 1, %0;" ::"n"(kConsumerWarps * 32) : "memory");   // everybody has counted
            uint32_t sum = 0;
#pragma unroll
            for (int l = 0; l < kLaneCopies; ++l)
            {
                uint32_t* w = s_hist + tid * kLaneCopies + ((l + tid) & (kLaneCopies - 1));
                sum += *w;
                *w = 0;
            }
            if (sum)
                atomicAdd(seg_counts + (size_t) seg * kLogWords + tid, sum);
            asm volatile("bar.sync 1, %0;" ::"n"(kConsumerWarps * 32) : "memory");   // bins are clean again
        };
        int seg = -1;
        Binner binner;
        binner.fast      = false;
        bool valid       = false;
        int formula      = 0;
        float formula_c  = 0.0f, formula_b = 0.0f;
        constexpr bool kIsBf16 = sizeof(T) == 2;
        for (int k = 0; k < my_tiles; ++k)
        {
            const int tile = t0 + k;
            const int nseg = seg_of(tile, seg < 0 ? 0 : seg);
            if (nseg != seg)
            {
                if (seg >= 0)
                    flush_bins(seg);
                seg                      = nseg;
                const ab_stats_state* st = states + p.state_index[seg];
                valid                    = st->initialized != 0;   // the host layer only defers calls on fixed ranges
                binner.dv                = make_divisor(st->bucket_size);
                binner.offset            = st->pdf_offset;
                binner.fast              = binner.dv.fast;
                if (kIsBf16)
                    formula = st->bf16_formula, formula_c = st->bf16_scale, formula_b = st->bf16_shift;
            }
            const int s = k % kStages;
            int64_t off;
            const int nb = (int) tile_bytes(seg, tile, off);
            mbar_wait(s_full + s, (uint32_t) ((k / kStages) & 1));
            if (valid)
            {
                const uint4* src = reinterpret_cast<const uint4*>(s_tiles + (size_t) s * kTileBytes);
                if (kIsBf16 && formula == 1)
                    consume_tile_bf16_formula(src, nb, tid, s_hist_lane, formula_c, formula_b);
                else if (binner.fast)
                    consume_tile<T, true>(src, nb, tid, s_hist_lane, binner);
                else
                    consume_tile<T, false>(src, nb, tid, s_hist_lane, binner);
            }
            __syncwarp();
            if (lane == 0)
                mbar_arrive(s_empty + s);
            if (valid && tile == p.first_tile[seg + 1] - 1)
            {
                // the sub-vector tail of this tensor (count % kV elements) belongs to whoever owns its last tile
                const T* in        = reinterpret_cast<const T*>(p.data[seg]);
                const int64_t body = (p.count[seg] / kV) * kV;
                if (body + tid < p.count[seg])
                    binner.count(s_hist_lane, Elem<T>::load(in + body + tid));
            }
        }
        if (seg >= 0)
            flush_bins(seg);
    }
    if (timer_slot != nullptr && tid == 0)
        atomicMax(timer_slot + 1, global_timer_ns());
}

// Second launch of a multi-tensor update: CTA s handles the record of call s if s is that record's FIRST call in the table,
// and then folds all of the record's calls in table order -- a reused module (a ReLU called three times per residual
// block) updates its quantizer several times per forward, and the running mean is order dependent. One thread per bin.
__global__ void __launch_bounds__(kBins)
    fold_segments_kernel(const __grid_constant__ MultiParams p, ab_stats_state* states, uint32_t* seg_counts, int log_only)
{
    const int s = blockIdx.x;
    if (s >= p.num_segments)
        return;
    const int idx = p.state_index[s];
    for (int j = 0; j < s; ++j)
        if (p.state_index[j] == idx)
            return;
    ab_stats_state* st = states + idx;
    const int b        = threadIdx.x;
    const bool valid   = st->initialized != 0;
    if (log_only)
    {
        // leave the raw counts where they are and complete the entries with their element counts (0: call not counted)
        if (b == 0)
        {
            for (int j = s; j < p.num_segments; ++j)
                if (p.state_index[j] == idx)
                {
                    const uint64_t c                           = valid ? (uint64_t) p.count[j] : 0;
                    seg_counts[(size_t) j * kLogWords + kBins]     = (uint32_t) c;
                    seg_counts[(size_t) j * kLogWords + kBins + 1] = (uint32_t) (c >> 32);
                }
            st->stats_updated = 1;
        }
        return;
    }
    double pdf           = st->pdf[b];
    int k                = st->iterations;
    const int pending    = st->pending;
    if (pending)   // a batch parked by the single-tensor kernel comes first
    {
        const int pp = st->write_parity ^ 1;
        pdf          = __ddiv_rn(__dadd_rn(__dmul_rn(pdf, (double) k), (double) st->hist[pp][b] / st->pending_count),
                                 (double) (k + 1));
        st->hist[pp][b] = 0;
        ++k;
    }
    if (valid)
        for (int j = s; j < p.num_segments; ++j)
            if (p.state_index[j] == idx)
            {
                uint32_t* e      = seg_counts + (size_t) j * kLogWords + b;
                const double prob = (double) *e / (double) p.count[j];
                pdf               = __ddiv_rn(__dadd_rn(__dmul_rn(pdf, (double) k), prob), (double) (k + 1));
                ++k;
                *e = 0;   // the scratch rows are handed back zeroed
            }
    st->pdf[b] = pdf;
    __syncthreads();   // every thread has read iterations / pending before they change
    if (b == 0)
    {
        st->iterations    = k;
        st->pending       = 0;
        st->stats_updated = 1;
    }
}

// One CTA per record, one thread per bin: replay the record's log entries (rows of kLogWords words anywhere in `log`) in
// the given order. record_begin is a CSR index into entry_rows. Entries whose element count is 0 are skipped.
__global__ void __launch_bounds__(kBins)
    fold_log_kernel(ab_stats_state* states, int64_t count, const uint32_t* __restrict__ log,
                    const int64_t* __restrict__ entry_rows, const int64_t* __restrict__ record_begin)
{
    const int64_t s = blockIdx.x;
    if (s >= count)
        return;
    ab_stats_state* st = states + s;
    const int b        = threadIdx.x;
    double pdf         = 0.0;
    int iterations     = 0;
    bool seen          = false;
    for (int64_t e = record_begin[s]; e < record_begin[s + 1]; ++e)
    {
        const uint32_t* entry = log + entry_rows[e] * kLogWords;
        const uint64_t cnt    = (uint64_t) entry[kBins] | ((uint64_t) entry[kBins + 1] << 32);
        if (cnt == 0)
            continue;
        seen              = true;
        const double prob = (double) entry[b] / (double) cnt;
        pdf               = __ddiv_rn(__dadd_rn(__dmul_rn(pdf, (double) iterations), prob), (double) (iterations + 1));
        ++iterations;
    }
    st->pdf[b]     = pdf;
    st->hist[0][b] = 0;
    st->hist[1][b] = 0;
    if (b == 0)
    {
        st->pending    = 0;
        st->iterations = iterations;
        if (seen)
            st->stats_updated = 1;
    }
}

template <typename K>
int resident_grid(K kernel, int threads, size_t smem)
{
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem) != cudaSuccess || per_sm <= 0)
        per_sm = 1;
    return per_sm * num_sms();
}

bool check_common(const void* in, int64_t count, int dtype, int quant_mode, const void* state)
{
    if (count < 0 || (count > 0 && in == nullptr) || state == nullptr)
    {
        set_error("null pointer or negative count");
        return false;
    }
    if (dtype != AB_F32 && dtype != AB_BF16)
    {
        set_error("unsupported dtype %d", dtype);
        return false;
    }
    if (quant_mode != AB_QUANTIZATION_TF && quant_mode != AB_QUANTIZATION_TF_ENHANCED)
    {
        set_error("unsupported quantization mode %d (tf, tf_enhanced, percentile and mse are on the hot path)", quant_mode);
        return false;
    }
    return true;
}

// ab_debug_hist_timer: slots handed to the histogram launches that follow, one {start, end} pair each
unsigned long long* g_timer_slots = nullptr;
int64_t g_timer_capacity = 0, g_timer_used = 0;

template <typename T>
int launch_update(const T* in, int64_t count, int quant_mode, ab_stats_state* st, uint32_t* batch_log, int flags,
                  cudaStream_t stream)
{
    constexpr int kV = Elem<T>::kPerVec;
    if (!(quant_mode == AB_QUANTIZATION_TF_ENHANCED && (flags & AB_STATS_RANGE_FIXED)))
    {
        auto k              = minmax_kernel<T>;
        const int64_t tiles = (count / kV + kMmThreads * kMmUnroll - 1) / (kMmThreads * kMmUnroll);
        static thread_local int resident = 0;   // occupancy query once per thread, not per call
        if (resident == 0)
            resident = resident_grid(k, kMmThreads, 0);
        int grid = resident;
        if (tiles < grid)
            grid = tiles < 1 ? 1 : (int) tiles;
        k<<<grid, kMmThreads, 0, stream>>>(in, count, quant_mode, st);
        AB_CUDA_CHECK(cudaGetLastError());
    }
    if (quant_mode == AB_QUANTIZATION_TF_ENHANCED)
    {
        // The histogram is launched with programmatic stream serialization (see the kernel's prologue): measured 0.2-0.65 us
        // per call and 0.1-0.2 ms per ResNet-50 calibration step. AB_PDL=0 falls back to a plain launch.
        static const bool pdl = [] {
            const char* e = getenv("AB_PDL");
            return e == nullptr || e[0] != '0';
        }();
        static const int reverse = [] {
            const char* e = getenv("AB_HIST_REVERSE");
            return (e == nullptr || e[0] != '0') ? 1 : 0;
        }();
        unsigned long long* timer_slot = nullptr;
        if (g_timer_slots != nullptr && g_timer_used < g_timer_capacity)
            timer_slot = g_timer_slots + 3 * g_timer_used++;
        static thread_local bool configured[2] = {false, false};
        const int which                        = sizeof(T) == 4 ? 0 : 1;
        if (!configured[which])
        {
            AB_CUDA_CHECK(cudaFuncSetAttribute(hist_kernel<T, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                               (int) kHistSmem));
            AB_CUDA_CHECK(cudaFuncSetAttribute(hist_kernel<T, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                               (int) kHistSmem));
            configured[which] = true;
        }
        const int64_t tiles = ((count / kV) * 16 + kTileBytes - 1) / kTileBytes;
        int grid            = num_sms();
        if (tiles < grid)
            grid = tiles < 16 ? 16 : (int) tiles;   // >= 16 CTAs: their keeper warps fold 512 parked bins in one pass
        if (pdl)
        {
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim            = dim3((unsigned) grid);
            cfg.blockDim           = dim3(kHistThreads);
            cfg.dynamicSmemBytes   = kHistSmem;
            cfg.stream             = stream;
            cudaLaunchAttribute attr[1];
            attr[0].id                                         = cudaLaunchAttributeProgrammaticStreamSerialization;
            attr[0].val.programmaticStreamSerializationAllowed = 1;
            cfg.attrs                                          = attr;
            cfg.numAttrs                                       = 1;
            AB_CUDA_CHECK(cudaLaunchKernelEx(&cfg, hist_kernel<T, true>, in, count, st, batch_log, reverse,
                                             timer_slot));
        }
        else
        {
            hist_kernel<T, false><<<grid, kHistThreads, kHistSmem, stream>>>(in, count, st, batch_log, reverse,
                                                                             timer_slot);
            AB_CUDA_CHECK(cudaGetLastError());
        }
    }
    return AB_OK;
}

template <typename T>
int launch_multi(const ab_stats_segment* segs, int n, ab_stats_state* states, uint32_t* seg_counts, int flags,
                 cudaStream_t stream)
{
    constexpr int kV = Elem<T>::kPerVec;
    MultiParams p;
    memset(&p, 0, sizeof(p));
    p.num_segments = n;
    int64_t total  = 0;
    for (int s = 0; s < n; ++s)
    {
        p.data[s]        = segs[s].data;
        p.count[s]       = segs[s].count;
        p.state_index[s] = segs[s].state_index;
        p.first_tile[s]  = (int32_t) total;
        total += ((segs[s].count / kV) * 16 + kTileBytes - 1) / kTileBytes;
        if (total > 0x7fffffff)
        {
            set_error("too many tiles in one multi-tensor call");
            return AB_ERR_INVALID;
        }
    }
    p.first_tile[n] = (int32_t) total;
    static thread_local bool configured[2] = {false, false};
    const int which                        = sizeof(T) == 4 ? 0 : 1;
    if (!configured[which])
    {
        AB_CUDA_CHECK(cudaFuncSetAttribute(hist_multi_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int) kHistSmem));
        configured[which] = true;
    }
    unsigned long long* timer_slot = nullptr;
    if (g_timer_slots != nullptr && g_timer_used < g_timer_capacity && !(flags & AB_STATS_MULTI_FOLD_ONLY))
        timer_slot = g_timer_slots + 3 * g_timer_used++;
    int grid = num_sms();
    if (total < grid)
        grid = (int) total;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim            = dim3((unsigned) grid);
    cfg.blockDim           = dim3(kHistThreads);
    cfg.dynamicSmemBytes   = kHistSmem;
    cfg.stream             = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id                                         = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs                                          = attr;
    cfg.numAttrs                                       = 1;
    if (!(flags & AB_STATS_MULTI_FOLD_ONLY))
        AB_CUDA_CHECK(cudaLaunchKernelEx(&cfg, hist_multi_kernel<T>, p, (const ab_stats_state*) states, seg_counts,
                                         timer_slot));
    if (!(flags & AB_STATS_MULTI_HIST_ONLY))
    {
        fold_segments_kernel<<<n, kBins, 0, stream>>>(p, states, seg_counts, (flags & AB_STATS_MULTI_LOG_ONLY) ? 1 : 0);
        AB_CUDA_CHECK(cudaGetLastError());
    }
    return AB_OK;
}

}   // namespace
}   // namespace ab

using namespace ab;

extern "C"
{
size_t ab_stats_state_bytes(void)
{
    return sizeof(ab_stats_state);
}

int64_t ab_debug_hist_timer(unsigned long long* slots, int64_t capacity)
{
    const int64_t used = g_timer_used;
    g_timer_slots      = slots;
    g_timer_capacity   = slots != nullptr ? capacity : 0;
    g_timer_used       = 0;
    return used;
}

int ab_stats_reset(ab_stats_state* states, int64_t count, void* stream)
{
    if (count < 0 || (count > 0 && states == nullptr))
    {
        set_error("null state pointer or negative count");
        return AB_ERR_INVALID;
    }
    if (count == 0)
        return AB_OK;
    if (count > 0x7fffffff)
    {
        set_error("too many states in one call");
        return AB_ERR_INVALID;
    }
    reset_kernel<<<(unsigned) count, 128, 0, (cudaStream_t) stream>>>(states, count);
    AB_CUDA_CHECK(cudaGetLastError());
    return AB_OK;
}

// the percentile and MSE analyzers keep exactly the tf_enhanced statistics (PercentileEncodingAnalyzer.cpp:69-75,
// MseEncodingAnalyzer.cpp:70-76 -> UpdatePdf)
static inline int stats_mode(int quant_mode)
{
    return (quant_mode == AB_QUANTIZATION_PERCENTILE || quant_mode == AB_QUANTIZATION_MSE) ? AB_QUANTIZATION_TF_ENHANCED
                                                                                            : quant_mode;
}

int ab_stats_update(const void* in, int64_t count, int dtype, int quant_mode, ab_stats_state* state,
                    uint32_t* batch_log_entry, int flags, void* stream)
{
    quant_mode = stats_mode(quant_mode);
    if (!check_common(in, count, dtype, quant_mode, state))
        return AB_ERR_INVALID;
    if (dtype == AB_F32)
        return launch_update((const float*) in, count, quant_mode, state, batch_log_entry, flags,
                             (cudaStream_t) stream);
    return launch_update((const __nv_bfloat16*) in, count, quant_mode, state, batch_log_entry, flags,
                         (cudaStream_t) stream);
}

int ab_stats_update_segmented(const void* in, int64_t num_segments, int64_t segment_len, int dtype, int quant_mode,
                              ab_stats_state* states, void* stream)
{
    if (num_segments < 0 || segment_len < 0)
    {
        set_error("negative segment count or length");
        return AB_ERR_INVALID;
    }
    quant_mode = stats_mode(quant_mode);
    if (!check_common(in, num_segments * segment_len, dtype, quant_mode, states))
        return AB_ERR_INVALID;
    if (num_segments == 0)
        return AB_OK;
    cudaStream_t st = (cudaStream_t) stream;
    if (dtype == AB_F32)
    {
        auto k   = segmented_kernel<float>;
        int grid = resident_grid(k, kSegThreads, 0);
        if (num_segments < grid)
            grid = (int) num_segments;
        k<<<grid, kSegThreads, 0, st>>>((const float*) in, num_segments, segment_len, quant_mode, states);
    }
    else
    {
        auto k   = segmented_kernel<__nv_bfloat16>;
        int grid = resident_grid(k, kSegThreads, 0);
        if (num_segments < grid)
            grid = (int) num_segments;
        k<<<grid, kSegThreads, 0, st>>>((const __nv_bfloat16*) in, num_segments, segment_len, quant_mode, states);
    }
    AB_CUDA_CHECK(cudaGetLastError());
    return AB_OK;
}

// reset -> updateStats -> computeEncoding (-> per-channel parameter blocks) for MANY parameter tensors with a handful of
// launches: what every wrapper of a model does with its weights before a training-mode forward, and once per calibration job
// (TEt/.../v1/qc_quantize_op.py:753-798) -- 4 launches per weight through ab_stats_refresh_encodings, 81 weights in
// MobileNet-v2, 54 in ResNet-50.
int ab_stats_refresh_encodings_multi(const ab_refresh_item* items, int num_items, int dtype, int quant_mode,
                                     ab_stats_state* states, int bw, int use_symmetric, int use_strict_symmetric,
                                     int use_unsigned_symmetric, double* enc_out, float* qdq4_out, float* params_out,
                                     void* stream)
{
    if (num_items < 0 || num_items > AB_REFRESH_MULTI_MAX_ITEMS ||
        (num_items > 0 && (items == nullptr || states == nullptr || enc_out == nullptr)))
    {
        set_error("null pointer, or more than %d items", AB_REFRESH_MULTI_MAX_ITEMS);
        return AB_ERR_INVALID;
    }
    if (dtype != AB_F32 && dtype != AB_BF16)
    {
        set_error("unsupported dtype %d", dtype);
        return AB_ERR_INVALID;
    }
    if (num_items == 0)
        return AB_OK;
    const int stats = stats_mode(quant_mode);
    if (stats != AB_QUANTIZATION_TF && stats != AB_QUANTIZATION_TF_ENHANCED)
    {
        set_error("unsupported quantization mode %d", quant_mode);
        return AB_ERR_INVALID;
    }
    constexpr int64_t kLargeSegment = 128 * 1024;   // a single segment this long gets the grid-wide single-tensor kernels
    ItemTable t;
    memset(&t, 0, sizeof(t));
    t.num_items   = num_items;
    int64_t total = 0;
    bool any_small = false;
    for (int i = 0; i < num_items; ++i)
    {
        if (items[i].data == nullptr || items[i].num_segments < 1 || items[i].segment_len < 1 ||
            items[i].first_record != total)
        {
            set_error("item %d: null data, empty tensor, or records that do not follow the previous item's", i);
            return AB_ERR_INVALID;
        }
        t.data[i]        = items[i].data;
        t.segment_len[i] = items[i].segment_len;
        t.first[i]       = (int32_t) total;
        t.skip[i]        = (items[i].num_segments == 1 && items[i].segment_len >= kLargeSegment) ? 1 : 0;
        any_small |= !t.skip[i];
        total += items[i].num_segments;
        if (total > 0x7fffffff)
        {
            set_error("too many records in one call");
            return AB_ERR_INVALID;
        }
    }
    t.first[num_items] = (int32_t) total;
    cudaStream_t st    = (cudaStream_t) stream;
    int rc             = AB_OK;
    if (stats == AB_QUANTIZATION_TF && quant_mode == AB_QUANTIZATION_TF)
    {
        // fused path: one warp per record (tf_refresh_kernel); a single huge segment still takes the grid-wide kernels
        if (any_small)
        {
            const int64_t ctas = (total + 7) / 8;
            const int grid     = (int) (ctas < (int64_t) 16 * num_sms() ? ctas : (int64_t) 16 * num_sms());
            if (dtype == AB_F32)
                tf_refresh_kernel<float><<<grid, 256, 0, st>>>(t, states, bw, use_symmetric, use_strict_symmetric,
                                                               use_unsigned_symmetric, enc_out, qdq4_out);
            else
                tf_refresh_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(t, states, bw, use_symmetric, use_strict_symmetric,
                                                                       use_unsigned_symmetric, enc_out, qdq4_out);
            AB_CUDA_CHECK(cudaGetLastError());
        }
        for (int i = 0; i < num_items; ++i)
            if (t.skip[i])
            {
                rc = ab_stats_refresh_encodings(items[i].data, 1, items[i].segment_len, dtype, quant_mode, states + t.first[i],
                                                bw, use_symmetric, use_strict_symmetric, use_unsigned_symmetric,
                                                enc_out + 5 * (int64_t) t.first[i],
                                                qdq4_out ? qdq4_out + 4 * (int64_t) t.first[i] : nullptr, nullptr, stream);
                if (rc != AB_OK)
                    return rc;
            }
        if (params_out != nullptr)
        {
            per_channel_params_multi_kernel<<<(unsigned) ((total + 127) / 128), 128, 0, st>>>(t, enc_out, bw, params_out);
            AB_CUDA_CHECK(cudaGetLastError());
        }
        return AB_OK;
    }
    rc = ab_stats_reset(states, total, stream);
    if (rc != AB_OK)
        return rc;
    if (any_small)
    {
        if (dtype == AB_F32)
        {
            auto k   = segmented_multi_kernel<float>;
            int grid = resident_grid(k, kSegThreads, 0);
            k<<<total < grid ? (int) total : grid, kSegThreads, 0, st>>>(t, stats, states);
        }
        else
        {
            auto k   = segmented_multi_kernel<__nv_bfloat16>;
            int grid = resident_grid(k, kSegThreads, 0);
            k<<<total < grid ? (int) total : grid, kSegThreads, 0, st>>>(t, stats, states);
        }
        AB_CUDA_CHECK(cudaGetLastError());
    }
    for (int i = 0; i < num_items; ++i)
        if (t.skip[i])
        {
            rc = ab_stats_update(items[i].data, items[i].segment_len, dtype, stats, states + t.first[i], nullptr, 0, stream);
            if (rc != AB_OK)
                return rc;
        }
    rc = quant_mode == AB_QUANTIZATION_PERCENTILE
             ? AB_ERR_UNSUPPORTED
             : ab_compute_encodings(states, total, quant_mode, bw, use_symmetric, use_strict_symmetric, use_unsigned_symmetric,
                                    enc_out, qdq4_out, stream);
    if (rc != AB_OK)
    {
        if (rc == AB_ERR_UNSUPPORTED)
            set_error("the percentile scheme is refreshed per tensor (it needs its percentile value)");
        return rc;
    }
    if (params_out != nullptr)
    {
        per_channel_params_multi_kernel<<<(unsigned) ((total + 127) / 128), 128, 0, st>>>(t, enc_out, bw, params_out);
        AB_CUDA_CHECK(cudaGetLastError());
    }
    return AB_OK;
}

int ab_stats_fold_log(ab_stats_state* states, int64_t count, const uint32_t* log, const int64_t* entry_rows,
                      const int64_t* record_begin, void* stream)
{
    if (count < 0 || (count > 0 && (states == nullptr || log == nullptr || entry_rows == nullptr || record_begin == nullptr)))
    {
        set_error("null pointer or negative count");
        return AB_ERR_INVALID;
    }
    if (count == 0)
        return AB_OK;
    fold_log_kernel<<<(unsigned) count, kBins, 0, (cudaStream_t) stream>>>(states, count, log, entry_rows, record_begin);
    AB_CUDA_CHECK(cudaGetLastError());
    return AB_OK;
}

int ab_stats_update_multi(const ab_stats_segment* segments, int num_segments, int dtype, ab_stats_state* states,
                          uint32_t* seg_counts, int flags, void* stream)
{
    if (num_segments < 0 || num_segments > AB_STATS_MULTI_MAX_SEGMENTS ||
        (num_segments > 0 && (segments == nullptr || states == nullptr || seg_counts == nullptr)))
    {
        set_error("null pointer, or more than %d segments", AB_STATS_MULTI_MAX_SEGMENTS);
        return AB_ERR_INVALID;
    }
    if (dtype != AB_F32 && dtype != AB_BF16)
    {
        set_error("unsupported dtype %d", dtype);
        return AB_ERR_INVALID;
    }
    const int64_t per_vec = dtype == AB_F32 ? 4 : 8;
    for (int s = 0; s < num_segments; ++s)
        if (segments[s].data == nullptr || segments[s].count < per_vec || segments[s].state_index < 0 ||
            (reinterpret_cast<uintptr_t>(segments[s].data) & 15u) != 0)
        {
            set_error("segment %d: needs a 16-byte aligned device pointer, at least one 128-bit vector of data and a "
                      "non-negative record index",
                      s);
            return AB_ERR_INVALID;
        }
    if (num_segments == 0)
        return AB_OK;
    if (dtype == AB_F32)
        return launch_multi<float>(segments, num_segments, states, seg_counts, flags, (cudaStream_t) stream);
    return launch_multi<__nv_bfloat16>(segments, num_segments, states, seg_counts, flags, (cudaStream_t) stream);
}

int ab_stats_init_range(ab_stats_state* states, int64_t count, const float* minmax, void* stream)
{
    if (count < 0 || (count > 0 && (states == nullptr || minmax == nullptr)))
    {
        set_error("null pointer or negative count");
        return AB_ERR_INVALID;
    }
    if (count == 0)
        return AB_OK;
    init_range_kernel<<<(unsigned) ((count + 127) / 128), 128, 0, (cudaStream_t) stream>>>(states, count, minmax);
    AB_CUDA_CHECK(cudaGetLastError());
    return AB_OK;
}

int ab_stats_fold_batches(ab_stats_state* states, int64_t count, const uint32_t* batch_log,
                          const int64_t* batch_offsets, int64_t num_batches, void* stream)
{
    if (count < 0 || num_batches < 0 ||
        (count > 0 && (states == nullptr || (num_batches > 0 && (batch_log == nullptr || batch_offsets == nullptr)))))
    {
        set_error("null pointer or negative count");
        return AB_ERR_INVALID;
    }
    if (count == 0)
        return AB_OK;
    fold_batches_kernel<<<(unsigned) count, kBins, 0, (cudaStream_t) stream>>>(states, count, batch_log,
                                                                               batch_offsets, num_batches);
    AB_CUDA_CHECK(cudaGetLastError());
    return AB_OK;
}
}   // extern "C"
