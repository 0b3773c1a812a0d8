// The entropy scheme (EntropyEncodingAnalyzer, DlQ/src/EntropyEncodingAnalyzer.cpp; QuantizationMode::QUANTIZATION_ENTROPY,
// reachable through libpymo's TensorQuantizer / EncodingAnalyzerForPython): a 512-bin histogram whose range GROWS with the
// data (older counts are redistributed when a batch exceeds it) and a KL-divergence search over shrinking windows of it.
//
// In the reference even the GPU build copies the tensor to the host and bins it there, one element at a time
// (math_functions.cpp:440-462, 476-560). Here one updateStats call is three launches on the caller's stream, no host sync:
//   1. min / max of the tensor (warp shuffles, one ordered-int atomic pair per CTA)                     -- 1 s bytes / element
//   2. one thread: the range bookkeeping (all-zero batch -> no-op; first batch fixes the range; a wider batch rescales the
//      512 counts exactly as the reference's double arithmetic does; entropy_math.h prepare_update)
//   3. binning: bin = min(size_t((x - min) / width), 511) in the reference's float arithmetic, counts privatised per lane in
//      shared memory (word bin * 32 + lane: the bank is the lane, no conflicts whatever the distribution), one 64-bit global
//      atomic per non-empty bin and CTA, the last CTA adds the batch's counts to the running histogram (doubles holding
//      integers: exact) and re-arms the scratch                                                         -- 1 s bytes / element
// computeEncoding reads the histogram back (4 KB) and runs the search on the host: see entropy_math.h for why.
//
// State: an ab_stats_state record initialised by ab_stats_reset, fields reused as follows -- pdf[] = the histogram,
// x_left0 / bucket_size_d = its min / max, initialized = "histogram exists", iterations, stats_updated, bucket_size /
// pdf_offset = the float bin width / float min of the current call, pending = "this call is a no-op", hist[][] = 512 64-bit
// batch counts, batch_min_bits / batch_max_bits / ticket = scratch (armed between calls).
#include <vector>

#include "common.cuh"
#include "entropy_math.h"

namespace ab
{
namespace
{
constexpr int kThreads    = 256;
constexpr int kUnroll     = 4;
constexpr int kPosInfBits = 0x7f800000;
constexpr int kNegInfBits = (int) 0xff800000 ^ 0x7fffffff;   // float_to_ordered(-inf)

// GetMin_cpu / GetMax_cpu (math_functions.cpp:327-347): val = +-DBL_MAX narrowed to float (= +-inf); std::min / std::max
// ignore a NaN element
template <typename T>
__global__ void __launch_bounds__(kThreads) entropy_minmax_kernel(const T* __restrict__ in, int64_t count, ab_stats_state* st)
{
    constexpr int kV      = Elem<T>::kPerVec;
    const int64_t num_vec = (reinterpret_cast<uintptr_t>(in) & 15u) == 0 ? count / kV : 0;
    float lo = __int_as_float(0x7f800000), hi = __int_as_float(0xff800000);
    const int64_t stride = (int64_t) gridDim.x * kThreads;
    for (int64_t v = (int64_t) blockIdx.x * kThreads + threadIdx.x; v < num_vec; v += stride)
    {
        float f[kV];
        Elem<T>::unpack(ldg_stream(reinterpret_cast<const uint4*>(in) + v), f);
#pragma unroll
        for (int k = 0; k < kV; ++k)
        {
            lo = (f[k] < lo) ? f[k] : lo;
            hi = (hi < f[k]) ? f[k] : hi;
        }
    }
    for (int64_t i = num_vec * kV + (int64_t) blockIdx.x * kThreads + threadIdx.x; i < count; i += stride)
    {
        const float x = Elem<T>::load(in + i);
        lo            = (x < lo) ? x : lo;
        hi            = (hi < x) ? x : hi;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
    {
        lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));   // no NaN can be in lo / hi
        hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    __shared__ float s_lo[kThreads / 32], s_hi[kThreads / 32];
    if ((threadIdx.x & 31) == 0)
        s_lo[threadIdx.x >> 5] = lo, s_hi[threadIdx.x >> 5] = hi;
    __syncthreads();
    if (threadIdx.x == 0)
    {
        for (int w = 1; w < kThreads / 32; ++w)
            lo = fminf(lo, s_lo[w]), hi = fmaxf(hi, s_hi[w]);
        atomicMin(&st->batch_min_bits, float_to_ordered(lo));
        atomicMax(&st->batch_max_bits, float_to_ordered(hi));
    }
}

__global__ void entropy_prepare_kernel(ab_stats_state* st)
{
    const float lo = ordered_to_float(st->batch_min_bits), hi = ordered_to_float(st->batch_max_bits);
    st->batch_min_bits = kPosInfBits;
    st->batch_max_bits = kNegInfBits;
    st->stats_updated  = 1;
    float width = 0.0f, mn = 0.0f;
    int init    = st->initialized;
    const bool go = ent::prepare_update(st->pdf, st->x_left0, st->bucket_size_d, init, lo, hi, &width, &mn);
    st->initialized = init;
    st->bucket_size = width;
    st->pdf_offset  = mn;
    st->pending     = go ? 0 : 1;
}

template <typename T>
__global__ void __launch_bounds__(kThreads) entropy_bin_kernel(const T* __restrict__ in, int64_t count, ab_stats_state* st)
{
    extern __shared__ uint32_t s_bins[];   // [512][32]
    __shared__ bool s_last;
    if (st->pending)                       // all-zero tensor: the call leaves the histogram alone (uniform for the grid)
        return;
    constexpr int kV  = Elem<T>::kPerVec;
    const float width = st->bucket_size, mn = st->pdf_offset;
    // (x - min) / width with the hoisted reciprocal (common.cuh: the very FFMA sequence div.rn.f32 executes, hence the same
    // bits; a numerator too small for it perturbs only quotients far below 1, which truncate to bin 0 either way), and the
    // float -> size_t cast of a quotient in [0, 2^22) as one round-toward-zero add: no MUFU, no F2I in the loop.
    const Divisor dv  = make_divisor(width);
    const bool fast   = dv.fast && width > 0.0f;
    auto bin_of       = [&](float x) -> int {
        const float d = __fsub_rn(x, mn);
        if (!fast)
            return width == 0.0f ? 0 : ent::bin_of_quotient(__fdiv_rn(d, width));
        const float q = div_fast(d, dv);
        if (q >= 0.0f && q < 4194304.0f)
            return min(__float_as_int(__fadd_rz(q, 8388608.0f)) & 0x7fffff, ent::kBins - 1);
        return ent::bin_of_quotient(q);
    };
    for (int i = threadIdx.x; i < ent::kBins * 32; i += kThreads)
        s_bins[i] = 0;
    __syncthreads();
    const int lane        = threadIdx.x & 31;
    const int64_t num_vec = (reinterpret_cast<uintptr_t>(in) & 15u) == 0 ? count / kV : 0;
    const int64_t stride  = (int64_t) gridDim.x * kThreads;
    for (int64_t v0 = (int64_t) blockIdx.x * kThreads + threadIdx.x; v0 < num_vec; v0 += stride * kUnroll)
    {
        uint4 raw[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u)
            if (v0 + u * stride < num_vec)
                raw[u] = ldg_stream(reinterpret_cast<const uint4*>(in) + v0 + u * stride);
#pragma unroll
        for (int u = 0; u < kUnroll; ++u)
            if (v0 + u * stride < num_vec)
            {
                float f[kV];
                Elem<T>::unpack(raw[u], f);
#pragma unroll
                for (int k = 0; k < kV; ++k)
                    atomicAdd(&s_bins[bin_of(f[k]) * 32 + lane], 1u);
            }
    }
    for (int64_t i = num_vec * kV + (int64_t) blockIdx.x * kThreads + threadIdx.x; i < count; i += stride)
        atomicAdd(&s_bins[bin_of(Elem<T>::load(in + i)) * 32 + lane], 1u);
    __syncthreads();
    unsigned long long* batch = reinterpret_cast<unsigned long long*>(&st->hist[0][0]);   // 512 x 64 bit
    for (int b = threadIdx.x; b < ent::kBins; b += kThreads)
    {
        uint32_t sum = 0;
#pragma unroll
        for (int l = 0; l < 32; ++l)
            sum += s_bins[b * 32 + ((l + b) & 31)];   // rotated: the threads of a warp read 32 different banks
        if (sum != 0)
            atomicAdd(batch + b, (unsigned long long) sum);
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0)
        s_last = atomicAdd(&st->ticket, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!s_last)
        return;
    __threadfence();
    for (int b = threadIdx.x; b < ent::kBins; b += kThreads)
    {
        const unsigned long long c = __ldcg(batch + b);
        st->pdf[b] += (double) c;   // tpp.histogram[newBin] += 1, c times: integers below 2^53, exact
        batch[b] = 0;
    }
    if (threadIdx.x == 0)
    {
        st->iterations += 1;
        st->ticket = 0;
    }
}

template <typename T>
int launch(const T* in, int64_t count, ab_stats_state* st, cudaStream_t stream)
{
    constexpr int kV = Elem<T>::kPerVec;
    const int sms    = num_sms();
    if (count > 0)
    {
        const int64_t tiles = (count / kV + kThreads - 1) / kThreads + 1;
        int grid            = 4 * sms;
        if (tiles < grid)
            grid = (int) tiles;
        entropy_minmax_kernel<T><<<grid, kThreads, 0, stream>>>(in, count, st);
        AB_CUDA_CHECK(cudaGetLastError());
    }
    entropy_prepare_kernel<<<1, 1, 0, stream>>>(st);
    AB_CUDA_CHECK(cudaGetLastError());
    if (count > 0)
    {
        constexpr size_t kSmem = ent::kBins * 32 * sizeof(uint32_t);   // 64 KB
        static bool configured[2] = {false, false};
        const int which = sizeof(T) == 2;
        if (!configured[which])
        {
            AB_CUDA_CHECK(cudaFuncSetAttribute(entropy_bin_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) kSmem));
            configured[which] = true;
        }
        const int64_t tiles = (count / kV + (int64_t) kThreads * kUnroll - 1) / ((int64_t) kThreads * kUnroll) + 1;
        int grid            = 3 * sms;   // 3 x 64 KB of shared memory per SM
        if (tiles < grid)
            grid = (int) tiles;
        entropy_bin_kernel<T><<<grid, kThreads, kSmem, stream>>>(in, count, st);
        AB_CUDA_CHECK(cudaGetLastError());
    }
    return AB_OK;
}

struct HostCopy
{
    double hist[ent::kBins];
    double mn, mx;
    int initialized, stats_updated, iterations;
};

int read_back(const ab_stats_state* st, HostCopy& h, cudaStream_t stream)
{
    static_assert(offsetof(ab_stats_state, pdf) == 0, "pdf first");
    struct Tail
    {
        double x_left0, bucket_size_d;
    } tail;
    int32_t flags[3];
    AB_CUDA_CHECK(cudaMemcpyAsync(h.hist, st->pdf, sizeof(h.hist), cudaMemcpyDeviceToHost, stream));
    AB_CUDA_CHECK(cudaMemcpyAsync(&tail, &st->x_left0, sizeof(tail), cudaMemcpyDeviceToHost, stream));
    AB_CUDA_CHECK(cudaMemcpyAsync(flags, &st->initialized, sizeof(flags), cudaMemcpyDeviceToHost, stream));
    AB_CUDA_CHECK(cudaStreamSynchronize(stream));
    static_assert(offsetof(ab_stats_state, stats_updated) == offsetof(ab_stats_state, initialized) + 4 &&
                      offsetof(ab_stats_state, iterations) == offsetof(ab_stats_state, initialized) + 8,
                  "initialized, stats_updated, iterations are consecutive");
    h.mn = tail.x_left0, h.mx = tail.bucket_size_d;
    h.initialized = flags[0], h.stats_updated = flags[1], h.iterations = flags[2];
    return AB_OK;
}

}   // namespace
}   // namespace ab

using namespace ab;

extern "C" int ab_entropy_update(const void* in, int64_t count, int dtype, ab_stats_state* state, void* stream)
{
    if (count < 0 || (count > 0 && in == nullptr) || state == nullptr)
    {
        set_error("null pointer or negative count");
        return AB_ERR_INVALID;
    }
    if (dtype == AB_F32)
        return launch((const float*) in, count, state, (cudaStream_t) stream);
    if (dtype == AB_BF16)
        return launch((const __nv_bfloat16*) in, count, state, (cudaStream_t) stream);
    set_error("unsupported dtype %d", dtype);
    return AB_ERR_INVALID;
}

extern "C" int ab_entropy_compute_encoding(const ab_stats_state* state, int bw, int use_symmetric, int use_strict_symmetric,
                                           int use_unsigned_symmetric, ab_encoding* out, void* stream)
{
    if (state == nullptr || out == nullptr)
    {
        set_error("null pointer");
        return AB_ERR_INVALID;
    }
    if (bw < 1 || bw > 32)
    {
        set_error("Invalid bitwidth: %d", bw);
        return AB_ERR_INVALID;
    }
    HostCopy h;
    const int rc = read_back(state, h, (cudaStream_t) stream);
    if (rc != AB_OK)
        return rc;
    ent::compute_encoding(h.hist, h.mn, h.mx, h.initialized, h.stats_updated, bw, use_symmetric != 0, use_strict_symmetric != 0,
                          use_unsigned_symmetric != 0, *out);
    return AB_OK;
}

extern "C" int ab_entropy_histogram(const ab_stats_state* state, double* hist512, double* min_max, int* info, void* stream)
{
    if (state == nullptr || hist512 == nullptr || min_max == nullptr)
    {
        set_error("null pointer");
        return AB_ERR_INVALID;
    }
    HostCopy h;
    const int rc = read_back(state, h, (cudaStream_t) stream);
    if (rc != AB_OK)
        return rc;
    for (int i = 0; i < ent::kBins; ++i)
        hist512[i] = h.hist[i];
    min_max[0] = h.mn, min_max[1] = h.mx;
    if (info != nullptr)
        info[0] = h.initialized, info[1] = h.stats_updated, info[2] = h.iterations;
    return AB_OK;
}
