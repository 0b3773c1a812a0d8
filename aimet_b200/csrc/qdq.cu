// Job 1 -- fused quantize-dequantize (QDQ), quantize-only and STE-backward kernels for sm_100a.
//
// All of these are pure streaming kernels bound by HBM bandwidth (8 B/element fp32 QDQ, 4 B/element bf16,
// 12 / 6 B/element for the backward). Design:
//   * 128-bit coalesced accesses (ld.global.nc.L1::no_allocate.v4 / st.global.L1::no_allocate.v4), kUnroll
//     independent loads in flight per thread before any arithmetic, so one CTA keeps 16 KB of reads outstanding;
//   * grid = min(tiles, resident CTAs per SM x 148 SMs), grid-stride over tiles: a whole number of waves;
//   * exact IEEE arithmetic (see common.cuh) -- the ~25 ALU instructions per element stay well under the
//     issue budget of a memory-bound kernel;
//   * per-channel: the CTA's tile covers a contiguous range of channels whose {min,max,delta,offset} are staged in
//     shared memory as float4 once per tile; the channel index is advanced incrementally, one integer division per
//     128-bit vector rather than one per element.
//
// Reference semantics: DlQ/src/trim_functions.cpp:140-218, 697-709 (CPU path = parity target);
// the reference's own GPU kernels are DlQ/src/trim_functions.cu:46-92.
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include "common.cuh"
#include "encoding_math.h"

namespace ab
{

#ifndef AB_QDQ_THREADS
#define AB_QDQ_THREADS 256
#endif
#ifndef AB_QDQ_UNROLL
#define AB_QDQ_UNROLL 4
#endif
constexpr int kThreads = AB_QDQ_THREADS;
constexpr int kUnroll  = AB_QDQ_UNROLL;   // 128-bit vectors per thread per tile
#ifndef AB_QDQ_MIN_BLOCKS
#define AB_QDQ_MIN_BLOCKS 1               // tuning hook: resident CTAs per SM the streaming kernels are compiled for
#endif

enum class Op
{
    kQdq,
    kQuantize
};

struct TensorArgs
{
    Enc4 enc;            // used when enc_dev == nullptr
    const float* enc_dev;   // optional device-resident {min,max,delta,offset}
    float shift;         // quantize-only: subtracted from the grid value
    uint64_t seed;
    int reverse;         // walk the tiles from the end of the tensor backwards (see per_tensor_body)
};

template <Op kOp, bool kStochastic>
__device__ __forceinline__ float apply(float x, const Enc4& e, float shift, uint64_t seed, uint64_t idx)
{
    const float q = quantize_value<kStochastic, kOp == Op::kQuantize>(x, e, seed, idx);
    if (kOp == Op::kQdq)
        return dequantize_value(q, e);
    return __fsub_rn(q, shift);   // out[i] -= shift (DlQ/src/trim_functions.cpp:216)
}

// ---------------------------------------------------------------------------------------------------------------
// per-tensor QDQ / quantize-only.  `in`/`out` are 16-byte aligned here (otherwise per_tensor_scalar_kernel runs).
// ---------------------------------------------------------------------------------------------------------------
template <typename T, Op kOp, bool kStochastic, bool kFast, bool kPos = false>
__device__ __forceinline__ void per_tensor_body(const T* __restrict__ in, T* __restrict__ out, int64_t count,
                                                const Enc4& e, const Divisor& dv, float shift, uint64_t seed,
                                                int reverse)
{
    const float m_off = __fsub_rn(12582912.0f, e.offset);   // kPos only
    constexpr int kV        = Elem<T>::kPerVec;
    const int64_t num_vec   = count / kV;
    const int64_t num_tiles = (num_vec + kThreads * kUnroll - 1) / (kThreads * kUnroll);

    // Tile order. The input was usually written by the kernel just before this one and the output is usually read from
    // its start by the kernel just after it. What the L2 holds of a tensor larger than itself is the part written last.
    // Walking backwards therefore reads the producer's tail from L2 before it is evicted, and leaves the HEAD of the
    // output as the most recent lines for the consumer; walking forwards does neither. Element-wise: order is free.
    for (int64_t t = blockIdx.x; t < num_tiles; t += gridDim.x)
    {
        const int64_t tile = reverse ? num_tiles - 1 - t : t;
        const int64_t v0 = tile * (kThreads * kUnroll) + threadIdx.x;
        uint4 raw[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u)
        {
            const int64_t v = v0 + (int64_t) u * kThreads;
            if (v < num_vec)
                raw[u] = ldg_stream(reinterpret_cast<const uint4*>(in) + v);
        }
#pragma unroll
        for (int u = 0; u < kUnroll; ++u)
        {
            const int64_t v = v0 + (int64_t) u * kThreads;
            if (v < num_vec)
            {
                float f[kV];
                Elem<T>::unpack(raw[u], f);
#pragma unroll
                for (int k = 0; k < kV; ++k)
                    f[k] = !kFast               ? apply<kOp, kStochastic>(f[k], e, shift, seed, (uint64_t) (v * kV + k))
                           : kOp == Op::kQdq    ? (kPos ? qdq_fast_pos(f[k], e, dv, m_off) : qdq_fast(f[k], e, dv))
                                                : __fsub_rn(quantize_fast(f[k], e, dv), shift);
                stg_stream(reinterpret_cast<uint4*>(out) + v, Elem<T>::pack(f));
            }
        }
    }
    // scalar tail (< one vector), handled by the first threads of block 0
    if (blockIdx.x == 0)
    {
        const int64_t i = num_vec * kV + threadIdx.x;
        if (i < count)
            Elem<T>::store(out + i, apply<kOp, kStochastic>(Elem<T>::load(in + i), e, shift, seed, (uint64_t) i));
    }
}

template <typename T, Op kOp, bool kStochastic>
__global__ void __launch_bounds__(kThreads, AB_QDQ_MIN_BLOCKS) per_tensor_kernel(const T* __restrict__ in, T* __restrict__ out,
                                                              int64_t count, TensorArgs args)
{
    Enc4 e = args.enc;
    if (args.enc_dev != nullptr)
    {
        const float4 p = *reinterpret_cast<const float4*>(args.enc_dev);
        e              = Enc4 {p.x, p.y, p.z, p.w};
    }
    const Divisor dv = make_divisor(e.delta);
    // Nearest rounding with an ordinary grid takes the XU-free path; the choice is uniform over the launch.
    if (!kStochastic && qdq_fast_ok(e, dv))
    {
        if (kOp == Op::kQdq && qdq_pos_ok(e, dv))
            per_tensor_body<T, kOp, kStochastic, true, true>(in, out, count, e, dv, args.shift, args.seed, args.reverse);
        else
            per_tensor_body<T, kOp, kStochastic, true>(in, out, count, e, dv, args.shift, args.seed, args.reverse);
    }
    else
        per_tensor_body<T, kOp, kStochastic, false>(in, out, count, e, dv, args.shift, args.seed, args.reverse);
}

// element-wise variant for tensors that are not 16-byte aligned
template <typename T, Op kOp, bool kStochastic>
__global__ void __launch_bounds__(kThreads) per_tensor_scalar_kernel(const T* __restrict__ in, T* __restrict__ out,
                                                                     int64_t count, TensorArgs args)
{
    Enc4 e = args.enc;
    if (args.enc_dev != nullptr)
    {
        const float4 p = *reinterpret_cast<const float4*>(args.enc_dev);
        e              = Enc4 {p.x, p.y, p.z, p.w};
    }
    const int64_t stride = (int64_t) gridDim.x * kThreads;
    for (int64_t i = (int64_t) blockIdx.x * kThreads + threadIdx.x; i < count; i += stride)
        Elem<T>::store(out + i, apply<kOp, kStochastic>(Elem<T>::load(in + i), e, args.shift, args.seed, (uint64_t) i));
}

// ---------------------------------------------------------------------------------------------------------------
// per-channel QDQ.  channel(i) = (i / per_channel) % num_channel   (DlQ/src/trim_functions.cpp:703)
// ---------------------------------------------------------------------------------------------------------------
constexpr int kSmemChannels = 1024;   // staged {min,max,delta,offset} + reciprocal: 20 KB of shared memory

struct ChannelArgs
{
    const float* params;   // device: [min | max | delta | offset], each num_channel long
    int64_t num_channel;
    int64_t per_channel;
    uint64_t seed;
    uint32_t div_mul, div_shift;   // n / per_channel == umulhi(n, div_mul) >> div_shift for n < 2^31 (per_channel > 1)
};

__device__ __forceinline__ Enc4 load_channel(const float* params, int64_t num_channel, int64_t c)
{
    return Enc4 {__ldg(params + c), __ldg(params + num_channel + c), __ldg(params + 2 * num_channel + c),
                 __ldg(params + 3 * num_channel + c)};
}

// Fast kernel (nearest rounding, channel length < 2^31 - 2^16). Any element count: only the tile base is 64-bit, and its
// one 64-bit division per tile is done by a single thread while the tile's data is already in flight.
//   * the tile's channels are staged once, with their refined reciprocal, in shared memory;
//   * inside the tile all index arithmetic is 32-bit and one multiply-high replaces the division per 128-bit vector;
//   * a vector that lies inside one channel (the common case) does a single 16-byte shared load and runs the straight-line
//     fast QDQ; only vectors that straddle a channel boundary step element by element.
template <typename T>
// (5 resident CTAs per SM = 48 registers: 0.87 against 0.78 of the HBM roofline at 64 MB fp32 with the compiler's own 60)
__global__ void __launch_bounds__(kThreads, AB_QDQ_MIN_BLOCKS > 5 ? AB_QDQ_MIN_BLOCKS : 5) per_channel_fast_kernel(const T* __restrict__ in, T* __restrict__ out,
                                                                    int64_t count, ChannelArgs args)
{
    constexpr int kV               = Elem<T>::kPerVec;
    constexpr uint32_t kVecPerTile = kThreads * kUnroll;
    constexpr uint32_t kTileLen    = kVecPerTile * kV;
    __shared__ float4 s_enc[kSmemChannels];
    __shared__ float s_rcp[kSmemChannels];
    __shared__ uint32_t s_rem0, s_c0;

    const int64_t num_vec   = count / kV;
    const int64_t num_tiles = (count + kTileLen - 1) / kTileLen;
    const uint32_t C        = (uint32_t) args.num_channel;
    const uint32_t L        = (uint32_t) args.per_channel;
    auto div_l              = [&](uint32_t n) { return L == 1 ? n : (__umulhi(n, args.div_mul) >> args.div_shift); };

    for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x)
    {
        const int64_t e0       = tile * kTileLen;
        const uint32_t tile_n  = (uint32_t) min((int64_t) kTileLen, count - e0);
        // the tile's data is requested first: the loads do not depend on the channel parameters, so their latency
        // overlaps the parameter staging below instead of following it
        const int64_t v0 = tile * kVecPerTile + threadIdx.x;
        uint4 raw[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u)
        {
            const int64_t v = v0 + (int64_t) u * kThreads;
            if (v < num_vec)
                raw[u] = ldg_stream(reinterpret_cast<const uint4*>(in) + v);
        }
        __syncthreads();   // previous tile's readers are done with the staged channels and s_rem0 / s_c0
        if (threadIdx.x == 0)
        {
            const int64_t g0 = e0 / (int64_t) L;      // un-wrapped channel counter of the tile's first element
            s_rem0           = (uint32_t) (e0 - g0 * (int64_t) L);
            s_c0             = (uint32_t) (g0 % (int64_t) C);
        }
        __syncthreads();
        const uint32_t rem0 = s_rem0, c0 = s_c0;
        const uint32_t span = div_l(rem0 + tile_n - 1) + 1;   // channels the tile touches (rem0 + tile_n < 2^31)
        bool ok             = span <= kSmemChannels;
        bool pos            = true;
        if (ok)
            for (uint32_t j = threadIdx.x; j < span; j += kThreads)
            {
                const Enc4 e     = load_channel(args.params, C, (c0 + j) % C);
                const Divisor dv = make_divisor(e.delta);
                s_enc[j]         = make_float4(e.mn, e.mx, e.delta, e.offset);
                s_rcp[j]         = dv.y;
                ok               = ok && qdq_fast_ok(e, dv);
                pos              = pos && qdq_pos_ok(e, dv);
            }
        const bool fast = __syncthreads_and(ok);   // every channel of the tile is staged and takes the fast arithmetic
        const bool fpos = __syncthreads_and(ok && pos);   // ... and none of them can produce a position below -0.5

#pragma unroll
        for (int u = 0; u < kUnroll; ++u)
        {
            const int64_t v = v0 + (int64_t) u * kThreads;
            if (v >= num_vec)
                continue;
            float f[kV];
            Elem<T>::unpack(raw[u], f);
            const uint32_t off = (threadIdx.x + u * kThreads) * kV + rem0;   // position counted from the first channel's start
            uint32_t j         = div_l(off);
            uint32_t rem       = off - j * L;
            if (fast)
            {
                float4 p   = s_enc[j];
                Enc4 e     = Enc4 {p.x, p.y, p.z, p.w};
                Divisor dv = Divisor {p.z, s_rcp[j], true};
                if (rem + kV <= L)
                {
                    if (fpos)
                    {
                        const float m_off = __fsub_rn(12582912.0f, e.offset);
#pragma unroll
                        for (int k = 0; k < kV; ++k)
                            f[k] = qdq_fast_pos(f[k], e, dv, m_off);
                    }
                    else
                    {
#pragma unroll
                        for (int k = 0; k < kV; ++k)
                            f[k] = qdq_fast(f[k], e, dv);
                    }
                }
                else
                {
#pragma unroll
                    for (int k = 0; k < kV; ++k)
                    {
                        f[k] = qdq_fast(f[k], e, dv);
                        if (++rem == L && k + 1 < kV)
                        {
                            rem = 0;
                            ++j;
                            p  = s_enc[j];
                            e  = Enc4 {p.x, p.y, p.z, p.w};
                            dv = Divisor {p.z, s_rcp[j], true};
                        }
                    }
                }
            }
            else
            {
                uint32_t c = (uint32_t) (((uint64_t) c0 + j) % C);
#pragma unroll
                for (int k = 0; k < kV; ++k)
                {
                    const Enc4 e = load_channel(args.params, C, c);
                    f[k]         = dequantize_value(quantize_value<false>(f[k], e, 0, 0), e);
                    if (++rem == L)
                    {
                        rem = 0;
                        c   = (c + 1 == C) ? 0 : c + 1;
                    }
                }
            }
            stg_stream(reinterpret_cast<uint4*>(out) + v, Elem<T>::pack(f));
        }
        // scalar tail of the whole tensor (only the last tile can have one)
        if (e0 + tile_n == count)
        {
            const int64_t i = num_vec * kV + threadIdx.x;
            if (i < count)
            {
                const Enc4 e = load_channel(args.params, C, (int64_t) ((i / (int64_t) L) % (int64_t) C));
                Elem<T>::store(out + i, dequantize_value(quantize_value<false>(Elem<T>::load(in + i), e, 0, 0), e));
            }
        }
    }
}

// General kernel: 64-bit indexing, either rounding mode.
template <typename T, bool kStochastic>
__global__ void __launch_bounds__(kThreads) per_channel_kernel(const T* __restrict__ in, T* __restrict__ out,
                                                               int64_t count, ChannelArgs args)
{
    constexpr int kV      = Elem<T>::kPerVec;
    const int64_t num_vec = count / kV;
    const int64_t C       = args.num_channel;
    const int64_t L       = args.per_channel;
    const int64_t stride  = (int64_t) gridDim.x * kThreads;
    for (int64_t v = (int64_t) blockIdx.x * kThreads + threadIdx.x; v < num_vec; v += stride)
    {
        float f[kV];
        Elem<T>::unpack(ldg_stream(reinterpret_cast<const uint4*>(in) + v), f);
        const int64_t i0 = v * kV;
        const int64_t g  = i0 / L;
        int64_t rem      = i0 - g * L;
        int64_t c        = g % C;
        Enc4 e           = load_channel(args.params, C, c);
#pragma unroll
        for (int k = 0; k < kV; ++k)
        {
            f[k] = dequantize_value(quantize_value<kStochastic>(f[k], e, args.seed, (uint64_t) (i0 + k)), e);
            if (++rem == L)
            {
                rem = 0;
                c   = (c + 1 == C) ? 0 : c + 1;
                e   = load_channel(args.params, C, c);
            }
        }
        stg_stream(reinterpret_cast<uint4*>(out) + v, Elem<T>::pack(f));
    }
    if (blockIdx.x == 0)
    {
        const int64_t i = num_vec * kV + threadIdx.x;
        if (i < count)
        {
            const Enc4 e = load_channel(args.params, C, (i / L) % C);
            Elem<T>::store(out + i, dequantize_value(quantize_value<kStochastic>(Elem<T>::load(in + i), e, args.seed,
                                                                                  (uint64_t) i),
                                                     e));
        }
    }
}

// per-channel parameter preparation on the device (one thread per channel), see em::per_channel_param
__global__ void per_channel_params_kernel(const double* __restrict__ enc5, int64_t num_channel, int bw,
                                          float* __restrict__ params)
{
    const int64_t c = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= num_channel)
        return;
    double steps = em::pow2(bw) - 1;
    if (enc5[0] == -enc5[1])   // decided from channel 0 only (ATQ:286-294)
        steps -= 1;
    em::per_channel_param(enc5[c * 5], enc5[c * 5 + 1], (float) steps, params[c], params[num_channel + c],
                          params[2 * num_channel + c], params[3 * num_channel + c]);
}

// ---------------------------------------------------------------------------------------------------------------
// STE backward: grad_in = grad * [min <= x <= max]   (quantsim_straight_through_grad.py:91-118)
// The product with 1.0f / 0.0f (not a select) keeps torch's `grad * mask` semantics for inf / NaN gradients.
// ---------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kThreads) ste_bwd_kernel(const T* __restrict__ x, const T* __restrict__ grad,
                                                           T* __restrict__ grad_in, int64_t count, float mn, float mx)
{
    constexpr int kV        = Elem<T>::kPerVec;
    constexpr int kU        = 2;   // two tensors are read: keep the same bytes in flight as the forward
    const int64_t num_vec   = count / kV;
    const int64_t num_tiles = (num_vec + kThreads * kU - 1) / (kThreads * kU);
    for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x)
    {
        const int64_t v0 = tile * (kThreads * kU) + threadIdx.x;
        uint4 rx[kU], rg[kU];
#pragma unroll
        for (int u = 0; u < kU; ++u)
        {
            const int64_t v = v0 + (int64_t) u * kThreads;
            if (v < num_vec)
            {
                rx[u] = ldg_stream(reinterpret_cast<const uint4*>(x) + v);
                rg[u] = ldg_stream(reinterpret_cast<const uint4*>(grad) + v);
            }
        }
#pragma unroll
        for (int u = 0; u < kU; ++u)
        {
            const int64_t v = v0 + (int64_t) u * kThreads;
            if (v < num_vec)
            {
                float fx[kV], fg[kV];
                Elem<T>::unpack(rx[u], fx);
                Elem<T>::unpack(rg[u], fg);
#pragma unroll
                for (int k = 0; k < kV; ++k)
                    fg[k] = __fmul_rn(fg[k], (mn <= fx[k] && fx[k] <= mx) ? 1.0f : 0.0f);
                stg_stream(reinterpret_cast<uint4*>(grad_in) + v, Elem<T>::pack(fg));
            }
        }
    }
    if (blockIdx.x == 0)
    {
        const int64_t i = num_vec * kV + threadIdx.x;
        if (i < count)
        {
            const float xv = Elem<T>::load(x + i);
            Elem<T>::store(grad_in + i, __fmul_rn(Elem<T>::load(grad + i), (mn <= xv && xv <= mx) ? 1.0f : 0.0f));
        }
    }
}

// Where the per-channel range comes from: two float arrays, or the rows {min, max, delta, offset, bw} (doubles) the grid
// search left on the device -- narrowed here exactly as `tensor.to(torch.float32)` narrows them (round to nearest even),
// and, for the reference's 0-dim comparison in a bf16 tensor's dtype, rounded once more to bf16.
struct RangeSrc
{
    const float* mins;
    const float* maxs;
    const double* enc5;   // non-null: use this instead of mins / maxs
    int bf16_round;
    __device__ __forceinline__ void load(int64_t c, float& mn, float& mx) const
    {
        if (enc5 == nullptr)
        {
            mn = __ldg(mins + c), mx = __ldg(maxs + c);
            return;
        }
        mn = __double2float_rn(__ldg(enc5 + 5 * c));
        mx = __double2float_rn(__ldg(enc5 + 5 * c + 1));
        if (bf16_round)
        {
            mn = __bfloat162float(__float2bfloat16_rn(mn));
            mx = __bfloat162float(__float2bfloat16_rn(mx));
        }
    }
};

template <typename T>
__global__ void __launch_bounds__(kThreads)
    ste_bwd_per_channel_kernel(const T* __restrict__ x, const T* __restrict__ grad, T* __restrict__ grad_in,
                               int64_t count, int64_t C, int64_t L, RangeSrc src)
{
    constexpr int kV        = Elem<T>::kPerVec;
    const int64_t num_vec   = count / kV;
    const int64_t stride    = (int64_t) gridDim.x * kThreads;
    for (int64_t v = (int64_t) blockIdx.x * kThreads + threadIdx.x; v < num_vec; v += stride)
    {
        const uint4 rx = ldg_stream(reinterpret_cast<const uint4*>(x) + v);
        const uint4 rg = ldg_stream(reinterpret_cast<const uint4*>(grad) + v);
        float fx[kV], fg[kV];
        Elem<T>::unpack(rx, fx);
        Elem<T>::unpack(rg, fg);
        const int64_t i0 = v * kV;
        int64_t g        = i0 / L;
        int64_t rem      = i0 - g * L;
        int64_t c        = g % C;
        float mn, mx;
        src.load(c, mn, mx);
#pragma unroll
        for (int k = 0; k < kV; ++k)
        {
            fg[k] = __fmul_rn(fg[k], (mn <= fx[k] && fx[k] <= mx) ? 1.0f : 0.0f);
            if (++rem == L)
            {
                rem = 0;
                c   = (c + 1 == C) ? 0 : c + 1;
                src.load(c, mn, mx);
            }
        }
        stg_stream(reinterpret_cast<uint4*>(grad_in) + v, Elem<T>::pack(fg));
    }
    if (blockIdx.x == 0)
    {
        const int64_t i = num_vec * kV + threadIdx.x;
        if (i < count)
        {
            const int64_t c = (i / L) % C;
            const float xv  = Elem<T>::load(x + i);
            float mn, mx;
            src.load(c, mn, mx);
            Elem<T>::store(grad_in + i, __fmul_rn(Elem<T>::load(grad + i), (mn <= xv && xv <= mx) ? 1.0f : 0.0f));
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------
static thread_local char g_error[512] = "";

void set_error(const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what)
{
    set_error("CUDA error %d (%s) in %s", (int) e, cudaGetErrorString(e), what);
    return AB_ERR_CUDA;
}

int num_sms()
{
    static int cached[64] = {0};
    int dev               = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64)
        return 148;
    if (cached[dev] == 0)
    {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
            n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

template <typename K>
static int grid_for(K kernel, int64_t tiles, int threads, size_t smem = 0)
{
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem) != cudaSuccess || per_sm <= 0)
        per_sm = 1;
    const int64_t resident = (int64_t) per_sm * num_sms();
    return (int) (tiles < 1 ? 1 : (tiles < resident ? tiles : resident));
}

static bool aligned16(const void* p)
{
    return (reinterpret_cast<uintptr_t>(p) & 15u) == 0;
}


template <typename T, Op kOp>
static int launch_per_tensor_typed(const T* in, T* out, int64_t count, const TensorArgs& a, bool stochastic,
                                   cudaStream_t st)
{
    constexpr int kV      = Elem<T>::kPerVec;
    const int64_t tiles   = (count / kV + kThreads * kUnroll - 1) / (kThreads * kUnroll);
    if (stochastic)
    {
        auto k = per_tensor_kernel<T, kOp, true>;
        k<<<grid_for(k, tiles, kThreads), kThreads, 0, st>>>(in, out, count, a);
    }
    else
    {
        auto k = per_tensor_kernel<T, kOp, false>;
        k<<<grid_for(k, tiles, kThreads), kThreads, 0, st>>>(in, out, count, a);
    }
    AB_CUDA_CHECK(cudaGetLastError());
    return AB_OK;
}

template <typename T, Op kOp>
static int launch_scalar_typed(const T* in, T* out, int64_t count, const TensorArgs& a, bool stochastic,
                               cudaStream_t st)
{
    const int64_t tiles = (count + kThreads - 1) / kThreads;
    if (stochastic)
    {
        auto k = per_tensor_scalar_kernel<T, kOp, true>;
        k<<<grid_for(k, tiles, kThreads), kThreads, 0, st>>>(in, out, count, a);
    }
    else
    {
        auto k = per_tensor_scalar_kernel<T, kOp, false>;
        k<<<grid_for(k, tiles, kThreads), kThreads, 0, st>>>(in, out, count, a);
    }
    AB_CUDA_CHECK(cudaGetLastError());
    return AB_OK;
}

template <Op kOp>
static int launch_per_tensor(const void* in, void* out, int64_t count, int dtype, TensorArgs a, int round_mode,
                             cudaStream_t st)
{
    if (count < 0 || (count > 0 && (in == nullptr || out == nullptr)))
    {
        set_error("null tensor pointer or negative count");
        return AB_ERR_INVALID;
    }
    static const int reverse = [] {
        const char* e = getenv("AB_QDQ_REVERSE");
        return (e == nullptr || e[0] != '0') ? 1 : 0;
    }();
    a.reverse = reverse;
    if (dtype != AB_F32 && dtype != AB_BF16)
    {
        set_error("unsupported dtype %d", dtype);
        return AB_ERR_UNSUPPORTED;
    }
    if (round_mode != AB_ROUND_NEAREST && round_mode != AB_ROUND_STOCHASTIC)
    {
        set_error("Unknown rounding mode.");   // DlQ/src/trim_functions.cpp:162
        return AB_ERR_INVALID;
    }
    if (count == 0)
        return AB_OK;
    const bool stochastic = round_mode == AB_ROUND_STOCHASTIC;
    // A view that starts at an odd element offset (x[1:]) is contiguous but not 16-byte aligned: such tensors take the
    // element-wise kernel (still coalesced, 4 / 2-byte accesses).
    const bool vec = aligned16(in) && aligned16(out);
    if (dtype == AB_F32)
        return vec ? launch_per_tensor_typed<float, kOp>((const float*) in, (float*) out, count, a, stochastic, st)
                   : launch_scalar_typed<float, kOp>((const float*) in, (float*) out, count, a, stochastic, st);
    return vec ? launch_per_tensor_typed<__nv_bfloat16, kOp>((const __nv_bfloat16*) in, (__nv_bfloat16*) out, count, a,
                                                             stochastic, st)
               : launch_scalar_typed<__nv_bfloat16, kOp>((const __nv_bfloat16*) in, (__nv_bfloat16*) out, count, a,
                                                         stochastic, st);
}

static Enc4 narrow(const ab_encoding& e)
{
    // the double encoding narrows to the tensor's float type at the call (DlQ/src/trim_functions.cpp:178)
    return Enc4 {(float) e.min, (float) e.max, (float) e.delta, (float) e.offset};
}

}   // namespace ab

using namespace ab;

extern "C"
{
const char* ab_last_error(void)
{
    return g_error;
}

int ab_version(void)
{
    return 100;
}

int ab_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess)
    {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int ab_qdq_per_tensor_fwd(const void* in, void* out, int64_t count, int dtype, double enc_min, double enc_max,
                          int bw, int round_mode, uint64_t seed, void* stream)
{
    ab_encoding e;
    em::fill_encoding_info(bw, enc_min, enc_max, e);
    TensorArgs a {narrow(e), nullptr, 0.0f, seed};
    return launch_per_tensor<Op::kQdq>(in, out, count, dtype, a, round_mode, (cudaStream_t) stream);
}

int ab_qdq_per_tensor_fwd_dev(const void* in, void* out, int64_t count, int dtype, const float* enc4,
                              int round_mode, uint64_t seed, void* stream)
{
    if (enc4 == nullptr || !aligned16(enc4))
    {
        set_error("enc4 must be a 16-byte aligned device pointer");
        return AB_ERR_INVALID;
    }
    TensorArgs a {Enc4 {0, 0, 1, 0}, enc4, 0.0f, seed};
    return launch_per_tensor<Op::kQdq>(in, out, count, dtype, a, round_mode, (cudaStream_t) stream);
}

int ab_quantize_to_grid(const void* in, void* out, int64_t count, int dtype, double enc_min, double enc_max, int bw,
                        int round_mode, int shift_to_signed, uint64_t seed, void* stream)
{
    ab_encoding e;
    em::fill_encoding_info(bw, enc_min, enc_max, e);
    // `unsigned int shift = pow(2, bw - 1)` then `out[i] -= shift` in float (DlQ/src/trim_functions.cpp:206-216)
    unsigned int shift = 0;
    if (shift_to_signed)
        shift = (unsigned int) em::pow2(e.bw - 1);
    TensorArgs a {narrow(e), nullptr, (float) shift, seed};
    return launch_per_tensor<Op::kQuantize>(in, out, count, dtype, a, round_mode, (cudaStream_t) stream);
}

int ab_qdq_per_channel_fwd(const void* in, void* out, int64_t num_channel, int64_t num_element,
                           int64_t num_element_per_channel, int dtype, const float* params, int round_mode,
                           uint64_t seed, void* stream)
{
    if (num_element < 0 || num_channel <= 0 || num_element_per_channel <= 0 || params == nullptr ||
        (num_element > 0 && (in == nullptr || out == nullptr)))
    {
        set_error("invalid per-channel arguments");
        return AB_ERR_INVALID;
    }
    if (dtype != AB_F32 && dtype != AB_BF16)
    {
        set_error("unsupported dtype %d", dtype);
        return AB_ERR_UNSUPPORTED;
    }
    if (round_mode != AB_ROUND_NEAREST && round_mode != AB_ROUND_STOCHASTIC)
    {
        set_error("Unknown rounding mode.");
        return AB_ERR_INVALID;
    }
    if (num_element == 0)
        return AB_OK;
    if (!aligned16(in) || !aligned16(out))
    {
        set_error("per-channel tensors must be 16-byte aligned");
        return AB_ERR_INVALID;
    }
    cudaStream_t st = (cudaStream_t) stream;
    ChannelArgs a {params, num_channel, num_element_per_channel, seed, 0, 0};
    const bool stochastic = round_mode == AB_ROUND_STOCHASTIC;
    const bool fast       = !stochastic && num_element_per_channel < (int64_t) 0x7fff0000 &&
                      num_channel < (int64_t) 0x7fffffff;
    // bf16 is issue-bound, and there the straight-line run kernel of broadcast.cu (no staging, no barriers: every vector
    // fetches its channel's four parameters through L1) is the faster one: 0.85 against 0.78 of the HBM roofline.
    // fp32: the staged kernel below is the faster one on large tensors (0.93-0.95 against 0.89-0.92 from 256 MB up), the run
    // kernel on small ones, where the staging barriers cost more than they save (16 MB: 6.9 against 8.2 us) -- and weights,
    // the tensors that are quantized per channel, are small.
    const bool run_bf16 = dtype == AB_BF16 && (num_element_per_channel % 8 == 0 || num_element_per_channel >= 1024);
    const bool run_fp32 = dtype == AB_F32 && num_element_per_channel % 4 == 0 && num_element_per_channel >= 512 &&
                          num_element <= (int64_t) 8 << 20;
    if (fast && (run_bf16 || run_fp32) && num_element <= num_channel * num_element_per_channel)
    {
        // the run kernel indexes with 32 bits: a tensor of 2^31 elements or more goes in slices of whole channels
        const int64_t channels_per_launch = std::max<int64_t>(1, (int64_t) 0x7fff0000 / num_element_per_channel);
        const int64_t slice               = channels_per_launch * num_element_per_channel;
        const size_t es                   = dtype == AB_BF16 ? 2 : 4;
        for (int64_t e0 = 0, c0 = 0; e0 < num_element; e0 += slice, c0 += channels_per_launch)
        {
            const int rc = launch_run_qdq(static_cast<const char*>(in) + e0 * es, static_cast<char*>(out) + e0 * es,
                                          std::min(slice, num_element - e0), num_element_per_channel, params + c0,
                                          params + num_channel + c0, params + 2 * num_channel + c0,
                                          params + 3 * num_channel + c0, dtype, st);
            if (rc != AB_OK)
                return rc;
        }
        return AB_OK;
    }
    if (fast)
    {
        // CUTLASS-style fast divmod constants, exact for dividends below 2^31
        const uint32_t d = (uint32_t) num_element_per_channel;
        if (d > 1)
        {
            uint32_t lg = 0;
            while ((1ull << lg) < d)
                ++lg;
            const uint32_t p = 31 + lg;
            a.div_mul        = (uint32_t) (((1ull << p) + d - 1) / d);
            a.div_shift      = p - 32;
        }
        if (dtype == AB_F32)
        {
            auto k              = per_channel_fast_kernel<float>;
            const int64_t tiles = (num_element + (int64_t) kThreads * kUnroll * 4 - 1) / ((int64_t) kThreads * kUnroll * 4);
            k<<<grid_for(k, tiles, kThreads), kThreads, 0, st>>>((const float*) in, (float*) out, num_element, a);
        }
        else
        {
            auto k              = per_channel_fast_kernel<__nv_bfloat16>;
            const int64_t tiles = (num_element + (int64_t) kThreads * kUnroll * 8 - 1) / ((int64_t) kThreads * kUnroll * 8);
            k<<<grid_for(k, tiles, kThreads), kThreads, 0, st>>>((const __nv_bfloat16*) in, (__nv_bfloat16*) out,
                                                                 num_element, a);
        }
        AB_CUDA_CHECK(cudaGetLastError());
        return AB_OK;
    }
    const int64_t es    = dtype == AB_F32 ? 4 : 8;
    const int64_t tiles = (num_element / es + kThreads - 1) / kThreads;
    if (dtype == AB_F32)
    {
        if (stochastic)
        {
            auto k = per_channel_kernel<float, true>;
            k<<<grid_for(k, tiles, kThreads), kThreads, 0, st>>>((const float*) in, (float*) out, num_element, a);
        }
        else
        {
            auto k = per_channel_kernel<float, false>;
            k<<<grid_for(k, tiles, kThreads), kThreads, 0, st>>>((const float*) in, (float*) out, num_element, a);
        }
    }
    else
    {
        if (stochastic)
        {
            auto k = per_channel_kernel<__nv_bfloat16, true>;
            k<<<grid_for(k, tiles, kThreads), kThreads, 0, st>>>((const __nv_bfloat16*) in, (__nv_bfloat16*) out,
                                                                 num_element, a);
        }
        else
        {
            auto k = per_channel_kernel<__nv_bfloat16, false>;
            k<<<grid_for(k, tiles, kThreads), kThreads, 0, st>>>((const __nv_bfloat16*) in, (__nv_bfloat16*) out,
                                                                 num_element, a);
        }
    }
    AB_CUDA_CHECK(cudaGetLastError());
    return AB_OK;
}

int ab_per_channel_params_dev(const double* enc5, int64_t num_channel, int bw, float* params, void* stream)
{
    if (enc5 == nullptr || params == nullptr || num_channel <= 0)
    {
        set_error("invalid per-channel arguments");
        return AB_ERR_INVALID;
    }
    per_channel_params_kernel<<<(unsigned) ((num_channel + 127) / 128), 128, 0, (cudaStream_t) stream>>>(enc5, num_channel,
                                                                                                       bw, params);
    AB_CUDA_CHECK(cudaGetLastError());
    return AB_OK;
}

int ab_qdq_ste_bwd(const void* x, const void* grad, void* grad_in, int64_t count, int dtype, float enc_min,
                   float enc_max, void* stream)
{
    if (count < 0 || (count > 0 && (x == nullptr || grad == nullptr || grad_in == nullptr)))
    {
        set_error("null tensor pointer or negative count");
        return AB_ERR_INVALID;
    }
    if (dtype != AB_F32 && dtype != AB_BF16)
    {
        set_error("unsupported dtype %d", dtype);
        return AB_ERR_UNSUPPORTED;
    }
    if (count == 0)
        return AB_OK;
    if (!aligned16(x) || !aligned16(grad) || !aligned16(grad_in))
    {
        set_error("STE tensors must be 16-byte aligned");
        return AB_ERR_INVALID;
    }
    cudaStream_t st = (cudaStream_t) stream;
    if (dtype == AB_F32)
    {
        auto k              = ste_bwd_kernel<float>;
        const int64_t tiles = (count / 4 + kThreads * 2 - 1) / (kThreads * 2);
        k<<<grid_for(k, tiles, kThreads), kThreads, 0, st>>>((const float*) x, (const float*) grad, (float*) grad_in,
                                                             count, enc_min, enc_max);
    }
    else
    {
        auto k              = ste_bwd_kernel<__nv_bfloat16>;
        const int64_t tiles = (count / 8 + kThreads * 2 - 1) / (kThreads * 2);
        k<<<grid_for(k, tiles, kThreads), kThreads, 0, st>>>((const __nv_bfloat16*) x, (const __nv_bfloat16*) grad,
                                                             (__nv_bfloat16*) grad_in, count, enc_min, enc_max);
    }
    AB_CUDA_CHECK(cudaGetLastError());
    return AB_OK;
}

static int launch_ste_per_channel(const void* x, const void* grad, void* grad_in, int64_t num_channel, int64_t num_element,
                                  int64_t num_element_per_channel, int dtype, RangeSrc src, void* stream);

int ab_qdq_ste_bwd_per_channel(const void* x, const void* grad, void* grad_in, int64_t num_channel,
                               int64_t num_element, int64_t num_element_per_channel, int dtype,
                               const float* enc_min, const float* enc_max, void* stream)
{
    if (enc_min == nullptr || enc_max == nullptr)
    {
        set_error("invalid per-channel arguments");
        return AB_ERR_INVALID;
    }
    return launch_ste_per_channel(x, grad, grad_in, num_channel, num_element, num_element_per_channel, dtype,
                                  RangeSrc {enc_min, enc_max, nullptr, 0}, stream);
}

int ab_qdq_ste_bwd_enc5(const void* x, const void* grad, void* grad_in, int64_t num_channel, int64_t num_element,
                        int64_t num_element_per_channel, int dtype, const double* enc5, int range_in_bf16, void* stream)
{
    if (enc5 == nullptr)
    {
        set_error("invalid per-channel arguments");
        return AB_ERR_INVALID;
    }
    return launch_ste_per_channel(x, grad, grad_in, num_channel, num_element, num_element_per_channel, dtype,
                                  RangeSrc {nullptr, nullptr, enc5, range_in_bf16 ? 1 : 0}, stream);
}

static int launch_ste_per_channel(const void* x, const void* grad, void* grad_in, int64_t num_channel, int64_t num_element,
                                  int64_t num_element_per_channel, int dtype, RangeSrc src, void* stream)
{
    if (num_element < 0 || num_channel <= 0 || num_element_per_channel <= 0 ||
        (num_element > 0 && (x == nullptr || grad == nullptr || grad_in == nullptr)))
    {
        set_error("invalid per-channel arguments");
        return AB_ERR_INVALID;
    }
    if (dtype != AB_F32 && dtype != AB_BF16)
    {
        set_error("unsupported dtype %d", dtype);
        return AB_ERR_UNSUPPORTED;
    }
    if (num_element == 0)
        return AB_OK;
    if (!aligned16(x) || !aligned16(grad) || !aligned16(grad_in))
    {
        set_error("STE tensors must be 16-byte aligned");
        return AB_ERR_INVALID;
    }
    cudaStream_t st = (cudaStream_t) stream;
    if (dtype == AB_F32)
    {
        auto k              = ste_bwd_per_channel_kernel<float>;
        const int64_t tiles = (num_element / 4 + kThreads - 1) / kThreads;
        k<<<grid_for(k, tiles, kThreads), kThreads, 0, st>>>((const float*) x, (const float*) grad, (float*) grad_in,
                                                             num_element, num_channel, num_element_per_channel, src);
    }
    else
    {
        auto k              = ste_bwd_per_channel_kernel<__nv_bfloat16>;
        const int64_t tiles = (num_element / 8 + kThreads - 1) / kThreads;
        k<<<grid_for(k, tiles, kThreads), kThreads, 0, st>>>((const __nv_bfloat16*) x, (const __nv_bfloat16*) grad,
                                                             (__nv_bfloat16*) grad_in, num_element, num_channel,
                                                             num_element_per_channel, src);
    }
    AB_CUDA_CHECK(cudaGetLastError());
    return AB_OK;
}

// ---- host helpers ---------------------------------------------------------------------------------------------
int ab_gate_min_max(double* enc_min, double* enc_max)
{
    if (!enc_min || !enc_max)
        return AB_ERR_INVALID;
    em::gate_min_max(*enc_min, *enc_max);
    return AB_OK;
}

int ab_fill_encoding_info(int bw, double enc_min, double enc_max, ab_encoding* out)
{
    if (!out)
        return AB_ERR_INVALID;
    em::fill_encoding_info(bw, enc_min, enc_max, *out);
    return AB_OK;
}

int ab_tf_compute_encoding(int bw, double mn, double mx, int use_symmetric, int use_strict_symmetric,
                           int use_unsigned_symmetric, ab_encoding* out)
{
    if (!out)
        return AB_ERR_INVALID;
    em::tf_encoding(bw, mn, mx, use_symmetric != 0, use_strict_symmetric != 0, use_unsigned_symmetric != 0, *out);
    return AB_OK;
}

int ab_tf_analyzer_encoding(int bw, double run_min, double run_max, int use_symmetric, int use_strict_symmetric,
                            int use_unsigned_symmetric, ab_encoding* out)
{
    if (!out)
        return AB_ERR_INVALID;
    em::tf_analyzer_encoding(bw, run_min, run_max, use_symmetric != 0, use_strict_symmetric != 0,
                             use_unsigned_symmetric != 0, *out);
    return AB_OK;
}

int ab_compute_partial_encoding(int bw, ab_encoding* e, int sym, int unsigned_sym, int strict)
{
    if (!e)
        return AB_ERR_INVALID;
    if (e->min == 0 && e->max == 0)
    {
        // computeMinMaxRangeFromDeltaOffset -- DlQ/src/quantization_utils.cpp:158-205
        if (e->bw == 0)
        {
            set_error("Encodings must have a valid non-zero bitwidth");
            return AB_ERR_INVALID;
        }
        if (e->delta == 0 && e->offset > 0)
        {
            set_error("Encoding must have a valid non-zero delta/offset if min and max are zero");
            return AB_ERR_INVALID;
        }
        double steps = em::pow2((uint8_t) bw) - 1;
        if (sym && strict)
            steps -= 1;
        e->min = e->offset * e->delta;
        if (sym && ((e->min < 0.0) || !unsigned_sym))
            e->max = e->delta * floor(steps / 2);
        else
            e->max = e->delta * steps + e->min;
        if (e->max - e->min < em::kEpsilon)
            em::gate_min_max(e->min, e->max);
        return AB_OK;
    }
    if (e->delta == 0)
    {
        // computeDeltaAndOffsetFromMinMax -- DlQ/src/quantization_utils.cpp:207-228
        if (e->bw == 0)
        {
            set_error("Encodings must have a valid non-zero bitwidth");
            return AB_ERR_INVALID;
        }
        const ab_encoding orig = *e;
        em::tf_encoding(bw, orig.min, orig.max, sym != 0, strict != 0, unsigned_sym != 0, *e);
        e->min = orig.min;
        e->max = orig.max;
        return AB_OK;
    }
    set_error("Cannot determine how to compute partial encoding");   // DlQ/src/TensorQuantizer.cpp:341
    return AB_ERR_INVALID;
}

int ab_per_channel_params(const double* enc_min, const double* enc_max, int num_channel, int bw, float* params)
{
    if (!enc_min || !enc_max || !params || num_channel <= 0)
        return AB_ERR_INVALID;
    double steps = em::pow2(bw) - 1;
    if (enc_min[0] == -enc_max[0])   // decided from channel 0 only (ATQ:286-294)
        steps -= 1;
    const float steps_f = (float) steps;
    for (int c = 0; c < num_channel; ++c)
        em::per_channel_param(enc_min[c], enc_max[c], steps_f, params[c], params[num_channel + c],
                              params[2 * num_channel + c], params[3 * num_channel + c]);
    return AB_OK;
}
}   // extern "C"
