// Quantize-dequantize with an encoding tensor BROADCAST over the input (blockwise / LPBQ weights, any mix of per-channel
// axes): every element uses the {min, max, delta, offset} at the position its index maps to in the encoding tensor.
//
// Reference: quantizeDequantizeBroadcast (DlQ/include/DlQuantization/Quantization.hpp:193-221,
// DlQ/src/trim_functions.cpp:633-662 = CPU parity target; the reference's own GPU twin is trim_functions.cu:96-122), used
// by the ONNX custom op (TrainingExtensions/onnx/src/AimetOpUtils.h:269). Per element it decomposes the flat index
// dimension by dimension (two integer divisions per dimension) and loads four scalars.
//
// Here the trailing dimensions along which the encoding is broadcast are folded into one run length `inner`: the encoding
// index only depends on i / inner. A 128-bit vector that lies inside one run (the common case: a block is >= 4 elements)
// resolves its encoding once -- 32-bit multiply-high divisions with host-prepared constants -- and runs the straight-line
// XU-free QDQ of common.cuh; vectors that straddle a run boundary, unaligned tensors and tensors of 2^31 elements or more
// take the element-wise path. Bytes: 2 s per element (+ the encoding tensor, which lives in L1/L2).
#include "common.cuh"

namespace ab
{
namespace
{

constexpr int kBcThreads = 256;
constexpr int kBcUnroll  = 4;
constexpr int kMaxDims   = 8;

struct Dim
{
    int64_t stride;       // elements of the OUTER index space (flat index / inner) per step along this dimension
    int64_t enc_stride;   // encoding elements per step along this dimension (0 = broadcast)
    uint32_t mul, shift;  // n / stride == umulhi(n, mul) >> shift for n < 2^31 (stride > 1)
};

struct BroadcastArgs
{
    Dim dims[kMaxDims];   // outer dimensions only, outermost first
    int num_dims;
    int linear;                          // the encoding tensor is laid out like the outer index space: index == i / inner
    int64_t inner;                       // run length over which the encoding index is constant
    uint32_t inner_mul, inner_shift;     // fast division by inner
    const float *mn, *mx, *delta, *offset;
};

__device__ __forceinline__ uint32_t fast_div(uint32_t n, int64_t d, uint32_t mul, uint32_t shift)
{
    return d == 1 ? n : (__umulhi(n, mul) >> shift);
}

// encoding index of outer position g (32-bit fast path)
__device__ __forceinline__ int64_t enc_index32(const BroadcastArgs& a, uint32_t g)
{
    if (a.linear)   // blockwise / per-channel encodings stored contiguously: the common case, no decomposition at all
        return (int64_t) g;
    int64_t idx = 0;
#pragma unroll 1
    for (int d = 0; d < a.num_dims; ++d)
    {
        const uint32_t q = fast_div(g, a.dims[d].stride, a.dims[d].mul, a.dims[d].shift);
        g -= q * (uint32_t) a.dims[d].stride;
        idx += a.dims[d].enc_stride * (int64_t) q;
    }
    return idx;
}
__device__ __forceinline__ int64_t enc_index64(const BroadcastArgs& a, int64_t g)
{
    if (a.linear)
        return g;
    int64_t idx = 0;
    for (int d = 0; d < a.num_dims; ++d)
    {
        const int64_t q = g / a.dims[d].stride;
        g -= q * a.dims[d].stride;
        idx += a.dims[d].enc_stride * q;
    }
    return idx;
}

__device__ __forceinline__ Enc4 load_enc(const BroadcastArgs& a, int64_t idx)
{
    return Enc4 {__ldg(a.mn + idx), __ldg(a.mx + idx), __ldg(a.delta + idx), __ldg(a.offset + idx)};
}

__device__ __forceinline__ float qdq_exact(float x, const Enc4& e)
{
    return dequantize_value(quantize_value<false>(x, e, 0, 0), e);
}

// kSimple: the encoding index is the run index itself (a.linear) and a run is a whole number of 128-bit vectors, so no
// vector straddles a run boundary -- blockwise and per-channel encodings with block sizes that are multiples of 4 (fp32) /
// 8 (bf16) elements. The per-vector work is then one multiply-high, four loads and the divisor set-up.
// kShared: linear encodings with runs at least as long as the 32 * kBcUnroll vectors a thread's chunk spans; with kSimple the
// run is a whole number of vectors, without it a run may end inside a vector (handled where it happens) -- first branch.
template <typename T, bool kSimple, bool kShared = false>
__global__ void __launch_bounds__(kBcThreads)
    broadcast_fast_kernel(const T* __restrict__ in, T* __restrict__ out, int64_t count, BroadcastArgs a)
{
    constexpr int kV        = Elem<T>::kPerVec;
    const int64_t num_vec   = count / kV;
    const int64_t num_tiles = (num_vec + kBcThreads * kBcUnroll - 1) / (kBcThreads * kBcUnroll);
    const uint32_t inner    = (uint32_t) a.inner;
    if constexpr (kShared)
    {
        // Long runs that are whole numbers of vectors (per-channel weights, blocks): every WARP takes kBcUnroll consecutive rows of
        // 32 vectors, so a thread's vectors lie within 32 * kBcUnroll * kV consecutive elements -- inside ONE run whenever
        // the run is at least that long and the chunk does not straddle its end. Then the thread fetches the encoding and
        // sets up the divisor once for all its vectors instead of once per vector (the per-vector set-up was 3 of the 17
        // instructions per bf16 element), and takes the shorter QDQ form where the grid allows it. A warp whose chunk does
        // straddle a run end redoes the set-up only where the run changes.
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        constexpr bool ragged = !kSimple;      // runs may end inside a vector (a.linear still holds: index == run index)
        for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x)
        {
            const int64_t v0 = tile * (kBcThreads * kBcUnroll) + (int64_t) warp * (32 * kBcUnroll) + lane;
            uint4 raw[kBcUnroll];
#pragma unroll
            for (int u = 0; u < kBcUnroll; ++u)
                if (v0 + u * 32 < num_vec)
                    raw[u] = ldg_stream(reinterpret_cast<const uint4*>(in) + v0 + u * 32);
            // does every lane's chunk lie inside one run? (the common case: a single fetch + set-up per thread, no per-vector index)
            // (first element of the first vector, LAST element of the last one: a ragged run length may split a vector)
            const int64_t v_last = min(v0 + (kBcUnroll - 1) * 32, num_vec - 1 - ((num_vec - 1 - v0) & 31));
            const uint32_t g0    = fast_div((uint32_t) ((v0 < num_vec ? v0 : 0) * kV), a.inner, a.inner_mul, a.inner_shift);
            const uint32_t g1    = fast_div((uint32_t) ((v0 < num_vec ? v_last : 0) * kV + (kV - 1)), a.inner, a.inner_mul,
                                            a.inner_shift);
            const bool uniform   = __all_sync(0xffffffffu, g0 == g1);   // (the tile loop is uniform: all 32 lanes are here)
            if (v0 >= num_vec)
                continue;
            if (uniform)
            {
                const Enc4 e     = load_enc(a, (int64_t) g0);
                const Divisor dv = make_divisor(e.delta);
                const bool ok    = qdq_fast_ok(e, dv);
                const bool pos   = ok && qdq_pos_ok(e, dv);
                const float moff = __fsub_rn(12582912.0f, e.offset);
#pragma unroll
                for (int u = 0; u < kBcUnroll; ++u)
                {
                    if (v0 + u * 32 >= num_vec)
                        continue;
                    float f[kV];
                    Elem<T>::unpack(raw[u], f);
                    if (pos)
                    {
#pragma unroll
                        for (int k = 0; k < kV; ++k)
                            f[k] = qdq_fast_pos(f[k], e, dv, moff);
                    }
                    else if (ok)
                    {
#pragma unroll
                        for (int k = 0; k < kV; ++k)
                            f[k] = qdq_fast(f[k], e, dv);
                    }
                    else
                    {
#pragma unroll
                        for (int k = 0; k < kV; ++k)
                            f[k] = qdq_exact(f[k], e);
                    }
                    stg_stream(reinterpret_cast<uint4*>(out) + v0 + u * 32, Elem<T>::pack(f));
                }
                continue;
            }
            // a run ends inside this warp's chunk: the encoding and the divisor set-up are redone only where the run changes
            // between two of a thread's vectors (only that short set-up diverges, the arithmetic below stays converged)
            uint32_t g_cur = 0xffffffffu;
            Enc4 e {};
            Divisor dv {};
            bool ok = false, pos = false;
            float moff = 0.0f;
#pragma unroll
            for (int u = 0; u < kBcUnroll; ++u)
            {
                const int64_t v = v0 + u * 32;
                if (v >= num_vec)
                    continue;
                const uint32_t g = fast_div((uint32_t) (v * kV), a.inner, a.inner_mul, a.inner_shift);
                if (ragged && (uint32_t) (v * kV) - g * inner + kV > inner)
                {
                    // the run ends inside this vector (run length not a whole number of vectors): element by element
                    float f[kV];
                    Elem<T>::unpack(raw[u], f);
                    uint32_t r = (uint32_t) (v * kV) - g * inner, gg = g;
                    Enc4 ee    = load_enc(a, (int64_t) gg);
#pragma unroll
                    for (int k = 0; k < kV; ++k)
                    {
                        f[k] = qdq_exact(f[k], ee);
                        if (++r == inner && k + 1 < kV)
                        {
                            r  = 0;
                            ee = load_enc(a, (int64_t) ++gg);
                        }
                    }
                    stg_stream(reinterpret_cast<uint4*>(out) + v, Elem<T>::pack(f));
                    continue;
                }
                if (g != g_cur)
                {
                    g_cur = g;
                    e     = load_enc(a, (int64_t) g);
                    dv    = make_divisor(e.delta);
                    ok    = qdq_fast_ok(e, dv);
                    pos   = ok && qdq_pos_ok(e, dv);
                    moff  = __fsub_rn(12582912.0f, e.offset);
                }
                float f[kV];
                Elem<T>::unpack(raw[u], f);
                if (pos)
                {
#pragma unroll
                    for (int k = 0; k < kV; ++k)
                        f[k] = qdq_fast_pos(f[k], e, dv, moff);
                }
                else if (ok)
                {
#pragma unroll
                    for (int k = 0; k < kV; ++k)
                        f[k] = qdq_fast(f[k], e, dv);
                }
                else
                {
#pragma unroll
                    for (int k = 0; k < kV; ++k)
                        f[k] = qdq_exact(f[k], e);
                }
                stg_stream(reinterpret_cast<uint4*>(out) + v, Elem<T>::pack(f));
            }
        }
        if (blockIdx.x == 0)
        {
            const int64_t i = num_vec * kV + threadIdx.x;
            if (i < count)
                Elem<T>::store(out + i, qdq_exact(Elem<T>::load(in + i), load_enc(a, enc_index64(a, i / a.inner))));
        }
        return;
    }
    for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x)
    {
        const int64_t v0 = tile * (kBcThreads * kBcUnroll) + threadIdx.x;
        uint4 raw[kBcUnroll];
#pragma unroll
        for (int u = 0; u < kBcUnroll; ++u)
        {
            const int64_t v = v0 + (int64_t) u * kBcThreads;
            if (v < num_vec)
                raw[u] = ldg_stream(reinterpret_cast<const uint4*>(in) + v);
        }
        // the encodings only depend on the vector's position: resolve and request them while the data is in flight
        uint32_t g[kBcUnroll], rem[kBcUnroll];
        Enc4 enc[kBcUnroll];
#pragma unroll
        for (int u = 0; u < kBcUnroll; ++u)
        {
            const int64_t v   = v0 + (int64_t) u * kBcThreads;
            const uint32_t i0 = (uint32_t) ((v < num_vec ? v : 0) * kV);
            g[u]              = fast_div(i0, a.inner, a.inner_mul, a.inner_shift);
            rem[u]            = kSimple ? 0u : i0 - g[u] * inner;
            enc[u]            = load_enc(a, kSimple ? (int64_t) g[u] : enc_index32(a, g[u]));
        }
#pragma unroll
        for (int u = 0; u < kBcUnroll; ++u)
        {
            const int64_t v = v0 + (int64_t) u * kBcThreads;
            if (v >= num_vec)
                continue;
            float f[kV];
            Elem<T>::unpack(raw[u], f);
            Enc4 e = enc[u];
            if (kSimple || rem[u] + kV <= inner)
            {
                const Divisor dv = make_divisor(e.delta);
                if (qdq_fast_ok(e, dv))
                {
#pragma unroll
                    for (int k = 0; k < kV; ++k)
                        f[k] = qdq_fast(f[k], e, dv);
                }
                else
                {
#pragma unroll
                    for (int k = 0; k < kV; ++k)
                        f[k] = qdq_exact(f[k], e);
                }
            }
            else
            {
                uint32_t r = rem[u], gg = g[u];
#pragma unroll
                for (int k = 0; k < kV; ++k)
                {
                    f[k] = qdq_exact(f[k], e);
                    if (++r == inner && k + 1 < kV)
                    {
                        r = 0;
                        e = load_enc(a, enc_index32(a, ++gg));
                    }
                }
            }
            stg_stream(reinterpret_cast<uint4*>(out) + v, Elem<T>::pack(f));
        }
    }
    if (blockIdx.x == 0)
    {
        const int64_t i = num_vec * kV + threadIdx.x;
        if (i < count)
            Elem<T>::store(out + i, qdq_exact(Elem<T>::load(in + i), load_enc(a, enc_index64(a, i / a.inner))));
    }
}

// any alignment, any size
template <typename T>
__global__ void __launch_bounds__(kBcThreads)
    broadcast_scalar_kernel(const T* __restrict__ in, T* __restrict__ out, int64_t count, BroadcastArgs a)
{
    const int64_t stride = (int64_t) gridDim.x * kBcThreads;
    for (int64_t i = (int64_t) blockIdx.x * kBcThreads + threadIdx.x; i < count; i += stride)
        Elem<T>::store(out + i, qdq_exact(Elem<T>::load(in + i), load_enc(a, enc_index64(a, i / a.inner))));
}

void magic(int64_t d, uint32_t& mul, uint32_t& shift)
{
    mul = 0, shift = 0;
    if (d > 1 && d < (int64_t) 0x7fff0000)
    {
        uint32_t lg = 0;
        while ((1ull << lg) < (uint64_t) d)
            ++lg;
        const uint32_t p = 31 + lg;
        mul              = (uint32_t) (((1ull << p) + (uint64_t) d - 1) / (uint64_t) d);
        shift            = p - 32;
    }
}

int grid_size(const void* kernel, int64_t tiles)
{
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kBcThreads, 0) != cudaSuccess || per_sm <= 0)
        per_sm = 1;
    int64_t grid = (int64_t) per_sm * num_sms();
    if (tiles < grid)
        grid = tiles;
    return (int) (grid < 1 ? 1 : grid);
}

template <typename T>
int launch(const void* in, void* out, int64_t count, const BroadcastArgs& a, bool fast, cudaStream_t st)
{
    const T* x = reinterpret_cast<const T*>(in);
    T* y       = reinterpret_cast<T*>(out);
    if (fast)
    {
        constexpr int kV    = Elem<T>::kPerVec;
        const int64_t tiles = (count / kV + kBcThreads * kBcUnroll - 1) / (kBcThreads * kBcUnroll);
        if (a.linear && a.inner % kV == 0 && a.inner >= (int64_t) 32 * kBcUnroll * kV)
            broadcast_fast_kernel<T, true, true>
                <<<grid_size((const void*) broadcast_fast_kernel<T, true, true>, tiles), kBcThreads, 0, st>>>(x, y, count, a);
        else if (a.linear && a.inner >= (int64_t) 32 * kBcUnroll * kV)       // long runs of a ragged length
            broadcast_fast_kernel<T, false, true>
                <<<grid_size((const void*) broadcast_fast_kernel<T, false, true>, tiles), kBcThreads, 0, st>>>(x, y, count, a);
        else if (a.linear && a.inner % kV == 0)
            broadcast_fast_kernel<T, true><<<grid_size((const void*) broadcast_fast_kernel<T, true>, tiles), kBcThreads, 0, st>>>(
                x, y, count, a);
        else
            broadcast_fast_kernel<T, false><<<grid_size((const void*) broadcast_fast_kernel<T, false>, tiles), kBcThreads, 0, st>>>(
                x, y, count, a);
    }
    else
    {
        const int64_t tiles = (count + kBcThreads - 1) / kBcThreads;
        broadcast_scalar_kernel<T><<<grid_size((const void*) broadcast_scalar_kernel<T>, tiles), kBcThreads, 0, st>>>(x, y, count,
                                                                                                                 a);
    }
    AB_CUDA_CHECK(cudaGetLastError());
    return AB_OK;
}

}   // namespace

int launch_run_qdq(const void* in, void* out, int64_t count, int64_t run, const float* mn, const float* mx,
                   const float* delta, const float* offset, int dtype, cudaStream_t stream)
{
    BroadcastArgs a {};
    a.mn = mn, a.mx = mx, a.delta = delta, a.offset = offset;
    a.inner    = run;
    a.linear   = 1;
    a.num_dims = 0;
    magic(a.inner, a.inner_mul, a.inner_shift);
    if (dtype == AB_F32)
        return launch<float>(in, out, count, a, true, stream);
    return launch<__nv_bfloat16>(in, out, count, a, true, stream);
}

}   // namespace ab

using namespace ab;

extern "C" int ab_qdq_broadcast_fwd(const void* in, void* out, int64_t num_element, int num_dims,
                                    const int64_t* input_strides, const int64_t* encoding_strides, const float* enc_min,
                                    const float* enc_max, const float* enc_delta, const float* enc_offset, int dtype,
                                    void* stream)
{
    if (num_element < 0 || num_dims < 1 || num_dims > kMaxDims || !input_strides || !encoding_strides)
    {
        set_error("bad geometry: %d dimensions (1..%d supported), %lld elements", num_dims, kMaxDims, (long long) num_element);
        return AB_ERR_INVALID;
    }
    if (dtype != AB_F32 && dtype != AB_BF16)
    {
        set_error("unsupported dtype %d", dtype);
        return AB_ERR_INVALID;
    }
    if (!enc_min || !enc_max || !enc_delta || !enc_offset || (num_element > 0 && (!in || !out)))
    {
        set_error("null pointer");
        return AB_ERR_INVALID;
    }
    for (int d = 0; d < num_dims; ++d)
        if (input_strides[d] < 1 || encoding_strides[d] < 0 || (d > 0 && input_strides[d - 1] % input_strides[d] != 0))
        {
            set_error("input strides must be those of a contiguous tensor, encoding strides non-negative");
            return AB_ERR_INVALID;
        }
    if (num_element == 0)
        return AB_OK;

    // fold the trailing broadcast dimensions into one run length
    BroadcastArgs a {};
    a.mn = enc_min, a.mx = enc_max, a.delta = enc_delta, a.offset = enc_offset;
    int last = -1;
    for (int d = 0; d < num_dims; ++d)
        if (encoding_strides[d] != 0)
            last = d;
    a.inner = (last < 0) ? (num_element > 0 ? num_element : 1) : input_strides[last];
    if (a.inner < 1)
        a.inner = 1;
    a.num_dims = 0;
    for (int d = 0; d <= last; ++d)
    {
        Dim& dim       = a.dims[a.num_dims++];
        dim.stride     = input_strides[d] / a.inner;
        dim.enc_stride = encoding_strides[d];
        magic(dim.stride, dim.mul, dim.shift);
    }
    magic(a.inner, a.inner_mul, a.inner_shift);
    // outer position g = sum_d q_d * stride_d and encoding index = sum_d q_d * enc_stride_d: identical when the strides are
    a.linear = 1;
    for (int d = 0; d < a.num_dims; ++d)
        if (a.dims[d].enc_stride != a.dims[d].stride)
        {
            // a dimension of extent 1 never contributes (q_d == 0), whatever its strides say
            const int64_t extent = (d == 0 ? num_element : input_strides[d - 1]) / input_strides[d];
            if (extent != 1)
                a.linear = 0;
        }

    const bool fast = ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15u) == 0 &&
                      num_element < (int64_t) 0x7fff0000;
    cudaStream_t st = (cudaStream_t) stream;
    if (dtype == AB_F32)
        return launch<float>(in, out, num_element, a, fast, st);
    return launch<__nv_bfloat16>(in, out, num_element, a, fast, st);
}
