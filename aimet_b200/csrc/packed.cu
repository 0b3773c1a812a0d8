// Packed integer export: the integer grid value of every element in the narrowest integer type that holds `bw` bits.
//
// Reference: ITensorQuantizationSim::quantizeTensorPacked (DlQ/src/TensorQuantizationSim.cpp:128-139) ->
// quantizeToFxpPackedCpu (DlQ/src/trim_functions.cpp:221-388). CPU only there ("GPU packed quantization not supported",
// :194-196). Unlike the float QDQ kernels, the reference computes this one in DOUBLE on the double encoding:
//     q = round(max(min((double) x, enc.max), enc.min) / enc.delta - enc.offset)
// unsigned: one uint8 per value below 8 bit (clamped to 2^bw - 1; the reference's sub-byte packing is compiled out),
// uint8 / uint16 / uint32 above; shiftToSigned: q -= 2^(bw-1) - 1 (not 2^(bw-1) as quantizeToFxp), stored as int8 (masked to
// the low bw bits below 8 bit), int8 / int16 / int32. Streaming kernel: 128-bit loads, one packed store per input vector;
// the divisions are IEEE double divisions (__ddiv_rn), so the kernel is FP64-issue bound at roughly the HBM rate -- an export
// path, not a hot one.
#include "common.cuh"
#include "encoding_math.h"

namespace ab
{
namespace
{
constexpr int kPackThreads = 256;

struct PackArgs
{
    double mn, mx, delta, offset, shift, top;   // top = 2^bw - 1
    int bw, is_signed;
};

// std::min / std::max as the reference calls them: the FIRST argument comes back when the comparison is false (NaN)
__device__ __forceinline__ double std_min(double a, double b) { return (b < a) ? b : a; }
__device__ __forceinline__ double std_max(double a, double b) { return (a < b) ? b : a; }
// x86 cvttsd2si: out-of-range and NaN give the "integer indefinite" value
__device__ __forceinline__ int32_t cvtt_i32(double v)
{
    if (!(v == v) || v >= 2147483648.0 || v < -2147483649.0)
        return INT32_MIN;
    return (int32_t) v;
}
__device__ __forceinline__ int64_t cvtt_i64(double v)
{
    if (!(v == v) || v >= 9223372036854775808.0 || v < -9223372036854775808.0)
        return INT64_MIN;
    return (int64_t) v;
}

// the stored bits of one value, zero-extended to 32 bits
__device__ __forceinline__ uint32_t pack_one(float x, const PackArgs& a)
{
    double q = std_max(std_min((double) x, a.mx), a.mn);
    q        = __dsub_rn(__ddiv_rn(q, a.delta), a.offset);
    q        = round(q);
    if (!a.is_signed)
    {
        if (a.bw < 8)
        {
            const uint8_t shr = (uint8_t) cvtt_i32(q);
            return (uint8_t) cvtt_i32(std_max(std_min((double) shr, a.top), 0.0));
        }
        if (a.bw == 8)
            return (uint8_t) cvtt_i32(std_max(std_min(q, 255.0), 0.0));
        if (a.bw == 16)
            return (uint16_t) cvtt_i32(std_max(std_min(q, 65535.0), 0.0));
        return (uint32_t) cvtt_i64(std_max(std_min(q, 4294967295.0), 0.0));
    }
    q = __dsub_rn(q, a.shift);
    if (a.bw < 8)
        return (uint8_t) ((int8_t) cvtt_i32(q) & (int8_t) (int) a.top);
    if (a.bw == 8)
        return (uint8_t) (int8_t) cvtt_i32(std_max(std_min(q, 127.0), -128.0));
    if (a.bw == 16)
        return (uint16_t) (int16_t) cvtt_i32(std_max(std_min(q, 32767.0), -32768.0));
    return (uint32_t) cvtt_i32(std_max(std_min(q, 2147483647.0), -2147483648.0));
}

template <int kOutBytes>
__device__ __forceinline__ void store4(uint8_t* out, int64_t first_elem, const uint32_t (&v)[4])
{
    if (kOutBytes == 1)
        *reinterpret_cast<uint32_t*>(out + first_elem) = v[0] | (v[1] << 8) | (v[2] << 16) | (v[3] << 24);
    else if (kOutBytes == 2)
        *reinterpret_cast<uint2*>(out + first_elem * 2) = make_uint2(v[0] | (v[1] << 16), v[2] | (v[3] << 16));
    else
        *reinterpret_cast<uint4*>(out + first_elem * 4) = make_uint4(v[0], v[1], v[2], v[3]);
}

template <typename T, int kOutBytes>
__global__ void __launch_bounds__(kPackThreads) pack_kernel(const T* __restrict__ in, uint8_t* __restrict__ out, int64_t count,
                                                            PackArgs a, int vector_ok)
{
    constexpr int kV = Elem<T>::kPerVec;
    const int64_t num_vec = vector_ok ? count / kV : 0;
    for (int64_t v = (int64_t) blockIdx.x * kPackThreads + threadIdx.x; v < num_vec; v += (int64_t) gridDim.x * kPackThreads)
    {
        float f[kV];
        Elem<T>::unpack(ldg_stream(reinterpret_cast<const uint4*>(in) + v), f);
#pragma unroll
        for (int g = 0; g < kV / 4; ++g)
        {
            uint32_t w[4];
#pragma unroll
            for (int k = 0; k < 4; ++k)
                w[k] = pack_one(f[g * 4 + k], a);
            store4<kOutBytes>(out, v * kV + g * 4, w);
        }
    }
    // what the vector body did not cover: the tail, or everything when a pointer is misaligned
    for (int64_t i = num_vec * kV + (int64_t) blockIdx.x * kPackThreads + threadIdx.x; i < count;
         i += (int64_t) gridDim.x * kPackThreads)
    {
        const uint32_t w = pack_one(Elem<T>::load(in + i), a);
        if (kOutBytes == 1)
            out[i] = (uint8_t) w;
        else if (kOutBytes == 2)
            reinterpret_cast<uint16_t*>(out)[i] = (uint16_t) w;
        else
            reinterpret_cast<uint32_t*>(out)[i] = w;
    }
}

template <typename T, int kOutBytes>
int launch_pack(const void* in, void* out, int64_t count, const PackArgs& a, cudaStream_t stream)
{
    const bool ok = ((reinterpret_cast<uintptr_t>(in) & 15u) == 0) && ((reinterpret_cast<uintptr_t>(out) & 15u) == 0);
    constexpr int kV = Elem<T>::kPerVec;
    int64_t blocks   = (count / kV + kPackThreads - 1) / kPackThreads;
    const int64_t cap = (int64_t) num_sms() * 8;
    if (blocks > cap)
        blocks = cap;
    if (blocks < 1)
        blocks = 1;
    pack_kernel<T, kOutBytes><<<(unsigned) blocks, kPackThreads, 0, stream>>>((const T*) in, (uint8_t*) out, count, a,
                                                                              ok ? 1 : 0);
    AB_CUDA_CHECK(cudaGetLastError());
    return AB_OK;
}
}   // namespace
}   // namespace ab

using namespace ab;

extern "C" int ab_quantize_to_packed(const void* in, void* out, int64_t count, int dtype, double enc_min, double enc_max,
                                     int bw, int shift_to_signed, void* stream)
{
    if (count < 0 || (count > 0 && (in == nullptr || out == nullptr)))
    {
        set_error("null tensor pointer or negative count");
        return AB_ERR_INVALID;
    }
    if (!(bw == 1 || bw == 2 || bw == 4 || bw == 8 || bw == 16 || bw == 32))
    {
        set_error("Bit-width needs to be power of two and between 1 and 32.");   // the reference's message
        return AB_ERR_INVALID;
    }
    if (dtype != AB_F32 && dtype != AB_BF16)
    {
        set_error("unsupported dtype %d", dtype);
        return AB_ERR_UNSUPPORTED;
    }
    if (count == 0)
        return AB_OK;
    ab_encoding e;
    em::fill_encoding_info(bw, enc_min, enc_max, e);
    PackArgs a;
    a.mn = e.min, a.mx = e.max, a.delta = e.delta, a.offset = e.offset;
    a.top       = em::pow2(bw) - 1;
    a.shift     = shift_to_signed ? em::pow2(bw - 1) - 1 : 0.0;
    a.bw        = bw;
    a.is_signed = shift_to_signed ? 1 : 0;
    cudaStream_t st = (cudaStream_t) stream;
    const int bytes = (bw > 8 ? bw : 8) / 8;
    if (dtype == AB_F32)
        return bytes == 1 ? launch_pack<float, 1>(in, out, count, a, st)
                          : bytes == 2 ? launch_pack<float, 2>(in, out, count, a, st) : launch_pack<float, 4>(in, out, count, a, st);
    return bytes == 1 ? launch_pack<__nv_bfloat16, 1>(in, out, count, a, st)
                      : bytes == 2 ? launch_pack<__nv_bfloat16, 2>(in, out, count, a, st)
                                   : launch_pack<__nv_bfloat16, 4>(in, out, count, a, st);
}
