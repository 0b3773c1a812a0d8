// TEST-ONLY device program: brute-force equivalence of the XU-free fast paths in aimet_b200/csrc/common.cuh against the
// plain IEEE formulation (__fdiv_rn, roundf) over billions of operand pairs. Prints the mismatch counts; exit code 1 if
// any. Built and run by tests/test_gpu_fastpath.py with nvcc on the GPU box.
#include <cstdio>
#include <cstdlib>

#include "../../aimet_b200/csrc/common.cuh"

namespace ab
{
void set_error(const char*, ...) {}
int cuda_fail(cudaError_t, const char*) { return -2; }
int num_sms() { return 148; }
}   // namespace ab
using namespace ab;

__device__ __forceinline__ uint64_t mix(uint64_t z)
{
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

// One divisor per block-iteration, many numerators per divisor. Numerators: k + f with f near 0, 0.5 and random, scaled
// by the divisor (so quotients sit on and around the rounding boundaries that matter), plus raw random floats.
__global__ void check_div(uint64_t seed, int iters, unsigned long long* bad_div, unsigned long long* bad_round,
                          unsigned long long* bad_qdq, unsigned long long* bad_bin)
{
    unsigned long long nd = 0, nr = 0, nq = 0, nb = 0;
    for (int it = 0; it < iters; ++it)
    {
        const uint64_t h = mix(seed + (uint64_t) blockIdx.x * 1000003ull + it);
        // divisor: random mantissa, exponent in [-64, 64]
        const uint32_t mant = (uint32_t) h & 0x7fffffu;
        const int ex        = (int) ((h >> 23) % 129) - 64;
        float d             = __uint_as_float(((uint32_t) (ex + 127) << 23) | mant);
        if ((h >> 40) & 1)
            d = __uint_as_float(((uint32_t) (ex + 127) << 23) | 0x7fffffu - (mant & 7));   // mantissa near all-ones
        const Divisor dv = make_divisor(d);
        const uint64_t g = mix(h ^ (threadIdx.x * 0x9E3779B97F4A7C15ull));
        float xs[4];
        const float k = (float) ((g >> 8) % 70000);
        xs[0]         = (k + 0.5f) * d;
        xs[1]         = __uint_as_float(__float_as_uint((k + 0.5f) * d) + (int) ((g >> 3) & 7) - 3);
        xs[2]         = (k + (float) ((g >> 32) & 0xffffff) * (1.0f / 16777216.0f)) * d;
        xs[3]         = __uint_as_float((uint32_t) (g >> 16) & 0x7fffffffu) * (((g >> 5) & 1) ? -1.f : 1.f);
#pragma unroll
        for (int j = 0; j < 4; ++j)
        {
            const float x = xs[j];
            if (!(fabsf(x) < 3.0e38f) || x == 0.0f)
                continue;
            const float ref = __fdiv_rn(x, d);
            const float got = div_fast(x, dv);
            // only quotients that can influence a grid point / bin matter (|q| >= 2^-20 and finite)
            if (fabsf(ref) >= 9.5e-7f && fabsf(ref) < 1.0e30f && __float_as_uint(ref) != __float_as_uint(got))
                ++nd;
            // rounding, on the quotient and on nearby ties
            if (fabsf(ref) < 4194304.0f)
            {
                if (__float_as_uint(roundf(ref)) != __float_as_uint(round_half_away_small(ref)))
                    ++nr;
                const float tie = truncf(ref) + copysignf(0.5f, ref);
                if (__float_as_uint(roundf(tie)) != __float_as_uint(round_half_away_small(tie)))
                    ++nr;
            }
        }
        // a full QDQ and a full bin computation with these parameters
        const float off  = -truncf((float) ((h >> 12) % 65536));
        const float stps = (float) ((1u << (2 + (h >> 50) % 15)) - 1);
        Enc4 e {__fmul_rn(off, d), __fmul_rn(__fadd_rn(off, stps), d), d, off};
        if (qdq_fast_ok(e, dv))
        {
#pragma unroll
            for (int j = 0; j < 4; ++j)
            {
                const float slow = dequantize_value(quantize_value<false>(xs[j], e, 0, 0), e);
                const float fast = qdq_fast(xs[j], e, dv);
                if (__float_as_uint(slow) != __float_as_uint(fast) && !(slow == 0.0f && fast == 0.0f))
                    ++nq;
                if (qdq_pos_ok(e, dv))   // the two-instructions-shorter form for grids whose positions stay above -0.5
                {
                    const float pos = qdq_fast_pos(xs[j], e, dv, __fsub_rn(12582912.0f, e.offset));
                    if (__float_as_uint(slow) != __float_as_uint(pos) && !(slow == 0.0f && pos == 0.0f))
                        ++nq;
                }
                // quantize-only: bit-exact INCLUDING the sign of zero, also for +-0 inputs
                const float xq = ((g >> (13 + j)) & 15) == 0 ? copysignf(0.0f, xs[j]) : xs[j];
                if (__float_as_uint(quantize_value<false, true>(xq, e, 0, 0)) != __float_as_uint(quantize_fast(xq, e, dv)))
                    ++nq;
            }
        }
        // histogram bin: fast formulation vs round(x / bucket - offset) with the x86 drop rules
        {
            constexpr float kMagic = 12582912.0f;
            const float offset     = (float) ((int) ((h >> 20) % 1024) - 512) + (float) ((h >> 44) & 0xff) / 256.0f;
#pragma unroll
            for (int j = 0; j < 4; ++j)
            {
                float x = xs[j] * (((g >> 7) & 1) ? 1.0f : 1.0f / 128.0f);
                if (((g >> 9) & 3) == 0)      // construct quotients that land exactly on k +- 0.5 relative to the offset
                    x = (offset + (float) ((int) ((g >> 11) % 516) - 2) + (((g >> 21) & 1) ? 0.5f : -0.5f)) * d;
                const float r  = round_half_away(__fsub_rn(__fdiv_rn(x, d), offset));
                const int slow = (r >= 0.0f && r < 512.0f) ? (int) r : -1;
                float v        = __fsub_rn(div_fast(x, dv), offset);
                v              = (v == -0.5f) ? -1.0f : v;
                const float t  = __fadd_rd(__fadd_rz(v, 0.5f), kMagic);
                const uint32_t idx = min(__float_as_uint(t) - __float_as_uint(kMagic), 512u);
                const int fast = idx < 512u ? (int) idx : -1;
                if (slow != fast)
                    ++nb;
            }
        }
    }
    if (nd) atomicAdd(bad_div, nd);
    if (nr) atomicAdd(bad_round, nr);
    if (nq) atomicAdd(bad_qdq, nq);
    if (nb) atomicAdd(bad_bin, nb);
}

int main(int argc, char** argv)
{
    const int iters = argc > 1 ? atoi(argv[1]) : 4096;
    unsigned long long* bad;
    cudaMalloc(&bad, 4 * sizeof(unsigned long long));
    cudaMemset(bad, 0, 4 * sizeof(unsigned long long));
    check_div<<<148 * 8, 256>>>(20261018ull, iters, bad, bad + 1, bad + 2, bad + 3);
    unsigned long long h[4];
    if (cudaMemcpy(h, bad, sizeof(h), cudaMemcpyDeviceToHost) != cudaSuccess)
    {
        printf("CUDA failure: %s\n", cudaGetErrorString(cudaGetLastError()));
        return 2;
    }
    const double pairs = 148.0 * 8 * 256 * iters * 4;
    printf("pairs=%.3g bad_div=%llu bad_round=%llu bad_qdq=%llu bad_bin=%llu\n", pairs, h[0], h[1], h[2], h[3]);
    return (h[0] | h[1] | h[2] | h[3]) ? 1 : 0;
}
