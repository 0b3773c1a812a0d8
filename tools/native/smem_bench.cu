// DEV TOOL (not part of the product): how many histogram samples per clock per SM can the shared-memory pipe take?
// Variants of the per-sample instruction sequence on register-generated data (no global loads), 148 CTAs x 512 threads,
// bins privatised per lane exactly like hist_kernel (word b*32+lane).
//   0: ATOMS only (index from an LCG)                 1: per-group table LDS.64 + integer affine + ATOMS (bf16 table path)
//   2: current float path (div_fast, magic rounding)  3: LDS.64 only (no atomics; result xor-folded)
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int kBins = 512, kLanes = 32, kThreads = 512;

template <int V>
__global__ void __launch_bounds__(kThreads, 1) bench(int iters, uint32_t* sink, float c0, float c1, float c2, float off)
{
    extern __shared__ __align__(16) uint8_t smem[];
    uint32_t* s_hist = reinterpret_cast<uint32_t*>(smem);
    uint2* s_tab     = reinterpret_cast<uint2*>(smem + (kBins + 1) * kLanes * 4);
    for (int i = threadIdx.x; i < (kBins + 1) * kLanes; i += kThreads)
        s_hist[i] = 0;
    for (int i = threadIdx.x; i < 512; i += kThreads)
        s_tab[i] = make_uint2(40000u + i * 17u, (uint32_t) (i * 2654435761u) >> 6);
    __syncthreads();
    const int lane        = threadIdx.x & 31;
    uint32_t* s_hist_lane = s_hist + lane;
    uint32_t w            = threadIdx.x * 2654435761u + blockIdx.x * 40503u + 12345u;
    uint32_t fold         = 0;
    for (int it = 0; it < iters; ++it)
    {
#pragma unroll
        for (int j = 0; j < 8; ++j)
        {
            w = w * 1664525u + 1013904223u;   // one 32-bit word = two bf16 samples
#pragma unroll
            for (int h = 0; h < 2; ++h)
            {
                if (V == 0)
                {
                    const uint32_t u = (h ? (w >> 20) : (w >> 7)) & 511u;
                    atomicAdd(s_hist_lane + u * kLanes, 1u);
                }
                else if (V == 1 || V == 3)
                {
                    // group = sign+exponent (9 bits) restricted to 8 distinct values, like real data
                    const uint32_t t = (h ? (w >> 20) : (w >> 4)) & 0x038u;
                    const uint2 ab   = *reinterpret_cast<const uint2*>(reinterpret_cast<const uint8_t*>(s_tab) + t);
                    const uint32_t m = h ? (w >> 16) : (w & 0xffffu);
                    const uint32_t v = m * ab.x + ab.y;
                    if (V == 1)
                    {
                        const uint32_t u = min(v >> 16, (uint32_t) kBins);
                        atomicAdd(s_hist_lane + u * kLanes, 1u);
                    }
                    else
                        fold ^= v;
                }
                else if (V == 2)
                {
                    const float x = __uint_as_float(h ? (w & 0xffff0000u) : (w << 16));
                    const float q0 = __fmul_rn(x, c0);
                    const float r  = __fmaf_rn(q0, c1, x);
                    float v        = __fsub_rn(__fmaf_rn(c0, r, q0), off);
                    v              = (v == -0.5f) ? -1.0f : v;
                    constexpr float kMagic = 12582912.0f;
                    const float t  = __fadd_rd(__fadd_rz(v, 0.5f), kMagic);
                    const uint32_t u = min(__float_as_uint(t) - __float_as_uint(kMagic), (uint32_t) kBins);
                    atomicAdd(s_hist_lane + u * kLanes, 1u);
                }
            }
        }
    }
    __syncthreads();
    uint32_t s = fold;
    for (int i = threadIdx.x; i < kBins * kLanes; i += kThreads)
        s += s_hist[i];
    if (s == 0xdeadbeefu)
        sink[0] = s;
}

template <int V>
void run(const char* name, int iters, uint32_t* sink)
{
    const size_t smem = (kBins + 1) * kLanes * 4 + 512 * 8;
    cudaFuncSetAttribute(bench<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    bench<V><<<148, kThreads, smem>>>(iters / 8, sink, 0.37f, -2.7f, 0.37f, 3.0f);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; ++r)
    {
        cudaEventRecord(a);
        bench<V><<<148, kThreads, smem>>>(iters, sink, 0.37f, -2.7f, 0.37f, 3.0f);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms;
        cudaEventElapsedTime(&ms, a, b);
        best = ms < best ? ms : best;
    }
    const double samples = 148.0 * kThreads * (double) iters * 16;
    const double per_ns_sm = samples / (best * 1e6) / 148.0;
    printf("%-28s %8.3f ms  %6.2f samples/ns/SM  (= %5.2f /clk at 1.965 GHz)  -> bf16 %5.0f GB/s, fp32 %5.0f GB/s\n", name, best,
           per_ns_sm, per_ns_sm / 1.965, per_ns_sm * 148 * 2, per_ns_sm * 148 * 4);
    if (cudaGetLastError() != cudaSuccess)
        printf("CUDA error\n");
}

int main()
{
    uint32_t* sink;
    cudaMalloc(&sink, 4);
    run<0>("atoms only", 4000, sink);
    run<1>("table lds64 + imad + atoms", 4000, sink);
    run<2>("float path + atoms", 4000, sink);
    run<3>("table lds64 only", 4000, sink);
    return 0;
}
