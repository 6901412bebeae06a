// Microbenchmark: the fc1 epilogue's per-element chain (bias add, round to f16 and back, tanh-form GELU through ex2 + rcp, pack
// to f16) on registers only, as a function of warps per SM sub-partition and with single pieces removed.
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/microbench/gelu_chain tools/microbench/gelu_chain.cu
#include <cstdio>
#include <cstdint>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

__device__ __forceinline__ float ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// 1/d for d in [1, 2^127) on the FMA pipe: magic-constant seed (12 % off) + three Newton steps (error ~4e-8)
__device__ __forceinline__ float rcp_newton(float d)
{
    float r = __int_as_float(0x7EF311C7 - __float_as_int(d));
    r = r * fmaf(-d, r, 2.0f);
    r = r * fmaf(-d, r, 2.0f);
    r = r * fmaf(-d, r, 2.0f);
    return r;
}

template <int MODE>
__device__ __forceinline__ float gelu(float x)
{
    if constexpr (MODE == 5)
    {
        const float w = fmaf(x * x, -0.10294323958083856f, -2.3022081981625516f);
        const float e = ex2(fminf(x * w, 126.0f));
        return x * rcp_newton(1.0f + e);
    }
    const float w = fmaf(x * x, -0.10294323958083856f, -2.3022081981625516f);
    float e;
    if constexpr (MODE == 2) e = x * w * 0.01f; else e = ex2(x * w);
    if constexpr (MODE == 1) return x * (1.0f + e) * 0.5f; else return x * rcp(1.0f + e);
}

template <int MODE> // 0 full, 1 no rcp, 2 no ex2, 3 no f16 round trip of the input, 4 round trip replaced by a Veltkamp split
__global__ void __launch_bounds__(512, 1) bench(int iters, long long *cycles, uint32_t *sink, float bias)
{
    uint32_t v[64];
#pragma unroll
    for (int i = 0; i < 64; ++i) v[i] = __float_as_uint(0.01f * (float)(((threadIdx.x * 37 + i * 11) & 255) - 128));
    uint32_t acc = 0;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it)
    {
        uint32_t packed[32];
#pragma unroll
        for (int e = 0; e < 32; ++e)
        {
            float x0 = __uint_as_float(v[2 * e]) + bias, x1 = __uint_as_float(v[2 * e + 1]) + bias;
            float r0 = x0, r1 = x1;
            if constexpr (MODE == 4)
            {
                // Veltkamp split: x rounded to an 11-bit significand with FP32 ops only (no f16 round trip)
                const float t0 = __fmul_rn(x0, 8193.0f), t1 = __fmul_rn(x1, 8193.0f);
                r0 = __fsub_rn(t0, __fsub_rn(t0, x0));
                r1 = __fsub_rn(t1, __fsub_rn(t1, x1));
            }
            else if constexpr (MODE != 3)
            {
                const float2 r = __half22float2(__floats2half2_rn(x0, x1));
                r0 = r.x; r1 = r.y;
            }
            const __half2 h = __floats2half2_rn(gelu<(MODE == 6) ? 0 : MODE>(r0), gelu<(MODE == 6) ? 5 : MODE>(r1));
            packed[e] = *reinterpret_cast<const uint32_t *>(&h);
        }
#pragma unroll
        for (int e = 0; e < 32; ++e) { acc ^= packed[e]; v[2 * e] ^= (packed[e] & 1u); v[2 * e + 1] ^= ((packed[e] >> 16) & 1u); }
    }
    __syncthreads();
    const long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
    if (acc == 0x12345678u) sink[threadIdx.x] = acc;
}

template <int MODE>
void run(int warps, const char *name)
{
    long long *d_c; uint32_t *d_s;
    cudaMalloc(&d_c, 8); cudaMalloc(&d_s, 4096);
    const int iters = 1000;
    bench<MODE><<<148, warps * 32>>>(iters, d_c, d_s, 0.03f);
    bench<MODE><<<148, warps * 32>>>(iters, d_c, d_s, 0.03f);
    cudaError_t e = cudaDeviceSynchronize();
    long long c = 0;
    cudaMemcpy(&c, d_c, 8, cudaMemcpyDeviceToHost);
    printf("%-12s warps/SM %2d : %6.2f clk per element per warp, %6.2f elements/clk/SM (%s)\n", name, warps, (double)c / (iters * 64.0),
           (double)iters * 64 * 32 * warps / c, cudaGetErrorString(e));
    cudaFree(d_c); cudaFree(d_s);
}

int main()
{
    for (int w : {4, 8, 16})
    {
        run<0>(w, "full");
        run<1>(w, "no-rcp");
        run<2>(w, "no-ex2");
        run<3>(w, "no-f16-trip");
        run<4>(w, "veltkamp");
        run<5>(w, "newton-all");
        run<6>(w, "newton-1/2");
    }
    return 0;
}
