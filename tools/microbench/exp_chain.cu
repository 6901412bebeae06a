// Microbenchmark: the soft-max pass-2 element chain (scale/shift, round to f16, exp2, round to f16, f32 sum) on registers only,
// as a function of warps per SM sub-partition.  Prints clocks per element per warp.
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/microbench/exp_chain tools/microbench/exp_chain.cu
#include <cstdio>
#include <cstdint>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

__device__ __forceinline__ float ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ void add_f32_f16(float &acc, unsigned short h) { asm("add.rn.f32.f16 %0, %1, %0;" : "+f"(acc) : "h"(h)); }

// 2^y on the FMA pipe: Cody-Waite split through the 1.5 * 2^23 magic constant, degree-4 polynomial on [-0.5, 0.5]
// (max relative error 2.7e-6), exponent inserted with an integer add
__device__ __forceinline__ float ex2_poly(float y)
{
    y = fmaxf(y, -126.0f);
    const float t = y + 12582912.0f;
    const float f = y - (t - 12582912.0f);
    float p = fmaf(f, 0.009570096619427204f, 0.05591785907745361f);
    p = fmaf(p, f, 0.240247443318367f);
    p = fmaf(p, f, 0.6931217908859253f);
    p = fmaf(p, f, 0.9999992847442627f);
    return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}

template <int MODE>
__device__ __forceinline__ uint32_t pair(float x0, float x1, float &lsum)
{
    if constexpr (MODE == 5) // exponentials on the FMA pipe (every call of this mode)
    {
        const float2 xr = __half22float2(__floats2half2_rn(x0, x1));
        const __half2 e = __floats2half2_rn(ex2_poly(xr.x * 1.4426950408889634f), ex2_poly(xr.y * 1.4426950408889634f));
        add_f32_f16(lsum, __half_as_ushort(__low2half(e)));
        add_f32_f16(lsum, __half_as_ushort(__high2half(e)));
        return *reinterpret_cast<const uint32_t *>(&e);
    }
    if constexpr (MODE == 0) // full chain
    {
        const float2 xr = __half22float2(__floats2half2_rn(x0, x1));
        const __half2 e = __floats2half2_rn(ex2(xr.x * 1.4426950408889634f), ex2(xr.y * 1.4426950408889634f));
        add_f32_f16(lsum, __half_as_ushort(__low2half(e)));
        add_f32_f16(lsum, __half_as_ushort(__high2half(e)));
        return *reinterpret_cast<const uint32_t *>(&e);
    }
    else if constexpr (MODE == 1) // no MUFU (replace by fma)
    {
        const float2 xr = __half22float2(__floats2half2_rn(x0, x1));
        const __half2 e = __floats2half2_rn(fmaf(xr.x, 1.4426950408889634f, 0.5f), fmaf(xr.y, 1.4426950408889634f, 0.5f));
        add_f32_f16(lsum, __half_as_ushort(__low2half(e)));
        add_f32_f16(lsum, __half_as_ushort(__high2half(e)));
        return *reinterpret_cast<const uint32_t *>(&e);
    }
    else if constexpr (MODE == 3) // rounding of the argument by a Veltkamp split instead of the f16 round trip
    {
        const float t0 = __fmul_rn(x0, 8193.0f), t1 = __fmul_rn(x1, 8193.0f);
        const float r0 = __fsub_rn(t0, __fsub_rn(t0, x0)), r1 = __fsub_rn(t1, __fsub_rn(t1, x1));
        const __half2 e = __floats2half2_rn(ex2(r0 * 1.4426950408889634f), ex2(r1 * 1.4426950408889634f));
        add_f32_f16(lsum, __half_as_ushort(__low2half(e)));
        add_f32_f16(lsum, __half_as_ushort(__high2half(e)));
        return *reinterpret_cast<const uint32_t *>(&e);
    }
    else if constexpr (MODE == 4) // Veltkamp argument + un-rounded f32 sum (no mixed-precision adds)
    {
        const float t0 = __fmul_rn(x0, 8193.0f), t1 = __fmul_rn(x1, 8193.0f);
        const float r0 = __fsub_rn(t0, __fsub_rn(t0, x0)), r1 = __fsub_rn(t1, __fsub_rn(t1, x1));
        const float e0 = ex2(r0 * 1.4426950408889634f), e1 = ex2(r1 * 1.4426950408889634f);
        const __half2 e = __floats2half2_rn(e0, e1);
        lsum += e0;
        lsum += e1;
        return *reinterpret_cast<const uint32_t *>(&e);
    }
    else // MUFU only
    {
        const float a = ex2(x0), b = ex2(x1);
        lsum += a;
        lsum += b;
        return __float_as_uint(a) ^ __float_as_uint(b);
    }
}

template <int MODE>
__global__ void __launch_bounds__(512, 1) bench(int iters, long long *cycles, uint32_t *sink, float scale, float mxs)
{
    uint32_t v[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(-0.01f * (float)((threadIdx.x * 37 + i * 11) & 255));
    float l0 = 0.f, l1 = 0.f, l2 = 0.f, l3 = 0.f;
    uint32_t acc = 0;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it)
    {
        uint32_t pk[16];
#pragma unroll
        for (int j = 0; j < 16; j += 4)
        {
            pk[j] = pair<(MODE == 6 || MODE == 7) ? 0 : MODE>(__fmaf_rn(__uint_as_float(v[2 * j]), scale, -mxs), __fmaf_rn(__uint_as_float(v[2 * j + 1]), scale, -mxs), l0);
            pk[j + 1] = pair<(MODE == 7) ? 5 : (MODE == 6 ? 0 : MODE)>(__fmaf_rn(__uint_as_float(v[2 * j + 2]), scale, -mxs), __fmaf_rn(__uint_as_float(v[2 * j + 3]), scale, -mxs), l1);
            pk[j + 2] = pair<(MODE == 6 || MODE == 7) ? 0 : MODE>(__fmaf_rn(__uint_as_float(v[2 * j + 4]), scale, -mxs), __fmaf_rn(__uint_as_float(v[2 * j + 5]), scale, -mxs), l2);
            pk[j + 3] = pair<(MODE == 6 || MODE == 7) ? 5 : MODE>(__fmaf_rn(__uint_as_float(v[2 * j + 6]), scale, -mxs), __fmaf_rn(__uint_as_float(v[2 * j + 7]), scale, -mxs), l3);
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) { acc ^= pk[j]; v[2 * j] ^= (pk[j] & 1u); v[2 * j + 1] ^= ((pk[j] >> 16) & 1u); } // every input changes every iteration
    }
    __syncthreads();
    const long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
    if (acc == 0x12345678u || l0 + l1 + l2 + l3 == 1.2345f) sink[threadIdx.x] = acc;
}

template <int MODE>
void run(int warps, const char *name)
{
    long long *d_c; uint32_t *d_s;
    cudaMalloc(&d_c, 8); cudaMalloc(&d_s, 4096);
    const int iters = 2000;
    bench<MODE><<<148, warps * 32>>>(iters, d_c, d_s, 0.125f, 0.3f);
    bench<MODE><<<148, warps * 32>>>(iters, d_c, d_s, 0.125f, 0.3f);
    cudaError_t e = cudaDeviceSynchronize();
    long long c = 0;
    cudaMemcpy(&c, d_c, 8, cudaMemcpyDeviceToHost);
    printf("%-10s warps/SM %2d : %6.2f clk per element per warp, %6.2f elements/clk/SM (%s)\n", name, warps, (double)c / (iters * 32.0),
           (double)iters * 32 * 32 * warps / c, cudaGetErrorString(e));
    cudaFree(d_c); cudaFree(d_s);
}

int main()
{
    for (int w : {4, 8, 16})
    {
        run<0>(w, "full");
        run<1>(w, "no-mufu");
        run<2>(w, "mufu-only");
        run<3>(w, "veltkamp");
        run<4>(w, "velt+fadd");
        run<5>(w, "poly-all");
        run<6>(w, "poly-1/4");
        run<7>(w, "poly-1/2");
    }
    return 0;
}
