// Microbenchmark: tcgen05.ld (TMEM -> registers) throughput per SM as a function of warps and load shape.
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/microbench/ldtm_bw tools/microbench/ldtm_bw.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int SHAPE> // 32, 64 or 128 columns per instruction
__device__ __forceinline__ uint32_t ld(uint32_t taddr)
{
    uint32_t acc = 0;
    if constexpr (SHAPE == 32)
    {
        uint32_t v[32];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                       "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                     : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int i = 0; i < 32; ++i) acc ^= v[i];
    }
    return acc;
}

// NB loads of 32 columns issued back to back, then one wait
template <int NB>
__device__ __forceinline__ uint32_t ld_batch(uint32_t taddr)
{
    uint32_t v[NB][32];
#pragma unroll
    for (int b = 0; b < NB; ++b)
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                     : "=r"(v[b][0]), "=r"(v[b][1]), "=r"(v[b][2]), "=r"(v[b][3]), "=r"(v[b][4]), "=r"(v[b][5]), "=r"(v[b][6]), "=r"(v[b][7]), "=r"(v[b][8]), "=r"(v[b][9]), "=r"(v[b][10]), "=r"(v[b][11]), "=r"(v[b][12]), "=r"(v[b][13]), "=r"(v[b][14]), "=r"(v[b][15]),
                       "=r"(v[b][16]), "=r"(v[b][17]), "=r"(v[b][18]), "=r"(v[b][19]), "=r"(v[b][20]), "=r"(v[b][21]), "=r"(v[b][22]), "=r"(v[b][23]), "=r"(v[b][24]), "=r"(v[b][25]), "=r"(v[b][26]), "=r"(v[b][27]), "=r"(v[b][28]), "=r"(v[b][29]), "=r"(v[b][30]), "=r"(v[b][31])
                     : "r"(taddr + 32 * b));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    uint32_t acc = 0;
#pragma unroll
    for (int b = 0; b < NB; ++b)
#pragma unroll
        for (int i = 0; i < 32; ++i) acc ^= v[b][i];
    return acc;
}

template <int NB>
__global__ void __launch_bounds__(512, 1) bench(int iters, long long *cycles, uint32_t *sink)
{
    __shared__ uint32_t tmem_ptr;
    const int warp = threadIdx.x >> 5;
    if (warp == 0)
    {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_ptr)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = tmem_ptr + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t acc = 0;
    __syncthreads();
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) acc ^= ld_batch<NB>(base + ((i * NB * 32) & 255));
    __syncthreads();
    const long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
    if (acc == 0x12345678u) sink[threadIdx.x] = acc;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_ptr) : "memory");
}

template <int NB>
void run(int warps)
{
    long long *d_c; uint32_t *d_s;
    cudaMalloc(&d_c, 8); cudaMalloc(&d_s, 4096);
    const int iters = 2000;
    bench<NB><<<148, warps * 32>>>(iters, d_c, d_s);
    bench<NB><<<148, warps * 32>>>(iters, d_c, d_s);
    cudaError_t e = cudaDeviceSynchronize();
    long long c = 0;
    cudaMemcpy(&c, d_c, 8, cudaMemcpyDeviceToHost);
    const double bytes = (double)iters * NB * 32 * 32 * 4 * warps;
    printf("warps %2d  batch %d x32 : %8lld clk  %7.1f B/clk/SM  %6.1f clk per x32 load per warp  (%s)\n", warps, NB, c, bytes / c,
           (double)c / (iters * NB), cudaGetErrorString(e));
    cudaFree(d_c); cudaFree(d_s);
}

int main()
{
    for (int w : {1, 2, 4, 8, 16})
    {
        run<1>(w);
        run<2>(w);
        run<3>(w);
    }
    return 0;
}
