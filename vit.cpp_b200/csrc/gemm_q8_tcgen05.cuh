// gemm_q8_tcgen05.cuh -- the reference's q8_0 x q8_0 linear layer on the INTEGER tensor cores: prototype (BASELINE.json configs[4]).
//
// Reference semantics (q8_0 model files): every mul_mat with a quantised weight first quantises its f32 activation rows to
// q8_0 blocks of 32 (quantize_row_q8_0, ggml-quants.c:702-790: d = amax / 127 stored as f16, q = rint(x * 127 / amax)) and then
// evaluates, per output element (ggml_vec_dot_q8_0_q8_0, ggml-quants.c:3521+; block layout ggml-quants.h:42-46)
//     y = sum over blocks b of  (d_w[b] * d_x[b]) * (sum over the block's 32 elements of q_w * q_x)      accumulated in f32.
// One f32 scale per 32-element block of BOTH operands means an int32 accumulator can span exactly one block:
//   * tcgen05.mma.kind::i8 has K = 32 per instruction == one q8_0 block, so every MMA goes into a FRESH accumulator slot
//     (4 slots x 128 int32 columns = all 512 TMEM columns, rotated), signalled per slot with tcgen05.commit;
//   * eight correction warps (two per TMEM lane quarter, 64 columns each) read each slot back (tcgen05.ld), turn the exact
//     integer sums into f32 (|sum| <= 32 * 127 * 127 < 2^22: one IADD of 1.5 * 2^23's bit pattern + one exact FADD, no I2F),
//     and run acc = fma(sum, d_w * d_x, acc) per element with packed FP32 (FMUL2 / FFMA2 / FADD2) -- the reference's own
//     arithmetic up to the order of the eight partial f32 lanes of its AVX2 kernel.
// Data layout (repacked once at upload -- the 34-byte block_q8_0 is TMA-hostile): int8 planes A [M][K] and W [N][K] (K contiguous,
// 128-byte SWIZZLE_128B rows = 4 blocks per pipeline stage), scale planes Ad [M][K/32] f32 and WdT [K/32][N] f32 (transposed so a
// block's 128 column scales are one contiguous shared-memory row), all four moved by TMA under one mbarrier per stage.
//
// What this prototype is for: the per-element correction costs three FP32-pipe operations + one integer add
// per block against one K = 32 MMA that takes 64 clocks for the whole 128 x 128 tile, so the kernel is bound by the FP32 pipe at
// ~1/6 of the int8 tensor rate -- measured below the f16 tensor-core path the engine runs for q8_0 files.  Kept as a tested kernel
// and a measured number, not wired into the forward schedule.
#pragma once
#include "ptx.cuh"

namespace vitb200 {

struct Q8GemmParams
{
    int M, N, K;       // K % 128 == 0 (4 blocks per stage), N % 4 == 0
    const float *bias; // [N]
    float *out;        // [M][ldo] f32
    int ldo;
};

constexpr int Q8_BM = 128, Q8_BN = 128, Q8_BK = 128; // tile; K bytes (= int8 elements) per stage
constexpr int Q8_STAGES = 5;
constexpr int Q8_THREADS = 320; // warp 0 TMA, warp 1 MMA, warps 2-9 correction
constexpr int Q8_A_BYTES = Q8_BM * Q8_BK, Q8_W_BYTES = Q8_BN * Q8_BK;
constexpr int Q8_AD_BYTES = Q8_BM * 4 * 4, Q8_WD_BYTES = 4 * Q8_BN * 4;
constexpr int Q8_STAGE_BYTES = Q8_A_BYTES + Q8_W_BYTES + Q8_AD_BYTES + Q8_WD_BYTES;
constexpr int Q8_SMEM_BYTES = 1024 + Q8_STAGES * Q8_STAGE_BYTES + 256;

namespace ptx_q8 {
// D[tmem, s32] (+)= A[smem desc, s8] * B[smem desc, s8]: kind::i8, K = 32 per instruction
__device__ __forceinline__ void tcgen05_mma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// instruction descriptor for kind::i8 (cute/arch/mma_sm100_desc.hpp): c_format S32 = 2 at bits [4,6), a/b format INT8 = 1 at
// bits 7 / 10, both operands K-major, N >> 3 at [17,23), M >> 4 at [24,29)
__host__ __device__ constexpr inline uint32_t umma_idesc_i8(int M, int N)
{
    return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
} // namespace ptx_q8

// f32 activations [M][K] -> q8_0: int8 plane + per-block scale (the f16-rounded d, widened back to f32 as the dot product uses it).
// Bit-exact restatement of quantize_row_q8_0's AVX2 branch (ggml-quants.c:722-742): IEEE divisions, round-half-even.
// Eight lanes per block (one float4 each); a warp covers 128 consecutive elements per step.
__global__ void quantize_q8_0_kernel(const float *__restrict__ x, int8_t *__restrict__ q, float *__restrict__ d_out, long long n_blocks)
{
    const long long gthread = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long blk = gthread >> 3;
    const int sub = (int)(gthread & 7);
    const bool ok = blk < n_blocks;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (ok) v = __ldg(reinterpret_cast<const float4 *>(x + blk * 32) + sub);
    float amax = fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w)));
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
    const float d = __fdiv_rn(amax, 127.f);
    const float id = amax != 0.0f ? __fdiv_rn(127.f, amax) : 0.0f;
    if (!ok) return;
    const int q0 = __float2int_rn(__fmul_rn(v.x, id)), q1 = __float2int_rn(__fmul_rn(v.y, id));
    const int q2 = __float2int_rn(__fmul_rn(v.z, id)), q3 = __float2int_rn(__fmul_rn(v.w, id));
    const uint32_t packed = (uint32_t)(q0 & 0xFF) | ((uint32_t)(q1 & 0xFF) << 8) | ((uint32_t)(q2 & 0xFF) << 16) | ((uint32_t)(q3 & 0xFF) << 24);
    reinterpret_cast<uint32_t *>(q + blk * 32)[sub] = packed;
    if (sub == 0) d_out[blk] = __half2float(__float2half_rn(d));
}

__global__ void __launch_bounds__(Q8_THREADS, 1)
gemm_q8_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
                       const __grid_constant__ CUtensorMap tmAd, const __grid_constant__ CUtensorMap tmWd, const Q8GemmParams p)
{
    extern __shared__ uint8_t q8_smem_raw[];
    const uint32_t smem_base = (ptx::smem_u32(q8_smem_raw) + 1023u) & ~1023u;
    uint8_t *smem = q8_smem_raw + (smem_base - ptx::smem_u32(q8_smem_raw));
    auto sA = [&](int s) { return smem_base + (uint32_t)s * Q8_STAGE_BYTES; };
    auto sW = [&](int s) { return sA(s) + Q8_A_BYTES; };
    auto sAd = [&](int s) { return sW(s) + Q8_W_BYTES; };
    auto sWd = [&](int s) { return sAd(s) + Q8_AD_BYTES; };
    const uint32_t bars = smem_base + Q8_STAGES * Q8_STAGE_BYTES;
    // barriers: full[S] (TMA bytes), empty_mma[S] (operands consumed), empty_sc[S] (scales consumed), slot_full[4], slot_empty[4], tmem ptr
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto empty_mma = [&](int s) { return bars + 8u * (Q8_STAGES + s); };
    auto empty_sc = [&](int s) { return bars + 8u * (2 * Q8_STAGES + s); };
    auto slot_full = [&](int j) { return bars + 8u * (3 * Q8_STAGES + j); };
    auto slot_empty = [&](int j) { return bars + 8u * (3 * Q8_STAGES + 4 + j); };
    const uint32_t tmem_ptr_addr = bars + 8u * (3 * Q8_STAGES + 8);
    volatile uint32_t *tmem_ptr_gen = reinterpret_cast<volatile uint32_t *>(smem + Q8_STAGES * Q8_STAGE_BYTES + 8 * (3 * Q8_STAGES + 8));

    const int warp_idx = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m_tiles = (p.M + Q8_BM - 1) / Q8_BM, n_tiles = (p.N + Q8_BN - 1) / Q8_BN;
    const int num_tiles = m_tiles * n_tiles;
    const int num_kb = p.K / Q8_BK;

    if (warp_idx == 0 && lane == 0)
    {
        ptx::prefetch_tensormap(&tmA);
        ptx::prefetch_tensormap(&tmW);
        ptx::prefetch_tensormap(&tmAd);
        ptx::prefetch_tensormap(&tmWd);
    }
    if (warp_idx == 1 && lane == 0)
    {
        for (int s = 0; s < Q8_STAGES; ++s)
        {
            ptx::mbar_init(full_bar(s), 1);
            ptx::mbar_init(empty_mma(s), 1);
            ptx::mbar_init(empty_sc(s), 8);
        }
        for (int j = 0; j < 4; ++j)
        {
            ptx::mbar_init(slot_full(j), 1);
            ptx::mbar_init(slot_empty(j), 8);
        }
        ptx::fence_barrier_init();
    }
    if (warp_idx == 2)
    {
        ptx::tcgen05_alloc(tmem_ptr_addr, 512);
        ptx::tcgen05_relinquish();
    }
    ptx::tcgen05_fence_before();
    __syncthreads();
    ptx::tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr_gen;

    if (warp_idx == 0)
    {
        // ===================== TMA producer: int8 operand tiles + both scale tiles of a stage under one barrier =====================
        if (lane == 0)
        {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x)
            {
                const int m0 = (tile / n_tiles) * Q8_BM, n0 = (tile % n_tiles) * Q8_BN;
                for (int kb = 0; kb < num_kb; ++kb)
                {
                    ptx::mbar_wait(empty_mma(stage), phase ^ 1);
                    ptx::mbar_wait(empty_sc(stage), phase ^ 1);
                    ptx::mbar_arrive_expect_tx(full_bar(stage), Q8_STAGE_BYTES);
                    ptx::tma_load_2d(sA(stage), &tmA, full_bar(stage), kb * Q8_BK, m0);
                    ptx::tma_load_2d(sW(stage), &tmW, full_bar(stage), kb * Q8_BK, n0);
                    ptx::tma_load_2d(sAd(stage), &tmAd, full_bar(stage), kb * 4, m0);
                    ptx::tma_load_2d(sWd(stage), &tmWd, full_bar(stage), n0, kb * 4);
                    if (++stage == Q8_STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
        __syncwarp();
    }
    else if (warp_idx == 1)
    {
        // ===================== MMA issuer: one K = 32 MMA per q8_0 block, each into its own accumulator slot =====================
        constexpr uint32_t idesc = ptx_q8::umma_idesc_i8(Q8_BM, Q8_BN);
        int stage = 0;
        uint32_t phase = 0;
        uint32_t n = 0; // stages issued so far: slot j is used once per stage
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x)
        {
            for (int kb = 0; kb < num_kb; ++kb, ++n)
            {
                ptx::mbar_wait(full_bar(stage), phase);
                ptx::tcgen05_fence_after();
                const uint64_t adesc = ptx::umma_desc_kmajor_sw128(sA(stage));
                const uint64_t bdesc = ptx::umma_desc_kmajor_sw128(sW(stage));
#pragma unroll
                for (int j = 0; j < 4; ++j)
                {
                    ptx::mbar_wait(slot_empty(j), (n & 1) ^ 1); // the correction warps have read the previous block out of this slot
                    ptx::tcgen05_fence_after();
                    if (ptx::elect_one())
                    {
                        // 32 int8 = 32 B along K inside the 128-B swizzle row: +2 in the (addr >> 4) field
                        ptx_q8::tcgen05_mma_i8(tmem_base + (uint32_t)(j * Q8_BN), adesc + 2 * j, bdesc + 2 * j, idesc, 0);
                        ptx::tcgen05_commit(slot_full(j));
                        if (j == 3) ptx::tcgen05_commit(empty_mma(stage));
                    }
                    __syncwarp();
                }
                if (++stage == Q8_STAGES) { stage = 0; phase ^= 1; }
            }
        }
        __syncwarp();
    }
    else
    {
        // ===================== correction warps: int32 block sums -> f32, x (d_w d_x), accumulate in registers =====================
        const int q = warp_idx & 3;          // TMEM lane quarter
        const int half = (warp_idx - 2) >> 2; // column half of the tile: 64 columns
        const int r = q * 32 + lane;          // row of the tile
        int stage = 0;
        uint32_t phase = 0;
        uint32_t n = 0;
        const uint64_t neg_magic = ptx::pack_f32x2(-12582912.0f, -12582912.0f);
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x)
        {
            const int m0 = (tile / n_tiles) * Q8_BM, n0 = (tile % n_tiles) * Q8_BN;
            uint64_t acc[32];
#pragma unroll
            for (int c = 0; c < 32; ++c) acc[c] = 0ull;
            for (int kb = 0; kb < num_kb; ++kb, ++n)
            {
                ptx::mbar_wait(full_bar(stage), phase); // scale tiles of this stage are in shared memory
                const float4 ad4 = *reinterpret_cast<const float4 *>(smem + (size_t)stage * Q8_STAGE_BYTES + Q8_A_BYTES + Q8_W_BYTES + r * 16);
                const float adj[4] = {ad4.x, ad4.y, ad4.z, ad4.w};
#pragma unroll
                for (int j = 0; j < 4; ++j)
                {
                    uint32_t v[64];
                    ptx::mbar_wait(slot_full(j), n & 1);
                    ptx::tcgen05_fence_after();
                    {
                        uint32_t(&v0)[32] = *reinterpret_cast<uint32_t(*)[32]>(&v[0]);
                        uint32_t(&v1)[32] = *reinterpret_cast<uint32_t(*)[32]>(&v[32]);
                        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(j * Q8_BN + half * 64);
                        ptx::tcgen05_ld_32x32b_x32(taddr, v0);
                        ptx::tcgen05_ld_32x32b_x32(taddr + 32, v1);
                    }
                    ptx::tcgen05_wait_ld();
                    ptx::tcgen05_fence_before();
                    __syncwarp();
                    if (lane == 0) ptx::mbar_arrive(slot_empty(j));
                    const float4 *wd4 = reinterpret_cast<const float4 *>(smem + (size_t)stage * Q8_STAGE_BYTES + Q8_A_BYTES + Q8_W_BYTES + Q8_AD_BYTES) +
                                        j * (Q8_BN / 4) + half * 16;
                    const uint64_t ad2 = ptx::pack_f32x2(adj[j], adj[j]);
#pragma unroll
                    for (int c4 = 0; c4 < 16; ++c4)
                    {
                        const float4 w = wd4[c4]; // warp-uniform address: broadcast
                        // exact int -> float: the bit pattern of 1.5 * 2^23 plus a |sum| < 2^22 integer IS the float 1.5 * 2^23 + sum
                        const uint64_t g0 = ptx::add_f32x2(ptx::pack_f32x2(__uint_as_float(v[4 * c4] + 0x4B400000u), __uint_as_float(v[4 * c4 + 1] + 0x4B400000u)), neg_magic);
                        const uint64_t g1 = ptx::add_f32x2(ptx::pack_f32x2(__uint_as_float(v[4 * c4 + 2] + 0x4B400000u), __uint_as_float(v[4 * c4 + 3] + 0x4B400000u)), neg_magic);
                        const uint64_t s0 = ptx::mul_f32x2(ptx::pack_f32x2(w.x, w.y), ad2); // d = d_w * d_x (ggml-quants.c:3521+)
                        const uint64_t s1 = ptx::mul_f32x2(ptx::pack_f32x2(w.z, w.w), ad2);
                        acc[2 * c4] = ptx::fma_f32x2(s0, g0, acc[2 * c4]);
                        acc[2 * c4 + 1] = ptx::fma_f32x2(s1, g1, acc[2 * c4 + 1]);
                    }
                }
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(empty_sc(stage));
                if (++stage == Q8_STAGES) { stage = 0; phase ^= 1; }
            }
            // bias + store: 64 consecutive f32 per thread-row
            const int grow = m0 + r;
            if (grow < p.M)
            {
                float *orow = p.out + (size_t)grow * p.ldo + n0 + half * 64;
#pragma unroll
                for (int c4 = 0; c4 < 16; ++c4)
                {
                    const int gcol = n0 + half * 64 + c4 * 4;
                    if (gcol < p.N)
                    {
                        const float4 b = __ldg(reinterpret_cast<const float4 *>(p.bias + gcol));
                        float a0, a1, a2, a3;
                        ptx::unpack_f32x2(acc[2 * c4], a0, a1);
                        ptx::unpack_f32x2(acc[2 * c4 + 1], a2, a3);
                        *reinterpret_cast<float4 *>(orow + c4 * 4) = make_float4(__fadd_rn(a0, b.x), __fadd_rn(a1, b.y), __fadd_rn(a2, b.z), __fadd_rn(a3, b.w));
                    }
                }
            }
        }
    }

    ptx::tcgen05_fence_before();
    __syncthreads();
    if (warp_idx == 2)
    {
        ptx::tcgen05_fence_after();
        ptx::tcgen05_dealloc(tmem_base, 512);
    }
}

} // namespace vitb200
