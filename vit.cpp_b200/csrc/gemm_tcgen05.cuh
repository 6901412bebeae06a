// gemm_tcgen05.cuh -- the linear-layer kernel: C[M,N] = A[M,K] * W[N,K]^T with a fused epilogue.
//
// Replaces ggml_mul_mat + ggml_add_inplace (+ ggml_gelu) (+ residual ggml_add) of the reference graph
// (reference vit.cpp:820-821, 868-873, 889-900, 927-928, and the conv GEMM of vit.cpp:772-797) and the CPU
// kernel behind them (ggml.c:9388-9597 with ggml_vec_dot_f16, ggml.c:1200-1236).
// Numerics follow the reference recipe: A = activations already rounded to f16 (RNE, ggml.c:9493-9506),
// W = the f16 weights as stored in the model file (ne0 = K contiguous, vit.cpp:531-543), fp32 accumulation.
//
// Structure (sm_100a): persistent CTA PAIRS (cluster 2x1x1, one pair per TPC; CG = 1 keeps a single-CTA variant), each pair
// computing 256 x BN tiles of C with tcgen05.mma.cta_group::2 (M = 256: 128 rows per CTA, W tile split along N between the CTAs):
//   warp 0        TMA producer : cp.async.bulk.tensor -- this CTA's 128 x 64 A tile + its HALF of the BN x 64 W tile per stage
//                                (SWIZZLE_128B), transaction bytes of both CTAs collected on the leader's mbarrier
//   warp 1        MMA issuer   : leader CTA only; warp-uniform loop, one elected lane issues 4 x tcgen05.mma.kind::f16 (K = 16) per
//                                stage; accumulators in TMEM (2 x BN columns, double buffered against the epilogue); smem slots and
//                                accumulators are released to BOTH CTAs by multicast tcgen05.commit
//   warps 2-5     (EPI_PATCH_GATHER_F32 only) A producers: gather f32 pixels -> f16 swizzled operand tiles, no im2col buffer
//   next 8 warps  epilogue     : tcgen05.ld -> registers -> bias / GELU / hi-lo split -> XOR-swizzled smem transpose -> 128-bit
//                                coalesced stores (two warps per TMEM lane quarter, alternating column passes)
//   or 4 warps    residual epilogue (proj, fc2): per-warp TMA ring streams the f32 residual tile in, adds in place, TMA-stores
// Pipelines: smem ring full/empty mbarriers (TMA <-> MMA), TMEM full/empty mbarriers (MMA <-> epilogue).
#pragma once
#include "ptx.cuh"

namespace vitb200 {

enum GemmEpilogue
{
    EPI_BIAS_F16 = 0,       // out f16 = f16(acc + bias)                              (qkv)
    EPI_BIAS_GELU_F16 = 1,  // out f16 = f16(gelu_tanh(f32(f16(acc + bias))))           (fc1; ggml.c:1418-1441)
    EPI_BIAS_RESID_F32 = 2, // out f32 = (acc + bias) + resid                         (proj, fc2; vit.cpp:873,900)
    EPI_PATCH_F32 = 3,      // out f32[token row] = (acc + bias) + pos_embed           (patch embed; vit.cpp:773-797)
    EPI_BIAS_F32 = 4,       // out f32 = acc + bias                                   (head logits)
    EPI_PATCH_GATHER_F32 = 5, // EPI_PATCH_F32 with the A operand gathered straight from the f32 HWC pixels (P = 16): no im2col buffer
    EPI_BIAS_F16_HILO = 6,  // x = acc + bias kept to ~22 significant bits as TWO f16 tensors: out = hi = f16(x), out2 = lo = f16(x - hi)
                            // (qkv in front of the tcgen05 attention: the reference feeds f32 q, k, v to both attention mat-muls,
                            // vit.cpp:848,858; hi + lo lets the f16 tensor cores reproduce that with split-precision MMAs)
};

struct GemmParams
{
    int M, N, K;        // valid extents (rows of A / rows of W / reduction)
    const float *bias;  // [N]
    void *out;          // f16 or f32, row-major, leading dimension ldo
    void *out2;         // EPI_BIAS_F16_HILO: the lo halves, same shape as out
    int ldo;
    unsigned long long store_policy; // f16 epilogues: L2 eviction hint of the output's TMA stores (0 = none; ptx::L2_EVICT_FIRST for outputs that are
                        // read once by the next kernel and are larger than the L2 anyway)
    int headmajor;      // f16 epilogues: 1 = the output is HEAD-MAJOR, [N / 64 planes][M rows][64] (tmX / tmO2 are 3-D maps: column, row, plane):
                        // every 64-column pass is one head's slice, stored as ONE contiguous 4-KB box instead of 32 row pieces of 128 B that lie
                        // ldo * 2 bytes apart -- what the attention kernels then read back as contiguous tiles (qkv)
    const float *resid; // EPI_BIAS_RESID_F32: [M][ldo] (may alias out)
    const float *pos;   // EPI_PATCH_F32: pos_embed [ntok][N]
    int np, ntok;       // EPI_PATCH_F32: patches / tokens per image
    const float *img;   // EPI_PATCH_GATHER_F32: images [B][S][S][3] f32 (image_f32 layout, vit.h:98-103)
    int S, G;           // EPI_PATCH_GATHER_F32: image side, patches per side
    // EPI_BIAS_RESID_F32 with a fused LayerNorm (N == hidden size, N % 128 == 0, N <= 1024): the LayerNorm that follows the residual add
    // (vit.cpp:881-885 after proj, vit.cpp:808-812 of the next block after fc2) is applied to every 32-row group as soon as ALL of
    // its column tiles have been stored -- by whichever epilogue warp stores the last one (ln_count: one counter per 32-row group,
    // zero on entry and left zero) -- while the rows are still in L2; ln_out receives the f16 rows [M][N].  NULL = no fusion.
    __half *ln_out;
    const float *ln_w, *ln_b;
    int *ln_count;
    float ln_eps;
    int ln_dbg;         // dev knob (VITB200_LN_DBG): bit 0 = epilogue side without the GPU-scope fences, bit 1 = LayerNorm side without
};

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;

// The residual epilogue (proj, fc2) streams the f32 residual tile through a per-warp TMA ring (kResidRing slots of
// 32 rows x 32 columns = 4 KB) and stores the result with TMA from the same slot, so ~64 KB of residual loads are in
// flight per SM without holding registers; it pays for the ring with one fewer operand stage.
// CG = CTAs per MMA (cta_group): 2 = a CTA pair on one TPC computes a 256 x BN tile with M = 256 tcgen05.mma; each CTA
// stages its own 128 A rows and HALF of the W tile (the pair shares both halves), which halves the per-SM shared-memory
// traffic of the B operand -- the 1-CTA kernel is shared-memory-bandwidth bound at ~70 % tensor-pipe utilisation.
// kDeepK (fc2, K = 4 D): the main loop of a tile is long, so a shallower residual ring (2 loads in flight per warp) still keeps up
// and its shared memory buys a fifth operand stage, which is what the tensor pipe is short of there.
// kF16Out (qkv, fc1): the bias is applied in the row-per-thread domain, i.e. every lane needs the same 64 values per column pass.  With
// 227 KB of shared memory carved out the SM has practically no L1, so reading them with __ldg put an L2 round trip (~300 clocks) in
// front of every 8-column chunk (ncu: long-scoreboard stalls on LDG.CONSTANT dominated the epilogue and capped the split-precision
// qkv epilogue at 56 % tensor-pipe activity).  Each epilogue warp therefore keeps the <= 128 bias values of its column passes in a
// private 512-B shared-memory strip, loaded one tile ahead (one 16-B load per lane, latency hidden behind the accumulator wait).
template <int BN, bool kResid, int CG, bool kGather = false, bool kDeepK = false, bool kF16Out = false>
struct GemmCfg
{
    static constexpr int kResidRing = kDeepK ? 3 : 5;
    // epilogue warps: 4 for the TMA-ring residual epilogue (HBM-bound), 8 otherwise (two per TMEM lane quarter, splitting
    // the columns) so the ALU-heavy f16 epilogues (bias, GELU, packing) have two warps per SM sub-partition to overlap
    static constexpr int kEpiWarps = kResid ? 4 : 8;
    static constexpr int kGatherWarps = kGather ? 4 : 0; // patch-embedding A producers (thread = one patch row of the tile)
    static constexpr int kFirstEpiWarp = 2 + kGatherWarps;
    static constexpr int kLnWarps = kResid ? 8 : 0;      // residual epilogue: dedicated LayerNorm warps behind a shared-memory work queue
    static constexpr int kLnItemRows = 16;               // rows per work item (half a 32-row group)
    static constexpr int kThreads = 32 * (kFirstEpiWarp + kEpiWarps + kLnWarps);
    static constexpr int A_BYTES = GEMM_BM * GEMM_BK * 2;
    static constexpr int B_ROWS = BN / CG;               // W rows staged by this CTA
    static constexpr int B_BYTES = B_ROWS * GEMM_BK * 2;
    static constexpr int STAGE_BYTES = kResid ? 4 * kResidRing * 4096 : kEpiWarps * 4096; // per epilogue warp: ring or transpose buffer
    static constexpr int BAR_BYTES = 512;
    // per-warp bias strips: 128 values (f16 epilogues), BN values (residual epilogue); + LayerNorm weight and bias (2 x 1024 floats)
    // for the residual epilogue's fused LayerNorm
    static constexpr int kLnMaxD = 1024;
    static constexpr int kLnQueue = 256;                 // work-queue slots (8-row items), with flow control
    static constexpr int BIAS_BYTES = kF16Out ? kEpiWarps * 512 : (kResid ? 4 * BN * 4 + 2 * kLnMaxD * 4 + kLnQueue * 4 + 64 : 0);
    static constexpr int kSmemLimit = 232448; // 227 KB per CTA
    static constexpr int kStagesFit = (kSmemLimit - 1024 - STAGE_BYTES - BAR_BYTES - BIAS_BYTES) / (A_BYTES + B_BYTES);
    static constexpr int kStagesFit8 = kStagesFit > 8 ? 8 : kStagesFit;
    static constexpr int kStages = kGather ? kStagesFit8 / 3 * 3 : kStagesFit8; // operand ring depth: whatever fits (gather fills 3 at a time)
    static constexpr int SMEM_BYTES = 1024 /*align slack*/ + kStages * (A_BYTES + B_BYTES) + STAGE_BYTES + BAR_BYTES + BIAS_BYTES;
    static constexpr int TMEM_COLS = 2 * BN;
    static_assert(SMEM_BYTES <= 232448, "exceeds the 227 KB per-CTA shared memory limit");
};

// GELU, tanh form (the reference's formula, ggml.c:1418-1424) on an f16-valued input, evaluated in f32 as
//   0.5 x (1 + tanh(u)) = x / (1 + e^{-2u}),   u = sqrt(2/pi) x (1 + 0.044715 x^2)
// (algebraically identical; 5 FMA-pipe ops + MUFU.EX2 + MUFU.RCP, relative error ~3e-7, i.e. the f16 rounding the caller
// applies next (table semantics, ggml.c:2197) differs from the host table in well under 0.1 % of inputs, by one ulp).
__device__ __forceinline__ float gelu_tanh_f32(float x)
{
    // -2u log2(e) = x (k0 + k1 x^2), k0 = -2 sqrt(2/pi) log2(e), k1 = 0.044715 k0: FMUL, FFMA, FMUL, EX2, FADD, RCP, FMUL
    const float w = fmaf(x * x, -0.10294323958083856f, -2.3022081981625516f);
    const float e = ptx::ex2_approx(x * w);                      // e^{-2u}
    return x * ptx::rcp_approx(1.0f + e);
}

// Two elements at a time with Blackwell's packed FP32 instructions (FMUL2 / FFMA2 / FADD2): the same separately rounded IEEE
// operations as gelu_tanh_f32 on each lane (bit-identical results), half the issue slots for the five FMA-pipe steps; the two
// MUFU ops per element stay scalar.  tools/microbench/chain_r02.cu: 36.9 instead of 39.5 clocks per element per warp at two
// epilogue warps per sub-partition.
__device__ __forceinline__ void gelu_tanh_f32x2(float x0, float x1, float &y0, float &y1)
{
    const uint64_t x = ptx::pack_f32x2(x0, x1);
    const uint64_t w = ptx::fma_f32x2(ptx::mul_f32x2(x, x), ptx::pack_f32x2(-0.10294323958083856f, -0.10294323958083856f),
                                      ptx::pack_f32x2(-2.3022081981625516f, -2.3022081981625516f));
    float a0, a1;
    ptx::unpack_f32x2(ptx::mul_f32x2(x, w), a0, a1);
    float d0, d1;
    ptx::unpack_f32x2(ptx::add_f32x2(ptx::pack_f32x2(ptx::ex2_approx(a0), ptx::ex2_approx(a1)), ptx::pack_f32x2(1.0f, 1.0f)), d0, d1);
    ptx::unpack_f32x2(ptx::mul_f32x2(x, ptx::pack_f32x2(ptx::rcp_approx(d0), ptx::rcp_approx(d1))), y0, y1);
}

// One warp normalises one row of X (f32, already complete in L2) into the f16 A operand of the next GEMM: the arithmetic of
// layernorm_f16_kernel (kernels.cuh) -- f32 two-pass mean / biased variance, then ((x - mean) * scale) * w + b as three separately
// rounded operations (the reference's NORM, MUL, ADD nodes, ggml.c:8959-9008, vit.cpp:808-812), RNE to f16 -- with w and b read from
// shared memory and the row read past L1 (ld.global.cg: other SMs wrote it).  nv = D / 128 float4 per lane (<= 8).
__device__ __forceinline__ void ln_row_load(const float *xrow, int nv, int lane, float4 (&v)[8])
{
#pragma unroll
    for (int j = 0; j < 8; ++j)
        if (j < nv) v[j] = __ldcg(reinterpret_cast<const float4 *>(xrow) + lane + 32 * j);
}
__device__ __forceinline__ void ln_row_finish(float4 (&v)[8], int nv, int D, int lane, const float4 *w4, const float4 *b4, float eps, __half *yrow)
{
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j)
        if (j < nv) sum += (v[j].x + v[j].y) + (v[j].z + v[j].w);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float mean = sum / (float)D;
    float sq = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j)
        if (j < nv)
        {
            v[j].x -= mean; v[j].y -= mean; v[j].z -= mean; v[j].w -= mean;
            sq += (v[j].x * v[j].x + v[j].y * v[j].y) + (v[j].z * v[j].z + v[j].w * v[j].w);
        }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    const float scale = 1.0f / sqrtf(sq / (float)D + eps);
    uint2 *yr = reinterpret_cast<uint2 *>(yrow);
#pragma unroll
    for (int j = 0; j < 8; ++j)
        if (j < nv)
        {
            const int i = lane + 32 * j;
            const float4 ww = w4[i], bb = b4[i];
            const __half2 h0 = __floats2half2_rn(__fadd_rn(__fmul_rn(__fmul_rn(v[j].x, scale), ww.x), bb.x), __fadd_rn(__fmul_rn(__fmul_rn(v[j].y, scale), ww.y), bb.y));
            const __half2 h1 = __floats2half2_rn(__fadd_rn(__fmul_rn(__fmul_rn(v[j].z, scale), ww.z), bb.z), __fadd_rn(__fmul_rn(__fmul_rn(v[j].w, scale), ww.w), bb.w));
            uint2 u;
            u.x = *reinterpret_cast<const uint32_t *>(&h0);
            u.y = *reinterpret_cast<const uint32_t *>(&h1);
            yr[i] = u;
        }
}

template <int BN, int EPI, int DEEPK, int CG>
__global__ void __launch_bounds__((GemmCfg<BN, EPI == EPI_BIAS_RESID_F32, CG, EPI == EPI_PATCH_GATHER_F32, DEEPK != 0,
                                           EPI == EPI_BIAS_F16 || EPI == EPI_BIAS_GELU_F16 || EPI == EPI_BIAS_F16_HILO>::kThreads), 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmO2, const GemmParams p)
// tmX: EPI_BIAS_RESID_F32 -- the f32 residual/output stream (32 x 32 boxes); f16 epilogues -- the f16 OUTPUT tensor (64-column x 32-row
// boxes, SWIZZLE_128B), tmO2 -- the lo tensor of EPI_BIAS_F16_HILO.  Unused maps are ignored.
{
    constexpr bool kResid = (EPI == EPI_BIAS_RESID_F32);
    constexpr bool kGather = (EPI == EPI_PATCH_GATHER_F32);
    constexpr bool kPatch = (EPI == EPI_PATCH_F32 || EPI == EPI_PATCH_GATHER_F32);
    constexpr bool kOutF16 = (EPI == EPI_BIAS_F16 || EPI == EPI_BIAS_GELU_F16 || EPI == EPI_BIAS_F16_HILO);
    constexpr bool kHiLo = (EPI == EPI_BIAS_F16_HILO);
    using Cfg = GemmCfg<BN, kResid, CG, kGather, DEEPK != 0, kOutF16>;
    static_assert(!kGather || (CG == 2 && Cfg::kStages % 3 == 0 && Cfg::kStages >= 3), "gathered patch embedding: CTA pairs, stages in threes");
    constexpr int TILE_M = GEMM_BM * CG; // rows of C per CTA group
    constexpr int kStages = Cfg::kStages;

    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u; // SWIZZLE_128B wants 1024-B aligned tiles
    uint8_t *smem = smem_raw + (smem_base - ptx::smem_u32(smem_raw));
    const uint32_t sA = smem_base;
    const uint32_t sB = sA + kStages * Cfg::A_BYTES;
    uint8_t *stg_base = smem + kStages * (Cfg::A_BYTES + Cfg::B_BYTES);
    const uint32_t bars = smem_base + kStages * (Cfg::A_BYTES + Cfg::B_BYTES) + Cfg::STAGE_BYTES;
    // barrier layout (8 B each): full[kStages], empty[kStages], tmem_full[2], tmem_empty[2], then tmem ptr
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto empty_bar = [&](int s) { return bars + 8u * (kStages + s); };
    auto tfull_bar = [&](int a) { return bars + 8u * (2 * kStages + a); };
    auto tempty_bar = [&](int a) { return bars + 8u * (2 * kStages + 2 + a); };
    const uint32_t tmem_ptr_addr = bars + 8u * (2 * kStages + 4);
    // residual ring "full" barriers: [4 epilogue warps][kResidRing], after the tmem pointer slot
    auto rfull_bar = [&](int w, int r) { return bars + 8u * (2 * kStages + 6 + w * Cfg::kResidRing + r); };
    volatile uint32_t *tmem_ptr_gen = reinterpret_cast<volatile uint32_t *>(smem + kStages * (Cfg::A_BYTES + Cfg::B_BYTES) + Cfg::STAGE_BYTES + 8 * (2 * kStages + 4));

    const int warp_idx = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    const int m_tiles = (p.M + TILE_M - 1) / TILE_M;
    const int n_tiles = (p.N + BN - 1) / BN;
    const int num_tiles = m_tiles * n_tiles;
    const int num_kb = (p.K + GEMM_BK - 1) / GEMM_BK;
    const uint32_t cta_rank = (CG == 2) ? ptx::cluster_ctarank() : 0u; // 0 = leader of the pair
    const int group_id = blockIdx.x / CG, num_groups = gridDim.x / CG;  // persistent tile loop runs per CTA group
    if constexpr (CG == 2) ptx::cluster_sync(); // both CTAs resident before the pair-wide TMEM allocation

    if (warp_idx == 0 && lane == 0)
    {
        ptx::prefetch_tensormap(&tmA);
        ptx::prefetch_tensormap(&tmB);
        if constexpr (kResid || kOutF16) ptx::prefetch_tensormap(&tmX);
        if constexpr (kHiLo) ptx::prefetch_tensormap(&tmO2);
    }
    if (warp_idx == 1 && lane == 0)
    {
        for (int s = 0; s < kStages; ++s)
        {
            // one (remote) arrival per producer of the pair (+ one per gather warp); the leader's barrier collects all bytes
            ptx::mbar_init(full_bar(s), CG * (1 + Cfg::kGatherWarps));
            ptx::mbar_init(empty_bar(s), 1);
        }
        for (int a = 0; a < 2; ++a)
        {
            ptx::mbar_init(tfull_bar(a), 1);
            ptx::mbar_init(tempty_bar(a), CG * Cfg::kEpiWarps); // one arrival per epilogue warp of the group (leader's barrier)
        }
        if constexpr (kResid)
        {
            for (int w = 0; w < 4; ++w)
                for (int r = 0; r < Cfg::kResidRing; ++r) ptx::mbar_init(rfull_bar(w, r), 1);
            // LayerNorm work-queue counters (reserved, published, claimed, consumed)
            int *qctl = reinterpret_cast<int *>(smem + kStages * (Cfg::A_BYTES + Cfg::B_BYTES) + Cfg::STAGE_BYTES + Cfg::BAR_BYTES + 4 * BN * 4 +
                                                2 * Cfg::kLnMaxD * 4 + Cfg::kLnQueue * 4);
            qctl[0] = qctl[1] = qctl[2] = qctl[3] = 0;
        }
        ptx::fence_barrier_init();
    }
    if (warp_idx == Cfg::kFirstEpiWarp)
    {
        if constexpr (CG == 2) { ptx::tcgen05_alloc_cg2(tmem_ptr_addr, Cfg::TMEM_COLS); ptx::tcgen05_relinquish_cg2(); }
        else { ptx::tcgen05_alloc(tmem_ptr_addr, Cfg::TMEM_COLS); ptx::tcgen05_relinquish(); }
    }
    ptx::tcgen05_fence_before();
    if constexpr (CG == 2) ptx::cluster_sync(); else __syncthreads(); // barrier inits visible to the peer before any remote arrive
    ptx::tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr_gen;
    // prologue done (it overlapped the previous kernel's tail under PDL); from here on the kernel touches upstream data
    ptx::grid_dep_launch();
    ptx::grid_dep_wait();

    if (warp_idx == 0)
    {
        // ===================== TMA producer =====================
        if (lane == 0)
        {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = group_id; tile < num_tiles; tile += num_groups)
            {
                const int m_blk = tile / n_tiles, n_blk = tile % n_tiles;
                const int a_row = m_blk * TILE_M + (int)cta_rank * GEMM_BM;   // this CTA's 128 rows of A
                const int b_row = n_blk * BN + (int)cta_rank * Cfg::B_ROWS;   // this CTA's share of the W tile
                for (int kb = 0; kb < num_kb; ++kb)
                {
                    ptx::mbar_wait(empty_bar(stage), phase ^ 1); // own slot free (the MMA commit is multicast to both CTAs)
                    if constexpr (kGather)
                    {
                        // W only; k-blocks visited as (ky group, channel): k = c*256 + kyg*64 (K order c*P*P + ky*P + kx, ggml.c:11597)
                        const int koff = (kb % 3) * 256 + (kb / 3) * 64;
                        if (cta_rank == 0) ptx::mbar_arrive_expect_tx(full_bar(stage), 2 * Cfg::B_BYTES);
                        else ptx::mbar_arrive_remote(full_bar(stage), 0);
                        ptx::tma_load_2d_cg2(sB + stage * Cfg::B_BYTES, &tmB, full_bar(stage), koff, b_row);
                    }
                    else if constexpr (CG == 2)
                    {
                        // all bytes of the pair are accounted on the LEADER's full barrier
                        if (cta_rank == 0) ptx::mbar_arrive_expect_tx(full_bar(stage), 2 * (Cfg::A_BYTES + Cfg::B_BYTES));
                        else ptx::mbar_arrive_remote(full_bar(stage), 0);
                        ptx::tma_load_2d_cg2(sA + stage * Cfg::A_BYTES, &tmA, full_bar(stage), kb * GEMM_BK, a_row);
                        ptx::tma_load_2d_cg2(sB + stage * Cfg::B_BYTES, &tmB, full_bar(stage), kb * GEMM_BK, b_row);
                    }
                    else
                    {
                        ptx::mbar_arrive_expect_tx(full_bar(stage), Cfg::A_BYTES + Cfg::B_BYTES);
                        ptx::tma_load_2d(sA + stage * Cfg::A_BYTES, &tmA, full_bar(stage), kb * GEMM_BK, a_row);
                        ptx::tma_load_2d(sB + stage * Cfg::B_BYTES, &tmB, full_bar(stage), kb * GEMM_BK, b_row);
                    }
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
            }
        }
        __syncwarp();
    }
    else if (warp_idx == 1)
    {
        // ===================== MMA issuer: whole warp runs the loop (uniform), one elected lane issues =====================
        if (cta_rank == 0) // the leader CTA issues for the whole group
        {
            constexpr uint32_t idesc = ptx::umma_idesc_f16(TILE_M, BN, /*a=f16*/ 0, /*b=f16*/ 0);
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int tile = group_id; tile < num_tiles; tile += num_groups, ++it)
            {
                const int as = it & 1;
                const uint32_t aphase = (it >> 1) & 1;
                ptx::mbar_wait(tempty_bar(as), aphase ^ 1); // epilogue has drained this accumulator
                ptx::tcgen05_fence_after();
                const uint32_t tmem_d = tmem_base + as * BN;
                for (int kb = 0; kb < num_kb; ++kb)
                {
                    ptx::mbar_wait(full_bar(stage), phase); // TMA bytes have landed
                    ptx::tcgen05_fence_after();
                    const uint64_t adesc = ptx::umma_desc_kmajor_sw128(sA + stage * Cfg::A_BYTES);
                    const uint64_t bdesc = ptx::umma_desc_kmajor_sw128(sB + stage * Cfg::B_BYTES);
                    if (ptx::elect_one())
                    {
#pragma unroll
                        for (int k = 0; k < GEMM_BK / 16; ++k)
                        {
                            // advance 16 f16 = 32 B along K inside the 128-B swizzle row: +2 in the (addr >> 4) field
                            if constexpr (CG == 2) ptx::tcgen05_mma_f16_cg2(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
                            else ptx::tcgen05_mma_f16(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
                        }
                        if constexpr (CG == 2)
                        {
                            ptx::tcgen05_commit_cg2(empty_bar(stage), 3); // frees the slot in BOTH CTAs when these MMAs retire
                            if (kb == num_kb - 1) ptx::tcgen05_commit_cg2(tfull_bar(as), 3); // accumulator complete, both epilogues
                        }
                        else
                        {
                            ptx::tcgen05_commit(empty_bar(stage)); // frees the smem slot when these MMAs retire
                            if (kb == num_kb - 1) ptx::tcgen05_commit(tfull_bar(as)); // accumulator complete
                        }
                    }
                    __syncwarp();
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
            }
        }
        __syncwarp();
    }
    else if (kGather && warp_idx < Cfg::kFirstEpiWarp)
    {
        // ===================== A producers: im2col-free gather (patch size 16) =====================
        // Replaces ggml_im2col (ggml.c:11528-11608) + the HWC->CHW copy (vit.cpp:759-768): stride == kernel makes im2col a pure
        // permutation, so each thread owns one patch (one A row): per group of 4 kernel rows it reads the 4 x 16 pixels x 3
        // channels (4 x 192 contiguous bytes, 128-bit loads), rounds to f16 (RNE, ggml.c:11599) and writes the three channels'
        // 64-wide K slices into three consecutive pipeline stages in the UMMA SWIZZLE_128B layout.
        if constexpr (kGather)
        {
            const int r = (warp_idx - 2) * 32 + lane; // row of this CTA's 128 x 64 A tile
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = group_id; tile < num_tiles; tile += num_groups)
            {
                const int m_blk = tile / n_tiles;
                const int m = m_blk * TILE_M + (int)cta_rank * GEMM_BM + r; // patch index (row of the virtual im2col matrix)
                const bool valid = m < p.M;
                const int img_i = valid ? m / p.np : 0, pp = valid ? m - img_i * p.np : 0;
                const int py = pp / p.G, px = pp - py * p.G;
                const float *patch0 = p.img + (((size_t)img_i * p.S + (size_t)py * 16) * p.S + (size_t)px * 16) * 3;
                for (int kg = 0; kg < num_kb / 3; ++kg) // kernel rows 4*kg .. 4*kg+3
                {
#pragma unroll
                    for (int u = 0; u < 3; ++u) ptx::mbar_wait(empty_bar(stage + u), phase ^ 1);
#pragma unroll
                    for (int ky = 0; ky < 4; ++ky)
                    {
                        float v[48];
                        const float4 *src = reinterpret_cast<const float4 *>(patch0 + (size_t)(kg * 4 + ky) * p.S * 3);
#pragma unroll
                        for (int i = 0; i < 12; ++i)
                        {
                            const float4 f = valid ? __ldg(src + i) : make_float4(0.f, 0.f, 0.f, 0.f);
                            v[4 * i] = f.x; v[4 * i + 1] = f.y; v[4 * i + 2] = f.z; v[4 * i + 3] = f.w;
                        }
#pragma unroll
                        for (int c = 0; c < 3; ++c)
                        {
#pragma unroll
                            for (int hh = 0; hh < 2; ++hh) // pixels 0..7 / 8..15 of the kernel row -> one 16-byte chunk each
                            {
                                uint32_t w[4];
#pragma unroll
                                for (int e = 0; e < 4; ++e)
                                {
                                    const __half2 h2 = __floats2half2_rn(v[(hh * 8 + 2 * e) * 3 + c], v[(hh * 8 + 2 * e + 1) * 3 + c]);
                                    w[e] = *reinterpret_cast<const uint32_t *>(&h2);
                                }
                                const int chunk = ky * 2 + hh; // 16-B chunk inside the 128-B K slice of this row
                                uint4 *dst = reinterpret_cast<uint4 *>(smem + (size_t)(stage + c) * Cfg::A_BYTES + r * 128 + ((chunk ^ (r & 7)) << 4));
                                *dst = make_uint4(w[0], w[1], w[2], w[3]);
                            }
                        }
                    }
                    ptx::fence_proxy_async_smem(); // generic-proxy writes -> visible to the tensor core's (async proxy) reads
                    __syncwarp();
                    if (lane == 0)
                    {
#pragma unroll
                        for (int u = 0; u < 3; ++u)
                        {
                            if (cta_rank == 0) ptx::mbar_arrive(full_bar(stage + u));
                            else ptx::mbar_arrive_remote_release(full_bar(stage + u), 0);
                        }
                    }
                    __syncwarp();
                    stage += 3;
                    if (stage == kStages) { stage = 0; phase ^= 1; }
                }
            }
        }
    }
    else if (kResid && warp_idx >= Cfg::kFirstEpiWarp + Cfg::kEpiWarps)
    {
        // ===================== LayerNorm warps (residual epilogue with a fused LayerNorm) =====================
        // Work items are 16-row halves of 32-row groups whose every column tile has been stored (published by the epilogue warps
        // below through a shared-memory queue); any LayerNorm warp takes the next item, reads the rows from L2 (they were written
        // moments ago), normalises them (ln_row_finish) and writes the f16 A operand of the next GEMM.  A -1 item ends a warp.
        if constexpr (kResid)
        {
            if (p.ln_out != nullptr)
            {
                uint8_t *extra = smem + kStages * (Cfg::A_BYTES + Cfg::B_BYTES) + Cfg::STAGE_BYTES + Cfg::BAR_BYTES;
                float *lnw = reinterpret_cast<float *>(extra + 4 * BN * 4);
                const float4 *lnw4 = reinterpret_cast<const float4 *>(lnw);
                const float4 *lnb4 = lnw4 + Cfg::kLnMaxD / 4;
                volatile int *qbuf = reinterpret_cast<volatile int *>(extra + 4 * BN * 4 + 2 * Cfg::kLnMaxD * 4);
                int *qctl = const_cast<int *>(qbuf) + Cfg::kLnQueue; // [0] reserved, [1] published (tail), [2] claimed (head), [3] consumed
                const int lw = warp_idx - (Cfg::kFirstEpiWarp + Cfg::kEpiWarps);
                for (int i = lw * 32 + lane; i < p.N; i += 32 * Cfg::kLnWarps) { lnw[i] = __ldg(p.ln_w + i); lnw[Cfg::kLnMaxD + i] = __ldg(p.ln_b + i); }
                ptx::named_bar_sync(1, 32 * Cfg::kLnWarps); // the LayerNorm warps only
                const int ln_nv = p.N >> 7;
                const float *X = reinterpret_cast<const float *>(p.out);
                for (;;)
                {
                    int row0 = 0;
                    if (lane == 0)
                    {
                        const int idx = atomicAdd(&qctl[2], 1);
                        while (*reinterpret_cast<volatile int *>(&qctl[1]) <= idx) __nanosleep(100);
                        row0 = qbuf[idx % Cfg::kLnQueue];
                        __threadfence_block();
                        atomicAdd(&qctl[3], 1);
                    }
                    row0 = __shfl_sync(0xffffffffu, row0, 0);
                    if (row0 < 0) break;
                    // Three rows in flight per warp: while the GEMM saturates HBM a load takes several microseconds to come back, and
                    // the eight warps together must keep ~10 rows (30 KB) in flight per SM to follow the rate at which rows complete.
                    const int rows = (p.ln_dbg & 4) ? 0 : min(Cfg::kLnItemRows, p.M - row0);
                    float4 va[8], vb[8], vc[8];
                    if (rows > 0) ln_row_load(X + (size_t)row0 * p.ldo, ln_nv, lane, va);
                    if (rows > 1) ln_row_load(X + (size_t)(row0 + 1) * p.ldo, ln_nv, lane, vb);
                    if (rows > 2) ln_row_load(X + (size_t)(row0 + 2) * p.ldo, ln_nv, lane, vc);
                    for (int r = 0; r < rows; r += 3)
                    {
                        ln_row_finish(va, ln_nv, p.N, lane, lnw4, lnb4, p.ln_eps, p.ln_out + (size_t)(row0 + r) * p.N);
                        if (r + 3 < rows) ln_row_load(X + (size_t)(row0 + r + 3) * p.ldo, ln_nv, lane, va);
                        if (r + 1 < rows) ln_row_finish(vb, ln_nv, p.N, lane, lnw4, lnb4, p.ln_eps, p.ln_out + (size_t)(row0 + r + 1) * p.N);
                        if (r + 4 < rows) ln_row_load(X + (size_t)(row0 + r + 4) * p.ldo, ln_nv, lane, vb);
                        if (r + 2 < rows) ln_row_finish(vc, ln_nv, p.N, lane, lnw4, lnb4, p.ln_eps, p.ln_out + (size_t)(row0 + r + 2) * p.N);
                        if (r + 5 < rows) ln_row_load(X + (size_t)(row0 + r + 5) * p.ldo, ln_nv, lane, vc);
                    }
                }
            }
        }
    }
    else
    {
        // ===================== epilogue warps (2..5) =====================
        const int q = warp_idx & 3; // TMEM lane quarter this warp may access
        const uint32_t sw = lane & 7;
        int it = 0;
        if constexpr (kResid)
        {
            // out = (acc + bias) + resid, f32.  Per warp: a ring of 4-KB slots, each one 32 rows x 32 f32 columns in the
            // TMA SWIZZLE_128B layout (16-B chunk c of row r at chunk position c ^ (r & 7)).  Lane 0 keeps kResidRing-1
            // residual loads in flight (across tile boundaries); every lane then adds its accumulator row segment in
            // place and lane 0 TMA-stores the slot.  In-place on X is safe: a chunk is stored only after its own load
            // completed, and prefetched chunks belong to other tiles / columns.
            constexpr int R = Cfg::kResidRing;
            const int ew = warp_idx - Cfg::kFirstEpiWarp;
            uint8_t *ring = stg_base + ew * (R * 4096);
            const uint32_t ring_u32 = ptx::smem_u32(ring);
            auto n_chunks_of = [&](int tile) { const int n0 = (tile % n_tiles) * BN; const int rem = p.N - n0; return (rem >= BN ? BN : rem + 31) / 32; };
            // prefetch cursor
            int pf_tile = group_id, pf_chunk = 0, pf_count = 0;
            auto issue_next = [&]() {
                if (pf_tile >= num_tiles) return;
                const int slot = pf_count % R;
                const int row0 = (pf_tile / n_tiles) * TILE_M + (int)cta_rank * GEMM_BM + q * 32;
                const int col0 = (pf_tile % n_tiles) * BN + pf_chunk * 32;
                ptx::mbar_arrive_expect_tx(rfull_bar(ew, slot), 4096);
                ptx::tma_load_2d(ring_u32 + slot * 4096, &tmX, rfull_bar(ew, slot), col0, row0);
                ++pf_count;
                if (++pf_chunk == n_chunks_of(pf_tile)) { pf_chunk = 0; pf_tile += num_groups; }
            };
            if (lane == 0)
                for (int i = 0; i < R - 1; ++i) issue_next();
            // shared-memory extras of this epilogue: per-warp bias strip (BN floats: with the 227 KB carve-out there is no L1 to speak
            // of, a __ldg per chunk was an L2 round trip), then the fused LayerNorm's weight and bias
            uint8_t *extra = smem + kStages * (Cfg::A_BYTES + Cfg::B_BYTES) + Cfg::STAGE_BYTES + Cfg::BAR_BYTES;
            float4 *bias_r4 = reinterpret_cast<float4 *>(extra) + ew * (BN / 4);
            const bool fuse_ln = p.ln_out != nullptr;
            volatile int *qbuf = reinterpret_cast<volatile int *>(extra + 4 * BN * 4 + 2 * Cfg::kLnMaxD * 4);
            int *qctl = const_cast<int *>(qbuf) + Cfg::kLnQueue; // [0] reserved, [1] published (tail), [2] claimed (head), [3] consumed
            auto load_bias_r = [&](int tile, int half) { // lane's float4 #half of the tile's bias (columns 4 (lane + 32 half) ..)
                const int col = (tile % n_tiles) * BN + (lane + 32 * half) * 4;
                return (tile < num_tiles && lane + 32 * half < BN / 4 && col < p.N) ? __ldg(reinterpret_cast<const float4 *>(p.bias + col))
                                                                                     : make_float4(0.f, 0.f, 0.f, 0.f);
            };
            float4 bnext0 = load_bias_r(group_id, 0), bnext1 = load_bias_r(group_id, 1);
            // lane 0: publish `k` items (first, first + 8, ...) or one sentinel to the LayerNorm warps: reserve, write, publish in order
            auto q_push = [&](int first, int k) {
                const int base = atomicAdd(&qctl[0], k);
                while (base + k - *reinterpret_cast<volatile int *>(&qctl[3]) > Cfg::kLnQueue) __nanosleep(100); // flow control
                for (int i = 0; i < k; ++i) qbuf[(base + i) % Cfg::kLnQueue] = first < 0 ? -1 : first + Cfg::kLnItemRows * i;
                while (*reinterpret_cast<volatile int *>(&qctl[1]) != base) __nanosleep(20);
                __threadfence_block();
                *reinterpret_cast<volatile int *>(&qctl[1]) = base + k;
            };
            // "this warp's stores of row group `row0` have all completed": bump the group's counter; the warp that brings it to
            // n_tiles (every column tile of those 32 rows is in L2 / HBM) hands the rows to the LayerNorm warps and resets the
            // counter.  `pending` = bulk groups committed AFTER that tile's last store (they may still be in flight).
            auto ln_signal = [&](int row0, int pending) {
                if (lane == 0 && !(p.ln_dbg & 8))
                {
                    if (pending >= 4) ptx::tma_store_wait_group<4>(); else ptx::tma_store_wait_group<0>();
                    if (!(p.ln_dbg & 1)) __threadfence();
                    if (atomicAdd(p.ln_count + (row0 >> 5), 1) == n_tiles - 1)
                    {
                        p.ln_count[row0 >> 5] = 0;
                        if (row0 < p.M) q_push(row0, min(32 / Cfg::kLnItemRows, (p.M - row0 + Cfg::kLnItemRows - 1) / Cfg::kLnItemRows));
                    }
                }
                __syncwarp();
            };
            int cons = 0;
            int prev_row0 = -1; // row group of the previous tile, not yet signalled
            for (int tile = group_id; tile < num_tiles; tile += num_groups, ++it)
            {
                const int m_blk = tile / n_tiles, n_blk = tile % n_tiles;
                const int as = it & 1;
                const uint32_t aphase = (it >> 1) & 1;
                bias_r4[lane] = bnext0;
                if (BN > 128) bias_r4[lane + 32] = bnext1;
                __syncwarp();
                bnext0 = load_bias_r(tile + num_groups, 0);
                bnext1 = load_bias_r(tile + num_groups, 1);
                ptx::mbar_wait(tfull_bar(as), aphase);
                ptx::tcgen05_fence_after();
                const int m0 = m_blk * TILE_M + (int)cta_rank * GEMM_BM + q * 32;
                const int n0 = n_blk * BN;
                const uint32_t tmem_acc = tmem_base + ((uint32_t)(q * 32) << 16) + as * BN;
                const int nch = n_chunks_of(tile);
#pragma unroll 1
                for (int c = 0; c < nch; ++c, ++cons)
                {
                    uint32_t v[32];
                    ptx::tcgen05_ld_32x32b_x32(tmem_acc + c * 32, v);
                    ptx::tcgen05_wait_ld();
                    if (c == nch - 1)
                    {
                        ptx::tcgen05_fence_before();
                        __syncwarp();
                        if (lane == 0) { if (cta_rank == 0) ptx::mbar_arrive(tempty_bar(as)); else ptx::mbar_arrive_remote(tempty_bar(as), 0); }
                    }
                    const int slot = cons % R;
                    ptx::mbar_wait(rfull_bar(ew, slot), (cons / R) & 1);
                    float4 *sl = reinterpret_cast<float4 *>(ring + slot * 4096);
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                    {
                        const float4 b = bias_r4[c * 8 + j]; // zeros for columns >= N
                        float4 r = sl[lane * 8 + (j ^ sw)];
                        r.x = __fadd_rn(__fadd_rn(__uint_as_float(v[4 * j + 0]), b.x), r.x); // (mul_mat + b) + inpL, vit.cpp:869,873
                        r.y = __fadd_rn(__fadd_rn(__uint_as_float(v[4 * j + 1]), b.y), r.y);
                        r.z = __fadd_rn(__fadd_rn(__uint_as_float(v[4 * j + 2]), b.z), r.z);
                        r.w = __fadd_rn(__fadd_rn(__uint_as_float(v[4 * j + 3]), b.w), r.w);
                        sl[lane * 8 + (j ^ sw)] = r;
                    }
                    ptx::fence_proxy_async_smem(); // generic-proxy smem writes -> visible to the TMA store
                    __syncwarp();
                    if (lane == 0)
                    {
                        ptx::tma_store_2d(&tmX, ring_u32 + slot * 4096, n0 + c * 32, m0);
                        ptx::tma_store_commit();
                        // the slot of the PREVIOUS chunk is the next prefetch target: its store must have drained smem
                        ptx::tma_store_wait_read<1>();
                        issue_next();
                    }
                    __syncwarp();
                    // the previous tile's stores have had four chunk times to complete: signal its row group now (no stall)
                    if (fuse_ln && prev_row0 >= 0 && (c == 3 || (c == nch - 1 && nch < 4)))
                    {
                        ln_signal(prev_row0, c == 3 ? 4 : 0);
                        prev_row0 = -1;
                    }
                }
                if (fuse_ln) prev_row0 = m0;
            }
            if (fuse_ln && prev_row0 >= 0) ln_signal(prev_row0, 0);
            if (fuse_ln && lane == 0) q_push(-1, Cfg::kLnWarps / 4); // sentinels: as many in total as there are LayerNorm warps
            if (lane == 0) ptx::tma_store_wait_read<0>();
            __syncwarp();
        }
        else
        {
        uint8_t *stg = stg_base + (warp_idx - Cfg::kFirstEpiWarp) * 4096;
        const int half_sel = (warp_idx - Cfg::kFirstEpiWarp) >> 2; // warps w and w+4 share a TMEM lane quarter and alternate column passes
        // f16 epilogues: this warp's bias strip.  Lane l holds (and pre-loads, one tile ahead) the 4 bias values of columns
        // 4 (l & 15) .. +3 of the warp's column pass half_sel + 2 (l >> 4); the strip is [pass][16 x float4].
        float4 *bias_s4 = reinterpret_cast<float4 *>(smem + kStages * (Cfg::A_BYTES + Cfg::B_BYTES) + Cfg::STAGE_BYTES + Cfg::BAR_BYTES) +
                          (warp_idx - Cfg::kFirstEpiWarp) * 32;
        auto load_bias = [&](int tile) {
            float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
            if constexpr (kOutF16)
            {
                const int cp = half_sel + 2 * (lane >> 4);
                const int col = (tile % n_tiles) * BN + cp * 64 + (lane & 15) * 4;
                if (tile < num_tiles && cp < BN / 64 && col < p.N) b = __ldg(reinterpret_cast<const float4 *>(p.bias + col)); // N % 8 == 0
            }
            return b;
        };
        float4 bias_next = load_bias(group_id);
        for (int tile = group_id; tile < num_tiles; tile += num_groups, ++it)
        {
            const int m_blk = tile / n_tiles, n_blk = tile % n_tiles;
            const int as = it & 1;
            const uint32_t aphase = (it >> 1) & 1;
            if constexpr (kOutF16)
            {
                bias_s4[lane] = bias_next;             // every lane is past the previous tile's reads (__syncwarp at the end of its last pass)
                __syncwarp();
                bias_next = load_bias(tile + num_groups); // in flight while this tile's accumulator is awaited and processed
            }
            ptx::mbar_wait(tfull_bar(as), aphase);
            ptx::tcgen05_fence_after();
            const int m0 = m_blk * TILE_M + (int)cta_rank * GEMM_BM + q * 32;
            const int n0 = n_blk * BN;
            const uint32_t tmem_acc = tmem_base + ((uint32_t)(q * 32) << 16) + as * BN;

            constexpr int CH = kOutF16 ? 64 : 32; // columns per pass: one 128-B row segment per thread-row
            constexpr int kPasses = BN / CH;
            static_assert(kPasses % 2 == 0, "column passes must split evenly between the two warps of a quarter");
#pragma unroll 1
            for (int cp = half_sel; cp < kPasses; cp += 2)
            {
                const int c = cp * CH;
                uint32_t v[CH];
                {
                    uint32_t(&v0)[32] = *reinterpret_cast<uint32_t(*)[32]>(&v[0]);
                    ptx::tcgen05_ld_32x32b_x32(tmem_acc + c, v0);
                    if constexpr (CH == 64)
                    {
                        uint32_t(&v1)[32] = *reinterpret_cast<uint32_t(*)[32]>(&v[32]);
                        ptx::tcgen05_ld_32x32b_x32(tmem_acc + c + 32, v1);
                    }
                }
                ptx::tcgen05_wait_ld();
                if (cp + 2 >= kPasses)
                {
                    // this warp's last TMEM read of this accumulator: hand it back to the MMA warp early
                    ptx::tcgen05_fence_before();
                    __syncwarp();
                    if (lane == 0) { if (cta_rank == 0) ptx::mbar_arrive(tempty_bar(as)); else ptx::mbar_arrive_remote(tempty_bar(as), 0); }
                }

                if constexpr (kOutF16)
                {
                    // math in the row-per-thread domain (bias is warp-uniform -> broadcast loads from the strip), pack to f16 into this
                    // warp's 32-row x 128-B staging box -- 16-B chunk j of row r at j ^ (r & 7), which is exactly TMA SWIZZLE_128B and
                    // bank-conflict free -- then ONE TMA store per box: rows >= M and columns >= N are clipped by the tensor map.  (The
                    // first version read the box back with LDS and issued 8 STG.128 per lane group: a third of the epilogue's
                    // instructions and half of its LSU shared-memory wavefronts; ncu had the split-precision qkv epilogue at 72 %
                    // tensor-pipe activity against 90 % for the plain one.)  The hi-lo epilogue runs the round trip twice over the same
                    // accumulator registers: part 0 emits hi = f16(x), part 1 recomputes x and hi and emits lo = f16(x - hi) (exact
                    // difference, |lo| <= 2^-11 |x|) to the second tensor.
                    const uint32_t stg_u32 = ptx::smem_u32(stg);
#pragma unroll
                    for (int part = 0; part < (kHiLo ? 2 : 1); ++part)
                    {
                    if constexpr (EPI == EPI_BIAS_GELU_F16)
                    {
                        // GELU: all the arithmetic of the box first, into registers; only then wait for the previous TMA store to have read
                        // the staging box and copy the packed rows in -- the store's shared-memory read overlaps the (long) GELU chain
                        // instead of preceding it.  A/B at batch 256: fc1 2.60 -> 2.46 ms per forward; the same reordering makes the
                        // hi-lo epilogue SLOWER (qkv 2.16 -> 2.33 ms: 160 instead of 124 registers, both parts' packed rows live), so
                        // that one keeps the wait-first order below (profiles/microbench_r02.md).
                        uint32_t packed[32];
#pragma unroll
                        for (int j = 0; j < 8; ++j) // 8 chunks of 8 columns (16 B of f16)
                        {
                            const float4 bl = bias_s4[(cp >> 1) * 16 + j * 2], bh = bias_s4[(cp >> 1) * 16 + j * 2 + 1];
                            const float bias8[8] = {bl.x, bl.y, bl.z, bl.w, bh.x, bh.y, bh.z, bh.w};
#pragma unroll
                            for (int e = 0; e < 4; ++e)
                            {
                                float x0, x1; // acc + bias, both lanes in one FADD2
                                ptx::unpack_f32x2(ptx::add_f32x2(ptx::pack_f32x2(__uint_as_float(v[j * 8 + e * 2]), __uint_as_float(v[j * 8 + e * 2 + 1])),
                                                                 ptx::pack_f32x2(bias8[e * 2], bias8[e * 2 + 1])), x0, x1);
                                const float2 r = __half22float2(__floats2half2_rn(x0, x1)); // ggml.c:1434-1441: y = f16(gelu(f32(f16(x))))
                                float g0, g1;
                                gelu_tanh_f32x2(r.x, r.y, g0, g1);
                                const __half2 h = __floats2half2_rn(g0, g1);
                                packed[j * 4 + e] = *reinterpret_cast<const uint32_t *>(&h);
                            }
                        }
                        if (lane == 0) ptx::tma_store_wait_read<0>(); // the previous box has left the staging buffer
                        __syncwarp();
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            ptx::st_shared_v4(stg_u32 + (uint32_t)lane * 128u + (uint32_t)((j ^ sw) << 4), packed[4 * j], packed[4 * j + 1], packed[4 * j + 2], packed[4 * j + 3]);
                    }
                    else
                    {
                        if (lane == 0) ptx::tma_store_wait_read<0>(); // the previous box has left the staging buffer
                        __syncwarp();
#pragma unroll
                        for (int j = 0; j < 8; ++j) // 8 chunks of 8 columns (16 B of f16)
                        {
                            uint32_t packed[4];
                            // the strip holds zeros for columns >= N (N % 8 == 0 for f16 outputs: a chunk of 8 columns is all-in or all-out)
                            const float4 bl = bias_s4[(cp >> 1) * 16 + j * 2], bh = bias_s4[(cp >> 1) * 16 + j * 2 + 1];
                            const float bias8[8] = {bl.x, bl.y, bl.z, bl.w, bh.x, bh.y, bh.z, bh.w};
#pragma unroll
                            for (int e = 0; e < 4; ++e)
                            {
                                float x0, x1; // acc + bias, both lanes in one FADD2
                                ptx::unpack_f32x2(ptx::add_f32x2(ptx::pack_f32x2(__uint_as_float(v[j * 8 + e * 2]), __uint_as_float(v[j * 8 + e * 2 + 1])),
                                                                 ptx::pack_f32x2(bias8[e * 2], bias8[e * 2 + 1])), x0, x1);
                                __half2 h;
                                if constexpr (EPI == EPI_BIAS_GELU_F16)
                                {
                                    // ggml.c:1434-1441: y = f16(gelu(f32(f16(x))))
                                    const float2 r = __half22float2(__floats2half2_rn(x0, x1));
                                    float g0, g1;
                                    gelu_tanh_f32x2(r.x, r.y, g0, g1);
                                    h = __floats2half2_rn(g0, g1);
                                }
                                else
                                {
                                    h = __floats2half2_rn(x0, x1);
                                    if (kHiLo && part == 1)
                                    {
                                        const float2 hf = __half22float2(h);
                                        h = __floats2half2_rn(__fsub_rn(x0, hf.x), __fsub_rn(x1, hf.y));
                                    }
                                }
                                packed[e] = *reinterpret_cast<uint32_t *>(&h);
                            }
                            ptx::st_shared_v4(stg_u32 + (uint32_t)lane * 128u + (uint32_t)((j ^ sw) << 4), packed[0], packed[1], packed[2], packed[3]);
                        }
                    }
                    ptx::fence_proxy_async_smem(); // generic-proxy writes -> visible to the TMA store (async proxy)
                    __syncwarp();
                    if (lane == 0)
                    {
                        if (p.store_policy)
                        {
                            if (p.headmajor) ptx::tma_store_3d_hint(part == 0 ? &tmX : &tmO2, stg_u32, 0, m0, (n0 + c) >> 6, p.store_policy);
                            else ptx::tma_store_2d_hint(part == 0 ? &tmX : &tmO2, stg_u32, n0 + c, m0, p.store_policy);
                        }
                        else if (p.headmajor) ptx::tma_store_3d(part == 0 ? &tmX : &tmO2, stg_u32, 0, m0, (n0 + c) >> 6);
                        else ptx::tma_store_2d(part == 0 ? &tmX : &tmO2, stg_u32, n0 + c, m0);
                        ptx::tma_store_commit();
                    }
                    }
                }
                else
                {
                    // raw accumulators through the transpose buffer; bias / residual / pos math on the coalesced side.  The bias and
                    // pos_embed operands do not depend on the accumulator: they are requested BEFORE the staging round trip (there is no
                    // L1 to speak of next to 227 KB of shared memory, so each of these loads is an L2 round trip that used to sit,
                    // serialised, between the LDS and the store of every 16 bytes).
                    float4 *stg4 = reinterpret_cast<float4 *>(stg);
                    const int ch = lane & 7;
                    const int gcol = n0 + c + ch * 4;
                    const bool col_ok = gcol < p.N;
                    const float4 b = col_ok ? __ldg(reinterpret_cast<const float4 *>(p.bias + gcol)) : make_float4(0.f, 0.f, 0.f, 0.f);
                    float4 pe[kPatch ? 8 : 1];
                    size_t orow_[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                    {
                        const int grow = m0 + i * 4 + (lane >> 3);
                        orow_[i] = (size_t)grow;
                        if constexpr (kPatch)
                        {
                            const int img = grow / p.np, pp = grow - img * p.np;
                            orow_[i] = (size_t)img * p.ntok + 1 + pp;
                            pe[i] = (grow < p.M && col_ok) ? __ldg(reinterpret_cast<const float4 *>(p.pos + (size_t)(1 + pp) * p.N + gcol))
                                                           : make_float4(0.f, 0.f, 0.f, 0.f);
                        }
                    }
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        stg4[lane * 8 + (j ^ sw)] = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                                                                __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
                    __syncwarp();
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                    {
                        const int row = i * 4 + (lane >> 3);
                        float4 a = stg4[row * 8 + (ch ^ (row & 7))];
                        const int grow = m0 + row;
                        if (grow < p.M && col_ok)
                        {
                            a.x = __fadd_rn(a.x, b.x); a.y = __fadd_rn(a.y, b.y); a.z = __fadd_rn(a.z, b.z); a.w = __fadd_rn(a.w, b.w);
                            if constexpr (EPI == EPI_BIAS_RESID_F32)
                            {
                                const float4 r = *reinterpret_cast<const float4 *>(p.resid + (size_t)grow * p.ldo + gcol);
                                a.x = __fadd_rn(a.x, r.x); a.y = __fadd_rn(a.y, r.y); a.z = __fadd_rn(a.z, r.z); a.w = __fadd_rn(a.w, r.w);
                            }
                            if constexpr (kPatch)
                            {
                                const float4 r = pe[i];
                                a.x = __fadd_rn(a.x, r.x); a.y = __fadd_rn(a.y, r.y); a.z = __fadd_rn(a.z, r.z); a.w = __fadd_rn(a.w, r.w);
                            }
                            *reinterpret_cast<float4 *>(reinterpret_cast<float *>(p.out) + orow_[i] * p.ldo + gcol) = a;
                        }
                    }
                    __syncwarp();
                }
            }
        }
        if constexpr (kOutF16)
        {
            if (lane == 0) ptx::tma_store_wait_all(); // shared memory must outlive the last box's read
            __syncwarp();
        }
        } // !kResid
    }

    // ===================== teardown =====================
    ptx::tcgen05_fence_before();
    if constexpr (CG == 2) ptx::cluster_sync(); else __syncthreads(); // the peer may still be reading its TMEM / signalling our barriers
    if (warp_idx == Cfg::kFirstEpiWarp)
    {
        ptx::tcgen05_fence_after();
        if constexpr (CG == 2) ptx::tcgen05_dealloc_cg2(tmem_base, Cfg::TMEM_COLS);
        else ptx::tcgen05_dealloc(tmem_base, Cfg::TMEM_COLS);
    }
}

} // namespace vitb200
