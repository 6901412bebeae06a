// attention_tcgen05.cuh -- fused attention for one (image, head) per CTA iteration on the 5th-gen tensor cores.
//
// Replaces the reference's 16 reshape/permute/CONT nodes + mul_mat(K,Q) + scale + soft_max_inplace + mul_mat(V,P) + head
// merge (reference vit.cpp:826-866; CPU kernels ggml.c:1163-1198 (f32 dot), 10498-10567 (soft-max)).
//
// Precision: the reference feeds F32 q, k, v to both attention mat-muls (vit.cpp:848,858).  The qkv GEMM therefore leaves every
// value as TWO f16 numbers, hi = f16(x) and lo = f16(x - hi) (hi + lo carries ~22 significant bits), and the tensor cores run
//   S = Ql Kh^T + Qh Kl^T + Qh Kh^T   (the Ql Kl^T term is below 2^-22 relative)        O = P Vl + P Vh   (P is exactly f16)
// with f32 accumulation in TMEM: rounding q, k or v to f16 alone costs ~17 % of the logit-parity budget (CPU experiment,
// DESIGN.md section 4).  p.hilo = 0 (lo tensors absent) runs the single-term products.
//
//   S = Q K^T      tcgen05.mma SS: A = Q tile (128 x 64, K-major, TMA SWIZZLE_128B), B = K (NKP x 64, K-major), D in TMEM
//   P = softmax    thread = query row = TMEM lane: TRUE row max (2 TMEM passes), e = f16(exp(f16(s/8 - max))) exactly the
//                  reference's table semantics (ggml.c:10547-10549), un-normalised P written back into the S columns as packed
//                  f16 (tcgen05.st), l = sum e in f32
//   O = P V        tcgen05.mma TS: A = P from TMEM, B = V (NKP x 64, MN-major, same TMA tile), D in TMEM; O * (1/l) -> f16,
//                  staged in shared memory (SWIZZLE_128B) and written with one TMA store per warp
//
// Q, K, V are read by TMA out of the HEAD-MAJOR QKV buffers the qkv GEMM's epilogue writes ([3 H planes][tokens][64]: plane h / H + h /
// 2 H + h; every operand tile is one contiguous run of rows): no split / transpose copies.  Persistent CTAs (1 per SM), 10 warps: warp 0 TMA producer, warp 1 MMA issuer, warps 2-5 / 6-9
// soft-max + epilogue warpgroups for query tile 0 / 1 (rows 0-127 / 128-255).  Shared memory holds ONE problem's operands
// (hi + lo: up to 176 KB) with a full/empty mbarrier pair per operand group -- Q tile 0, Q tile 1, K, V -- so each group of the
// next problem is re-loaded as soon as the last MMA reading it has retired (Q/K right after the scores, V after P V), i.e.
// a whole problem period before it is needed.
// Supports N <= 224 tokens (keys padded to NKP = ceil16(N) <= 224: S0 at TMEM cols [0,224), S1 at [224,448), O at [448,512)).
#pragma once
#include "kernels.cuh"

namespace vitb200 {

struct AttnTcParams
{
    int N, D, H;  // tokens per image, hidden, heads
    int n_problems; // B * H
    int NKP;      // keys padded to a multiple of 16
    int n_mtiles; // 1 or 2 query tiles of 128 rows
    int kv_bytes; // NKP * 128 rounded up to 1024
    float scale;  // 1/sqrt(64)
    unsigned long long load_policy; // L2 eviction hint of the operand loads (0 = none): q, k, v are read exactly once
    int hilo;     // 1: split-precision operands (lo tensors present); 0: hi only
    int reverse;  // 1: problems are taken last image first -- the qkv GEMM wrote its output in ascending row order, so the tail of the
                  // QKV buffers is what the 126 MB L2 still holds when this kernel starts (like layernorm_f16_kernel)
    long long *trace; // dev only (VITB200_ATTN_TRACE): clock64 stamps of CTA 0, [problem < 16][slot < 32]; NULL in production
};

#define ATT_TRACE(slot) do { if (p.trace && blockIdx.x == 0 && i < 16 && lane == 0) p.trace[i * 32 + (slot)] = clock64(); } while (0)

constexpr int ATT_TC_THREADS = 320;
constexpr int ATT_TC_SCOL1 = 224, ATT_TC_OCOL = 448;

__device__ __forceinline__ uint32_t att_exp_pair(float x0, float x1, float &lsum)
{
    // e = f16(exp(f32(f16(x))))  (ggml.c:10547-10549), two elements at a time
    const float2 xr = __half22float2(__floats2half2_rn(x0, x1));
    const __half2 e = __floats2half2_rn(ptx::ex2_approx(xr.x * 1.4426950408889634f), ptx::ex2_approx(xr.y * 1.4426950408889634f));
    ptx::add_f32_f16(lsum, __half_as_ushort(__low2half(e)));
    ptx::add_f32_f16(lsum, __half_as_ushort(__high2half(e)));
    return *reinterpret_cast<const uint32_t *>(&e);
}
// The same from two raw scores: x = s * scale - max as ONE packed FFMA2 and the log2(e) multiply as one FMUL2 (Blackwell packed
// FP32: half the issue slots of the scalar forms; each lane is a separately rounded IEEE operation, results are bit-identical).
__device__ __forceinline__ uint32_t att_exp_pair_raw(uint32_t s0, uint32_t s1, uint64_t scale2, uint64_t nmax2, float &lsum)
{
    float x0, x1;
    ptx::unpack_f32x2(ptx::fma_f32x2(ptx::pack_f32x2(__uint_as_float(s0), __uint_as_float(s1)), scale2, nmax2), x0, x1);
    const float2 xr = __half22float2(__floats2half2_rn(x0, x1));
    float y0, y1;
    ptx::unpack_f32x2(ptx::mul_f32x2(ptx::pack_f32x2(xr.x, xr.y), ptx::pack_f32x2(1.4426950408889634f, 1.4426950408889634f)), y0, y1);
    const __half2 e = __floats2half2_rn(ptx::ex2_approx(y0), ptx::ex2_approx(y1));
    ptx::add_f32_f16(lsum, __half_as_ushort(__low2half(e)));
    ptx::add_f32_f16(lsum, __half_as_ushort(__high2half(e)));
    return *reinterpret_cast<const uint32_t *>(&e);
}

// dynamic shared memory of attention_tc_kernel: alignment slack + 4 Q tiles (hi/lo x 2 query tiles) + K, V (hi/lo) + 8 store boxes + barriers
__host__ __device__ inline int attention_tc_smem_bytes(int kv_bytes) { return 1024 + 4 * 16384 + 4 * kv_bytes + 8 * 4096 + 256; }

__global__ void __launch_bounds__(ATT_TC_THREADS, 1)
attention_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
                    const __grid_constant__ CUtensorMap tmQl, const __grid_constant__ CUtensorMap tmKVl,
                    const __grid_constant__ CUtensorMap tmO, const AttnTcParams p)
{
    extern __shared__ uint8_t att_tc_smem_raw[];
    const uint32_t smem_base = (ptx::smem_u32(att_tc_smem_raw) + 1023u) & ~1023u;
    uint8_t *smem = att_tc_smem_raw + (smem_base - ptx::smem_u32(att_tc_smem_raw));
    // operand layout (one problem): Qh tile 0 | Qh tile 1 | Ql tile 0 | Ql tile 1 | Kh | Kl | Vh | Vl   (hl = 0 hi, 1 lo)
    auto sQ = [&](int hl, int t) { return smem_base + (uint32_t)(hl * 2 + t) * 16384u; };
    auto sK = [&](int hl) { return smem_base + 4u * 16384u + (uint32_t)hl * (uint32_t)p.kv_bytes; };
    auto sV = [&](int hl) { return smem_base + 4u * 16384u + (uint32_t)(2 + hl) * (uint32_t)p.kv_bytes; };
    const uint32_t stage_out = smem_base + 4u * 16384u + 4u * (uint32_t)p.kv_bytes; // 8 x 4 KB: one 32-row x 128-B SWIZZLE_128B store box per soft-max warp
    const uint32_t bars = stage_out + 8 * 4096;
    // barriers: q_full[2], q_empty[2], k_full, k_empty, v_full, v_empty, s_full[2], p_ready[2], o_full[2], o_empty[2], tmem ptr
    auto q_full = [&](int t) { return bars + 8u * t; };
    auto q_empty = [&](int t) { return bars + 8u * (2 + t); };
    const uint32_t k_full = bars + 8u * 4, k_empty = bars + 8u * 5, v_full = bars + 8u * 6, v_empty = bars + 8u * 7;
    auto s_full = [&](int t) { return bars + 8u * (8 + t); };
    auto p_ready = [&](int t) { return bars + 8u * (10 + t); };
    auto o_full = [&](int t) { return bars + 8u * (12 + t); };
    auto o_empty = [&](int t) { return bars + 8u * (14 + t); };
    const uint32_t tmem_ptr_addr = bars + 8u * 16;
    volatile uint32_t *tmem_ptr_gen = reinterpret_cast<volatile uint32_t *>(smem + 4 * 16384 + 4 * p.kv_bytes + 8 * 4096 + 8 * 16);

    const int warp_idx = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool hilo = p.hilo != 0;
    if (warp_idx == 0 && lane == 0)
    {
        ptx::prefetch_tensormap(&tmQ);
        ptx::prefetch_tensormap(&tmKV);
        ptx::prefetch_tensormap(&tmO);
        if (hilo) { ptx::prefetch_tensormap(&tmQl); ptx::prefetch_tensormap(&tmKVl); }
    }
    if (warp_idx == 1 && lane == 0)
    {
        for (int i = 0; i < 2; ++i)
        {
            ptx::mbar_init(q_full(i), 1);
            ptx::mbar_init(q_empty(i), 1);
            ptx::mbar_init(s_full(i), 1);
            ptx::mbar_init(p_ready(i), 4);
            ptx::mbar_init(o_full(i), 1);
            ptx::mbar_init(o_empty(i), 4);
        }
        ptx::mbar_init(k_full, 1);
        ptx::mbar_init(k_empty, 1);
        ptx::mbar_init(v_full, 1);
        ptx::mbar_init(v_empty, 1);
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
    ptx::grid_dep_launch(); // PDL: the prologue above overlapped the previous kernel's tail
    ptx::grid_dep_wait();
    const uint32_t scol[2] = {0u, (uint32_t)ATT_TC_SCOL1};

    if (warp_idx == 0)
    {
        // ===================== TMA producer =====================
        // Per problem: K, Q tile 0, Q tile 1, then V -- each group as soon as its previous contents have been consumed (the
        // *_empty barriers are tcgen05.commit arrivals behind the last MMA that read the group).
        if (lane == 0)
        {
            const uint32_t kv_tx = (uint32_t)(p.NKP * 128) * (hilo ? 2u : 1u), q_tx = 16384u * (hilo ? 2u : 1u);
            // L2 prefetch of every operand tile of one problem: shared memory holds a single problem, so the real loads can only be
            // issued late (when the previous problem's MMAs retire); prefetching a problem ahead moves the HBM latency and the
            // bandwidth burst off that critical window -- the loads then come out of L2.
            auto prefetch = [&](int prob) {
                const int rp = p.reverse ? p.n_problems - 1 - prob : prob;
                const int b = rp / p.H, h = rp - b * p.H, row0 = b * p.N;
                for (int hl = 0; hl < (hilo ? 2 : 1); ++hl)
                {
                    const CUtensorMap *mkv = hl ? &tmKVl : &tmKV, *mq = hl ? &tmQl : &tmQ;
                    ptx::tma_prefetch_3d(mkv, 0, row0, p.H + h);
                    for (int t = 0; t < p.n_mtiles; ++t) ptx::tma_prefetch_3d(mq, 0, row0 + t * 128, h);
                    ptx::tma_prefetch_3d(mkv, 0, row0, 2 * p.H + h);
                }
            };
            auto ld3 = [&](uint32_t dst, const CUtensorMap *m, uint32_t bar, int c0, int c1, int c2) {
                if (p.load_policy) ptx::tma_load_3d_hint(dst, m, bar, c0, c1, c2, p.load_policy);
                else ptx::tma_load_3d(dst, m, bar, c0, c1, c2);
            };
            int i = 0;
            for (int prob = blockIdx.x; prob < p.n_problems; prob += gridDim.x, ++i)
            {
                const uint32_t ph = (uint32_t)(i & 1);
                const int rp = p.reverse ? p.n_problems - 1 - prob : prob; // (image, head) pairs are walked LAST FIRST, see AttnTcParams::reverse
                const int b = rp / p.H, h = rp - b * p.H;
                const int row0 = b * p.N;
                ptx::mbar_wait(k_empty, ph ^ 1);
                ptx::mbar_arrive_expect_tx(k_full, kv_tx);
                ld3(sK(0), &tmKV, k_full, 0, row0, p.H + h);
                if (hilo) ld3(sK(1), &tmKVl, k_full, 0, row0, p.H + h);
                for (int t = 0; t < p.n_mtiles; ++t)
                {
                    ptx::mbar_wait(q_empty(t), ph ^ 1);
                    ptx::mbar_arrive_expect_tx(q_full(t), q_tx);
                    ld3(sQ(0, t), &tmQ, q_full(t), 0, row0 + t * 128, h);
                    if (hilo) ld3(sQ(1, t), &tmQl, q_full(t), 0, row0 + t * 128, h);
                }
                // K/Q of this problem are on their way; now ask L2 for everything the NEXT problem will need
                if (prob + (int)gridDim.x < p.n_problems) prefetch(prob + (int)gridDim.x);
                ptx::mbar_wait(v_empty, ph ^ 1);
                ptx::mbar_arrive_expect_tx(v_full, kv_tx);
                ld3(sV(0), &tmKV, v_full, 0, row0, 2 * p.H + h);
                if (hilo) ld3(sV(1), &tmKVl, v_full, 0, row0, 2 * p.H + h);
            }
        }
        __syncwarp();
    }
    else if (warp_idx == 1)
    {
        // ===================== MMA issuer: the warp runs the loop uniformly, one elected lane issues =====================
        {
            const uint32_t idesc_s = ptx::umma_idesc_f16(128, p.NKP, 0, 0, 0, 0);
            const uint32_t idesc_o = ptx::umma_idesc_f16(128, 64, 0, 0, 0, /*B (V) is MN-major*/ 1);
            const int ksteps = p.NKP / 16;
            const int last_t = p.n_mtiles - 1;
            // S_t = Q_t K^T of problem `j` (its operands must have landed): small cross terms first, then the hi x hi product
            auto issue_s = [&](int t, int j) {
                if (t == 0) ptx::mbar_wait(k_full, (uint32_t)(j & 1));
                ptx::mbar_wait(q_full(t), (uint32_t)(j & 1));
                ptx::tcgen05_fence_after();
                const uint64_t kh = ptx::umma_desc_kmajor_sw128(sK(0)), kl = ptx::umma_desc_kmajor_sw128(sK(1));
                const uint64_t qh = ptx::umma_desc_kmajor_sw128(sQ(0, t)), ql = ptx::umma_desc_kmajor_sw128(sQ(1, t));
                if (ptx::elect_one())
                {
                    if (hilo)
                    {
#pragma unroll
                        for (int k = 0; k < 4; ++k) ptx::tcgen05_mma_f16(tmem_base + scol[t], ql + 2 * k, kh + 2 * k, idesc_s, k > 0);
#pragma unroll
                        for (int k = 0; k < 4; ++k) ptx::tcgen05_mma_f16(tmem_base + scol[t], qh + 2 * k, kl + 2 * k, idesc_s, 1);
                    }
#pragma unroll
                    for (int k = 0; k < 4; ++k) ptx::tcgen05_mma_f16(tmem_base + scol[t], qh + 2 * k, kh + 2 * k, idesc_s, (hilo || k > 0) ? 1 : 0);
                    ptx::tcgen05_commit(s_full(t));
                    ptx::tcgen05_commit(q_empty(t));              // Q tile t may be overwritten once these MMAs retire
                    if (t == last_t) ptx::tcgen05_commit(k_empty); // ... and K after the last tile's scores
                }
                __syncwarp();
            };
            int i = 0;
            const int first = blockIdx.x;
            if (first < p.n_problems)
                for (int t = 0; t < p.n_mtiles; ++t) issue_s(t, 0);
            for (int prob = first; prob < p.n_problems; prob += gridDim.x, ++i)
            {
                const bool has_next = prob + (int)gridDim.x < p.n_problems;
                const uint64_t vh = ptx::umma_desc_mnmajor_sw128(sV(0), (uint32_t)p.kv_bytes);
                const uint64_t vl = ptx::umma_desc_mnmajor_sw128(sV(1), (uint32_t)p.kv_bytes);
                for (int t = 0; t < p.n_mtiles; ++t)
                {
                    ptx::mbar_wait(p_ready(t), i & 1); // P_t is in TMEM (and the soft-max warps are done with S_t)
                    // the single O accumulator must have been drained by its previous user
                    if (t == 0) { if (i > 0) ptx::mbar_wait(o_empty(last_t), (i - 1) & 1); }
                    else ptx::mbar_wait(o_empty(0), i & 1);
                    if (t == 0) ptx::mbar_wait(v_full, (uint32_t)(i & 1));
                    ptx::tcgen05_fence_after();
                    if (ptx::elect_one())
                    {
                        // 16 keys per step: 8 TMEM columns of packed f16, 16 rows (2048 B) of V; O = P Vl + P Vh
                        if (hilo)
                            for (int j = 0; j < ksteps; ++j)
                                ptx::tcgen05_mma_f16_ts(tmem_base + ATT_TC_OCOL, tmem_base + scol[t] + 8 * j, vl + (uint64_t)(j * 128), idesc_o, j > 0);
                        for (int j = 0; j < ksteps; ++j)
                            ptx::tcgen05_mma_f16_ts(tmem_base + ATT_TC_OCOL, tmem_base + scol[t] + 8 * j, vh + (uint64_t)(j * 128), idesc_o, (hilo || j > 0) ? 1 : 0);
                        ptx::tcgen05_commit(o_full(t));
                        if (t == last_t) ptx::tcgen05_commit(v_empty); // all MMAs reading V have retired
                    }
                    __syncwarp();
                    // S_t of the NEXT problem goes into the pipe right behind P_t V: the in-order tensor pipe runs it after P_t V has
                    // consumed the aliased P_t columns, so tile t's warpgroup finds its next scores ready as soon as it has drained O
                    if (has_next) issue_s(t, i + 1);
                }
            }
        }
        __syncwarp();
    }
    else
    {
        // ===================== soft-max + epilogue warpgroups =====================
        const int t = (warp_idx - 2) >> 2; // query tile
        const int q = warp_idx & 3;        // TMEM lane quarter
        if (t < p.n_mtiles)
        {
            const bool warp_valid = (t * 128 + q * 32) < p.N;  // warp owns at least one real query row
            const uint32_t t_s = tmem_base + ((uint32_t)(q * 32) << 16) + scol[t];
            const uint32_t t_o = tmem_base + ((uint32_t)(q * 32) << 16) + ATT_TC_OCOL;
            const int n32 = p.NKP >> 5;
            const bool tail16 = (p.NKP & 16) != 0;
            const uint64_t scale2 = ptx::pack_f32x2(p.scale, p.scale);
            int i = 0;
            for (int prob = blockIdx.x; prob < p.n_problems; prob += gridDim.x, ++i)
            {
                const int rp = p.reverse ? p.n_problems - 1 - prob : prob; // (image, head) pairs are walked LAST FIRST, see AttnTcParams::reverse
                const int b = rp / p.H, h = rp - b * p.H;
                if (q == 0) ATT_TRACE(8 * t + 0);
                ptx::mbar_wait(s_full(t), i & 1);
                ptx::tcgen05_fence_after();
                if (q == 0) ATT_TRACE(8 * t + 1);
                float lsum = 0.f;
                if (warp_valid)
                {
                    // ---- pass 1: true row maximum over the valid keys (ggml.c:10533-10534)
                    float mx = -INFINITY;
                    {
                        float ma = -INFINITY, mb = -INFINITY;
                        int c = 0;
#pragma unroll 1
                        for (; c + 2 <= (p.N >> 5); c += 2) // two mask-free chunks per iteration
                        {
                            uint32_t va[32], vb[32];
                            ptx::tcgen05_ld_32x32b_x32(t_s + c * 32, va);
                            ptx::tcgen05_ld_32x32b_x32(t_s + c * 32 + 32, vb);
                            ptx::tcgen05_wait_ld();
#pragma unroll
                            for (int j = 0; j < 32; ++j) { ma = fmaxf(ma, __uint_as_float(va[j])); mb = fmaxf(mb, __uint_as_float(vb[j])); }
                        }
                        mx = fmaxf(ma, mb);
#pragma unroll 1
                        for (; c < n32; ++c)
                        {
                            uint32_t v[32];
                            ptx::tcgen05_ld_32x32b_x32(t_s + c * 32, v);
                            ptx::tcgen05_wait_ld();
#pragma unroll
                            for (int j = 0; j < 32; ++j)
                                if (c * 32 + j < p.N) mx = fmaxf(mx, __uint_as_float(v[j]));
                        }
                    }
                    if (tail16)
                    {
                        uint32_t v[16];
                        ptx::tcgen05_ld_32x32b_x16(t_s + n32 * 32, v);
                        ptx::tcgen05_wait_ld();
#pragma unroll
                        for (int j = 0; j < 16; ++j)
                            if (n32 * 32 + j < p.N) mx = fmaxf(mx, __uint_as_float(v[j]));
                    }
                    if (q == 0) ATT_TRACE(8 * t + 2);
                    const float mxs = mx * p.scale; // ggml_scale_inplace (vit.cpp:851-854); exact, scale = 1/8
                    const uint64_t nmax2 = ptx::pack_f32x2(-mxs, -mxs);
                    // ---- pass 2: P = f16(exp(f16(s*scale - max))) -> packed f16 into the S columns, l = sum P.
                    // Full (mask-free) chunks go two at a time with four independent partial sums so one warp keeps the
                    // MUFU / conversion latencies overlapped; the chunk(s) straddling N take the masked path.
                    const int n_full = p.N >> 5; // chunks with all 32 keys valid
                    float l0 = 0.f, l1 = 0.f, l2s = 0.f, l3 = 0.f;
                    int c = 0;
                    // (software-pipelined variants -- the next chunk's tcgen05.ld in flight during the exponentials -- were measured twice and
                    // dropped: with 64-column buffers at 168 registers the loop spilled (1.7x slower); with 32-column buffers and no spills, in
                    // this 10-warp layout or in a 12-warp layout whose soft-max warpgroups take 224 registers through setmaxnreg, the pass
                    // got 8 % SLOWER (1.76 -> 1.87-1.91 ms per forward): the other soft-max warp of the sub-partition already covers the
                    // TMEM round trip, and the 64-column steps give the scheduler more independent work; profiles/microbench_r02.md)
#pragma unroll 1 // (ptxas otherwise unrolls x4 with three peeled copies: 213 KB of SASS, the live part no longer fits the instruction cache)
                    for (; c + 2 <= n_full; c += 2)
                    {
                        uint32_t va[32], vb[32], pa[16], pb[16];
                        ptx::tcgen05_ld_32x32b_x32(t_s + c * 32, va);
                        ptx::tcgen05_ld_32x32b_x32(t_s + c * 32 + 32, vb);
                        ptx::tcgen05_wait_ld();
#pragma unroll
                        for (int j = 0; j < 16; j += 2)
                        {
                            pa[j] = att_exp_pair_raw(va[2 * j], va[2 * j + 1], scale2, nmax2, l0);
                            pb[j] = att_exp_pair_raw(vb[2 * j], vb[2 * j + 1], scale2, nmax2, l1);
                            pa[j + 1] = att_exp_pair_raw(va[2 * j + 2], va[2 * j + 3], scale2, nmax2, l2s);
                            pb[j + 1] = att_exp_pair_raw(vb[2 * j + 2], vb[2 * j + 3], scale2, nmax2, l3);
                        }
                        ptx::tcgen05_st_32x32b_x16(t_s + c * 16, pa);
                        ptx::tcgen05_st_32x32b_x16(t_s + c * 16 + 16, pb);
                    }
                    lsum = (l0 + l1) + (l2s + l3);
#pragma unroll 1
                    for (; c < n32; ++c)
                    {
                        uint32_t v[32], pk[16];
                        ptx::tcgen05_ld_32x32b_x32(t_s + c * 32, v);
                        ptx::tcgen05_wait_ld();
                        if (c * 32 + 32 <= p.N)
                        {
#pragma unroll
                            for (int j = 0; j < 16; ++j)
                                pk[j] = att_exp_pair(__fmaf_rn(__uint_as_float(v[2 * j]), p.scale, -mxs),
                                                     __fmaf_rn(__uint_as_float(v[2 * j + 1]), p.scale, -mxs), lsum);
                        }
                        else
                        {
#pragma unroll
                            for (int j = 0; j < 16; ++j)
                            {
                                const int key = c * 32 + 2 * j;
                                float l2 = 0.f;
                                uint32_t e = att_exp_pair(__fmaf_rn(__uint_as_float(v[2 * j]), p.scale, -mxs),
                                                          __fmaf_rn(__uint_as_float(v[2 * j + 1]), p.scale, -mxs), l2);
                                if (key + 1 >= p.N) // masked keys contribute exactly zero
                                {
                                    if (key >= p.N) { e = 0u; l2 = 0.f; }
                                    else { e &= 0xFFFFu; l2 = __half2float(__ushort_as_half((unsigned short)(e & 0xFFFFu))); }
                                }
                                pk[j] = e;
                                lsum += l2;
                            }
                        }
                        ptx::tcgen05_st_32x32b_x16(t_s + c * 16, pk);
                    }
                    if (tail16)
                    {
                        uint32_t v[16], pk[8];
                        ptx::tcgen05_ld_32x32b_x16(t_s + n32 * 32, v);
                        ptx::tcgen05_wait_ld();
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                        {
                            const int key = n32 * 32 + 2 * j;
                            float l2 = 0.f;
                            uint32_t e = att_exp_pair(__fmaf_rn(__uint_as_float(v[2 * j]), p.scale, -mxs),
                                                      __fmaf_rn(__uint_as_float(v[2 * j + 1]), p.scale, -mxs), l2);
                            if (key + 1 >= p.N)
                            {
                                if (key >= p.N) { e = 0u; l2 = 0.f; }
                                else { e &= 0xFFFFu; l2 = __half2float(__ushort_as_half((unsigned short)(e & 0xFFFFu))); }
                            }
                            pk[j] = e;
                            lsum += l2;
                        }
                        ptx::tcgen05_st_32x32b_x8(t_s + n32 * 16, pk);
                    }
                    ptx::tcgen05_wait_st();
                }
                ptx::tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(p_ready(t));
                if (q == 0) ATT_TRACE(8 * t + 3);

                // ---- O_t = P_t V is complete: drain, release, normalise, store
                ptx::mbar_wait(o_full(t), i & 1);
                ptx::tcgen05_fence_after();
                if (q == 0) ATT_TRACE(8 * t + 4);
                uint32_t o[64];
                if (warp_valid)
                {
                    uint32_t(&o0)[32] = *reinterpret_cast<uint32_t(*)[32]>(&o[0]);
                    uint32_t(&o1)[32] = *reinterpret_cast<uint32_t(*)[32]>(&o[32]);
                    ptx::tcgen05_ld_32x32b_x32(t_o, o0);
                    ptx::tcgen05_ld_32x32b_x32(t_o + 32, o1);
                    ptx::tcgen05_wait_ld();
                }
                ptx::tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(o_empty(t));
                if (q == 0) ATT_TRACE(8 * t + 5);
                if (warp_valid)
                {
                    // normalise, f16, into this warp's 32-row x 128-B staging box (16-B chunk j of row r at j ^ (r & 7) =
                    // SWIZZLE_128B, conflict-free), then ONE TMA store; token rows >= N are clipped by the tensor map
                    const float inv = 1.0f / lsum; // p_i = e_i * (1/sum)  (ggml.c:10556-10558)
                    const uint32_t sbox = stage_out + (uint32_t)(warp_idx - 2) * 4096u;
                    if (lane == 0) ptx::tma_store_wait_read<0>(); // the previous problem's store has left the box
                    __syncwarp();
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                    {
                        uint32_t w[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e)
                        {
                            const __half2 hv = __floats2half2_rn(__uint_as_float(o[j * 8 + 2 * e]) * inv, __uint_as_float(o[j * 8 + 2 * e + 1]) * inv);
                            w[e] = *reinterpret_cast<const uint32_t *>(&hv);
                        }
                        ptx::st_shared_v4(sbox + (uint32_t)lane * 128u + (uint32_t)((j ^ (lane & 7)) << 4), w[0], w[1], w[2], w[3]);
                    }
                    ptx::fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0)
                    {
                        ptx::tma_store_3d(&tmO, sbox, h * 64, t * 128 + q * 32, b);
                        ptx::tma_store_commit();
                    }
                }
            }
            if (lane == 0) ptx::tma_store_wait_all();
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
