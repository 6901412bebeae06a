// attention_tcgen05_long.cuh -- fused attention on the 5th-gen tensor cores for sequences that do not fit the single-block
// kernel (224 < N <= 640 tokens -- K and V of one head, N x 128 B each, stay in shared memory; ViT-L/16-384 has N = 577).
//
// Same arithmetic as attention_tcgen05.cuh (reference vit.cpp:826-866; soft-max ggml.c:10498-10567): the TRUE row maximum is
// taken before any exponential, e = f16(exp(f16(s/8 - max))), l = sum e, O = (P V) * (1/l).  A row of scores (NKP f32) is
// wider than tensor memory, so the keys are cut into nb blocks of <= 96 and the work on one 128-row query tile runs in
// two sweeps over the blocks:
//   sweep A   S_j = Q K_j^T (tcgen05.mma SS, D in TMEM)  ->  row max only (S_j is discarded)
//   sweep B   S_j again (the tensor pipe is otherwise idle: recomputing costs 1/3 more attention flops, no memory traffic)
//             ->  P_j = exp(...) written over S_j as packed f16  ->  O += P_j V_j (tcgen05.mma TS, accumulated in TMEM)
// One CTA per SM owns one (image, head) at a time: K and V of the head stay resident in shared memory (TMA, SWIZZLE_128B,
// 64-row boxes), the Q tiles stream through a 2-deep ring.  Two soft-max warpgroups share a query tile: warpgroup w takes the
// key blocks j = w, w+2, ... and owns two of the four 96-column TMEM buffers (+ 64 columns of O), so the scores of its next
// block are already waiting when it finishes the current one (the MMA round trip hides behind the exponentials); row maxima
// and row sums are exchanged through shared memory between the sweeps.
// Warp roles (10 warps): 0 TMA producer, 1 MMA issuer (one thread, fixed in-order script per tile), 2-5 warpgroup 0,
// 6-9 warpgroup 1 (TMEM lane quarter = warp index mod 4).  Warpgroup 0 also normalises and stores O (swizzled staging + one
// TMA store per warp; token rows >= N are clipped by the 3-D tensor map).
#pragma once
#include "attention_tcgen05.cuh"

namespace vitb200 {

constexpr int ATT_LONG_THREADS = 320;

struct AttnLongParams
{
    int N, D, H;     // tokens per image, hidden, heads
    int n_problems;  // B * H
    int NKP;         // keys padded to a multiple of 16
    int kv_rows;     // NKP rounded up to the 64-row TMA box
    int n_tiles;     // query tiles of 128 rows
    int nb;          // key blocks (even, 4..ATT_LONG_MAX_BLOCKS) of <= 96 keys
    int key0[9];     // first key of block j (att_long_block_key0), key0[nb] = NKP: read from the constant bank, no divisions on the device
    float scale;     // 1/sqrt(64)
    long long *trace; // dev only (VITB200_ATTN_TRACE): clock64 stamps of CTA 0, [tile < 16][slot < 32] (warpgroup w writes slots 8 w ..); NULL in production
};

#define ATT_LONG_TRACE(slot) do { if (p.trace && blockIdx.x == 0 && tile_seq < 16 && q == 0 && lane == 0) p.trace[tile_seq * 32 + 8 * w + (slot)] = clock64(); } while (0)

constexpr int ATT_LONG_BUF_COLS = 96, ATT_LONG_OCOL = 384, ATT_LONG_MAX_BLOCKS = 8, ATT_LONG_MAX_KEYS = 640;

// TMEM buffer of key block j: warpgroup j & 1 owns buffers 2 (j & 1) and 2 (j & 1) + 1 and alternates between them
__host__ __device__ __forceinline__ int att_long_buf(int j) { return ((j & 1) << 1) | ((j >> 1) & 1); }

// Key blocks: the ceil(NKP/32) 32-key chunks are dealt out evenly (the first blocks get the extra chunk), so every block is a
// whole number of chunks except the last, which ends at NKP (its final chunk may be 16 keys wide and holds the keys >= N).
__host__ __device__ __forceinline__ int att_long_block_key0(int nkp, int nb, int j)
{
    const int chunks = (nkp + 31) >> 5, base = chunks / nb, extra = chunks % nb;
    return (j * base + (j < extra ? j : extra)) * 32;
}
__host__ __device__ __forceinline__ int att_long_block_keys(int nkp, int nb, int j)
{
    const int k0 = att_long_block_key0(nkp, nb, j), k1 = j + 1 < nb ? att_long_block_key0(nkp, nb, j + 1) : nkp;
    return k1 - k0;
}

__global__ void __launch_bounds__(ATT_LONG_THREADS, 1)
attention_tc_long_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV64,
                         const __grid_constant__ CUtensorMap tmO, const __grid_constant__ AttnLongParams p)
{
    extern __shared__ uint8_t att_long_smem_raw[];
    const uint32_t smem_base = (ptx::smem_u32(att_long_smem_raw) + 1023u) & ~1023u;
    uint8_t *smem = att_long_smem_raw + (smem_base - ptx::smem_u32(att_long_smem_raw));
    const uint32_t kv_bytes = (uint32_t)p.kv_rows * 128u;
    const uint32_t sK = smem_base, sV = smem_base + kv_bytes;
    const uint32_t sQ0 = sV + kv_bytes;                  // 2 x 16 KB
    const uint32_t stage_out = sQ0 + 2 * 16384;          // 4 x 4 KB store boxes (warpgroup 0)
    const uint32_t ex_off = 2 * kv_bytes + 2 * 16384 + 4 * 4096;
    float *mx_ex = reinterpret_cast<float *>(smem + ex_off);         // [2][128]
    float *l_ex = mx_ex + 256;                                       // [2][128]
    const uint32_t bars = smem_base + ex_off + 2048;
    // barriers: k_full, k_empty, v_full, o_full, o_empty, q_full[2], q_empty[2], s_full[4], s_free[4], p_ready[4], v_empty, tmem ptr
    const uint32_t k_full = bars, k_empty = bars + 8, v_full = bars + 16, o_full = bars + 24, o_empty = bars + 32, v_empty = bars + 8u * 21;
    auto q_full = [&](int i) { return bars + 8u * (5 + i); };
    auto q_empty = [&](int i) { return bars + 8u * (7 + i); };
    auto s_full = [&](int b) { return bars + 8u * (9 + b); };
    auto s_free = [&](int b) { return bars + 8u * (13 + b); };
    auto p_ready = [&](int b) { return bars + 8u * (17 + b); };
    const uint32_t tmem_ptr_addr = bars + 8u * 22;
    volatile uint32_t *tmem_ptr_gen = reinterpret_cast<volatile uint32_t *>(smem + ex_off + 2048 + 8 * 22);

    const int warp_idx = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp_idx == 0 && lane == 0)
    {
        ptx::prefetch_tensormap(&tmQ);
        ptx::prefetch_tensormap(&tmKV64);
        ptx::prefetch_tensormap(&tmO);
    }
    if (warp_idx == 1 && lane == 0)
    {
        ptx::mbar_init(k_full, 1);
        ptx::mbar_init(v_full, 1);
        ptx::mbar_init(k_empty, 1);
        ptx::mbar_init(v_empty, 1);
        for (int i = 0; i < 2; ++i)
        {
            ptx::mbar_init(q_full(i), 1);
            ptx::mbar_init(q_empty(i), 1);
        }
        for (int i = 0; i < 4; ++i)
        {
            ptx::mbar_init(s_full(i), 1);
            ptx::mbar_init(s_free(i), 4);
            ptx::mbar_init(p_ready(i), 4);
        }
        ptx::mbar_init(o_full, 1);
        ptx::mbar_init(o_empty, 4);
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
    const int nb = p.nb;

    if (warp_idx == 0)
    {
        // ===================== TMA producer: K, V of the head once, then its Q tiles through a 2-deep ring =====================
        if (lane == 0)
        {
            int ip = 0, iq = 0;
            const int nbox = p.kv_rows / 64;
            for (int prob = blockIdx.x; prob < p.n_problems; prob += gridDim.x, ++ip)
            {
                const int b = prob / p.H, h = prob - b * p.H;
                const int row0 = b * p.N;
                ptx::mbar_wait(k_empty, (ip & 1) ^ 1); // K is free once the previous head's last scores are done: this load
                                                        // overlaps that head's remaining exponentials, P V and epilogue
                ptx::mbar_arrive_expect_tx(k_full, (uint32_t)(nbox * 8192));
                for (int x = 0; x < nbox; ++x) ptx::tma_load_2d(sK + x * 8192, &tmKV64, k_full, p.D + h * 64, row0 + x * 64);
                for (int ti = 0; ti < p.n_tiles; ++ti, ++iq)
                {
                    const int qb = iq & 1;
                    ptx::mbar_wait(q_empty(qb), ((iq >> 1) & 1) ^ 1);
                    ptx::mbar_arrive_expect_tx(q_full(qb), 16384u);
                    ptx::tma_load_2d(sQ0 + qb * 16384, &tmQ, q_full(qb), h * 64, row0 + ti * 128);
                    if (ti == 0) // V is first needed in sweep B of the first tile: it queues behind K and the first Q tile
                    {
                        ptx::mbar_wait(v_empty, (ip & 1) ^ 1);
                        ptx::mbar_arrive_expect_tx(v_full, (uint32_t)(nbox * 8192));
                        for (int x = 0; x < nbox; ++x) ptx::tma_load_2d(sV + x * 8192, &tmKV64, v_full, 2 * p.D + h * 64, row0 + x * 64);
                    }
                }
            }
        }
        __syncwarp();
    }
    else if (warp_idx == 1)
    {
        // ===================== MMA issuer: a fixed in-order script per query tile.  The whole warp runs it (so descriptors stay in
        // uniform registers and the MMAs go out back to back); one elected lane issues. =====================
        {
            const uint32_t idesc_o = ptx::umma_idesc_f16(128, 64, 0, 0, 0, /*B (V) is MN-major*/ 1);
            uint32_t c_sfree[4] = {0, 0, 0, 0}, c_pready[4] = {0, 0, 0, 0}, c_oempty = 0;
            int ip = 0, iq = 0;
            bool first_tile_of_cta = true;
            auto issue_s = [&](int qb, int j) {
                const int bb = att_long_buf(j), key0 = p.key0[j], kb = p.key0[j + 1] - key0;
                const uint32_t idesc_s = ptx::umma_idesc_f16(128, kb, 0, 0, 0, 0);
                const uint64_t kdesc = ptx::umma_desc_kmajor_sw128(sK + (uint32_t)key0 * 128u);
                const uint64_t qdesc = ptx::umma_desc_kmajor_sw128(sQ0 + qb * 16384);
                if (ptx::elect_one())
                {
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        ptx::tcgen05_mma_f16(tmem_base + (uint32_t)(bb * ATT_LONG_BUF_COLS), qdesc + 2 * k, kdesc + 2 * k, idesc_s, k > 0);
                    ptx::tcgen05_commit(s_full(bb));
                }
                __syncwarp();
            };
            // P_j V_j: 16 keys per step = 8 TMEM columns of packed f16 (A) x 16 rows (2048 B) of V (B), accumulated into O
            auto issue_pv = [&](int j, bool first) {
                const int bb = att_long_buf(j), key0 = p.key0[j], kb = p.key0[j + 1] - key0;
                const uint64_t vdesc = ptx::umma_desc_mnmajor_sw128(sV, kv_bytes) + (uint64_t)((key0 / 16) * 128);
                const uint32_t ta = tmem_base + (uint32_t)(bb * ATT_LONG_BUF_COLS);
                if (ptx::elect_one())
                {
                    const int ks = kb / 16;
                    ptx::tcgen05_mma_f16_ts(tmem_base + ATT_LONG_OCOL, ta, vdesc, idesc_o, first ? 0u : 1u);
                    for (int k = 1; k < ks; ++k)
                        ptx::tcgen05_mma_f16_ts(tmem_base + ATT_LONG_OCOL, ta + 8 * k, vdesc + (uint64_t)(k * 128), idesc_o, 1u);
                }
                __syncwarp();
            };
            for (int prob = blockIdx.x; prob < p.n_problems; prob += gridDim.x, ++ip)
            {
                ptx::mbar_wait(k_full, ip & 1);
                bool v_ready = false;
                for (int ti = 0; ti < p.n_tiles; ++ti, ++iq)
                {
                    const int qb = iq & 1;
                    ptx::mbar_wait(q_full(qb), (iq >> 1) & 1);
                    ptx::tcgen05_fence_after();
                    // sweep A: scores for the row maximum.  A buffer is reusable once its warpgroup has read the block before.
                    for (int j = 0; j < nb; ++j)
                    {
                        const int bb = att_long_buf(j);
                        if (j >= 4) { ptx::mbar_wait(s_free(bb), c_sfree[bb]++ & 1); ptx::tcgen05_fence_after(); }
                        issue_s(qb, j);
                    }
                    // sweep B: scores again, now followed by P_j V_j.  Buffer(j) holds P_{j-4} until P_{j-4} V_{j-4} has been issued
                    // (the tensor pipe runs in issue order, so S_j may follow it immediately).
                    bool first = true;
                    for (int j = 0; j < nb + 4; ++j)
                    {
                        const int bb = att_long_buf(j);
                        if (j < 4) { ptx::mbar_wait(s_free(bb), c_sfree[bb]++ & 1); ptx::tcgen05_fence_after(); }
                        else
                        {
                            ptx::mbar_wait(p_ready(bb), c_pready[bb]++ & 1);
                            if (first && !first_tile_of_cta) ptx::mbar_wait(o_empty, c_oempty++ & 1); // previous tile's O drained
                            if (!v_ready) { ptx::mbar_wait(v_full, ip & 1); v_ready = true; }
                            ptx::tcgen05_fence_after();
                            issue_pv(j - 4, first);
                            first = false;
                        }
                        if (j < nb) issue_s(qb, j);
                        if (j == nb - 1 && ptx::elect_one())
                        {
                            ptx::tcgen05_commit(q_empty(qb)); // every MMA reading this Q tile is issued
                            if (ti == p.n_tiles - 1) ptx::tcgen05_commit(k_empty); // ... and, on the head's last tile, every MMA reading K
                        }
                    }
                    if (ptx::elect_one()) ptx::tcgen05_commit(o_full);
                    __syncwarp();
                    first_tile_of_cta = false;
                }
                if (ptx::elect_one()) ptx::tcgen05_commit(v_empty); // V may be overwritten once every MMA of this head has retired
                __syncwarp();
            }
        }
    }
    else
    {
        // ===================== soft-max warpgroups =====================
        const int w = (warp_idx - 2) >> 2; // warpgroup = TMEM buffer = parity of the key blocks it owns
        const int q = warp_idx & 3;        // TMEM lane quarter
        const int row_in_tile = q * 32 + lane;
        const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16);
        const uint32_t t_o = t_lane + ATT_LONG_OCOL;
        uint32_t c_sfull[2] = {0, 0}, c_ofull = 0; // s_full parity of this warpgroup's two buffers
        int tile_seq = 0; // tiles this CTA has processed (trace index)
        for (int prob = blockIdx.x; prob < p.n_problems; prob += gridDim.x)
        {
            const int b = prob / p.H, h = prob - b * p.H;
            for (int ti = 0; ti < p.n_tiles; ++ti, ++tile_seq)
            {
                const bool warp_valid = (ti * 128 + q * 32) < p.N; // warp owns at least one real query row
                ATT_LONG_TRACE(0);
                // ---- sweep A: true row maximum over the valid keys (ggml.c:10533-10534), this warpgroup's blocks
                float mx = -INFINITY;
                for (int j = w; j < nb; j += 2)
                {
                    const int key0 = p.key0[j], kb = p.key0[j + 1] - key0;
                    const int bb = att_long_buf(j);
                    const uint32_t t_s = t_lane + (uint32_t)(bb * ATT_LONG_BUF_COLS);
                    ptx::mbar_wait(s_full(bb), c_sfull[bb & 1]++ & 1);
                    ptx::tcgen05_fence_after();
                    if (warp_valid)
                    {
                        const int valid = p.N - key0 < kb ? p.N - key0 : kb; // valid keys in this block (only the last block is short)
                        float ma = -INFINITY, mb = -INFINITY;
                        int c = 0;
                        for (; (c + 2) * 32 <= valid; c += 2) // two mask-free chunks per iteration
                        {
                            uint32_t va[32], vb[32];
                            ptx::tcgen05_ld_32x32b_x32(t_s + c * 32, va);
                            ptx::tcgen05_ld_32x32b_x32(t_s + c * 32 + 32, vb);
                            ptx::tcgen05_wait_ld();
#pragma unroll
                            for (int x = 0; x < 32; ++x) { ma = fmaxf(ma, __uint_as_float(va[x])); mb = fmaxf(mb, __uint_as_float(vb[x])); }
                        }
                        mx = fmaxf(mx, fmaxf(ma, mb));
                        for (; c * 32 < valid; ++c) // leftover chunks: 16 columns at a time so no column past the block is read
                        {
#pragma unroll
                            for (int hh = 0; hh < 2; ++hh)
                            {
                                const int col = c * 32 + hh * 16;
                                if (col < valid)
                                {
                                    uint32_t v[16];
                                    ptx::tcgen05_ld_32x32b_x16(t_s + col, v);
                                    ptx::tcgen05_wait_ld();
#pragma unroll
                                    for (int x = 0; x < 16; ++x)
                                        if (col + x < valid) mx = fmaxf(mx, __uint_as_float(v[x]));
                                }
                            }
                        }
                    }
                    ptx::tcgen05_fence_before();
                    __syncwarp();
                    if (lane == 0) ptx::mbar_arrive(s_free(bb));
                }
                ATT_LONG_TRACE(1);
                mx_ex[w * 128 + row_in_tile] = mx;
                ptx::named_bar_sync(1, 256);
                ATT_LONG_TRACE(2);
                mx = fmaxf(mx_ex[row_in_tile], mx_ex[128 + row_in_tile]);
                const float mxs = mx * p.scale; // ggml_scale_inplace (vit.cpp:851-854); exact, scale = 1/8
                const uint64_t scale2 = ptx::pack_f32x2(p.scale, p.scale), nmax2 = ptx::pack_f32x2(-mxs, -mxs);

                // ---- sweep B: P = f16(exp(f16(s*scale - max))) over S in place, l = sum P
                float l0 = 0.f, l1 = 0.f, l2s = 0.f, l3 = 0.f;
                for (int j = w; j < nb; j += 2)
                {
                    const int key0 = p.key0[j], kb = p.key0[j + 1] - key0;
                    const int bb = att_long_buf(j);
                    const uint32_t t_s = t_lane + (uint32_t)(bb * ATT_LONG_BUF_COLS);
                    ptx::mbar_wait(s_full(bb), c_sfull[bb & 1]++ & 1);
                    ptx::tcgen05_fence_after();
                    if (warp_valid)
                    {
                        const int valid = p.N - key0 < kb ? p.N - key0 : kb;
                        // Mask-free 32-key chunks two at a time (two tcgen05.ld.x32, one wait, 64 exponentials with packed FP32 scale / shift /
                        // log2(e), four partial sums) -- the single-block kernel's loop.  The first version here went 16 keys at a time with
                        // the next granule's load in flight; the phase trace (tools/attn_long_trace.py) had it at 26-29 clocks per key
                        // against 18 for this form: per warp the TMEM round trip is already covered by the other soft-max warp of the
                        // sub-partition, and short steps only add waits and stores (same finding as profiles/microbench_r02.md).
                        const int ng = kb >> 4, nc32 = valid >> 5; // 16-key granules in the block / chunks with all 32 keys valid
                        int c = 0;
#pragma unroll 1
                        for (; c + 2 <= nc32; c += 2)
                        {
                            uint32_t va[32], vb[32], pa[16], pb[16];
                            ptx::tcgen05_ld_32x32b_x32(t_s + c * 32, va);
                            ptx::tcgen05_ld_32x32b_x32(t_s + c * 32 + 32, vb);
                            ptx::tcgen05_wait_ld();
#pragma unroll
                            for (int x = 0; x < 16; x += 2)
                            {
                                pa[x] = att_exp_pair_raw(va[2 * x], va[2 * x + 1], scale2, nmax2, l0);
                                pb[x] = att_exp_pair_raw(vb[2 * x], vb[2 * x + 1], scale2, nmax2, l1);
                                pa[x + 1] = att_exp_pair_raw(va[2 * x + 2], va[2 * x + 3], scale2, nmax2, l2s);
                                pb[x + 1] = att_exp_pair_raw(vb[2 * x + 2], vb[2 * x + 3], scale2, nmax2, l3);
                            }
                            ptx::tcgen05_st_32x32b_x16(t_s + c * 16, pa);
                            ptx::tcgen05_st_32x32b_x16(t_s + c * 16 + 16, pb);
                        }
#pragma unroll 1
                        for (; c < nc32; ++c)
                        {
                            uint32_t va[32], pa[16];
                            ptx::tcgen05_ld_32x32b_x32(t_s + c * 32, va);
                            ptx::tcgen05_wait_ld();
#pragma unroll
                            for (int x = 0; x < 16; x += 4)
                            {
                                pa[x] = att_exp_pair_raw(va[2 * x], va[2 * x + 1], scale2, nmax2, l0);
                                pa[x + 1] = att_exp_pair_raw(va[2 * x + 2], va[2 * x + 3], scale2, nmax2, l1);
                                pa[x + 2] = att_exp_pair_raw(va[2 * x + 4], va[2 * x + 5], scale2, nmax2, l2s);
                                pa[x + 3] = att_exp_pair_raw(va[2 * x + 6], va[2 * x + 7], scale2, nmax2, l3);
                            }
                            ptx::tcgen05_st_32x32b_x16(t_s + c * 16, pa);
                        }
                        uint32_t va[16], pe[8];
                        int g = 0;
                        // the granule straddling N (keys >= N get exactly zero) and the all-padding granules behind it
                        for (g = nc32 * 2; g < ng; ++g)
                        {
                            const int col = g * 16;
                            if (col < valid)
                            {
                                ptx::tcgen05_ld_32x32b_x16(t_s + col, va);
                                ptx::tcgen05_wait_ld();
                            }
#pragma unroll
                            for (int x = 0; x < 8; ++x)
                            {
                                const int key = col + 2 * x;
                                uint32_t e = 0u;
                                float l2 = 0.f;
                                if (key < valid)
                                {
                                    e = att_exp_pair(__fmaf_rn(__uint_as_float(va[2 * x]), p.scale, -mxs),
                                                     key + 1 < valid ? __fmaf_rn(__uint_as_float(va[2 * x + 1]), p.scale, -mxs) : 0.f, l2);
                                    if (key + 1 >= valid) { e &= 0xFFFFu; l2 = __half2float(__ushort_as_half((unsigned short)(e & 0xFFFFu))); }
                                }
                                pe[x] = e;
                                l0 += l2;
                            }
                            ptx::tcgen05_st_32x32b_x8(t_s + g * 8, pe);
                        }
                        ptx::tcgen05_wait_st();
                    }
                    ptx::tcgen05_fence_before();
                    __syncwarp();
                    if (lane == 0) ptx::mbar_arrive(p_ready(bb));
                    if (p.trace && blockIdx.x == 0 && tile_seq < 16 && q == 0 && lane == 0 && (j >> 1) < 4) p.trace[tile_seq * 32 + 16 + 4 * w + (j >> 1)] = clock64();
                }
                ATT_LONG_TRACE(3);
                l_ex[w * 128 + row_in_tile] = (l0 + l1) + (l2s + l3);
                ptx::named_bar_sync(2, 256);
                ATT_LONG_TRACE(4);

                // ---- O = P V is complete: warpgroup 0 drains, releases, normalises and stores
                if (w == 0)
                {
                    const float lsum = l_ex[row_in_tile] + l_ex[128 + row_in_tile];
                    ptx::mbar_wait(o_full, c_ofull++ & 1);
                    ptx::tcgen05_fence_after();
                    ATT_LONG_TRACE(5);
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
                    if (lane == 0) ptx::mbar_arrive(o_empty);
                    if (warp_valid)
                    {
                        const float inv = 1.0f / lsum; // p_i = e_i * (1/sum)  (ggml.c:10556-10558)
                        const uint32_t sbox = stage_out + (uint32_t)q * 4096u;
                        if (lane == 0) ptx::tma_store_wait_read<0>(); // the previous tile's store has left the box
                        __syncwarp();
#pragma unroll
                        for (int x = 0; x < 8; ++x)
                        {
                            uint32_t wv[4];
#pragma unroll
                            for (int e = 0; e < 4; ++e)
                            {
                                const __half2 hv = __floats2half2_rn(__uint_as_float(o[x * 8 + 2 * e]) * inv, __uint_as_float(o[x * 8 + 2 * e + 1]) * inv);
                                wv[e] = *reinterpret_cast<const uint32_t *>(&hv);
                            }
                            ptx::st_shared_v4(sbox + (uint32_t)lane * 128u + (uint32_t)((x ^ (lane & 7)) << 4), wv[0], wv[1], wv[2], wv[3]);
                        }
                        ptx::fence_proxy_async_smem();
                        __syncwarp();
                        if (lane == 0)
                        {
                            ptx::tma_store_3d(&tmO, sbox, h * 64, ti * 128 + q * 32, b);
                            ptx::tma_store_commit();
                        }
                    }
                    ATT_LONG_TRACE(6);
                }
            }
        }
        if (w == 0 && lane == 0) ptx::tma_store_wait_all();
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
