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
// 64-row boxes).  ONE WARPGROUP PER QUERY TILE, the single-block kernel's structure: warpgroup w owns every second query tile of the
// CTA's tile sequence -- both sweeps over ALL key blocks, its own score buffers and O accumulator (TMEM: 2 x 192 + 2 x 64 = 512
// columns), its own Q slot, its own MMA-issuing warp running a fixed in-order script per tile, its own epilogue -- so nothing is
// exchanged between the warpgroups, and the phases in which one of them does not exponentiate (row-max sweep, waiting for O, epilogue)
// are filled by the other's exponentials.  Sweep A streams 64-key blocks through four 64-column buffers (the warpgroup's 192 score
// columns + its O columns, free until the tile's first P V): it is pure MMA round-trip latency, and four blocks in flight hide most of
// it; sweep B uses the nb <= 96-key blocks through two 96-column buffers (the next scores are waiting when a block's exponentials end).
// The first version had the two warpgroups SHARE every tile (alternate key blocks, row maxima / sums exchanged through shared memory,
// one MMA issuer): 16.8 ms of attention per ViT-L/16-384 forward at batch 128 against 14.1 ms for this one (tools/attn_long_trace.py: of
// a 14 700-clock tile, 5 500 were the max sweep, the wait for O and the epilogue, with the MUFU idle).  What is left: a head has five
// query tiles, so one warpgroup gets three and the other waits about one tile time at the head boundary (K is reloaded only when both
// issuers have committed their last scores of the head; a second K buffer does not fit next to 2 x 80 KB of K and V).
// Warps (11): 0 TMA producer, 1 MMA issuer of warpgroup 0, 2-5 warpgroup 0, 6-9 warpgroup 1, 10 MMA issuer of warpgroup 1 (TMEM lane
// quarter = warp index mod 4).  Epilogue: O * (1/l) -> f16 -> swizzled staging box -> one TMA store per warp; token rows >= N are
// clipped by the 3-D tensor map.
#pragma once
#include "attention_tcgen05.cuh"

namespace vitb200 {


struct AttnLongParams
{
    int N, D, H;     // tokens per image, hidden, heads
    int n_problems;  // B * H
    int NKP;         // keys padded to a multiple of 16
    int kv_rows;     // NKP rounded up to the 64-row TMA box
    int n_tiles;     // query tiles of 128 rows
    int nb;          // sweep B's key blocks (2..ATT_LONG_MAX_BLOCKS) of <= 96 keys
    int key0[9];     // first key of block j (att_long_block_key0), key0[nb] = NKP: read from the constant bank, no divisions on the device
    float scale;     // 1/sqrt(64)
    unsigned long long load_policy; // L2 eviction hint of the operand loads (0 = none): q, k, v are read exactly once
    long long *trace; // dev only (VITB200_ATTN_TRACE): clock64 stamps of CTA 0, [tile < 16][slot < 32] (warpgroup w writes slots 8 w ..); NULL in production
};

#define ATT_LONG_TRACE(slot) do { if (p.trace && blockIdx.x == 0 && tile_seq < 16 && q == 0 && lane == 0) p.trace[tile_seq * 32 + 8 * w + (slot)] = clock64(); } while (0)

constexpr int ATT_LONG_BUF_COLS = 96, ATT_LONG_OCOL = 384, ATT_LONG_MAX_BLOCKS = 8, ATT_LONG_MAX_KEYS = 640;

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

constexpr int ATT_LONG_THREADS = 352;
__host__ __device__ inline int attention_tc_long_smem_bytes(int kv_rows) { return 1024 + 2 * kv_rows * 128 + 2 * 16384 + 8 * 4096 + 512; }

__global__ void __launch_bounds__(ATT_LONG_THREADS, 1)
attention_tc_long_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV64,
                          const __grid_constant__ CUtensorMap tmO, const __grid_constant__ AttnLongParams p)
{
    extern __shared__ uint8_t att_long_smem_raw[];
    const uint32_t smem_base = (ptx::smem_u32(att_long_smem_raw) + 1023u) & ~1023u;
    uint8_t *smem = att_long_smem_raw + (smem_base - ptx::smem_u32(att_long_smem_raw));
    const uint32_t kv_bytes = (uint32_t)p.kv_rows * 128u;
    const uint32_t sK = smem_base, sV = smem_base + kv_bytes;
    const uint32_t sQ0 = sV + kv_bytes;                  // 2 x 16 KB: slot w belongs to warpgroup w
    const uint32_t stage_out = sQ0 + 2 * 16384;          // 8 x 4 KB store boxes
    const uint32_t bar_off = 2 * kv_bytes + 2 * 16384 + 8 * 4096;
    const uint32_t bars = smem_base + bar_off;
    // barriers: k_full, k_empty (2), v_full, v_empty (2), q_full[2], q_empty[2], o_full[2], o_empty[2] (4), s_full[2][2], s_free[2][2] (4),
    // p_ready[2][2] (4), tmem ptr
    const uint32_t k_full = bars, k_empty = bars + 8, v_full = bars + 16, v_empty = bars + 24;
    auto q_full = [&](int w) { return bars + 8u * (4 + w); };
    auto q_empty = [&](int w) { return bars + 8u * (6 + w); };
    auto o_full = [&](int w) { return bars + 8u * (8 + w); };
    auto o_empty = [&](int w) { return bars + 8u * (10 + w); };
    auto s_full = [&](int w, int bb) { return bars + 8u * (12 + 2 * w + bb); };
    auto s_free = [&](int w, int bb) { return bars + 8u * (16 + 2 * w + bb); };
    auto p_ready = [&](int w, int bb) { return bars + 8u * (20 + 2 * w + bb); };
    // sweep A runs over 64-key blocks through FOUR buffers per warpgroup (its two 96-column score buffers cut into three 64-column ones + its
    // O accumulator's 64 columns, free until the first P V of the tile): the row-max sweep is pure round-trip latency (an MMA + ~250 clocks of
    // tcgen05.ld / FMNMX per block), and two buffers kept only one block in flight (5 900 clocks per tile, tools/attn_long_trace.py)
    auto sa_full = [&](int w, int bb) { return bars + 8u * (24 + 4 * w + bb); };
    auto sa_free = [&](int w, int bb) { return bars + 8u * (32 + 4 * w + bb); };
    const uint32_t tmem_ptr_addr = bars + 8u * 40;
    volatile uint32_t *tmem_ptr_gen = reinterpret_cast<volatile uint32_t *>(smem + bar_off + 8 * 40);

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
        ptx::mbar_init(k_empty, 2);
        ptx::mbar_init(v_empty, 2);
        for (int w = 0; w < 2; ++w)
        {
            ptx::mbar_init(q_full(w), 1);
            ptx::mbar_init(q_empty(w), 1);
            ptx::mbar_init(o_full(w), 1);
            ptx::mbar_init(o_empty(w), 4);
            for (int bb = 0; bb < 2; ++bb)
            {
                ptx::mbar_init(s_full(w, bb), 1);
                ptx::mbar_init(s_free(w, bb), 4);
                ptx::mbar_init(p_ready(w, bb), 4);
            }
            for (int bb = 0; bb < 4; ++bb)
            {
                ptx::mbar_init(sa_full(w, bb), 1);
                ptx::mbar_init(sa_free(w, bb), 4);
            }
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
    ptx::grid_dep_launch(); // PDL: the prologue above overlapped the previous kernel's tail
    ptx::grid_dep_wait();
    const int nb = p.nb;
    // TMEM: warpgroup w's score buffers at columns 192 w + 96 bb, its O accumulator at 384 + 64 w
    auto scol = [&](int w, int bb) { return (uint32_t)(192 * w + ATT_LONG_BUF_COLS * bb); };
    auto ocol = [&](int w) { return (uint32_t)(ATT_LONG_OCOL + 64 * w); };
    auto sacol = [&](int w, int bb) { return bb < 3 ? (uint32_t)(192 * w + 64 * bb) : ocol(w); }; // sweep A's 64-column buffers
    const int ka = (p.NKP + 63) >> 6;                                                              // sweep A's 64-key blocks

    if (warp_idx == 0)
    {
        // ===================== TMA producer: K, V of the head once; Q tile number s of the CTA goes to slot s & 1 =====================
        if (lane == 0)
        {
            auto ld3 = [&](uint32_t dst, const CUtensorMap *m, uint32_t bar, int c0, int c1, int c2) {
                if (p.load_policy) ptx::tma_load_3d_hint(dst, m, bar, c0, c1, c2, p.load_policy);
                else ptx::tma_load_3d(dst, m, bar, c0, c1, c2);
            };
            int ip = 0, iq = 0;
            const int nbox = p.kv_rows / 64;
            for (int prob = blockIdx.x; prob < p.n_problems; prob += gridDim.x, ++ip)
            {
                const int b = prob / p.H, h = prob - b * p.H;
                const int row0 = b * p.N;
                ptx::mbar_wait(k_empty, (ip & 1) ^ 1); // both issuers have committed their last scores of the previous head
                ptx::mbar_arrive_expect_tx(k_full, (uint32_t)(nbox * 8192));
                for (int x = 0; x < nbox; ++x) ld3(sK + x * 8192, &tmKV64, k_full, 0, row0 + x * 64, p.H + h);
                for (int ti = 0; ti < p.n_tiles; ++ti, ++iq)
                {
                    const int qb = iq & 1;
                    ptx::mbar_wait(q_empty(qb), ((iq >> 1) & 1) ^ 1);
                    ptx::mbar_arrive_expect_tx(q_full(qb), 16384u);
                    ld3(sQ0 + qb * 16384, &tmQ, q_full(qb), 0, row0 + ti * 128, h);
                    if (ti == 0)
                    {
                        ptx::mbar_wait(v_empty, (ip & 1) ^ 1);
                        ptx::mbar_arrive_expect_tx(v_full, (uint32_t)(nbox * 8192));
                        for (int x = 0; x < nbox; ++x) ld3(sV + x * 8192, &tmKV64, v_full, 0, row0 + x * 64, 2 * p.H + h);
                    }
                }
            }
        }
        __syncwarp();
    }
    else if (warp_idx == 1 || warp_idx == 10)
    {
        // ===================== MMA issuer of warpgroup w: a fixed in-order script per query tile of that warpgroup =====================
        const int w = warp_idx == 1 ? 0 : 1;
        const uint32_t idesc_o = ptx::umma_idesc_f16(128, 64, 0, 0, 0, /*B (V) is MN-major*/ 1);
        uint32_t c_pready[2] = {0, 0}, c_oempty = 0, c_qfull = 0, sa_phase = 0; // sa_phase: bit bb = parity of sa_free(w, bb)
        bool first_tile_of_warp = true;
        auto issue_s = [&](int j) {
            const int bb = j & 1, key0 = p.key0[j], kb = p.key0[j + 1] - key0;
            const uint32_t idesc_s = ptx::umma_idesc_f16(128, kb, 0, 0, 0, 0);
            const uint64_t kdesc = ptx::umma_desc_kmajor_sw128(sK + (uint32_t)key0 * 128u);
            const uint64_t qdesc = ptx::umma_desc_kmajor_sw128(sQ0 + w * 16384);
            if (ptx::elect_one())
            {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    ptx::tcgen05_mma_f16(tmem_base + scol(w, bb), qdesc + 2 * k, kdesc + 2 * k, idesc_s, k > 0);
                ptx::tcgen05_commit(s_full(w, bb));
            }
            __syncwarp();
        };
        auto issue_pv = [&](int j, bool first) {
            const int bb = j & 1, key0 = p.key0[j], kb = p.key0[j + 1] - key0;
            const uint64_t vdesc = ptx::umma_desc_mnmajor_sw128(sV, kv_bytes) + (uint64_t)((key0 / 16) * 128);
            const uint32_t ta = tmem_base + scol(w, bb);
            if (ptx::elect_one())
            {
                const int ks = kb / 16;
                ptx::tcgen05_mma_f16_ts(tmem_base + ocol(w), ta, vdesc, idesc_o, first ? 0u : 1u);
                for (int k = 1; k < ks; ++k)
                    ptx::tcgen05_mma_f16_ts(tmem_base + ocol(w), ta + 8 * k, vdesc + (uint64_t)(k * 128), idesc_o, 1u);
            }
            __syncwarp();
        };
        int ip = 0, iq = 0;
        for (int prob = blockIdx.x; prob < p.n_problems; prob += gridDim.x, ++ip)
        {
            bool kv_seen = false;
            for (int ti = 0; ti < p.n_tiles; ++ti, ++iq)
            {
                if ((iq & 1) != w) continue; // the other warpgroup's tile
                const bool last_of_head = ti + 2 >= p.n_tiles; // this warpgroup's last tile of the head
                if (!kv_seen) { ptx::mbar_wait(k_full, ip & 1); }
                ptx::mbar_wait(q_full(w), c_qfull++ & 1);
                ptx::tcgen05_fence_after();
                // sweep A: scores for the row maximum, 64 keys at a time through four buffers.  Buffers 0-2 overlap the previous tile's P
                // (its P V were issued by this warp: in order on the tensor pipe); buffer 3 is the O accumulator, free once drained.
                {
                    const uint64_t qdesc = ptx::umma_desc_kmajor_sw128(sQ0 + w * 16384);
                    for (int j = 0; j < ka; ++j)
                    {
                        const int bb = j & 3;
                        if (j >= 4) { ptx::mbar_wait(sa_free(w, bb), (sa_phase >> bb) & 1); sa_phase ^= 1u << bb; ptx::tcgen05_fence_after(); }
                        else if (j == 3 && !first_tile_of_warp) { ptx::mbar_wait(o_empty(w), c_oempty++ & 1); ptx::tcgen05_fence_after(); }
                        const int key0 = j * 64, kb = p.NKP - key0 < 64 ? p.NKP - key0 : 64;
                        const uint32_t idesc_a = ptx::umma_idesc_f16(128, kb, 0, 0, 0, 0);
                        const uint64_t kdesc = ptx::umma_desc_kmajor_sw128(sK + (uint32_t)key0 * 128u);
                        if (ptx::elect_one())
                        {
#pragma unroll
                            for (int k = 0; k < 4; ++k)
                                ptx::tcgen05_mma_f16(tmem_base + sacol(w, bb), qdesc + 2 * k, kdesc + 2 * k, idesc_a, k > 0);
                            ptx::tcgen05_commit(sa_full(w, bb));
                        }
                        __syncwarp();
                    }
                    // every sweep-A block has been read before sweep B's scores overwrite the buffers (and P V the O columns)
                    for (int j = ka; j < ka + 4; ++j)
                    {
                        const int bb = j & 3;
                        if (j - 4 >= 0) { ptx::mbar_wait(sa_free(w, bb), (sa_phase >> bb) & 1); sa_phase ^= 1u << bb; }
                    }
                    if (ka < 4 && !first_tile_of_warp) ptx::mbar_wait(o_empty(w), c_oempty++ & 1); // (tiny sequences: buffer 3 never used)
                    ptx::tcgen05_fence_after();
                }
                // sweep B: scores again, now followed by P_j V_j.  Buffer(j) holds P_{j-2} until P_{j-2} V_{j-2} has been issued (the tensor
                // pipe runs in issue order, so S_j may follow it immediately).
                bool first = true;
                for (int j = 0; j < nb + 2; ++j)
                {
                    const int bb = j & 1;
                    if (j >= 2)
                    {
                        ptx::mbar_wait(p_ready(w, bb), c_pready[bb]++ & 1);
                        if (!kv_seen) { ptx::mbar_wait(v_full, ip & 1); kv_seen = true; }
                        ptx::tcgen05_fence_after();
                        issue_pv(j - 2, first);
                        first = false;
                    }
                    if (j < nb) issue_s(j);
                    if (j == nb - 1 && ptx::elect_one())
                    {
                        ptx::tcgen05_commit(q_empty(w));                // every MMA reading this Q tile is issued
                        if (last_of_head) ptx::tcgen05_commit(k_empty); // ... and every MMA of this issuer reading K
                    }
                    __syncwarp();
                }
                if (ptx::elect_one())
                {
                    ptx::tcgen05_commit(o_full(w));
                    if (last_of_head) ptx::tcgen05_commit(v_empty);     // every MMA of this issuer reading V has been issued before this commit
                }
                __syncwarp();
                first_tile_of_warp = false;
            }
        }
    }
    else
    {
        // ===================== soft-max + epilogue warpgroup w: every second query tile, both sweeps over all key blocks =====================
        const int w = (warp_idx - 2) >> 2;
        const int q = warp_idx & 3;        // TMEM lane quarter
        const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16);
        const uint32_t t_o = t_lane + ocol(w);
        uint32_t c_sfull[2] = {0, 0}, c_ofull = 0, sa_phase = 0; // sa_phase: bit bb = parity of sa_full(w, bb)
        const uint64_t scale2 = ptx::pack_f32x2(p.scale, p.scale);
        int iq = 0, tile_seq = 0;
        for (int prob = blockIdx.x; prob < p.n_problems; prob += gridDim.x)
        {
            const int b = prob / p.H, h = prob - b * p.H;
            for (int ti = 0; ti < p.n_tiles; ++ti, ++iq)
            {
                if ((iq & 1) != w) continue;
                const bool warp_valid = (ti * 128 + q * 32) < p.N; // warp owns at least one real query row
                ATT_LONG_TRACE(0);
                // ---- sweep A: true row maximum over the valid keys (ggml.c:10533-10534)
                float mx = -INFINITY;
                for (int j = 0; j < ka; ++j)
                {
                    const int key0 = j * 64, kb = p.NKP - key0 < 64 ? p.NKP - key0 : 64;
                    const int bb = j & 3;
                    const uint32_t t_s = t_lane + sacol(w, bb);
                    ptx::mbar_wait(sa_full(w, bb), (sa_phase >> bb) & 1);
                    sa_phase ^= 1u << bb;
                    ptx::tcgen05_fence_after();
                    if (warp_valid)
                    {
                        const int valid = p.N - key0 < kb ? p.N - key0 : kb; // valid keys in this block (only the last block is short)
                        if (valid >= 64)
                        {
                            uint32_t va[32], vb[32];
                            ptx::tcgen05_ld_32x32b_x32(t_s, va);
                            ptx::tcgen05_ld_32x32b_x32(t_s + 32, vb);
                            ptx::tcgen05_wait_ld();
                            float ma = -INFINITY, mb = -INFINITY;
#pragma unroll
                            for (int x = 0; x < 32; ++x) { ma = fmaxf(ma, __uint_as_float(va[x])); mb = fmaxf(mb, __uint_as_float(vb[x])); }
                            mx = fmaxf(mx, fmaxf(ma, mb));
                        }
                        else
                        {
                            for (int col = 0; col < valid; col += 16) // 16 columns at a time so no column past the block is read
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
                    ptx::tcgen05_fence_before();
                    __syncwarp();
                    if (lane == 0) ptx::mbar_arrive(sa_free(w, bb));
                }
                ATT_LONG_TRACE(1);
                const float mxs = mx * p.scale; // ggml_scale_inplace (vit.cpp:851-854); exact, scale = 1/8
                const uint64_t nmax2 = ptx::pack_f32x2(-mxs, -mxs);

                // ---- sweep B: P = f16(exp(f16(s*scale - max))) over S in place, l = sum P
                float l0 = 0.f, l1 = 0.f, l2s = 0.f, l3 = 0.f;
                for (int j = 0; j < nb; ++j)
                {
                    const int key0 = p.key0[j], kb = p.key0[j + 1] - key0;
                    const int bb = j & 1;
                    const uint32_t t_s = t_lane + scol(w, bb);
                    ptx::mbar_wait(s_full(w, bb), c_sfull[bb]++ & 1);
                    ptx::tcgen05_fence_after();
                    if (warp_valid)
                    {
                        const int valid = p.N - key0 < kb ? p.N - key0 : kb;
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
                        // the granule straddling N (keys >= N get exactly zero) and the all-padding granules behind it
                        for (int g = nc32 * 2; g < ng; ++g)
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
                    if (lane == 0) ptx::mbar_arrive(p_ready(w, bb));
                }
                ATT_LONG_TRACE(3);
                const float lsum = (l0 + l1) + (l2s + l3);

                // ---- O = P V is complete: drain, release, normalise, store
                ptx::mbar_wait(o_full(w), c_ofull++ & 1);
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
                if (lane == 0) ptx::mbar_arrive(o_empty(w));
                if (warp_valid)
                {
                    const float inv = 1.0f / lsum; // p_i = e_i * (1/sum)  (ggml.c:10556-10558)
                    const uint32_t sbox = stage_out + (uint32_t)(w * 4 + q) * 4096u;
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
                ++tile_seq;
            }
        }
        if (lane == 0) ptx::tma_store_wait_all();
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
