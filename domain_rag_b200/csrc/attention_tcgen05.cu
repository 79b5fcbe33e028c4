// Non-causal attention for the Flux MMDiT blocks (head dim 128, joint text+image sequence) and the
// CLIP ViT blocks (head dim 64) on tcgen05:   out = softmax(q k^T / sqrt(hd)) v,  q,k,v bf16 [B][H][S][hd]
// (F.scaled_dot_product_attention inside diffusers' Flux attention processor, reached from pipe(...) at
//  batch_generate_flux_kshot.py:467-474 / outpainting_updown_sampling_redux.py:1246-1257; and
//  nn.MultiheadAttention inside clip.encode_image, retrieval/clip100_resnet_style_all_shots.py:171.)
//
// One CTA per (256-query block, batch*head) = two 128-row query tiles A and B that ping-pong on the
// tensor core, flash-attention style loop over 128-key tiles shared by both:
//   warp 0        TMA: Q_A, Q_B once; K/V tiles into 2-stage rings (3-D tensor maps, SWIZZLE_128B, rows
//                 past the sequence end zero-filled by TMA);
//   warp 1        MMA issuer (one thread): S_X = Q_X K_j^T (UMMA 128x128x16, SS) into TMEM, and
//                 O_X += P_X V_j with P read straight from TMEM (TS form) and V consumed MN-major from
//                 its row-major tile (one UMMA 128xHDx16 per 16 keys, the 64-wide halves chained by the descriptor LBO);
//   warps 4-7 / 8-11  softmax warpgroup of tile A / B: thread = query row. Pass 1 reads the score row
//                 from TMEM for the running max, lazy rescale of the TMEM-resident O (only when the max
//                 grew by > 8 in the exp2 domain), pass 2 re-reads the scores, exponentiates and writes
//                 P as packed bf16 IN PLACE over the first 64 columns of its own S buffer.
// While one warpgroup does softmax the tensor core works for the other tile. The tensor pipe executes
// MMAs in issue order, so S_X(j+1) = Q_X K_{j+1}^T may be issued right after O_X += P_X(j) V_j even
// though it overwrites P_X(j).
// TMEM columns: S_A/P_A [0,128)  S_B/P_B [128,256)  O_A [256,256+hd)  O_B [384,384+hd).
#include "attention_common.cuh"

namespace drag {

// PP (ping-pong) = two query tiles per CTA, one CTA per SM: the long-sequence Flux shape. !PP = one query tile per
// CTA, 256 threads, 256 TMEM columns and <= 32 K registers, so TWO CTAs share an SM: medium head-dim-64 sequences (258-512
// keys, e.g. ViT-L/14 at 336 px has 577 -> PP) are a latency chain per CTA (prologue, first TMA, n x [QK -> softmax -> PV],
// epilogue), and a second resident CTA overlaps part of it. Up to 257 keys never get here: attention_row.cu.
template <int HD, bool PP>
struct AttnCfg {
    static constexpr int NH = HD / 64;
    static constexpr int QT = PP ? 2 : 1;                          // query tiles per CTA
    static constexpr int THREADS = PP ? AT_THREADS : 256;
    static constexpr int TMEM_COLS = PP ? 512 : 256;
    static constexpr int O_COL = PP ? 256 : 128;                   // first accumulator column
    static constexpr int TILE_BYTES = NH * AT_HALF_BYTES;          // one Q / K / V tile
    static constexpr int Q_OFF = 0;                                // Q_A, Q_B
    static constexpr int K_OFF = QT * TILE_BYTES;                  // 2 stages
    static constexpr int V_OFF = K_OFF + 2 * TILE_BYTES;           // 2 stages
    static constexpr int BAR_OFF = V_OFF + 2 * TILE_BYTES;
    static constexpr int SMEM = BAR_OFF + 256 + 1024;
};

template <int HD, bool PP>
__global__ void __launch_bounds__(PP ? AT_THREADS : 256, PP ? 1 : 2)
attention_tcgen05_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                         const __grid_constant__ CUtensorMap tmV, AttnArgs a) {
    using Cfg = AttnCfg<HD, PP>;
    constexpr int TILE_BYTES = Cfg::TILE_BYTES;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::BAR_OFF);
    uint64_t* q_full = bars + 0;
    uint64_t* k_full = bars + 1;    // [2]
    uint64_t* k_empty = bars + 3;   // [2]
    uint64_t* v_full = bars + 5;    // [2]
    uint64_t* v_empty = bars + 7;   // [2]
    uint64_t* s_full = bars + 9;    // [2] per query tile
    uint64_t* p_full = bars + 11;   // [2]
    uint64_t* pv_done = bars + 13;  // [2]
    uint64_t* p_half = bars + 15;   // [2] keys [0,64) of P_x written (p_full: the whole tile)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 18);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * Cfg::QT * AT_TILE;
    const int bh = blockIdx.y;
    const int n_tiles = (a.S + AT_TILE - 1) / AT_TILE;
    const bool has_b = PP && (q0 + AT_TILE) < a.S;     // second query tile holds at least one row
    // Short sequences (!PP: CLIP ViT, 50 / 197 / 257 tokens) pay for every padded key: ViT-L/14's 257 = 2 x 128 + 1 made the
    // third key tile a full 128-column tile with ONE live column. There the last tile is narrow: Q K^T with N = the valid
    // keys rounded up to 16, exponentials for the valid keys rounded up to 32, P V over the same 16-key steps.
    const int tail = a.S - (n_tiles - 1) * AT_TILE;                     // valid keys of the last tile, 1..128
    const int tail_ks = PP ? AT_TILE / 16 : (tail + 15) / 16;           // 16-key MMA steps of the last tile
    const int tail_cols = PP ? AT_TILE : ((tail + 31) & ~31);           // score columns the softmax touches there

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmQ);
        tma_prefetch_desc(&tmK);
        tma_prefetch_desc(&tmV);
        mbar_init(q_full, 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&k_full[i], 1);
            mbar_init(&k_empty[i], 1);
            mbar_init(&v_full[i], 1);
            mbar_init(&v_empty[i], 1);
            mbar_init(&s_full[i], 1);
            mbar_init(&p_full[i], 4);
            mbar_init(&p_half[i], 4);
            mbar_init(&pv_done[i], 1);
        }
        fence_mbar_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // register re-balancing: the data-movement warpgroup gives registers to the two softmax warpgroups
    if (warp < 4) {
    setmaxnreg_dec<PP ? 96 : 56>();

    // Both loops run on all 32 lanes with warp-uniform control flow; only the TMA / MMA / commit instructions are
    // predicated on one elected lane, so shared-memory addresses and UMMA descriptors stay in uniform registers. A
    // single-lane issuer paid ~4 R2UR per MMA and, with 32-64-cycle MMAs, was the bottleneck of the whole kernel.
    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer
        if (elect_one()) {
            mbar_arrive_expect_tx(q_full, (has_b ? 2 : 1) * TILE_BYTES);
#pragma unroll
            for (int hf = 0; hf < Cfg::NH; ++hf) {
                tma_load_3d(smem + Cfg::Q_OFF + hf * AT_HALF_BYTES, &tmQ, hf * 64, q0, bh, q_full);
                if (has_b)
                    tma_load_3d(smem + Cfg::Q_OFF + TILE_BYTES + hf * AT_HALF_BYTES, &tmQ, hf * 64, q0 + AT_TILE, bh, q_full);
            }
        }
        __syncwarp();
        for (int j = 0; j < n_tiles; ++j) {
            const int st = j & 1, par = (j >> 1) & 1;
            uint8_t* kd = smem + Cfg::K_OFF + st * TILE_BYTES;
            uint8_t* vd = smem + Cfg::V_OFF + st * TILE_BYTES;
            mbar_wait(&k_empty[st], par ^ 1);
            if (elect_one()) {
                mbar_arrive_expect_tx(&k_full[st], TILE_BYTES);
#pragma unroll
                for (int hf = 0; hf < Cfg::NH; ++hf)
                    tma_load_3d(kd + hf * AT_HALF_BYTES, &tmK, hf * 64, j * AT_TILE, bh, &k_full[st]);
            }
            __syncwarp();
            mbar_wait(&v_empty[st], par ^ 1);
            if (elect_one()) {
                mbar_arrive_expect_tx(&v_full[st], TILE_BYTES);
#pragma unroll
                for (int hf = 0; hf < Cfg::NH; ++hf)
                    tma_load_3d(vd + hf * AT_HALF_BYTES, &tmV, hf * 64, j * AT_TILE, bh, &v_full[st]);
            }
            __syncwarp();
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer
        constexpr uint32_t idesc_qk = umma_idesc_bf16(128, 128, 0, 0);
        // O_x += P_x V: A = P from TMEM (K-major), B = V consumed MN-major from its row-major tile. One UMMA of N = HD
        // per 16 keys: the two 64-wide halves of V are consecutive MN atoms AT_HALF_BYTES apart (descriptor LBO).
        constexpr uint32_t idesc_pv = umma_idesc_bf16(128, HD, 0, 1);
        const uint64_t q_desc0 = umma_desc_k_sw128(smem_u32(smem + Cfg::Q_OFF));
        const uint64_t k_desc0 = umma_desc_k_sw128(smem_u32(smem + Cfg::K_OFF));
        const uint64_t v_desc0 = umma_desc_mn_sw128(smem_u32(smem + Cfg::V_OFF), Cfg::NH > 1 ? AT_HALF_BYTES : 0, 1024);
        const uint32_t idesc_qk_tail = umma_idesc_bf16(128, static_cast<uint32_t>(tail_ks) * 16, 0, 0);
        auto issue_qk = [&](int x, int st, bool last_tile) {     // S_x = Q_x K^T  (K tile already waited for)
            const uint64_t qd = q_desc0 + ((x * TILE_BYTES) >> 4), kd = k_desc0 + ((st * TILE_BYTES) >> 4);
            const uint32_t idesc = (!PP && last_tile) ? idesc_qk_tail : idesc_qk;
            if (elect_one()) {
#pragma unroll
                for (int ks = 0; ks < HD / 16; ++ks) {
                    const uint32_t off = ((ks >> 2) * AT_HALF_BYTES + (ks & 3) * 32) >> 4;
                    tc_mma_f16(tmem_base + x * 128, qd + off, kd + off, idesc, ks != 0);
                }
                tc_commit(&s_full[x]);
            }
            __syncwarp();
        };
        // O_x += P_x V, keys [half*64, half*64+64) of the tile   (P_x: packed bf16 in S_x columns [0,64))
        auto issue_pv = [&](int x, int st, int j, int half, bool last) {
            const uint64_t vd = v_desc0 + ((st * TILE_BYTES) >> 4);
            const int nks = (!PP && j == n_tiles - 1) ? tail_ks : AT_TILE / 16;     // 16-key steps of this tile
            if (elect_one()) {
#pragma unroll
                for (int kk = 0; kk < AT_TILE / 32; ++kk) {    // 16 kv rows per k-step = 2048 B inside each half
                    const int ks = half * (AT_TILE / 32) + kk;
                    if (PP || ks < nks)
                        tc_mma_f16_ts(tmem_base + Cfg::O_COL + x * 128, tmem_base + x * 128 + ks * 8, vd + ((ks * 2048) >> 4),
                                      idesc_pv, (j | ks) != 0);
                }
                if (last) tc_commit(&pv_done[x]);
            }
            __syncwarp();
        };
        mbar_wait(q_full, 0);
        mbar_wait(&k_full[0], 0);
        tc_fence_after();
        issue_qk(0, 0, n_tiles == 1);
        if (has_b) issue_qk(1, 0, n_tiles == 1);
        if (elect_one()) tc_commit(&k_empty[0]);
        __syncwarp();
        for (int j = 0; j < n_tiles; ++j) {
            const int st = j & 1, par = (j >> 1) & 1;
            const int nst = (j + 1) & 1, npar = ((j + 1) >> 1) & 1;
            const bool more = (j + 1) < n_tiles;
            mbar_wait(&v_full[st], par);
            mbar_wait(&p_half[0], j & 1);
            tc_fence_after();
            issue_pv(0, st, j, 0, false);
            mbar_wait(&p_full[0], j & 1);
            tc_fence_after();
            issue_pv(0, st, j, 1, true);
            if (more) {
                mbar_wait(&k_full[nst], npar);
                tc_fence_after();
                issue_qk(0, nst, j + 2 == n_tiles);
            }
            if (has_b) {
                mbar_wait(&p_half[1], j & 1);
                tc_fence_after();
                issue_pv(1, st, j, 0, false);
                mbar_wait(&p_full[1], j & 1);
                tc_fence_after();
                issue_pv(1, st, j, 1, true);
                if (more) issue_qk(1, nst, j + 2 == n_tiles);
            }
            if (elect_one()) {
                tc_commit(&v_empty[st]);
                if (more) tc_commit(&k_empty[nst]);
            }
            __syncwarp();
        }
    }
    } else {
    setmaxnreg_inc<200>();
    if (warp < 8 || has_b) {
        // ------------------------------------------------------------------ softmax + epilogue
        const int x = (warp >= 8) ? 1 : 0;                 // query tile of this warpgroup
        const int quarter = warp & 3;
        const int r = quarter * 32 + lane;                 // query row inside the tile
        const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
        const uint32_t t_s = t_lane + x * 128;             // S / P
        const uint32_t t_o = t_lane + Cfg::O_COL + x * 128;  // O
        float m = -INFINITY, l = 0.f;
        // !PP: a warp whose 32 query rows all lie past the sequence end (ViT-L/14: 127 of the 128 rows of the third query
        // tile) only keeps the barrier protocol going - its rows of P and O are never stored, so they may hold anything.
        // It paces itself on s_full so that its arrivals land in the right barrier phase.
        const bool warp_dead = !PP && (q0 + quarter * 32) >= a.S;
        if (warp_dead) {
            for (int j = 0; j < n_tiles; ++j) {
                mbar_wait(&s_full[x], j & 1);
                if (lane == 0) {
                    mbar_arrive(&p_half[x]);
                    mbar_arrive(&p_full[x]);
                }
            }
        }
        for (int j = 0; j < (warp_dead ? 0 : n_tiles); ++j) {
            mbar_wait(&s_full[x], j & 1);
            tc_fence_after();
            const int kv_valid = a.S - j * AT_TILE;        // columns >= kv_valid are past the sequence end
            const int ncol = (!PP && j == n_tiles - 1) ? tail_cols : AT_TILE;     // score columns in use (warp-uniform)
            // the whole 128-wide score row of this thread in registers: ONE TMEM pass per key tile
            uint32_t v[AT_TILE];
#pragma unroll
            for (int c = 0; c < AT_TILE; c += 32) {
                if (PP || c < ncol) tmem_ld_32x32_ptr(t_s + c, &v[c]);
            }
            tmem_ld_wait();
            if (kv_valid < AT_TILE) {
#pragma unroll
                for (int i = 0; i < AT_TILE; ++i)
                    if (i >= kv_valid) v[i] = 0xff800000u;   // -inf: exp2 -> 0, never the maximum
            }
            // four independent chains (the FMNMX3 latency chain was as long as the exponential pass)
            float m0 = fmaxf(__uint_as_float(v[0]), __uint_as_float(v[1]));
            float m1 = fmaxf(__uint_as_float(v[2]), __uint_as_float(v[3]));
            float m2 = fmaxf(__uint_as_float(v[4]), __uint_as_float(v[5]));
            float m3 = fmaxf(__uint_as_float(v[6]), __uint_as_float(v[7]));
#pragma unroll
            for (int i = 8; i < AT_TILE; i += 8) {
                m0 = fmax3(m0, __uint_as_float(v[i]), __uint_as_float(v[i + 1]));
                m1 = fmax3(m1, __uint_as_float(v[i + 2]), __uint_as_float(v[i + 3]));
                m2 = fmax3(m2, __uint_as_float(v[i + 4]), __uint_as_float(v[i + 5]));
                m3 = fmax3(m3, __uint_as_float(v[i + 6]), __uint_as_float(v[i + 7]));
            }
            m0 = fmaxf(m0, m2);
            m1 = fmaxf(m1, m3);
            const float m_new = fmaxf(m, fmaxf(m0, m1) * a.scale_log2);
            if (__any_sync(0xffffffffu, m_new > m + AT_RESCALE_THRESHOLD)) {
                const float alpha = ex2_approx(m - m_new);   // 0 on the first tile (m = -inf)
                if (j > 0) {
                    mbar_wait(&pv_done[x], (j - 1) & 1);     // O_x += P_x(j-1) V_{j-1} has landed
                    tc_fence_after();
#pragma unroll 1
                    for (int c = 0; c < HD / 32; ++c) {
                        uint32_t o[32];
                        tmem_ld_32x32(t_o + c * 32, o);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
                        tmem_st_32x32(t_o + c * 32, o);
                    }
                }
                l *= alpha;
                m = m_new;
            }
            const float2 sc2 = splat2(a.scale_log2), nm2 = splat2(-m);
            float2 sum_a = splat2(0.f), sum_b = splat2(0.f);
            // P as packed bf16 written IN PLACE over the first 64 columns of S. Packed fp32x2 arithmetic throughout; of
            // every eight pairs six take their exponentials from the MUFU pipe and two from a Cody-Waite + cubic
            // polynomial on the FMA pipe (exp2 is the co-bottleneck of the tensor core at head dim 128). The first half
            // of the row is published on its own barrier so that the P V product of keys [0,64) starts while the second
            // half is still being exponentiated.
#pragma unroll
            for (int c = 0; c < AT_TILE; c += 32) {
                if (!PP && c >= ncol) break;                 // narrow last tile: nothing but padding from here on
                uint32_t packed[16];
#pragma unroll
                for (int pr = 0; pr < 16; ++pr) {            // pairs of row elements
                    const float2 x = ffma2(make_float2(__uint_as_float(v[c + 2 * pr]), __uint_as_float(v[c + 2 * pr + 1])),
                                           sc2, nm2);
                    const float2 e = ((AT_POLY_MASK >> (pr & 7)) & 1) ? ex2_poly2(x)
                                                                      : make_float2(ex2_approx(x.x), ex2_approx(x.y));
                    if (pr & 1) sum_b = fadd2(sum_b, e);
                    else sum_a = fadd2(sum_a, e);
                    packed[pr] = pack_bf16x2(e);
                }
                tmem_st_32x16(t_s + (c >> 1), packed);
                if (c == 32) {                       // keys [0,64) of this tile are in TMEM
                    tmem_st_wait();
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&p_half[x]);
                }
            }
            const float2 sum2 = fadd2(sum_a, sum_b);
            const float sum0 = sum2.x, sum1 = sum2.y;
            l += sum0 + sum1;
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (!PP && ncol < 64) mbar_arrive(&p_half[x]);     // the c == 32 checkpoint above was never reached
                mbar_arrive(&p_full[x]);
            }
        }
        if (!warp_dead) {
        mbar_wait(&pv_done[x], (n_tiles - 1) & 1);
        tc_fence_after();
        const float inv = 1.f / l;
        const int srow = q0 + x * AT_TILE + r;
        const int b = bh / a.H, h = bh - b * a.H;
        __nv_bfloat16* orow = (srow < a.split)
            ? a.out0 + (static_cast<size_t>(b) * a.split + srow) * a.ld0 + h * HD
            : a.out1 + (static_cast<size_t>(b) * (a.S - a.split) + (srow - a.split)) * a.ld1 + h * HD;
#pragma unroll 1
        for (int c = 0; c < HD / 32; ++c) {
            uint32_t v[32];
            tmem_ld_32x32(t_o + c * 32, v);
            tmem_ld_wait();
            if (srow < a.S) {
#pragma unroll
                for (int i = 0; i < 32; i += 8) {
                    __nv_bfloat162 p0 = __floats2bfloat162_rn(__uint_as_float(v[i]) * inv, __uint_as_float(v[i + 1]) * inv);
                    __nv_bfloat162 p1 = __floats2bfloat162_rn(__uint_as_float(v[i + 2]) * inv, __uint_as_float(v[i + 3]) * inv);
                    __nv_bfloat162 p2 = __floats2bfloat162_rn(__uint_as_float(v[i + 4]) * inv, __uint_as_float(v[i + 5]) * inv);
                    __nv_bfloat162 p3 = __floats2bfloat162_rn(__uint_as_float(v[i + 6]) * inv, __uint_as_float(v[i + 7]) * inv);
                    uint4 u;
                    u.x = *reinterpret_cast<uint32_t*>(&p0);
                    u.y = *reinterpret_cast<uint32_t*>(&p1);
                    u.z = *reinterpret_cast<uint32_t*>(&p2);
                    u.w = *reinterpret_cast<uint32_t*>(&p3);
                    *reinterpret_cast<uint4*>(orow + c * 32 + i) = u;
                }
            }
        }
        }
    }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
    }
}

// ---------------------------------------------------------------------------------------------- split-row variant
// Same pipeline as the ping-pong kernel (two query tiles per CTA, S / P / O in TMEM), but every query tile has TWO softmax
// warpgroups: thread (row r, half hf) owns score columns [64 hf, 64 hf + 64) of its row. In the one-thread-per-row kernel the
// S -> P latency of a tile (TMEM load of 128 columns, max chain, 128 exponentials, pack, store: ~1400 cycles on a single
// warp per sub-partition) is longer than the ~1024 cycles the tensor core needs for the OTHER tile's two MMAs, so the two
// chains alternate instead of overlapping and the tensor pipe idles 28 % of the time. Two warps per sub-partition halve the
// per-thread instruction stream and keep the MUFU pipe fed from two instruction streams. The row maximum is exchanged
// through shared memory (one 256-thread named barrier per key tile, which also orders the partner's S reads before this
// thread's in-place P writes); keys [0,64) are exactly half 0's work, so `p_half` / `p_full` are simply "half 0 done" /
// "half 1 done" and the MMA issuer is unchanged. O is rescaled and stored in halves. 20 warps: {TMA, MMA, 2 idle} + 4 x 4.
// MEASURED (profiles/r02_attention_split_ab.txt): correct (same parity bar) but 13 % slower than one thread per row - 1115 vs
// 1285 TFLOP/s at B = 4, S = 5337: the exchange barrier couples the two halves and the 104-register budget spills. Kept
// behind drag_debug_set key 7 as a documented negative result; the default path is the one-thread-per-row kernel.
constexpr int AT_SPLIT_THREADS = 640;
template <int HD>
// Register re-balancing: setmaxnreg.inc can only take what setmaxnreg.dec of the same CTA has released (registers of the SM
// that were never allocated to the CTA do not count; a request beyond the released pool waits forever - the first two
// versions of this kernel hung there). 640 threads launch with 96 registers; the data-movement warpgroup drops to 64
// (releases 128 x 32 = 4096) and the 512 softmax threads rise to 104 (take 512 x 8 = 4096).
__global__ void __launch_bounds__(AT_SPLIT_THREADS, 1)
attention_tcgen05_split_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                         const __grid_constant__ CUtensorMap tmV, AttnArgs a) {
    constexpr bool PP = true;
    using Cfg = AttnCfg<HD, true>;
    constexpr int TILE_BYTES = Cfg::TILE_BYTES;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::BAR_OFF);
    uint64_t* q_full = bars + 0;
    uint64_t* k_full = bars + 1;    // [2]
    uint64_t* k_empty = bars + 3;   // [2]
    uint64_t* v_full = bars + 5;    // [2]
    uint64_t* v_empty = bars + 7;   // [2]
    uint64_t* s_full = bars + 9;    // [2] per query tile
    uint64_t* p_full = bars + 11;   // [2]
    uint64_t* pv_done = bars + 13;  // [2]
    uint64_t* p_half = bars + 15;   // [2] keys [0,64) of P_x written (p_full: the whole tile)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 18);
    float* mx = reinterpret_cast<float*>(smem + Cfg::BAR_OFF + 256);   // [2 parity][2 tiles][2 halves][128 rows] row maxima
    float* lx = mx + 2 * 2 * 2 * 128;                                  // [2 tiles][2 halves][128 rows] row sums (epilogue)

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * Cfg::QT * AT_TILE;
    const int bh = blockIdx.y;
    const int n_tiles = (a.S + AT_TILE - 1) / AT_TILE;
    const bool has_b = PP && (q0 + AT_TILE) < a.S;     // second query tile holds at least one row

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmQ);
        tma_prefetch_desc(&tmK);
        tma_prefetch_desc(&tmV);
        mbar_init(q_full, 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&k_full[i], 1);
            mbar_init(&k_empty[i], 1);
            mbar_init(&v_full[i], 1);
            mbar_init(&v_empty[i], 1);
            mbar_init(&s_full[i], 1);
            mbar_init(&p_full[i], 4);
            mbar_init(&p_half[i], 4);
            mbar_init(&pv_done[i], 1);
        }
        fence_mbar_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // register re-balancing: the data-movement warpgroup gives registers to the two softmax warpgroups
    if (warp < 4) {
    setmaxnreg_dec<64>();

    // Both loops run on all 32 lanes with warp-uniform control flow; only the TMA / MMA / commit instructions are
    // predicated on one elected lane, so shared-memory addresses and UMMA descriptors stay in uniform registers. A
    // single-lane issuer paid ~4 R2UR per MMA and, with 32-64-cycle MMAs, was the bottleneck of the whole kernel.
    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer
        if (elect_one()) {
            mbar_arrive_expect_tx(q_full, (has_b ? 2 : 1) * TILE_BYTES);
#pragma unroll
            for (int hf = 0; hf < Cfg::NH; ++hf) {
                tma_load_3d(smem + Cfg::Q_OFF + hf * AT_HALF_BYTES, &tmQ, hf * 64, q0, bh, q_full);
                if (has_b)
                    tma_load_3d(smem + Cfg::Q_OFF + TILE_BYTES + hf * AT_HALF_BYTES, &tmQ, hf * 64, q0 + AT_TILE, bh, q_full);
            }
        }
        __syncwarp();
        for (int j = 0; j < n_tiles; ++j) {
            const int st = j & 1, par = (j >> 1) & 1;
            uint8_t* kd = smem + Cfg::K_OFF + st * TILE_BYTES;
            uint8_t* vd = smem + Cfg::V_OFF + st * TILE_BYTES;
            mbar_wait(&k_empty[st], par ^ 1);
            if (elect_one()) {
                mbar_arrive_expect_tx(&k_full[st], TILE_BYTES);
#pragma unroll
                for (int hf = 0; hf < Cfg::NH; ++hf)
                    tma_load_3d(kd + hf * AT_HALF_BYTES, &tmK, hf * 64, j * AT_TILE, bh, &k_full[st]);
            }
            __syncwarp();
            mbar_wait(&v_empty[st], par ^ 1);
            if (elect_one()) {
                mbar_arrive_expect_tx(&v_full[st], TILE_BYTES);
#pragma unroll
                for (int hf = 0; hf < Cfg::NH; ++hf)
                    tma_load_3d(vd + hf * AT_HALF_BYTES, &tmV, hf * 64, j * AT_TILE, bh, &v_full[st]);
            }
            __syncwarp();
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer
        constexpr uint32_t idesc_qk = umma_idesc_bf16(128, 128, 0, 0);
        // O_x += P_x V: A = P from TMEM (K-major), B = V consumed MN-major from its row-major tile. One UMMA of N = HD
        // per 16 keys: the two 64-wide halves of V are consecutive MN atoms AT_HALF_BYTES apart (descriptor LBO).
        constexpr uint32_t idesc_pv = umma_idesc_bf16(128, HD, 0, 1);
        const uint64_t q_desc0 = umma_desc_k_sw128(smem_u32(smem + Cfg::Q_OFF));
        const uint64_t k_desc0 = umma_desc_k_sw128(smem_u32(smem + Cfg::K_OFF));
        const uint64_t v_desc0 = umma_desc_mn_sw128(smem_u32(smem + Cfg::V_OFF), Cfg::NH > 1 ? AT_HALF_BYTES : 0, 1024);
        auto issue_qk = [&](int x, int st) {     // S_x = Q_x K^T  (K tile already waited for)
            const uint64_t qd = q_desc0 + ((x * TILE_BYTES) >> 4), kd = k_desc0 + ((st * TILE_BYTES) >> 4);
            if (elect_one()) {
#pragma unroll
                for (int ks = 0; ks < HD / 16; ++ks) {
                    const uint32_t off = ((ks >> 2) * AT_HALF_BYTES + (ks & 3) * 32) >> 4;
                    tc_mma_f16(tmem_base + x * 128, qd + off, kd + off, idesc_qk, ks != 0);
                }
                tc_commit(&s_full[x]);
            }
            __syncwarp();
        };
        // O_x += P_x V, keys [half*64, half*64+64) of the tile   (P_x: packed bf16 in S_x columns [0,64))
        auto issue_pv = [&](int x, int st, int j, int half, bool last) {
            const uint64_t vd = v_desc0 + ((st * TILE_BYTES) >> 4);
            if (elect_one()) {
#pragma unroll
                for (int kk = 0; kk < AT_TILE / 32; ++kk) {    // 16 kv rows per k-step = 2048 B inside each half
                    const int ks = half * (AT_TILE / 32) + kk;
                    tc_mma_f16_ts(tmem_base + Cfg::O_COL + x * 128, tmem_base + x * 128 + ks * 8, vd + ((ks * 2048) >> 4), idesc_pv,
                                  (j | ks) != 0);
                }
                if (last) tc_commit(&pv_done[x]);
            }
            __syncwarp();
        };
        mbar_wait(q_full, 0);
        mbar_wait(&k_full[0], 0);
        tc_fence_after();
        issue_qk(0, 0);
        if (has_b) issue_qk(1, 0);
        if (elect_one()) tc_commit(&k_empty[0]);
        __syncwarp();
        for (int j = 0; j < n_tiles; ++j) {
            const int st = j & 1, par = (j >> 1) & 1;
            const int nst = (j + 1) & 1, npar = ((j + 1) >> 1) & 1;
            const bool more = (j + 1) < n_tiles;
            mbar_wait(&v_full[st], par);
            mbar_wait(&p_half[0], j & 1);
            tc_fence_after();
            issue_pv(0, st, j, 0, false);
            mbar_wait(&p_full[0], j & 1);
            tc_fence_after();
            issue_pv(0, st, j, 1, true);
            if (more) {
                mbar_wait(&k_full[nst], npar);
                tc_fence_after();
                issue_qk(0, nst);
            }
            if (has_b) {
                mbar_wait(&p_half[1], j & 1);
                tc_fence_after();
                issue_pv(1, st, j, 0, false);
                mbar_wait(&p_full[1], j & 1);
                tc_fence_after();
                issue_pv(1, st, j, 1, true);
                if (more) issue_qk(1, nst);
            }
            if (elect_one()) {
                tc_commit(&v_empty[st]);
                if (more) tc_commit(&k_empty[nst]);
            }
            __syncwarp();
        }
    }
    } else {
    setmaxnreg_inc<104>();
    const int wg = (warp - 4) >> 2;                        // 0..3
    const int x = wg >> 1, hf = wg & 1;                    // query tile, column half
    if (x == 0 || has_b) {
        // ------------------------------------------------------------------ softmax + epilogue (half a row per thread)
        const int quarter = warp & 3;
        const int r = quarter * 32 + lane;                 // query row inside the tile
        const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
        const uint32_t t_s = t_lane + x * 128;             // S / P of the tile
        const uint32_t t_o = t_lane + Cfg::O_COL + x * 128 + hf * (HD / 2);  // this thread's half of O
        uint64_t* p_done = hf ? &p_full[x] : &p_half[x];
        float m = -INFINITY, l = 0.f;
        for (int j = 0; j < n_tiles; ++j) {
            mbar_wait(&s_full[x], j & 1);
            tc_fence_after();
            const int kv_valid = a.S - j * AT_TILE - hf * 64;   // columns >= kv_valid of THIS half are past the sequence end
            uint32_t v[64];
            tmem_ld_32x32_ptr(t_s + hf * 64, &v[0]);
            tmem_ld_32x32_ptr(t_s + hf * 64 + 32, &v[32]);
            tmem_ld_wait();
            if (kv_valid < 64) {
#pragma unroll
                for (int i = 0; i < 64; ++i)
                    if (i >= kv_valid) v[i] = 0xff800000u;   // -inf: exp2 -> 0, never the maximum
            }
            float m0 = fmaxf(__uint_as_float(v[0]), __uint_as_float(v[1]));
            float m1 = fmaxf(__uint_as_float(v[2]), __uint_as_float(v[3]));
            float m2 = fmaxf(__uint_as_float(v[4]), __uint_as_float(v[5]));
            float m3 = fmaxf(__uint_as_float(v[6]), __uint_as_float(v[7]));
#pragma unroll
            for (int i = 8; i < 64; i += 8) {
                m0 = fmax3(m0, __uint_as_float(v[i]), __uint_as_float(v[i + 1]));
                m1 = fmax3(m1, __uint_as_float(v[i + 2]), __uint_as_float(v[i + 3]));
                m2 = fmax3(m2, __uint_as_float(v[i + 4]), __uint_as_float(v[i + 5]));
                m3 = fmax3(m3, __uint_as_float(v[i + 6]), __uint_as_float(v[i + 7]));
            }
            const float m_loc = fmaxf(fmaxf(m0, m2), fmaxf(m1, m3));
            // exchange with the thread that owns the other half of this row. The barrier also guarantees that the partner
            // has finished reading its S columns before this thread overwrites them with P.
            float* mrow = mx + (((j & 1) * 2 + x) * 2) * 128;
            mrow[hf * 128 + r] = m_loc;
            asm volatile("bar.sync %0, 256;" ::"r"(1 + x) : "memory");
            const float m_new = fmaxf(m, fmaxf(m_loc, mrow[(hf ^ 1) * 128 + r]) * a.scale_log2);
            if (__any_sync(0xffffffffu, m_new > m + AT_RESCALE_THRESHOLD)) {
                const float alpha = ex2_approx(m - m_new);   // 0 on the first tile (m = -inf)
                if (j > 0) {
                    mbar_wait(&pv_done[x], (j - 1) & 1);     // O_x += P_x(j-1) V_{j-1} has landed
                    tc_fence_after();
#pragma unroll 1
                    for (int c = 0; c < HD / 64; ++c) {
                        uint32_t o[32];
                        tmem_ld_32x32(t_o + c * 32, o);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
                        tmem_st_32x32(t_o + c * 32, o);
                    }
                }
                l *= alpha;
                m = m_new;
            }
            const float2 sc2 = splat2(a.scale_log2), nm2 = splat2(-m);
            float2 sum_a = splat2(0.f), sum_b = splat2(0.f);
#pragma unroll
            for (int c = 0; c < 64; c += 32) {
                uint32_t packed[16];
#pragma unroll
                for (int pr = 0; pr < 16; ++pr) {            // pairs of row elements
                    const float2 xx = ffma2(make_float2(__uint_as_float(v[c + 2 * pr]), __uint_as_float(v[c + 2 * pr + 1])),
                                            sc2, nm2);
                    const float2 e = ((AT_POLY_MASK >> (pr & 7)) & 1) ? ex2_poly2(xx)
                                                                      : make_float2(ex2_approx(xx.x), ex2_approx(xx.y));
                    if (pr & 1) sum_b = fadd2(sum_b, e);
                    else sum_a = fadd2(sum_a, e);
                    packed[pr] = pack_bf16x2(e);
                }
                tmem_st_32x16(t_s + ((hf * 64 + c) >> 1), packed);    // P: packed bf16, keys [64 hf + c, +32)
            }
            const float2 sum2 = fadd2(sum_a, sum_b);
            l += sum2.x + sum2.y;
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(p_done);            // half 0 -> p_half (keys [0,64)), half 1 -> p_full
        }
        // total row sum = both halves
        lx[(x * 2 + hf) * 128 + r] = l;
        asm volatile("bar.sync %0, 256;" ::"r"(1 + x) : "memory");
        l += lx[(x * 2 + (hf ^ 1)) * 128 + r];
        mbar_wait(&pv_done[x], (n_tiles - 1) & 1);
        tc_fence_after();
        const float inv = 1.f / l;
        const int srow = q0 + x * AT_TILE + r;
        const int b = bh / a.H, h = bh - b * a.H;
        __nv_bfloat16* orow = (srow < a.split)
            ? a.out0 + (static_cast<size_t>(b) * a.split + srow) * a.ld0 + h * HD
            : a.out1 + (static_cast<size_t>(b) * (a.S - a.split) + (srow - a.split)) * a.ld1 + h * HD;
        orow += hf * (HD / 2);
#pragma unroll 1
        for (int c = 0; c < HD / 64; ++c) {
            uint32_t v[32];
            tmem_ld_32x32(t_o + c * 32, v);
            tmem_ld_wait();
            if (srow < a.S) {
#pragma unroll
                for (int i = 0; i < 32; i += 8) {
                    __nv_bfloat162 p0 = __floats2bfloat162_rn(__uint_as_float(v[i]) * inv, __uint_as_float(v[i + 1]) * inv);
                    __nv_bfloat162 p1 = __floats2bfloat162_rn(__uint_as_float(v[i + 2]) * inv, __uint_as_float(v[i + 3]) * inv);
                    __nv_bfloat162 p2 = __floats2bfloat162_rn(__uint_as_float(v[i + 4]) * inv, __uint_as_float(v[i + 5]) * inv);
                    __nv_bfloat162 p3 = __floats2bfloat162_rn(__uint_as_float(v[i + 6]) * inv, __uint_as_float(v[i + 7]) * inv);
                    uint4 u;
                    u.x = *reinterpret_cast<uint32_t*>(&p0);
                    u.y = *reinterpret_cast<uint32_t*>(&p1);
                    u.z = *reinterpret_cast<uint32_t*>(&p2);
                    u.w = *reinterpret_cast<uint32_t*>(&p3);
                    *reinterpret_cast<uint4*>(orow + c * 32 + i) = u;
                }
            }
        }
    }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
    }
}

template <int HD, bool PP>
static int launch_attention(const __nv_bfloat16* q, const __nv_bfloat16* k, const __nv_bfloat16* v, int B, int H, int S,
                            AttnArgs a, cudaStream_t st) {
    using Cfg = AttnCfg<HD, PP>;
    CUtensorMap tq, tk, tv;
    const uint64_t bh = static_cast<uint64_t>(B) * H;
    int rc = make_tmap_bf16_3d(&tq, q, HD, S, bh, HD, static_cast<uint64_t>(S) * HD, 64, AT_TILE);
    if (rc) return rc;
    rc = make_tmap_bf16_3d(&tk, k, HD, S, bh, HD, static_cast<uint64_t>(S) * HD, 64, AT_TILE);
    if (rc) return rc;
    rc = make_tmap_bf16_3d(&tv, v, HD, S, bh, HD, static_cast<uint64_t>(S) * HD, 64, AT_TILE);
    if (rc) return rc;
    static bool attr_set = false;
    if (!attr_set) {
        DRAG_CUDA(cudaFuncSetAttribute(attention_tcgen05_kernel<HD, PP>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       Cfg::SMEM));
        if (!PP)   // two CTAs per SM need 2 x SMEM of shared memory: ask for the largest carve-out
            DRAG_CUDA(cudaFuncSetAttribute(attention_tcgen05_kernel<HD, PP>,
                                           cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        attr_set = true;
    }
    a.scale_log2 = 1.4426950408889634f / sqrtf(static_cast<float>(HD));
    dim3 grid((S + Cfg::QT * AT_TILE - 1) / (Cfg::QT * AT_TILE), static_cast<unsigned>(bh));
    const int slot = prof_begin(PROF_ATTENTION, 4.0 * B * H * static_cast<double>(S) * S * HD, st);
    attention_tcgen05_kernel<HD, PP><<<grid, Cfg::THREADS, Cfg::SMEM, st>>>(tq, tk, tv, a); count_launch();
    prof_end(slot, st);
    DRAG_CUDA(cudaGetLastError());
    return DRAG_OK;
}

template <int HD>
static int launch_attention_split(const __nv_bfloat16* q, const __nv_bfloat16* k, const __nv_bfloat16* v, int B, int H, int S,
                                  AttnArgs a, cudaStream_t st) {
    using Cfg = AttnCfg<HD, true>;
    constexpr int SMEM = Cfg::SMEM + (2 * 2 * 2 + 2 * 2) * 128 * 4;       // + row-maximum / row-sum exchange
    CUtensorMap tq, tk, tv;
    const uint64_t bh = static_cast<uint64_t>(B) * H;
    int rc = make_tmap_bf16_3d(&tq, q, HD, S, bh, HD, static_cast<uint64_t>(S) * HD, 64, AT_TILE);
    if (rc) return rc;
    rc = make_tmap_bf16_3d(&tk, k, HD, S, bh, HD, static_cast<uint64_t>(S) * HD, 64, AT_TILE);
    if (rc) return rc;
    rc = make_tmap_bf16_3d(&tv, v, HD, S, bh, HD, static_cast<uint64_t>(S) * HD, 64, AT_TILE);
    if (rc) return rc;
    static bool attr_set = false;
    if (!attr_set) {
        DRAG_CUDA(cudaFuncSetAttribute(attention_tcgen05_split_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
        attr_set = true;
    }
    a.scale_log2 = 1.4426950408889634f / sqrtf(static_cast<float>(HD));
    dim3 grid((S + 2 * AT_TILE - 1) / (2 * AT_TILE), static_cast<unsigned>(bh));
    const int slot = prof_begin(PROF_ATTENTION, 4.0 * B * H * static_cast<double>(S) * S * HD, st);
    attention_tcgen05_split_kernel<HD><<<grid, AT_SPLIT_THREADS, SMEM, st>>>(tq, tk, tv, a); count_launch();
    prof_end(slot, st);
    DRAG_CUDA(cudaGetLastError());
    return DRAG_OK;
}

// Debug knobs kept for ABI stability (drag_debug_set keys 1/2); unused by the current kernel.
uint32_t g_attn_v_lbo = 0, g_attn_v_sbo = 1024;
int g_attn_force_pp = 0;      // drag_debug_set key 5: 1 = always the two-tile ping-pong kernel (A/B comparisons)
int g_attn_no_row = 0;        // drag_debug_set key 10: 1 = head dim 64 never takes the whole-row kernel (A/B comparisons)
int g_attn_split = 0;         // drag_debug_set key 7: head dim 128: 1 = split-row kernel (two softmax warpgroups per query tile),
                              // 0 = one thread per row (default: the split kernel measured 13 % SLOWER, see the kernel)

int attention_bf16(const __nv_bfloat16* q, const __nv_bfloat16* k, const __nv_bfloat16* v, int B, int H, int S,
                   int head_dim, int split, __nv_bfloat16* out0, int ld0, __nv_bfloat16* out1, int ld1,
                   cudaStream_t st) {
    DRAG_REQUIRE(q && k && v, "attention: null pointer");
    DRAG_REQUIRE(head_dim == 64 || head_dim == 128, "attention: head_dim must be 64 or 128");
    DRAG_REQUIRE(B >= 1 && H >= 1 && S >= 1 && split >= 0 && split <= S, "attention: bad sizes");
    DRAG_REQUIRE((split == 0 || out0) && (split == S || out1), "attention: null output");
    DRAG_REQUIRE(ld0 % 8 == 0 && ld1 % 8 == 0, "attention: output leading dims must be multiples of 8");
    AttnArgs a;
    a.out0 = out0; a.out1 = out1; a.ld0 = ld0; a.ld1 = ld1; a.split = split;
    a.S = S;
    a.H = H;
    a.scale_log2 = 0.f;
    if (head_dim == 128)
        return g_attn_split ? launch_attention_split<128>(q, k, v, B, H, S, a, st)
                            : launch_attention<128, true>(q, k, v, B, H, S, a, st);
    // head dim 64 = the CLIP ViT towers: up to 256 (+ AR_TAIL) keys -> the whole-row kernel; longer sequences the tiled ones
    // (two single-tile CTAs per SM up to 512 keys, the ping-pong beyond)
    if (attention_row_eligible(head_dim, S) && !g_attn_force_pp && !g_attn_no_row)
        return launch_attention_row_any(q, k, v, B, H, S, a, st);
    if (S <= 4 * AT_TILE && !g_attn_force_pp) return launch_attention<64, false>(q, k, v, B, H, S, a, st);
    return launch_attention<64, true>(q, k, v, B, H, S, a, st);
}

}  // namespace drag
