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
#include <cuda.h>

#include "common.cuh"
#include "flux_ops.cuh"
#include "prof.cuh"
#include "ptx.cuh"

namespace drag {

constexpr int AT_THREADS = 384;                        // 12 warps (warps 2,3 idle: keeps the softmax
                                                       // warpgroups aligned to TMEM lane quarters)
constexpr int AT_TILE = 128;
constexpr int AT_HALF_BYTES = AT_TILE * 64 * 2;        // 16 KB: 128 rows x 64 bf16
constexpr float AT_RESCALE_THRESHOLD = 8.0f;           // log2 domain
// Which of every 8 element pairs take exp2 from the FMA-pipe polynomial instead of MUFU.EX2 (bit i = pair i). MUFU
// issues 4 lanes / clock / sub-partition: 128 exponentials per row tile would keep the XU pipe busy for as long as the
// tensor core needs for the tile's two MMAs. Measured (S = 5337, batch 4): 2 of 8 -> 1324 TFLOP/s, 3 of 8 -> 1289.
constexpr uint32_t AT_POLY_MASK = 0x88;                // pairs 3, 7

// PP (ping-pong) = two query tiles per CTA, one CTA per SM: the long-sequence Flux shape. !PP = one query tile per
// CTA, 256 threads, 256 TMEM columns and <= 32 K registers, so TWO CTAs share an SM: short sequences (CLIP ViT, 50-257
// tokens = 1-3 key tiles) are a latency chain per CTA (prologue, first TMA, 3 x [QK -> softmax -> PV], epilogue), and
// a second resident CTA overlaps all of it, including launch and drain.
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

struct AttnArgs {
    __nv_bfloat16* out0;   // tokens [0, split): row b*split + s, leading dim ld0
    __nv_bfloat16* out1;   // tokens [split, S): row b*(S-split) + s-split, leading dim ld1
    int ld0, ld1, split;
    int S, H;
    float scale_log2;      // log2(e) / sqrt(head_dim)
};

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// 2^x for x <= ~8 on the FMA/ALU pipes: x = n + f with n = round(x), f in [-0.5, 0.5];
// 2^f by a cubic with max relative error 1.0e-4 (bf16 P has 2^-9), 2^n by adding n to the exponent field.
__device__ __forceinline__ float ex2_poly(float x) {
    x = fmaxf(x, -125.f);
    const float t = x + 12582912.f;                    // 1.5 * 2^23: the low mantissa bits of t hold round(x)
    const float f = x - (t - 12582912.f);
    float p = fmaf(0.05500871315598488f, f, 0.24221068620681763f);
    p = fmaf(p, f, 0.6932829022407532f);
    p = fmaf(p, f, 1.0f);
    return __uint_as_float(__float_as_uint(p) + (__float_as_uint(t) << 23));
}
__device__ __forceinline__ float fmax3(float a, float b, float c) {
    float d;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}
// Packed fp32x2 arithmetic (FFMA2 / FADD2 on sm_100): one issue slot for two row elements. The softmax warps are
// issue-bound (two of them share every SM sub-partition while the tensor core waits for P), so halving the FMA /
// ADD instruction count shortens the S -> P latency directly.
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ float2 fadd2(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 splat2(float x) { return make_float2(x, x); }
// ex2_poly on a pair.
__device__ __forceinline__ float2 ex2_poly2(float2 x) {
    x.x = fmaxf(x.x, -125.f);
    x.y = fmaxf(x.y, -125.f);
    const float2 t = fadd2(x, splat2(12582912.f));
    const float2 r = fadd2(t, splat2(-12582912.f));            // round(x)
    const float2 f = ffma2(r, splat2(-1.f), x);                 // x - round(x) in [-0.5, 0.5]
    float2 p = ffma2(splat2(0.05500871315598488f), f, splat2(0.24221068620681763f));
    p = ffma2(p, f, splat2(0.6932829022407532f));
    p = ffma2(p, f, splat2(1.0f));
    p.x = __uint_as_float(__float_as_uint(p.x) + (__float_as_uint(t.x) << 23));
    p.y = __uint_as_float(__float_as_uint(p.y) + (__float_as_uint(t.y) << 23));
    return p;
}
__device__ __forceinline__ uint32_t pack_bf16x2(float2 p) {
    __nv_bfloat162 pk = __floats2bfloat162_rn(p.x, p.y);
    return *reinterpret_cast<uint32_t*>(&pk);
}
// Register re-balancing between warpgroups (all four warps of a warpgroup execute it).
template <int N>
__device__ __forceinline__ void setmaxnreg_inc() {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N));
}
template <int N>
__device__ __forceinline__ void setmaxnreg_dec() {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N));
}

// registers -> TMEM, 32 lanes x 16 columns
__device__ __forceinline__ void tmem_st_32x16(uint32_t taddr, const uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
        "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
        "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
        : "memory");
}

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

// ---------------------------------------------------------------------------------------------- whole-row kernel (head dim 64)
// CLIP ViT sequences are 50 / 197 / 257 tokens. The tiled kernels above walk 128-key tiles with an online softmax: per
// query tile three QK -> softmax -> PV round trips through mbarriers plus prologue and epilogue - a LATENCY chain of ~13 k
// cycles per CTA in which the tensor core is busy 25 % of the time, and ViT-L/14's 257 = 2 x 128 + 1 pays a whole third key
// tile for one key (measured: trimming that tile's arithmetic alone changed nothing - the chain, not the math, is the cost).
// Here the whole score row lives in TMEM at once: ONE UMMA sequence S = Q K^T with N = up to 256 keys, a two-pass softmax
// (exact row maximum, no running rescale), ONE sequence O = P V, one barrier round trip per query tile. Key 256 (AR_TAIL = 1:
// the 257th token of ViT-L/14) never touch the tensor core: their scores are 64-term dot products on the
// CUDA cores, their P V contribution is added to O in the epilogue in fp32.
// TMEM (256 columns, two CTAs per SM): S [0,256) -> P packed bf16 [0,128) in place; O [128,192) reuses dead S columns.
constexpr int AR_MAIN = 256;
constexpr int AR_TAIL = 1;
struct AttnRowCfg {
    static constexpr int Q_OFF = 0;                           // 128 rows x 128 B
    static constexpr int K_OFF = AT_HALF_BYTES;               // 256 rows x 128 B (two TMA boxes back to back)
    static constexpr int V_OFF = K_OFF + 2 * AT_HALF_BYTES;
    static constexpr int BAR_OFF = V_OFF + 2 * AT_HALF_BYTES;
    static constexpr int KT_OFF = BAR_OFF + 256;              // tail keys: AR_TAIL rows x 128 B, unswizzled
    static constexpr int VT_OFF = KT_OFF + AR_TAIL * 128;
    static constexpr int SMEM = VT_OFF + AR_TAIL * 128 + 1024;
    static constexpr int TMEM_COLS = 256;
    static constexpr int O_COL = 128;
};
struct AttnRowArgs {
    AttnArgs a;
    const __nv_bfloat16 *k, *v;        // [B*H][S][64]: rows 256.. are bulk-copied next to the TMA tiles
    int prefetch_stride;               // this CTA warms L2 for the CTA `prefetch_stride` positions later in launch order
    int x_first;                       // first query tile of this launch (the persistent kernel leaves only tile 2 to this one)
};

__device__ __forceinline__ void bf16x8_to_float(const uint4& u, float* f) {
    const __nv_bfloat162* p = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 t = __bfloat1622float2(p[i]);
        f[2 * i] = t.x;
        f[2 * i + 1] = t.y;
    }
}

__global__ void __launch_bounds__(256, 2)
attention_row_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                     const __grid_constant__ CUtensorMap tmV, AttnRowArgs ar) {
    using Cfg = AttnRowCfg;
    constexpr int HD = 64;
    const AttnArgs& a = ar.a;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::BAR_OFF);
    uint64_t* qk_full = bars + 0;
    uint64_t* v_full = bars + 1;
    uint64_t* s_full = bars + 2;
    uint64_t* p_full = bars + 3;
    uint64_t* pv_done = bars + 4;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 6);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q0 = (blockIdx.x + ar.x_first) * AT_TILE;
    const int bh = blockIdx.y;
    const int n_main = min(a.S, AR_MAIN);                 // keys on the tensor core
    const int n_tail = a.S - n_main;                      // keys on the CUDA cores (0..AR_TAIL)
    const int n_mma = (n_main + 15) & ~15;                // N of Q K^T, K extent of P V
    const int k_boxes = (n_main + AT_TILE - 1) / AT_TILE;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmQ);
        tma_prefetch_desc(&tmK);
        tma_prefetch_desc(&tmV);
        mbar_init(qk_full, 1);
        mbar_init(v_full, 1);
        mbar_init(s_full, 1);
        mbar_init(p_full, 4);
        mbar_init(pv_done, 1);
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

    if (warp < 4) {
        setmaxnreg_dec<56>();
        if (warp == 0) {
            // ------------------------------------------------------------------ TMA producer: everything at once
            if (elect_one()) {
                const size_t tail_off = (static_cast<size_t>(bh) * a.S + n_main) * HD;
                mbar_arrive_expect_tx(qk_full, (1 + k_boxes) * AT_HALF_BYTES + n_tail * 128);
                tma_load_3d(smem + Cfg::Q_OFF, &tmQ, 0, q0, bh, qk_full);
                for (int i = 0; i < k_boxes; ++i)
                    tma_load_3d(smem + Cfg::K_OFF + i * AT_HALF_BYTES, &tmK, 0, i * AT_TILE, bh, qk_full);
                if (n_tail > 0) bulk_g2s(smem + Cfg::KT_OFF, ar.k + tail_off, n_tail * 128, qk_full);
                mbar_arrive_expect_tx(v_full, k_boxes * AT_HALF_BYTES + n_tail * 128);
                for (int i = 0; i < k_boxes; ++i)
                    tma_load_3d(smem + Cfg::V_OFF + i * AT_HALF_BYTES, &tmV, 0, i * AT_TILE, bh, v_full);
                if (n_tail > 0) bulk_g2s(smem + Cfg::VT_OFF, ar.v + tail_off, n_tail * 128, v_full);
                // A CTA lives ~7 us, of which the first ~2.5 were spent waiting for these tiles to come from HBM (ncu: 37 %
                // of the softmax warps' samples). Two CTAs per SM cannot hide that, so every CTA pulls the tiles of the CTA
                // that will take over a slot about one CTA lifetime from now into L2: its loads then cost an L2 hit.
                const long long tgt = static_cast<long long>(bh) * gridDim.x + blockIdx.x + ar.prefetch_stride;
                const int tbh = static_cast<int>(tgt / gridDim.x), tx = static_cast<int>(tgt - static_cast<long long>(tbh) * gridDim.x);
                if (ar.prefetch_stride > 0 && tbh < static_cast<int>(gridDim.y)) {
                    tma_prefetch_l2_3d(&tmQ, 0, tx * AT_TILE, tbh);
                    if (tx == 0) {                         // K / V are shared by the query tiles of one (batch, head)
                        for (int i = 0; i < k_boxes; ++i) {
                            tma_prefetch_l2_3d(&tmK, 0, i * AT_TILE, tbh);
                            tma_prefetch_l2_3d(&tmV, 0, i * AT_TILE, tbh);
                        }
                    }
                }
            }
            __syncwarp();
        } else if (warp == 1) {
            // ------------------------------------------------------------------ MMA issuer
            const uint32_t idesc_qk = umma_idesc_bf16(128, static_cast<uint32_t>(n_mma), 0, 0);
            constexpr uint32_t idesc_pv = umma_idesc_bf16(128, HD, 0, 1);
            const uint64_t qd = umma_desc_k_sw128(smem_u32(smem + Cfg::Q_OFF));
            const uint64_t kd = umma_desc_k_sw128(smem_u32(smem + Cfg::K_OFF));
            const uint64_t vd = umma_desc_mn_sw128(smem_u32(smem + Cfg::V_OFF), 0, 1024);
            mbar_wait(qk_full, 0);
            tc_fence_after();
            if (elect_one()) {
#pragma unroll
                for (int ks = 0; ks < HD / 16; ++ks)
                    tc_mma_f16(tmem_base, qd + ((ks * 32) >> 4), kd + ((ks * 32) >> 4), idesc_qk, ks != 0);
                tc_commit(s_full);
            }
            __syncwarp();
            mbar_wait(v_full, 0);
            mbar_wait(p_full, 0);
            tc_fence_after();
            if (elect_one()) {
                const int nks = n_mma / 16;
                for (int ks = 0; ks < nks; ++ks)          // 16 keys per step: 8 packed P columns, 2048 B of V
                    tc_mma_f16_ts(tmem_base + Cfg::O_COL, tmem_base + ks * 8, vd + ((ks * 2048) >> 4), idesc_pv, ks != 0);
                tc_commit(pv_done);
            }
            __syncwarp();
        }
    } else {
        setmaxnreg_inc<200>();
        const int quarter = warp & 3;
        const int r = quarter * 32 + lane;
        const int srow = q0 + r;
        if (q0 + quarter * 32 >= a.S) {
            // no live query row in this warp: its rows of P and O are never stored and may hold anything
            if (lane == 0) mbar_arrive(p_full);
        } else {
            const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
            const uint32_t t_s = t_lane;
            const uint32_t t_o = t_lane + Cfg::O_COL;
            // scores of the tail keys on the CUDA cores, before the tensor core has anything to show
            float st[AR_TAIL];
#pragma unroll
            for (int t = 0; t < AR_TAIL; ++t) st[t] = -INFINITY;
            if (n_tail > 0) {
                // q row r of the SWIZZLE_128B tile: 16-byte chunk j sits at chunk position j ^ (r & 7) of its 128-byte row
                mbar_wait(qk_full, 0);
                const uint8_t* qrow = smem + Cfg::Q_OFF + r * 128;
                uint4 qv[HD / 8];
#pragma unroll
                for (int i = 0; i < HD / 8; ++i) qv[i] = *reinterpret_cast<const uint4*>(qrow + ((i ^ (r & 7)) << 4));
#pragma unroll
                for (int t = 0; t < AR_TAIL; ++t) {
                    if (t < n_tail) {
                        const uint4* kp = reinterpret_cast<const uint4*>(smem + Cfg::KT_OFF + t * 128);
                        float acc0 = 0.f, acc1 = 0.f;
#pragma unroll
                        for (int i = 0; i < HD / 8; ++i) {
                            float qf[8], kf[8];
                            bf16x8_to_float(qv[i], qf);
                            bf16x8_to_float(kp[i], kf);
#pragma unroll
                            for (int e = 0; e < 8; e += 2) {
                                acc0 = fmaf(qf[e], kf[e], acc0);
                                acc1 = fmaf(qf[e + 1], kf[e + 1], acc1);
                            }
                        }
                        st[t] = acc0 + acc1;
                    }
                }
            }
            const int ncol = (n_main + 31) & ~31;          // score columns read (warp-uniform)
            mbar_wait(s_full, 0);
            tc_fence_after();
            // pass 1: the exact row maximum
            static_assert(AR_TAIL == 1, "one tail key seeds the first maximum chain");
            float m0 = st[0], m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
#pragma unroll
            for (int c = 0; c < AR_MAIN; c += 64) {
                if (c < ncol) {
                    uint32_t v[64];
                    tmem_ld_32x32_ptr(t_s + c, &v[0]);
                    tmem_ld_32x32_ptr(t_s + c + 32, &v[32]);
                    tmem_ld_wait();
                    if (c + 64 > n_main) {
#pragma unroll
                        for (int i = 0; i < 64; ++i)
                            if (c + i >= n_main) v[i] = 0xff800000u;
                    }
#pragma unroll
                    for (int i = 0; i < 64; i += 8) {
                        m0 = fmax3(m0, __uint_as_float(v[i]), __uint_as_float(v[i + 1]));
                        m1 = fmax3(m1, __uint_as_float(v[i + 2]), __uint_as_float(v[i + 3]));
                        m2 = fmax3(m2, __uint_as_float(v[i + 4]), __uint_as_float(v[i + 5]));
                        m3 = fmax3(m3, __uint_as_float(v[i + 6]), __uint_as_float(v[i + 7]));
                    }
                }
            }
            const float m = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3)) * a.scale_log2;
            // pass 2: P = exp2(s * scale - m) as packed bf16 IN PLACE (columns [c/2, c/2 + 16) were read in an earlier step)
            const float2 sc2 = splat2(a.scale_log2), nm2 = splat2(-m);
            float2 sum_a = splat2(0.f), sum_b = splat2(0.f);
#pragma unroll
            for (int c = 0; c < AR_MAIN; c += 32) {
                if (c < ncol) {
                    uint32_t v[32];
                    tmem_ld_32x32(t_s + c, v);
                    tmem_ld_wait();
                    if (c + 32 > n_main) {
#pragma unroll
                        for (int i = 0; i < 32; ++i)
                            if (c + i >= n_main) v[i] = 0xff800000u;
                    }
                    uint32_t packed[16];
#pragma unroll
                    for (int pr = 0; pr < 16; ++pr) {
                        const float2 x = ffma2(make_float2(__uint_as_float(v[2 * pr]), __uint_as_float(v[2 * pr + 1])), sc2, nm2);
                        const float2 e = ((AT_POLY_MASK >> (pr & 7)) & 1) ? ex2_poly2(x)
                                                                          : make_float2(ex2_approx(x.x), ex2_approx(x.y));
                        if (pr & 1) sum_b = fadd2(sum_b, e);
                        else sum_a = fadd2(sum_a, e);
                        packed[pr] = pack_bf16x2(e);
                    }
                    tmem_st_32x16(t_s + (c >> 1), packed);
                }
            }
            const float2 sum2 = fadd2(sum_a, sum_b);
            float l = sum2.x + sum2.y;
            float pt[AR_TAIL];
#pragma unroll
            for (int t = 0; t < AR_TAIL; ++t) {
                pt[t] = (t < n_tail) ? ex2_approx(fmaf(st[t], a.scale_log2, -m)) : 0.f;
                l += pt[t];
            }
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(p_full);

            mbar_wait(pv_done, 0);
            tc_fence_after();
            if (n_tail > 0) mbar_wait(v_full, 0);            // long complete; makes the bulk-copied tail rows visible here
            const float inv = 1.f / l;
            const int b = bh / a.H, h = bh - b * a.H;
            __nv_bfloat16* orow = (srow < a.split)
                ? a.out0 + (static_cast<size_t>(b) * a.split + srow) * a.ld0 + h * HD
                : a.out1 + (static_cast<size_t>(b) * (a.S - a.split) + (srow - a.split)) * a.ld1 + h * HD;
#pragma unroll 1
            for (int c = 0; c < HD / 32; ++c) {
                uint32_t v[32];
                tmem_ld_32x32(t_o + c * 32, v);
                tmem_ld_wait();
                float o[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) o[i] = __uint_as_float(v[i]);
#pragma unroll
                for (int t = 0; t < AR_TAIL; ++t) {          // the tail keys' share of P V, fp32
                    if (t < n_tail) {
                        const uint4* vp = reinterpret_cast<const uint4*>(smem + Cfg::VT_OFF + t * 128 + c * 64);
                        const float p = pt[t];
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            float vf[8];
                            bf16x8_to_float(vp[i], vf);
#pragma unroll
                            for (int e = 0; e < 8; ++e) o[i * 8 + e] = fmaf(p, vf[e], o[i * 8 + e]);
                        }
                    }
                }
                if (srow < a.S) {
#pragma unroll
                    for (int i = 0; i < 32; i += 8) {
                        __nv_bfloat162 p0 = __floats2bfloat162_rn(o[i] * inv, o[i + 1] * inv);
                        __nv_bfloat162 p1 = __floats2bfloat162_rn(o[i + 2] * inv, o[i + 3] * inv);
                        __nv_bfloat162 p2 = __floats2bfloat162_rn(o[i + 4] * inv, o[i + 5] * inv);
                        __nv_bfloat162 p3 = __floats2bfloat162_rn(o[i + 6] * inv, o[i + 7] * inv);
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

int g_attn_row_prefetch = 1;   // drag_debug_set key 11: 0 = the whole-row kernel does not warm L2 for later CTAs (A/B comparisons)
static int launch_attention_row(const __nv_bfloat16* q, const __nv_bfloat16* k, const __nv_bfloat16* v, int B, int H, int S,
                                AttnArgs a, cudaStream_t st) {
    using Cfg = AttnRowCfg;
    constexpr int HD = 64;
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
        DRAG_CUDA(cudaFuncSetAttribute(attention_row_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
        DRAG_CUDA(cudaFuncSetAttribute(attention_row_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        attr_set = true;
    }
    a.scale_log2 = 1.4426950408889634f / sqrtf(static_cast<float>(HD));
    AttnRowArgs ar;
    ar.a = a; ar.k = k; ar.v = v;
    static int sm_count = 0;
    if (sm_count == 0) {
        int dev = 0;
        DRAG_CUDA(cudaGetDevice(&dev));
        DRAG_CUDA(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev));
    }
    ar.prefetch_stride = g_attn_row_prefetch ? 2 * sm_count : 0;       // two resident CTAs per SM
    ar.x_first = 0;
    dim3 grid((S + AT_TILE - 1) / AT_TILE, static_cast<unsigned>(bh));
    const int slot = prof_begin(PROF_ATTENTION, 4.0 * B * H * static_cast<double>(S) * S * HD, st);
    attention_row_kernel<<<grid, 256, Cfg::SMEM, st>>>(tq, tk, tv, ar); count_launch();
    prof_end(slot, st);
    DRAG_CUDA(cudaGetLastError());
    return DRAG_OK;
}

// ---------------------------------------------------------------------------------------------- persistent whole-row kernel
// ncu of the kernel above at the ViT-L/14 shape (500 x 16 heads x 257 tokens): DRAM 27 %, L2 27 %, issue slots 26 %, tensor
// pipe 16 % - nothing is busy. A CTA lives ~13 k cycles of which its softmax warps compute for a quarter; the rest is the
// serial chain launch -> barrier init -> TMEM alloc -> HBM round trip -> Q K^T -> ... -> dealloc, and two resident CTAs cannot
// overlap it. This kernel keeps ONE CTA per SM alive over many (batch, head) items and takes every fixed cost off the chain:
//   warp 0      TMA: Q tiles 0 and 1, K, V (+ the rows past 256) of item i+1 into the other shared-memory slot while item i
//               is being computed (two 96 KB slots);
//   warp 1      MMA: S_A = Q_0 K^T and S_B = Q_1 K^T (N = up to 256) into the two halves of TMEM, later O_x = P_x V;
//   warps 4-7   softmax + epilogue of query tile 0,  warps 8-11 of query tile 1 - both tiles of a head at the same time,
//               K and V loaded once for both.
// TMEM: tile x owns columns [256 x, 256 x + 256): S -> P (packed bf16, first 128) in place, O in [128,192) of the same half.
//   warps 2, 3  query row 256 (ViT-L/14's 257th token does not fit two 128-row tiles, and a third MMA tile for one row
//               would double a warpgroup's work): one warp does that row's attention on the CUDA cores straight from the K / V
//               tiles in shared memory - 257 64-term dot products, a warp-wide softmax, 257 rank-1 updates of 64 outputs
//               (even items on warp 2, odd items on warp 3).
struct AttnRow2Cfg {
    static constexpr int SLOT_Q = 0;                               // Q tile 0 | Q tile 1
    static constexpr int SLOT_K = 2 * AT_HALF_BYTES;
    static constexpr int SLOT_V = SLOT_K + 2 * AT_HALF_BYTES;
    static constexpr int SLOT_BYTES = SLOT_V + 2 * AT_HALF_BYTES;  // 96 KB
    static constexpr int TAIL_OFF = 2 * SLOT_BYTES;                // per slot: K row 256 | V row 256 | Q row 256, 128 B each
    static constexpr int TAIL_BYTES = 3 * AR_TAIL * 128;
    static constexpr int BAR_OFF = TAIL_OFF + 2 * TAIL_BYTES;
    static constexpr int SCRATCH_OFF = BAR_OFF + 256;              // 2 warps x 264 floats: scores / probabilities of row 256
    static constexpr int SMEM = SCRATCH_OFF + 2 * 264 * 4 + 1024;
    static constexpr int TMEM_COLS = 512;
    static constexpr int THREADS = 384;
};
struct AttnRow2Args {
    AttnArgs a;
    const __nv_bfloat16 *q, *k, *v;
    int n_items;                       // B * H
    int stagger;                       // 1 = hold query tile 1 back by half an item (see the MMA warp)
};

template <uint32_t POLY_MASK>
__global__ void __launch_bounds__(AttnRow2Cfg::THREADS, 1)
attention_row2_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                      const __grid_constant__ CUtensorMap tmV, AttnRow2Args ar) {
    using Cfg = AttnRow2Cfg;
    constexpr int HD = 64;
    const AttnArgs& a = ar.a;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::BAR_OFF);
    uint64_t* qk_full = bars + 0;     // [2] per slot
    uint64_t* v_full = bars + 2;      // [2] per slot
    uint64_t* slot_free = bars + 4;   // [2] per slot: all eight softmax warps are done with the item in it
    uint64_t* s_full = bars + 6;      // [2] per query tile
    uint64_t* p_full = bars + 8;      // [2]
    uint64_t* pv_done = bars + 10;    // [2]
    uint64_t* tmem_free = bars + 12;  // [2] per query tile: O has been read, the half may be overwritten
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 15);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_main = min(a.S, AR_MAIN);
    const int n_tail = a.S - n_main;
    const int n_mma = (n_main + 15) & ~15;
    const int k_boxes = (n_main + AT_TILE - 1) / AT_TILE;
    const int n_my = (ar.n_items - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmQ);
        tma_prefetch_desc(&tmK);
        tma_prefetch_desc(&tmV);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&qk_full[i], 1);
            mbar_init(&v_full[i], 1);
            mbar_init(&slot_free[i], 8 + (n_tail > 0 ? 1 : 0));     // + the warp that computes query row 256
            mbar_init(&s_full[i], 1);
            mbar_init(&p_full[i], 4);
            mbar_init(&pv_done[i], 1);
            mbar_init(&tmem_free[i], 4);
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

    if (warp < 4) {
        setmaxnreg_dec<104>();   // 128 x (168 - 104) released = 256 x (200 - 168) taken by the softmax warps
        if (warp == 0) {
            // ------------------------------------------------------------------ TMA producer, one item ahead
            for (int k = 0; k < n_my; ++k) {
                const int s = k & 1;
                const int bh = blockIdx.x + k * gridDim.x;
                uint8_t* slot = smem + s * Cfg::SLOT_BYTES;
                uint8_t* tail = smem + Cfg::TAIL_OFF + s * Cfg::TAIL_BYTES;
                mbar_wait(&slot_free[s], ((k >> 1) & 1) ^ 1);
                if (elect_one()) {
                    const size_t tail_off = (static_cast<size_t>(bh) * a.S + n_main) * HD;
                    mbar_arrive_expect_tx(&qk_full[s], (2 + k_boxes) * AT_HALF_BYTES + 2 * n_tail * 128);
                    tma_load_3d(slot + Cfg::SLOT_Q, &tmQ, 0, 0, bh, &qk_full[s]);
                    tma_load_3d(slot + Cfg::SLOT_Q + AT_HALF_BYTES, &tmQ, 0, AT_TILE, bh, &qk_full[s]);
                    for (int i = 0; i < k_boxes; ++i)
                        tma_load_3d(slot + Cfg::SLOT_K + i * AT_HALF_BYTES, &tmK, 0, i * AT_TILE, bh, &qk_full[s]);
                    if (n_tail > 0) {
                        bulk_g2s(tail, ar.k + tail_off, n_tail * 128, &qk_full[s]);
                        bulk_g2s(tail + 2 * AR_TAIL * 128, ar.q + tail_off, n_tail * 128, &qk_full[s]);
                    }
                    mbar_arrive_expect_tx(&v_full[s], k_boxes * AT_HALF_BYTES + n_tail * 128);
                    for (int i = 0; i < k_boxes; ++i)
                        tma_load_3d(slot + Cfg::SLOT_V + i * AT_HALF_BYTES, &tmV, 0, i * AT_TILE, bh, &v_full[s]);
                    if (n_tail > 0) bulk_g2s(tail + AR_TAIL * 128, ar.v + tail_off, n_tail * 128, &v_full[s]);
                }
                __syncwarp();
            }
        } else if (warp == 1) {
            // ------------------------------------------------------------------ MMA issuer
            const uint32_t idesc_qk = umma_idesc_bf16(128, static_cast<uint32_t>(n_mma), 0, 0);
            constexpr uint32_t idesc_pv = umma_idesc_bf16(128, HD, 0, 1);
            const int nks = n_mma / 16;
            // Event-driven: each query tile is a two-state machine (Q K^T wanted / P V wanted) served as soon as its
            // barriers allow, so the two softmax warpgroups need not run in lockstep. With ar.stagger tile 1 is held back
            // until tile 0's first P V is issued: from then on one warpgroup computes while the other waits for the
            // tensor core and for its barriers, instead of both doing the same thing at the same time.
            int kx[2] = {0, 0};
            int stage[2] = {0, 0};
            bool b_enabled = ar.stagger == 0;
            while (kx[0] < n_my || kx[1] < n_my) {
#pragma unroll
                for (int x = 0; x < 2; ++x) {
                    const int k = kx[x];
                    if (k >= n_my || (x == 1 && !b_enabled)) continue;
                    const int s = k & 1, ph = (k >> 1) & 1, kp = k & 1;
                    const uint32_t slot = smem_u32(smem + s * Cfg::SLOT_BYTES);
                    if (stage[x] == 0) {
                        uint32_t ok = mbar_try_wait(&qk_full[s], ph) && mbar_try_wait(&tmem_free[x], kp ^ 1);
                        ok = __shfl_sync(0xffffffffu, ok, 0);
                        if (!ok) continue;
                        tc_fence_after();
                        const uint64_t kd = umma_desc_k_sw128(slot + Cfg::SLOT_K);
                        const uint64_t qd = umma_desc_k_sw128(slot + Cfg::SLOT_Q + x * AT_HALF_BYTES);
                        if (elect_one()) {
#pragma unroll
                            for (int ks = 0; ks < HD / 16; ++ks)
                                tc_mma_f16(tmem_base + x * 256, qd + ((ks * 32) >> 4), kd + ((ks * 32) >> 4), idesc_qk, ks != 0);
                            tc_commit(&s_full[x]);
                        }
                        __syncwarp();
                        stage[x] = 1;
                    } else {
                        uint32_t ok = mbar_try_wait(&v_full[s], ph) && mbar_try_wait(&p_full[x], kp);
                        ok = __shfl_sync(0xffffffffu, ok, 0);
                        if (!ok) continue;
                        tc_fence_after();
                        const uint64_t vd = umma_desc_mn_sw128(slot + Cfg::SLOT_V, 0, 1024);
                        if (elect_one()) {
                            for (int ks = 0; ks < nks; ++ks)
                                tc_mma_f16_ts(tmem_base + x * 256 + 128, tmem_base + x * 256 + ks * 8, vd + ((ks * 2048) >> 4),
                                              idesc_pv, ks != 0);
                            tc_commit(&pv_done[x]);
                        }
                        __syncwarp();
                        stage[x] = 0;
                        kx[x] = k + 1;
                        b_enabled = true;
                    }
                }
            }
        } else if (n_tail > 0) {
            // ------------------------------------------------------------------ query row 256 on the CUDA cores
            // lane L scores keys L, L + 32, ..., L + 224 (key 256 on every lane) into a per-warp scratch row, then owns
            // output dims 2L, 2L + 1. K and V tiles are SWIZZLE_128B: 16-byte chunk c of row j sits at chunk position
            // c ^ (j & 7). Loops stay rolled: this code runs once per item on one warp and must not evict the softmax loops
            // of the other ten from the instruction cache.
            float* prob = reinterpret_cast<float*>(smem + Cfg::SCRATCH_OFF) + (warp - 2) * 264;
            uint32_t voff[8];                                   // byte offset of dims 2L, 2L+1 inside a V row, per (row & 7)
#pragma unroll
            for (int m8 = 0; m8 < 8; ++m8) voff[m8] = ((((lane >> 2) ^ m8) << 4) + (lane & 3) * 4);
            for (int k = warp - 2; k < n_my; k += 2) {
                const int s = k & 1, ph = (k >> 1) & 1;
                const int bh = blockIdx.x + k * gridDim.x;
                const uint8_t* slot = smem + s * Cfg::SLOT_BYTES;
                const uint8_t* tail = smem + Cfg::TAIL_OFF + s * Cfg::TAIL_BYTES;
                mbar_wait(&qk_full[s], ph);
                float qf[HD];
#pragma unroll
                for (int i = 0; i < HD / 8; ++i)
                    bf16x8_to_float(*reinterpret_cast<const uint4*>(tail + 2 * AR_TAIL * 128 + i * 16), &qf[i * 8]);
                float mx;
                {
                    float acc0 = 0.f, acc1 = 0.f;
#pragma unroll
                    for (int c = 0; c < HD / 8; ++c) {
                        float kf[8];
                        bf16x8_to_float(*reinterpret_cast<const uint4*>(tail + c * 16), kf);
#pragma unroll
                        for (int e = 0; e < 8; e += 2) {
                            acc0 = fmaf(qf[c * 8 + e], kf[e], acc0);
                            acc1 = fmaf(qf[c * 8 + e + 1], kf[e + 1], acc1);
                        }
                    }
                    mx = (acc0 + acc1) * a.scale_log2;          // key 256: the same on every lane
                }
                const float s_tail = mx;
#pragma unroll 1
                for (int i = 0; i < 8; ++i) {
                    const uint8_t* krow = slot + Cfg::SLOT_K + (lane + 32 * i) * 128;
                    float acc0 = 0.f, acc1 = 0.f;
#pragma unroll
                    for (int c = 0; c < HD / 8; ++c) {
                        float kf[8];
                        bf16x8_to_float(*reinterpret_cast<const uint4*>(krow + ((c ^ (lane & 7)) << 4)), kf);
#pragma unroll
                        for (int e = 0; e < 8; e += 2) {
                            acc0 = fmaf(qf[c * 8 + e], kf[e], acc0);
                            acc1 = fmaf(qf[c * 8 + e + 1], kf[e + 1], acc1);
                        }
                    }
                    const float sj = (acc0 + acc1) * a.scale_log2;
                    prob[lane + 32 * i] = sj;
                    mx = fmaxf(mx, sj);
                }
#pragma unroll
                for (int d = 16; d >= 1; d >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, d));
                float lsum = 0.f;
#pragma unroll
                for (int i = 0; i < 8; ++i) {                   // each lane revisits its own eight entries
                    const float e = ex2_approx(prob[lane + 32 * i] - mx);
                    prob[lane + 32 * i] = e;
                    lsum += e;
                }
#pragma unroll
                for (int d = 16; d >= 1; d >>= 1) lsum += __shfl_xor_sync(0xffffffffu, lsum, d);
                const float p_tail = ex2_approx(s_tail - mx);
                lsum += p_tail;
                __syncwarp();                                   // prob[] complete before anyone reads across lanes
                mbar_wait(&v_full[s], ph);
                float o0 = 0.f, o1 = 0.f, o2 = 0.f, o3 = 0.f;
                const uint8_t* vbase = slot + Cfg::SLOT_V;
#pragma unroll 1
                for (int j0 = 0; j0 < AR_MAIN; j0 += 8) {
                    const float4 pa = *reinterpret_cast<const float4*>(prob + j0);       // broadcast reads
                    const float4 pb = *reinterpret_cast<const float4*>(prob + j0 + 4);
                    const float pj[8] = {pa.x, pa.y, pa.z, pa.w, pb.x, pb.y, pb.z, pb.w};
#pragma unroll
                    for (int jj = 0; jj < 8; jj += 2) {
                        const uint32_t u0 = *reinterpret_cast<const uint32_t*>(vbase + (j0 + jj) * 128 + voff[jj]);
                        const uint32_t u1 = *reinterpret_cast<const uint32_t*>(vbase + (j0 + jj + 1) * 128 + voff[jj + 1]);
                        o0 = fmaf(pj[jj], __uint_as_float(u0 << 16), o0);
                        o1 = fmaf(pj[jj], __uint_as_float(u0 & 0xffff0000u), o1);
                        o2 = fmaf(pj[jj + 1], __uint_as_float(u1 << 16), o2);
                        o3 = fmaf(pj[jj + 1], __uint_as_float(u1 & 0xffff0000u), o3);
                    }
                }
                {
                    const uint32_t u = *reinterpret_cast<const uint32_t*>(tail + AR_TAIL * 128 + lane * 4);
                    o0 = fmaf(p_tail, __uint_as_float(u << 16), o0);
                    o1 = fmaf(p_tail, __uint_as_float(u & 0xffff0000u), o1);
                }
                __syncwarp();                                   // prob[] is rewritten by the next item of this warp
                if (lane == 0) mbar_arrive(&slot_free[s]);       // every shared-memory read of this item is done
                const float inv = 1.f / lsum;
                const int srow = 2 * AT_TILE;
                const int b = bh / a.H, h = bh - b * a.H;
                __nv_bfloat16* orow = (srow < a.split)
                    ? a.out0 + (static_cast<size_t>(b) * a.split + srow) * a.ld0 + h * HD
                    : a.out1 + (static_cast<size_t>(b) * (a.S - a.split) + (srow - a.split)) * a.ld1 + h * HD;
                *reinterpret_cast<__nv_bfloat162*>(orow + 2 * lane) = __floats2bfloat162_rn((o0 + o2) * inv, (o1 + o3) * inv);
            }
        }
    } else {
        setmaxnreg_inc<200>();
        const int x = (warp >= 8) ? 1 : 0;
        const int quarter = warp & 3;
        const int r = quarter * 32 + lane;
        const int srow = x * AT_TILE + r;
        const int rows_here = min(a.S, 2 * AT_TILE);                  // query rows this kernel covers
        const bool live = (x * AT_TILE + quarter * 32) < rows_here;
        const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
        const uint32_t t_s = t_lane + x * 256;
        const uint32_t t_o = t_s + 128;
        const int ncol = (n_main + 31) & ~31;
        for (int k = 0; k < n_my; ++k) {
            const int s = k & 1, ph = (k >> 1) & 1, kp = k & 1;
            const int bh = blockIdx.x + k * gridDim.x;
            if (!live) {
                // no live query row in this warp: keep the barrier phases moving, in step with the MMA warp
                mbar_wait(&s_full[x], kp);
                if (lane == 0) mbar_arrive(&p_full[x]);
                mbar_wait(&pv_done[x], kp);
                if (lane == 0) {
                    mbar_arrive(&tmem_free[x]);
                    mbar_arrive(&slot_free[s]);
                }
                continue;
            }
            const uint8_t* slot = smem + s * Cfg::SLOT_BYTES;
            const uint8_t* tail = smem + Cfg::TAIL_OFF + s * Cfg::TAIL_BYTES;
            float st[AR_TAIL];
#pragma unroll
            for (int t = 0; t < AR_TAIL; ++t) st[t] = -INFINITY;
            if (n_tail > 0) {
                mbar_wait(&qk_full[s], ph);
                const uint8_t* qrow = slot + Cfg::SLOT_Q + x * AT_HALF_BYTES + r * 128;
                uint4 qv[HD / 8];
#pragma unroll
                for (int i = 0; i < HD / 8; ++i) qv[i] = *reinterpret_cast<const uint4*>(qrow + ((i ^ (r & 7)) << 4));
#pragma unroll
                for (int t = 0; t < AR_TAIL; ++t) {
                    if (t < n_tail) {
                        const uint4* kp4 = reinterpret_cast<const uint4*>(tail + t * 128);
                        float acc0 = 0.f, acc1 = 0.f;
#pragma unroll
                        for (int i = 0; i < HD / 8; ++i) {
                            float qf[8], kf[8];
                            bf16x8_to_float(qv[i], qf);
                            bf16x8_to_float(kp4[i], kf);
#pragma unroll
                            for (int e = 0; e < 8; e += 2) {
                                acc0 = fmaf(qf[e], kf[e], acc0);
                                acc1 = fmaf(qf[e + 1], kf[e + 1], acc1);
                            }
                        }
                        st[t] = acc0 + acc1;
                    }
                }
            }
            mbar_wait(&s_full[x], kp);
            tc_fence_after();
            float m0 = st[0], m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
            // rolled on purpose: fully unrolled, the two passes were 4 000 instructions and the warps stalled on instruction
            // fetch 12 % of the time (two warpgroups in different phases thrash the instruction cache)
#pragma unroll 1
            for (int c = 0; c < ncol; c += 64) {
                uint32_t v[64];
                tmem_ld_32x32_ptr(t_s + c, &v[0]);
                tmem_ld_32x32_ptr(t_s + c + 32, &v[32]);
                tmem_ld_wait();
                if (c + 64 > n_main) {
#pragma unroll
                    for (int i = 0; i < 64; ++i)
                        if (c + i >= n_main) v[i] = 0xff800000u;
                }
#pragma unroll
                for (int i = 0; i < 64; i += 8) {
                    m0 = fmax3(m0, __uint_as_float(v[i]), __uint_as_float(v[i + 1]));
                    m1 = fmax3(m1, __uint_as_float(v[i + 2]), __uint_as_float(v[i + 3]));
                    m2 = fmax3(m2, __uint_as_float(v[i + 4]), __uint_as_float(v[i + 5]));
                    m3 = fmax3(m3, __uint_as_float(v[i + 6]), __uint_as_float(v[i + 7]));
                }
            }
            const float m = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3)) * a.scale_log2;
            const float2 sc2 = splat2(a.scale_log2), nm2 = splat2(-m);
            float2 sum_a = splat2(0.f), sum_b = splat2(0.f);
#pragma unroll 1
            for (int c = 0; c < ncol; c += 64) {          // 64 columns per trip: one TMEM round trip, two P stores
                uint32_t v[64];
                tmem_ld_32x32_ptr(t_s + c, &v[0]);
                tmem_ld_32x32_ptr(t_s + c + 32, &v[32]);   // (ncol is a multiple of 32: the second half may be padding)
                tmem_ld_wait();
                if (c + 64 > n_main) {
#pragma unroll
                    for (int i = 0; i < 64; ++i)
                        if (c + i >= n_main) v[i] = 0xff800000u;
                }
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
                    if (c + hh * 32 < ncol) {
                        uint32_t packed[16];
#pragma unroll
                        for (int pr = 0; pr < 16; ++pr) {
                            const float2 xx = ffma2(make_float2(__uint_as_float(v[hh * 32 + 2 * pr]),
                                                                __uint_as_float(v[hh * 32 + 2 * pr + 1])), sc2, nm2);
                            const float2 e = ((POLY_MASK >> (pr & 7)) & 1) ? ex2_poly2(xx)
                                                                              : make_float2(ex2_approx(xx.x), ex2_approx(xx.y));
                            if (pr & 1) sum_b = fadd2(sum_b, e);
                            else sum_a = fadd2(sum_a, e);
                            packed[pr] = pack_bf16x2(e);
                        }
                        tmem_st_32x16(t_s + ((c + hh * 32) >> 1), packed);
                    }
                }
            }
            const float2 sum2 = fadd2(sum_a, sum_b);
            float l = sum2.x + sum2.y;
            float pt[AR_TAIL];
#pragma unroll
            for (int t = 0; t < AR_TAIL; ++t) {
                pt[t] = (t < n_tail) ? ex2_approx(fmaf(st[t], a.scale_log2, -m)) : 0.f;
                l += pt[t];
            }
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&p_full[x]);

            mbar_wait(&pv_done[x], kp);
            tc_fence_after();
            if (n_tail > 0) mbar_wait(&v_full[s], ph);
            const float inv = 1.f / l;
            const int b = bh / a.H, h = bh - b * a.H;
            __nv_bfloat16* orow = (srow < a.split)
                ? a.out0 + (static_cast<size_t>(b) * a.split + srow) * a.ld0 + h * HD
                : a.out1 + (static_cast<size_t>(b) * (a.S - a.split) + (srow - a.split)) * a.ld1 + h * HD;
            uint32_t ov[HD];
            tmem_ld_32x32_ptr(t_o, &ov[0]);
            tmem_ld_32x32_ptr(t_o + 32, &ov[32]);
            tmem_ld_wait();
            float o[HD];
#pragma unroll
            for (int i = 0; i < HD; ++i) o[i] = __uint_as_float(ov[i]);
#pragma unroll
            for (int t = 0; t < AR_TAIL; ++t) {
                if (t < n_tail) {
                    const uint4* vp = reinterpret_cast<const uint4*>(tail + AR_TAIL * 128 + t * 128);
                    const float p = pt[t];
#pragma unroll
                    for (int i = 0; i < HD / 8; ++i) {
                        float vf[8];
                        bf16x8_to_float(vp[i], vf);
#pragma unroll
                        for (int e = 0; e < 8; ++e) o[i * 8 + e] = fmaf(p, vf[e], o[i * 8 + e]);
                    }
                }
            }
            // this warp is done with the TMEM half and the shared-memory slot: let the next items in
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(&tmem_free[x]);
                mbar_arrive(&slot_free[s]);
            }
            if (srow < rows_here) {
#pragma unroll
                for (int i = 0; i < HD; i += 8) {
                    __nv_bfloat162 p0 = __floats2bfloat162_rn(o[i] * inv, o[i + 1] * inv);
                    __nv_bfloat162 p1 = __floats2bfloat162_rn(o[i + 2] * inv, o[i + 3] * inv);
                    __nv_bfloat162 p2 = __floats2bfloat162_rn(o[i + 4] * inv, o[i + 5] * inv);
                    __nv_bfloat162 p3 = __floats2bfloat162_rn(o[i + 6] * inv, o[i + 7] * inv);
                    uint4 u;
                    u.x = *reinterpret_cast<uint32_t*>(&p0);
                    u.y = *reinterpret_cast<uint32_t*>(&p1);
                    u.z = *reinterpret_cast<uint32_t*>(&p2);
                    u.w = *reinterpret_cast<uint32_t*>(&p3);
                    *reinterpret_cast<uint4*>(orow + i) = u;
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

int g_attn_row_stagger = 1;     // drag_debug_set key 13: 0 = both query tiles of the persistent kernel start together
int g_attn_row_poly = 0;        // drag_debug_set key 14: 1 = the persistent kernel takes 2 of 8 exponentials from the FMA-pipe polynomial
int g_attn_row_persistent = 1;  // drag_debug_set key 12: 0 = 129..260 keys take the one-tile-per-CTA whole-row kernel (A/B comparisons)
static int launch_attention_row2(const __nv_bfloat16* q, const __nv_bfloat16* k, const __nv_bfloat16* v, int B, int H, int S,
                                 AttnArgs a, cudaStream_t st) {
    using Cfg = AttnRow2Cfg;
    constexpr int HD = 64;
    CUtensorMap tq, tk, tv;
    const uint64_t bh = static_cast<uint64_t>(B) * H;
    int rc = make_tmap_bf16_3d(&tq, q, HD, S, bh, HD, static_cast<uint64_t>(S) * HD, 64, AT_TILE);
    if (rc) return rc;
    rc = make_tmap_bf16_3d(&tk, k, HD, S, bh, HD, static_cast<uint64_t>(S) * HD, 64, AT_TILE);
    if (rc) return rc;
    rc = make_tmap_bf16_3d(&tv, v, HD, S, bh, HD, static_cast<uint64_t>(S) * HD, 64, AT_TILE);
    if (rc) return rc;
    static bool attr_set = false;
    static int sm_count = 0;
    if (!attr_set) {
        int dev = 0;
        DRAG_CUDA(cudaGetDevice(&dev));
        DRAG_CUDA(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev));
        DRAG_CUDA(cudaFuncSetAttribute(attention_row2_kernel<0u>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
        DRAG_CUDA(cudaFuncSetAttribute(attention_row2_kernel<AT_POLY_MASK>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
        attr_set = true;
    }
    a.scale_log2 = 1.4426950408889634f / sqrtf(static_cast<float>(HD));
    AttnRow2Args ar;
    ar.a = a; ar.q = q; ar.k = k; ar.v = v;
    ar.n_items = static_cast<int>(bh);
    ar.stagger = g_attn_row_stagger;
    const unsigned grid = static_cast<unsigned>(bh < static_cast<uint64_t>(sm_count) ? bh : sm_count);
    const int slot = prof_begin(PROF_ATTENTION, 4.0 * B * H * static_cast<double>(S) * S * HD, st);
    // The softmax warps of this kernel are issue- / latency-bound, not MUFU-bound (ncu: XU pipe 28 %, issue slots 45 %): every
    // exponential on MUFU.EX2 is fewer instructions than the polynomial mix of the head-dim-128 kernel.
    if (g_attn_row_poly) attention_row2_kernel<AT_POLY_MASK><<<grid, Cfg::THREADS, Cfg::SMEM, st>>>(tq, tk, tv, ar);
    else attention_row2_kernel<0u><<<grid, Cfg::THREADS, Cfg::SMEM, st>>>(tq, tk, tv, ar);
    count_launch();
    DRAG_CUDA(cudaGetLastError());
    prof_end(slot, st);
    return DRAG_OK;
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
    if (S <= AR_MAIN + AR_TAIL && !g_attn_force_pp && !g_attn_no_row) {
        if (S > AT_TILE && g_attn_row_persistent) return launch_attention_row2(q, k, v, B, H, S, a, st);
        return launch_attention_row(q, k, v, B, H, S, a, st);
    }
    if (S <= 4 * AT_TILE && !g_attn_force_pp) return launch_attention<64, false>(q, k, v, B, H, S, a, st);
    return launch_attention<64, true>(q, k, v, B, H, S, a, st);
}

}  // namespace drag
