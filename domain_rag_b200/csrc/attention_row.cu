// Head-dim-64 attention for short sequences (CLIP ViT: 50 / 197 / 257 tokens) on tcgen05: the whole score row in TMEM at once.
// nn.MultiheadAttention inside clip.encode_image, retrieval/clip100_resnet_style_all_shots.py:171.
#include "attention_common.cuh"

namespace drag {

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
    static constexpr int XCH_OFF = SCRATCH_OFF + 2 * 264 * 4;      // split form: 2 tiles x 128 rows x {max0, max1, sum0, sum1, p_tail}
    static constexpr int SMEM = XCH_OFF + 2 * 128 * 5 * 4 + 1024;
    static constexpr int TMEM_COLS = 512;
    static constexpr int THREADS = 384;
    static constexpr int THREADS_SPLIT = 640;                      // {TMA, MMA, 2 x row 256} + 4 x 4 softmax warps
};
struct AttnRow2Args {
    AttnArgs a;
    const __nv_bfloat16 *q, *k, *v;
    int n_items;                       // B * H
    int stagger;                       // 1 = hold query tile 1 back by half an item (see the MMA warp)
    int pair;                          // 1 (S <= 128): an item is TWO heads, tile x = head 2 * item + x with its own Q / K / V
};

// SPLIT: every score row is shared by TWO threads (columns [0,128) and [128,256)), 16 softmax warps instead of 8. The
// one-thread-per-row form runs its SM sub-partitions at 45 % issue utilisation: two dependent instruction streams per
// sub-partition cannot cover MUFU / TMEM latency, and when one warpgroup waits for the tensor core only one stream is left.
// The halves exchange row maximum and row sum through shared memory (one 64-thread named barrier each); half 1 keeps its P in
// its own S columns ([128,192), so it never overwrites scores half 0 still reads) and O moves to [192,256).
template <uint32_t POLY_MASK, bool SPLIT>
__global__ void __launch_bounds__(SPLIT ? AttnRow2Cfg::THREADS_SPLIT : AttnRow2Cfg::THREADS, 1)
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
    // S <= 128 (ViT-B/32: 50 tokens): one 128-row tile per head, so an item is a PAIR of heads - tile x works on head
    // 2 * item + x with its own Q / K / V boxes in the same slot layout (Q_x, K box x, V box x); nothing is shared.
    const bool pair = ar.pair != 0;
    const int n_work = pair ? (ar.n_items + 1) / 2 : ar.n_items;
    const int n_my = (n_work - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmQ);
        tma_prefetch_desc(&tmK);
        tma_prefetch_desc(&tmV);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&qk_full[i], 1);
            mbar_init(&v_full[i], 1);
            mbar_init(&slot_free[i], (SPLIT ? 16 : 8) + (n_tail > 0 ? 1 : 0));     // + the warp that computes query row 256
            mbar_init(&s_full[i], 1);
            mbar_init(&p_full[i], SPLIT ? 8 : 4);
            mbar_init(&pv_done[i], 1);
            mbar_init(&tmem_free[i], SPLIT ? 8 : 4);
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
        // setmaxnreg.inc can only take what .dec of the same CTA released: 384 threads launch with 168 registers, 128 x (168 -
        // 104) released = 256 x (200 - 168) taken; 640 threads launch with 96: 128 x (96 - 64) released = 512 x (104 - 96) taken
        if constexpr (SPLIT) setmaxnreg_dec<64>();
        else setmaxnreg_dec<104>();
        if (warp == 0) {
            // ------------------------------------------------------------------ TMA producer, one item ahead
            for (int k = 0; k < n_my; ++k) {
                const int s = k & 1;
                const int bh = blockIdx.x + k * gridDim.x;
                uint8_t* slot = smem + s * Cfg::SLOT_BYTES;
                uint8_t* tail = smem + Cfg::TAIL_OFF + s * Cfg::TAIL_BYTES;
                mbar_wait(&slot_free[s], ((k >> 1) & 1) ^ 1);
                if (pair) {
                    if (elect_one()) {
                        const int n_heads = (2 * bh + 1 < ar.n_items) ? 2 : 1;
                        mbar_arrive_expect_tx(&qk_full[s], 2 * n_heads * AT_HALF_BYTES);
                        for (int x = 0; x < n_heads; ++x) {
                            tma_load_3d(slot + Cfg::SLOT_Q + x * AT_HALF_BYTES, &tmQ, 0, 0, 2 * bh + x, &qk_full[s]);
                            tma_load_3d(slot + Cfg::SLOT_K + x * AT_HALF_BYTES, &tmK, 0, 0, 2 * bh + x, &qk_full[s]);
                        }
                        mbar_arrive_expect_tx(&v_full[s], n_heads * AT_HALF_BYTES);
                        for (int x = 0; x < n_heads; ++x)
                            tma_load_3d(slot + Cfg::SLOT_V + x * AT_HALF_BYTES, &tmV, 0, 0, 2 * bh + x, &v_full[s]);
                    }
                } else if (elect_one()) {
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
                        const uint64_t kd = umma_desc_k_sw128(slot + Cfg::SLOT_K + (pair ? x * AT_HALF_BYTES : 0));
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
                        const uint64_t vd = umma_desc_mn_sw128(slot + Cfg::SLOT_V + (pair ? x * AT_HALF_BYTES : 0), 0, 1024);
                        if (elect_one()) {
                            for (int ks = 0; ks < nks; ++ks) {
                                // P of keys [128,256) sits at columns [128,192) in the split form, O at [192,256)
                                const int p_col = (SPLIT && ks >= 8) ? 128 + (ks - 8) * 8 : ks * 8;
                                tc_mma_f16_ts(tmem_base + x * 256 + (SPLIT ? 192 : 128), tmem_base + x * 256 + p_col,
                                              vd + ((ks * 2048) >> 4), idesc_pv, ks != 0);
                            }
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
                const uint8_t* qrow = tail + 2 * AR_TAIL * 128;       // q row 256 is re-read (broadcast) per key: no 64 live registers
                float mx;
                {
                    float acc0 = 0.f, acc1 = 0.f;
#pragma unroll
                    for (int c = 0; c < HD / 8; ++c) {
                        float kf[8], qf[8];
                        bf16x8_to_float(*reinterpret_cast<const uint4*>(tail + c * 16), kf);
                        bf16x8_to_float(*reinterpret_cast<const uint4*>(qrow + c * 16), qf);
#pragma unroll
                        for (int e = 0; e < 8; e += 2) {
                            acc0 = fmaf(qf[e], kf[e], acc0);
                            acc1 = fmaf(qf[e + 1], kf[e + 1], acc1);
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
                        float kf[8], qf[8];
                        bf16x8_to_float(*reinterpret_cast<const uint4*>(krow + ((c ^ (lane & 7)) << 4)), kf);
                        bf16x8_to_float(*reinterpret_cast<const uint4*>(qrow + c * 16), qf);
#pragma unroll
                        for (int e = 0; e < 8; e += 2) {
                            acc0 = fmaf(qf[e], kf[e], acc0);
                            acc1 = fmaf(qf[e + 1], kf[e + 1], acc1);
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
    } else if constexpr (SPLIT) {
        setmaxnreg_inc<104>();
        const int w4 = warp - 4;
        const int x = w4 >> 3, hf = (w4 >> 2) & 1, quarter = w4 & 3;
        const int r = quarter * 32 + lane;
        const int srow = pair ? r : x * AT_TILE + r;
        const int rows_here = pair ? a.S : min(a.S, 2 * AT_TILE);
        const bool live = (pair ? quarter * 32 : x * AT_TILE + quarter * 32) < rows_here;
        const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
        const uint32_t t_s = t_lane + x * 256;
        const uint32_t t_p = t_s + hf * 128;                          // this half's P: in place over its own S columns
        const uint32_t t_o = t_s + 192 + hf * 32;                     // this half's 32 output columns
        const int ncol = (n_main + 31) & ~31;
        const int c_lo = hf * 128, c_hi = min(ncol, c_lo + 128);      // this thread's score columns
        float* xch = reinterpret_cast<float*>(smem + Cfg::XCH_OFF) + (x * AT_TILE + r) * 5;   // max0 max1 sum0 sum1 p_tail
        const int bar_id = 1 + x * 4 + quarter;                       // the two warps that share these 32 rows
        for (int k = 0; k < n_my; ++k) {
            const int s = k & 1, ph = (k >> 1) & 1, kp = k & 1;
            const int it = blockIdx.x + k * gridDim.x;
            const int bh = pair ? 2 * it + x : it;                    // (batch, head) of this tile
            if (!live || bh >= ar.n_items) {                          // (an odd head count leaves tile 1 of the last item empty)
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
            float st = -INFINITY;
            if (n_tail > 0 && hf == 0) {                              // key 256: half 0 scores it
                mbar_wait(&qk_full[s], ph);
                const uint8_t* qrow = slot + Cfg::SLOT_Q + x * AT_HALF_BYTES + r * 128;
                const uint4* kp4 = reinterpret_cast<const uint4*>(tail);
                float acc0 = 0.f, acc1 = 0.f;
#pragma unroll
                for (int i = 0; i < HD / 8; ++i) {
                    float qf[8], kf[8];
                    bf16x8_to_float(*reinterpret_cast<const uint4*>(qrow + ((i ^ (r & 7)) << 4)), qf);
                    bf16x8_to_float(kp4[i], kf);
#pragma unroll
                    for (int e = 0; e < 8; e += 2) {
                        acc0 = fmaf(qf[e], kf[e], acc0);
                        acc1 = fmaf(qf[e + 1], kf[e + 1], acc1);
                    }
                }
                st = acc0 + acc1;
            }
            mbar_wait(&s_full[x], kp);
            tc_fence_after();
            float m0 = st, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
#pragma unroll 1
            for (int c = c_lo; c < c_hi; c += 32) {
                uint32_t v[32];
                tmem_ld_32x32(t_s + c, v);
                tmem_ld_wait();
                if (c + 32 > n_main) {
#pragma unroll
                    for (int i = 0; i < 32; ++i)
                        if (c + i >= n_main) v[i] = 0xff800000u;
                }
#pragma unroll
                for (int i = 0; i < 32; i += 8) {
                    m0 = fmax3(m0, __uint_as_float(v[i]), __uint_as_float(v[i + 1]));
                    m1 = fmax3(m1, __uint_as_float(v[i + 2]), __uint_as_float(v[i + 3]));
                    m2 = fmax3(m2, __uint_as_float(v[i + 4]), __uint_as_float(v[i + 5]));
                    m3 = fmax3(m3, __uint_as_float(v[i + 6]), __uint_as_float(v[i + 7]));
                }
            }
            const float m_mine = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
            xch[hf] = m_mine;
            asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory");
            const float m = fmaxf(m_mine, xch[hf ^ 1]) * a.scale_log2;
            const float2 sc2 = splat2(a.scale_log2), nm2 = splat2(-m);
            float2 sum_a = splat2(0.f), sum_b = splat2(0.f);
#pragma unroll 1
            for (int c = c_lo; c < c_hi; c += 32) {
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
                    const float2 xx = ffma2(make_float2(__uint_as_float(v[2 * pr]), __uint_as_float(v[2 * pr + 1])), sc2, nm2);
                    const float2 e = ((POLY_MASK >> (pr & 7)) & 1) ? ex2_poly2(xx)
                                                                   : make_float2(ex2_approx(xx.x), ex2_approx(xx.y));
                    if (pr & 1) sum_b = fadd2(sum_b, e);
                    else sum_a = fadd2(sum_a, e);
                    packed[pr] = pack_bf16x2(e);
                }
                tmem_st_32x16(t_p + ((c - c_lo) >> 1), packed);
            }
            const float2 sum2 = fadd2(sum_a, sum_b);
            float pt = (n_tail > 0 && hf == 0) ? ex2_approx(fmaf(st, a.scale_log2, -m)) : 0.f;
            const float l_mine = sum2.x + sum2.y + pt;
            xch[2 + hf] = l_mine;
            if (hf == 0) xch[4] = pt;
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&p_full[x]);
            asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory");
            const float l = l_mine + xch[2 + (hf ^ 1)];
            if (hf == 1) pt = xch[4];

            mbar_wait(&pv_done[x], kp);
            tc_fence_after();
            if (n_tail > 0) mbar_wait(&v_full[s], ph);
            const float inv = 1.f / l;
            const int b = bh / a.H, h = bh - b * a.H;
            __nv_bfloat16* orow = ((srow < a.split)
                ? a.out0 + (static_cast<size_t>(b) * a.split + srow) * a.ld0 + h * HD
                : a.out1 + (static_cast<size_t>(b) * (a.S - a.split) + (srow - a.split)) * a.ld1 + h * HD) + hf * 32;
            uint32_t ov[32];
            tmem_ld_32x32(t_o, ov);
            tmem_ld_wait();
            float o[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = __uint_as_float(ov[i]);
            if (n_tail > 0) {
                const uint4* vp = reinterpret_cast<const uint4*>(tail + AR_TAIL * 128 + hf * 64);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    float vf[8];
                    bf16x8_to_float(vp[i], vf);
#pragma unroll
                    for (int e = 0; e < 8; ++e) o[i * 8 + e] = fmaf(pt, vf[e], o[i * 8 + e]);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(&tmem_free[x]);
                mbar_arrive(&slot_free[s]);
            }
            if (srow < rows_here) {
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
                    *reinterpret_cast<uint4*>(orow + i) = u;
                }
            }
        }
    } else {
        setmaxnreg_inc<200>();
        const int x = (warp >= 8) ? 1 : 0;
        const int quarter = warp & 3;
        const int r = quarter * 32 + lane;
        const int srow = pair ? r : x * AT_TILE + r;                  // pair mode: tile x is its own head, rows 0..S-1
        const int rows_here = pair ? a.S : min(a.S, 2 * AT_TILE);     // query rows this kernel covers
        const bool live = (pair ? quarter * 32 : x * AT_TILE + quarter * 32) < rows_here;
        const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
        const uint32_t t_s = t_lane + x * 256;
        const uint32_t t_o = t_s + 128;
        const int ncol = (n_main + 31) & ~31;
        for (int k = 0; k < n_my; ++k) {
            const int s = k & 1, ph = (k >> 1) & 1, kp = k & 1;
            const int it = blockIdx.x + k * gridDim.x;
            const int bh = pair ? 2 * it + x : it;                    // (batch, head) of this tile
            if (!live || bh >= ar.n_items) {                          // (an odd head count leaves tile 1 of the last item empty)
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
// drag_debug_set key 16: 1 = two softmax threads per score row in the persistent kernel (SPLIT). MEASURED at the ViT-L/14 shape:
// 0.381 ms against 0.326 ms with one thread per row (ViT-B/16 shape: 0.198 vs 0.202) - the second negative result of this
// kind (see attention_tcgen05_split_kernel): 16 softmax warps at 104 registers, two named barriers and an exchange through
// shared memory per item cost more than the extra instruction streams hide. Kept for A/B; the default is one thread per row.
int g_attn_row_split = 0;
int g_attn_row_pair = 1;        // drag_debug_set key 17: 0 = up to 128 keys take the one-tile-per-CTA kernel instead of the persistent
                                // kernel in pair mode (two heads per item)
int g_attn_row_persistent = 1;  // drag_debug_set key 12: 0 = 129..260 keys take the one-tile-per-CTA whole-row kernel (A/B comparisons)
static int launch_attention_row2(const __nv_bfloat16* q, const __nv_bfloat16* k, const __nv_bfloat16* v, int B, int H, int S,
                                 AttnArgs a, bool pair, cudaStream_t st) {
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
        DRAG_CUDA(cudaFuncSetAttribute(attention_row2_kernel<0u, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
        DRAG_CUDA(cudaFuncSetAttribute(attention_row2_kernel<AT_POLY_MASK, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
        DRAG_CUDA(cudaFuncSetAttribute(attention_row2_kernel<0u, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
        attr_set = true;
    }
    a.scale_log2 = 1.4426950408889634f / sqrtf(static_cast<float>(HD));
    AttnRow2Args ar;
    ar.a = a; ar.q = q; ar.k = k; ar.v = v;
    ar.n_items = static_cast<int>(bh);
    ar.stagger = g_attn_row_stagger;
    ar.pair = pair ? 1 : 0;
    const uint64_t n_work = pair ? (bh + 1) / 2 : bh;
    const unsigned grid = static_cast<unsigned>(n_work < static_cast<uint64_t>(sm_count) ? n_work : sm_count);
    const int slot = prof_begin(PROF_ATTENTION, 4.0 * B * H * static_cast<double>(S) * S * HD, st);
    // The softmax warps of this kernel are issue- / latency-bound, not MUFU-bound (ncu: XU pipe 28 %, issue slots 45 %): every
    // exponential on MUFU.EX2 is fewer instructions than the polynomial mix of the head-dim-128 kernel.
    if (g_attn_row_split) attention_row2_kernel<0u, true><<<grid, Cfg::THREADS_SPLIT, Cfg::SMEM, st>>>(tq, tk, tv, ar);
    else if (g_attn_row_poly) attention_row2_kernel<AT_POLY_MASK, false><<<grid, Cfg::THREADS, Cfg::SMEM, st>>>(tq, tk, tv, ar);
    else attention_row2_kernel<0u, false><<<grid, Cfg::THREADS, Cfg::SMEM, st>>>(tq, tk, tv, ar);
    count_launch();
    DRAG_CUDA(cudaGetLastError());
    prof_end(slot, st);
    return DRAG_OK;
}

bool attention_row_eligible(int head_dim, int S) { return head_dim == 64 && S <= AR_MAIN + AR_TAIL; }

int launch_attention_row_any(const __nv_bfloat16* q, const __nv_bfloat16* k, const __nv_bfloat16* v, int B, int H, int S,
                             AttnArgs a, cudaStream_t st) {
    if (S > AT_TILE && g_attn_row_persistent) return launch_attention_row2(q, k, v, B, H, S, a, false, st);
    if (S <= AT_TILE && g_attn_row_persistent && g_attn_row_pair) return launch_attention_row2(q, k, v, B, H, S, a, true, st);
    return launch_attention_row(q, k, v, B, H, S, a, st);
}

}  // namespace drag
