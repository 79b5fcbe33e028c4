// Non-causal joint (text + image) attention for the Flux MMDiT blocks, head dim 128, on tcgen05.
//   out[b][s][h*128 + :] = softmax(q k^T / sqrt(128)) v      q,k,v bf16 [B][H][S][128]
// (F.scaled_dot_product_attention inside diffusers' Flux attention processor, reached from
//  pipe(...) at batch_generate_flux_kshot.py:467-474 / outpainting_updown_sampling_redux.py:1246-1257.)
//
// One CTA per (128-query tile, batch*head); flash-attention style loop over 128-key tiles:
//   warp 0   TMA: Q once, K/V tiles into 2-stage rings (3-D tensor maps, SWIZZLE_128B, rows past
//            the sequence end are zero-filled by TMA);
//   warp 1   MMA issuer: S[j%2] = Q K_j^T (UMMA 128x128x16, both operands K-major) into one of two
//            TMEM score buffers, so QK^T of tile j+1 overlaps the softmax of tile j; then
//            O += P_j V_j (UMMA 128x64x16 twice per k-step, V consumed MN-major straight from the
//            row-major tile - no transpose);
//   warps 2-5 softmax: thread = query row. tcgen05.ld the score row, online max/sum in the exp2
//            domain, lazy rescale of the TMEM-resident O (only when the running max grows by > 8),
//            P written to shared memory as bf16 in the swizzled K-major layout the MMA reads.
// TMEM: S0 [0,128) S1 [128,256) O [256,384).
#include <cuda.h>

#include "common.cuh"
#include "flux_ops.cuh"
#include "prof.cuh"
#include "ptx.cuh"

namespace drag {

constexpr int AT_THREADS = 192;
constexpr int AT_TILE = 128;
constexpr int AT_HALF_BYTES = AT_TILE * 64 * 2;        // 16 KB: 128 rows x 64 bf16
constexpr float AT_RESCALE_THRESHOLD = 8.0f;           // log2 domain

// Head dim HD = 128 (Flux) or 64 (CLIP ViT): tiles are HD/64 swizzled halves of 128 rows x 64 bf16.
template <int HD>
struct AttnCfg {
    static constexpr int NH = HD / 64;
    static constexpr int TILE_BYTES = NH * AT_HALF_BYTES;          // Q / K / V tile
    static constexpr int P_BYTES = 2 * AT_HALF_BYTES;              // P is always 128 x 128
    static constexpr int Q_OFF = 0;
    static constexpr int K_OFF = TILE_BYTES;
    static constexpr int V_OFF = K_OFF + 2 * TILE_BYTES;
    static constexpr int P_OFF = V_OFF + 2 * TILE_BYTES;
    static constexpr int BAR_OFF = P_OFF + P_BYTES;
    static constexpr int SMEM = BAR_OFF + 256 + 1024;
};

struct AttnArgs {
    __nv_bfloat16* out0;   // tokens [0, split): row b*split + s, leading dim ld0
    __nv_bfloat16* out1;   // tokens [split, S): row b*(S-split) + s-split, leading dim ld1
    int ld0, ld1, split;
    int S, H;
    float scale_log2;      // log2(e) / sqrt(head_dim)
    uint32_t v_lbo, v_sbo; // MN-major V descriptor strides (bytes)
};

template <int HD>
__global__ void __launch_bounds__(AT_THREADS, 1)
attention_tcgen05_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                         const __grid_constant__ CUtensorMap tmV, AttnArgs a) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
    using Cfg = AttnCfg<HD>;
    constexpr int AT_HD = HD;
    constexpr int AT_TILE_BYTES = Cfg::TILE_BYTES;
    constexpr int AT_Q_OFF = Cfg::Q_OFF, AT_K_OFF = Cfg::K_OFF, AT_V_OFF = Cfg::V_OFF, AT_P_OFF = Cfg::P_OFF;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::BAR_OFF);
    uint64_t* q_full = bars + 0;
    uint64_t* k_full = bars + 1;    // [2]
    uint64_t* k_empty = bars + 3;   // [2]
    uint64_t* v_full = bars + 5;    // [2]
    uint64_t* v_empty = bars + 7;   // [2]
    uint64_t* s_full = bars + 9;    // [2]
    uint64_t* s_empty = bars + 11;  // [2]
    uint64_t* p_full = bars + 13;
    uint64_t* pv_done = bars + 14;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * AT_TILE;
    const int bh = blockIdx.y;
    const int n_tiles = (a.S + AT_TILE - 1) / AT_TILE;

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
            mbar_init(&s_empty[i], 4);
        }
        mbar_init(p_full, 4);
        mbar_init(pv_done, 1);
        fence_mbar_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0 && lane == 0) {
        // ------------------------------------------------------------------ TMA producer
        mbar_arrive_expect_tx(q_full, AT_TILE_BYTES);
#pragma unroll
        for (int hf = 0; hf < Cfg::NH; ++hf)
            tma_load_3d(smem + AT_Q_OFF + hf * AT_HALF_BYTES, &tmQ, hf * 64, q0, bh, q_full);
        for (int j = 0; j < n_tiles; ++j) {
            const int st = j & 1, par = (j >> 1) & 1;
            uint8_t* kd = smem + AT_K_OFF + st * AT_TILE_BYTES;
            uint8_t* vd = smem + AT_V_OFF + st * AT_TILE_BYTES;
            mbar_wait(&k_empty[st], par ^ 1);
            mbar_arrive_expect_tx(&k_full[st], AT_TILE_BYTES);
#pragma unroll
            for (int hf = 0; hf < Cfg::NH; ++hf)
                tma_load_3d(kd + hf * AT_HALF_BYTES, &tmK, hf * 64, j * AT_TILE, bh, &k_full[st]);
            mbar_wait(&v_empty[st], par ^ 1);
            mbar_arrive_expect_tx(&v_full[st], AT_TILE_BYTES);
#pragma unroll
            for (int hf = 0; hf < Cfg::NH; ++hf)
                tma_load_3d(vd + hf * AT_HALF_BYTES, &tmV, hf * 64, j * AT_TILE, bh, &v_full[st]);
        }
    } else if (warp == 1 && lane == 0) {
        // ------------------------------------------------------------------ MMA issuer
        constexpr uint32_t idesc_qk = umma_idesc_bf16(128, 128, 0, 0);
        constexpr uint32_t idesc_pv = umma_idesc_bf16(128, 64, 0, 1);   // B (V) is MN-major
        const uint32_t q_addr = smem_u32(smem + AT_Q_OFF);
        const uint32_t p_addr = smem_u32(smem + AT_P_OFF);
        auto issue_qk = [&](int t) {
            const int st = t & 1, par = (t >> 1) & 1;
            mbar_wait(&k_full[st], par);
            mbar_wait(&s_empty[st], par ^ 1);
            tc_fence_after();
            const uint32_t k_addr = smem_u32(smem + AT_K_OFF + st * AT_TILE_BYTES);
#pragma unroll
            for (int ks = 0; ks < AT_HD / 16; ++ks) {
                const uint32_t off = (ks >> 2) * AT_HALF_BYTES + (ks & 3) * 32;
                tc_mma_f16(tmem_base + st * 128, umma_desc_k_sw128(q_addr + off), umma_desc_k_sw128(k_addr + off),
                           idesc_qk, ks != 0);
            }
            tc_commit(&s_full[st]);
            tc_commit(&k_empty[st]);
        };
        mbar_wait(q_full, 0);
        issue_qk(0);
        for (int j = 0; j < n_tiles; ++j) {
            if (j + 1 < n_tiles) issue_qk(j + 1);
            const int st = j & 1, par = (j >> 1) & 1;
            mbar_wait(p_full, j & 1);
            mbar_wait(&v_full[st], par);
            tc_fence_after();
            const uint32_t v_addr = smem_u32(smem + AT_V_OFF + st * AT_TILE_BYTES);
#pragma unroll
            for (int ks = 0; ks < AT_TILE / 16; ++ks) {
                const uint64_t pd = umma_desc_k_sw128(p_addr + (ks >> 2) * AT_HALF_BYTES + (ks & 3) * 32);
#pragma unroll
                for (int nh = 0; nh < Cfg::NH; ++nh) {
                    // V half nh: [128 kv rows][64 hd], 128-byte rows; 16 kv rows per k-step = 2048 B
                    const uint64_t vdsc = umma_desc_mn_sw128(v_addr + nh * AT_HALF_BYTES + ks * 2048, a.v_lbo, a.v_sbo);
                    tc_mma_f16(tmem_base + 256 + nh * 64, pd, vdsc, idesc_pv, (j | ks) != 0);
                }
            }
            tc_commit(&v_empty[st]);
            tc_commit(pv_done);
        }
    } else if (warp >= 2) {
        // ------------------------------------------------------------------ softmax + epilogue
        const int quarter = warp & 3;
        const int r = quarter * 32 + lane;                 // query row inside the tile
        const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
        uint8_t* p_row = smem + AT_P_OFF + r * 128;
        float m = -INFINITY, l = 0.f;
        for (int j = 0; j < n_tiles; ++j) {
            const int st = j & 1, par = (j >> 1) & 1;
            mbar_wait(&s_full[st], par);
            tc_fence_after();
            const uint32_t t_s = t_lane + st * 128;
            const int kv_valid = a.S - j * AT_TILE;        // columns >= kv_valid are past the sequence end
            float mx = -INFINITY;
#pragma unroll 1
            for (int c = 0; c < 4; ++c) {
                uint32_t v[32];
                tmem_ld_32x32(t_s + c * 32, v);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const float sv = (c * 32 + i < kv_valid) ? __uint_as_float(v[i]) : -INFINITY;
                    mx = fmaxf(mx, sv);
                }
            }
            const float m_new = fmaxf(m, mx * a.scale_log2);
            bool waited = false;
            if (__any_sync(0xffffffffu, m_new > m + AT_RESCALE_THRESHOLD)) {
                const float alpha = exp2f(m - m_new);      // 0 on the first tile (m = -inf)
                if (j > 0) {
                    mbar_wait(pv_done, (j - 1) & 1);       // O += P_{j-1} V_{j-1} has landed
                    tc_fence_after();
                    waited = true;
#pragma unroll 1
                    for (int c = 0; c < HD / 32; ++c) {
                        uint32_t v[32];
                        tmem_ld_32x32(t_lane + 256 + c * 32, v);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * alpha);
                        tmem_st_32x32(t_lane + 256 + c * 32, v);
                    }
                    tmem_st_wait();
                }
                l *= alpha;
                m = m_new;
            }
            if (j > 0 && !waited) mbar_wait(pv_done, (j - 1) & 1);   // P buffer is free again
            float rowsum = 0.f;
#pragma unroll 1
            for (int c = 0; c < 4; ++c) {
                uint32_t v[32];
                tmem_ld_32x32(t_s + c * 32, v);
                tmem_ld_wait();
                uint32_t packed[16];
#pragma unroll
                for (int i = 0; i < 32; i += 2) {
                    float p0 = (c * 32 + i < kv_valid) ? exp2f(__uint_as_float(v[i]) * a.scale_log2 - m) : 0.f;
                    float p1 = (c * 32 + i + 1 < kv_valid) ? exp2f(__uint_as_float(v[i + 1]) * a.scale_log2 - m) : 0.f;
                    rowsum += p0 + p1;
                    __nv_bfloat162 pk = __floats2bfloat162_rn(p0, p1);
                    packed[i >> 1] = *reinterpret_cast<uint32_t*>(&pk);
                }
                // kv columns [c*32, c*32+32) -> half c/2, 16-byte chunks (c%2)*4 + g, XOR-swizzled by row
                uint8_t* dst = p_row + (c >> 1) * AT_HALF_BYTES;
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    const int chunk = ((c & 1) * 4 + g) ^ (r & 7);
                    *reinterpret_cast<uint4*>(dst + chunk * 16) =
                        make_uint4(packed[4 * g], packed[4 * g + 1], packed[4 * g + 2], packed[4 * g + 3]);
                }
            }
            l += rowsum;
            tc_fence_before();
            fence_proxy_async();           // generic-proxy smem writes -> visible to the MMA (async proxy)
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(&s_empty[st]);
                mbar_arrive(p_full);
            }
        }
        mbar_wait(pv_done, (n_tiles - 1) & 1);
        tc_fence_after();
        const float inv = 1.f / l;
        const int srow = q0 + r;
        const int b = bh / a.H, h = bh - b * a.H;
        __nv_bfloat16* orow = (srow < a.split)
            ? a.out0 + (static_cast<size_t>(b) * a.split + srow) * a.ld0 + h * AT_HD
            : a.out1 + (static_cast<size_t>(b) * (a.S - a.split) + (srow - a.split)) * a.ld1 + h * AT_HD;
#pragma unroll 1
        for (int c = 0; c < HD / 32; ++c) {
            uint32_t v[32];
            tmem_ld_32x32(t_lane + 256 + c * 32, v);
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
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

// Debug knob (drag_debug_set): MN-major V descriptor strides.
uint32_t g_attn_v_lbo = 0, g_attn_v_sbo = 1024;

template <int HD>
static int launch_attention(const __nv_bfloat16* q, const __nv_bfloat16* k, const __nv_bfloat16* v, int B, int H, int S,
                            AttnArgs a, cudaStream_t st) {
    using Cfg = AttnCfg<HD>;
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
        DRAG_CUDA(cudaFuncSetAttribute(attention_tcgen05_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       Cfg::SMEM));
        attr_set = true;
    }
    a.scale_log2 = 1.4426950408889634f / sqrtf(static_cast<float>(HD));
    dim3 grid((S + AT_TILE - 1) / AT_TILE, static_cast<unsigned>(bh));
    const int slot = prof_begin(PROF_ATTENTION, 4.0 * B * H * static_cast<double>(S) * S * HD, st);
    attention_tcgen05_kernel<HD><<<grid, AT_THREADS, Cfg::SMEM, st>>>(tq, tk, tv, a);
    prof_end(slot, st);
    DRAG_CUDA(cudaGetLastError());
    return DRAG_OK;
}

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
    a.v_lbo = g_attn_v_lbo; a.v_sbo = g_attn_v_sbo;
    if (head_dim == 128) return launch_attention<128>(q, k, v, B, H, S, a, st);
    return launch_attention<64>(q, k, v, B, H, S, a, st);
}

}  // namespace drag
