// bf16 GEMM core on 5th-gen tensor cores: C[M,N] = A[M,K] * W[N,K]^T with fused epilogues.
//
// Persistent, warp-specialised, one CTA per SM (grid = min(tiles, #SM)):
//   warp 0 (1 lane)  TMA producer: A/W tiles (128 x 64 and BN x 64 bf16, SWIZZLE_128B) into a
//                    STAGES-deep shared-memory ring, completion by mbarrier transaction bytes;
//   warp 1 (1 lane)  MMA issuer: tcgen05.mma cta_group::1 kind::f16, UMMA 128 x BN x 16, fp32
//                    accumulators in TMEM (2 accumulator stages x BN columns), tcgen05.commit
//                    releases ring slots and publishes finished accumulators;
//   warps 2..9       epilogue (two warps per TMEM lane quarter, each owning half of the tile's columns):
//                    tcgen05.ld (32 lanes x 32 columns per instruction, double buffered), fused
//                    bias / activation / gate*x+residual / per-head RMSNorm+RoPE, bf16 stores.
// The epilogue of tile i overlaps the MMAs of tile i+1 through the double-buffered accumulator.
// Tile order: see tile_coords (row-fastest while A fits L2, column groups beyond that).
#include <cuda.h>
#include <cudaTypedefs.h>

#include <mutex>

#include "flux_ops.cuh"
#include "gemm.cuh"
#include "prof.cuh"
#include "ptx.cuh"

namespace drag {

constexpr int G_BM = 128;
constexpr int G_BK = 64;
// 12 warps = 3 warpgroups: {TMA, MMA, 2 idle} + 8 epilogue warps. The idle warps buy a legal `setmaxnreg`: a 16 K-register
// sub-partition hosts 3 warps, i.e. 168 registers each at launch; the data-movement warpgroup drops to 40 and the two
// epilogue warpgroups take 232, enough to keep a whole 128-wide head row plus prefetched RoPE / bias operands in
// registers (the 10-warp layout was capped at 168 for every warp and spilled as soon as the epilogue prefetched).
constexpr int G_THREADS = 384;
constexpr int G_EPI_WARPS = 8;
constexpr int G_EPI_WARP0 = 4;
constexpr int G_REGS_MOVE = 72, G_REGS_EPI = 216;
template <int N>
__device__ __forceinline__ void g_setmaxnreg_inc() {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N));
}
template <int N>
__device__ __forceinline__ void g_setmaxnreg_dec() {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N));
}

template <int BN>
struct GemmCfg {
    static constexpr int A_BYTES = G_BM * G_BK * 2;
    static constexpr int B_BYTES = BN * G_BK * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int STAGES = (BN == 256) ? 4 : (BN == 128 ? 6 : 8);
    static constexpr int TMEM_COLS = (2 * BN < 32) ? 32 : 2 * BN;  // two accumulator stages
    static constexpr int SMEM = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
};

struct GemmShape {
    int M, N, K;
    int num_m, num_n, num_k;
    int group_m;                                // tile raster: groups of group_m row tiles, see tile_coords
    int group_n;                                // > 0: groups of group_n COLUMN tiles instead (W resident, A streams)
    // Implicit-GEMM convolution over an NHWC activation (conv = 0: plain row-major A). The A tile of 128 output
    // pixels is one 4-D TMA box (64 channels x tw x th pixels) at a tap-dependent offset; image borders are the
    // TMA out-of-bounds zero fill. K runs over taps x channel blocks (weights [C_out][tap][C_in]).
    int conv, cblocks, ksize, stride, pad;      // cblocks = C_in / 64
    int Wo, Ho, tw, th, tiles_x, tiles_y;       // output size, pixel tile, tiles per image row / per image column
};

// Tile raster. Tiles are visited in groups of `group_m` row tiles: inside a group the row tile runs fastest, then the
// column tile, so the CTAs working at the same time cover group_m x (#CTAs / group_m) tiles, the group's A rows
// (group_m x tile rows x K, sized by the host to ~32 MB) stay in L2 while every column tile of W streams past them
// once. With the plain row-fastest order every column tile re-read ALL of A, which outgrows the 126 MB L2 as soon as
// M x K x 2 B does (batched Flux steps, the K = 15360 projection): DRAM traffic of N/256 x |A| instead of |A|.
__device__ __forceinline__ void tile_coords(const GemmShape& sh, int t, int& m_blk, int& n_blk) {
    if (sh.group_n > 0) {
        // column groups: the group's W rows (group_n x BN x K) stay in L2, the column tile runs fastest so that the CTAs
        // working at the same time share each A row tile, and A streams through once per group. Cheaper than row groups
        // when N is small and K large (the N = 3072 projections: |W| + |A| x groups_n  <  |A| + |W| x groups_m).
        const int per_group = sh.group_n * sh.num_m;
        const int g = t / per_group, r = t - g * per_group;
        const int n0 = g * sh.group_n;
        const int gn = min(sh.group_n, sh.num_n - n0);
        m_blk = r / gn;
        n_blk = n0 + (r - m_blk * gn);
        return;
    }
    const int per_group = sh.group_m * sh.num_n;
    const int g = t / per_group, r = t - g * per_group;
    const int m0 = g * sh.group_m;
    const int gm = min(sh.group_m, sh.num_m - m0);
    n_blk = r / gm;
    m_blk = m0 + (r - n_blk * gm);
}

// First output row (linear pixel index) and number of valid rows of 128-row A tile `mt`.
__device__ __forceinline__ void tile_rows(const GemmShape& sh, int mt, int& base, int& count, int& b, int& y0, int& x0) {
    if (!sh.conv) {
        base = mt * 128;
        count = sh.M - base;
        b = y0 = x0 = 0;
        return;
    }
    const int per_img = sh.tiles_x * sh.tiles_y;
    b = mt / per_img;
    const int r = mt - b * per_img;
    const int ty = r / sh.tiles_x, tx = r - ty * sh.tiles_x;
    y0 = ty * sh.th;
    x0 = tx * sh.tw;
    base = (b * sh.Ho + y0) * sh.Wo + x0;       // th > 1 only with tw == Wo: the tile is th full rows, still linear
    count = (sh.th > 1) ? min(128, (sh.Ho - y0) * sh.Wo) : min(sh.tw, sh.Wo - x0);
    if (b * sh.Ho * sh.Wo >= sh.M) count = 0;
}

__device__ __forceinline__ void load_a_tile(const GemmShape& sh, const CUtensorMap* tmA, void* dst, int kb, int m_row,
                                            int b, int y0, int x0, uint64_t* bar) {
    if (!sh.conv) {
        tma_load_2d(dst, tmA, kb * 64, m_row, bar);
    } else {
        const int tap = kb / sh.cblocks, cb = kb - tap * sh.cblocks;
        const int dy = tap / sh.ksize, dx = tap - dy * sh.ksize;
        tma_load_4d(dst, tmA, cb * 64, x0 * sh.stride + dx - sh.pad, y0 * sh.stride + dy - sh.pad, b, bar);
    }
}

// Activations of the epilogues with one MUFU.EX2 and one MUFU.RCP each and no IEEE-division sequence (which costs
// ~10 instructions and a slow-path branch per element and slowed the GELU GEMM by 14 %).
__device__ __forceinline__ float rcp_approx(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float ex2_approx_f(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// gelu_tanh(x) = 0.5 x (1 + tanh(u)), u = sqrt(2/pi) (x + 0.044715 x^3); with e = exp(2u): = x - x / (e + 1).
__device__ __forceinline__ float gelu_tanh_f(float x) {
    const float k1 = 2.f * 0.7978845608028654f * 1.4426950408889634f;      // 2 sqrt(2/pi) log2(e)
    const float k3 = k1 * 0.044715f;
    const float e = ex2_approx_f(x * fmaf(k3, x * x, k1));
    return fmaf(-x, rcp_approx(e + 1.f), x);
}
__device__ __forceinline__ float sigmoid_f(float x) { return rcp_approx(1.f + ex2_approx_f(-1.4426950408889634f * x)); }
// QuickGELU x * sigmoid(1.702 x) = x / (1 + 2^(-1.702 log2(e) x)) with ONE MUFU op per element. The fc GEMM of a ViT block has
// K = 1024: a 128 x 256 tile is 4096 tensor-core cycles, and two MUFU ops per output element (ex2 + rcp, 16 lanes / clock / SM)
// are 4096 cycles as well - the epilogue could not hide behind the next tile's MMAs (ncu: tensor pipe 68 %). The reciprocal of
// d = 1 + 2^t in [1, 2^126] therefore runs on the FMA pipe: exponent-flip seed (5 % off) and two Newton steps, relative error
// 7e-6 - three orders below the bf16 rounding of the stored value. t is clamped so that d stays finite (x < -50: result 0).
__device__ __forceinline__ float rcp_fma(float d) {
    float r = __uint_as_float(0x7EF311C7u - __float_as_uint(d));
    r = r * fmaf(-d, r, 2.f);
    r = r * fmaf(-d, r, 2.f);
    return r;
}
__device__ __forceinline__ float quick_gelu_f(float x) {
    const float e = ex2_approx_f(fminf(-2.4554669595930157f * x, 125.f));     // 1.702 * log2(e)
    return x * rcp_fma(1.f + e);
}
__device__ __forceinline__ float bf16_round(float x) { return __bfloat162float(__float2bfloat16(x)); }

__device__ __forceinline__ void store_bf16x8(__nv_bfloat16* dst, const float* v) {
    __nv_bfloat162 p0 = __floats2bfloat162_rn(v[0], v[1]);
    __nv_bfloat162 p1 = __floats2bfloat162_rn(v[2], v[3]);
    __nv_bfloat162 p2 = __floats2bfloat162_rn(v[4], v[5]);
    __nv_bfloat162 p3 = __floats2bfloat162_rn(v[6], v[7]);
    uint4 u;
    u.x = *reinterpret_cast<uint32_t*>(&p0);
    u.y = *reinterpret_cast<uint32_t*>(&p1);
    u.z = *reinterpret_cast<uint32_t*>(&p2);
    u.w = *reinterpret_cast<uint32_t*>(&p3);
    *reinterpret_cast<uint4*>(dst) = u;
}

// 16 bf16 = one full 32-byte sector per lane in ONE store (STG.256, sm_100). In the row-per-thread epilogue a warp store
// touches 32 different rows: with 16-byte stores every instruction writes 32 HALF sectors and the K = 1024 GEMMs of the ViT
// towers (output bytes large next to their 64 k-steps) were bound by the store path at 41-71 % tensor-pipe utilisation.
__device__ __forceinline__ void store_bf16x16(__nv_bfloat16* dst, const float* v) {
    uint32_t u[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        __nv_bfloat162 p = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
        u[i] = *reinterpret_cast<uint32_t*>(&p);
    }
    asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(dst), "r"(u[0]), "r"(u[1]), "r"(u[2]),
                 "r"(u[3]), "r"(u[4]), "r"(u[5]), "r"(u[6]), "r"(u[7])
                 : "memory");
}
// store 32 consecutive bf16 of one row: 2 x 32 B when the destination allows it, else 4 x 16 B
__device__ __forceinline__ void store_row32(__nv_bfloat16* dst, const float* v, bool wide) {
    if (wide) {
        store_bf16x16(dst, v);
        store_bf16x16(dst + 16, v + 16);
    } else {
#pragma unroll
        for (int j = 0; j < 32; j += 8) store_bf16x8(dst + j, v + j);
    }
}

__device__ __forceinline__ void load_bf16x8(const __nv_bfloat16* src, float* v) {
    uint4 u = *reinterpret_cast<const uint4*>(src);
    const __nv_bfloat162* p = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float2 f = __bfloat1622float2(p[i]);
        v[2 * i] = f.x;
        v[2 * i + 1] = f.y;
    }
}

// 16 bf16 (one 32-byte sector) per lane in one load (LDG.256): the residual rows of the gate*x+residual epilogue.
__device__ __forceinline__ void load_bf16x16(const __nv_bfloat16* src, float* v) {
    uint32_t u[8];
    asm volatile("ld.global.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7])
                 : "l"(src)
                 : "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u[i]));
        v[2 * i] = f.x;
        v[2 * i + 1] = f.y;
    }
}

// Generic epilogue on one 32-column chunk held in registers (this thread = one row).
__device__ __forceinline__ void epilogue_chunk(const GemmEpi& e, float (&v)[32], int row, int col0, int N, float ln_mean,
                                               float ln_rstd, float& st_sum, float& st_sq) {
    if (e.ln_stats) {           // folded LayerNorm of the A rows: rstd * (acc - mean * s[n]) + c[n]
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
            const float4 s4 = __ldg(reinterpret_cast<const float4*>(e.ln_s + col0 + j));
            const float4 c4 = __ldg(reinterpret_cast<const float4*>(e.ln_c + col0 + j));
            v[j] = fmaf(ln_rstd, fmaf(-ln_mean, s4.x, v[j]), c4.x);
            v[j + 1] = fmaf(ln_rstd, fmaf(-ln_mean, s4.y, v[j + 1]), c4.y);
            v[j + 2] = fmaf(ln_rstd, fmaf(-ln_mean, s4.z, v[j + 2]), c4.z);
            v[j + 3] = fmaf(ln_rstd, fmaf(-ln_mean, s4.w, v[j + 3]), c4.w);
        }
    }
    if (e.bias) {
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
            float b[8];
            load_bf16x8(e.bias + col0 + j, b);
#pragma unroll
            for (int t = 0; t < 8; ++t) v[j + t] += b[t];
        }
    }
    if (e.mode == EPI_GELU_TANH) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = gelu_tanh_f(v[j]);
    } else if (e.mode == EPI_QUICK_GELU) {
        if (e.rcp_mufu == 3) {
            // one MUFU.RCP for two elements: 1 / (d0 d1), then s0 = r d1, s1 = r d0 (d = 1 + 2^t, t clamped to 60 so that
            // the product stays finite; x < -24 gives x * 2^-60 instead of x * 2^-2.46|x|: both vanish next to any other term of the next GEMM)
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
                const float d0 = 1.f + ex2_approx_f(fminf(-2.4554669595930157f * v[j], 60.f));
                const float d1 = 1.f + ex2_approx_f(fminf(-2.4554669595930157f * v[j + 1], 60.f));
                const float r = rcp_approx(d0 * d1);
                v[j] = v[j] * (r * d1);
                v[j + 1] = v[j + 1] * (r * d0);
            }
        } else if (e.rcp_mufu == 4) {
            // one MUFU.RCP for four elements (t clamped to 30)
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
                const float d0 = 1.f + ex2_approx_f(fminf(-2.4554669595930157f * v[j], 30.f));
                const float d1 = 1.f + ex2_approx_f(fminf(-2.4554669595930157f * v[j + 1], 30.f));
                const float d2 = 1.f + ex2_approx_f(fminf(-2.4554669595930157f * v[j + 2], 30.f));
                const float d3 = 1.f + ex2_approx_f(fminf(-2.4554669595930157f * v[j + 3], 30.f));
                const float d01 = d0 * d1, d23 = d2 * d3;
                const float r = rcp_approx(d01 * d23);
                const float r01 = r * d23, r23 = r * d01;          // 1 / (d0 d1), 1 / (d2 d3)
                v[j] = v[j] * (r01 * d1);
                v[j + 1] = v[j + 1] * (r01 * d0);
                v[j + 2] = v[j + 2] * (r23 * d3);
                v[j + 3] = v[j + 3] * (r23 * d2);
            }
        } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const bool mufu = e.rcp_mufu == 1 || (e.rcp_mufu == 2 && (j & 3) != 3);      // 2: three of four on MUFU
                v[j] = mufu ? v[j] * sigmoid_f(1.702f * v[j]) : quick_gelu_f(v[j]);
            }
        }
    } else if (e.mode == EPI_SILU) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = v[j] * sigmoid_f(v[j]);
    } else if (e.mode == EPI_GATE_RESID) {
        const int b = row / e.rows_per_batch;
        float rr[32];
        if (e.wide_ld) {
            load_bf16x16(e.resid + static_cast<size_t>(row) * e.ldr + col0, rr);
            load_bf16x16(e.resid + static_cast<size_t>(row) * e.ldr + col0 + 16, rr + 16);
        } else {
#pragma unroll
            for (int j = 0; j < 32; j += 8) load_bf16x8(e.resid + static_cast<size_t>(row) * e.ldr + col0 + j, rr + j);
        }
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
            const float* r = rr + j;
            if (e.gate) {
                float g[8];
                load_bf16x8(e.gate + static_cast<size_t>(b) * e.gate_ld + col0 + j, g);
#pragma unroll
                for (int t = 0; t < 8; ++t) v[j + t] = r[t] + g[t] * bf16_round(v[j + t]);
            } else {
#pragma unroll
                for (int t = 0; t < 8; ++t) v[j + t] = r[t] + bf16_round(v[j + t]);
            }
        }
    }
    if (e.mode == EPI_QKV_SPLIT) {
        // 32-column chunk lies inside one head (head_dim is a multiple of 32)
        const int hw = e.heads * e.head_dim;
        const int which = col0 / hw;
        const int rem = col0 - which * hw;
        const int head = rem / e.head_dim, c = rem - head * e.head_dim;
        const int b = row / e.rows_per_batch;
        const int pos = e.tok_offset + (row - b * e.rows_per_batch);
        __nv_bfloat16* base = which == 0 ? e.q_out : (which == 1 ? e.k_out : e.v_out);
        __nv_bfloat16* dst = base + ((static_cast<size_t>(b) * e.heads + head) * e.s_total + pos) * e.head_dim + c;
        store_row32(dst, v, e.wide_st != 0);
    } else if (e.mode == EPI_BIAS_F32) {
        float* dst = e.out_f32 + static_cast<size_t>(row) * e.ldo + col0;
#pragma unroll
        for (int j = 0; j < 32; j += 4)
            *reinterpret_cast<float4*>(dst + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
    } else {
        __nv_bfloat16* dst = e.out + static_cast<size_t>(row) * e.ldo + col0;
        store_row32(dst, v, e.wide_st != 0);
        if (e.stats_out) {      // moments of the values as stored (bf16), for the LayerNorm folded into the next GEMM
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const float r = bf16_round(v[j]);
                st_sum += r;
                st_sq = fmaf(r, r, st_sq);
            }
        }
    }
}

// Epilogue of one accumulator tile: TMEM -> registers -> fused op -> global. Thread = one output row; the
// 8 epilogue warps are two groups of 4 (one warp per TMEM lane quarter each): group `half` owns columns
// [half*BN/2, (half+1)*BN/2) of the tile, so every SM sub-partition has two warps to hide TMEM / global latency.
// t_addr: TMEM address of this warp's lane quarter at column 0 of the accumulator.
// Folded LayerNorm: mean / rstd of A row `row` from its partial moments. Called BEFORE the accumulator wait so that the (up
// to 8, all in flight at once) loads overlap the tile's MMAs; summed in a fixed order (deterministic).
__device__ __forceinline__ void ln_row_moments(const GemmEpi& epi, int row, bool row_ok, float& ln_mean, float& ln_rstd) {
    ln_mean = 0.f;
    ln_rstd = 1.f;
    if (!epi.ln_stats || !row_ok) return;
    const float2* ps = reinterpret_cast<const float2*>(epi.ln_stats) + static_cast<size_t>(row) * epi.ln_parts;
    float2 t[8];
#pragma unroll
    for (int p = 0; p < 8; ++p) t[p] = (p < epi.ln_parts) ? __ldg(ps + p) : make_float2(0.f, 0.f);
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int p = 0; p < 8; ++p) {
        s1 += t[p].x;
        s2 += t[p].y;
    }
    const float inv_k = 1.f / static_cast<float>(epi.ln_k);
    ln_mean = s1 * inv_k;
    ln_rstd = rsqrtf(fmaxf(s2 * inv_k - ln_mean * ln_mean, 0.f) + epi.ln_eps);
}

// RPRE: the instantiation for the residual epilogue of the short-K GEMMs (ViT out-projection / MLP-down: K <= 2048, a main
// loop of 4096 cycles per tile at K = 1024). Loading the residual where it is used makes that epilogue a chain of BN / 64
// dependent global round trips per tile - longer than the main loop it has to hide behind (ncu: tensor pipe 57 % on the
// out-projection). Here every 32-byte load of this thread's row segment is issued up front, before the first TMEM read, and
// bias + residual (+ gate) is the only epilogue compiled in, which is what makes room for the 16 x BN / 64 extra registers
// (the same prefetch inside the generic epilogue spilled 184-330 bytes in every instantiation, the Flux ones included).
template <int BN, bool RPRE = false>
__device__ __forceinline__ void epilogue_tile(const GemmEpi& epi, const GemmShape& sh, uint32_t t_addr, int row,
                                              bool row_ok, int n_blk, int half, float ln_mean, float ln_rstd) {
    if constexpr (RPRE) {
        constexpr int NC = BN / 64;
        const int c0 = half * (BN / 2);
        const int colbase = n_blk * BN + c0;
        uint32_t rp[NC][16];
        if (row_ok) {
#pragma unroll
            for (int i = 0; i < NC; ++i) {
                if (colbase + i * 32 < sh.N) {
                    const __nv_bfloat16* src = epi.resid + static_cast<size_t>(row) * epi.ldr + colbase + i * 32;
#pragma unroll
                    for (int hh = 0; hh < 2; ++hh)
                        asm volatile("ld.global.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                                     : "=r"(rp[i][hh * 8 + 0]), "=r"(rp[i][hh * 8 + 1]), "=r"(rp[i][hh * 8 + 2]),
                                       "=r"(rp[i][hh * 8 + 3]), "=r"(rp[i][hh * 8 + 4]), "=r"(rp[i][hh * 8 + 5]),
                                       "=r"(rp[i][hh * 8 + 6]), "=r"(rp[i][hh * 8 + 7])
                                     : "l"(src + hh * 16)
                                     : "memory");
                }
            }
        }
        float st_sum = 0.f, st_sq = 0.f;
        uint32_t ra[32], rb[32];
        tmem_ld_32x32(t_addr + c0, ra);
        const int b = row / epi.rows_per_batch;
#pragma unroll
        for (int i = 0; i < NC; ++i) {
            tmem_ld_wait();
            uint32_t (&cur)[32] = (i & 1) ? rb : ra;
            uint32_t (&nxt)[32] = (i & 1) ? ra : rb;
            if (i + 1 < NC) tmem_ld_32x32(t_addr + c0 + (i + 1) * 32, nxt);
            const int col0 = colbase + i * 32;
            if (row_ok && col0 < sh.N) {
                float v[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(cur[j]);
                if (epi.bias) {
#pragma unroll
                    for (int j = 0; j < 32; j += 8) {
                        float bb[8];
                        load_bf16x8(epi.bias + col0 + j, bb);
#pragma unroll
                        for (int t = 0; t < 8; ++t) v[j + t] += bb[t];
                    }
                }
#pragma unroll
                for (int j = 0; j < 32; j += 8) {
                    float g[8];
                    if (epi.gate) load_bf16x8(epi.gate + static_cast<size_t>(b) * epi.gate_ld + col0 + j, g);
#pragma unroll
                    for (int t = 0; t < 8; t += 2) {
                        const float2 r2 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&rp[i][(j + t) >> 1]));
                        v[j + t] = epi.gate ? r2.x + g[t] * bf16_round(v[j + t]) : r2.x + bf16_round(v[j + t]);
                        v[j + t + 1] = epi.gate ? r2.y + g[t + 1] * bf16_round(v[j + t + 1]) : r2.y + bf16_round(v[j + t + 1]);
                    }
                }
                store_row32(epi.out + static_cast<size_t>(row) * epi.ldo + col0, v, true);
                if (epi.stats_out) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const float r = bf16_round(v[j]);
                        st_sum += r;
                        st_sq = fmaf(r, r, st_sq);
                    }
                }
            }
        }
        if (epi.stats_out && row_ok && colbase < sh.N) {
            const int part = colbase / (BN / 2);
            *reinterpret_cast<float2*>(epi.stats_out + (static_cast<size_t>(row) * epi.stats_parts + part) * 2) =
                make_float2(st_sum, st_sq);
        }
        return;
    }
    if (epi.mode == EPI_QKV_ROPE) {
        // BN covers BN/128 whole heads (one per warp group at BN = 256); columns [0,H*128) q, [H*128,2H*128) k, rest v.
        // The whole 128-wide head row lives in registers: one TMEM pass for sum of squares, RMSNorm, RoPE and store.
        constexpr int hd = 128;
        if (BN < 256 && half != 0) return;
        const int h0 = (BN >= 256) ? half * hd : 0;
        const int b = row / epi.rows_per_batch;
        const int pos = epi.tok_offset + (row - b * epi.rows_per_batch);
        const int col_h = n_blk * BN + h0;
        const int which = col_h / (epi.heads * hd);          // 0 q, 1 k, 2 v
        const int head = (col_h - which * epi.heads * hd) / hd;
        __nv_bfloat16* dst_base = (which == 0 ? epi.q_out : (which == 1 ? epi.k_out : epi.v_out));
        // RoPE tables of this row: 2 x 64 fp32 = 32 x 16 B that no other row shares. Loading them where they are used made
        // the epilogue a chain of 16 dependent L2 round trips per tile (tensor pipe 84 % busy on the QKV GEMM vs 98 % on
        // the others), so they are software-pipelined PF iterations ahead, the first PF before the TMEM read.
        constexpr int PF = 6;
        const bool rope = which < 2;
        const float* cs = epi.rope_cos + static_cast<size_t>(pos) * (hd / 2);
        const float* sn = epi.rope_sin + static_cast<size_t>(pos) * (hd / 2);
        float4 cbuf[PF], sbuf[PF];
        if (rope && row_ok) {
#pragma unroll
            for (int p = 0; p < PF; ++p) {
                cbuf[p] = __ldg(reinterpret_cast<const float4*>(cs) + p);
                sbuf[p] = __ldg(reinterpret_cast<const float4*>(sn) + p);
            }
        }
        uint32_t r[hd];
#pragma unroll
        for (int c = 0; c < hd; c += 32) tmem_ld_32x32_ptr(t_addr + h0 + c, &r[c]);
        tmem_ld_wait();
        float* x = reinterpret_cast<float*>(r);      // in place: x[j] = bf16(acc + bias)
        float ss = 0.f;
        uint4 bnext = epi.bias ? __ldg(reinterpret_cast<const uint4*>(epi.bias + col_h)) : make_uint4(0, 0, 0, 0);
#pragma unroll
        for (int j = 0; j < hd; j += 8) {
            const uint4 bcur = bnext;
            if (epi.bias && j + 8 < hd) bnext = __ldg(reinterpret_cast<const uint4*>(epi.bias + col_h + j + 8));
            const __nv_bfloat162* bp = reinterpret_cast<const __nv_bfloat162*>(&bcur);
#pragma unroll
            for (int tt = 0; tt < 4; ++tt) {
                const float2 bf = __bfloat1622float2(bp[tt]);
                const float x0 = bf16_round(__uint_as_float(r[j + 2 * tt]) + bf.x);
                const float x1 = bf16_round(__uint_as_float(r[j + 2 * tt + 1]) + bf.y);
                r[j + 2 * tt] = __float_as_uint(x0);
                r[j + 2 * tt + 1] = __float_as_uint(x1);
                ss = fmaf(x0, x0, ss);
                ss = fmaf(x1, x1, ss);
            }
        }
        if (!row_ok) return;
        __nv_bfloat16* dst = dst_base + ((static_cast<size_t>(b) * epi.heads + head) * epi.s_total + pos) * hd;
        if (rope) {
            const float inv_rms = rsqrtf(ss * (1.f / hd) + epi.rms_eps);
            const __nv_bfloat16* nw = (which == 0) ? epi.q_norm_w : epi.k_norm_w;
            uint4 wnext = __ldg(reinterpret_cast<const uint4*>(nw));
#pragma unroll
            for (int j = 0; j < hd; j += 8) {
                const uint4 wcur = wnext;
                if (j + 8 < hd) wnext = __ldg(reinterpret_cast<const uint4*>(nw + j + 8));
                const __nv_bfloat162* wp = reinterpret_cast<const __nv_bfloat162*>(&wcur);
                const float4 c4 = cbuf[(j / 8) % PF], s4 = sbuf[(j / 8) % PF];
                if (j / 8 + PF < hd / 8) {
                    cbuf[(j / 8) % PF] = __ldg(reinterpret_cast<const float4*>(cs) + j / 8 + PF);
                    sbuf[(j / 8) % PF] = __ldg(reinterpret_cast<const float4*>(sn) + j / 8 + PF);
                }
                const float cc[4] = {c4.x, c4.y, c4.z, c4.w};
                const float sv[4] = {s4.x, s4.y, s4.z, s4.w};
                float o[8];
#pragma unroll
                for (int tt = 0; tt < 4; ++tt) {
                    const float2 w2 = __bfloat1622float2(wp[tt]);
                    const float x0 = bf16_round(bf16_round(x[j + 2 * tt] * inv_rms) * w2.x);
                    const float x1 = bf16_round(bf16_round(x[j + 2 * tt + 1] * inv_rms) * w2.y);
                    o[2 * tt] = x0 * cc[tt] - x1 * sv[tt];
                    o[2 * tt + 1] = x1 * cc[tt] + x0 * sv[tt];
                }
                store_bf16x8(dst + j, o);
            }
        } else {
#pragma unroll
            for (int j = 0; j < hd; j += 8) store_bf16x8(dst + j, &x[j]);
        }
    } else {
        // generic: BN/2 columns per warp group in 32-column chunks, the TMEM load of chunk i+1 in flight while
        // chunk i is processed (tcgen05.wait::ld covers every earlier load of this thread)
        constexpr int NC = BN / 64;
        const int c0 = half * (BN / 2);
        float st_sum = 0.f, st_sq = 0.f;
        uint32_t ra[32], rb[32];
        tmem_ld_32x32(t_addr + c0, ra);
#pragma unroll
        for (int i = 0; i < NC; ++i) {
            tmem_ld_wait();
            uint32_t (&cur)[32] = (i & 1) ? rb : ra;
            uint32_t (&nxt)[32] = (i & 1) ? ra : rb;
            if (i + 1 < NC) tmem_ld_32x32(t_addr + c0 + (i + 1) * 32, nxt);
            const int col0 = n_blk * BN + c0 + i * 32;
            if (row_ok && col0 < sh.N) {
                float v[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(cur[j]);
                epilogue_chunk(epi, v, row, col0, sh.N, ln_mean, ln_rstd, st_sum, st_sq);
            }
        }
        if (epi.stats_out && row_ok && n_blk * BN + c0 < sh.N) {
            const int part = (n_blk * BN + c0) / (BN / 2);
            *reinterpret_cast<float2*>(epi.stats_out + (static_cast<size_t>(row) * epi.stats_parts + part) * 2) =
                make_float2(st_sum, st_sq);
        }
    }
}

template <int BN>
__global__ void __launch_bounds__(G_THREADS, 1)
gemm_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                         GemmShape sh, GemmEpi epi) {
    using Cfg = GemmCfg<BN>;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);   // SWIZZLE_128B tiles: 1024-B aligned
    uint8_t* tiles = smem;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + Cfg::STAGES * Cfg::STAGE_BYTES);
    uint64_t* empty = full + Cfg::STAGES;
    uint64_t* tmem_full = empty + Cfg::STAGES;
    uint64_t* tmem_empty = tmem_full + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int num_tiles = sh.num_m * sh.num_n;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        for (int i = 0; i < Cfg::STAGES; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tmem_full[i], 1);
            mbar_init(&tmem_empty[i], G_EPI_WARPS);   // one arrival per epilogue warp
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

    // The producer and issuer loops run on ALL lanes of their warp with warp-uniform control flow and only the
    // asynchronous instruction itself is predicated on one elected lane: addresses, descriptors and coordinates then
    // live in uniform registers (UTMALDG / UTCHMMA take UR operands) instead of costing an R2UR each, which is what
    // bounds a single-lane issuer at ~70 % tensor occupancy.
    if (warp < G_EPI_WARP0) {
    g_setmaxnreg_dec<G_REGS_MOVE>();
    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer
        int stage = 0, phase = 0;
        for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
            int m_blk, n_blk;
            tile_coords(sh, t, m_blk, n_blk);
            int base, count, ib, iy, ix;
            tile_rows(sh, m_blk, base, count, ib, iy, ix);
            for (int kb = 0; kb < sh.num_k; ++kb) {
                mbar_wait(&empty[stage], phase ^ 1);
                uint8_t* a_dst = tiles + stage * Cfg::STAGE_BYTES;
                uint8_t* b_dst = a_dst + Cfg::A_BYTES;
                if (elect_one()) {
                    mbar_arrive_expect_tx(&full[stage], Cfg::STAGE_BYTES);
                    load_a_tile(sh, &tmA, a_dst, kb, m_blk * G_BM, ib, iy, ix, &full[stage]);
                    tma_load_2d(b_dst, &tmB, kb * G_BK, n_blk * BN, &full[stage]);
                }
                __syncwarp();
                if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer
        constexpr uint32_t idesc = umma_idesc_bf16(G_BM, BN);
        int stage = 0, phase = 0;
        int acc = 0, acc_phase = 0;
        for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
            mbar_wait(&tmem_empty[acc], acc_phase ^ 1);   // epilogue drained this accumulator
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + acc * BN;
            for (int kb = 0; kb < sh.num_k; ++kb) {
                mbar_wait(&full[stage], phase);
                tc_fence_after();
                const uint32_t a_addr = smem_u32(tiles + stage * Cfg::STAGE_BYTES);
                const uint32_t b_addr = a_addr + Cfg::A_BYTES;
                const uint64_t a_desc = umma_desc_k_sw128(a_addr);
                const uint64_t b_desc = umma_desc_k_sw128(b_addr);
                if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < G_BK / 16; ++k) {
                        // advance 16 bf16 = 32 bytes along K inside the swizzle atom: +2 in the addr>>4 field
                        tc_mma_f16(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0);
                    }
                    tc_commit(&empty[stage]);          // slot reusable once these MMAs have read it
                }
                __syncwarp();
                if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
            }
            if (elect_one()) tc_commit(&tmem_full[acc]);            // accumulator complete
            __syncwarp();
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    }
    } else {
        g_setmaxnreg_inc<G_REGS_EPI>();
        // ------------------------------------------------------------------ epilogue
        const int quarter = warp & 3;              // TMEM lane quarter this warp may access
        int acc = 0, acc_phase = 0;
        for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
            int m_blk, n_blk;
            tile_coords(sh, t, m_blk, n_blk);
            int base, count, ib, iy, ix;
            tile_rows(sh, m_blk, base, count, ib, iy, ix);
            const int r_in = quarter * 32 + lane;
            float ln_mean, ln_rstd;
            ln_row_moments(epi, base + r_in, r_in < count, ln_mean, ln_rstd);
            mbar_wait(&tmem_full[acc], acc_phase);
            tc_fence_after();
            const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * BN;
            epilogue_tile<BN>(epi, sh, t_addr, base + r_in, r_in < count, n_blk, (warp - G_EPI_WARP0) >> 2, ln_mean, ln_rstd);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[acc]);
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
    }
}


// ------------------------------------------------------------------------------ CTA-pair variant
// Two CTAs of one cluster (the two SMs of a TPC) compute one 256 x BN tile with tcgen05.mma.cta_group::2:
// each CTA stages ITS 128 rows of A and ITS BN/2 rows of W (32 KB per stage at BN = 256 instead of 48 KB),
// the leader CTA's single issuing thread drives both tensor cores, and each CTA's TMEM receives the 128 x BN
// accumulator of its own rows. Per-SM L2->smem traffic drops by a third (the 1-CTA kernel sits on the L2
// bandwidth cap), shared-memory operand reads by a quarter. Barriers: full[] lives in the leader and counts the
// TMA bytes of both CTAs; empty[] / tmem_full[] exist in both CTAs and are signalled by multicast commits;
// tmem_empty[] lives in the leader and collects the 8 + 8 epilogue warps of the pair.
template <int BN>
struct Gemm2Cfg {
    static constexpr int A_BYTES = G_BM * G_BK * 2;
    static constexpr int B_BYTES = (BN / 2) * G_BK * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int STAGES = (BN == 256) ? 6 : 8;
    static constexpr int TMEM_COLS = 2 * BN;
    static constexpr int SMEM = STAGES * STAGE_BYTES + 1024 + 256;
};

template <int BN, bool RPRE = false>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(G_THREADS, 1)
gemm_bf16_tcgen05_2cta_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                              GemmShape sh, GemmEpi epi) {
    using Cfg = Gemm2Cfg<BN>;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
    uint8_t* tiles = smem;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + Cfg::STAGES * Cfg::STAGE_BYTES);
    uint64_t* empty = full + Cfg::STAGES;
    uint64_t* tmem_full = empty + Cfg::STAGES;
    uint64_t* tmem_empty = tmem_full + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();          // 0 = leader
    const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;
    const int num_tiles = sh.num_m * sh.num_n;        // num_m counts 256-row tiles here

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        for (int i = 0; i < Cfg::STAGES; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tmem_full[i], 1);
            mbar_init(&tmem_empty[i], 2 * G_EPI_WARPS);   // the epilogue warps of both CTAs of the pair
        }
        fence_mbar_init();
    }
    if (warp == 1) {
        tmem_alloc_2cta(tmem_slot, Cfg::TMEM_COLS);
        tmem_relinquish_2cta();
    }
    tc_fence_before();
    cluster_sync_all();                               // barriers of both CTAs initialised, TMEM allocated
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < G_EPI_WARP0) {
    g_setmaxnreg_dec<G_REGS_MOVE>();
    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer (both CTAs; warp-uniform, see above)
        int stage = 0, phase = 0;
        for (int t = pair; t < num_tiles; t += num_pairs) {
            int m_blk, n_blk;
            tile_coords(sh, t, m_blk, n_blk);
            const int a_row = m_blk * (2 * G_BM) + rank * G_BM;
            const int b_row = n_blk * BN + rank * (BN / 2);
            int base, count, ib, iy, ix;
            tile_rows(sh, m_blk * 2 + rank, base, count, ib, iy, ix);
            for (int kb = 0; kb < sh.num_k; ++kb) {
                mbar_wait_cluster(&empty[stage], phase ^ 1);
                uint8_t* a_dst = tiles + stage * Cfg::STAGE_BYTES;
                uint8_t* b_dst = a_dst + Cfg::A_BYTES;
                const uint32_t full_leader = mapa_shared(smem_u32(&full[stage]), 0);
                int c1 = a_row, c2 = 0, c3 = 0, c0 = kb * G_BK;
                if (sh.conv) {
                    const int tap = kb / sh.cblocks, cb = kb - tap * sh.cblocks;
                    const int dy = tap / sh.ksize, dx = tap - dy * sh.ksize;
                    c0 = cb * 64; c1 = ix * sh.stride + dx - sh.pad; c2 = iy * sh.stride + dy - sh.pad; c3 = ib;
                }
                if (elect_one()) {
                    if (rank == 0) mbar_arrive_expect_tx(&full[stage], 2 * Cfg::STAGE_BYTES);
                    if (!sh.conv) tma_load_2d_2cta(a_dst, &tmA, c0, c1, full_leader);
                    else tma_load_4d_2cta(a_dst, &tmA, c0, c1, c2, c3, full_leader);
                    tma_load_2d_2cta(b_dst, &tmB, kb * G_BK, b_row, full_leader);
                }
                __syncwarp();
                if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1 && rank == 0) {
        // ------------------------------------------------------------------ MMA issuer (leader only; warp-uniform)
        constexpr uint32_t idesc = umma_idesc_bf16(2 * G_BM, BN);
        int stage = 0, phase = 0;
        int acc = 0, acc_phase = 0;
        for (int t = pair; t < num_tiles; t += num_pairs) {
            mbar_wait_cluster(&tmem_empty[acc], acc_phase ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + acc * BN;
            for (int kb = 0; kb < sh.num_k; ++kb) {
                mbar_wait_cluster(&full[stage], phase);
                tc_fence_after();
                const uint32_t a_addr = smem_u32(tiles + stage * Cfg::STAGE_BYTES);
                const uint32_t b_addr = a_addr + Cfg::A_BYTES;
                const uint64_t a_desc = umma_desc_k_sw128(a_addr);
                const uint64_t b_desc = umma_desc_k_sw128(b_addr);
                if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < G_BK / 16; ++k)
                        tc_mma_f16_2cta(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0);
                    tc_commit_2cta(&empty[stage], 3);
                }
                __syncwarp();
                if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
            }
            if (elect_one()) tc_commit_2cta(&tmem_full[acc], 3);
            __syncwarp();
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    }
    } else {
        g_setmaxnreg_inc<G_REGS_EPI>();
        // ------------------------------------------------------------------ epilogue (both CTAs, own rows)
        const int quarter = warp & 3;
        int acc = 0, acc_phase = 0;
        for (int t = pair; t < num_tiles; t += num_pairs) {
            int m_blk, n_blk;
            tile_coords(sh, t, m_blk, n_blk);
            int base, count, ib, iy, ix;
            tile_rows(sh, m_blk * 2 + rank, base, count, ib, iy, ix);
            const int r_in = quarter * 32 + lane;
            float ln_mean, ln_rstd;
            ln_row_moments(epi, base + r_in, r_in < count, ln_mean, ln_rstd);
            mbar_wait_cluster(&tmem_full[acc], acc_phase);
            tc_fence_after();
            const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * BN;
            epilogue_tile<BN, RPRE>(epi, sh, t_addr, base + r_in, r_in < count, n_blk, (warp - G_EPI_WARP0) >> 2, ln_mean, ln_rstd);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(mapa_shared(smem_u32(&tmem_empty[acc]), 0));
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    }
    tc_fence_before();
    cluster_sync_all();                               // the pair's MMAs, TMA writes and remote arrivals are done
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc_2cta(tmem_base, Cfg::TMEM_COLS);
    }
}

// ---------------------------------------------------------------------------------------- host
static PFN_cuTensorMapEncodeTiled_v12000 g_encode = nullptr;
static std::once_flag g_encode_once;

static void load_encode() {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
        g_encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
}

// 2-D bf16 row-major [rows][cols] with leading dimension ld (elements); box = 64 cols x box_rows.
int make_tmap_bf16_2d(CUtensorMap* map, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld,
                      uint32_t box_rows) {
    std::call_once(g_encode_once, load_encode);
    if (!g_encode) return fail(DRAG_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {ld * 2};
    cuuint32_t box[2] = {64, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = g_encode(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box,
                          estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(DRAG_ERR_CUDA, "cuTensorMapEncodeTiled failed: " + std::to_string(r));
    return DRAG_OK;
}

// 3-D bf16 tensor [d2][d1][d0] (d0 contiguous); box = box0 x box1 x 1, SWIZZLE_128B (box0 * 2 bytes == 128).
int make_tmap_bf16_3d(CUtensorMap* map, const void* ptr, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t stride1_elems,
                      uint64_t stride2_elems, uint32_t box0, uint32_t box1) {
    std::call_once(g_encode_once, load_encode);
    if (!g_encode) return fail(DRAG_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
    cuuint64_t dims[3] = {d0, d1, d2};
    cuuint64_t strides[2] = {stride1_elems * 2, stride2_elems * 2};
    cuuint32_t box[3] = {box0, box1, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = g_encode(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box,
                          estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(DRAG_ERR_CUDA, "cuTensorMapEncodeTiled(3d) failed: " + std::to_string(r));
    return DRAG_OK;
}

// 4-D bf16 NHWC activation [B][H][W][C]: box = 64 channels x tw x th pixels of one image, traversal stride `stride`
// along W and H (strided convolutions), SWIZZLE_128B, out-of-bounds elements read as zero (= the conv padding).
int make_tmap_bf16_nhwc(CUtensorMap* map, const void* ptr, uint64_t B, uint64_t H, uint64_t W, uint64_t C, uint32_t tw,
                        uint32_t th, uint32_t stride) {
    std::call_once(g_encode_once, load_encode);
    if (!g_encode) return fail(DRAG_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
    cuuint64_t dims[4] = {C, W, H, B};
    cuuint64_t strides[3] = {C * 2, W * C * 2, H * W * C * 2};
    cuuint32_t box[4] = {64, tw * stride, th * stride, 1};
    cuuint32_t estr[4] = {1, stride, stride, 1};
    if (box[1] > 256 || box[2] > 256) return fail(DRAG_ERR_INVALID, "conv2d: pixel tile too large for a TMA box");
    CUresult r = g_encode(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(DRAG_ERR_CUDA, "cuTensorMapEncodeTiled(nhwc) failed: " + std::to_string(r));
    return DRAG_OK;
}

template <int BN>
static int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmShape& sh, const GemmEpi& epi,
                       cudaStream_t st) {
    using Cfg = GemmCfg<BN>;
    static bool attr_set = false;
    if (!attr_set) {
        DRAG_CUDA(cudaFuncSetAttribute(gemm_bf16_tcgen05_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       Cfg::SMEM));
        attr_set = true;
    }
    int sms = device_sm_count();
    if (sms <= 0) sms = 148;
    int tiles = sh.num_m * sh.num_n;
    int grid = tiles < sms ? tiles : sms;
    gemm_bf16_tcgen05_kernel<BN><<<grid, G_THREADS, Cfg::SMEM, st>>>(tmA, tmB, sh, epi); count_launch();
    DRAG_CUDA(cudaGetLastError());
    return DRAG_OK;
}


template <int BN, bool RPRE = false>
static int launch_gemm_2cta(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmShape& sh, const GemmEpi& epi,
                            cudaStream_t st) {
    using Cfg = Gemm2Cfg<BN>;
    static bool attr_set = false;
    if (!attr_set) {
        DRAG_CUDA(cudaFuncSetAttribute(gemm_bf16_tcgen05_2cta_kernel<BN, RPRE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       Cfg::SMEM));
        attr_set = true;
    }
    int sms = device_sm_count();
    if (sms <= 0) sms = 148;
    const int pairs_max = sms / 2;
    const int tiles = sh.num_m * sh.num_n;
    const int pairs = tiles < pairs_max ? tiles : pairs_max;
    gemm_bf16_tcgen05_2cta_kernel<BN, RPRE><<<2 * pairs, G_THREADS, Cfg::SMEM, st>>>(tmA, tmB, sh, epi); count_launch();
    DRAG_CUDA(cudaGetLastError());
    return DRAG_OK;
}


// drag_debug_set key 15: QuickGELU reciprocal: 3 = one MUFU.RCP shared by two elements (default), 4 = by four, 1 = one per
// element, 0 = FMA-pipe Newton iteration, 2 = one element of four on the FMA pipe. The fc GEMM of a ViT block (K = 1024) has a
// 4096-cycle main loop per tile and, at two MUFU ops per element, a 4096-cycle epilogue. MEASURED in the C2 job, same box:
// 0 -> 5342 images/s, 2 -> 5516-5535, 1 -> 5554-5573 (Newton costs more issue slots than the MUFU slots it frees); on another
// box 1 -> 5470-5501, 3 -> 5786-5795: sharing the reciprocal removes a quarter of the MUFU work for two multiplies.
int g_gemm_quick_gelu_mufu = 3;
int g_gemm_resid_prefetch = 1;    // drag_debug_set key 18: 0 = the short-K residual GEMMs use the generic epilogue (A/B comparisons)
int g_gemm_no_wide_st = 0;   // drag_debug_set key 9: 1 = 16-byte epilogue stores only (A/B comparisons)
int g_gemm_force_1cta = 0;   // drag_debug_set key 3: 1 = always use the single-CTA kernel (A/B comparisons)
int g_gemm_group_n = 0;      // drag_debug_set key 6: > 0 = column-group raster with this many column tiles per group
int g_gemm_group_m = 0;      // drag_debug_set key 4: > 0 = force the raster group size (1 << 20 = plain row-fastest order)

// Shared launch logic: builds the operand tensor maps (A from a row-major matrix unless a ready map is given) and
// picks the CTA-pair kernel whenever there is more than one 128-row tile of work and N tiles evenly.
static int dispatch_gemm(const __nv_bfloat16* A, const CUtensorMap* tmA_ready, int M, int K, int lda,
                         const __nv_bfloat16* W, int ldw, GemmShape sh, int bn, int m_tiles, const GemmEpi& epi_in,
                         cudaStream_t st) {
    const int N = sh.N;
    GemmEpi epi = epi_in;
    // 32-byte stores need 32-byte aligned row segments: base pointers and row pitches (32-column chunks start at multiples of 64 B)
    if (epi.mode == EPI_QKV_SPLIT)
        epi.wide_st = ((reinterpret_cast<uintptr_t>(epi.q_out) | reinterpret_cast<uintptr_t>(epi.k_out) |
                        reinterpret_cast<uintptr_t>(epi.v_out)) & 31) == 0 && (epi.head_dim % 16) == 0;
    else if (epi.out)
        epi.wide_st = (reinterpret_cast<uintptr_t>(epi.out) & 31) == 0 && (epi.ldo % 16) == 0;
    epi.wide_ld = epi.resid && (reinterpret_cast<uintptr_t>(epi.resid) & 31) == 0 && (epi.ldr % 16) == 0;
    // Measured (profiles/r02_gemm_epilogue_store_width.txt): 32-byte stores lift the store-bound K = 1024 shapes of the ViT
    // blocks by 12-35 % (QKV 1109 -> 1495 TFLOP/s) and do nothing for K >= 3072, where the epilogue hides behind 48+ k-steps
    // (Flux MLP-up 1486 vs 1405): wide accesses only where the k loop is short.
    if (K > 2048) epi.wide_st = epi.wide_ld = 0;
    if (g_gemm_no_wide_st) epi.wide_st = epi.wide_ld = 0;
    epi.rcp_mufu = g_gemm_quick_gelu_mufu;
    const bool pair_ok = !g_gemm_force_1cta && m_tiles > 1 && (bn == 256 || bn == 128) && N % bn == 0;
    CUtensorMap tmA, tmB;
    int rc;
    if (tmA_ready) tmA = *tmA_ready;
    else if ((rc = make_tmap_bf16_2d(&tmA, A, M, K, lda, G_BM))) return rc;
    const int slot = prof_begin(PROF_GEMM, 2.0 * M * static_cast<double>(N) * K, st);
    // Tile raster (tile_coords). While all of A fits comfortably in L2 (<= 48 MB) the plain row-fastest order is best: W
    // streams once, A stays resident (1345 vs 1299 TFLOP/s at M = 5337, K = 3072). Beyond that, COLUMN groups: up to 24
    // column tiles whose W rows take <= ~96 MB stay L2-resident (they are re-touched by every wave), the column tile
    // runs fastest so concurrent CTAs share each A row tile, and A streams through once per group. Measured at M = 21348
    // (profiles/r01_gemm_raster*): row-fastest 1083, row groups 1318 / 1255 / 1271, column groups 1355 / 1323 / 1312
    // TFLOP/s at (N, K) = (12288, 3072) / (3072, 15360) / (3072, 12288). L2 eviction hints (A evict-first, W evict-last)
    // made every shape slower (1163-1264) and are not used.
    const long long tile_bytes = static_cast<long long>(pair_ok ? 2 * G_BM : G_BM) * K * 2;
    const int num_m_tiles = pair_ok ? ceil_div(m_tiles, 2) : m_tiles;
    const int num_n_tiles = pair_ok ? N / bn : ceil_div(N, bn);
    int gm = num_m_tiles, gn = 0;
    if (tile_bytes * num_m_tiles > (48ll << 20)) {
        const long long col_bytes = static_cast<long long>(bn) * K * 2;
        int gmax = static_cast<int>((96ll << 20) / (col_bytes > 0 ? col_bytes : 1));
        gmax = gmax < 1 ? 1 : (gmax > 24 ? 24 : gmax);
        gn = ceil_div(num_n_tiles, ceil_div(num_n_tiles, gmax));
    }
    if (g_gemm_group_m > 0) { gm = g_gemm_group_m; gn = 0; }
    if (g_gemm_group_n > 0) gn = g_gemm_group_n;
    sh.group_m = gm;
    sh.group_n = gn;
    if (pair_ok) {
        sh.num_m = ceil_div(m_tiles, 2);
        sh.num_n = N / bn;
        if ((rc = make_tmap_bf16_2d(&tmB, W, N, K, ldw, bn / 2))) return rc;
        const bool rpre = bn == 256 && epi.mode == EPI_GATE_RESID && epi.wide_ld && epi.wide_st && !epi.ln_stats &&
                          g_gemm_resid_prefetch;      // (wide accesses are only granted for K <= 2048, see above)
        rc = rpre ? launch_gemm_2cta<256, true>(tmA, tmB, sh, epi, st)
                  : (bn == 256) ? launch_gemm_2cta<256>(tmA, tmB, sh, epi, st) : launch_gemm_2cta<128>(tmA, tmB, sh, epi, st);
    } else {
        sh.num_m = m_tiles;
        sh.num_n = ceil_div(N, bn);
        if ((rc = make_tmap_bf16_2d(&tmB, W, N, K, ldw, bn))) return rc;
        switch (bn) {
            case 256: rc = launch_gemm<256>(tmA, tmB, sh, epi, st); break;
            case 128: rc = launch_gemm<128>(tmA, tmB, sh, epi, st); break;
            default:  rc = launch_gemm<64>(tmA, tmB, sh, epi, st); break;
        }
    }
    prof_end(slot, st);
    return rc;
}

int gemm_bf16(const __nv_bfloat16* A, int lda, const __nv_bfloat16* W, int ldw, int M, int N, int K,
              const GemmEpi& epi, cudaStream_t st) {
    DRAG_REQUIRE(A && W, "gemm: null operand");
    DRAG_REQUIRE(M >= 1 && N >= 1 && K >= 1, "gemm: empty problem");
    DRAG_REQUIRE(K % 8 == 0 && lda % 8 == 0 && ldw % 8 == 0, "gemm: K, lda, ldw must be multiples of 8");
    DRAG_REQUIRE((reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(W) & 15) == 0,
                 "gemm: operands must be 16-byte aligned");
    DRAG_REQUIRE(N % 32 == 0, "gemm: N must be a multiple of 32");
    int bn = 256;
    if (epi.mode == EPI_QKV_ROPE) {
        DRAG_REQUIRE(N == 3 * epi.heads * 128, "gemm qkv epilogue: N must be 3*heads*128");
        DRAG_REQUIRE(epi.q_out && epi.k_out && epi.v_out && epi.rope_cos && epi.rope_sin && epi.q_norm_w &&
                         epi.k_norm_w, "gemm qkv epilogue: null pointer");
        bn = (N % 256 == 0) ? 256 : 128;
    } else if (epi.mode == EPI_QKV_SPLIT) {
        DRAG_REQUIRE((epi.head_dim == 64 || epi.head_dim == 128) && N == 3 * epi.heads * epi.head_dim,
                     "gemm qkv-split epilogue: N must be 3*heads*head_dim");
        DRAG_REQUIRE(epi.q_out && epi.k_out && epi.v_out, "gemm qkv-split epilogue: null pointer");
        if (N % 256 != 0 || N <= 256) bn = (N % 128 == 0 && N > 128) ? 128 : 64;
    } else {
        DRAG_REQUIRE(epi.out || epi.out_f32, "gemm: null output");
        if (N % 256 != 0 || N <= 256) bn = (N % 128 == 0 && N > 128) ? 128 : 64;
        if (N % bn != 0 && N > bn) bn = 64;   // N % 32 == 0: tail columns masked per 32-column chunk
    }
    if (epi.stats_out) {
        DRAG_REQUIRE(epi.mode == EPI_BIAS || epi.mode == EPI_GATE_RESID, "gemm: row statistics need a bf16 row-major output");
        DRAG_REQUIRE(N % (bn / 2) == 0 && epi.stats_parts == N / (bn / 2),
                     "gemm: stats_parts must equal N / (BN/2) = " + std::to_string(N / (bn / 2)));
    }
    if (epi.ln_stats) {
        DRAG_REQUIRE(epi.mode != EPI_QKV_ROPE, "gemm: folded LayerNorm is not available with the RoPE epilogue");
        DRAG_REQUIRE(epi.ln_s && epi.ln_c && epi.ln_parts >= 1 && epi.ln_parts <= 8 && epi.ln_k == K && !epi.bias,
                     "gemm: folded LayerNorm needs ln_s, ln_c, ln_parts, ln_k == K and no separate bias");
    }
    GemmShape sh{};
    sh.M = M; sh.N = N; sh.K = K;
    sh.num_k = ceil_div(K, G_BK);
    return dispatch_gemm(A, nullptr, M, K, lda, W, ldw, sh, bn, ceil_div(M, G_BM), epi, st);
}

// Implicit-GEMM convolution: out[b][y][x][:] = epilogue(sum_{tap,c} in[b][y*s+dy-pad][x*s+dx-pad][c] * w[:][tap][c]).
// in bf16 NHWC [B][H][W][C_in] (C_in % 64 == 0), w bf16 [C_out][ksize*ksize][C_in], out NHWC [B][Ho][Wo][C_out]
// written through the usual epilogues with row = linear output pixel. ksize 1 or 3; stride 1 or 2; pad = left/top
// padding (right/bottom are the zero fill of whatever the window reaches past the border).
int conv2d_nhwc_bf16(const __nv_bfloat16* in, int B, int H, int W, int C_in, const __nv_bfloat16* w, int C_out,
                     int ksize, int stride, int pad, int Ho, int Wo, const GemmEpi& epi, cudaStream_t st) {
    DRAG_REQUIRE(in && w, "conv2d: null operand");
    DRAG_REQUIRE(B >= 1 && H >= 1 && W >= 1 && Ho >= 1 && Wo >= 1, "conv2d: empty problem");
    DRAG_REQUIRE(C_in % 64 == 0 && C_out % 32 == 0, "conv2d: C_in must be a multiple of 64, C_out of 32");
    DRAG_REQUIRE((ksize == 1 || ksize == 3) && (stride == 1 || stride == 2), "conv2d: ksize 1|3, stride 1|2");
    DRAG_REQUIRE(epi.mode != EPI_QKV_ROPE && epi.mode != EPI_QKV_SPLIT, "conv2d: unsupported epilogue");
    DRAG_REQUIRE(epi.out || epi.out_f32, "conv2d: null output");
    GemmShape sh{};
    sh.M = B * Ho * Wo; sh.N = C_out; sh.K = ksize * ksize * C_in;
    sh.num_k = sh.K / G_BK;
    sh.conv = 1; sh.cblocks = C_in / 64; sh.ksize = ksize; sh.stride = stride; sh.pad = pad;
    sh.Wo = Wo; sh.Ho = Ho;
    if (Wo < 128 && 128 % Wo == 0) {           // th full rows per tile
        sh.tw = Wo; sh.th = 128 / Wo; sh.tiles_x = 1; sh.tiles_y = ceil_div(Ho, sh.th);
    } else {                                     // one row segment of up to 128 pixels per tile
        sh.tw = 128; sh.th = 1; sh.tiles_x = ceil_div(Wo, 128); sh.tiles_y = Ho;
    }
    const int m_tiles = B * sh.tiles_x * sh.tiles_y;
    int bn = 256;
    if (C_out % 256 != 0 || C_out <= 256) bn = (C_out % 128 == 0 && C_out > 128) ? 128 : 64;
    if (C_out % bn != 0 && C_out > bn) bn = 64;
    CUtensorMap tmA;
    int rc = make_tmap_bf16_nhwc(&tmA, in, B, H, W, C_in, sh.tw, sh.th, stride);
    if (rc) return rc;
    return dispatch_gemm(nullptr, &tmA, sh.M, sh.K, 0, w, sh.K, sh, bn, m_tiles, epi, st);
}

}  // namespace drag
