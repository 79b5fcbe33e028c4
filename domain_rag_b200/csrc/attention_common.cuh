// Shared by the attention kernels (attention_tcgen05.cu: tiled online-softmax kernels; attention_row.cu: whole-row kernels
// for short head-dim-64 sequences): tile constants, the argument block and the softmax arithmetic helpers.
#pragma once
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

// Whole-row kernels for head dim 64 and at most 257 keys (attention_row.cu); picks the persistent or the one-tile form.
bool attention_row_eligible(int head_dim, int S);
int launch_attention_row_any(const __nv_bfloat16* q, const __nv_bfloat16* k, const __nv_bfloat16* v, int B, int H, int S,
                             AttnArgs a, cudaStream_t st);

}  // namespace drag
