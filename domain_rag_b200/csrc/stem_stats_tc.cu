// ResNet-50 stem + style statistics on the 5th-gen tensor cores (retrieval/clip100_resnet_style_all_shots.py:51-74, 180-203):
// conv 7x7 s2 p3 (eval BatchNorm folded) -> ReLU -> maxpool 3x3 s2 p1 -> per-channel mean / sqrt(unbiased var + eps).
//
// The CUDA-core kernel (stem_stats.cu) is bound by the FP32 pipe: 308 MFLOP per image = 8.2 ms for the 707 images of a C2
// re-rank batch, 1 % of what HBM would allow. Here the convolution is an implicit GEMM with fp32-class accuracy from
// split-bf16 operands:   x = xh + xl, w = wh + wl (bf16 each);   x.w ~= xh.wh + xh.wl + xl.wh   (fp32 accumulate in TMEM;
// the dropped xl.wl term is 2^-16 relative), three UMMAs per k-step.
//
// One persistent CTA per SM walks images; per image one conv row (128 pixels = the UMMA M dimension) at a time:
//   D[cy][128 px][64 ch] = sum over the 7 input rows y = 2cy-3+ky of  X[y][128 px][32] . W[ky][64 ch][32]^T
//   with the per-input-row K block k = ci*8 + kx (kx = 7 and ci = 3 are zero padding): the im2col block of an input row is
//   built ONCE in shared memory (hi and lo tiles) and reused by the up to four conv rows that touch it. Two input rows
//   share one 128-byte-swizzled [128][64] K-major tile (the layout TMA would write), ring of 5 row pairs.
//   warps 0-3   loaders (thread = pixel): 7-tap windows from global/L1 (all loads of a row pair in flight at once), split to
//               bf16 hi/lo, swizzled 16-byte stores, fence.proxy.async, arrive on the pair's `xfull` barrier. The folded
//               BatchNorm bias rides in the GEMM: the padding tap (kx = 7) of channel 0 is a column of ones in A and holds the
//               bias in the ky = 3 weight block, so every conv row receives it exactly once;
//   warp 4      MMA issuer: 7 x 2 k-steps x 3 split terms = 42 UMMA 128x64x16 per conv row into one of 8 TMEM accumulators,
//               tcgen05.commit -> `dfull`; releases row pairs (`xempty`) as conv rows retire;
//   warps 8-15  pooling + statistics, two warpgroups splitting the CHANNELS (0-31 / 32-63; the profile of the one-warpgroup
//               version showed this stage, not the tensor core, bounding the kernel): thread = pixel. For pooled row py the
//               three conv rows 2py-1..2py+1 are read from TMEM (vertical max), the horizontal 3-max comes from warp shuffles
//               (+ a shared-memory hand-off at warp edges), ReLU after the max (monotonic), then shifted running moments per
//               channel: even lanes own the first 16 channels of their pooled pixel, odd lanes the other 16. fp64 merge.
// Neither the 64x128x128 conv map nor the 64x64x64 pooled map ever leaves the SM. Input: fp32 [B][3][256][256] in [0,1] (the
// reference contract) or uint8 pixels (scaled by 1/255 in the loader).
#include "common.cuh"
#include "index.cuh"
#include "ptx.cuh"

namespace drag {

constexpr int TS_THREADS = 512;                 // 16 warps: 0-3 loaders, 4 MMA, 5-7 idle, 8-11 / 12-15 pooling (channels
                                                // 0-31 / 32-63)
constexpr int TS_RING_MAX = 10;                 // input-row pairs resident: 5 (fp32 input: hi + lo tile per pair) or 10 (uint8 input:
                                                // the pixel VALUES 0..255 are exact in bf16 - one tile per pair, 1/255 folded into W)
constexpr int TS_XTILE = 128 * 128;             // [128 px][64 K] bf16 = 16 KB
constexpr int TS_WTILE = 64 * 128;              // [64 ch][64 K] bf16 = 8 KB (two ky per tile)
constexpr int TS_DBUF = 8;                      // TMEM accumulators (64 columns each)
constexpr int TS_X_OFF = 0;                                         // ring: [slot][hi|lo]
constexpr int TS_W_OFF = TS_X_OFF + 5 * 2 * TS_XTILE;               // [hi|lo][4 ky pairs]
constexpr int TS_EDGE_OFF = TS_W_OFF + 2 * 4 * TS_WTILE;            // [2 parity][4 warps][64] fp32 warp-edge hand-off; the
                                                                    // end-of-image reduction buffer [4 warps][2][64] aliases it
constexpr int TS_BAR_OFF = TS_EDGE_OFF + 2 * 4 * 64 * 4;
constexpr int TS_SMEM = TS_BAR_OFF + 320;                           // 231 744 B of the 232 448 B a CTA may own
static_assert(TS_SMEM <= 232448, "stem_stats_tc: shared memory budget");

// byte offset of element (row r, k) of a K-major SWIZZLE_128B tile (rows of 64 bf16 = 128 B, 8-row atoms of 1024 B)
__device__ __forceinline__ uint32_t sw128_chunk(int r, int chunk16) {
    return static_cast<uint32_t>((r >> 3) * 1024 + (r & 7) * 128 + ((chunk16 ^ (r & 7)) << 4));
}

__device__ __forceinline__ void split_bf16x8(const float (&x)[8], uint4& hi, uint4& lo) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const __nv_bfloat16 h0 = __float2bfloat16(x[2 * i]), h1 = __float2bfloat16(x[2 * i + 1]);
        const __nv_bfloat16 l0 = __float2bfloat16(x[2 * i] - __bfloat162float(h0));
        const __nv_bfloat16 l1 = __float2bfloat16(x[2 * i + 1] - __bfloat162float(h1));
        h[i] = static_cast<uint32_t>(__bfloat16_as_ushort(h0)) | (static_cast<uint32_t>(__bfloat16_as_ushort(h1)) << 16);
        l[i] = static_cast<uint32_t>(__bfloat16_as_ushort(l0)) | (static_cast<uint32_t>(__bfloat16_as_ushort(l1)) << 16);
    }
    hi = make_uint4(h[0], h[1], h[2], h[3]);
    lo = make_uint4(l[0], l[1], l[2], l[3]);
}

__device__ __forceinline__ void named_bar_sync(int id, int threads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}
__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t* v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}

template <bool U8>
__global__ void __launch_bounds__(TS_THREADS, 1)
stem_stats_tc_kernel(const void* __restrict__ img_v, const float* __restrict__ w_fold, const float* __restrict__ b_fold,
                     float eps, float* __restrict__ out, int B) {
    extern __shared__ __align__(1024) uint8_t smem[];          // SWIZZLE_128B tiles need 1024-byte aligned bases
    uint8_t* xring = smem + TS_X_OFF;
    uint8_t* wt = smem + TS_W_OFF;
    float* edge = reinterpret_cast<float*>(smem + TS_EDGE_OFF);        // [2][4][64]
    float* red = edge;                                                 // [4 warps][sum | sumsq][64 ch], end of image only
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + TS_BAR_OFF);
    constexpr int TS_RING = U8 ? 10 : 5;     // ring slots
    constexpr int TS_TILES = U8 ? 1 : 2;     // operand tiles per slot (hi [, lo])
    uint64_t* xfull = bars;                      // [ring]  count 128 (loader threads)
    uint64_t* xempty = bars + TS_RING_MAX;       // [ring]  count 1 (tcgen05.commit)
    uint64_t* dfull = bars + 2 * TS_RING_MAX;    // [8]  count 1 (tcgen05.commit)
    uint64_t* dempty = dfull + TS_DBUF;      // [8]  count 8 (one per pooling warp)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(dempty + TS_DBUF);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n_img = (B - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);

    if (tid == 0) {
        if (smem_u32(smem) & 1023u) __trap();                  // swizzled tiles would be read from the wrong banks: fail loudly
        for (int i = 0; i < TS_RING; ++i) {
            mbar_init(&xfull[i], 128);
            mbar_init(&xempty[i], 1);
        }
        for (int i = 0; i < TS_DBUF; ++i) {
            mbar_init(&dfull[i], 1);
            mbar_init(&dempty[i], 8);                        // every warp of both pooling warpgroups
        }
        fence_mbar_init();
    }
    if (warp == 4) {
        tmem_alloc(tmem_slot, 512);
        tmem_relinquish();
    }
    // Weight tiles (persistent across images): tile t holds ky = 2t in K [0,32) and ky = 2t+1 in K [32,64); k = ci*8 + kx.
    for (int i = tid; i < 4 * 64 * 8; i += TS_THREADS) {       // (tile, channel row, 16-byte chunk)
        const int t = i >> 9, ch = (i >> 3) & 63, c = i & 7;
        const int ky = 2 * t + (c >> 2), ci = c & 3;
        float x[8];
#pragma unroll
        for (int kx = 0; kx < 8; ++kx) {
            x[kx] = (ky < 7 && ci < 3 && kx < 7) ? w_fold[((ch * 3 + ci) * 7 + ky) * 7 + kx] : 0.f;
            if (U8) x[kx] = __fdiv_rn(x[kx], 255.f);          // uint8 input: A holds the pixel values, the / 255 of :193 lives here
        }
        if (ky == 3 && ci == 0) x[7] = b_fold[ch];           // multiplies the column of ones the loaders put at (ci 0, kx 7)
        uint4 hi, lo;
        split_bf16x8(x, hi, lo);
        const uint32_t off = static_cast<uint32_t>(t) * TS_WTILE + sw128_chunk(ch, c);
        *reinterpret_cast<uint4*>(wt + off) = hi;
        *reinterpret_cast<uint4*>(wt + 4 * TS_WTILE + off) = lo;
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < 8) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 96;");      // 256 threads release 32 each; the 256 pooling threads take 32 each
    if (warp < 4) {
        // ------------------------------------------------------------------ loaders: thread = conv pixel px
        const int px = tid;
        // the zero chunks (ci = 3 of either row) of every ring slot are written once: no later store touches them
        for (int slot = 0; slot < TS_RING * TS_TILES; ++slot) {
            uint8_t* t0 = xring + static_cast<size_t>(slot) * TS_XTILE;
            *reinterpret_cast<uint4*>(t0 + sw128_chunk(px, 3)) = make_uint4(0, 0, 0, 0);
            *reinterpret_cast<uint4*>(t0 + sw128_chunk(px, 7)) = make_uint4(0, 0, 0, 0);
        }
        // window of conv pixel px: input columns 2px-3 .. 2px+3, fetched as the 4 aligned pairs covering 2px-4 .. 2px+3
        // (rows have 256 columns, so a pair is either wholly inside the row or wholly padding)
        const int x0 = 2 * px - 4;
        const bool ok0 = x0 >= 0, ok3 = x0 + 6 < 256;       // pairs 1, 2 are always inside
        for (int it = 0; it < n_img; ++it) {
            const size_t b = static_cast<size_t>(blockIdx.x) + static_cast<size_t>(it) * gridDim.x;
            for (int p = 0; p < 128; ++p) {
                const int P = it * 128 + p, slot = P % TS_RING;
                // all 24 loads of the pair in flight before the first conversion: ONE memory latency per row pair
                if (U8) {
                    // uint8 pixels: 4 x 16-bit loads per (row, channel) window; the VALUES go to the tensor core as exact bf16
                    unsigned short q[6][4];
#pragma unroll
                    for (int c = 0; c < 6; ++c) {
                        const int h = c / 3, ci = c - h * 3;
                        const uint8_t* src = static_cast<const uint8_t*>(img_v) + ((b * 3 + ci) * 256 + (2 * p + h)) * 256;
                        q[c][0] = ok0 ? __ldg(reinterpret_cast<const unsigned short*>(src + x0)) : static_cast<unsigned short>(0);
                        q[c][1] = (x0 + 2 >= 0) ? __ldg(reinterpret_cast<const unsigned short*>(src + x0 + 2)) : static_cast<unsigned short>(0);
                        q[c][2] = __ldg(reinterpret_cast<const unsigned short*>(src + x0 + 4));
                        q[c][3] = ok3 ? __ldg(reinterpret_cast<const unsigned short*>(src + x0 + 6)) : static_cast<unsigned short>(0);
                    }
                    mbar_wait(&xempty[slot], ((P / TS_RING) & 1) ^ 1);
                    uint8_t* thi = xring + static_cast<size_t>(slot) * TS_XTILE;
#pragma unroll
                    for (int c = 0; c < 6; ++c) {
                        const int h = c / 3, ci = c - h * 3;
                        // window bytes 1..7 are taps kx = 0..6; integers 0..255 convert to bf16 exactly
                        float f[8];
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            f[2 * j] = static_cast<float>(q[c][j] & 0xff);
                            f[2 * j + 1] = static_cast<float>(q[c][j] >> 8);
                        }
                        uint32_t u[4];
#pragma unroll
                        for (int j = 0; j < 3; ++j) {
                            __nv_bfloat162 pk = __floats2bfloat162_rn(f[2 * j + 1], f[2 * j + 2]);
                            u[j] = *reinterpret_cast<uint32_t*>(&pk);
                        }
                        __nv_bfloat162 pl = __floats2bfloat162_rn(f[7], ci == 0 ? 1.f : 0.f);   // tap 6 | the ones column (bias)
                        u[3] = *reinterpret_cast<uint32_t*>(&pl);
                        *reinterpret_cast<uint4*>(thi + sw128_chunk(px, h * 4 + ci)) = make_uint4(u[0], u[1], u[2], u[3]);
                    }
                } else {
                float w[6][8];
#pragma unroll
                for (int c = 0; c < 6; ++c) {                 // (row of the pair, channel)
                    const int h = c / 3, ci = c - h * 3;
                    const float* src = static_cast<const float*>(img_v) + ((b * 3 + ci) * 256 + (2 * p + h)) * 256;
                    const float2 z = make_float2(0.f, 0.f);
                    const float2 a0 = ok0 ? __ldg(reinterpret_cast<const float2*>(src + x0)) : z;
                    const float2 a1 = (x0 + 2 >= 0) ? __ldg(reinterpret_cast<const float2*>(src + x0 + 2)) : z;
                    const float2 a2 = __ldg(reinterpret_cast<const float2*>(src + x0 + 4));
                    const float2 a3 = ok3 ? __ldg(reinterpret_cast<const float2*>(src + x0 + 6)) : z;
                    w[c][0] = a0.x; w[c][1] = a0.y; w[c][2] = a1.x; w[c][3] = a1.y;
                    w[c][4] = a2.x; w[c][5] = a2.y; w[c][6] = a3.x; w[c][7] = a3.y;
                }
                mbar_wait(&xempty[slot], ((P / TS_RING) & 1) ^ 1);
                uint8_t* thi = xring + static_cast<size_t>(slot) * 2 * TS_XTILE;
                uint8_t* tlo = thi + TS_XTILE;
#pragma unroll
                for (int c = 0; c < 6; ++c) {
                    const int h = c / 3, ci = c - h * 3;
                    float x[8];                               // taps kx = 0..6 are window elements 1..7; kx = 7 is padding
#pragma unroll
                    for (int kx = 0; kx < 7; ++kx) x[kx] = w[c][kx + 1];
                    x[7] = (ci == 0) ? 1.f : 0.f;             // the ones column that carries the bias (weights: ky 3, ci 0, kx 7)
                    uint4 hi, lo;
                    split_bf16x8(x, hi, lo);
                    const uint32_t off = sw128_chunk(px, h * 4 + ci);
                    *reinterpret_cast<uint4*>(thi + off) = hi;
                    *reinterpret_cast<uint4*>(tlo + off) = lo;
                }
                }
                fence_proxy_async();                          // generic-proxy stores -> visible to the tensor core's async proxy
                mbar_arrive(&xfull[slot]);
            }
        }
    } else if (warp == 4) {
        // ------------------------------------------------------------------ MMA issuer
        // Everything the issue loop needs is kept in running counters updated with add / compare / select only (no % or /),
        // and the tap loop is fully unrolled, so that ring slots, phases and UMMA descriptors live in UNIFORM registers:
        // descriptor arithmetic in vector registers costs an R2UR per operand and made this loop, not the tensor core, the
        // bottleneck (148 cycles per 32-cycle MMA in the first version).
        constexpr uint32_t idesc = umma_idesc_bf16(128, 64, 0, 0);
        const uint64_t x_desc0 = umma_desc_k_sw128(smem_u32(xring)), w_desc0 = umma_desc_k_sw128(smem_u32(wt));
        int s0 = TS_RING - 2;                                  // ring slot of pair (cy - 2); pairs run on across images (128 % ring
                                                               // != 0 is fine: slots are just a running counter mod ring)
        int w_slot = 0, w_phase = 0;                           // next pair to wait for: its slot and barrier phase
        int buf = 0, d_phase = 0;                              // accumulator of the current conv row and its phase
        for (int it = 0; it < n_img; ++it) {
            int pairs_ready = -1;                              // highest pair of this image already waited for
            for (int cy = 0; cy < 128; ++cy) {
                mbar_wait(&dempty[buf], d_phase ^ 1);
                const int p_hi = (cy + 1 < 127) ? cy + 1 : 127;
                while (pairs_ready < p_hi) {
                    ++pairs_ready;
                    mbar_wait(&xfull[w_slot], w_phase);
                    if (++w_slot == TS_RING) { w_slot = 0; w_phase ^= 1; }
                }
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + buf * 64;
                uint32_t acc = 0;
#pragma unroll
                for (int ky = 0; ky < 7; ++ky) {
                    // input row y = 2cy - 3 + ky lives in pair (cy - 2) + (ky + 1) / 2, half (ky + 1) & 1
                    const bool valid = (2 * cy - 3 + ky >= 0) && (2 * cy - 3 + ky <= 255);   // padding rows contribute nothing
                    int slot = s0 + ((ky + 1) >> 1);
                    slot = slot >= TS_RING ? slot - TS_RING : slot;
                    const uint64_t ah = x_desc0 + static_cast<uint64_t>((slot * TS_TILES * TS_XTILE + ((ky + 1) & 1) * 64) >> 4);
                    const uint64_t al = ah + (TS_XTILE >> 4);     // lo tile: fp32 input only
                    const uint64_t bh = w_desc0 + (((ky >> 1) * TS_WTILE + (ky & 1) * 64) >> 4);
                    const uint64_t bl = bh + ((4 * TS_WTILE) >> 4);
                    if (valid) {
                        if (elect_one()) {
                            tc_mma_f16(d_tmem, ah, bh, idesc, acc);
                            tc_mma_f16(d_tmem, ah, bl, idesc, 1);
                            if (!U8) tc_mma_f16(d_tmem, al, bh, idesc, 1);     // uint8 pixels are exact in bf16: no lo operand
                            tc_mma_f16(d_tmem, ah + 2, bh + 2, idesc, 1);      // second k-step: +32 bytes along K
                            tc_mma_f16(d_tmem, ah + 2, bl + 2, idesc, 1);
                            if (!U8) tc_mma_f16(d_tmem, al + 2, bh + 2, idesc, 1);
                        }
                        __syncwarp();
                        acc = 1;
                    }
                }
                if (elect_one()) {
                    tc_commit(&dfull[buf]);
                    // conv row cy was the last user of pair cy-2 (slot s0); the image's last row also retires pairs 126, 127
                    if (cy >= 2) tc_commit(&xempty[s0]);
                    if (cy == 127) {
                        const int s1 = (s0 + 1 >= TS_RING) ? s0 + 1 - TS_RING : s0 + 1;
                        const int s2 = (s0 + 2 >= TS_RING) ? s0 + 2 - TS_RING : s0 + 2;
                        tc_commit(&xempty[s1]);
                        tc_commit(&xempty[s2]);
                    }
                }
                __syncwarp();
                if (++s0 == TS_RING) s0 = 0;
                if (++buf == TS_DBUF) { buf = 0; d_phase ^= 1; }
            }
        }
    }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 160;");
        // ------------------------------------------------------------------ pooling + statistics: thread = conv pixel px
        const int g = (warp - 8) >> 2;                        // channel half: 32 g .. 32 g + 31
        const int w = warp & 3;                               // TMEM lane quarter
        const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(w * 32) << 16) + g * 32;
        const bool odd = lane & 1;
        float* edge_g = edge + g * (2 * 4 * 32);              // [2 parity][4 warps][32] of this warpgroup
        float* red_g = edge_g;                                // end of image: [4 warps][mean | M2][32] aliases it
        const int t = tid - 256 - g * 128;                    // thread index inside the warpgroup
        for (int it = 0; it < n_img; ++it) {
            const size_t b = static_cast<size_t>(blockIdx.x) + static_cast<size_t>(it) * gridDim.x;
            // Moments of the thread's 64 pooled values per channel, accumulated around a per-thread SHIFT (its first pooled
            // value): sum of (x - shift) and of (x - shift)^2 stay small for smooth maps, so fp32 accumulation does not
            // cancel when the variance is formed (a constant image has var ~ 1e-6 next to mean^2 ~ 1).
            float shift[16], sum[16], sq[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) shift[i] = sum[i] = sq[i] = 0.f;
            for (int py = 0; py < 64; ++py) {
                const int R2 = it * 128 + 2 * py + 1, R1 = R2 - 1, R0 = R2 - 2;
                mbar_wait(&dfull[R2 % TS_DBUF], (R2 / TS_DBUF) & 1);   // MMAs retire in order: rows R1, R0 are complete too
                tc_fence_after();
                float vm[32];
#pragma unroll
                for (int c = 0; c < 32; c += 16) {
                    uint32_t r1[16], r2[16], r0[16];
                    tmem_ld_32x16(t_lane + (R1 % TS_DBUF) * 64 + c, r1);
                    tmem_ld_32x16(t_lane + (R2 % TS_DBUF) * 64 + c, r2);
                    if (py > 0) tmem_ld_32x16(t_lane + (R0 % TS_DBUF) * 64 + c, r0);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        float v = fmaxf(__uint_as_float(r1[i]), __uint_as_float(r2[i]));
                        if (py > 0) v = fmaxf(v, __uint_as_float(r0[i]));
                        vm[c + i] = v;
                    }
                }
                // rows R0 and R1 are dead now (R2 is the next pooled row's R0); the image's last conv row retires with py = 63
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    if (py > 0) mbar_arrive(&dempty[R0 % TS_DBUF]);
                    mbar_arrive(&dempty[R1 % TS_DBUF]);
                    if (py == 63) mbar_arrive(&dempty[R2 % TS_DBUF]);
                }
                // hand the last pixel of this warp to the next warp (its lanes 0 / 1 need px - 1 / px - 2)
                float* e = edge_g + ((py & 1) * 4 + w) * 32;
                if (lane == 31) {
#pragma unroll
                    for (int i = 0; i < 32; i += 4) *reinterpret_cast<float4*>(e + i) = make_float4(vm[i], vm[i + 1], vm[i + 2], vm[i + 3]);
                }
                named_bar_sync(1 + g, 128);
                const float* ep = edge_g + ((py & 1) * 4 + (w > 0 ? w - 1 : 0)) * 32;
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    // even lane (pooled pixel centre px): channel i      from lanes l-1, l, l+1
                    // odd lane  (px = centre + 1):        channel 16 + i from lanes l-2, l-1, l
                    const float a_e = __shfl_up_sync(0xffffffffu, vm[i], 1);
                    const float c_e = __shfl_down_sync(0xffffffffu, vm[i], 1);
                    const float a_o = __shfl_up_sync(0xffffffffu, vm[16 + i], 2);
                    const float b_o = __shfl_up_sync(0xffffffffu, vm[16 + i], 1);
                    float left, mid, right;
                    if (!odd) {
                        left = (lane == 0) ? (w > 0 ? ep[i] : -INFINITY) : a_e;
                        mid = vm[i];
                        right = c_e;
                    } else {
                        left = (lane == 1) ? (w > 0 ? ep[16 + i] : -INFINITY) : a_o;
                        mid = b_o;
                        right = vm[16 + i];
                    }
                    const float pooled = fmaxf(fmaxf(fmaxf(left, mid), right), 0.f);   // ReLU commutes with max; bias is in the GEMM
                    if (py == 0) shift[i] = pooled;
                    const float dlt = pooled - shift[i];
                    sum[i] += dlt;
                    sq[i] = fmaf(dlt, dlt, sq[i]);
                }
            }
            // image done. Per thread and channel: n = 64 values, mean = shift + S / n, M2 = Q - S^2 / n (fp64). Partials are
            // merged pairwise (Chan et al.): equal counts at every level -> mean' = (ma + mb) / 2, M2' = M2a + M2b +
            // (mb - ma)^2 * n / 2: first across the 16 same-parity lanes of the warp (butterfly), then across the 4 warps.
            double cnt = 64.0;
            float mean_f[16], m2_f[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                double mean = static_cast<double>(shift[i]) + static_cast<double>(sum[i]) / 64.0;
                double m2 = static_cast<double>(sq[i]) - static_cast<double>(sum[i]) * static_cast<double>(sum[i]) / 64.0;
                double n = 64.0;
#pragma unroll
                for (int o = 2; o < 32; o <<= 1) {
                    const double mo = __shfl_xor_sync(0xffffffffu, mean, o);
                    const double qo = __shfl_xor_sync(0xffffffffu, m2, o);
                    const double dl = mo - mean;
                    m2 = m2 + qo + dl * dl * (n * 0.5);
                    mean = 0.5 * (mean + mo);
                    n *= 2.0;
                }
                cnt = n;
                mean_f[i] = static_cast<float>(mean);
                m2_f[i] = static_cast<float>(m2);
            }
            named_bar_sync(1 + g, 128);                        // the last pooled row's readers of the edge buffer are done
            if (lane < 2) {
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    red_g[(w * 2 + 0) * 32 + lane * 16 + i] = mean_f[i];
                    red_g[(w * 2 + 1) * 32 + lane * 16 + i] = m2_f[i];
                }
            }
            named_bar_sync(1 + g, 128);
            if (t < 32) {
                double mean = static_cast<double>(red_g[(0 * 2 + 0) * 32 + t]), m2 = static_cast<double>(red_g[(0 * 2 + 1) * 32 + t]);
                double n = cnt;                                // 1024 values per warp partial
                for (int ww = 1; ww < 4; ++ww) {
                    const double mb = static_cast<double>(red_g[(ww * 2 + 0) * 32 + t]), qb = static_cast<double>(red_g[(ww * 2 + 1) * 32 + t]);
                    const double dl = mb - mean, nt = n + cnt;
                    m2 = m2 + qb + dl * dl * (n * cnt / nt);
                    mean = mean + dl * (cnt / nt);
                    n = nt;
                }
                double var = m2 / (n - 1.0);
                if (var < 0.0) var = 0.0;
                out[b * 128 + g * 32 + t] = static_cast<float>(mean);
                out[b * 128 + 64 + g * 32 + t] = static_cast<float>(sqrt(var + static_cast<double>(eps)));
            }
            named_bar_sync(1 + g, 128);                        // `red` aliases the edge buffer the next image writes
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 4) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

int g_stem_force_ffma = 0;     // drag_debug_set key 8: 1 = the FP32 CUDA-core kernel of stem_stats.cu (A/B comparisons)

int stem_stats_ffma_device(const float* img, int B, int H, int W, const float* w_fold, const float* b_fold, float eps,
                           float* out, cudaStream_t st);

// img_kind 0: fp32 [B][3][256][256] in [0,1]; 1: uint8 [B][3][256][256] (divided by 255 in the loader)
int stem_stats_any_device(const void* img, int img_kind, int B, int H, int W, const float* w_fold, const float* b_fold,
                          float eps, float* out, cudaStream_t st) {
    DRAG_REQUIRE(img && w_fold && b_fold && out, "stem_stats: null pointer");
    DRAG_REQUIRE(H == 256 && W == 256, "stem_stats: input must be 256x256 (reference resize)");
    DRAG_REQUIRE(B >= 0 && (img_kind == 0 || img_kind == 1), "stem_stats: bad arguments");
    DRAG_REQUIRE((reinterpret_cast<uintptr_t>(img) & 7) == 0, "stem_stats: image pointer must be 8-byte aligned");
    if (B == 0) return DRAG_OK;
    if (g_stem_force_ffma && img_kind == 0)
        return stem_stats_ffma_device(static_cast<const float*>(img), B, H, W, w_fold, b_fold, eps, out, st);
    int sms = device_sm_count();
    if (sms <= 0) sms = 148;
    const int grid = B < sms ? B : sms;
    if (img_kind == 0) {
        DRAG_CUDA(cudaFuncSetAttribute(stem_stats_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TS_SMEM));
        stem_stats_tc_kernel<false><<<grid, TS_THREADS, TS_SMEM, st>>>(img, w_fold, b_fold, eps, out, B);
    } else {
        DRAG_CUDA(cudaFuncSetAttribute(stem_stats_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TS_SMEM));
        stem_stats_tc_kernel<true><<<grid, TS_THREADS, TS_SMEM, st>>>(img, w_fold, b_fold, eps, out, B);
    }
    count_launch();
    DRAG_CUDA(cudaGetLastError());
    return DRAG_OK;
}

int stem_stats_device(const float* img, int B, int H, int W, const float* w_fold, const float* b_fold, float eps, float* out,
                      cudaStream_t st) {
    return stem_stats_any_device(img, 0, B, H, W, w_fold, b_fold, eps, out, st);
}

}  // namespace drag
