// ResNet-50 stem (conv 7x7 s2 p3, eval-mode BatchNorm folded into the filter, ReLU, maxpool 3x3 s2
// p1) fused with the per-channel mean / sqrt(unbiased var + eps) "style" statistics the reference
// computes from it (retrieval/clip100_resnet_style_all_shots.py:51-74, 180-203).
//
// One CTA per image walks the 128 conv rows top to bottom; only the last three conv rows live in
// shared memory (ring), a pooled row is emitted after every odd conv row and folded straight into
// per-channel running sums, so neither the 64x128x128 conv map nor the 64x64x64 pooled map is ever
// written to HBM. fp32 FFMA throughout (the reference runs this in fp32), fp64 for the final
// moment combination. Traffic: 786 432 B in + 512 B out per image.
#include "common.cuh"

namespace drag {

constexpr int ST_H = 256, ST_W = 256;          // input size (reference resizes to 256x256)
constexpr int ST_HC = 128, ST_WC = 128;        // conv output
constexpr int ST_HP = 64, ST_WP = 64;          // pooled output
constexpr int ST_C = 64;
constexpr int ST_THREADS = 256;
constexpr int ST_INW = ST_W + 8;               // padded input row: x in [-3, W+3) -> index x+3, +2 slack
constexpr int ST_SMEM = (147 * ST_C + ST_C + 3 * 7 * ST_INW + 3 * ST_WC * ST_C) * 4;

__global__ void __launch_bounds__(ST_THREADS, 1)
stem_stats_kernel(const float* __restrict__ img, const float* __restrict__ w_fold,
                  const float* __restrict__ b_fold, float eps, float* __restrict__ out) {
    extern __shared__ __align__(16) float sm[];
    float* w_s = sm;                          // [147][64]  tap-major, channel contiguous
    float* b_s = w_s + 147 * ST_C;            // [64]
    float* in_s = b_s + ST_C;                 // [3][7][ST_INW]
    float* ring = in_s + 3 * 7 * ST_INW;      // [3][128][64]
    __shared__ double red_sum[4][ST_C];
    __shared__ double red_sq[4][ST_C];

    const int tid = threadIdx.x;
    const int b = blockIdx.x;
    const float* im = img + static_cast<size_t>(b) * 3 * ST_H * ST_W;

    // w_fold is [64][3][7][7] (OIHW); transpose to [tap][c]
    for (int i = tid; i < 147 * ST_C; i += ST_THREADS) {
        int c = i / 147, tap = i - c * 147;
        w_s[tap * ST_C + c] = w_fold[i];
    }
    if (tid < ST_C) b_s[tid] = b_fold[tid];

    const int cg = tid & 15;   // channels cg*4 .. cg*4+3
    const int xg = tid >> 4;   // conv x positions xg*8 .. xg*8+7
    const int pc = tid & 63;   // pooling: channel
    const int pg = tid >> 6;   // pooling: x group (16 pooled columns)
    double acc_sum = 0.0, acc_sq = 0.0;

    for (int cy = 0; cy < ST_HC; ++cy) {
        __syncthreads();  // previous iteration's readers of in_s / ring are done
        // input rows 2cy-3 .. 2cy+3, zero padded
        for (int i = tid; i < 3 * 7 * ST_INW; i += ST_THREADS) {
            int ci = i / (7 * ST_INW);
            int rem = i - ci * 7 * ST_INW;
            int ky = rem / ST_INW;
            int xx = rem - ky * ST_INW - 3;
            int yy = 2 * cy - 3 + ky;
            float v = 0.f;
            if (yy >= 0 && yy < ST_H && xx >= 0 && xx < ST_W) v = im[(ci * ST_H + yy) * ST_W + xx];
            in_s[i] = v;
        }
        __syncthreads();
        float acc[8][4];
#pragma unroll
        for (int p = 0; p < 8; ++p)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[p][c] = 0.f;
        for (int ci = 0; ci < 3; ++ci) {
            for (int ky = 0; ky < 7; ++ky) {
                const float* row = in_s + (ci * 7 + ky) * ST_INW + xg * 16;  // x = 2*(xg*8+p) + kx - 3 -> idx 2*(xg*8+p)+kx
                float iv[21];
#pragma unroll
                for (int i = 0; i < 21; ++i) iv[i] = row[i];
#pragma unroll
                for (int kx = 0; kx < 7; ++kx) {
                    const float4 w4 =
                        *reinterpret_cast<const float4*>(w_s + ((ci * 7 + ky) * 7 + kx) * ST_C + cg * 4);
#pragma unroll
                    for (int p = 0; p < 8; ++p) {
                        const float x = iv[2 * p + kx];
                        acc[p][0] = fmaf(x, w4.x, acc[p][0]);
                        acc[p][1] = fmaf(x, w4.y, acc[p][1]);
                        acc[p][2] = fmaf(x, w4.z, acc[p][2]);
                        acc[p][3] = fmaf(x, w4.w, acc[p][3]);
                    }
                }
            }
        }
        {
            const float4 bb = *reinterpret_cast<const float4*>(b_s + cg * 4);
            float* dst = ring + ((cy % 3) * ST_WC + xg * 8) * ST_C + cg * 4;
#pragma unroll
            for (int p = 0; p < 8; ++p) {
                float4 o;
                o.x = fmaxf(acc[p][0] + bb.x, 0.f);
                o.y = fmaxf(acc[p][1] + bb.y, 0.f);
                o.z = fmaxf(acc[p][2] + bb.z, 0.f);
                o.w = fmaxf(acc[p][3] + bb.w, 0.f);
                *reinterpret_cast<float4*>(dst + p * ST_C) = o;
            }
        }
        if (cy & 1) {
            __syncthreads();
            // pooled row py = (cy-1)/2 covers conv rows cy-2 (if >= 0), cy-1, cy
            for (int i = 0; i < 16; ++i) {
                const int px = pg * 16 + i;
                float m = 0.f;  // post-ReLU values are >= 0 and every window holds a valid element
                for (int dy = 0; dy < 3; ++dy) {
                    const int ry = cy - 2 + dy;
                    if (ry < 0) continue;
                    const float* rr = ring + (ry % 3) * ST_WC * ST_C;
                    for (int dx = -1; dx <= 1; ++dx) {
                        const int rx = 2 * px + dx;
                        if (rx < 0 || rx >= ST_WC) continue;
                        m = fmaxf(m, rr[rx * ST_C + pc]);
                    }
                }
                const double md = static_cast<double>(m);
                acc_sum += md;
                acc_sq = fma(md, md, acc_sq);
            }
        }
    }
    red_sum[pg][pc] = acc_sum;
    red_sq[pg][pc] = acc_sq;
    __syncthreads();
    if (tid < ST_C) {
        const double n = static_cast<double>(ST_HP * ST_WP);
        double s = red_sum[0][tid] + red_sum[1][tid] + red_sum[2][tid] + red_sum[3][tid];
        double ss = red_sq[0][tid] + red_sq[1][tid] + red_sq[2][tid] + red_sq[3][tid];
        double mean = s / n;
        double var = (ss - n * mean * mean) / (n - 1.0);
        if (var < 0.0) var = 0.0;
        out[static_cast<size_t>(b) * 2 * ST_C + tid] = static_cast<float>(mean);
        out[static_cast<size_t>(b) * 2 * ST_C + ST_C + tid] =
            static_cast<float>(sqrt(var + static_cast<double>(eps)));
    }
}

// FP32 CUDA-core path: kept for A/B comparisons against the tensor-core kernel (stem_stats_tc.cu; drag_debug_set key 8).
int stem_stats_ffma_device(const float* img, int B, int H, int W, const float* w_fold,
                           const float* b_fold, float eps, float* out, cudaStream_t st) {
    DRAG_REQUIRE(img && w_fold && b_fold && out, "stem_stats: null pointer");
    DRAG_REQUIRE(H == ST_H && W == ST_W, "stem_stats: input must be 256x256 (reference resize)");
    DRAG_REQUIRE(B >= 0, "stem_stats: negative batch");
    if (B == 0) return DRAG_OK;
    DRAG_CUDA(cudaFuncSetAttribute(stem_stats_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   ST_SMEM));
    stem_stats_kernel<<<B, ST_THREADS, ST_SMEM, st>>>(img, w_fold, b_fold, eps, out); count_launch();
    DRAG_CUDA(cudaGetLastError());
    return DRAG_OK;
}

}  // namespace drag
