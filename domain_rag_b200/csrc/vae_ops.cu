// HBM-bound kernels of the Flux VAE path (AutoencoderKL decode / encode around the implicit-GEMM convolutions):
// GroupNorm(32)+SiLU on NHWC activations, nearest 2x upsampling, row softmax of the mid-block attention scores,
// NCHW <-> NHWC layout changes fused with the latent scale/shift, and the image post-process
// (VaeImageProcessor: x/2+0.5, clamp, x255, round) that produces the uint8 pixels of `pipe(...).images`
// (reference: batch_generate_flux_kshot.py:467-474, outpainting_updown_sampling_redux.py:1246-1257).
// bf16 storage, fp32 arithmetic, 16-byte vector accesses, fixed-order (deterministic) reductions.
#include "common.cuh"
#include "vae_ops.cuh"

namespace drag {

__device__ __forceinline__ void vld8(const __nv_bfloat16* p, float* v) {
    uint4 u = *reinterpret_cast<const uint4*>(p);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float2 f = __bfloat1622float2(h[i]);
        v[2 * i] = f.x;
        v[2 * i + 1] = f.y;
    }
}
__device__ __forceinline__ void vst8(__nv_bfloat16* p, const float* v) {
    __nv_bfloat162 h[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    *reinterpret_cast<uint4*>(p) = *reinterpret_cast<uint4*>(h);
}

// ------------------------------------------------------------------------------------------ GroupNorm
// x [B][HW][C]. Pass 1: per (image, pixel chunk) per-channel sum and sum of squares -> part[b][chunk][2][C].
// One thread owns one channel octet for a strided subset of the chunk's pixels; a shared-memory tree finishes.
constexpr int GN_THREADS = 256;

__global__ void __launch_bounds__(GN_THREADS) gn_partial_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ part,
                                                                int HW, int C, int chunk_px) {
    extern __shared__ float sm[];                     // [GN_THREADS][16]
    const int b = blockIdx.y, chunk = blockIdx.x;
    const int octs = C / 8;
    const int oct = threadIdx.x % octs, sub = threadIdx.x / octs, nsub = GN_THREADS / octs;
    const int p0 = chunk * chunk_px, p1 = min(HW, p0 + chunk_px);
    float s[8] = {0, 0, 0, 0, 0, 0, 0, 0}, q[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (sub < nsub) {
        const __nv_bfloat16* base = x + (static_cast<size_t>(b) * HW) * C + oct * 8;
#pragma unroll 4
        for (int p = p0 + sub; p < p1; p += nsub) {
            float v[8];
            vld8(base + static_cast<size_t>(p) * C, v);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                s[i] += v[i];
                q[i] = fmaf(v[i], v[i], q[i]);
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        sm[threadIdx.x * 16 + i] = s[i];
        sm[threadIdx.x * 16 + 8 + i] = q[i];
    }
    __syncthreads();
    // threads [0, C): channel c sums over the nsub pixel subsets in a fixed order
    for (int c = threadIdx.x; c < 2 * C; c += GN_THREADS) {
        const int which = c / C, ch = c - which * C;
        const int o = ch / 8, e = ch - o * 8;
        float acc = 0.f;
        for (int k = 0; k < nsub; ++k) acc += sm[(k * octs + o) * 16 + which * 8 + e];
        part[((static_cast<size_t>(b) * gridDim.x + chunk) * 2 + which) * C + ch] = acc;
    }
}

// Pass 2: stats[b][g] = (mean, rstd) from the partials (double accumulation, fixed order).
__global__ void gn_finalize_kernel(const float* __restrict__ part, float* __restrict__ stats, int chunks, int C,
                                   int groups, int HW, float eps) {
    const int b = blockIdx.y, g = blockIdx.x;
    const int cpg = C / groups;
    double s = 0.0, q = 0.0;
    for (int i = threadIdx.x; i < chunks * cpg; i += 32) {
        const int chunk = i / cpg, ch = g * cpg + (i - chunk * cpg);
        s += part[((static_cast<size_t>(b) * chunks + chunk) * 2 + 0) * C + ch];
        q += part[((static_cast<size_t>(b) * chunks + chunk) * 2 + 1) * C + ch];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o);
        q += __shfl_xor_sync(0xffffffffu, q, o);
    }
    if (threadIdx.x == 0) {
        const double n = static_cast<double>(HW) * cpg;
        const double mean = s / n;
        const double var = fmax(q / n - mean * mean, 0.0);
        stats[(b * groups + g) * 2 + 0] = static_cast<float>(mean);
        stats[(b * groups + g) * 2 + 1] = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
    }
}

// Pass 3: y = (x - mean) * rstd * gamma + beta, optional SiLU.
__global__ void __launch_bounds__(256) gn_apply_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y,
                                                       const float* __restrict__ stats, const __nv_bfloat16* __restrict__ gamma,
                                                       const __nv_bfloat16* __restrict__ beta, int HW, int C, int groups,
                                                       int silu, size_t total_oct) {
    const size_t i = static_cast<size_t>(blockIdx.x) * 256 + threadIdx.x;
    if (i >= total_oct) return;
    const int octs = C / 8;
    const int oct = static_cast<int>(i % octs);
    const size_t px = i / octs;
    const int b = static_cast<int>(px / HW);
    const int cpg = C / groups;
    float v[8], gm[8], bt[8];
    vld8(x + i * 8, v);
    vld8(gamma + oct * 8, gm);
    vld8(beta + oct * 8, bt);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        const int g = (oct * 8 + e) / cpg;
        const float mean = stats[(b * groups + g) * 2], rstd = stats[(b * groups + g) * 2 + 1];
        float t = (v[e] - mean) * rstd * gm[e] + bt[e];
        if (silu) t = t / (1.f + __expf(-t));
        v[e] = t;
    }
    vst8(y + i * 8, v);
}

int groupnorm_nhwc(const __nv_bfloat16* x, __nv_bfloat16* y, int B, int HW, int C, int groups, const __nv_bfloat16* gamma,
                   const __nv_bfloat16* beta, float eps, int silu, float* workspace, size_t workspace_floats,
                   cudaStream_t st) {
    DRAG_REQUIRE(x && y && gamma && beta && workspace, "groupnorm: null pointer");
    DRAG_REQUIRE(B >= 1 && HW >= 1 && C % 8 == 0 && C % groups == 0 && C / 8 <= GN_THREADS && GN_THREADS % (C / 8) == 0,
                 "groupnorm: unsupported channel count");
    int chunks = HW / 256;
    if (chunks < 1) chunks = 1;
    if (chunks > 1024) chunks = 1024;
    const int chunk_px = ceil_div(HW, chunks);
    chunks = ceil_div(HW, chunk_px);
    const size_t need = static_cast<size_t>(B) * chunks * 2 * C + static_cast<size_t>(B) * groups * 2;
    DRAG_REQUIRE(workspace_floats >= need, "groupnorm: workspace too small");
    float* part = workspace;
    float* stats = workspace + static_cast<size_t>(B) * chunks * 2 * C;
    gn_partial_kernel<<<dim3(chunks, B), GN_THREADS, GN_THREADS * 16 * sizeof(float), st>>>(x, part, HW, C, chunk_px); count_launch();
    gn_finalize_kernel<<<dim3(groups, B), 32, 0, st>>>(part, stats, chunks, C, groups, HW, eps); count_launch();
    const size_t total_oct = static_cast<size_t>(B) * HW * (C / 8);
    gn_apply_kernel<<<static_cast<unsigned>((total_oct + 255) / 256), 256, 0, st>>>(x, y, stats, gamma, beta, HW, C, groups,
                                                                                   silu, total_oct); count_launch();
    DRAG_CUDA(cudaGetLastError());
    return DRAG_OK;
}

// ------------------------------------------------------------------------------------------ upsample
__global__ void __launch_bounds__(256) upsample2x_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y,
                                                         int H, int W, int C, size_t total_oct) {
    const size_t i = static_cast<size_t>(blockIdx.x) * 256 + threadIdx.x;   // output octet
    if (i >= total_oct) return;
    const int octs = C / 8;
    const int oct = static_cast<int>(i % octs);
    size_t p = i / octs;
    const int xo = static_cast<int>(p % (2 * W));
    p /= 2 * W;
    const int yo = static_cast<int>(p % (2 * H));
    const size_t b = p / (2 * H);
    const uint4 v = *reinterpret_cast<const uint4*>(x + ((b * H + (yo >> 1)) * W + (xo >> 1)) * C + oct * 8);
    *reinterpret_cast<uint4*>(y + i * 8) = v;
}

int upsample2x_nhwc(const __nv_bfloat16* x, __nv_bfloat16* y, int B, int H, int W, int C, cudaStream_t st) {
    DRAG_REQUIRE(x && y && C % 8 == 0, "upsample2x: bad arguments");
    const size_t total_oct = static_cast<size_t>(B) * 4 * H * W * (C / 8);
    upsample2x_kernel<<<static_cast<unsigned>((total_oct + 255) / 256), 256, 0, st>>>(x, y, H, W, C, total_oct); count_launch();
    DRAG_CUDA(cudaGetLastError());
    return DRAG_OK;
}

// ------------------------------------------------------------------------------------------ softmax
// p[r][c] = softmax_c(s[r][c]) in bf16; one block per row, three passes over a row that stays in L2.
__global__ void __launch_bounds__(256) softmax_rows_kernel(const float* __restrict__ s, size_t ld_s,
                                                           __nv_bfloat16* __restrict__ p, size_t ld_p, int cols) {
    __shared__ float red[8];
    __shared__ float bcast;
    const float* sr = s + static_cast<size_t>(blockIdx.x) * ld_s;
    __nv_bfloat16* pr = p + static_cast<size_t>(blockIdx.x) * ld_p;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float m = -INFINITY;
    for (int c = threadIdx.x * 4; c < cols; c += 1024) {
        const float4 v = *reinterpret_cast<const float4*>(sr + c);
        m = fmaxf(m, fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w)));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane == 0) red[warp] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = red[0];
        for (int i = 1; i < 8; ++i) t = fmaxf(t, red[i]);
        bcast = t;
    }
    __syncthreads();
    m = bcast;
    float sum = 0.f;
    for (int c = threadIdx.x * 4; c < cols; c += 1024) {
        const float4 v = *reinterpret_cast<const float4*>(sr + c);
        sum += __expf(v.x - m) + __expf(v.y - m) + __expf(v.z - m) + __expf(v.w - m);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    __syncthreads();
    if (lane == 0) red[warp] = sum;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int i = 0; i < 8; ++i) t += red[i];
        bcast = 1.f / t;
    }
    __syncthreads();
    const float inv = bcast;
    for (int c = threadIdx.x * 4; c < cols; c += 1024) {
        const float4 v = *reinterpret_cast<const float4*>(sr + c);
        __nv_bfloat162 a = __floats2bfloat162_rn(__expf(v.x - m) * inv, __expf(v.y - m) * inv);
        __nv_bfloat162 b = __floats2bfloat162_rn(__expf(v.z - m) * inv, __expf(v.w - m) * inv);
        uint2 u;
        u.x = *reinterpret_cast<uint32_t*>(&a);
        u.y = *reinterpret_cast<uint32_t*>(&b);
        *reinterpret_cast<uint2*>(pr + c) = u;
    }
}

int softmax_rows(const float* s, size_t ld_s, __nv_bfloat16* p, size_t ld_p, int rows, int cols, cudaStream_t st) {
    DRAG_REQUIRE(s && p && rows >= 1 && cols >= 4 && cols % 4 == 0 && ld_s % 4 == 0 && ld_p % 4 == 0, "softmax_rows: bad arguments");
    softmax_rows_kernel<<<rows, 256, 0, st>>>(s, ld_s, p, ld_p, cols); count_launch();
    DRAG_CUDA(cudaGetLastError());
    return DRAG_OK;
}

// ------------------------------------------------------------------------------------------ layouts
// out[b][y][x][c] = in[b][c][y][x] * scale + shift for c < C, 0 for C <= c < C_pad (NCHW -> channel-padded NHWC bf16).
template <typename T>
__global__ void __launch_bounds__(256) nchw_to_nhwc_kernel(const T* __restrict__ in, __nv_bfloat16* __restrict__ out, int C,
                                                           int HW, int C_pad, float scale, float shift, size_t total) {
    const size_t i = static_cast<size_t>(blockIdx.x) * 256 + threadIdx.x;   // (b, pixel, c_pad)
    if (i >= total) return;
    const int c = static_cast<int>(i % C_pad);
    const size_t bp = i / C_pad;
    const size_t b = bp / HW, px = bp - b * HW;
    float v = 0.f;
    if (c < C) v = static_cast<float>(in[(b * C + c) * HW + px]) * scale + shift;
    out[i] = __float2bfloat16(v);
}

int nchw_to_nhwc_pad(const void* in, int in_is_f32, __nv_bfloat16* out, int B, int C, int H, int W, int C_pad, float scale,
                     float shift, cudaStream_t st) {
    DRAG_REQUIRE(in && out && C_pad >= C, "nchw_to_nhwc: bad arguments");
    const size_t total = static_cast<size_t>(B) * H * W * C_pad;
    const unsigned grid = static_cast<unsigned>((total + 255) / 256);
    if (in_is_f32)
        nchw_to_nhwc_kernel<float><<<grid, 256, 0, st>>>(static_cast<const float*>(in), out, C, H * W, C_pad, scale, shift, total);
    else
        nchw_to_nhwc_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(in), out, C, H * W, C_pad,
                                                                 scale, shift, total);
    count_launch();
    DRAG_CUDA(cudaGetLastError());
    return DRAG_OK;
}

// out[b][c][y][x] (fp32) = in[b][y][x][c] * scale + shift for the first C channels of an NHWC row of `ld` elements.
template <typename T>
__global__ void __launch_bounds__(256) nhwc_to_nchw_kernel(const T* __restrict__ in, int ld, float* __restrict__ out, int C,
                                                           int HW, float scale, float shift, size_t total) {
    const size_t i = static_cast<size_t>(blockIdx.x) * 256 + threadIdx.x;   // (b, c, pixel)
    if (i >= total) return;
    const size_t px = i % HW;
    const size_t bc = i / HW;
    const size_t b = bc / C, c = bc - b * C;
    out[i] = static_cast<float>(in[(b * HW + px) * ld + c]) * scale + shift;
}

int nhwc_to_nchw_f32(const void* in, int in_is_f32, int ld, float* out, int B, int C, int H, int W, float scale, float shift,
                     cudaStream_t st) {
    DRAG_REQUIRE(in && out && ld >= C, "nhwc_to_nchw: bad arguments");
    const size_t total = static_cast<size_t>(B) * C * H * W;
    const unsigned grid = static_cast<unsigned>((total + 255) / 256);
    if (in_is_f32)
        nhwc_to_nchw_kernel<float><<<grid, 256, 0, st>>>(static_cast<const float*>(in), ld, out, C, H * W, scale, shift, total);
    else
        nhwc_to_nchw_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(in), ld, out, C, H * W, scale,
                                                                 shift, total);
    count_launch();
    DRAG_CUDA(cudaGetLastError());
    return DRAG_OK;
}

// uint8 RGB pixels from the decoder output (fp32 NHWC rows of `ld` floats, first 3 used): clamp(x/2 + 0.5, 0, 1) * 255, rounded
// half to even like torch.round / numpy.round in VaeImageProcessor.
__global__ void __launch_bounds__(256) image_postprocess_kernel(const float* __restrict__ in, int ld, uint8_t* __restrict__ out,
                                                                size_t pixels) {
    const size_t i = static_cast<size_t>(blockIdx.x) * 256 + threadIdx.x;
    if (i >= pixels) return;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float v = fminf(fmaxf(in[i * ld + c] * 0.5f + 0.5f, 0.f), 1.f);
        out[i * 3 + c] = static_cast<uint8_t>(__float2int_rn(v * 255.f));
    }
}

int image_postprocess_u8(const float* in, int ld, uint8_t* out, size_t pixels, cudaStream_t st) {
    DRAG_REQUIRE(in && out && ld >= 3, "image_postprocess: bad arguments");
    image_postprocess_kernel<<<static_cast<unsigned>((pixels + 255) / 256), 256, 0, st>>>(in, ld, out, pixels); count_launch();
    DRAG_CUDA(cudaGetLastError());
    return DRAG_OK;
}

// NHWC bf16 (channel padded) in [-1, 1] from uint8 RGB pixels: x / 255 * 2 - 1 (VaeImageProcessor.preprocess). With a
// mask (uint8 per pixel, non-zero = repaint) the masked pixels become 0, i.e. init_image * (1 - mask) of FluxFillPipeline.
__global__ void __launch_bounds__(256) image_preprocess_kernel(const uint8_t* __restrict__ in, const uint8_t* __restrict__ mask,
                                                               __nv_bfloat16* __restrict__ out, int C_pad, size_t total) {
    const size_t i = static_cast<size_t>(blockIdx.x) * 256 + threadIdx.x;   // (pixel, c_pad)
    if (i >= total) return;
    const int c = static_cast<int>(i % C_pad);
    const size_t px = i / C_pad;
    float v = 0.f;
    if (c < 3 && !(mask && mask[px])) v = static_cast<float>(in[px * 3 + c]) / 255.f * 2.f - 1.f;
    out[i] = __float2bfloat16(v);
}

int image_preprocess_u8(const uint8_t* in, const uint8_t* mask, __nv_bfloat16* out, size_t pixels, int C_pad, cudaStream_t st) {
    DRAG_REQUIRE(in && out && C_pad >= 3, "image_preprocess: bad arguments");
    const size_t total = pixels * C_pad;
    image_preprocess_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, st>>>(in, mask, out, C_pad, total); count_launch();
    DRAG_CUDA(cudaGetLastError());
    return DRAG_OK;
}

// out = a * x + b * y on bf16 vectors (fp32 arithmetic): the flow-match noising sigma * noise + (1 - sigma) * image_latents
// of the img2img / fill pipelines (scheduler.scale_noise).
__global__ void __launch_bounds__(256) axpby_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ y,
                                                    float a, float b, __nv_bfloat16* __restrict__ out, size_t n) {
    const size_t i = static_cast<size_t>(blockIdx.x) * 256 + threadIdx.x;
    if (i < n) out[i] = __float2bfloat16(a * __bfloat162float(x[i]) + b * __bfloat162float(y[i]));
}

int axpby_bf16(const __nv_bfloat16* x, const __nv_bfloat16* y, float a, float b, __nv_bfloat16* out, size_t n, cudaStream_t st) {
    DRAG_REQUIRE(x && y && out, "axpby: null pointer");
    axpby_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, st>>>(x, y, a, b, out, n); count_launch();
    DRAG_CUDA(cudaGetLastError());
    return DRAG_OK;
}

}  // namespace drag
