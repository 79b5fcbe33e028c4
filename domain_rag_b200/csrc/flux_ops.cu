// HBM-bound elementwise / row kernels around the GEMM and attention cores of the ViT and Flux paths:
// LayerNorm (+AdaLN modulate or affine), sinusoidal timestep embedding, SiLU of the summed
// conditioning vector, flow-match Euler update, Redux prompt blend, L2 normalisation.
// All bf16 storage, fp32 arithmetic; 16-byte vector accesses; one warp per row for the row kernels.
#include "common.cuh"
#include "flux_ops.cuh"

namespace drag {

__device__ __forceinline__ void ld8(const __nv_bfloat16* p, float* v) {
    uint4 u = *reinterpret_cast<const uint4*>(p);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float2 f = __bfloat1622float2(h[i]);
        v[2 * i] = f.x;
        v[2 * i + 1] = f.y;
    }
}
__device__ __forceinline__ void st8(__nv_bfloat16* p, const float* v) {
    __nv_bfloat162 h[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    *reinterpret_cast<uint4*>(p) = *reinterpret_cast<uint4*>(h);
}
__device__ __forceinline__ float wsum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// out[row] = LN(x[row]) * mul + add.
//   mul_add_one = 1: mul = 1 + mulp[b*mul_ld + c]   (AdaLN: scale)     add = addp[b*add_ld + c] (shift)
//   mul_add_one = 0: mul = mulp[c] (gamma)                              add = addp[c] (beta); either may be null
// b = row / rows_per_batch. Two exact passes for mean / variance (second pass hits L1/L2).
__global__ void __launch_bounds__(256) layernorm_kernel(const __nv_bfloat16* __restrict__ x, int ldx,
                                                        __nv_bfloat16* __restrict__ out, int ldo, int M, int d,
                                                        const __nv_bfloat16* __restrict__ mulp, int mul_ld,
                                                        const __nv_bfloat16* __restrict__ addp, int add_ld,
                                                        int mul_add_one, int rows_per_batch, float eps) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= M) return;
    const __nv_bfloat16* xr = x + static_cast<size_t>(row) * ldx;
    float s = 0.f;
    for (int c = lane * 8; c < d; c += 256) {
        float v[8];
        ld8(xr + c, v);
#pragma unroll
        for (int i = 0; i < 8; ++i) s += v[i];
    }
    const float mean = wsum(s) / d;
    float ss = 0.f;
    for (int c = lane * 8; c < d; c += 256) {
        float v[8];
        ld8(xr + c, v);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float t = v[i] - mean;
            ss = fmaf(t, t, ss);
        }
    }
    const float rstd = rsqrtf(wsum(ss) / d + eps);
    const int b = row / rows_per_batch;
    const __nv_bfloat16* mr = mulp ? mulp + static_cast<size_t>(mul_add_one ? b : 0) * mul_ld : nullptr;
    const __nv_bfloat16* ar = addp ? addp + static_cast<size_t>(mul_add_one ? b : 0) * add_ld : nullptr;
    __nv_bfloat16* orow = out + static_cast<size_t>(row) * ldo;
    for (int c = lane * 8; c < d; c += 256) {
        float v[8], mu[8], ad[8];
        ld8(xr + c, v);
        if (mr) ld8(mr + c, mu);
        if (ar) ld8(ar + c, ad);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float y = (v[i] - mean) * rstd;
            if (mr) y *= (mul_add_one ? 1.f + mu[i] : mu[i]);
            if (ar) y += ad[i];
            v[i] = y;
        }
        st8(orow + c, v);
    }
}

// Same arithmetic with the row held in registers (d = NCH * 256: one warp per row, NCH 16-byte loads per lane, all in flight
// at once): x crosses the memory system ONCE instead of three times. Row widths of the towers on the path: 768 / 1024 (CLIP
// ViT), 3072 (Flux). The statistics are the same two exact passes (mean, then centred sum of squares) - now over registers.
template <int NCH>
__global__ void __launch_bounds__(256, 2) layernorm_reg_kernel(const __nv_bfloat16* __restrict__ x, int ldx,
                                                            __nv_bfloat16* __restrict__ out, int ldo, int M,
                                                            const __nv_bfloat16* __restrict__ mulp, int mul_ld,
                                                            const __nv_bfloat16* __restrict__ addp, int add_ld,
                                                            int mul_add_one, int rows_per_batch, float eps) {
    constexpr int d = NCH * 256;
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= M) return;
    const __nv_bfloat16* xr = x + static_cast<size_t>(row) * ldx + lane * 8;
    uint4 raw[NCH];
#pragma unroll
    for (int k = 0; k < NCH; ++k) raw[k] = *reinterpret_cast<const uint4*>(xr + k * 256);
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < NCH; ++k) {
        const __nv_bfloat162* p2 = reinterpret_cast<const __nv_bfloat162*>(&raw[k]);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float2 f = __bfloat1622float2(p2[i]);
            s += f.x;                        // same element order as the three-pass kernel
            s += f.y;
        }
    }
    const float mean = wsum(s) / d;
    float ss = 0.f;
#pragma unroll
    for (int k = 0; k < NCH; ++k) {
        const __nv_bfloat162* p2 = reinterpret_cast<const __nv_bfloat162*>(&raw[k]);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float2 f = __bfloat1622float2(p2[i]);
            const float t0 = f.x - mean, t1 = f.y - mean;
            ss = fmaf(t0, t0, ss);
            ss = fmaf(t1, t1, ss);
        }
    }
    const float rstd = rsqrtf(wsum(ss) / d + eps);
    const int b = row / rows_per_batch;
    const __nv_bfloat16* mr = mulp ? mulp + static_cast<size_t>(mul_add_one ? b : 0) * mul_ld + lane * 8 : nullptr;
    const __nv_bfloat16* ar = addp ? addp + static_cast<size_t>(mul_add_one ? b : 0) * add_ld + lane * 8 : nullptr;
    __nv_bfloat16* orow = out + static_cast<size_t>(row) * ldo + lane * 8;
#pragma unroll
    for (int k = 0; k < NCH; ++k) {
        float v[8], mu[8], ad[8];
        const __nv_bfloat162* p2 = reinterpret_cast<const __nv_bfloat162*>(&raw[k]);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float2 f = __bfloat1622float2(p2[i]);
            v[2 * i] = f.x;
            v[2 * i + 1] = f.y;
        }
        if (mr) ld8(mr + k * 256, mu);
        if (ar) ld8(ar + k * 256, ad);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float y = (v[i] - mean) * rstd;
            if (mr) y *= (mul_add_one ? 1.f + mu[i] : mu[i]);
            if (ar) y += ad[i];
            v[i] = y;
        }
        st8(orow + k * 256, v);
    }
}

int layernorm_bf16(const __nv_bfloat16* x, int ldx, __nv_bfloat16* out, int ldo, int M, int d,
                   const __nv_bfloat16* mul, int mul_ld, const __nv_bfloat16* add, int add_ld, int mul_add_one,
                   int rows_per_batch, float eps, cudaStream_t st) {
    DRAG_REQUIRE(x && out && M >= 1 && d >= 8 && d % 8 == 0 && ldx % 8 == 0 && ldo % 8 == 0, "layernorm: bad arguments");
    if (rows_per_batch <= 0) rows_per_batch = 1 << 30;
    const bool aligned = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out)) & 15) == 0;
#define DRAG_LN_REG(NCH)                                                                                                   \
    layernorm_reg_kernel<NCH><<<ceil_div(M, 8), 256, 0, st>>>(x, ldx, out, ldo, M, mul, mul_ld, add, add_ld, mul_add_one, \
                                                              rows_per_batch, eps)
    if (aligned && (d == 768 || d == 1024 || d == 3072)) {
        if (d == 768) DRAG_LN_REG(3);
        else if (d == 1024) DRAG_LN_REG(4);
        else DRAG_LN_REG(12);
        count_launch();
        DRAG_CUDA(cudaGetLastError());
        return DRAG_OK;
    }
#undef DRAG_LN_REG
    layernorm_kernel<<<ceil_div(M, 8), 256, 0, st>>>(x, ldx, out, ldo, M, d, mul, mul_ld, add, add_ld, mul_add_one,
                                                     rows_per_batch, eps); count_launch();
    DRAG_CUDA(cudaGetLastError());
    return DRAG_OK;
}

// emb[b][0:128] = cos(1000 t_b w_i), emb[b][128:256] = sin(...), w_i = exp(-ln(1e4) i / 128).
__global__ void timestep_embed_kernel(const float* __restrict__ t, __nv_bfloat16* __restrict__ out, int B) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * 128) return;
    const int b = i >> 7, j = i & 127;
    const float freq = expf(-9.210340371976184f * static_cast<float>(j) / 128.f);
    const float arg = 1000.f * t[b] * freq;
    float sn, cs;
    sincosf(arg, &sn, &cs);
    out[b * 256 + j] = __float2bfloat16(cs);
    out[b * 256 + 128 + j] = __float2bfloat16(sn);
}
int timestep_embed(const float* t_dev, __nv_bfloat16* out, int B, cudaStream_t st) {
    DRAG_REQUIRE(t_dev && out && B >= 1, "timestep_embed: bad arguments");
    timestep_embed_kernel<<<ceil_div(B * 128, 128), 128, 0, st>>>(t_dev, out, B); count_launch();
    DRAG_CUDA(cudaGetLastError());
    return DRAG_OK;
}

// out = silu(a + b + c)  (b, c optional) - the conditioning vector fed to every modulation Linear.
__global__ void sum_silu_kernel(const __nv_bfloat16* a, const __nv_bfloat16* b, const __nv_bfloat16* c,
                                __nv_bfloat16* out, int n, int apply_silu) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    // each stage rounds to bf16 like the reference's bf16 tensor adds
    float v = __bfloat162float(a[i]);
    if (b) v = __bfloat162float(__float2bfloat16(v + __bfloat162float(b[i])));
    if (c) v = __bfloat162float(__float2bfloat16(v + __bfloat162float(c[i])));
    if (apply_silu) v = v / (1.f + __expf(-v));
    out[i] = __float2bfloat16(v);
}
int sum_silu(const __nv_bfloat16* a, const __nv_bfloat16* b, const __nv_bfloat16* c, __nv_bfloat16* out, int n,
             int apply_silu, cudaStream_t st) {
    DRAG_REQUIRE(a && out && n >= 1, "sum_silu: bad arguments");
    sum_silu_kernel<<<ceil_div(n, 256), 256, 0, st>>>(a, b, c, out, n, apply_silu); count_launch();
    DRAG_CUDA(cudaGetLastError());
    return DRAG_OK;
}

// x[r][c] = bf16(float(x[r][c]) + dsigma * float(v[r][c]))   flow-match Euler step on a strided view
__global__ void euler_step_kernel(__nv_bfloat16* x, int ldx, const __nv_bfloat16* v, int ldv, int rows, int cols,
                                  float dsigma) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    const int cpr = cols / 8;
    if (i >= static_cast<int64_t>(rows) * cpr) return;
    const int r = static_cast<int>(i / cpr), c = static_cast<int>(i - static_cast<int64_t>(r) * cpr) * 8;
    float xv[8], vv[8];
    ld8(x + static_cast<size_t>(r) * ldx + c, xv);
    ld8(v + static_cast<size_t>(r) * ldv + c, vv);
#pragma unroll
    for (int j = 0; j < 8; ++j) xv[j] = __fadd_rn(xv[j], __fmul_rn(dsigma, vv[j]));   // unfused, like torch
    st8(x + static_cast<size_t>(r) * ldx + c, xv);
}
int euler_step(__nv_bfloat16* x, int ldx, const __nv_bfloat16* v, int ldv, int rows, int cols, float dsigma,
               cudaStream_t st) {
    DRAG_REQUIRE(x && v && rows >= 1 && cols % 8 == 0 && ldx % 8 == 0 && ldv % 8 == 0, "euler_step: bad arguments");
    const int64_t n = static_cast<int64_t>(rows) * (cols / 8);
    euler_step_kernel<<<ceil_div(n, 256), 256, 0, st>>>(x, ldx, v, ldv, rows, cols, dsigma); count_launch();
    DRAG_CUDA(cudaGetLastError());
    return DRAG_OK;
}

// Flux 2x2 latent packing (diffusers _pack_latents / _unpack_latents): z [B][C][h][w] <-> token s = (y/2)(w/2) + x/2,
// channel c*4 + (y%2)*2 + (x%2). One thread moves the four values of one (token, c): an 8-byte store on the packed side.
__global__ void pack_latents_kernel(const __nv_bfloat16* __restrict__ z, int C, int h, int w, __nv_bfloat16* __restrict__ out,
                                    int64_t ldo, int ch_off, int64_t total) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int w2 = w / 2, S = (h / 2) * w2;
    const int c = static_cast<int>(i % C);
    const int64_t bs = i / C;
    const int s = static_cast<int>(bs % S), b = static_cast<int>(bs / S);
    const int sy = s / w2, sx = s - sy * w2;
    const __nv_bfloat16* src = z + ((static_cast<size_t>(b) * C + c) * h + 2 * sy) * w + 2 * sx;
    const __nv_bfloat162 r0 = *reinterpret_cast<const __nv_bfloat162*>(src);
    const __nv_bfloat162 r1 = *reinterpret_cast<const __nv_bfloat162*>(src + w);
    __nv_bfloat162* dst = reinterpret_cast<__nv_bfloat162*>(out + static_cast<size_t>(bs) * ldo + ch_off + c * 4);
    dst[0] = r0;
    dst[1] = r1;
}
__global__ void unpack_latents_kernel(const __nv_bfloat16* __restrict__ x, int64_t ldx, int C, int h, int w,
                                      __nv_bfloat16* __restrict__ z, int64_t total) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int w2 = w / 2, h2 = h / 2;
    const int sx = static_cast<int>(i % w2);
    int64_t r = i / w2;
    const int sy = static_cast<int>(r % h2);
    r /= h2;
    const int c = static_cast<int>(r % C), b = static_cast<int>(r / C);
    const __nv_bfloat162* src = reinterpret_cast<const __nv_bfloat162*>(
        x + (static_cast<size_t>(b) * h2 * w2 + static_cast<size_t>(sy) * w2 + sx) * ldx + c * 4);
    __nv_bfloat16* dst = z + ((static_cast<size_t>(b) * C + c) * h + 2 * sy) * w + 2 * sx;
    *reinterpret_cast<__nv_bfloat162*>(dst) = src[0];
    *reinterpret_cast<__nv_bfloat162*>(dst + w) = src[1];
}
int pack_latents(const __nv_bfloat16* z, int B, int C, int h, int w, __nv_bfloat16* out, int64_t ldo, int ch_off,
                 cudaStream_t st) {
    DRAG_REQUIRE(z && out && B >= 1 && C >= 1 && h >= 2 && w >= 2 && h % 2 == 0 && w % 2 == 0 && ldo % 4 == 0 &&
                     ch_off % 4 == 0 && ch_off + 4 * C <= ldo, "pack_latents: bad arguments");
    const int64_t total = static_cast<int64_t>(B) * (h / 2) * (w / 2) * C;
    pack_latents_kernel<<<ceil_div(total, 256), 256, 0, st>>>(z, C, h, w, out, ldo, ch_off, total); count_launch();
    DRAG_CUDA(cudaGetLastError());
    return DRAG_OK;
}
int unpack_latents(const __nv_bfloat16* x, int64_t ldx, int B, int C, int h, int w, __nv_bfloat16* z, cudaStream_t st) {
    DRAG_REQUIRE(x && z && B >= 1 && C >= 1 && h >= 2 && w >= 2 && h % 2 == 0 && w % 2 == 0 && ldx % 4 == 0 && 4 * C <= ldx,
                 "unpack_latents: bad arguments");
    const int64_t total = static_cast<int64_t>(B) * C * (h / 2) * (w / 2);
    unpack_latents_kernel<<<ceil_div(total, 256), 256, 0, st>>>(x, ldx, C, h, w, z, total); count_launch();
    DRAG_CUDA(cudaGetLastError());
    return DRAG_OK;
}

// Flux-Fill transformer input (FluxFillPipeline.prepare_mask_latents + the channel concat of its denoising loop):
// x[b][s][0:64] = packed latents (copied when given), [64:128] = 2x2-packed masked-image latents, [128:384] = the 8x8 block
// of mask pixels under each latent pixel as 64 channels, 2x2-packed. 96 work items per token, four channels each.
__global__ void pack_fill_inputs_kernel(const __nv_bfloat16* __restrict__ latents, int64_t ld_lat,
                                        const __nv_bfloat16* __restrict__ masked, const uint8_t* __restrict__ mask, int h,
                                        int w, __nv_bfloat16* __restrict__ x, int64_t ldx, int64_t total) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int w2 = w / 2, S = (h / 2) * w2;
    const int q = static_cast<int>(i % 96);
    const int64_t bs = i / 96;
    const int s = static_cast<int>(bs % S), b = static_cast<int>(bs / S);
    const int sy = s / w2, sx = s - sy * w2;
    __nv_bfloat162* dst = reinterpret_cast<__nv_bfloat162*>(x + static_cast<size_t>(bs) * ldx + q * 4);
    if (q < 16) {
        if (latents == nullptr) return;
        const __nv_bfloat162* src = reinterpret_cast<const __nv_bfloat162*>(latents + static_cast<size_t>(bs) * ld_lat + q * 4);
        dst[0] = src[0];
        dst[1] = src[1];
    } else if (q < 32) {
        const int c = q - 16;
        const __nv_bfloat16* src = masked + ((static_cast<size_t>(b) * 16 + c) * h + 2 * sy) * w + 2 * sx;
        dst[0] = *reinterpret_cast<const __nv_bfloat162*>(src);
        dst[1] = *reinterpret_cast<const __nv_bfloat162*>(src + w);
    } else {
        const int m = q - 32, my = m >> 3, mx = m & 7;
        const size_t W = static_cast<size_t>(w) * 8;
        const uint8_t* src = mask + (static_cast<size_t>(b) * h * 8 + static_cast<size_t>(2 * sy) * 8 + my) * W +
                             static_cast<size_t>(2 * sx) * 8 + mx;
        const float v00 = src[0] ? 1.f : 0.f, v01 = src[8] ? 1.f : 0.f;
        const float v10 = src[8 * W] ? 1.f : 0.f, v11 = src[8 * W + 8] ? 1.f : 0.f;
        dst[0] = __floats2bfloat162_rn(v00, v01);
        dst[1] = __floats2bfloat162_rn(v10, v11);
    }
}
int pack_fill_inputs(const __nv_bfloat16* latents, int64_t ld_lat, const __nv_bfloat16* masked, const uint8_t* mask, int B,
                     int h, int w, __nv_bfloat16* x, int64_t ldx, cudaStream_t st) {
    DRAG_REQUIRE(masked && mask && x && B >= 1 && h >= 2 && w >= 2 && h % 2 == 0 && w % 2 == 0 && ldx >= 384 &&
                     ldx % 4 == 0 && (latents == nullptr || (ld_lat >= 64 && ld_lat % 4 == 0)),
                 "pack_fill_inputs: bad arguments");
    const int64_t total = static_cast<int64_t>(B) * (h / 2) * (w / 2) * 96;
    pack_fill_inputs_kernel<<<ceil_div(total, 256), 256, 0, st>>>(latents, ld_lat, masked, mask, h, w, x, ldx, total);
    count_launch();
    DRAG_CUDA(cudaGetLastError());
    return DRAG_OK;
}

// Redux prompt blend: out_embeds[t][c] = sum_b s_embed[b] * (t < n_txt ? txt[b][t][c] : img[b][t-n_txt][c]);
// out_pooled[c] = sum_b s_pool[b] * pooled[b][c]. Products and the running sum round to bf16 like
// the reference's bf16 tensor ops.
__global__ void redux_blend_kernel(const __nv_bfloat16* txt, const __nv_bfloat16* img, const __nv_bfloat16* pooled,
                                   const float* s_embed, const float* s_pool, __nv_bfloat16* out_embeds,
                                   __nv_bfloat16* out_pooled, int B, int n_txt, int n_img, int dim, int pooled_dim) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    const int64_t n_embed = static_cast<int64_t>(n_txt + n_img) * dim;
    if (i < n_embed) {
        const int t = static_cast<int>(i / dim), c = static_cast<int>(i - static_cast<int64_t>(t) * dim);
        float acc = 0.f;
        for (int b = 0; b < B; ++b) {
            const __nv_bfloat16 e = (t < n_txt) ? txt[(static_cast<size_t>(b) * n_txt + t) * dim + c]
                                                : img[(static_cast<size_t>(b) * n_img + (t - n_txt)) * dim + c];
            const float prod = __bfloat162float(__float2bfloat16(__bfloat162float(e) * __bfloat162float(__float2bfloat16(s_embed[b]))));
            acc = (b == 0) ? prod : __bfloat162float(__float2bfloat16(acc + prod));
        }
        out_embeds[i] = __float2bfloat16(acc);
    } else if (i < n_embed + pooled_dim) {
        const int c = static_cast<int>(i - n_embed);
        float acc = 0.f;
        for (int b = 0; b < B; ++b) {
            const float prod = __bfloat162float(__float2bfloat16(
                __bfloat162float(pooled[static_cast<size_t>(b) * pooled_dim + c]) * __bfloat162float(__float2bfloat16(s_pool[b]))));
            acc = (b == 0) ? prod : __bfloat162float(__float2bfloat16(acc + prod));
        }
        out_pooled[c] = __float2bfloat16(acc);
    }
}
int redux_blend(const __nv_bfloat16* txt, const __nv_bfloat16* img, const __nv_bfloat16* pooled, const float* s_embed,
                const float* s_pool, __nv_bfloat16* out_embeds, __nv_bfloat16* out_pooled, int B, int n_txt, int n_img,
                int dim, int pooled_dim, cudaStream_t st) {
    DRAG_REQUIRE(txt && img && pooled && s_embed && s_pool && out_embeds && out_pooled && B >= 1, "redux_blend: bad arguments");
    const int64_t n = static_cast<int64_t>(n_txt + n_img) * dim + pooled_dim;
    redux_blend_kernel<<<ceil_div(n, 256), 256, 0, st>>>(txt, img, pooled, s_embed, s_pool, out_embeds, out_pooled, B,
                                                         n_txt, n_img, dim, pooled_dim); count_launch();
    DRAG_CUDA(cudaGetLastError());
    return DRAG_OK;
}

// out[r] = x[r] / ||x[r]||_2  (fp32 in/out) - the caller-side normalisation of retrieval...:172.
__global__ void l2_normalize_kernel(const float* x, float* out, int rows, int d) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= rows) return;
    float ss = 0.f;
    for (int c = lane; c < d; c += 32) {
        const float v = x[static_cast<size_t>(row) * d + c];
        ss = fmaf(v, v, ss);
    }
    const float inv = 1.f / sqrtf(wsum(ss));
    for (int c = lane; c < d; c += 32) out[static_cast<size_t>(row) * d + c] = x[static_cast<size_t>(row) * d + c] * inv;
}
int l2_normalize(const float* x, float* out, int rows, int d, cudaStream_t st) {
    DRAG_REQUIRE(x && out && rows >= 1 && d >= 1, "l2_normalize: bad arguments");
    l2_normalize_kernel<<<ceil_div(rows, 8), 256, 0, st>>>(x, out, rows, d); count_launch();
    DRAG_CUDA(cudaGetLastError());
    return DRAG_OK;
}


// ---------------------------------------------------------------------------------------- ViT glue
// Patch extraction for the stride == kernel conv of the ViT stem (conv1 of OpenAI CLIP's
// VisionTransformer): img fp32 [B][3][R][R] -> patches bf16 [B*g*g][kpad], column (c, py, px),
// zero padded from 3*p*p to kpad so the GEMM's K is TMA friendly.
__global__ void vit_patchify_kernel(const float* __restrict__ img, __nv_bfloat16* __restrict__ out, int B, int R,
                                    int p, int g, int kpad) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    const int64_t total = static_cast<int64_t>(B) * g * g * kpad;
    if (i >= total) return;
    const int col = static_cast<int>(i % kpad);
    const int64_t row = i / kpad;
    const int gx = static_cast<int>(row % g), gy = static_cast<int>((row / g) % g), b = static_cast<int>(row / (g * g));
    float v = 0.f;
    if (col < 3 * p * p) {
        const int c = col / (p * p), r2 = col - c * p * p, py = r2 / p, px = r2 - py * p;
        v = img[((static_cast<size_t>(b) * 3 + c) * R + gy * p + py) * R + gx * p + px];
    }
    out[i] = __float2bfloat16(v);
}
int vit_patchify(const float* img, __nv_bfloat16* out, int B, int R, int p, int kpad, cudaStream_t st) {
    DRAG_REQUIRE(img && out && B >= 1 && p >= 1 && R >= p && kpad >= 3 * p * p && kpad % 8 == 0, "vit_patchify: bad arguments");
    // stride == kernel, no padding: floor(R / p) patches per side, trailing pixels unused (SigLIP 384 / 14 = 27)
    const int g = R / p;
    const int64_t total = static_cast<int64_t>(B) * g * g * kpad;
    vit_patchify_kernel<<<ceil_div(total, 256), 256, 0, st>>>(img, out, B, R, p, g, kpad); count_launch();
    DRAG_CUDA(cudaGetLastError());
    return DRAG_OK;
}

// x[b][0] = cls + pos[0]; x[b][1+i] = patch_emb[b][i] + pos[1+i]   (bf16, width w, L = n_patch + 1)
__global__ void vit_assemble_kernel(const __nv_bfloat16* __restrict__ pe, const __nv_bfloat16* __restrict__ cls,
                                    const __nv_bfloat16* __restrict__ pos, __nv_bfloat16* __restrict__ x, int B,
                                    int n_patch, int w) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    const int L = n_patch + 1, w8 = w / 8;
    if (i >= static_cast<int64_t>(B) * L * w8) return;
    const int c = static_cast<int>(i % w8) * 8;
    const int t = static_cast<int>((i / w8) % L), b = static_cast<int>(i / (static_cast<int64_t>(w8) * L));
    float a[8], pz[8];
    if (t == 0) ld8(cls + c, a);
    else ld8(pe + (static_cast<size_t>(b) * n_patch + t - 1) * w + c, a);
    ld8(pos + static_cast<size_t>(t) * w + c, pz);
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] += pz[j];
    st8(x + (static_cast<size_t>(b) * L + t) * w + c, a);
}
int vit_assemble(const __nv_bfloat16* patch_emb, const __nv_bfloat16* cls, const __nv_bfloat16* pos, __nv_bfloat16* x,
                 int B, int n_patch, int w, cudaStream_t st) {
    DRAG_REQUIRE(patch_emb && cls && pos && x && B >= 1 && w % 8 == 0, "vit_assemble: bad arguments");
    const int64_t total = static_cast<int64_t>(B) * (n_patch + 1) * (w / 8);
    vit_assemble_kernel<<<ceil_div(total, 256), 256, 0, st>>>(patch_emb, cls, pos, x, B, n_patch, w); count_launch();
    DRAG_CUDA(cudaGetLastError());
    return DRAG_OK;
}

}  // namespace drag
