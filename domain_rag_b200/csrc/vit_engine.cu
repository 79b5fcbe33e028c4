// CLIP ViT image tower orchestrated in C++ over the tcgen05 GEMM, the head-dim-64 tcgen05 attention and the row kernels:
// the `model.encode_image(x)` + L2-normalise of the reference's retrieval loops (retrieval/clip100_resnet_style_all_shots.py:
// 161-177, 270-296, 326-349; clip.load at :209). One C call per batch instead of ~15 Python -> ctypes launches per
// transformer block (ViT-B/32 jobs were launch-bound from Python: 3 400 launches for a 183 ms job).
//
// Input: either the normalised float tensor `preprocess` produces (fp32 [B][3][R][R], the reference contract) or the raw
// uint8 pixels [B][3][R][R] (a quarter of the PCIe bytes) - then ToTensor + Normalize run inside the patch
// extraction kernel with the same two IEEE divisions torchvision performs ((u8 / 255 - mean) / std), so both paths feed
// identical bf16 patches to the GEMM.
// Layout (bf16): patches [B*g*g][kpad] -> tokens h [B][L][w], L = g*g + 1 (class token first); q/k/v [B][H][L][64];
// MLP hidden [B*L][4w]. The workspace is sized for `max_batch` images; larger calls are processed in chunks.
#include <vector>

#include "common.cuh"
#include "flux_ops.cuh"
#include "gemm.cuh"
#include "vit_engine.cuh"

namespace drag {

// uint8 [B][3][R][R] -> bf16 patches [B*g*g][kpad] with ToTensor + Normalize fused (fp32, IEEE division like torchvision)
__global__ void vit_patchify_u8_kernel(const uint8_t* __restrict__ img, __nv_bfloat16* __restrict__ out, int B, int R,
                                       int p, int g, int kpad, float m0, float m1, float m2, float s0, float s1, float s2) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    const int64_t total = static_cast<int64_t>(B) * g * g * kpad;
    if (i >= total) return;
    const int col = static_cast<int>(i % kpad);
    const int64_t row = i / kpad;
    const int gx = static_cast<int>(row % g), gy = static_cast<int>((row / g) % g), b = static_cast<int>(row / (g * g));
    float v = 0.f;
    if (col < 3 * p * p) {
        const int c = col / (p * p), r2 = col - c * p * p, py = r2 / p, px = r2 - py * p;
        const float u = static_cast<float>(img[((static_cast<size_t>(b) * 3 + c) * R + gy * p + py) * R + gx * p + px]);
        const float mean = c == 0 ? m0 : (c == 1 ? m1 : m2), sd = c == 0 ? s0 : (c == 1 ? s1 : s2);
        v = __fdiv_rn(__fsub_rn(__fdiv_rn(u, 255.f), mean), sd);
    }
    out[i] = __float2bfloat16(v);
}

// (sum, sum of squares) of each bf16 row into part 0 of the row's partial-moment slots, the other parts zeroed: the
// input of the first block's folded LayerNorm (later blocks get their moments from the producing GEMM's epilogue).
__global__ void vit_row_stats_kernel(const __nv_bfloat16* __restrict__ x, int ld, int M, int d, int parts,
                                     float* __restrict__ stats) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= M) return;
    float s = 0.f, ss = 0.f;
    for (int c = lane * 2; c < d; c += 64) {
        const float2 v = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(x + static_cast<size_t>(row) * ld + c));
        s += v.x + v.y;
        ss = fmaf(v.x, v.x, fmaf(v.y, v.y, ss));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o);
        ss += __shfl_xor_sync(0xffffffffu, ss, o);
    }
    float2* dst = reinterpret_cast<float2*>(stats) + static_cast<size_t>(row) * parts;
    for (int p = lane; p < parts; p += 32) dst[p] = (p == 0) ? make_float2(s, ss) : make_float2(0.f, 0.f);
}

static int alloc_bytes(void** p, size_t n) {
    DRAG_CUDA(cudaMalloc(p, n));
    return DRAG_OK;
}

int vit_create(const VitCfg& cfg, VitEngine** out) {
    DRAG_REQUIRE(out, "vit_create: null out");
    DRAG_REQUIRE(cfg.width % 256 == 0 && cfg.heads >= 1 && cfg.width == cfg.heads * 64,
                 "vit_create: width must be heads * 64 and a multiple of 256 (CLIP ViT towers: 768, 1024)");
    DRAG_REQUIRE(cfg.layers >= 1 && cfg.patch >= 1 && cfg.image >= cfg.patch && cfg.out_dim % 32 == 0 && cfg.max_batch >= 1,
                 "vit_create: bad configuration");
    VitEngine* e = new VitEngine();
    e->cfg = cfg;
    e->grid = cfg.image / cfg.patch;
    e->tokens = e->grid * e->grid + 1;
    e->kpad = (3 * cfg.patch * cfg.patch + 7) / 8 * 8;
    e->blocks.resize(cfg.layers);
    const size_t B = cfg.max_batch, L = e->tokens, w = cfg.width, np = static_cast<size_t>(e->grid) * e->grid;
    int rc = 0;
    rc |= alloc_bytes(reinterpret_cast<void**>(&e->patches), B * np * e->kpad * 2);
    rc |= alloc_bytes(reinterpret_cast<void**>(&e->pe), B * np * w * 2);
    rc |= alloc_bytes(reinterpret_cast<void**>(&e->h), B * L * w * 2);
    rc |= alloc_bytes(reinterpret_cast<void**>(&e->y), B * L * w * 2);
    rc |= alloc_bytes(reinterpret_cast<void**>(&e->q), B * L * w * 2);
    rc |= alloc_bytes(reinterpret_cast<void**>(&e->k), B * L * w * 2);
    rc |= alloc_bytes(reinterpret_cast<void**>(&e->v), B * L * w * 2);
    rc |= alloc_bytes(reinterpret_cast<void**>(&e->a), B * L * w * 2);
    rc |= alloc_bytes(reinterpret_cast<void**>(&e->u), B * L * 4 * w * 2);
    rc |= alloc_bytes(reinterpret_cast<void**>(&e->cls_ln), B * w * 2);
    rc |= alloc_bytes(reinterpret_cast<void**>(&e->emb), B * cfg.out_dim * 4);
    rc |= alloc_bytes(reinterpret_cast<void**>(&e->stats_a), B * L * (w / 128) * 2 * 4);
    rc |= alloc_bytes(reinterpret_cast<void**>(&e->stats_b), B * L * (w / 128) * 2 * 4);
    if (rc) {
        vit_destroy(e);
        return fail(DRAG_ERR_CUDA, "vit_create: workspace allocation failed");
    }
    *out = e;
    return DRAG_OK;
}

int vit_destroy(VitEngine* e) {
    if (!e) return DRAG_OK;
    void* bufs[] = {e->patches, e->pe, e->h, e->y, e->q, e->k, e->v, e->a, e->u, e->cls_ln, e->emb, e->stats_a, e->stats_b};
    for (void* b : bufs)
        if (b) cudaFree(b);
    delete e;
    return DRAG_OK;
}

int vit_set_weights(VitEngine* e, const void* const* ptrs, int n) {
    DRAG_REQUIRE(e && ptrs, "vit_set_weights: null pointer");
    const int expect = 8 + 18 * e->cfg.layers;
    DRAG_REQUIRE(n == expect, "vit_set_weights: expected " + std::to_string(expect) + " pointers");
    for (int j = 0; j < n; ++j) DRAG_REQUIRE(ptrs[j], "vit_set_weights: null weight pointer at slot " + std::to_string(j));
    int i = 0;
    auto next = [&]() { return static_cast<const __nv_bfloat16*>(ptrs[i++]); };
    e->conv_w = next(); e->cls = next(); e->pos = next();
    e->ln_pre_w = next(); e->ln_pre_b = next(); e->ln_post_w = next(); e->ln_post_b = next(); e->proj_t = next();
    for (VitBlockW& b : e->blocks) {
        b.ln1_w = next(); b.ln1_b = next(); b.qkv_w = next(); b.qkv_b = next(); b.out_w = next(); b.out_b = next();
        b.ln2_w = next(); b.ln2_b = next(); b.fc_w = next(); b.fc_b = next(); b.proj_w = next(); b.proj_b = next();
        b.qkv_wf = next(); b.qkv_s = reinterpret_cast<const float*>(next()); b.qkv_c = reinterpret_cast<const float*>(next());
        b.fc_wf = next(); b.fc_s = reinterpret_cast<const float*>(next()); b.fc_c = reinterpret_cast<const float*>(next());
    }
    e->weights_set = true;
    return DRAG_OK;
}

#define VX(call)             \
    do {                     \
        int _rc = (call);    \
        if (_rc) return _rc; \
    } while (0)

static int lin(const __nv_bfloat16* A, int lda, const __nv_bfloat16* W, int K, int M, int N, const __nv_bfloat16* bias,
               int mode, __nv_bfloat16* out, int ldo, const __nv_bfloat16* resid, cudaStream_t st) {
    GemmEpi ep;
    ep.mode = mode;
    ep.bias = bias;
    ep.out = out;
    ep.ldo = ldo;
    ep.resid = resid;
    ep.ldr = ldo;
    return gemm_bf16(A, lda, W, K, M, N, K, ep, st);
}

static int encode_chunk(VitEngine* e, const void* img, int img_kind, int B, float* out, int normalize, cudaStream_t st) {
    const VitCfg& c = e->cfg;
    const int w = c.width, H = c.heads, L = e->tokens, g = e->grid, np = g * g, M = B * L;
    if (img_kind == 0) {
        VX(vit_patchify(static_cast<const float*>(img), e->patches, B, c.image, c.patch, e->kpad, st));
    } else {
        const int64_t total = static_cast<int64_t>(B) * np * e->kpad;
        vit_patchify_u8_kernel<<<ceil_div(total, 256), 256, 0, st>>>(static_cast<const uint8_t*>(img), e->patches, B, c.image,
                                                                     c.patch, g, e->kpad, c.mean[0], c.mean[1], c.mean[2],
                                                                     c.std[0], c.std[1], c.std[2]);
        count_launch();
        DRAG_CUDA(cudaGetLastError());
    }
    VX(lin(e->patches, e->kpad, e->conv_w, e->kpad, B * np, w, nullptr, EPI_BIAS, e->pe, w, nullptr, st));
    VX(vit_assemble(e->pe, e->cls, e->pos, e->h, B, np, w, st));
    VX(layernorm_bf16(e->h, w, e->h, w, M, w, e->ln_pre_w, 0, e->ln_pre_b, 0, 0, 0, 1e-5f, st));
    const int parts = w / 128;          // the N = w GEMMs run 256-wide tiles: one partial per 128 columns (BN / 2)
    if (e->fold_ln) {
        vit_row_stats_kernel<<<ceil_div(M, 8), 256, 0, st>>>(e->h, w, M, w, parts, e->stats_a);
        count_launch();
        DRAG_CUDA(cudaGetLastError());
    }
    for (const VitBlockW& b : e->blocks) {
        GemmEpi qe;
        qe.mode = EPI_QKV_SPLIT;
        qe.q_out = e->q; qe.k_out = e->k; qe.v_out = e->v;
        qe.heads = H;
        qe.head_dim = 64;
        qe.s_total = L;
        qe.tok_offset = 0;
        qe.rows_per_batch = L;
        if (e->fold_ln) {
            // ln_1 lives in the QKV GEMM's epilogue (moments of h from the previous producer), ln_2 in the MLP-up GEMM's;
            // the two residual GEMMs emit the moments of the rows they store
            qe.ln_stats = e->stats_a; qe.ln_parts = parts; qe.ln_k = w; qe.ln_s = b.qkv_s; qe.ln_c = b.qkv_c; qe.ln_eps = 1e-5f;
            VX(gemm_bf16(e->h, w, b.qkv_wf, w, M, 3 * w, w, qe, st));
            VX(attention_bf16(e->q, e->k, e->v, B, H, L, 64, 0, nullptr, 8, e->a, w, st));
            GemmEpi oe;
            oe.mode = EPI_GATE_RESID; oe.bias = b.out_b; oe.out = e->h; oe.ldo = w; oe.resid = e->h; oe.ldr = w;
            oe.stats_out = e->stats_b; oe.stats_parts = parts;
            VX(gemm_bf16(e->a, w, b.out_w, w, M, w, w, oe, st));
            GemmEpi fe;
            fe.mode = EPI_QUICK_GELU; fe.out = e->u; fe.ldo = 4 * w;
            fe.ln_stats = e->stats_b; fe.ln_parts = parts; fe.ln_k = w; fe.ln_s = b.fc_s; fe.ln_c = b.fc_c; fe.ln_eps = 1e-5f;
            VX(gemm_bf16(e->h, w, b.fc_wf, w, M, 4 * w, w, fe, st));
            GemmEpi pe2;
            pe2.mode = EPI_GATE_RESID; pe2.bias = b.proj_b; pe2.out = e->h; pe2.ldo = w; pe2.resid = e->h; pe2.ldr = w;
            pe2.stats_out = e->stats_a; pe2.stats_parts = parts;
            VX(gemm_bf16(e->u, 4 * w, b.proj_w, 4 * w, M, w, 4 * w, pe2, st));
        } else {
            VX(layernorm_bf16(e->h, w, e->y, w, M, w, b.ln1_w, 0, b.ln1_b, 0, 0, 0, 1e-5f, st));
            qe.bias = b.qkv_b;
            VX(gemm_bf16(e->y, w, b.qkv_w, w, M, 3 * w, w, qe, st));
            VX(attention_bf16(e->q, e->k, e->v, B, H, L, 64, 0, nullptr, 8, e->a, w, st));
            VX(lin(e->a, w, b.out_w, w, M, w, b.out_b, EPI_GATE_RESID, e->h, w, e->h, st));
            VX(layernorm_bf16(e->h, w, e->y, w, M, w, b.ln2_w, 0, b.ln2_b, 0, 0, 0, 1e-5f, st));
            VX(lin(e->y, w, b.fc_w, w, M, 4 * w, b.fc_b, EPI_QUICK_GELU, e->u, 4 * w, nullptr, st));
            VX(lin(e->u, 4 * w, b.proj_w, 4 * w, M, w, b.proj_b, EPI_GATE_RESID, e->h, w, e->h, st));
        }
    }
    // class-token rows (stride L*w) -> ln_post -> projection (fp32 out) -> optional L2 normalise
    VX(layernorm_bf16(e->h, L * w, e->cls_ln, w, B, w, e->ln_post_w, 0, e->ln_post_b, 0, 0, 0, 1e-5f, st));
    GemmEpi pe;
    pe.mode = EPI_BIAS_F32;
    pe.out_f32 = normalize ? e->emb : out;
    pe.ldo = c.out_dim;
    VX(gemm_bf16(e->cls_ln, w, e->proj_t, w, B, c.out_dim, w, pe, st));
    if (normalize) VX(l2_normalize(e->emb, out, B, c.out_dim, st));
    return DRAG_OK;
}

int vit_set_option(VitEngine* e, int key, int value) {
    DRAG_REQUIRE(e, "vit_set_option: null engine");
    DRAG_REQUIRE(key == 1, "vit_set_option: unknown key (1 = fold LayerNorm into the GEMMs: 1 on, 0 off)");
    e->fold_ln = value ? 1 : 0;
    return DRAG_OK;
}

int vit_encode(VitEngine* e, const void* img, int img_kind, int B, float* out, int normalize, cudaStream_t st) {
    DRAG_REQUIRE(e && e->weights_set, "vit_encode: weights not set");
    DRAG_REQUIRE(img && out && B >= 0, "vit_encode: bad arguments");
    DRAG_REQUIRE(img_kind == 0 || img_kind == 1, "vit_encode: img_kind must be 0 (fp32 normalised) or 1 (uint8 pixels)");
    const size_t per_img = static_cast<size_t>(3) * e->cfg.image * e->cfg.image * (img_kind == 0 ? 4 : 1);
    for (int b0 = 0; b0 < B; b0 += e->cfg.max_batch) {
        const int nb = (B - b0 < e->cfg.max_batch) ? (B - b0) : e->cfg.max_batch;
        VX(encode_chunk(e, static_cast<const uint8_t*>(img) + per_img * b0, img_kind, nb,
                        out + static_cast<size_t>(b0) * e->cfg.out_dim, normalize, st));
    }
    return DRAG_OK;
}

}  // namespace drag
