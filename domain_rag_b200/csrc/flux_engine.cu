// Flux MMDiT forward (19 double + 38 single blocks for FLUX.1-dev / Fill-dev) orchestrated in C++ over
// the tcgen05 GEMM, the tcgen05 attention and the row kernels - the FluxTransformer2DModel.forward
// that diffusers' pipelines call once per denoising step (reference call sites:
// batch_generate_flux_kshot.py:467-474, outpainting_updown_sampling_redux.py:1246-1257).
//
// Data layout in HBM (bf16): img stream [B*S_img][d], txt stream [B*S_txt][d] through the double
// blocks; joint z [B][S_txt+S_img][d] through the single blocks; q/k/v [B][H][S][128] written directly
// by the QKV GEMM epilogue (RMSNorm + RoPE fused, text tokens first); attention output written
// token-major so that it is the A operand of the following projection (single blocks: columns
// [0,d) of the [M][5d] buffer whose columns [d,5d) receive GELU(mlp) - concat-free).
// Every step launches: 1 modulation GEMM for all 57 blocks + the final layer (temb is token
// independent), then per double block 2 x (LN-mod, QKV GEMM, out GEMM, LN-mod, MLP-up, MLP-down) + 1
// attention; per single block LN-mod, QKV GEMM, MLP GEMM, attention, out GEMM.
#include <vector>

#include "common.cuh"
#include "flux_engine.cuh"
#include "flux_ops.cuh"
#include "gemm.cuh"

namespace drag {

static int alloc_bf16(__nv_bfloat16** p, size_t n) {
    DRAG_CUDA(cudaMalloc(p, n * sizeof(__nv_bfloat16)));
    return DRAG_OK;
}

int flux_create(const FluxCfg& cfg, FluxEngine** out) {
    DRAG_REQUIRE(out, "flux_create: null out");
    DRAG_REQUIRE(cfg.d == cfg.heads * 128, "flux_create: d must equal heads*128");
    DRAG_REQUIRE(cfg.d % 256 == 0 && cfg.in_channels % 8 == 0 && cfg.txt_dim % 8 == 0 && cfg.pooled_dim % 8 == 0 &&
                     cfg.out_channels % 32 == 0, "flux_create: unsupported dimensions");
    DRAG_REQUIRE(cfg.max_batch >= 1 && cfg.max_img_tokens >= 1 && cfg.txt_tokens >= 1, "flux_create: bad capacity");
    FluxEngine* e = new FluxEngine();
    e->cfg = cfg;
    e->dimg.resize(cfg.n_double);
    e->dtxt.resize(cfg.n_double);
    e->single.resize(cfg.n_single);
    const size_t d = cfg.d, B = cfg.max_batch, Si = cfg.max_img_tokens, St = cfg.txt_tokens, S = Si + St;
    e->n_mod = static_cast<size_t>(cfg.n_double) * 12 * d + static_cast<size_t>(cfg.n_single) * 3 * d + 2 * d;
    int rc = 0;
    rc |= alloc_bf16(&e->img, B * Si * d);
    rc |= alloc_bf16(&e->txt, B * St * d);
    rc |= alloc_bf16(&e->z, B * S * d);
    rc |= alloc_bf16(&e->h, B * S * d);
    rc |= alloc_bf16(&e->q, B * S * d);
    rc |= alloc_bf16(&e->k, B * S * d);
    rc |= alloc_bf16(&e->v, B * S * d);
    rc |= alloc_bf16(&e->attn_img, B * Si * d);
    rc |= alloc_bf16(&e->attn_txt, B * St * d);
    rc |= alloc_bf16(&e->wide, B * S * 5 * d);
    rc |= alloc_bf16(&e->mod, B * e->n_mod);
    rc |= alloc_bf16(&e->temb, B * 256 * 2);
    rc |= alloc_bf16(&e->vec_tmp, B * d * 8);
    if (rc) {
        delete e;
        return fail(DRAG_ERR_CUDA, "flux_create: workspace allocation failed");
    }
    *out = e;
    return DRAG_OK;
}

int flux_destroy(FluxEngine* e) {
    if (!e) return DRAG_OK;
    __nv_bfloat16* bufs[] = {e->img, e->txt, e->z, e->h, e->q, e->k, e->v, e->attn_img, e->attn_txt, e->wide,
                             e->mod, e->temb, e->vec_tmp};
    for (auto* b : bufs)
        if (b) cudaFree(b);
    delete e;
    return DRAG_OK;
}

int flux_set_weights(FluxEngine* e, const void* const* ptrs, int n) {
    DRAG_REQUIRE(e && ptrs, "flux_set_weights: null pointer");
    const int expect = 20 + 20 * e->cfg.n_double + 8 * e->cfg.n_single;
    DRAG_REQUIRE(n == expect, "flux_set_weights: expected " + std::to_string(expect) + " pointers");
    int i = 0;
    auto next = [&]() { return static_cast<const __nv_bfloat16*>(ptrs[i++]); };
    e->x_in_w = next(); e->x_in_b = next(); e->ctx_in_w = next(); e->ctx_in_b = next();
    e->t_w1 = next(); e->t_b1 = next(); e->t_w2 = next(); e->t_b2 = next();
    e->g_w1 = next(); e->g_b1 = next(); e->g_w2 = next(); e->g_b2 = next();
    e->p_w1 = next(); e->p_b1 = next(); e->p_w2 = next(); e->p_b2 = next();
    e->mod_w = next(); e->mod_b = next(); e->final_w = next(); e->final_b = next();
    for (int b = 0; b < e->cfg.n_double; ++b) {
        for (FluxStreamW* s : {&e->dimg[b], &e->dtxt[b]}) {
            s->qkv_w = next(); s->qkv_b = next(); s->qnorm = next(); s->knorm = next();
            s->out_w = next(); s->out_b = next(); s->mlp1_w = next(); s->mlp1_b = next();
            s->mlp2_w = next(); s->mlp2_b = next();
        }
    }
    for (int b = 0; b < e->cfg.n_single; ++b) {
        FluxSingleW& s = e->single[b];
        s.qkv_w = next(); s.qkv_b = next(); s.qnorm = next(); s.knorm = next();
        s.mlp_w = next(); s.mlp_b = next(); s.out_w = next(); s.out_b = next();
    }
    for (int j = 0; j < n; ++j) {
        const bool optional = (j >= 8 && j < 12 && !e->cfg.guidance);
        DRAG_REQUIRE(ptrs[j] != nullptr || optional, "flux_set_weights: null weight pointer at slot " + std::to_string(j));
    }
    e->weights_set = true;
    return DRAG_OK;
}

#define FX(call)                 \
    do {                         \
        int _rc = (call);        \
        if (_rc) return _rc;     \
    } while (0)

static int linear(const __nv_bfloat16* A, int lda, const __nv_bfloat16* W, int K, int M, int N, const __nv_bfloat16* bias,
                  int mode, __nv_bfloat16* out, int ldo, cudaStream_t st) {
    GemmEpi ep;
    ep.mode = mode;
    ep.bias = bias;
    ep.out = out;
    ep.ldo = ldo;
    return gemm_bf16(A, lda, W, K, M, N, K, ep, st);
}

static int gated_linear(const __nv_bfloat16* A, int lda, const __nv_bfloat16* W, int K, int M, int N,
                        const __nv_bfloat16* bias, __nv_bfloat16* x, int ldx, const __nv_bfloat16* gate, int gate_ld,
                        int rows_per_batch, cudaStream_t st) {
    GemmEpi ep;
    ep.mode = EPI_GATE_RESID;
    ep.bias = bias;
    ep.out = x;
    ep.ldo = ldx;
    ep.resid = x;
    ep.ldr = ldx;
    ep.gate = gate;
    ep.gate_ld = gate_ld;
    ep.rows_per_batch = rows_per_batch;
    return gemm_bf16(A, lda, W, K, M, N, K, ep, st);
}

static int qkv_linear(const FluxEngine* e, const __nv_bfloat16* A, const __nv_bfloat16* W, const __nv_bfloat16* bias,
                      const __nv_bfloat16* qn, const __nv_bfloat16* kn, int M, int rows_per_batch, int tok_offset,
                      int S, const float* cosv, const float* sinv, cudaStream_t st) {
    GemmEpi ep;
    ep.mode = EPI_QKV_ROPE;
    ep.bias = bias;
    ep.q_out = e->q; ep.k_out = e->k; ep.v_out = e->v;
    ep.q_norm_w = qn; ep.k_norm_w = kn;
    ep.rope_cos = cosv; ep.rope_sin = sinv;
    ep.heads = e->cfg.heads;
    ep.s_total = S;
    ep.tok_offset = tok_offset;
    ep.rows_per_batch = rows_per_batch;
    ep.rms_eps = 1e-6f;
    return gemm_bf16(A, e->cfg.d, W, e->cfg.d, M, 3 * e->cfg.d, e->cfg.d, ep, st);
}

// x [B*S_img][ldx] (first in_channels columns used), ctx [B*S_txt][txt_dim], pooled [B][pooled_dim],
// t_dev/g_dev fp32 [B] on the device, rope tables fp32 [S_txt+S_img][64]; v_out bf16 [B*S_img][out_channels].
int flux_forward(FluxEngine* e, const __nv_bfloat16* x, int ldx, const __nv_bfloat16* ctx, const __nv_bfloat16* pooled,
                 const float* t_dev, const float* g_dev, const float* rope_cos, const float* rope_sin, int B, int S_img,
                 __nv_bfloat16* v_out, int ldv, int n_double_run, int n_single_run, cudaStream_t st) {
    DRAG_REQUIRE(e && e->weights_set, "flux_forward: weights not set");
    DRAG_REQUIRE(x && ctx && pooled && t_dev && rope_cos && rope_sin && v_out, "flux_forward: null pointer");
    const FluxCfg& c = e->cfg;
    DRAG_REQUIRE(B >= 1 && B <= c.max_batch && S_img >= 1 && S_img <= c.max_img_tokens, "flux_forward: exceeds capacity");
    DRAG_REQUIRE(!c.guidance || g_dev, "flux_forward: guidance model needs g_dev");
    const int d = c.d, St = c.txt_tokens, Si = S_img, S = St + Si, H = c.heads;
    const int Mi = B * Si, Mt = B * St, M = B * S;
    const int n_mod = static_cast<int>(e->n_mod);
    const int nd = (n_double_run < 0 || n_double_run > c.n_double) ? c.n_double : n_double_run;
    const int ns = (n_single_run < 0 || n_single_run > c.n_single) ? c.n_single : n_single_run;

    // ---- conditioning vector and every block's modulation in one GEMM
    __nv_bfloat16* t_emb = e->temb;
    __nv_bfloat16* g_emb = e->temb + static_cast<size_t>(B) * 256;
    __nv_bfloat16* t1 = e->vec_tmp;
    __nv_bfloat16* t2 = t1 + static_cast<size_t>(B) * d;
    __nv_bfloat16* g1 = t2 + static_cast<size_t>(B) * d;
    __nv_bfloat16* g2 = g1 + static_cast<size_t>(B) * d;
    __nv_bfloat16* p1 = g2 + static_cast<size_t>(B) * d;
    __nv_bfloat16* p2 = p1 + static_cast<size_t>(B) * d;
    __nv_bfloat16* vec = p2 + static_cast<size_t>(B) * d;
    FX(timestep_embed(t_dev, t_emb, B, st));
    FX(linear(t_emb, 256, e->t_w1, 256, B, d, e->t_b1, EPI_SILU, t1, d, st));
    FX(linear(t1, d, e->t_w2, d, B, d, e->t_b2, EPI_BIAS, t2, d, st));
    if (c.guidance) {
        FX(timestep_embed(g_dev, g_emb, B, st));
        FX(linear(g_emb, 256, e->g_w1, 256, B, d, e->g_b1, EPI_SILU, g1, d, st));
        FX(linear(g1, d, e->g_w2, d, B, d, e->g_b2, EPI_BIAS, g2, d, st));
    }
    FX(linear(pooled, c.pooled_dim, e->p_w1, c.pooled_dim, B, d, e->p_b1, EPI_SILU, p1, d, st));
    FX(linear(p1, d, e->p_w2, d, B, d, e->p_b2, EPI_BIAS, p2, d, st));
    FX(sum_silu(t2, c.guidance ? g2 : nullptr, p2, vec, B * d, 1, st));
    FX(linear(vec, d, e->mod_w, d, B, n_mod, e->mod_b, EPI_BIAS, e->mod, n_mod, st));

    // ---- input embedders
    FX(linear(x, ldx, e->x_in_w, c.in_channels, Mi, d, e->x_in_b, EPI_BIAS, e->img, d, st));
    FX(linear(ctx, c.txt_dim, e->ctx_in_w, c.txt_dim, Mt, d, e->ctx_in_b, EPI_BIAS, e->txt, d, st));

    // ---- double-stream blocks
    for (int i = 0; i < nd; ++i) {
        const __nv_bfloat16* mi = e->mod + static_cast<size_t>(i) * 12 * d;   // img: shift1 scale1 gate1 shift2 scale2 gate2
        const __nv_bfloat16* mt = mi + 6 * d;                                  // txt: same order
        const FluxStreamW& wi = e->dimg[i];
        const FluxStreamW& wt = e->dtxt[i];
        __nv_bfloat16* h_img = e->h;
        __nv_bfloat16* h_txt = e->h + static_cast<size_t>(Mi) * d;
        FX(layernorm_bf16(e->img, d, h_img, d, Mi, d, mi + d, n_mod, mi, n_mod, 1, Si, 1e-6f, st));
        FX(layernorm_bf16(e->txt, d, h_txt, d, Mt, d, mt + d, n_mod, mt, n_mod, 1, St, 1e-6f, st));
        FX(qkv_linear(e, h_txt, wt.qkv_w, wt.qkv_b, wt.qnorm, wt.knorm, Mt, St, 0, S, rope_cos, rope_sin, st));
        FX(qkv_linear(e, h_img, wi.qkv_w, wi.qkv_b, wi.qnorm, wi.knorm, Mi, Si, St, S, rope_cos, rope_sin, st));
        FX(attention_bf16(e->q, e->k, e->v, B, H, S, 128, St, e->attn_txt, d, e->attn_img, d, st));
        // image stream
        FX(gated_linear(e->attn_img, d, wi.out_w, d, Mi, d, wi.out_b, e->img, d, mi + 2 * d, n_mod, Si, st));
        FX(layernorm_bf16(e->img, d, h_img, d, Mi, d, mi + 4 * d, n_mod, mi + 3 * d, n_mod, 1, Si, 1e-6f, st));
        FX(linear(h_img, d, wi.mlp1_w, d, Mi, 4 * d, wi.mlp1_b, EPI_GELU_TANH, e->wide, 4 * d, st));
        FX(gated_linear(e->wide, 4 * d, wi.mlp2_w, 4 * d, Mi, d, wi.mlp2_b, e->img, d, mi + 5 * d, n_mod, Si, st));
        // text stream
        FX(gated_linear(e->attn_txt, d, wt.out_w, d, Mt, d, wt.out_b, e->txt, d, mt + 2 * d, n_mod, St, st));
        FX(layernorm_bf16(e->txt, d, h_txt, d, Mt, d, mt + 4 * d, n_mod, mt + 3 * d, n_mod, 1, St, 1e-6f, st));
        FX(linear(h_txt, d, wt.mlp1_w, d, Mt, 4 * d, wt.mlp1_b, EPI_GELU_TANH, e->wide, 4 * d, st));
        FX(gated_linear(e->wide, 4 * d, wt.mlp2_w, 4 * d, Mt, d, wt.mlp2_b, e->txt, d, mt + 5 * d, n_mod, St, st));
    }

    // ---- z = cat(txt, img) along the sequence, per batch element
    DRAG_CUDA(cudaMemcpy2DAsync(e->z, static_cast<size_t>(S) * d * 2, e->txt, static_cast<size_t>(St) * d * 2,
                                static_cast<size_t>(St) * d * 2, B, cudaMemcpyDeviceToDevice, st));
    DRAG_CUDA(cudaMemcpy2DAsync(e->z + static_cast<size_t>(St) * d, static_cast<size_t>(S) * d * 2, e->img,
                                static_cast<size_t>(Si) * d * 2, static_cast<size_t>(Si) * d * 2, B,
                                cudaMemcpyDeviceToDevice, st));

    // ---- single-stream blocks
    const __nv_bfloat16* mod_single = e->mod + static_cast<size_t>(c.n_double) * 12 * d;
    for (int i = 0; i < ns; ++i) {
        const __nv_bfloat16* ms = mod_single + static_cast<size_t>(i) * 3 * d;   // shift scale gate
        const FluxSingleW& w = e->single[i];
        FX(layernorm_bf16(e->z, d, e->h, d, M, d, ms + d, n_mod, ms, n_mod, 1, S, 1e-6f, st));
        FX(qkv_linear(e, e->h, w.qkv_w, w.qkv_b, w.qnorm, w.knorm, M, S, 0, S, rope_cos, rope_sin, st));
        FX(linear(e->h, d, w.mlp_w, d, M, 4 * d, w.mlp_b, EPI_GELU_TANH, e->wide + d, 5 * d, st));
        FX(attention_bf16(e->q, e->k, e->v, B, H, S, 128, 0, nullptr, 8, e->wide, 5 * d, st));
        FX(gated_linear(e->wide, 5 * d, w.out_w, 5 * d, M, d, w.out_b, e->z, d, ms + 2 * d, n_mod, S, st));
    }

    // ---- final layer on the image rows: LN * (1 + scale) + shift, then Linear(d -> out_channels)
    const __nv_bfloat16* mf = e->mod + (static_cast<size_t>(c.n_double) * 12 + static_cast<size_t>(c.n_single) * 3) * d;
    for (int b = 0; b < B; ++b) {
        const __nv_bfloat16* zr = e->z + (static_cast<size_t>(b) * S + St) * d;
        FX(layernorm_bf16(zr, d, e->h + static_cast<size_t>(b) * Si * d, d, Si, d, mf + static_cast<size_t>(b) * n_mod, n_mod,
                          mf + d + static_cast<size_t>(b) * n_mod, n_mod, 1, 1 << 30, 1e-6f, st));
    }
    FX(linear(e->h, d, e->final_w, d, Mi, c.out_channels, e->final_b, EPI_BIAS, v_out, ldv, st));
    return DRAG_OK;
}

}  // namespace drag
