// Host-side state of one Flux MMDiT engine (weights are caller-owned device pointers; the engine
// owns only its activation workspace).
#pragma once
#include <cuda_bf16.h>

#include <vector>

#include "common.cuh"

namespace drag {

struct FluxCfg {
    int in_channels, d, heads, n_double, n_single, txt_dim, pooled_dim, out_channels, guidance;
    int max_batch, max_img_tokens, txt_tokens;
};

struct FluxStreamW {
    const __nv_bfloat16 *qkv_w, *qkv_b, *qnorm, *knorm, *out_w, *out_b, *mlp1_w, *mlp1_b, *mlp2_w, *mlp2_b;
};
struct FluxSingleW {
    const __nv_bfloat16 *qkv_w, *qkv_b, *qnorm, *knorm, *mlp_w, *mlp_b, *out_w, *out_b;
};

struct FluxEngine {
    FluxCfg cfg;
    bool weights_set = false;
    size_t n_mod = 0;
    const __nv_bfloat16 *x_in_w, *x_in_b, *ctx_in_w, *ctx_in_b;
    const __nv_bfloat16 *t_w1, *t_b1, *t_w2, *t_b2, *g_w1, *g_b1, *g_w2, *g_b2, *p_w1, *p_b1, *p_w2, *p_b2;
    const __nv_bfloat16 *mod_w, *mod_b, *final_w, *final_b;
    std::vector<FluxStreamW> dimg, dtxt;
    std::vector<FluxSingleW> single;
    // workspace
    __nv_bfloat16 *img = nullptr, *txt = nullptr, *z = nullptr, *h = nullptr, *q = nullptr, *k = nullptr, *v = nullptr;
    __nv_bfloat16 *attn_img = nullptr, *attn_txt = nullptr, *wide = nullptr, *mod = nullptr, *temb = nullptr;
    __nv_bfloat16 *vec_tmp = nullptr;
};

int flux_create(const FluxCfg& cfg, FluxEngine** out);
int flux_destroy(FluxEngine* e);
int flux_set_weights(FluxEngine* e, const void* const* ptrs, int n);
int flux_forward(FluxEngine* e, const __nv_bfloat16* x, int ldx, const __nv_bfloat16* ctx, const __nv_bfloat16* pooled,
                 const float* t_dev, const float* g_dev, const float* rope_cos, const float* rope_sin, int B, int S_img,
                 __nv_bfloat16* v_out, int ldv, int n_double_run, int n_single_run, cudaStream_t st);

}  // namespace drag
