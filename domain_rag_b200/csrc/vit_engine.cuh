// Host-side state of one CLIP ViT image-tower engine (weights are caller-owned bf16 device pointers; the engine owns only
// its activation workspace, sized for cfg.max_batch images).
#pragma once
#include <cuda_bf16.h>

#include <vector>

#include "common.cuh"

namespace drag {

struct VitCfg {
    int width, layers, heads, patch, image, out_dim, max_batch;
    float mean[3], std[3];          // Normalize constants of the uint8 input path (CLIP: 0.4814.. / 0.2686..)
};

struct VitBlockW {
    const __nv_bfloat16 *ln1_w, *ln1_b, *qkv_w, *qkv_b, *out_w, *out_b, *ln2_w, *ln2_b, *fc_w, *fc_b, *proj_w, *proj_b;
    // LayerNorm folded into the GEMM that consumes it (see GemmEpi): gamma-scaled weights and the two fp32 vectors
    const __nv_bfloat16 *qkv_wf, *fc_wf;
    const float *qkv_s, *qkv_c, *fc_s, *fc_c;
};

struct VitEngine {
    VitCfg cfg;
    int grid = 0, tokens = 0, kpad = 0;
    bool weights_set = false;
    const __nv_bfloat16 *conv_w = nullptr, *cls = nullptr, *pos = nullptr, *ln_pre_w = nullptr, *ln_pre_b = nullptr,
                        *ln_post_w = nullptr, *ln_post_b = nullptr, *proj_t = nullptr;
    std::vector<VitBlockW> blocks;
    __nv_bfloat16 *patches = nullptr, *pe = nullptr, *h = nullptr, *y = nullptr, *q = nullptr, *k = nullptr, *v = nullptr,
                  *a = nullptr, *u = nullptr, *cls_ln = nullptr;
    float* emb = nullptr;
    float *stats_a = nullptr, *stats_b = nullptr;   // [max_batch * tokens][width / 128][2] row moments (LN folding)
    int fold_ln = 0;                                  // 1: ln_1 / ln_2 folded into the GEMM epilogues (drag_vit_set_option key 1).
                                                      // Measured on C2 (ViT-L/14): no faster than the separate kernels - off by default
};

int vit_create(const VitCfg& cfg, VitEngine** out);
int vit_destroy(VitEngine* e);
int vit_set_weights(VitEngine* e, const void* const* ptrs, int n);
// img_kind 0: fp32 [B][3][R][R] already normalised (what `preprocess` returns); 1: uint8 [B][3][R][R] raw pixels.
int vit_set_option(VitEngine* e, int key, int value);
int vit_encode(VitEngine* e, const void* img, int img_kind, int B, float* out, int normalize, cudaStream_t st);

}  // namespace drag
