// Declarations of the row/elementwise kernels, the attention core and TMA descriptor helpers.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>

#include "common.cuh"

namespace drag {

int layernorm_bf16(const __nv_bfloat16* x, int ldx, __nv_bfloat16* out, int ldo, int M, int d,
                   const __nv_bfloat16* mul, int mul_ld, const __nv_bfloat16* add, int add_ld, int mul_add_one,
                   int rows_per_batch, float eps, cudaStream_t st);
int timestep_embed(const float* t_dev, __nv_bfloat16* out, int B, cudaStream_t st);
int sum_silu(const __nv_bfloat16* a, const __nv_bfloat16* b, const __nv_bfloat16* c, __nv_bfloat16* out, int n,
             int apply_silu, cudaStream_t st);
int euler_step(__nv_bfloat16* x, int ldx, const __nv_bfloat16* v, int ldv, int rows, int cols, float dsigma,
               cudaStream_t st);
int pack_latents(const __nv_bfloat16* z, int B, int C, int h, int w, __nv_bfloat16* out, int64_t ldo, int ch_off,
                 cudaStream_t st);
int unpack_latents(const __nv_bfloat16* x, int64_t ldx, int B, int C, int h, int w, __nv_bfloat16* z, cudaStream_t st);
int pack_fill_inputs(const __nv_bfloat16* latents, int64_t ld_lat, const __nv_bfloat16* masked, const uint8_t* mask, int B,
                     int h, int w, __nv_bfloat16* x, int64_t ldx, cudaStream_t st);
int redux_blend(const __nv_bfloat16* txt, const __nv_bfloat16* img, const __nv_bfloat16* pooled, const float* s_embed,
                const float* s_pool, __nv_bfloat16* out_embeds, __nv_bfloat16* out_pooled, int B, int n_txt, int n_img,
                int dim, int pooled_dim, cudaStream_t st);
int vit_patchify(const float* img, __nv_bfloat16* out, int B, int R, int p, int kpad, cudaStream_t st);
int vit_assemble(const __nv_bfloat16* patch_emb, const __nv_bfloat16* cls, const __nv_bfloat16* pos, __nv_bfloat16* x,
                 int B, int n_patch, int w, cudaStream_t st);
int l2_normalize(const float* x, float* out, int rows, int d, cudaStream_t st);

// q,k,v bf16 [B][H][S][head_dim], head_dim 64 or 128. Token s < split goes to out0 row (b*split + s) with leading dim ld0,
// token s >= split to out1 row (b*(S-split) + s-split) with leading dim ld1; head h at column h*head_dim.
int attention_bf16(const __nv_bfloat16* q, const __nv_bfloat16* k, const __nv_bfloat16* v, int B, int H, int S,
                   int head_dim, int split, __nv_bfloat16* out0, int ld0, __nv_bfloat16* out1, int ld1, cudaStream_t st);

int make_tmap_bf16_2d(CUtensorMap* map, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows);
int make_tmap_bf16_3d(CUtensorMap* map, const void* ptr, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t stride1_elems,
                      uint64_t stride2_elems, uint32_t box0, uint32_t box1);

}  // namespace drag
