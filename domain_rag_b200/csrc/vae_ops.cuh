// Declarations of the VAE-path kernels (vae_ops.cu).
#pragma once
#include <cuda_bf16.h>
#include <stddef.h>
#include <stdint.h>

#include "common.cuh"

namespace drag {

int groupnorm_nhwc(const __nv_bfloat16* x, __nv_bfloat16* y, int B, int HW, int C, int groups, const __nv_bfloat16* gamma,
                   const __nv_bfloat16* beta, float eps, int silu, float* workspace, size_t workspace_floats,
                   cudaStream_t st);
int upsample2x_nhwc(const __nv_bfloat16* x, __nv_bfloat16* y, int B, int H, int W, int C, cudaStream_t st);
int softmax_rows(const float* s, size_t ld_s, __nv_bfloat16* p, size_t ld_p, int rows, int cols, cudaStream_t st);
int nchw_to_nhwc_pad(const void* in, int in_is_f32, __nv_bfloat16* out, int B, int C, int H, int W, int C_pad, float scale,
                     float shift, cudaStream_t st);
int nhwc_to_nchw_f32(const void* in, int in_is_f32, int ld, float* out, int B, int C, int H, int W, float scale, float shift,
                     cudaStream_t st);
int image_postprocess_u8(const float* in, int ld, uint8_t* out, size_t pixels, cudaStream_t st);
int image_preprocess_u8(const uint8_t* in, const uint8_t* mask, __nv_bfloat16* out, size_t pixels, int C_pad, cudaStream_t st);
int axpby_bf16(const __nv_bfloat16* x, const __nv_bfloat16* y, float a, float b, __nv_bfloat16* out, size_t n, cudaStream_t st);

}  // namespace drag
