// Host/device interface of the tcgen05 GEMM core (gemm_tcgen05.cu).
#pragma once
#include <cuda_bf16.h>

#include "common.cuh"

namespace drag {

enum GemmEpiMode : int {
    EPI_BIAS = 0,        // out = acc + bias                        (bf16)
    EPI_GELU_TANH = 1,   // out = gelu_tanh(acc + bias)             (Flux MLP)
    EPI_QUICK_GELU = 2,  // out = x * sigmoid(1.702 x)              (OpenAI CLIP MLP)
    EPI_SILU = 3,        // out = x * sigmoid(x)                    (time/guidance/pooled MLPs, Redux)
    EPI_GATE_RESID = 4,  // out = resid + gate[b][n] * (acc + bias) (gate == null: plain residual)
    EPI_QKV_ROPE = 5,    // per-head RMSNorm(q,k) + RoPE, scatter to [B][H][S][128]
    EPI_BIAS_F32 = 6,    // out_f32 = acc + bias
    EPI_QKV_SPLIT = 7,   // acc + bias scattered to q/k/v [B][H][S][head_dim] (no norm / rope; CLIP ViT)
};

struct GemmEpi {
    int mode = EPI_BIAS;
    const __nv_bfloat16* bias = nullptr;  // [N] or null
    __nv_bfloat16* out = nullptr;         // row-major, leading dimension ldo (elements)
    float* out_f32 = nullptr;
    int ldo = 0;
    // EPI_GATE_RESID
    const __nv_bfloat16* resid = nullptr;
    int ldr = 0;
    const __nv_bfloat16* gate = nullptr;  // [B][gate_ld]
    int gate_ld = 0;
    int rows_per_batch = 1 << 30;         // batch index of a row = row / rows_per_batch
    // EPI_QKV_ROPE: N == 3 * heads * 128, columns [0,H*128) = q, then k, then v
    __nv_bfloat16* q_out = nullptr;       // [B][H][S_total][128]
    __nv_bfloat16* k_out = nullptr;
    __nv_bfloat16* v_out = nullptr;
    const __nv_bfloat16* q_norm_w = nullptr;  // [128]
    const __nv_bfloat16* k_norm_w = nullptr;  // [128]
    const float* rope_cos = nullptr;      // [S_total][64]
    const float* rope_sin = nullptr;      // [S_total][64]
    int heads = 0, s_total = 0, tok_offset = 0;
    int head_dim = 128;                   // EPI_QKV_SPLIT: 64 or 128
    float rms_eps = 1e-6f;
    // LayerNorm folded into the GEMMs on either side of it (CLIP ViT blocks; any mode that goes through the generic epilogue).
    //   producer: stats_out != null -> every epilogue thread also writes the (sum, sum of squares) of the bf16-rounded values
    //             it stores for its row and its BN/2 columns: stats_out[(row * stats_parts + part) * 2 + {0,1}], part =
    //             column / (BN/2), stats_parts = N / (BN/2) (checked by gemm_bf16). Deterministic: no atomics.
    //   consumer: ln_stats != null -> A holds the RAW rows x (not normalised); W holds gamma-scaled weights W' = W * gamma.
    //             With mean / rstd from the ln_parts partials over ln_k columns:
    //                 LN(x) W^T + b  =  rstd * (x W'^T - mean * ln_s) + ln_c,   ln_s[n] = sum_k W'[n][k],
    //                                                                          ln_c[n] = sum_k beta[k] W[n][k] + b[n]
    //             applied to the accumulator before the activation / scatter (bias must be null: it lives in ln_c).
    float* stats_out = nullptr;
    int stats_parts = 0;
    const float* ln_stats = nullptr;
    int ln_parts = 0, ln_k = 0;
    const float* ln_s = nullptr;          // [N] fp32
    const float* ln_c = nullptr;          // [N] fp32
    float ln_eps = 1e-5f;
    int rcp_mufu = 0;                     // set by the launcher: QuickGELU takes its reciprocal from MUFU.RCP (drag_debug_set key 15)
    int wide_st = 0;                      // set by the launcher: 32-byte (STG.256) row stores are legal for this output
    int wide_ld = 0;                      //                      32-byte (LDG.256) loads of the residual rows
};

// C[M,N] = A[M,K] (bf16 row-major, lda) x W[N,K]^T (bf16 row-major, ldw), fp32 accumulate in TMEM.
// Requirements: K % 8 == 0, lda % 8 == 0, ldw % 8 == 0, A and W 16-byte aligned.
int gemm_bf16(const __nv_bfloat16* A, int lda, const __nv_bfloat16* W, int ldw, int M, int N, int K,
              const GemmEpi& epi, cudaStream_t st);

// Implicit-GEMM convolution on the same kernels (A tiles fetched as 4-D TMA boxes of the NHWC activation).
int conv2d_nhwc_bf16(const __nv_bfloat16* in, int B, int H, int W, int C_in, const __nv_bfloat16* w, int C_out,
                     int ksize, int stride, int pad, int Ho, int Wo, const GemmEpi& epi, cudaStream_t st);

}  // namespace drag
