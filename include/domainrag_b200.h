/*
 * libdomainrag_b200.so - C ABI of the B200-native Domain-RAG retrieve-then-compose hot path.
 *
 * Every entry point returns an int status (0 = ok); the message for the last failure on the calling
 * thread is available from drag_last_error(). Unless a function says "host", every pointer is a
 * DEVICE pointer valid on the handle's device, and work is enqueued on `stream` (a cudaStream_t
 * passed as void*; NULL = the legacy default stream) without synchronising. The library never
 * frees caller memory.  No torch types cross this boundary.
 *
 * The reference (LiYu0524/Domain-RAG) has no FFI of its own: its hot path is seven calls into
 * third-party Python packages. Each group below names the call site it replaces
 * (paths relative to the reference root). INTEGRATION.md shows the ctypes binding.
 */
#ifndef DOMAINRAG_B200_H
#define DOMAINRAG_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DRAG_OK 0
#define DRAG_ERR_INVALID 1
#define DRAG_ERR_CUDA 2
#define DRAG_ERR_UNSUPPORTED 3
#define DRAG_ERR_NO_DEVICE 4

/* ---- library ---------------------------------------------------------------------------- */
const char* drag_last_error(void);
int drag_version(void);
/* Number of visible CUDA devices and the SM count / name of one of them. */
int drag_device_count(int* count);
int drag_device_info(int device, int* sm_count, int* cc_major, int* cc_minor, char* name, int name_len);

/* ---- exact inner-product index ---------------------------------------------------------------
 * Replaces faiss.IndexFlatIP(d) / index.add(X) / index.search(q, k) at
 * retrieval/clip100_resnet_style_all_shots.py:425-434 (clip_first_stage_retrieval).
 * The corpus stays resident in HBM across queries (the reference rebuilds the index per query).
 * Results: scores descending; equal scores -> lower id first; missing slots -> (-FLT_MAX, -1).
 * Limits: 1 <= k <= 1024; fewer than 2^32-1 rows per device. */
typedef struct drag_index drag_index_t;
int drag_index_create(int d, int device, drag_index_t** out);
int drag_index_destroy(drag_index_t* h);
/* Copy N rows of fp32 [N][d] into index-owned device memory; ids are base_id .. base_id+N-1.
 * x_on_device = 0: X is a host pointer (the faiss contract), 1: X is a device pointer. */
int drag_index_add(drag_index_t* h, const float* X, int64_t N, int64_t base_id, int x_on_device, void* stream);
/* Zero-copy variant: the index scans caller-owned device memory (must outlive the index). */
int drag_index_adopt(drag_index_t* h, const float* X_dev, int64_t N, int64_t base_id);
int drag_index_reset(drag_index_t* h);
int drag_index_ntotal(drag_index_t* h, int64_t* ntotal);
/* HOST buffers, synchronous - the drop-in for index.search(): q [nq][d] -> D [nq][k], I [nq][k]. */
int drag_index_search(drag_index_t* h, const float* q_host, int nq, int k, float* D_host, int64_t* I_host);
/* DEVICE buffers, asynchronous on `stream`. */
int drag_index_search_device(drag_index_t* h, const float* q_dev, int nq, int k, float* D_dev, int64_t* I_dev, void* stream);
/* Geometry of the most recent scan launch (for roofline arithmetic in bench.py). */
int drag_index_last_launch(drag_index_t* h, int* grid, int* stages, int* rows_per_stage, int* nq_batch);
/* Optional CUDA-event bracket around the scan kernel launches of the next searches (the merge
 * kernel is outside it); drag_index_last_scan_ms synchronises on the closing event. */
int drag_index_set_timing(drag_index_t* h, int enable);
int drag_index_last_scan_ms(drag_index_t* h, float* ms);
/* Merge `lists` per-shard top-k_in lists per query (after the all-gather of per-shard results,
 * SURVEY 8e) into the global top-k_out: scores/ids [nq][lists][k_in] -> D/I [nq][k_out].
 * ids < 0 mark empty slots. lists*k_in <= 8192. */
int drag_topk_merge_device(const float* scores, const int64_t* ids, int nq, int lists, int k_in, int k_out,
                           float* D_dev, int64_t* I_dev, void* stream);

/* The exchange step of the sharded search as ONE kernel over NVLink peer memory (no NCCL collective): every rank pushes its
 * per-query (score, id) lists into every peer's symmetric buffer with P2P stores, raises a flag, waits for the peers' flags
 * and merges - bit-identical to drag_topk_merge_device over an all-gather. peer_bufs: HOST array of `world` device pointers,
 * entry p = rank p's buffer as mapped into this process (torch.distributed._symmetric_memory or cudaIpc), each of
 * drag_topk_exchange_buffer_bytes(world, nq_cap, k_cap) bytes and zero-initialised once. epoch: 1, 2, 3, ... - the same on
 * every rank, incremented per call (the call is a collective). nq <= nq_cap <= 64 (one resident CTA per query), k <= k_cap,
 * world * k <= 8192, world <= 16. */
int drag_topk_exchange_buffer_bytes(int world, int nq_cap, int k_cap, int64_t* bytes);
int drag_topk_exchange_merge(const float* D_loc, const int64_t* I_loc, int nq, int k, void* const* peer_bufs, int world,
                             int rank, int nq_cap, int k_cap, uint32_t epoch, float* D_dev, int64_t* I_dev, void* stream);
/* The whole sharded search of one rank in one call: drag_index_search_device over this rank's rows into the caller's
 * workspace D_ws / I_ws [nq][k], then drag_topk_exchange_merge. Collective over the ranks that own `peer_bufs`; replaces
 * index.search(q, k) of retrieval/clip100_resnet_style_all_shots.py:431-434 when the corpus is row-sharded. */
int drag_index_search_sharded(drag_index_t* h, const float* q_dev, int nq, int k, void* const* peer_bufs, int world, int rank,
                              int nq_cap, int k_cap, uint32_t epoch, float* D_ws, int64_t* I_ws, float* D_dev,
                              int64_t* I_dev, void* stream);

/* ---- ResNet-50 stem + style statistics --------------------------------------------------------
 * Replaces ResNetEncoder()(x) + calc_mean_std (retrieval/clip100_resnet_style_all_shots.py:51-74,
 * 197-200): img fp32 [B][3][256][256] in [0,1] -> out fp32 [B][128] = cat(mean[64], std[64]),
 * std = sqrt(unbiased var + eps). w_fold [64][3][7][7] / b_fold [64] are conv1 with eval-mode bn1
 * folded in. Implicit GEMM on the tcgen05 tensor cores with split-bf16 operands (x = xh + xl, w = wh + wl; three
 * MMAs per k-step, fp32 accumulation: 2^-16 relative per product), pooling and moments fused; the feature maps never
 * leave the SM (csrc/stem_stats_tc.cu). */
int drag_stem_stats(const float* img_dev, int B, int H, int W, const float* w_fold_dev, const float* b_fold_dev,
                    float eps, float* out_dev, void* stream);
/* Same with the raw uint8 pixels [B][3][256][256] (what cv2.resize returns, channel-major): the / 255 of :193 runs in the
 * kernel's loader (IEEE fp32 division), a quarter of the PCIe / HBM bytes of the float tensor. */
int drag_stem_stats_u8(const uint8_t* img_dev, int B, int H, int W, const float* w_fold_dev, const float* b_fold_dev,
                       float eps, float* out_dev, void* stream);

/* ---- bf16 GEMM core (tcgen05 / TMEM / TMA) ------------------------------------------------------
 * out[M][N] = epilogue(A[M][K] * W[N][K]^T + bias): the nn.Linear calls inside clip.encode_image
 * (retrieval/clip100_resnet_style_all_shots.py:171) and inside the diffusers Flux transformer that
 * pipe(...) runs (batch_generate_flux_kshot.py:467-474, outpainting_updown_sampling_redux.py:1246-1257).
 * A, W, bias, out, resid, gate are bf16 device pointers; fp32 accumulation.
 * epi_mode: 0 bias, 1 gelu_tanh, 2 quick_gelu, 3 silu, 4 out = resid + gate[row/rows_per_batch][n] * (..)
 * (gate NULL = plain residual add), 6 fp32 output (out is float*). K, lda, ldw multiples of 8; N of 32. */
int drag_gemm_bf16(const void* A, int lda, const void* W, int ldw, int M, int N, int K, int epi_mode,
                   const void* bias, void* out, int ldo, const void* resid, int ldr, const void* gate,
                   int gate_ld, int rows_per_batch, void* stream);
/* QKV projection fused with per-head RMSNorm(q), RMSNorm(k) and rotary embedding, scattered into the
 * attention layout: q/k/v_out bf16 [B][heads][s_total][128] at token tok_offset + (row % rows_per_batch).
 * W is [3*heads*128][K] (q rows, then k, then v); rope_cos/sin fp32 [s_total][64]. */
int drag_gemm_qkv_rope(const void* A, int lda, const void* W, int ldw, int M, int K, int heads, const void* bias,
                       void* q_out, void* k_out, void* v_out, const void* q_norm_w, const void* k_norm_w,
                       const float* rope_cos, const float* rope_sin, int s_total, int tok_offset,
                       int rows_per_batch, float rms_eps, void* stream);

/* ---- attention / row kernels of the Flux and ViT paths ------------------------------------------
 * Non-causal attention, head_dim 128 (F.scaled_dot_product_attention in the diffusers Flux blocks) or 64
 * (nn.MultiheadAttention in OpenAI CLIP's ViT): q,k,v bf16 [B][H][S][head_dim]; token s < split is written to out0 row (b*split+s), token s >= split to out1
 * row (b*(S-split)+s-split); head h occupies columns [h*head_dim,(h+1)*head_dim) of a row with leading dim ld0/ld1. */
int drag_attention_bf16(const void* q, const void* k, const void* v, int B, int H, int S, int head_dim, int split,
                        void* out0, int ld0, void* out1, int ld1, void* stream);
/* out = LayerNorm(x) (no affine, eps) then * (1 + mul[b]) + add[b] when adaln = 1 (AdaLN modulate, b =
 * row / rows_per_batch, per-batch vectors with stride mul_ld / add_ld), or * mul + add per channel when
 * adaln = 0 (plain affine LayerNorm; either pointer may be NULL). bf16, d % 8 == 0. */
int drag_layernorm_bf16(const void* x, int ldx, void* out, int ldo, int M, int d, const void* mul, int mul_ld,
                        const void* add, int add_ld, int adaln, int rows_per_batch, float eps, void* stream);
/* Sinusoidal embedding of 1000*t: out bf16 [B][256] = [cos | sin]; t fp32 [B] on the device. */
int drag_timestep_embed(const float* t_dev, void* out, int B, void* stream);
/* Flow-match Euler update on a strided bf16 view: x += dsigma * v (fp32 arithmetic). */
int drag_euler_step(void* x, int ldx, const void* v, int ldv, int rows, int cols, float dsigma, void* stream);
/* Flux 2x2 latent packing (diffusers FluxPipeline._pack_latents / _unpack_latents, inside the pipe(...) calls of
 * batch_generate_flux_kshot.py:467-474 and outpainting_updown_sampling_redux.py:1246-1257): z bf16 [B][C][h][w] <->
 * token s = (y/2)(w/2) + x/2, channel ch_off + c*4 + (y%2)*2 + (x%2) of a row with leading dim ldo / ldx. */
int drag_pack_latents(const void* z, int B, int C, int h, int w, void* out, int64_t ldo, int ch_off, void* stream);
int drag_unpack_latents(const void* x, int64_t ldx, int B, int C, int h, int w, void* z, void* stream);
/* Flux-Fill transformer input (FluxFillPipeline.prepare_mask_latents + the per-step channel concat; outpainting...:1246-1257):
 * x bf16 [B][(h/2)(w/2)][ldx >= 384]: channels 0:64 = `latents` (already packed, leading dim ld_lat; NULL leaves them
 * untouched), 64:128 = packed `masked_latents` bf16 [B][16][h][w], 128:384 = the 8x8 block of `mask` (uint8 [B][8h][8w],
 * non-zero = repaint) under every latent pixel as 64 channels, packed the same way. One launch. */
int drag_pack_fill_inputs(const void* latents, int64_t ld_lat, const void* masked_latents, const uint8_t* mask, int B, int h,
                          int w, void* x, int64_t ldx, void* stream);
/* FluxPriorReduxPipeline output blend (batch_generate_flux_kshot.py:459-465, outpainting...:1237-1243):
 * out_embeds[1][n_txt+n_img][dim] = sum_b s_embed[b] * cat(txt[b], img[b]); out_pooled = sum_b s_pool[b]*pooled[b]. */
int drag_redux_blend(const void* txt, const void* img, const void* pooled, const float* s_embed_dev,
                     const float* s_pool_dev, void* out_embeds, void* out_pooled, int B, int n_txt, int n_img, int dim,
                     int pooled_dim, void* stream);
/* out[r] = x[r] / ||x[r]||_2, fp32 (image_embedding / image_embedding.norm(dim=-1), retrieval...:172). */
int drag_l2_normalize(const float* x, float* out, int rows, int d, void* stream);

/* ---- CLIP ViT image encoder pieces (model.encode_image, retrieval/clip100_resnet_style_all_shots.py:171) ---
 * QKV projection scattered head-major for the attention kernel: q/k/v_out bf16 [B][heads][s_total][head_dim]. */
int drag_gemm_qkv_split(const void* A, int lda, const void* W, int ldw, int M, int K, int heads, int head_dim,
                        const void* bias, void* q_out, void* k_out, void* v_out, int s_total, int tok_offset,
                        int rows_per_batch, void* stream);
/* conv1 (kernel == stride == patch, no bias) as patch extraction + GEMM: img fp32 [B][3][R][R] ->
 * bf16 [B*(R/patch)^2][kpad], columns (c, py, px) zero padded to kpad (multiple of 8). */
int drag_vit_patchify(const float* img, void* out, int B, int R, int patch, int kpad, void* stream);
/* x[b][0] = class_embedding + pos[0]; x[b][1+i] = patch_emb[b][i] + pos[1+i]  (bf16 [B][n_patch+1][w]). */
int drag_vit_assemble(const void* patch_emb, const void* cls, const void* pos, void* x, int B, int n_patch, int w,
                      void* stream);

/* ---- CLIP ViT image tower (one call per batch) -----------------------------------------------------
 * model.encode_image(x) [+ x / x.norm(dim=-1)] of the reference's embedding loops (retrieval/clip100_resnet_style_all_shots.py:
 * 161-177, 270-296, 326-349; model from clip.load at :209): the whole tower - patch embedding, class token + positions,
 * ln_pre, `layers` pre-LN blocks (QKV GEMM -> head-dim-64 attention -> out projection + residual -> QuickGELU MLP +
 * residual), ln_post of the class token, projection, optional L2 normalise - orchestrated inside the library.
 * Weights: caller-owned device pointers in the order of domain_rag_b200/clip.py::ENGINE_ORDER (8 globals, 18 per block:
 * the 12 bf16 tensors of the block, then the LayerNorm-folded QKV / MLP-up weights (bf16) each with its fp32 s and c vectors). img_kind 0: fp32 [B][3][image][image] already normalised (what `preprocess` returns); img_kind 1: raw uint8
 * pixels [B][3][image][image] - ToTensor + Normalize((u8/255 - mean) / std, IEEE fp32) run inside the patch kernel.
 * out fp32 [B][out_dim]. B may exceed max_batch (processed in chunks of max_batch). */
typedef struct drag_vit drag_vit_t;
typedef struct {
    int width, layers, heads, patch, image, out_dim, max_batch;
    float mean[3], std[3];
} drag_vit_config;
int drag_vit_create(const drag_vit_config* cfg, drag_vit_t** out);
int drag_vit_destroy(drag_vit_t* h);
int drag_vit_set_weights(drag_vit_t* h, const void* const* ptrs, int n);
/* key 1: fold ln_1 / ln_2 into the GEMMs around them (default 0 - measured no faster on C2; kept for A/B): the residual GEMMs' epilogues emit per-row moments, the QKV
 * and MLP-up GEMMs run on the raw residual rows with gamma-scaled weights and apply rstd * (acc - mean * s) + c. 0 = separate
 * LayerNorm kernels (A/B comparisons). */
int drag_vit_set_option(drag_vit_t* h, int key, int value);
int drag_vit_encode(drag_vit_t* h, const void* img, int img_kind, int B, float* out, int l2_normalize, void* stream);

/* ---- Flux MMDiT engine --------------------------------------------------------------------------
 * One FluxTransformer2DModel.forward per call (the denoising step inside pipe(...) /
 * pipe_fill(...): batch_generate_flux_kshot.py:467-474, outpainting_updown_sampling_redux.py:1246-1257).
 * Weights are caller-owned bf16 device pointers in the canonical order documented in
 * domain_rag_b200/flux.py (param_order); the engine owns its activation workspace. */
typedef struct drag_flux drag_flux_t;
typedef struct {
    int in_channels, d, heads, n_double, n_single, txt_dim, pooled_dim, out_channels, guidance;
    int max_batch, max_img_tokens, txt_tokens;
} drag_flux_config;
int drag_flux_create(const drag_flux_config* cfg, drag_flux_t** out);
int drag_flux_destroy(drag_flux_t* h);
int drag_flux_set_weights(drag_flux_t* h, const void* const* ptrs, int n);
/* x bf16 [B*S_img][ldx] (first in_channels columns), ctx bf16 [B*txt_tokens][txt_dim], pooled bf16
 * [B][pooled_dim], t_dev / g_dev fp32 [B] (sigma and guidance scale), rope tables fp32 [txt_tokens+S_img][64];
 * v_out bf16 [B*S_img][ldv]. n_double_run / n_single_run < 0 run every block. */
int drag_flux_forward(drag_flux_t* h, const void* x, int ldx, const void* ctx, const void* pooled, const float* t_dev,
                      const float* g_dev, const float* rope_cos, const float* rope_sin, int B, int S_img, void* v_out,
                      int ldv, int n_double_run, int n_single_run, void* stream);

/* ---- Flux VAE path (AutoencoderKL decode after / encode before the sampling loop, and VaeImageProcessor) --------
 * What `pipe(...).images` and `pipe_fill(image=..., mask_image=...)` run around the transformer
 * (batch_generate_flux_kshot.py:467-474, outpainting_updown_sampling_redux.py:1246-1257). Activations are bf16 NHWC.
 * Convolution as implicit GEMM on the tcgen05 core: in [B][H][W][C_in] (C_in % 64 == 0), w [C_out][ksize*ksize][C_in]
 * (tap-major), out [B][Ho][Wo][C_out]; ksize 1|3, stride 1|2, pad = left/top padding (the rest is zero fill).
 * epi_mode: 0 bias, 3 silu, 4 out = resid + (acc + bias), 6 fp32 output. */
int drag_conv2d_nhwc(const void* in, int B, int H, int W, int C_in, const void* w, int C_out, int ksize, int stride,
                     int pad, int Ho, int Wo, int epi_mode, const void* bias, void* out, const void* resid, void* stream);
/* GroupNorm over [HW x C/groups] per image and group, affine, optional SiLU. workspace: fp32 scratch of at least
 * B*1024*2*C + B*groups*2 floats. */
int drag_groupnorm_nhwc(const void* x, void* y, int B, int HW, int C, int groups, const void* gamma, const void* beta,
                        float eps, int silu, float* workspace, int64_t workspace_floats, void* stream);
int drag_upsample2x_nhwc(const void* x, void* y, int B, int H, int W, int C, void* stream);
/* p[r][:] = softmax(s[r][:]) (fp32 scores -> bf16 probabilities), cols % 4 == 0. */
int drag_softmax_rows(const float* s, int64_t ld_s, void* p, int64_t ld_p, int rows, int cols, void* stream);
/* out[b][y][x][c] = in[b][c][y][x] * scale + shift (c < C), 0 for padded channels; in fp32 or bf16. */
int drag_nchw_to_nhwc_pad(const void* in, int in_is_f32, void* out, int B, int C, int H, int W, int C_pad, float scale,
                          float shift, void* stream);
/* out fp32 [B][C][H][W] = in[b][y][x][c] * scale + shift from NHWC rows of ld elements (fp32 or bf16). */
int drag_nhwc_to_nchw_f32(const void* in, int in_is_f32, int ld, float* out, int B, int C, int H, int W, float scale,
                          float shift, void* stream);
/* uint8 RGB [pixels][3] = round(clamp(x / 2 + 0.5, 0, 1) * 255) from fp32 NHWC rows of ld floats. */
int drag_image_postprocess_u8(const float* in, int ld, uint8_t* out, int64_t pixels, void* stream);
/* bf16 NHWC [pixels][C_pad] = u8 / 255 * 2 - 1 (channels >= 3 zero). mask (optional, uint8 per pixel): non-zero pixels
 * are written as 0 = init_image * (1 - mask) of FluxFillPipeline. */
int drag_image_preprocess_u8(const uint8_t* in, const uint8_t* mask, void* out, int64_t pixels, int C_pad, void* stream);
/* out = a * x + b * y (bf16, fp32 arithmetic): scheduler.scale_noise of the img2img / fill pipelines. */
int drag_axpby_bf16(const void* x, const void* y, float a, float b, void* out, int64_t n, void* stream);

/* Per-launch CUDA-event timing of the heavy kernels on their launching stream (bench.py roofline):
 * class 0 = tcgen05 GEMM (work = 2*M*N*K flops), class 1 = attention (work = 4*B*H*S*S*128 flops).
 * drag_prof_collect synchronises, returns summed milliseconds / work / launch counts per class, and resets. */
int drag_prof_enable(int on);
int drag_prof_collect(double* ms, double* work, int* count, int n_classes);

/* Number of CUDA kernels this library has launched in the calling process since the last reset (host-side counter bumped at
 * every launch site; single-threaded use as everywhere in this ABI). bench.py reports it as `gpu_launches`. */
int drag_launch_count(int64_t* count, int reset);

/* Debug knobs for bring-up and A/B comparisons (key 1/2: unused; key 3: 1 = force the single-CTA GEMM kernel
 * instead of the CTA-pair cta_group::2 kernel; key 4: > 0 = force the GEMM tile-raster group size, 1 << 20 = plain
 * row-fastest order; key 5: 1 = head-dim-64 attention always on the two-tile ping-pong kernel; key 6: > 0 = force the
 * column-group raster with that many column tiles per group; key 7: head-dim-128 attention:
 * 1 = split-row kernel, two softmax warpgroups per query tile (measured slower), 0 = one thread per row (default); key 8: 1 = stem statistics on the FP32 CUDA-core kernel instead of
 * the tensor-core kernel; key 9: 1 = GEMM epilogues store 16 bytes per lane instead of 32; key 10: 1 = head-dim-64
 * attention never takes the whole-row kernel, i.e. the tiled online-softmax kernels also for <= 260 keys; key 11: 0 = the whole-row kernel
 * does not prefetch the tiles of later CTAs into L2; key 12: 0 = 129..260 keys take the one-tile-per-CTA whole-row kernel
 * instead of the persistent one; key 13: 0 = its two query tiles start together instead of half an item apart; key 14: 1 = it
 * takes 2 of 8 exponentials from the FMA-pipe polynomial like the head-dim-128 kernel; key 15: QuickGELU reciprocal 3 = one MUFU.RCP
 * per two elements (default), 4 = per four, 1 = per element, 0 = FMA-pipe Newton iteration, 2 = one element of four on the FMA pipe; key 16: 1 = the
 * persistent head-dim-64 attention kernel shares every score row between two softmax threads - measured slower; key 17: 0 = up to 128 keys take
 * the one-tile-per-CTA whole-row kernel instead of the persistent kernel with two heads per item; key 18: 0 = the residual
 * GEMMs with K <= 2048 use the generic epilogue instead of the instantiation that requests the residual rows up front).
 * The environment variable DRAG_DEBUG_SET="key=value,key=value" applies the same knobs when the Python binding loads the library. */
int drag_debug_set(int key, int value);

#ifdef __cplusplus
}
#endif
#endif /* DOMAINRAG_B200_H */
