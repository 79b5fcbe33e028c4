// extern "C" surface of libdomainrag_b200.so (declared in include/domainrag_b200.h).
#include <string.h>

#include <string>
#include <vector>

#include "../../include/domainrag_b200.h"
#include "common.cuh"
#include "flux_engine.cuh"
#include "vit_engine.cuh"
#include "flux_ops.cuh"
#include "gemm.cuh"
#include "vae_ops.cuh"
#include "index.cuh"

namespace drag {

static thread_local std::string g_last_error;
void set_error(const std::string& msg) { g_last_error = msg; }
int fail(int code, const std::string& msg) {
    g_last_error = msg;
    return code;
}
int device_sm_count() {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
    return n;
}

long long g_launch_count = 0;
extern uint32_t g_attn_v_lbo, g_attn_v_sbo;
extern int g_gemm_force_1cta;
extern int g_gemm_group_m;
extern int g_attn_force_pp;
extern int g_stem_force_ffma;
extern int g_attn_split;
extern int g_attn_no_row;
extern int g_attn_row_prefetch;
extern int g_attn_row_persistent;
extern int g_attn_row_stagger;
extern int g_attn_row_poly;
extern int g_attn_row_split;
extern int g_attn_row_pair;
extern int g_gemm_no_wide_st;
extern int g_gemm_quick_gelu_mufu;
extern int g_gemm_resid_prefetch;
int stem_stats_any_device(const void* img, int img_kind, int B, int H, int W, const float* w_fold, const float* b_fold,
                          float eps, float* out, cudaStream_t st);
extern int g_gemm_group_n;
int prof_enable(int on);
int prof_collect(double* ms, double* work, int* count, int n_classes);

}  // namespace drag

using namespace drag;

extern "C" {

const char* drag_last_error(void) { return g_last_error.c_str(); }
int drag_version(void) { return 100; }

int drag_device_count(int* count) {
    DRAG_REQUIRE(count, "drag_device_count: null pointer");
    cudaError_t e = cudaGetDeviceCount(count);
    if (e != cudaSuccess) {
        *count = 0;
        return fail(DRAG_ERR_NO_DEVICE, std::string("cudaGetDeviceCount: ") + cudaGetErrorString(e));
    }
    return DRAG_OK;
}

int drag_device_info(int device, int* sm_count, int* cc_major, int* cc_minor, char* name, int name_len) {
    cudaDeviceProp p;
    DRAG_CUDA(cudaGetDeviceProperties(&p, device));
    if (sm_count) *sm_count = p.multiProcessorCount;
    if (cc_major) *cc_major = p.major;
    if (cc_minor) *cc_minor = p.minor;
    if (name && name_len > 0) {
        strncpy(name, p.name, name_len - 1);
        name[name_len - 1] = 0;
    }
    return DRAG_OK;
}

// ------------------------------------------------------------------------------------ index
int drag_index_create(int d, int device, drag_index_t** out) {
    DRAG_REQUIRE(out, "drag_index_create: null out pointer");
    DRAG_REQUIRE(d >= 1 && d <= 16384, "drag_index_create: need 1 <= d <= 16384");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail(DRAG_ERR_NO_DEVICE, "drag_index_create: no CUDA device (this library has no CPU path)");
    DRAG_REQUIRE(device >= 0 && device < ndev, "drag_index_create: bad device ordinal");
    Index* ix = new Index();
    ix->d = d;
    ix->device = device;
    DRAG_CUDA(cudaDeviceGetAttribute(&ix->sm_count, cudaDevAttrMultiProcessorCount, device));
    *out = reinterpret_cast<drag_index_t*>(ix);
    return DRAG_OK;
}

int drag_index_reset(drag_index_t* h) {
    DRAG_REQUIRE(h, "drag_index_reset: null index");
    Index* ix = reinterpret_cast<Index*>(h);
    DRAG_CUDA(cudaSetDevice(ix->device));
    for (Segment& s : ix->segs)
        if (s.owned && s.X) DRAG_CUDA(cudaFree(s.X));
    ix->segs.clear();
    ix->ntotal = 0;
    ix->seg_tab_n = -1;
    return DRAG_OK;
}

int drag_index_destroy(drag_index_t* h) {
    if (!h) return DRAG_OK;
    Index* ix = reinterpret_cast<Index*>(h);
    int rc = drag_index_reset(h);
    if (ix->partial) cudaFree(ix->partial);
    if (ix->qdev) cudaFree(ix->qdev);
    if (ix->Ddev) cudaFree(ix->Ddev);
    if (ix->Idev) cudaFree(ix->Idev);
    if (ix->seg_start) cudaFree(ix->seg_start);
    if (ix->seg_base) cudaFree(ix->seg_base);
    if (ix->ev0) cudaEventDestroy(ix->ev0);
    if (ix->ev1) cudaEventDestroy(ix->ev1);
    delete ix;
    return rc;
}

int drag_index_add(drag_index_t* h, const float* X, int64_t N, int64_t base_id, int x_on_device, void* stream) {
    DRAG_REQUIRE(h, "drag_index_add: null index");
    DRAG_REQUIRE(N >= 0, "drag_index_add: negative row count");
    if (N == 0) return DRAG_OK;
    DRAG_REQUIRE(X, "drag_index_add: null data pointer");
    Index* ix = reinterpret_cast<Index*>(h);
    DRAG_REQUIRE(ix->ntotal + N < 0xFFFFFFFFll, "drag_index_add: more than 2^32-1 rows on one device");
    DRAG_CUDA(cudaSetDevice(ix->device));
    Segment s;
    s.N = N;
    s.base_id = base_id;
    s.owned = true;
    const size_t bytes = static_cast<size_t>(N) * ix->d * sizeof(float);
    DRAG_CUDA(cudaMalloc(&s.X, bytes));
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    cudaError_t e = cudaMemcpyAsync(s.X, X, bytes, x_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess && !x_on_device) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) {
        cudaFree(s.X);
        return fail(DRAG_ERR_CUDA, std::string("drag_index_add copy: ") + cudaGetErrorString(e));
    }
    ix->segs.push_back(s);
    ix->ntotal += N;
    ix->seg_tab_n = -1;
    return DRAG_OK;
}

int drag_index_adopt(drag_index_t* h, const float* X_dev, int64_t N, int64_t base_id) {
    DRAG_REQUIRE(h, "drag_index_adopt: null index");
    DRAG_REQUIRE(N >= 0, "drag_index_adopt: negative row count");
    if (N == 0) return DRAG_OK;
    DRAG_REQUIRE(X_dev, "drag_index_adopt: null data pointer");
    Index* ix = reinterpret_cast<Index*>(h);
    DRAG_REQUIRE(ix->ntotal + N < 0xFFFFFFFFll, "drag_index_adopt: more than 2^32-1 rows on one device");
    Segment s;
    s.X = const_cast<float*>(X_dev);
    s.N = N;
    s.base_id = base_id;
    s.owned = false;
    ix->segs.push_back(s);
    ix->ntotal += N;
    ix->seg_tab_n = -1;
    return DRAG_OK;
}

int drag_index_ntotal(drag_index_t* h, int64_t* ntotal) {
    DRAG_REQUIRE(h && ntotal, "drag_index_ntotal: null pointer");
    *ntotal = reinterpret_cast<Index*>(h)->ntotal;
    return DRAG_OK;
}

int drag_index_search_device(drag_index_t* h, const float* q_dev, int nq, int k, float* D_dev, int64_t* I_dev,
                             void* stream) {
    return index_search_device(reinterpret_cast<Index*>(h), q_dev, nq, k, D_dev, I_dev,
                               reinterpret_cast<cudaStream_t>(stream));
}

int drag_index_search(drag_index_t* h, const float* q_host, int nq, int k, float* D_host, int64_t* I_host) {
    DRAG_REQUIRE(h, "drag_index_search: null index");
    DRAG_REQUIRE(nq >= 0 && k >= 1, "drag_index_search: bad nq/k");
    if (nq == 0) return DRAG_OK;
    DRAG_REQUIRE(q_host && D_host && I_host, "drag_index_search: null pointer");
    Index* ix = reinterpret_cast<Index*>(h);
    DRAG_CUDA(cudaSetDevice(ix->device));
    int rc = index_ensure_io(ix, nq, k);
    if (rc) return rc;
    cudaStream_t st = 0;
    DRAG_CUDA(cudaMemcpyAsync(ix->qdev, q_host, static_cast<size_t>(nq) * ix->d * sizeof(float),
                              cudaMemcpyHostToDevice, st));
    rc = index_search_device(ix, ix->qdev, nq, k, ix->Ddev, ix->Idev, st);
    if (rc) return rc;
    DRAG_CUDA(cudaMemcpyAsync(D_host, ix->Ddev, static_cast<size_t>(nq) * k * sizeof(float),
                              cudaMemcpyDeviceToHost, st));
    DRAG_CUDA(cudaMemcpyAsync(I_host, ix->Idev, static_cast<size_t>(nq) * k * sizeof(int64_t),
                              cudaMemcpyDeviceToHost, st));
    DRAG_CUDA(cudaStreamSynchronize(st));
    return DRAG_OK;
}

int drag_index_last_launch(drag_index_t* h, int* grid, int* stages, int* rows_per_stage, int* nq_batch) {
    DRAG_REQUIRE(h, "drag_index_last_launch: null index");
    Index* ix = reinterpret_cast<Index*>(h);
    if (grid) *grid = ix->last_grid;
    if (stages) *stages = ix->last_stages;
    if (rows_per_stage) *rows_per_stage = ix->last_rps;
    if (nq_batch) *nq_batch = ix->last_nqb;
    return DRAG_OK;
}

int drag_index_set_timing(drag_index_t* h, int enable) {
    DRAG_REQUIRE(h, "drag_index_set_timing: null index");
    Index* ix = reinterpret_cast<Index*>(h);
    DRAG_CUDA(cudaSetDevice(ix->device));
    if (enable && !ix->ev0) {
        DRAG_CUDA(cudaEventCreate(&ix->ev0));
        DRAG_CUDA(cudaEventCreate(&ix->ev1));
    }
    ix->timing = enable != 0;
    return DRAG_OK;
}

int drag_index_last_scan_ms(drag_index_t* h, float* ms) {
    DRAG_REQUIRE(h && ms, "drag_index_last_scan_ms: null pointer");
    Index* ix = reinterpret_cast<Index*>(h);
    DRAG_REQUIRE(ix->timing && ix->ev0, "drag_index_last_scan_ms: timing not enabled");
    DRAG_CUDA(cudaEventSynchronize(ix->ev1));
    DRAG_CUDA(cudaEventElapsedTime(ms, ix->ev0, ix->ev1));
    return DRAG_OK;
}

int drag_topk_merge_device(const float* scores, const int64_t* ids, int nq, int lists, int k_in, int k_out,
                           float* D_dev, int64_t* I_dev, void* stream) {
    return merge_pairs_device(scores, ids, nq, lists, k_in, k_out, D_dev, I_dev,
                              reinterpret_cast<cudaStream_t>(stream));
}

// ------------------------------------------------------------------------------------ stem
int drag_stem_stats(const float* img_dev, int B, int H, int W, const float* w_fold_dev, const float* b_fold_dev,
                    float eps, float* out_dev, void* stream) {
    return stem_stats_device(img_dev, B, H, W, w_fold_dev, b_fold_dev, eps, out_dev,
                             reinterpret_cast<cudaStream_t>(stream));
}
int drag_stem_stats_u8(const uint8_t* img_dev, int B, int H, int W, const float* w_fold_dev, const float* b_fold_dev,
                       float eps, float* out_dev, void* stream) {
    return stem_stats_any_device(img_dev, 1, B, H, W, w_fold_dev, b_fold_dev, eps, out_dev, reinterpret_cast<cudaStream_t>(stream));
}

// ------------------------------------------------------------------------------------ gemm
int drag_gemm_bf16(const void* A, int lda, const void* W, int ldw, int M, int N, int K, int epi_mode,
                   const void* bias, void* out, int ldo, const void* resid, int ldr, const void* gate,
                   int gate_ld, int rows_per_batch, void* stream) {
    DRAG_REQUIRE(epi_mode >= 0 && epi_mode <= 6 && epi_mode != EPI_QKV_ROPE, "drag_gemm_bf16: bad epi_mode");
    GemmEpi e;
    e.mode = epi_mode;
    e.bias = static_cast<const __nv_bfloat16*>(bias);
    if (epi_mode == EPI_BIAS_F32) e.out_f32 = static_cast<float*>(out);
    else e.out = static_cast<__nv_bfloat16*>(out);
    e.ldo = ldo;
    e.resid = static_cast<const __nv_bfloat16*>(resid);
    e.ldr = ldr;
    e.gate = static_cast<const __nv_bfloat16*>(gate);
    e.gate_ld = gate_ld;
    e.rows_per_batch = rows_per_batch > 0 ? rows_per_batch : (1 << 30);
    if (epi_mode == EPI_GATE_RESID) DRAG_REQUIRE(resid, "drag_gemm_bf16: gated residual needs resid");
    return gemm_bf16(static_cast<const __nv_bfloat16*>(A), lda, static_cast<const __nv_bfloat16*>(W), ldw, M, N, K,
                     e, reinterpret_cast<cudaStream_t>(stream));
}

int drag_gemm_qkv_rope(const void* A, int lda, const void* W, int ldw, int M, int K, int heads, const void* bias,
                       void* q_out, void* k_out, void* v_out, const void* q_norm_w, const void* k_norm_w,
                       const float* rope_cos, const float* rope_sin, int s_total, int tok_offset,
                       int rows_per_batch, float rms_eps, void* stream) {
    GemmEpi e;
    e.mode = EPI_QKV_ROPE;
    e.bias = static_cast<const __nv_bfloat16*>(bias);
    e.q_out = static_cast<__nv_bfloat16*>(q_out);
    e.k_out = static_cast<__nv_bfloat16*>(k_out);
    e.v_out = static_cast<__nv_bfloat16*>(v_out);
    e.q_norm_w = static_cast<const __nv_bfloat16*>(q_norm_w);
    e.k_norm_w = static_cast<const __nv_bfloat16*>(k_norm_w);
    e.rope_cos = rope_cos;
    e.rope_sin = rope_sin;
    e.heads = heads;
    e.s_total = s_total;
    e.tok_offset = tok_offset;
    e.rows_per_batch = rows_per_batch > 0 ? rows_per_batch : (1 << 30);
    e.rms_eps = rms_eps;
    return gemm_bf16(static_cast<const __nv_bfloat16*>(A), lda, static_cast<const __nv_bfloat16*>(W), ldw, M,
                     3 * heads * 128, K, e, reinterpret_cast<cudaStream_t>(stream));
}

// ------------------------------------------------------------------------------------ row kernels
#define BF(p) static_cast<const __nv_bfloat16*>(p)
#define BFM(p) static_cast<__nv_bfloat16*>(p)
#define ST(p) reinterpret_cast<cudaStream_t>(p)

int drag_attention_bf16(const void* q, const void* k, const void* v, int B, int H, int S, int head_dim, int split,
                        void* out0, int ld0, void* out1, int ld1, void* stream) {
    return attention_bf16(BF(q), BF(k), BF(v), B, H, S, head_dim, split, BFM(out0), ld0, BFM(out1), ld1, ST(stream));
}
int drag_layernorm_bf16(const void* x, int ldx, void* out, int ldo, int M, int d, const void* mul, int mul_ld,
                        const void* add, int add_ld, int adaln, int rows_per_batch, float eps, void* stream) {
    return layernorm_bf16(BF(x), ldx, BFM(out), ldo, M, d, BF(mul), mul_ld, BF(add), add_ld, adaln, rows_per_batch, eps,
                          ST(stream));
}
int drag_timestep_embed(const float* t_dev, void* out, int B, void* stream) {
    return timestep_embed(t_dev, BFM(out), B, ST(stream));
}
int drag_euler_step(void* x, int ldx, const void* v, int ldv, int rows, int cols, float dsigma, void* stream) {
    return euler_step(BFM(x), ldx, BF(v), ldv, rows, cols, dsigma, ST(stream));
}
int drag_pack_latents(const void* z, int B, int C, int h, int w, void* out, int64_t ldo, int ch_off, void* stream) {
    return pack_latents(BF(z), B, C, h, w, BFM(out), ldo, ch_off, ST(stream));
}
int drag_unpack_latents(const void* x, int64_t ldx, int B, int C, int h, int w, void* z, void* stream) {
    return unpack_latents(BF(x), ldx, B, C, h, w, BFM(z), ST(stream));
}
int drag_pack_fill_inputs(const void* latents, int64_t ld_lat, const void* masked_latents, const uint8_t* mask, int B, int h,
                          int w, void* x, int64_t ldx, void* stream) {
    return pack_fill_inputs(BF(latents), ld_lat, BF(masked_latents), mask, B, h, w, BFM(x), ldx, ST(stream));
}
int drag_redux_blend(const void* txt, const void* img, const void* pooled, const float* s_embed_dev,
                     const float* s_pool_dev, void* out_embeds, void* out_pooled, int B, int n_txt, int n_img, int dim,
                     int pooled_dim, void* stream) {
    return redux_blend(BF(txt), BF(img), BF(pooled), s_embed_dev, s_pool_dev, BFM(out_embeds), BFM(out_pooled), B,
                       n_txt, n_img, dim, pooled_dim, ST(stream));
}
int drag_l2_normalize(const float* x, float* out, int rows, int d, void* stream) {
    return l2_normalize(x, out, rows, d, ST(stream));
}

int drag_gemm_qkv_split(const void* A, int lda, const void* W, int ldw, int M, int K, int heads, int head_dim,
                        const void* bias, void* q_out, void* k_out, void* v_out, int s_total, int tok_offset,
                        int rows_per_batch, void* stream) {
    GemmEpi e;
    e.mode = EPI_QKV_SPLIT;
    e.bias = BF(bias);
    e.q_out = BFM(q_out); e.k_out = BFM(k_out); e.v_out = BFM(v_out);
    e.heads = heads;
    e.head_dim = head_dim;
    e.s_total = s_total;
    e.tok_offset = tok_offset;
    e.rows_per_batch = rows_per_batch > 0 ? rows_per_batch : (1 << 30);
    return gemm_bf16(BF(A), lda, BF(W), ldw, M, 3 * heads * head_dim, K, e, ST(stream));
}
int drag_vit_patchify(const float* img, void* out, int B, int R, int patch, int kpad, void* stream) {
    return vit_patchify(img, BFM(out), B, R, patch, kpad, ST(stream));
}
int drag_vit_assemble(const void* patch_emb, const void* cls, const void* pos, void* x, int B, int n_patch, int w,
                      void* stream) {
    return vit_assemble(BF(patch_emb), BF(cls), BF(pos), BFM(x), B, n_patch, w, ST(stream));
}

int drag_topk_exchange_buffer_bytes(int world, int nq_cap, int k_cap, int64_t* bytes) {
    DRAG_REQUIRE(bytes && world >= 1 && nq_cap >= 1 && k_cap >= 1, "drag_topk_exchange_buffer_bytes: bad arguments");
    *bytes = static_cast<int64_t>(topk_exchange_buffer_bytes(world, nq_cap, k_cap));
    return DRAG_OK;
}
int drag_topk_exchange_merge(const float* D_loc, const int64_t* I_loc, int nq, int k, void* const* peer_bufs, int world,
                             int rank, int nq_cap, int k_cap, uint32_t epoch, float* D_dev, int64_t* I_dev, void* stream) {
    return topk_exchange_merge(D_loc, I_loc, nq, k, peer_bufs, world, rank, nq_cap, k_cap, epoch, D_dev, I_dev, ST(stream));
}
int drag_index_search_sharded(drag_index_t* h, const float* q_dev, int nq, int k, void* const* peer_bufs, int world, int rank,
                              int nq_cap, int k_cap, uint32_t epoch, float* D_ws, int64_t* I_ws, float* D_dev,
                              int64_t* I_dev, void* stream) {
    DRAG_REQUIRE(h && D_ws && I_ws, "drag_index_search_sharded: null pointer");
    const int rc = index_search_device(reinterpret_cast<Index*>(h), q_dev, nq, k, D_ws, I_ws, ST(stream));
    if (rc != DRAG_OK) return rc;
    return topk_exchange_merge(D_ws, I_ws, nq, k, peer_bufs, world, rank, nq_cap, k_cap, epoch, D_dev, I_dev, ST(stream));
}

// ------------------------------------------------------------------------------------ CLIP ViT engine
int drag_vit_create(const drag_vit_config* cfg, drag_vit_t** out) {
    DRAG_REQUIRE(cfg && out, "drag_vit_create: null pointer");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail(DRAG_ERR_NO_DEVICE, "drag_vit_create: no CUDA device (this library has no CPU path)");
    VitCfg c;
    c.width = cfg->width; c.layers = cfg->layers; c.heads = cfg->heads; c.patch = cfg->patch; c.image = cfg->image;
    c.out_dim = cfg->out_dim; c.max_batch = cfg->max_batch;
    for (int i = 0; i < 3; ++i) { c.mean[i] = cfg->mean[i]; c.std[i] = cfg->std[i]; }
    VitEngine* e = nullptr;
    int rc = vit_create(c, &e);
    if (rc) return rc;
    *out = reinterpret_cast<drag_vit_t*>(e);
    return DRAG_OK;
}
int drag_vit_destroy(drag_vit_t* h) { return vit_destroy(reinterpret_cast<VitEngine*>(h)); }
int drag_vit_set_weights(drag_vit_t* h, const void* const* ptrs, int n) {
    return vit_set_weights(reinterpret_cast<VitEngine*>(h), ptrs, n);
}
int drag_vit_set_option(drag_vit_t* h, int key, int value) { return vit_set_option(reinterpret_cast<VitEngine*>(h), key, value); }
int drag_vit_encode(drag_vit_t* h, const void* img, int img_kind, int B, float* out, int l2_normalize, void* stream) {
    return vit_encode(reinterpret_cast<VitEngine*>(h), img, img_kind, B, out, l2_normalize, ST(stream));
}

// ------------------------------------------------------------------------------------ flux engine
int drag_flux_create(const drag_flux_config* cfg, drag_flux_t** out) {
    DRAG_REQUIRE(cfg && out, "drag_flux_create: null pointer");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail(DRAG_ERR_NO_DEVICE, "drag_flux_create: no CUDA device (this library has no CPU path)");
    FluxCfg c;
    c.in_channels = cfg->in_channels; c.d = cfg->d; c.heads = cfg->heads; c.n_double = cfg->n_double;
    c.n_single = cfg->n_single; c.txt_dim = cfg->txt_dim; c.pooled_dim = cfg->pooled_dim;
    c.out_channels = cfg->out_channels; c.guidance = cfg->guidance; c.max_batch = cfg->max_batch;
    c.max_img_tokens = cfg->max_img_tokens; c.txt_tokens = cfg->txt_tokens;
    FluxEngine* e = nullptr;
    int rc = flux_create(c, &e);
    if (rc) return rc;
    *out = reinterpret_cast<drag_flux_t*>(e);
    return DRAG_OK;
}
int drag_flux_destroy(drag_flux_t* h) { return flux_destroy(reinterpret_cast<FluxEngine*>(h)); }
int drag_flux_set_weights(drag_flux_t* h, const void* const* ptrs, int n) {
    return flux_set_weights(reinterpret_cast<FluxEngine*>(h), ptrs, n);
}
int drag_flux_forward(drag_flux_t* h, const void* x, int ldx, const void* ctx, const void* pooled, const float* t_dev,
                      const float* g_dev, const float* rope_cos, const float* rope_sin, int B, int S_img, void* v_out,
                      int ldv, int n_double_run, int n_single_run, void* stream) {
    return flux_forward(reinterpret_cast<FluxEngine*>(h), BF(x), ldx, BF(ctx), BF(pooled), t_dev, g_dev, rope_cos,
                        rope_sin, B, S_img, BFM(v_out), ldv, n_double_run, n_single_run, ST(stream));
}

int drag_prof_enable(int on) { return prof_enable(on); }
int drag_prof_collect(double* ms, double* work, int* count, int n_classes) {
    DRAG_REQUIRE(ms && work && count && n_classes >= 1, "drag_prof_collect: bad arguments");
    return prof_collect(ms, work, count, n_classes);
}

// ------------------------------------------------------------------------------------ VAE path
int drag_conv2d_nhwc(const void* in, int B, int H, int W, int C_in, const void* w, int C_out, int ksize, int stride,
                     int pad, int Ho, int Wo, int epi_mode, const void* bias, void* out, const void* resid, void* stream) {
    DRAG_REQUIRE(epi_mode == EPI_BIAS || epi_mode == EPI_SILU || epi_mode == EPI_GATE_RESID || epi_mode == EPI_BIAS_F32,
                 "drag_conv2d_nhwc: epi_mode must be 0 (bias), 3 (silu), 4 (+residual) or 6 (fp32 out)");
    GemmEpi e;
    e.mode = epi_mode;
    e.bias = static_cast<const __nv_bfloat16*>(bias);
    if (epi_mode == EPI_BIAS_F32) e.out_f32 = static_cast<float*>(out);
    else e.out = static_cast<__nv_bfloat16*>(out);
    e.ldo = C_out;
    e.resid = static_cast<const __nv_bfloat16*>(resid);
    e.ldr = C_out;
    if (epi_mode == EPI_GATE_RESID) DRAG_REQUIRE(resid, "drag_conv2d_nhwc: residual mode needs resid");
    return conv2d_nhwc_bf16(static_cast<const __nv_bfloat16*>(in), B, H, W, C_in, static_cast<const __nv_bfloat16*>(w), C_out,
                            ksize, stride, pad, Ho, Wo, e, reinterpret_cast<cudaStream_t>(stream));
}

int drag_groupnorm_nhwc(const void* x, void* y, int B, int HW, int C, int groups, const void* gamma, const void* beta,
                        float eps, int silu, float* workspace, int64_t workspace_floats, void* stream) {
    return groupnorm_nhwc(static_cast<const __nv_bfloat16*>(x), static_cast<__nv_bfloat16*>(y), B, HW, C, groups,
                          static_cast<const __nv_bfloat16*>(gamma), static_cast<const __nv_bfloat16*>(beta), eps, silu,
                          workspace, static_cast<size_t>(workspace_floats), reinterpret_cast<cudaStream_t>(stream));
}

int drag_upsample2x_nhwc(const void* x, void* y, int B, int H, int W, int C, void* stream) {
    return upsample2x_nhwc(static_cast<const __nv_bfloat16*>(x), static_cast<__nv_bfloat16*>(y), B, H, W, C,
                           reinterpret_cast<cudaStream_t>(stream));
}

int drag_softmax_rows(const float* s, int64_t ld_s, void* p, int64_t ld_p, int rows, int cols, void* stream) {
    return softmax_rows(s, static_cast<size_t>(ld_s), static_cast<__nv_bfloat16*>(p), static_cast<size_t>(ld_p), rows, cols,
                        reinterpret_cast<cudaStream_t>(stream));
}

int drag_nchw_to_nhwc_pad(const void* in, int in_is_f32, void* out, int B, int C, int H, int W, int C_pad, float scale,
                          float shift, void* stream) {
    return nchw_to_nhwc_pad(in, in_is_f32, static_cast<__nv_bfloat16*>(out), B, C, H, W, C_pad, scale, shift,
                            reinterpret_cast<cudaStream_t>(stream));
}

int drag_nhwc_to_nchw_f32(const void* in, int in_is_f32, int ld, float* out, int B, int C, int H, int W, float scale,
                          float shift, void* stream) {
    return nhwc_to_nchw_f32(in, in_is_f32, ld, out, B, C, H, W, scale, shift, reinterpret_cast<cudaStream_t>(stream));
}

int drag_image_postprocess_u8(const float* in, int ld, uint8_t* out, int64_t pixels, void* stream) {
    return image_postprocess_u8(in, ld, out, static_cast<size_t>(pixels), reinterpret_cast<cudaStream_t>(stream));
}

int drag_image_preprocess_u8(const uint8_t* in, const uint8_t* mask, void* out, int64_t pixels, int C_pad, void* stream) {
    return image_preprocess_u8(in, mask, static_cast<__nv_bfloat16*>(out), static_cast<size_t>(pixels), C_pad,
                               reinterpret_cast<cudaStream_t>(stream));
}

int drag_axpby_bf16(const void* x, const void* y, float a, float b, void* out, int64_t n, void* stream) {
    return axpby_bf16(static_cast<const __nv_bfloat16*>(x), static_cast<const __nv_bfloat16*>(y), a, b,
                      static_cast<__nv_bfloat16*>(out), static_cast<size_t>(n), reinterpret_cast<cudaStream_t>(stream));
}

int drag_launch_count(int64_t* count, int reset) {
    if (!count) return fail(DRAG_ERR_INVALID, "drag_launch_count: null pointer");
    *count = g_launch_count;
    if (reset) g_launch_count = 0;
    return DRAG_OK;
}

int drag_debug_set(int key, int value) {
    if (key == 1) g_attn_v_lbo = static_cast<uint32_t>(value);
    else if (key == 2) g_attn_v_sbo = static_cast<uint32_t>(value);
    else if (key == 3) g_gemm_force_1cta = value;
    else if (key == 4) g_gemm_group_m = value;
    else if (key == 5) g_attn_force_pp = value;
    else if (key == 6) g_gemm_group_n = value;
    else if (key == 7) g_attn_split = value;
    else if (key == 8) g_stem_force_ffma = value;
    else if (key == 9) g_gemm_no_wide_st = value;
    else if (key == 10) g_attn_no_row = value;
    else if (key == 11) g_attn_row_prefetch = value;
    else if (key == 12) g_attn_row_persistent = value;
    else if (key == 13) g_attn_row_stagger = value;
    else if (key == 14) g_attn_row_poly = value;
    else if (key == 15) g_gemm_quick_gelu_mufu = value;
    else if (key == 16) g_attn_row_split = value;
    else if (key == 17) g_attn_row_pair = value;
    else if (key == 18) g_gemm_resid_prefetch = value;
    else return fail(DRAG_ERR_INVALID, "drag_debug_set: unknown key");
    return DRAG_OK;
}

}  // extern "C"
