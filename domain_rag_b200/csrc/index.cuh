// Host-side state of one device-resident inner-product index (shared by topk_scan.cu and capi.cu).
#pragma once
#include <vector>

#include "common.cuh"

namespace drag {

struct Segment {
    float* X;          // device, [N][d] row-major fp32
    int64_t N;
    int64_t base_id;   // id of row 0
    bool owned;        // index-owned (cudaFree on reset) or adopted caller memory
};

struct Index {
    int d = 0;
    int device = 0;
    int sm_count = 148;
    std::vector<Segment> segs;
    int64_t ntotal = 0;
    // workspaces, grown on demand (never on the steady-state hot path)
    uint64_t* partial = nullptr; size_t partial_cap = 0;     // per-CTA key lists
    float* qdev = nullptr;       size_t qdev_cap = 0;        // staging for the host-buffer API
    float* Ddev = nullptr;       int64_t* Idev = nullptr; size_t out_cap = 0;
    uint32_t* seg_start = nullptr; int64_t* seg_base = nullptr; int seg_tab_n = -1;
    // geometry of the last scan launch (roofline arithmetic in bench.py)
    int last_grid = 0, last_stages = 0, last_rps = 0, last_nqb = 0;
    // optional CUDA-event bracket around the scan launches (excludes the merge kernel)
    bool timing = false;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
};

int index_search_device(Index* ix, const float* q, int nq, int k, float* D, int64_t* I, cudaStream_t st);
int index_ensure_io(Index* ix, int nq, int k);
int merge_pairs_device(const float* scores, const int64_t* ids, int nq, int lists, int k_in, int k_out,
                       float* D, int64_t* I, cudaStream_t st);
// Fused exchange + merge of per-shard top-k lists over NVLink peer memory (topk_exchange.cu).
size_t topk_exchange_buffer_bytes(int world, int nq_cap, int k_cap);
int topk_exchange_merge(const float* D_loc, const int64_t* I_loc, int nq, int k, void* const* peer_bufs, int world, int rank,
                        int nq_cap, int k_cap, uint32_t epoch, float* D, int64_t* I, cudaStream_t st);
int stem_stats_device(const float* img, int B, int H, int W, const float* w_fold, const float* b_fold,
                      float eps, float* out, cudaStream_t st);

}  // namespace drag
