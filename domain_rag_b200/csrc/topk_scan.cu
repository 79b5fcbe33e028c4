// Exact inner-product scan x top-k over a device-resident fp32 corpus (the faiss.IndexFlatIP
// add/search pair the reference builds per query: retrieval/clip100_resnet_style_all_shots.py:425-434).
//
// HBM-bound: N*D*4 bytes are read exactly once per query batch. One persistent CTA per SM:
//   warp 0      : producer - one elected lane streams contiguous row blocks (<=16 KB) of X into a
//                 shared-memory ring with 1-D bulk async copies (cp.async.bulk, SASS UBLKCP),
//                 completion tracked by mbarrier transaction bytes;
//   warps 1..12 : consumers - each owns whole ring stages round-robin, computes the dot products
//                 with a fixed summation order (per-lane sequential FMA, then xor-shuffle tree, so a
//                 row's score does not depend on grid size or sharding), and keeps a per-CTA
//                 running top-k: candidates above the CTA's current k-th key go to a warp-private
//                 buffer that is merged into the CTA list by a warp bitonic sort under a per-query
//                 lock.
// A second tiny kernel merges the per-CTA lists with an 8-pass radix select over 64-bit keys.
// Keys are (orderable fp32 score << 32) | (0xFFFFFFFF - row ordinal): descending key order is
// "score descending, lower index first", the tie rule the oracle uses (oracle/ip_topk.py).
#include <float.h>

#include <vector>

#include "common.cuh"
#include "index.cuh"
#include "ptx.cuh"

namespace drag {

constexpr int SCAN_CW = 12;                      // consumer warps per CTA
constexpr int SCAN_THREADS = 32 * (SCAN_CW + 1); // + producer warp
constexpr int SCAN_WB = 64;                      // capacity of the pending buffer per (warp, query)
constexpr int SCAN_WB_FLUSH = 32;                // merge into the CTA list once this many are pending
constexpr int SCAN_STAGE_TARGET = 16384;         // bytes per ring stage
constexpr int SCAN_MAX_STAGES = 24;
constexpr int SCAN_SMEM_BUDGET = 227 * 1024;
constexpr int TOPK_KMAX = 1024;

__device__ __forceinline__ uint32_t order_f32(float f) {
    if (f != f) f = -INFINITY;  // NaN ranks last
    f += 0.0f;                  // -0 -> +0
    uint32_t b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float unorder_f32(uint32_t u) {
    uint32_t b = (u & 0x80000000u) ? (u & 0x7FFFFFFFu) : ~u;
    return __uint_as_float(b);
}
__device__ __forceinline__ uint64_t make_key(float score, uint32_t ordinal) {
    return (static_cast<uint64_t>(order_f32(score)) << 32) | (0xFFFFFFFFu - ordinal);
}

struct ScanSmem {
    float* q;              // [nq][D]
    uint64_t* lists;       // [nq][2][k]   ping-pong, current copy sorted descending, zero padded
    uint64_t* thr;         // [nq]         current k-th key (0 = list not full)
    uint64_t* wbuf;        // [CW][nq][WB]
    int* wcnt;             // [CW][nq]
    int* locks;            // [nq]
    int* cur;              // [nq]         which ping-pong copy is current
    int nq, k;
};

__device__ __forceinline__ uint64_t shfl_xor_u64(uint64_t v, int m) {
    uint32_t lo = __shfl_xor_sync(0xffffffffu, static_cast<uint32_t>(v), m);
    uint32_t hi = __shfl_xor_sync(0xffffffffu, static_cast<uint32_t>(v >> 32), m);
    return (static_cast<uint64_t>(hi) << 32) | lo;
}

// Merge this warp's pending candidates (<= 64, unsorted) into the CTA's sorted top-k list:
// register bitonic sort of the candidates (element i = r*32 + lane), then every element of both
// sorted sequences finds its merged rank by binary search in the other one (keys are unique) and
// is scattered into the other ping-pong copy. ~1 us instead of a full smem sort.
// Lock protocol: test-and-test-and-set with back-off. Waiters poll with a plain shared load and
// sleep between polls, so they do not saturate the shared-memory atomic unit the lock holder's own
// loads go through (a tight atomicCAS spin by 11 warps slowed every merge ~10x).
__device__ __forceinline__ bool try_lock(int* lock, int lane) {
    int got = 0;
    if (lane == 0) got = (*reinterpret_cast<volatile int*>(lock) == 0 && atomicCAS(lock, 0, 1) == 0) ? 1 : 0;
    return __shfl_sync(0xffffffffu, got, 0) != 0;
}

// blocking = false: give up if another warp holds the list (caller retries after its next stage).
__device__ __forceinline__ bool flush_candidates(const ScanSmem& s, int cw, int q, int lane,
                                                 bool blocking = true) {
    const int n = s.wcnt[cw * s.nq + q];
    if (n == 0) return true;
    int* lock = &s.locks[q];
    while (!try_lock(lock, lane)) {
        if (!blocking) return false;
        __nanosleep(200);
    }
    __threadfence_block();
    uint64_t* wb = s.wbuf + (static_cast<size_t>(cw) * s.nq + q) * SCAN_WB;
    uint64_t e0 = (lane < n) ? wb[lane] : 0ull;
    uint64_t e1 = (lane + 32 < n) ? wb[lane + 32] : 0ull;
#pragma unroll
    for (int k2 = 2; k2 <= 64; k2 <<= 1) {
#pragma unroll
        for (int j = k2 >> 1; j > 0; j >>= 1) {
            if (j == 32) {
                const uint64_t mx = e0 > e1 ? e0 : e1, mn = e0 > e1 ? e1 : e0;   // k2 == 64: descending
                e0 = mx;
                e1 = mn;
            } else {
                const uint64_t p0 = shfl_xor_u64(e0, j), p1 = shfl_xor_u64(e1, j);
                const bool lower = (lane & j) == 0;
                const bool desc0 = (lane & k2) == 0;            // element index lane
                const bool desc1 = ((lane + 32) & k2) == 0;     // element index lane + 32
                e0 = ((lower == desc0) == (e0 > p0)) ? e0 : p0;
                e1 = ((lower == desc1) == (e1 > p1)) ? e1 : p1;
            }
        }
    }
    wb[lane] = e0;
    wb[lane + 32] = e1;
    __syncwarp();
    const int cur = s.cur[q];
    const uint64_t* src = s.lists + (static_cast<size_t>(q) * 2 + cur) * s.k;
    uint64_t* dst = s.lists + (static_cast<size_t>(q) * 2 + (cur ^ 1)) * s.k;
    // candidates: rank = own index + #list entries greater
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const uint64_t b = r ? e1 : e0;
        if (b != 0ull) {
            int lo = 0, hi = s.k;
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (src[mid] > b) lo = mid + 1; else hi = mid;
            }
            const int rank = r * 32 + lane + lo;
            if (rank < s.k) dst[rank] = b;
        }
    }
    // list entries: rank = own index + #candidates greater
    for (int i = lane; i < s.k; i += 32) {
        const uint64_t a = src[i];
        if (a == 0ull) break;   // zero padding from here on (sorted descending)
        int lo = 0, hi = SCAN_WB;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (wb[mid] > a) lo = mid + 1; else hi = mid;
        }
        const int rank = i + lo;
        if (rank < s.k) dst[rank] = a;
    }
    __syncwarp();
    if (lane == 0) {
        reinterpret_cast<volatile uint64_t*>(s.thr)[q] = dst[s.k - 1];
        s.cur[q] = cur ^ 1;
        s.wcnt[cw * s.nq + q] = 0;
    }
    __threadfence_block();
    __syncwarp();
    if (lane == 0) atomicExch(lock, 0);
    __syncwarp();
    return true;
}

// Append one accepted candidate to the warp-private pending buffer (no merge here: the caller
// flushes between ring stages, after the stage has been handed back to the producer).
__device__ __forceinline__ void push_candidate(const ScanSmem& s, int cw, int q, uint64_t key,
                                               int lane) {
    const int cnt = s.wcnt[cw * s.nq + q];
    __syncwarp();
    if (lane == 0) {
        s.wbuf[(static_cast<size_t>(cw) * s.nq + q) * SCAN_WB + cnt] = key;
        s.wcnt[cw * s.nq + q] = cnt + 1;
    }
    __syncwarp();
}

// Called between ring stages. `next_rows` bounds how many candidates the next stage can add per
// query: the merge is only forced (blocking) when the pending buffer could overflow.
__device__ __forceinline__ void flush_if_needed(const ScanSmem& s, int cw, int lane, int next_rows) {
    for (int q = 0; q < s.nq; ++q) {
        const int cnt = s.wcnt[cw * s.nq + q];
        if (cnt >= SCAN_WB_FLUSH) flush_candidates(s, cw, q, lane, cnt + next_rows > SCAN_WB);
    }
}

__device__ __forceinline__ float warp_sum(float v) {
    v += __shfl_xor_sync(0xffffffffu, v, 16);
    v += __shfl_xor_sync(0xffffffffu, v, 8);
    v += __shfl_xor_sync(0xffffffffu, v, 4);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    return v;
}

__device__ __forceinline__ float dot4(const float4& x, const float4& q, float acc) {
    acc = fmaf(x.x, q.x, acc);
    acc = fmaf(x.y, q.y, acc);
    acc = fmaf(x.z, q.z, acc);
    acc = fmaf(x.w, q.w, acc);
    return acc;
}

// NV > 0: D == NV*128, row fragments live in registers (4 rows at a time) and are reused across
// the query batch. NV == 0: any D % 4 == 0, operands re-read from shared memory.
template <int NV>
__device__ __forceinline__ void consume_rows(const ScanSmem& s, const float* stage, int rows, int D,
                                             uint32_t ord0, int cw, int lane) {
    const float4* xs = reinterpret_cast<const float4*>(stage);
    const int D4 = D >> 2;
    const volatile uint64_t* thr = s.thr;
    if constexpr (NV > 0) {
        constexpr int RB = (NV <= 4) ? 4 : 2;  // rows held in registers at once
        for (int r = 0; r < rows; r += RB) {
            float4 x[RB][NV];
#pragma unroll
            for (int rr = 0; rr < RB; ++rr) {
                int row = min(r + rr, rows - 1);
#pragma unroll
                for (int it = 0; it < NV; ++it) x[rr][it] = xs[row * D4 + it * 32 + lane];
            }
            for (int q = 0; q < s.nq; ++q) {
                const float4* qs = reinterpret_cast<const float4*>(s.q + static_cast<size_t>(q) * D);
                float acc[RB];
#pragma unroll
                for (int rr = 0; rr < RB; ++rr) acc[rr] = 0.f;
#pragma unroll
                for (int it = 0; it < NV; ++it) {
                    float4 qv = qs[it * 32 + lane];
#pragma unroll
                    for (int rr = 0; rr < RB; ++rr) acc[rr] = dot4(x[rr][it], qv, acc[rr]);
                }
#pragma unroll
                for (int rr = 0; rr < RB; ++rr) acc[rr] = warp_sum(acc[rr]);
                uint64_t t = thr[q];
#pragma unroll
                for (int rr = 0; rr < RB; ++rr) {
                    if (r + rr < rows) {
                        uint64_t key = make_key(acc[rr], ord0 + r + rr);
                        if (key > t) {
                            push_candidate(s, cw, q, key, lane);
                            t = thr[q];
                        }
                    }
                }
            }
        }
    } else {
        for (int r = 0; r < rows; ++r) {
            for (int q = 0; q < s.nq; ++q) {
                const float4* qs = reinterpret_cast<const float4*>(s.q + static_cast<size_t>(q) * D);
                float acc = 0.f;
                for (int v = lane; v < D4; v += 32) acc = dot4(xs[r * D4 + v], qs[v], acc);
                acc = warp_sum(acc);
                uint64_t key = make_key(acc, ord0 + r);
                if (key > thr[q]) push_candidate(s, cw, q, key, lane);
            }
        }
    }
}

struct ScanArgs {
    const float* X;       // [N][D] device, 16-byte aligned rows
    const float* Q;       // [nq][D] device
    uint64_t* partial;    // [nq][lists_total][k] keys
    int64_t N;
    int D, nq, k;
    int rps;              // rows per ring stage
    int stages;
    int chunks_per_cta;
    int lists_total, list_off;  // this launch fills lists [list_off, list_off + gridDim.x)
    uint32_t ord_base;          // ordinal of row 0 of this segment
};

template <int NV>
__global__ void __launch_bounds__(SCAN_THREADS, 1) ip_scan_topk_kernel(ScanArgs a) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int stage_bytes = a.rps * a.D * 4;
    // carve shared memory
    uint8_t* p = smem;
    float* ring = reinterpret_cast<float*>(p);
    p += static_cast<size_t>(a.stages) * stage_bytes;
    ScanSmem s;
    s.nq = a.nq; s.k = a.k;
    s.lists = reinterpret_cast<uint64_t*>(p); p += static_cast<size_t>(a.nq) * 2 * a.k * 8;
    s.wbuf = reinterpret_cast<uint64_t*>(p);  p += static_cast<size_t>(SCAN_CW) * a.nq * SCAN_WB * 8;
    s.thr = reinterpret_cast<uint64_t*>(p);   p += static_cast<size_t>((a.nq + 1) & ~1) * 8;
    uint64_t* full = reinterpret_cast<uint64_t*>(p);  p += SCAN_MAX_STAGES * 8;
    uint64_t* empty = reinterpret_cast<uint64_t*>(p); p += SCAN_MAX_STAGES * 8;
    s.q = reinterpret_cast<float*>(p);        p += static_cast<size_t>(a.nq) * a.D * 4;
    s.wcnt = reinterpret_cast<int*>(p);       p += static_cast<size_t>(SCAN_CW) * a.nq * 4;
    s.locks = reinterpret_cast<int*>(p);      p += static_cast<size_t>(a.nq) * 4;
    s.cur = reinterpret_cast<int*>(p);

    // chunk range of this CTA
    const int64_t total_chunks = (a.N + a.rps - 1) / a.rps;
    const int64_t c_begin = static_cast<int64_t>(blockIdx.x) * a.chunks_per_cta;
    int64_t c_end = c_begin + a.chunks_per_cta;
    if (c_end > total_chunks) c_end = total_chunks;
    const int nchunks = (c_end > c_begin) ? static_cast<int>(c_end - c_begin) : 0;

    for (int i = threadIdx.x; i < a.nq * 2 * a.k; i += blockDim.x) s.lists[i] = 0ull;
    for (int i = threadIdx.x; i < a.nq * a.D; i += blockDim.x) s.q[i] = a.Q[i];
    for (int i = threadIdx.x; i < SCAN_CW * a.nq; i += blockDim.x) s.wcnt[i] = 0;
    if (threadIdx.x < a.nq) {
        s.thr[threadIdx.x] = 0ull;
        s.locks[threadIdx.x] = 0;
        s.cur[threadIdx.x] = 0;
    }
    if (threadIdx.x == 0) {
        for (int i = 0; i < a.stages; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], 1);
        }
        fence_mbar_init();
    }
    __syncthreads();

    if (warp == 0) {
        if (lane == 0) {
            for (int c = 0; c < nchunks; ++c) {
                const int st = c % a.stages, it = c / a.stages;
                if (it > 0) mbar_wait(&empty[st], (it - 1) & 1);
                const int64_t row0 = (c_begin + c) * a.rps;
                int64_t rows = a.N - row0;
                if (rows > a.rps) rows = a.rps;
                const uint32_t bytes = static_cast<uint32_t>(rows) * a.D * 4;
                mbar_arrive_expect_tx(&full[st], bytes);
                bulk_g2s(reinterpret_cast<uint8_t*>(ring) + static_cast<size_t>(st) * stage_bytes,
                         a.X + row0 * a.D, bytes, &full[st]);
            }
        }
        return;
    }

    // Stage st is owned by consumer warp st % CW for the whole kernel: the owner observes every
    // phase of full[st] in order, so the parity wait can never alias a phase it skipped.
    const int cw = warp - 1;
    for (int base = 0, it = 0; base < nchunks; base += a.stages, ++it) {
        for (int st = cw; st < a.stages; st += SCAN_CW) {
            const int c = base + st;
            if (c >= nchunks) break;
            mbar_wait(&full[st], it & 1);
            const int64_t row0 = (c_begin + c) * a.rps;
            int64_t rows = a.N - row0;
            if (rows > a.rps) rows = a.rps;
            const float* stage = reinterpret_cast<const float*>(reinterpret_cast<uint8_t*>(ring) +
                                                                static_cast<size_t>(st) * stage_bytes);
            consume_rows<NV>(s, stage, static_cast<int>(rows), a.D,
                             a.ord_base + static_cast<uint32_t>(row0), cw, lane);
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[st]);   // hand the stage back before any list merge
            flush_if_needed(s, cw, lane, a.rps);
        }
    }
    for (int q = 0; q < a.nq; ++q) flush_candidates(s, cw, q, lane);
    asm volatile("bar.sync 1, %0;" ::"r"(SCAN_CW * 32) : "memory");
    const int ct = threadIdx.x - 32;
    for (int i = ct; i < a.nq * a.k; i += SCAN_CW * 32) {
        const int q = i / a.k, j = i - q * a.k;
        a.partial[(static_cast<size_t>(q) * a.lists_total + a.list_off + blockIdx.x) * a.k + j] =
            s.lists[(static_cast<size_t>(q) * 2 + s.cur[q]) * a.k + j];
    }
}

// Fallback for D % 4 != 0 or unaligned corpora: consumers read global memory directly.
__global__ void __launch_bounds__(SCAN_CW * 32, 1) ip_scan_topk_direct_kernel(ScanArgs a) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int cw = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint8_t* p = smem;
    ScanSmem s;
    s.nq = a.nq; s.k = a.k;
    s.lists = reinterpret_cast<uint64_t*>(p); p += static_cast<size_t>(a.nq) * 2 * a.k * 8;
    s.wbuf = reinterpret_cast<uint64_t*>(p);  p += static_cast<size_t>(SCAN_CW) * a.nq * SCAN_WB * 8;
    s.thr = reinterpret_cast<uint64_t*>(p);   p += static_cast<size_t>((a.nq + 1) & ~1) * 8;
    s.q = reinterpret_cast<float*>(p);        p += static_cast<size_t>(a.nq) * a.D * 4;
    s.wcnt = reinterpret_cast<int*>(p);       p += static_cast<size_t>(SCAN_CW) * a.nq * 4;
    s.locks = reinterpret_cast<int*>(p);      p += static_cast<size_t>(a.nq) * 4;
    s.cur = reinterpret_cast<int*>(p);
    for (int i = threadIdx.x; i < a.nq * 2 * a.k; i += blockDim.x) s.lists[i] = 0ull;
    for (int i = threadIdx.x; i < a.nq * a.D; i += blockDim.x) s.q[i] = a.Q[i];
    for (int i = threadIdx.x; i < SCAN_CW * a.nq; i += blockDim.x) s.wcnt[i] = 0;
    if (threadIdx.x < a.nq) {
        s.thr[threadIdx.x] = 0ull;
        s.locks[threadIdx.x] = 0;
        s.cur[threadIdx.x] = 0;
    }
    __syncthreads();
    const int64_t rows_per_cta = static_cast<int64_t>(a.chunks_per_cta) * a.rps;
    const int64_t r_begin = blockIdx.x * rows_per_cta;
    int64_t r_end = r_begin + rows_per_cta;
    if (r_end > a.N) r_end = a.N;
    const volatile uint64_t* thr = s.thr;
    for (int64_t r = r_begin + cw; r < r_end; r += SCAN_CW) {
        const float* x = a.X + r * a.D;
        for (int q = 0; q < a.nq; ++q) {
            const float* qs = s.q + static_cast<size_t>(q) * a.D;
            float acc = 0.f;
            for (int v = lane; v < a.D; v += 32) acc = fmaf(x[v], qs[v], acc);
            acc = warp_sum(acc);
            uint64_t key = make_key(acc, a.ord_base + static_cast<uint32_t>(r));
            if (key > thr[q]) push_candidate(s, cw, q, key, lane);
        }
        flush_if_needed(s, cw, lane, 1);
    }
    for (int q = 0; q < a.nq; ++q) flush_candidates(s, cw, q, lane);
    __syncthreads();
    for (int i = threadIdx.x; i < a.nq * a.k; i += blockDim.x) {
        const int q = i / a.k, j = i - q * a.k;
        a.partial[(static_cast<size_t>(q) * a.lists_total + a.list_off + blockIdx.x) * a.k + j] =
            s.lists[(static_cast<size_t>(q) * 2 + s.cur[q]) * a.k + j];
    }
}

// ---------------------------------------------------------------------------------------------
// Merge of per-CTA key lists: radix select of the k-th largest 64-bit key, gather, bitonic sort.
struct MergeArgs {
    const uint64_t* partial;   // [nq][lists][k]
    int lists, k, kpad;        // kpad = pow2 >= k
    const uint32_t* seg_start; // [nseg+1] cumulative ordinals
    const int64_t* seg_base;   // [nseg]
    int nseg;
    float* D;                  // [nq][k]
    int64_t* I;                // [nq][k]
};

constexpr int MERGE_THREADS = 1024;

__global__ void __launch_bounds__(MERGE_THREADS, 1) topk_merge_keys_kernel(MergeArgs a) {
    __shared__ uint32_t hist[256];
    __shared__ uint32_t suf_local[256];
    __shared__ uint32_t wtot[8];
    __shared__ uint64_t sel[TOPK_KMAX];
    __shared__ uint64_t s_prefix, s_mask;
    __shared__ int s_remaining, s_count;
    const int q = blockIdx.x, tid = threadIdx.x;
    const uint64_t* keys = a.partial + static_cast<size_t>(q) * a.lists * a.k;
    const int T = a.lists * a.k;

    if (tid == 0) {
        s_prefix = 0ull;
        s_mask = 0ull;
        s_remaining = a.k;
        s_count = 0;
    }
    __syncthreads();
    // k-th largest key (0 if fewer than k valid keys exist; valid keys are never 0)
    for (int pass = 7; pass >= 0; --pass) {
        if (tid < 256) hist[tid] = 0;
        __syncthreads();
        const uint64_t prefix = s_prefix, mask = s_mask;
        const uint32_t rem = static_cast<uint32_t>(s_remaining);
        const int shift = pass * 8;
        for (int i = tid; i < T; i += MERGE_THREADS) {
            uint64_t key = keys[i];
            if ((key & mask) == prefix) atomicAdd(&hist[(key >> shift) & 0xFF], 1u);
        }
        __syncthreads();
        // suffix sums over the 256 digit buckets: warp shuffles inside each of the 8 warps, then
        // the totals of the higher warps; the selected bucket is the one whose suffix sum first
        // reaches the number of keys still wanted.
        if (tid < 256) {
            const uint32_t h = hist[tid];
            uint32_t v = h;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                uint32_t o = __shfl_down_sync(0xffffffffu, v, off);
                if ((tid & 31) + off < 32) v += o;
            }
            if ((tid & 31) == 0) wtot[tid >> 5] = v;
            suf_local[tid] = v;
        }
        __syncthreads();
        if (tid < 256) {
            uint32_t higher = 0;
            for (int w = (tid >> 5) + 1; w < 8; ++w) higher += wtot[w];
            const uint32_t incl = suf_local[tid] + higher;      // sum over buckets >= tid
            const uint32_t excl = incl - hist[tid];             // sum over buckets >  tid
            if ((incl >= rem && excl < rem) || (tid == 0 && incl < rem)) {
                s_remaining = static_cast<int>(rem - (incl >= rem ? excl : 0));
                s_prefix = prefix | (static_cast<uint64_t>(tid) << shift);
                s_mask = mask | (0xFFull << shift);
            }
        }
        __syncthreads();
    }
    uint64_t kth = s_prefix;
    if (kth == 0ull) kth = 1ull;  // fewer than k valid keys: take every valid one
    for (int i = tid; i < a.kpad; i += MERGE_THREADS) sel[i] = 0ull;
    __syncthreads();
    for (int i = tid; i < T; i += MERGE_THREADS) {
        uint64_t key = keys[i];
        if (key >= kth) {
            int pos = atomicAdd(&s_count, 1);
            if (pos < a.kpad) sel[pos] = key;
        }
    }
    __syncthreads();
    // bitonic sort of kpad keys, descending
    for (int k2 = 2; k2 <= a.kpad; k2 <<= 1) {
        for (int j = k2 >> 1; j > 0; j >>= 1) {
            for (int t = tid; t < (a.kpad >> 1); t += MERGE_THREADS) {
                int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                int l = i | j;
                bool desc = ((i & k2) == 0);
                uint64_t x = sel[i], y = sel[l];
                bool swap = desc ? (x < y) : (x > y);
                if (swap) {
                    sel[i] = y;
                    sel[l] = x;
                }
            }
            __syncthreads();
        }
    }
    for (int j = tid; j < a.k; j += MERGE_THREADS) {
        uint64_t key = sel[j];
        float score = -FLT_MAX;
        int64_t id = -1;
        if (key != 0ull) {
            score = unorder_f32(static_cast<uint32_t>(key >> 32));
            uint32_t ord = 0xFFFFFFFFu - static_cast<uint32_t>(key & 0xFFFFFFFFull);
            int sg = 0;
            while (sg + 1 < a.nseg && ord >= a.seg_start[sg + 1]) ++sg;
            id = a.seg_base[sg] + static_cast<int64_t>(ord - a.seg_start[sg]);
        }
        a.D[static_cast<size_t>(q) * a.k + j] = score;
        a.I[static_cast<size_t>(q) * a.k + j] = id;
    }
}

// Merge of (score, id) lists gathered from several shards (the NCCL all-gather of per-shard
// top-k): order by score descending, id ascending; id < 0 marks an empty slot.
constexpr int PAIRS_MAX = 8192;
__global__ void __launch_bounds__(1024, 1)
topk_merge_pairs_kernel(const float* scores, const int64_t* ids, int lists, int k_in, int k_out,
                        int npad, float* D, int64_t* I) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint32_t* sk = reinterpret_cast<uint32_t*>(smem);                       // [npad] orderable score
    int64_t* si = reinterpret_cast<int64_t*>(smem + sizeof(uint32_t) * PAIRS_MAX);  // [npad]
    const int q = blockIdx.x, tid = threadIdx.x;
    const int T = lists * k_in;
    for (int i = tid; i < npad; i += blockDim.x) {
        if (i < T) {
            int64_t id = ids[static_cast<size_t>(q) * T + i];
            si[i] = id;
            sk[i] = (id < 0) ? 0u : order_f32(scores[static_cast<size_t>(q) * T + i]);
        } else {
            si[i] = -1;
            sk[i] = 0u;
        }
    }
    __syncthreads();
    auto before = [](uint32_t ka, int64_t ia, uint32_t kb, int64_t ib) {
        // true if a must come before b: valid first, score desc, id asc
        bool va = ia >= 0, vb = ib >= 0;
        if (va != vb) return va;
        if (ka != kb) return ka > kb;
        return ia < ib;
    };
    for (int k2 = 2; k2 <= npad; k2 <<= 1) {
        for (int j = k2 >> 1; j > 0; j >>= 1) {
            for (int t = tid; t < (npad >> 1); t += blockDim.x) {
                int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                int l = i | j;
                bool fwd = ((i & k2) == 0);
                uint32_t kx = sk[i], ky = sk[l];
                int64_t ix = si[i], iy = si[l];
                bool swap = fwd ? before(ky, iy, kx, ix) : before(kx, ix, ky, iy);
                if (swap) {
                    sk[i] = ky; sk[l] = kx;
                    si[i] = iy; si[l] = ix;
                }
            }
            __syncthreads();
        }
    }
    for (int j = tid; j < k_out; j += blockDim.x) {
        bool valid = (j < npad) && si[j] >= 0;
        D[static_cast<size_t>(q) * k_out + j] = valid ? unorder_f32(sk[j]) : -FLT_MAX;
        I[static_cast<size_t>(q) * k_out + j] = valid ? si[j] : -1;
    }
}

// ---------------------------------------------------------------------------------------------
// Host side: the index object behind drag_index_*.
static int next_pow2(int v) {
    int p = 1;
    while (p < v) p <<= 1;
    return p;
}

struct ScanPlan {
    bool bulk;     // false: direct kernel
    int nv;        // template selector (0 generic)
    int rps, stages, nqb;
    size_t smem;
};

static size_t scan_fixed_smem(int nqb, int D, int k) {
    return static_cast<size_t>(nqb) * 2 * k * 8 + static_cast<size_t>(nqb) * 4 + static_cast<size_t>(SCAN_CW) * nqb * SCAN_WB * 8 +
           static_cast<size_t>(nqb) * 8 + 2 * SCAN_MAX_STAGES * 8 + static_cast<size_t>(nqb) * D * 4 +
           static_cast<size_t>(SCAN_CW) * nqb * 4 + static_cast<size_t>(nqb) * 4 + 256;
}

static int make_plan(int D, int nq, int k, bool aligned, ScanPlan* plan) {
    plan->rps = 8; plan->stages = 0; plan->smem = 0;
    plan->bulk = aligned && (D % 4 == 0) && (static_cast<size_t>(D) * 4 <= 65536);
    plan->nv = (plan->bulk && D % 128 == 0 && D / 128 >= 1 && D / 128 <= 8) ? D / 128 : 0;
    int nqb = nq < 8 ? nq : 8;
    while (nqb > 1 && scan_fixed_smem(nqb, D, k) > SCAN_SMEM_BUDGET / 2) nqb >>= 1;
    plan->nqb = nqb;
    size_t fixed = scan_fixed_smem(nqb, D, k);
    if (fixed > static_cast<size_t>(SCAN_SMEM_BUDGET))
        return fail(DRAG_ERR_UNSUPPORTED, "index search: k/d too large for shared memory");
    if (plan->bulk) {
        int rps = SCAN_STAGE_TARGET / (D * 4);
        if (rps < 1) rps = 1;
        if (rps > 32) rps = 32;
        if (plan->nv > 0) rps = (rps / 4) * 4 > 0 ? (rps / 4) * 4 : 4;
        plan->rps = rps;
        size_t stage_bytes = static_cast<size_t>(rps) * D * 4;
        int stages = static_cast<int>((SCAN_SMEM_BUDGET - fixed) / stage_bytes);
        if (stages > SCAN_MAX_STAGES) stages = SCAN_MAX_STAGES;
        if (stages > SCAN_CW) stages = (stages / SCAN_CW) * SCAN_CW;  // balanced stage ownership
        if (stages < 2) {
            plan->bulk = false;
        } else {
            plan->stages = stages;
            plan->smem = fixed + stage_bytes * stages;
        }
    }
    if (!plan->bulk) {
        plan->nv = 0;
        plan->rps = 8;
        plan->stages = 0;
        plan->smem = fixed;
    }
    return DRAG_OK;
}

template <int NV>
static cudaError_t launch_scan(const ScanArgs& a, int grid, size_t smem, cudaStream_t st) {
    cudaError_t e = cudaFuncSetAttribute(ip_scan_topk_kernel<NV>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(smem));
    if (e != cudaSuccess) return e;
    ip_scan_topk_kernel<NV><<<grid, SCAN_THREADS, smem, st>>>(a); count_launch();
    return cudaGetLastError();
}

static int grid_for_segment(const Index* ix, const ScanPlan& plan, int64_t N, int* chunks_per_cta) {
    int64_t total_chunks = (N + plan.rps - 1) / plan.rps;
    int grid = ix->sm_count;
    if (total_chunks < grid) grid = static_cast<int>(total_chunks);
    if (grid < 1) grid = 1;
    *chunks_per_cta = static_cast<int>((total_chunks + grid - 1) / grid);
    grid = static_cast<int>((total_chunks + *chunks_per_cta - 1) / *chunks_per_cta);
    if (grid < 1) grid = 1;
    return grid;
}

// Device-pointer search on `stream`: q [nq][d], D [nq][k], I [nq][k] all on the device.
int index_search_device(Index* ix, const float* q, int nq, int k, float* D, int64_t* I,
                        cudaStream_t st) {
    DRAG_REQUIRE(ix != nullptr, "index_search: null index");
    DRAG_REQUIRE(nq >= 0 && k >= 1 && k <= TOPK_KMAX, "index_search: need 1 <= k <= 1024");
    if (nq == 0) return DRAG_OK;
    DRAG_REQUIRE(q && D && I, "index_search: null pointer");
    DRAG_REQUIRE(ix->ntotal < 0xFFFFFFFFll, "index_search: more than 2^32-1 rows on one device");
    DRAG_CUDA(cudaSetDevice(ix->device));

    bool aligned = true;
    for (const Segment& sg : ix->segs)
        if ((reinterpret_cast<uintptr_t>(sg.X) & 15) != 0) aligned = false;
    ScanPlan plan;
    int rc = make_plan(ix->d, nq, k, aligned, &plan);
    if (rc) return rc;

    // list layout across segments
    std::vector<int> grids(ix->segs.size()), cpcs(ix->segs.size());
    int lists_total = 0;
    for (size_t i = 0; i < ix->segs.size(); ++i) {
        grids[i] = ix->segs[i].N > 0 ? grid_for_segment(ix, plan, ix->segs[i].N, &cpcs[i]) : 0;
        lists_total += grids[i];
    }
    const int nseg = static_cast<int>(ix->segs.size());
    size_t need = static_cast<size_t>(nq) * (lists_total > 0 ? lists_total : 1) * k;
    if (need > ix->partial_cap) {
        if (ix->partial) DRAG_CUDA(cudaFree(ix->partial));
        DRAG_CUDA(cudaMalloc(&ix->partial, need * sizeof(uint64_t)));
        ix->partial_cap = need;
    }
    if (ix->seg_tab_n != nseg) {
        if (ix->seg_start) DRAG_CUDA(cudaFree(ix->seg_start));
        if (ix->seg_base) DRAG_CUDA(cudaFree(ix->seg_base));
        std::vector<uint32_t> ss(nseg + 1, 0);
        std::vector<int64_t> sb(nseg > 0 ? nseg : 1, 0);
        for (int i = 0; i < nseg; ++i) {
            ss[i + 1] = ss[i] + static_cast<uint32_t>(ix->segs[i].N);
            sb[i] = ix->segs[i].base_id;
        }
        DRAG_CUDA(cudaMalloc(&ix->seg_start, (nseg + 1) * sizeof(uint32_t)));
        DRAG_CUDA(cudaMalloc(&ix->seg_base, sb.size() * sizeof(int64_t)));
        DRAG_CUDA(cudaMemcpy(ix->seg_start, ss.data(), (nseg + 1) * sizeof(uint32_t),
                             cudaMemcpyHostToDevice));
        DRAG_CUDA(cudaMemcpy(ix->seg_base, sb.data(), sb.size() * sizeof(int64_t),
                             cudaMemcpyHostToDevice));
        ix->seg_tab_n = nseg;
    }
    if (lists_total == 0) {
        // empty index: a single all-empty list makes the merge emit (-FLT_MAX, -1)
        DRAG_CUDA(cudaMemsetAsync(ix->partial, 0, need * sizeof(uint64_t), st));
        lists_total = 1;
    } else {
        if (ix->timing) DRAG_CUDA(cudaEventRecord(ix->ev0, st));
        for (int q0 = 0; q0 < nq; q0 += plan.nqb) {
            const int nqb = (nq - q0 < plan.nqb) ? (nq - q0) : plan.nqb;
            int list_off = 0;
            uint32_t ord = 0;
            for (int i = 0; i < nseg; ++i) {
                const Segment& sg = ix->segs[i];
                if (sg.N > 0) {
                    ScanArgs a;
                    a.X = sg.X;
                    a.Q = q + static_cast<size_t>(q0) * ix->d;
                    a.partial = ix->partial + static_cast<size_t>(q0) * lists_total * k;
                    a.N = sg.N; a.D = ix->d; a.nq = nqb; a.k = k;
                    a.rps = plan.rps; a.stages = plan.stages; a.chunks_per_cta = cpcs[i];
                    a.lists_total = lists_total; a.list_off = list_off; a.ord_base = ord;
                    // smem depends on the batch width only through `fixed`; keep the plan's size
                    cudaError_t e;
                    if (!plan.bulk) {
                        e = cudaFuncSetAttribute(ip_scan_topk_direct_kernel,
                                                 cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                 static_cast<int>(plan.smem));
                        if (e == cudaSuccess) {
                            ip_scan_topk_direct_kernel<<<grids[i], SCAN_CW * 32, plan.smem, st>>>(a); count_launch();
                            e = cudaGetLastError();
                        }
                    } else {
                        switch (plan.nv) {
                            case 1: e = launch_scan<1>(a, grids[i], plan.smem, st); break;
                            case 2: e = launch_scan<2>(a, grids[i], plan.smem, st); break;
                            case 3: e = launch_scan<3>(a, grids[i], plan.smem, st); break;
                            case 4: e = launch_scan<4>(a, grids[i], plan.smem, st); break;
                            case 5: e = launch_scan<5>(a, grids[i], plan.smem, st); break;
                            case 6: e = launch_scan<6>(a, grids[i], plan.smem, st); break;
                            case 7: e = launch_scan<7>(a, grids[i], plan.smem, st); break;
                            case 8: e = launch_scan<8>(a, grids[i], plan.smem, st); break;
                            default: e = launch_scan<0>(a, grids[i], plan.smem, st); break;
                        }
                    }
                    if (e != cudaSuccess)
                        return fail(DRAG_ERR_CUDA, std::string("ip_scan_topk launch: ") +
                                                       cudaGetErrorString(e));
                    ix->last_grid = grids[i];
                }
                list_off += grids[i];
                ord += static_cast<uint32_t>(sg.N);
            }
        }
    }
    if (ix->timing && !ix->segs.empty()) DRAG_CUDA(cudaEventRecord(ix->ev1, st));
    ix->last_stages = plan.stages; ix->last_rps = plan.rps; ix->last_nqb = plan.nqb;
    MergeArgs m;
    m.partial = ix->partial; m.lists = lists_total; m.k = k; m.kpad = next_pow2(k);
    m.seg_start = ix->seg_start; m.seg_base = ix->seg_base; m.nseg = nseg > 0 ? nseg : 1;
    m.D = D; m.I = I;
    topk_merge_keys_kernel<<<nq, MERGE_THREADS, 0, st>>>(m); count_launch();
    DRAG_CUDA(cudaGetLastError());
    return DRAG_OK;
}

int index_ensure_io(Index* ix, int nq, int k) {
    size_t qn = static_cast<size_t>(nq) * ix->d;
    if (qn > ix->qdev_cap) {
        if (ix->qdev) DRAG_CUDA(cudaFree(ix->qdev));
        DRAG_CUDA(cudaMalloc(&ix->qdev, qn * sizeof(float)));
        ix->qdev_cap = qn;
    }
    size_t on = static_cast<size_t>(nq) * k;
    if (on > ix->out_cap) {
        if (ix->Ddev) DRAG_CUDA(cudaFree(ix->Ddev));
        if (ix->Idev) DRAG_CUDA(cudaFree(ix->Idev));
        DRAG_CUDA(cudaMalloc(&ix->Ddev, on * sizeof(float)));
        DRAG_CUDA(cudaMalloc(&ix->Idev, on * sizeof(int64_t)));
        ix->out_cap = on;
    }
    return DRAG_OK;
}

int merge_pairs_device(const float* scores, const int64_t* ids, int nq, int lists, int k_in,
                       int k_out, float* D, int64_t* I, cudaStream_t st) {
    DRAG_REQUIRE(scores && ids && D && I, "topk_merge: null pointer");
    DRAG_REQUIRE(nq >= 0 && lists >= 1 && k_in >= 1 && k_out >= 1, "topk_merge: bad sizes");
    if (nq == 0) return DRAG_OK;
    const int T = lists * k_in;
    DRAG_REQUIRE(T <= PAIRS_MAX, "topk_merge: lists*k exceeds 8192");
    const int npad = next_pow2(T);
    const size_t smem = sizeof(uint32_t) * PAIRS_MAX + sizeof(int64_t) * PAIRS_MAX;
    DRAG_CUDA(cudaFuncSetAttribute(topk_merge_pairs_kernel,
                                   cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   static_cast<int>(smem)));
    topk_merge_pairs_kernel<<<nq, 1024, smem, st>>>(scores, ids, lists, k_in, k_out, npad, D, I); count_launch();
    DRAG_CUDA(cudaGetLastError());
    return DRAG_OK;
}

}  // namespace drag
