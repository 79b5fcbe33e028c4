// Sharded exact top-k: the exchange step of SURVEY 8e ("all-gather of per-shard top-k -> merge") as ONE kernel over
// NVLink peer memory instead of a packed NCCL all-gather followed by a merge kernel. The per-query result is nq * k * 16
// bytes per rank (1.6 KB at nq = 1, k = 100): the NCCL path is pure latency - a collective launch, its proxy / kernel
// hand-shake and the merge launch cost more than the 39 us shard scan they follow at N = 1 M / 8 GPUs.
//
// Every rank owns one symmetric buffer (torch.distributed._symmetric_memory: the same allocation mapped into every peer's
// address space):   slots [2 parities][world][nq_cap][k_cap] x {u32 orderable score, u32 pad, i64 id}   + flags
//                   [2][world][nq_cap] u32.
// CTA q of rank r:  (1) PUSH  its local (score, id) list for query q into slot [parity][r][q] of EVERY rank with plain P2P
//                       stores (8 x 1.6 KB over NVSwitch), __threadfence_system, then a release store of `epoch` to
//                       flag [parity][r][q] on every rank;
//                   (2) WAIT  until its own flags [parity][*][q] have reached `epoch` (acquire loads, local memory);
//                   (3) MERGE the world lists from its own slots with the (score desc, id asc) bitonic network of the
//                       NCCL path's merge kernel -> D, I. Bit-identical to a single index over the whole corpus.
// Parity = epoch & 1 double-buffers the slots: a peer may already be pushing search e+1 while this rank still merges
// search e; it cannot reach e+2 before this rank has pushed e+1, i.e. finished reading e. All ranks must call the
// exchange the same number of times (it is a collective); kernels on different GPUs only ever wait for pushes, never
// for each other's completion, so there is no co-scheduling requirement.
#include <cfloat>

#include "common.cuh"
#include "index.cuh"

namespace drag {

constexpr int XCH_MAX_WORLD = 16;
constexpr int XCH_THREADS = 1024;
constexpr int XCH_PAIRS_MAX = 8192;

struct XchArgs {
    uint8_t* peer[XCH_MAX_WORLD];      // base of every rank's symmetric buffer, as mapped on THIS device
    const float* D_loc;                // [nq][k]
    const int64_t* I_loc;              // [nq][k]
    float* D;                          // [nq][k_out]
    int64_t* I;
    int world, rank, nq_cap, k_cap, k, k_out, npad;
    uint32_t epoch;
};

__device__ __forceinline__ uint32_t xch_order_f32(float f) {
    if (f != f) f = -INFINITY;
    f += 0.0f;
    const uint32_t b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float xch_unorder_f32(uint32_t u) {
    const uint32_t b = (u & 0x80000000u) ? (u & 0x7FFFFFFFu) : ~u;
    return __uint_as_float(b);
}
__device__ __forceinline__ void st_release_sys_u32(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys_u32(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__host__ __device__ inline size_t xch_slot_bytes(int world, int nq_cap, int k_cap) {
    return static_cast<size_t>(2) * world * nq_cap * k_cap * 16;
}

__global__ void __launch_bounds__(XCH_THREADS, 1) topk_exchange_merge_kernel(XchArgs a) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint32_t* sk = reinterpret_cast<uint32_t*>(smem);                                   // [npad] orderable score
    int64_t* si = reinterpret_cast<int64_t*>(smem + sizeof(uint32_t) * XCH_PAIRS_MAX);  // [npad]
    const int q = blockIdx.x, tid = threadIdx.x;
    const int par = a.epoch & 1;
    const size_t slot_bytes = xch_slot_bytes(a.world, a.nq_cap, a.k_cap);
    const size_t my_slot = ((static_cast<size_t>(par) * a.world + a.rank) * a.nq_cap + q) * a.k_cap * 16;
    const size_t my_flag = slot_bytes + ((static_cast<size_t>(par) * a.world + a.rank) * a.nq_cap + q) * 4;

    // (1) push this rank's list for query q to every rank (own buffer included)
    for (int i = tid; i < a.world * a.k; i += XCH_THREADS) {
        const int p = i / a.k, j = i - p * a.k;
        const int64_t id = a.I_loc[static_cast<size_t>(q) * a.k + j];
        const uint32_t key = (id < 0) ? 0u : xch_order_f32(a.D_loc[static_cast<size_t>(q) * a.k + j]);
        uint4 v;
        v.x = key;
        v.y = 0u;
        v.z = static_cast<uint32_t>(static_cast<uint64_t>(id));
        v.w = static_cast<uint32_t>(static_cast<uint64_t>(id) >> 32);
        *reinterpret_cast<uint4*>(a.peer[p] + my_slot + static_cast<size_t>(j) * 16) = v;
    }
    __syncthreads();
    if (tid < a.world) {
        __threadfence_system();
        st_release_sys_u32(reinterpret_cast<uint32_t*>(a.peer[tid] + my_flag), a.epoch);
    }
    // (2) wait for every rank's push of this query (flags are monotonic: a peer may already be one search ahead)
    if (tid < a.world) {
        const uint32_t* f = reinterpret_cast<const uint32_t*>(
            a.peer[a.rank] + slot_bytes + ((static_cast<size_t>(par) * a.world + tid) * a.nq_cap + q) * 4);
        while (static_cast<int32_t>(ld_acquire_sys_u32(f) - a.epoch) < 0) __nanosleep(64);
    }
    __syncthreads();
    // (3) merge world x k pairs from this rank's own slots
    const int T = a.world * a.k;
    for (int i = tid; i < a.npad; i += XCH_THREADS) {
        if (i < T) {
            const int p = i / a.k, j = i - p * a.k;
            const uint8_t* src = a.peer[a.rank] + ((static_cast<size_t>(par) * a.world + p) * a.nq_cap + q) * a.k_cap * 16 +
                                 static_cast<size_t>(j) * 16;
            const uint4 v = __ldcg(reinterpret_cast<const uint4*>(src));   // L2 is the coherence point for peer writes; skip L1
            const int64_t id = static_cast<int64_t>((static_cast<uint64_t>(v.w) << 32) | v.z);
            si[i] = id;
            sk[i] = (id < 0) ? 0u : v.x;
        } else {
            si[i] = -1;
            sk[i] = 0u;
        }
    }
    __syncthreads();
    auto before = [](uint32_t ka, int64_t ia, uint32_t kb, int64_t ib) {
        const bool va = ia >= 0, vb = ib >= 0;       // valid first, score desc, id asc
        if (va != vb) return va;
        if (ka != kb) return ka > kb;
        return ia < ib;
    };
    for (int k2 = 2; k2 <= a.npad; k2 <<= 1) {
        for (int j = k2 >> 1; j > 0; j >>= 1) {
            for (int t = tid; t < (a.npad >> 1); t += XCH_THREADS) {
                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                const int l = i | j;
                const bool fwd = ((i & k2) == 0);
                const uint32_t kx = sk[i], ky = sk[l];
                const int64_t ix = si[i], iy = si[l];
                const bool swap = fwd ? before(ky, iy, kx, ix) : before(kx, ix, ky, iy);
                if (swap) {
                    sk[i] = ky; sk[l] = kx;
                    si[i] = iy; si[l] = ix;
                }
            }
            __syncthreads();
        }
    }
    for (int j = tid; j < a.k_out; j += XCH_THREADS) {
        const bool valid = (j < a.npad) && si[j] >= 0;
        a.D[static_cast<size_t>(q) * a.k_out + j] = valid ? xch_unorder_f32(sk[j]) : -FLT_MAX;
        a.I[static_cast<size_t>(q) * a.k_out + j] = valid ? si[j] : -1;
    }
}

size_t topk_exchange_buffer_bytes(int world, int nq_cap, int k_cap) {
    return xch_slot_bytes(world, nq_cap, k_cap) + static_cast<size_t>(2) * world * nq_cap * 4;
}

int topk_exchange_merge(const float* D_loc, const int64_t* I_loc, int nq, int k, void* const* peer_bufs, int world, int rank,
                        int nq_cap, int k_cap, uint32_t epoch, float* D, int64_t* I, cudaStream_t st) {
    DRAG_REQUIRE(D_loc && I_loc && peer_bufs && D && I, "topk_exchange: null pointer");
    DRAG_REQUIRE(world >= 1 && world <= XCH_MAX_WORLD && rank >= 0 && rank < world, "topk_exchange: bad world / rank");
    DRAG_REQUIRE(nq >= 0 && nq <= nq_cap && k >= 1 && k <= k_cap, "topk_exchange: nq / k exceed the buffer capacity");
    DRAG_REQUIRE(static_cast<long long>(world) * k <= XCH_PAIRS_MAX, "topk_exchange: world * k exceeds 8192");
    DRAG_REQUIRE(epoch != 0, "topk_exchange: epoch starts at 1 (flags are zero-initialised)");
    if (nq == 0) return DRAG_OK;
    XchArgs a{};
    for (int p = 0; p < world; ++p) {
        DRAG_REQUIRE(peer_bufs[p], "topk_exchange: null peer buffer");
        a.peer[p] = static_cast<uint8_t*>(peer_bufs[p]);
    }
    a.D_loc = D_loc; a.I_loc = I_loc; a.D = D; a.I = I;
    a.world = world; a.rank = rank; a.nq_cap = nq_cap; a.k_cap = k_cap; a.k = k; a.k_out = k;
    int npad = 1;
    while (npad < world * k) npad <<= 1;
    a.npad = npad;
    a.epoch = epoch;
    const size_t smem = sizeof(uint32_t) * XCH_PAIRS_MAX + sizeof(int64_t) * XCH_PAIRS_MAX;
    static bool attr_set = false;
    if (!attr_set) {
        DRAG_CUDA(cudaFuncSetAttribute(topk_exchange_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(smem)));
        attr_set = true;
    }
    topk_exchange_merge_kernel<<<nq, XCH_THREADS, smem, st>>>(a); count_launch();
    DRAG_CUDA(cudaGetLastError());
    return DRAG_OK;
}

}  // namespace drag
