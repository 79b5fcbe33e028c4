#include <vector>

#include "common.cuh"
#include "prof.cuh"

namespace drag {

struct ProfRec {
    int cls;
    double work;
    cudaEvent_t e0, e1;
};
static bool g_prof_on = false;
static std::vector<ProfRec> g_recs;
static std::vector<cudaEvent_t> g_pool;

bool prof_enabled() { return g_prof_on; }

static cudaEvent_t get_event() {
    if (!g_pool.empty()) {
        cudaEvent_t e = g_pool.back();
        g_pool.pop_back();
        return e;
    }
    cudaEvent_t e;
    cudaEventCreate(&e);
    return e;
}

int prof_begin(int cls, double work, cudaStream_t st) {
    if (!g_prof_on) return -1;
    ProfRec r;
    r.cls = cls;
    r.work = work;
    r.e0 = get_event();
    r.e1 = get_event();
    cudaEventRecord(r.e0, st);
    g_recs.push_back(r);
    return static_cast<int>(g_recs.size()) - 1;
}

void prof_end(int slot, cudaStream_t st) {
    if (slot < 0) return;
    cudaEventRecord(g_recs[slot].e1, st);
}

int prof_enable(int on) {
    g_prof_on = on != 0;
    return DRAG_OK;
}

// Synchronises, sums elapsed time / work / launches per class, and clears the records.
int prof_collect(double* ms, double* work, int* count, int n_classes) {
    for (int c = 0; c < n_classes; ++c) {
        ms[c] = 0;
        work[c] = 0;
        count[c] = 0;
    }
    for (ProfRec& r : g_recs) {
        DRAG_CUDA(cudaEventSynchronize(r.e1));
        float t = 0.f;
        DRAG_CUDA(cudaEventElapsedTime(&t, r.e0, r.e1));
        if (r.cls < n_classes) {
            ms[r.cls] += t;
            work[r.cls] += r.work;
            count[r.cls] += 1;
        }
        g_pool.push_back(r.e0);
        g_pool.push_back(r.e1);
    }
    g_recs.clear();
    return DRAG_OK;
}

}  // namespace drag
