// Optional per-kernel-family timing with CUDA events on the launching stream (bench.py "roofline").
// Disabled by default: when off, prof_begin/prof_end are a single predictable branch.
#include <vector>
#include "gvl_internal.h"
#include "../../include/gvl.h"

namespace gvl {

namespace {
struct Rec {
    cudaEvent_t a, b;
    double work;
};
bool g_on = false;
std::vector<Rec> g_recs[GVL_PROF_KINDS];
std::vector<cudaEvent_t> g_pool;

cudaEvent_t get_event() {
    if (!g_pool.empty()) {
        cudaEvent_t e = g_pool.back();
        g_pool.pop_back();
        return e;
    }
    cudaEvent_t e;
    cudaEventCreate(&e);
    return e;
}
}  // namespace

bool prof_enabled() { return g_on; }

void prof_begin(int kind, double work, cudaStream_t s) {
    if (!g_on) return;
    Rec r;
    r.a = get_event();
    r.b = get_event();
    r.work = work;
    cudaEventRecord(r.a, s);
    g_recs[kind].push_back(r);
}

void prof_end(int kind, cudaStream_t s) {
    if (!g_on || g_recs[kind].empty()) return;
    cudaEventRecord(g_recs[kind].back().b, s);
}

}  // namespace gvl

extern "C" {

int gvl_profile_enable(int on) {
    gvl::g_on = on != 0;
    return GVL_OK;
}

// Synchronises the device, sums (elapsed ms, work units, launches) of one kernel family and clears it.
// work units: FLOPs for GVL_PROF_GEMM / GVL_PROF_ATTN, bytes for GVL_PROF_GEMV / GVL_PROF_DECODE_ATTN.
int gvl_profile_collect(int kind, double* total_ms, double* total_work, long long* launches) {
    if (kind < 0 || kind >= GVL_PROF_KINDS || !total_ms || !total_work || !launches) return GVL_ERR_ARG;
    if (cudaDeviceSynchronize() != cudaSuccess) return GVL_ERR_CUDA;
    double ms = 0, work = 0;
    for (auto& r : gvl::g_recs[kind]) {
        float t = 0;
        if (cudaEventElapsedTime(&t, r.a, r.b) == cudaSuccess) ms += t;
        work += r.work;
        gvl::g_pool.push_back(r.a);
        gvl::g_pool.push_back(r.b);
    }
    *total_ms = ms;
    *total_work = work;
    *launches = (long long)gvl::g_recs[kind].size();
    gvl::g_recs[kind].clear();
    return GVL_OK;
}

}  // extern "C"
