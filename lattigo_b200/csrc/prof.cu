// prof.cu -- launch accounting and optional CUDA-event profiling of kernel classes.
// bench.py uses it for `gpu_launches` and for the live roofline figure (algorithmic bytes / event time of the
// dominant kernel class, measured on the launching stream). Disabled by default: zero events recorded.
#include <atomic>
#include <mutex>
#include <vector>
#include "../../include/lattigo_b200.h"
#include "engine.h"

namespace lgpu {

static std::atomic<unsigned long long> g_launches{0};
static std::atomic<int> g_prof_on{0};
struct ProfRec { int k; cudaEvent_t a, b; double bytes; int kernels; };
static std::mutex g_mu;
static std::vector<ProfRec> g_recs;

bool profiling_on() { return g_prof_on.load(std::memory_order_relaxed) != 0; }
void count_launch(int n) { g_launches.fetch_add((unsigned long long)n, std::memory_order_relaxed); }

ProfScope::ProfScope(int k, cudaStream_t st, double bytes, int kernels) : k_(k), st_(st), on_(false), bytes_(bytes), kernels_(kernels) {
    count_launch(kernels);
    if (g_prof_on.load(std::memory_order_relaxed)) {
        if (cudaEventCreate(&a_) == cudaSuccess && cudaEventCreate(&b_) == cudaSuccess) {
            cudaEventRecord(a_, st_);
            on_ = true;
        }
    }
}
ProfScope::~ProfScope() {
    if (!on_) return;
    cudaEventRecord(b_, st_);
    std::lock_guard<std::mutex> lk(g_mu);
    g_recs.push_back(ProfRec{k_, a_, b_, bytes_, kernels_});
}

}  // namespace lgpu

using namespace lgpu;

extern "C" {

unsigned long long lgpu_launch_count(void) { return g_launches.load(); }

int lgpu_profile_enable(int on) {
    g_prof_on.store(on ? 1 : 0);
    return 0;
}

// Sums and clears the recorded scopes. Arrays have LGPU_KCLASS_COUNT entries.
int lgpu_profile_read(double* ms, double* bytes, unsigned long long* scopes, unsigned long long* kernels) {
    std::vector<ProfRec> recs;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        recs.swap(g_recs);
    }
    for (int i = 0; i < LGPU_KCLASS_COUNT; i++) { ms[i] = 0; bytes[i] = 0; scopes[i] = 0; kernels[i] = 0; }
    for (auto& r : recs) {
        cudaEventSynchronize(r.b);
        float t = 0;
        cudaEventElapsedTime(&t, r.a, r.b);
        if (r.k >= 0 && r.k < LGPU_KCLASS_COUNT) { ms[r.k] += t; bytes[r.k] += r.bytes; scopes[r.k]++; kernels[r.k] += r.kernels; }
        cudaEventDestroy(r.a);
        cudaEventDestroy(r.b);
    }
    return 0;
}

}  // extern "C"
