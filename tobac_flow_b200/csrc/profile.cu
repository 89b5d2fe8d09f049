// Launch accounting for bench.py: per kernel class, the number of launches, the algorithmic bytes they moved and
// (when enabled) their device time measured with CUDA events recorded on the launching stream.
#include <stdlib.h>

#include <mutex>
#include <vector>

#include "tf_common.cuh"

namespace tf {

namespace {
struct Rec { cudaEvent_t a, b; int klass; };
std::mutex g_mu;
bool g_enabled = false;
std::vector<Rec> g_recs;
std::vector<cudaEvent_t> g_pool;
double g_bytes[KC_COUNT] = {0};
long long g_launches[KC_COUNT] = {0};
double g_ms[KC_COUNT] = {0};
const size_t kMaxRecs = 1 << 16;

cudaEvent_t get_event() {
    if (!g_pool.empty()) { cudaEvent_t e = g_pool.back(); g_pool.pop_back(); return e; }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
}

void drain_locked() {
    // TF_PROFILE_DUMP=file: one line per timed launch group (class, start relative to the first group, duration, ms)
    static const char* dump_path = getenv("TF_PROFILE_DUMP");
    FILE* dump = (dump_path && !g_recs.empty()) ? fopen(dump_path, "a") : nullptr;
    for (Rec& r : g_recs) {
        float ms = 0.f;
        if (cudaEventSynchronize(r.b) == cudaSuccess && cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess)
            g_ms[r.klass] += ms;
        if (dump) {
            float t0 = 0.f;
            cudaEventElapsedTime(&t0, g_recs.front().a, r.a);
            fprintf(dump, "%d %.4f %.4f\n", r.klass, t0, ms);
        }
        g_pool.push_back(r.a);
        g_pool.push_back(r.b);
    }
    if (dump) fclose(dump);
    g_recs.clear();
}
}  // namespace

LaunchTimer::LaunchTimer(int klass, double bytes, cudaStream_t s, int n_launches) : klass_(klass), s_(s), b_(nullptr) {
    std::lock_guard<std::mutex> lk(g_mu);
    g_launches[klass] += n_launches;
    g_bytes[klass] += bytes;
    if (g_enabled && g_recs.size() < kMaxRecs) {
        cudaEvent_t a = get_event();
        b_ = get_event();
        if (a && b_) {
            cudaEventRecord(a, s);
            g_recs.push_back(Rec{a, b_, klass});
        } else {
            b_ = nullptr;
        }
    }
}

LaunchTimer::~LaunchTimer() {
    if (b_) cudaEventRecord(b_, s_);
}

}  // namespace tf

using namespace tf;

extern "C" int tf_profile_enable(int on) {
    std::lock_guard<std::mutex> lk(g_mu);
    drain_locked();
    g_enabled = on != 0;
    return TF_OK;
}

extern "C" int tf_profile_reset(void) {
    std::lock_guard<std::mutex> lk(g_mu);
    drain_locked();
    for (int i = 0; i < KC_COUNT; ++i) { g_bytes[i] = 0; g_launches[i] = 0; g_ms[i] = 0; }
    return TF_OK;
}

extern "C" int tf_profile_read(int klass, double* total_ms, double* total_bytes, long long* launches) {
    if (klass < 0 || klass >= KC_COUNT) { set_error("tf_profile_read: unknown kernel class %d", klass); return TF_ERR_INVALID_ARGUMENT; }
    std::lock_guard<std::mutex> lk(g_mu);
    drain_locked();
    if (total_ms) *total_ms = g_ms[klass];
    if (total_bytes) *total_bytes = g_bytes[klass];
    if (launches) *launches = g_launches[klass];
    return TF_OK;
}
