// Library-wide state: version, thread-local error text, launch counter.
#include <atomic>
#include <map>
#include <mutex>
#include <stdarg.h>
#include <string>
#include <vector>

#include <string.h>
#include <algorithm>

#include "common.cuh"

namespace fr {
static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

// ---- optional per-kernel timing (CUDA events on the launch stream), off by default
static std::atomic<int> g_prof{0};
struct Rec { const char *name; cudaEvent_t a, b; };
static std::vector<Rec> g_recs;
static std::mutex g_mu;

bool profiling() { return g_prof.load(std::memory_order_relaxed) != 0; }

LaunchTimer::LaunchTimer(const char *name, cudaStream_t st) : name_(name), st_(st), on_(profiling()) {
    if (!on_) return;
    cudaEventCreate(&a_);
    cudaEventCreate(&b_);
    cudaEventRecord(a_, st_);
}
LaunchTimer::~LaunchTimer() {
    if (!on_) return;
    cudaEventRecord(b_, st_);
    std::lock_guard<std::mutex> lk(g_mu);
    g_recs.push_back({name_, a_, b_});
}
}  // namespace fr

extern "C" int fr_profile_enable(int on) {
    fr::g_prof.store(on ? 1 : 0);
    return FR_OK;
}

// Synchronises, folds the recorded launches per kernel name and writes
// "name,launches,total_us\n" lines into buf (truncated to cap); clears the records.
extern "C" int fr_profile_dump(char *buf, int64_t cap) {
    std::lock_guard<std::mutex> lk(fr::g_mu);
    std::map<std::string, std::pair<long, double>> agg;
    for (auto &r : fr::g_recs) {
        cudaEventSynchronize(r.b);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, r.a, r.b);
        auto &e = agg[r.name];
        e.first += 1;
        e.second += ms * 1e3;
        cudaEventDestroy(r.a);
        cudaEventDestroy(r.b);
    }
    fr::g_recs.clear();
    std::string out;
    for (auto &kv : agg) out += kv.first + "," + std::to_string(kv.second.first) + "," + std::to_string(kv.second.second) + "\n";
    if (buf && cap > 0) {
        const size_t n = std::min<size_t>(out.size(), (size_t)cap - 1);
        memcpy(buf, out.data(), n);
        buf[n] = 0;
    }
    return FR_OK;
}

extern "C" int fr_version(void) { return 100; }
extern "C" const char *fr_last_error(void) { return fr::g_err; }
extern "C" int64_t fr_launch_count(void) { return fr::g_launches.load(std::memory_order_relaxed); }
