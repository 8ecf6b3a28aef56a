// Library-wide state: version, thread-local error text, launch counter.
#include <atomic>
#include <stdarg.h>

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
}  // namespace fr

extern "C" int fr_version(void) { return 100; }
extern "C" const char *fr_last_error(void) { return fr::g_err; }
extern "C" int64_t fr_launch_count(void) { return fr::g_launches.load(std::memory_order_relaxed); }
