// librloa_b200: error reporting, version, launch counter (include/rloa_b200.h).
#include "common.cuh"

namespace rloa {

static thread_local char g_err[512] = "";
std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

}  // namespace rloa

extern "C" const char* rloa_last_error(void) { return rloa::g_err; }
extern "C" int rloa_version(void) { return 100; }
extern "C" uint64_t rloa_launch_count(void) { return rloa::g_launches.load(std::memory_order_relaxed); }
