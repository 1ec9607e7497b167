// librloa_b200: error reporting, version, launch counter, developer trace (include/rloa_b200.h).
#include "common.cuh"

#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

namespace rloa {

static thread_local char g_err[512] = "";
std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

// ---- developer trace: RLOA_TRACE=1 records one CUDA event after every launch on the stream given to
// rloa_trace_set_stream (eager launches on that one stream only: side-stream launches show up as short gaps);
// rloa_trace_dump prints the mean interval that ends at each launch site.  Off by default, one branch per launch. ----
struct TraceRec {
    cudaEvent_t ev;
    const char* file;
    int line;
};
static std::vector<TraceRec> g_trace;
static int g_trace_on = -1;
static cudaStream_t g_trace_stream = nullptr;

void trace_mark(const char* file, int line) {
    if (g_trace_on < 0) {
        const char* e = getenv("RLOA_TRACE");
        g_trace_on = (e != nullptr && e[0] == '1') ? 1 : 0;
    }
    if (!g_trace_on || g_trace.size() > 200000) return;
    TraceRec r{nullptr, file, line};
    if (cudaEventCreate(&r.ev) != cudaSuccess) return;
    if (cudaEventRecord(r.ev, g_trace_stream) != cudaSuccess) {
        cudaGetLastError();
        cudaEventDestroy(r.ev);
        return;
    }
    g_trace.push_back(r);
}

}  // namespace rloa

extern "C" const char* rloa_last_error(void) { return rloa::g_err; }
extern "C" int rloa_version(void) { return 100; }
extern "C" uint64_t rloa_launch_count(void) { return rloa::g_launches.load(std::memory_order_relaxed); }

// not part of the public header: developer tooling (tools/trace_step.py)
extern "C" void rloa_trace_set_stream(void* stream) { rloa::g_trace_stream = reinterpret_cast<cudaStream_t>(stream); }
extern "C" void rloa_trace_dump(int skip_first) {
    using namespace rloa;
    cudaDeviceSynchronize();
    struct Acc {
        double ms = 0;
        int n = 0;
    };
    std::map<std::pair<std::string, int>, Acc> acc;
    std::vector<std::pair<std::string, int>> order;
    for (size_t i = (size_t)skip_first + 1; i < g_trace.size(); i++) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, g_trace[i - 1].ev, g_trace[i].ev) != cudaSuccess) {
            cudaGetLastError();
            continue;
        }
        const char* f = strrchr(g_trace[i].file, '/');
        auto key = std::make_pair(std::string(f ? f + 1 : g_trace[i].file), g_trace[i].line);
        if (!acc.count(key)) order.push_back(key);
        acc[key].ms += ms;
        acc[key].n++;
    }
    double total = 0;
    for (auto& k : order) {
        printf("%-20s:%5d  n=%5d  avg %7.2f us\n", k.first.c_str(), k.second, acc[k].n, acc[k].ms / acc[k].n * 1e3);
        total += acc[k].ms / acc[k].n * 1e3;
    }
    printf("sum of the averages: %.1f us\n", total);
    for (auto& r : g_trace) cudaEventDestroy(r.ev);
    g_trace.clear();
}
