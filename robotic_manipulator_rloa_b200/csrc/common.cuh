// Shared host-side plumbing of librloa_b200.so: error reporting and the launch counter.
#pragma once
#include <cuda_runtime.h>
#include <atomic>
#include <cstdarg>
#include <cstdint>
#include <cstdio>

#include "rloa_b200.h"

namespace rloa {

void set_error(const char* fmt, ...);
extern std::atomic<uint64_t> g_launches;

inline int fail(int code, const char* what) {
    set_error("%s", what);
    return code;
}

#define RLOA_CUDA(call)                                                                            \
    do {                                                                                           \
        cudaError_t err__ = (call);                                                                \
        if (err__ != cudaSuccess) {                                                                \
            ::rloa::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(err__), __FILE__, \
                              __LINE__);                                                           \
            return RLOA_ERR_CUDA;                                                                  \
        }                                                                                          \
    } while (0)

// developer trace (RLOA_TRACE=1, eager launches only): one CUDA event after every launch, intervals printed by
// rloa_trace_dump() — warm per-kernel times inside the real loop, which ncu (cold, serialised) cannot give
void trace_mark(const char* file, int line);

// after a <<<>>> launch: count it and surface launch-configuration errors
#define RLOA_LAUNCHED()                                                                            \
    do {                                                                                           \
        ::rloa::g_launches.fetch_add(1, std::memory_order_relaxed);                                \
        ::rloa::trace_mark(__FILE__, __LINE__);                                                    \
        cudaError_t err__ = cudaPeekAtLastError();                                                 \
        if (err__ != cudaSuccess) {                                                                \
            ::rloa::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(err__),       \
                              __FILE__, __LINE__);                                                 \
            return RLOA_ERR_CUDA;                                                                  \
        }                                                                                          \
    } while (0)

#define RLOA_REQUIRE(cond, msg)                          \
    do {                                                 \
        if (!(cond)) return ::rloa::fail(RLOA_ERR_INVALID, msg); \
    } while (0)

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

}  // namespace rloa
