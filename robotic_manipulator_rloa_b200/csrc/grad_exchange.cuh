// NVLink peer-memory gradient exchange: definitions shared by grad_exchange.cu (the 80-block publish / reduce + Adam kernels)
// and naf_learn_cluster.cu (the same exchange per gradient slice inside the fused learn kernel's tail).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "optim.cuh"

struct rloa_xchg;

namespace rloa {

constexpr int kXchgSlices = 8;             // flag sets per rank: slice 0 = the whole-gradient kernels, 0..7 = the cluster CTAs
constexpr int kXchgMaxWorld = 16;
constexpr int kXchgPushRegions = 8;        // ranks of the push-model exchange of the fused learn kernel (one NVSwitch box)
constexpr long long kXchgTimeoutCycles = 4000000000ll;      // ~2 s at 1.9 GHz

struct XchgPeers {
    const float* grad[kXchgMaxWorld];                        // mapped peer blocks (own block at index rank)
    unsigned long long* ready[kXchgMaxWorld];                // ready[r] = slot array in rank r's block (write slot `rank`)
    unsigned long long* done[kXchgMaxWorld];
    int world, rank;
};

__device__ __forceinline__ unsigned long long ld_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// peer data loads: system-scope (never served from a stale line), volatile so they stay below the flag wait, but
// without a memory clobber so a batch of them is issued back to back and the NVLink round trips overlap
__device__ __forceinline__ float ld_peer_f32(const float* p) {
    float v;
    asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ float4 ld_peer_f32x4(const float* p) {
    float4 v;
    asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}


// the mapped peer blocks of an exchange handle and its sticky time-out flag (device pointer)
void xchg_peers(const rloa_xchg* x, XchgPeers* out, int** status);
int xchg_connected_world(const rloa_xchg* x);

}  // namespace rloa
