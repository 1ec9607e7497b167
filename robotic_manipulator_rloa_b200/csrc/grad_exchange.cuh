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


// ---- tagged words of the fused learn kernel's exchange (naf_learn_cluster.cu) ----------------------------------------
// Every fp32 value travels as one 64-bit word {value bits | update number << 32} written with a single 8-byte store: the
// receiver polls the word itself until the tag is the current update's - no flag, no fence, no round trip behind the
// data (what NCCL's LL protocol does).  Layout of a block, in floats from its start:
//   [0, n)                       the rank's gradient, pulled by the peers (whole-gradient kernels of grad_exchange.cu)
//   [n, n + 16 n)                round 1: kXchgPushRegions regions of n words; region s word e = element e as summed by rank s
//   [17 n, 19 n)                 round 2: n words, the reduced elements written by their owner ranks
__host__ __device__ constexpr size_t xchg_ll1_offset(int n) { return (size_t)n; }
__host__ __device__ constexpr size_t xchg_ll2_offset(int n) { return (size_t)n * (1 + 2 * kXchgPushRegions); }
__host__ __device__ constexpr size_t xchg_block_floats(int n) { return (size_t)n * (3 + 2 * kXchgPushRegions); }

__device__ __forceinline__ unsigned long long ll_word(float v, unsigned tag) {
    return ((unsigned long long)tag << 32) | (unsigned long long)__float_as_uint(v);
}
__device__ __forceinline__ void st_ll(unsigned long long* p, float v, unsigned tag) {
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(ll_word(v, tag)) : "memory");
}
__device__ __forceinline__ void st_ll4(unsigned long long* p, float4 v, unsigned tag) {      // p 32-byte aligned
    asm volatile("st.relaxed.sys.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(ll_word(v.x, tag)), "l"(ll_word(v.y, tag)) : "memory");
    asm volatile("st.relaxed.sys.global.v2.u64 [%0], {%1, %2};" ::"l"(p + 2), "l"(ll_word(v.z, tag)), "l"(ll_word(v.w, tag)) : "memory");
}
// one attempt: true when the word(s) carry the tag
__device__ __forceinline__ bool ld_ll(const unsigned long long* p, unsigned tag, float& v) {
    unsigned long long w;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(w) : "l"(p) : "memory");
    v = __uint_as_float((unsigned)w);
    return (unsigned)(w >> 32) == tag;
}
__device__ __forceinline__ bool ld_ll4(const unsigned long long* p, unsigned tag, float4& v) {
    unsigned long long w0, w1, w2, w3;
    asm volatile("ld.relaxed.sys.global.v2.u64 {%0, %1}, [%2];" : "=l"(w0), "=l"(w1) : "l"(p) : "memory");
    asm volatile("ld.relaxed.sys.global.v2.u64 {%0, %1}, [%2];" : "=l"(w2), "=l"(w3) : "l"(p + 2) : "memory");
    v = make_float4(__uint_as_float((unsigned)w0), __uint_as_float((unsigned)w1), __uint_as_float((unsigned)w2), __uint_as_float((unsigned)w3));
    return (unsigned)(w0 >> 32) == tag && (unsigned)(w1 >> 32) == tag && (unsigned)(w2 >> 32) == tag && (unsigned)(w3 >> 32) == tag;
}

// the mapped peer blocks of an exchange handle and its sticky time-out flag (device pointer)
void xchg_peers(const rloa_xchg* x, XchgPeers* out, int** status);
int xchg_connected_world(const rloa_xchg* x);

}  // namespace rloa
