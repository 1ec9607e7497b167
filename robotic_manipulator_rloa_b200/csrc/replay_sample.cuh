// Replay sampler shared by rloa_replay_sample (replay.cu) and the fused learn kernel (naf_learn_cluster.cu): ReplayBuffer.sample
// (reference utils/replay_buffer.py:47-67, random.sample = without replacement) as a keyed bijection of the live window.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "rloa_b200.h"

namespace rloa {

__device__ __forceinline__ uint32_t mix32(uint32_t x) {
    x ^= x >> 16; x *= 0x85EBCA6Bu; x ^= x >> 13; x *= 0xC2B2AE35u; x ^= x >> 16;
    return x;
}
// keyed bijection of [0, 2^bits): 6-round Feistel over (hi: bits - bits/2, lo: bits/2) with alternating halves
__device__ __forceinline__ uint32_t feistel(uint32_t x, int bits, uint32_t k0, uint32_t k1) {
    const int lb = bits >> 1, hb = bits - lb;
    uint32_t lo = x & ((1u << lb) - 1u), hi = x >> lb;
#pragma unroll
    for (int r = 0; r < 6; r++) {
        if ((r & 1) == 0) hi = (hi ^ mix32(lo * 0x9E3779B9u + k0 + r)) & ((1u << hb) - 1u);
        else lo = (lo ^ mix32(hi * 0x7FEB352Du + k1 + r)) & ((1u << lb) - 1u);
    }
    return (hi << lb) | lo;
}

// ring slot of sample number `i` of the batch drawn with (seed, draw): distinct for distinct i < live
// `cur`: transitions ever appended (the ring's cursor, or the cursor as it will be once pending rows are committed)
__device__ __forceinline__ size_t replay_sample_slot_at(const rloa_replay& rb, long long cur, int i, unsigned long long seed,
                                                        unsigned long long draw) {
    const uint32_t live = (uint32_t)(cur < rb.capacity ? cur : rb.capacity);
    if (live == 0) return 0;
    int bits = 1;
    while ((1u << bits) < live) bits++;
    if (bits < 2) bits = 2;
    const uint32_t k0 = mix32((uint32_t)seed ^ mix32((uint32_t)draw)), k1 = mix32((uint32_t)(seed >> 32) + 0x68E31DA4u ^ (uint32_t)(draw >> 32) ^ k0);
    uint32_t x = (uint32_t)i % live;
    do {
        x = feistel(x, bits, k0, k1);
    } while (x >= live);                          // cycle walking keeps the map a bijection of [0, live)
    // x counts from the oldest live transition, like indexing the deque
    return (size_t)(((cur < rb.capacity ? 0 : cur) + x) % rb.capacity);
}
__device__ __forceinline__ size_t replay_sample_slot(const rloa_replay& rb, int i, unsigned long long seed, unsigned long long draw) {
    return replay_sample_slot_at(rb, *rb.cursor, i, seed, draw);
}

}  // namespace rloa
