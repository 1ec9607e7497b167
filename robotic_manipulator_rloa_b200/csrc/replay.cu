// librloa_b200: replay ring in HBM (include/rloa_b200.h).
//
// Replaces utils/replay_buffer.py of the reference: `deque(maxlen)` of host namedtuples becomes a
// struct-of-arrays ring resident in HBM (append = coalesced row copies at (cursor + rank) % capacity),
// `random.sample(memory, k)` becomes a keyed Feistel permutation of the live window with cycle walking
// (k distinct slots, no rejection table, no host round trip), and the np.stack / .to(device) copies
// (replay_buffer.py:57-65) become one gather kernel.
#include "common.cuh"
#include "replay_sample.cuh"

namespace rloa {

constexpr int kAppendRows = 64;        // transition rows per block
constexpr int kAppendThreads = 256;

__global__ void __launch_bounds__(kAppendRows) replay_count_kernel(const uint8_t* __restrict__ valid, int n,
                                                                   int* __restrict__ block_counts) {
    const int i = blockIdx.x * kAppendRows + threadIdx.x;
    const int v = (i < n && valid[i] != 0) ? 1 : 0;
    const int c = __syncthreads_count(v);
    if (threadIdx.x == 0) block_counts[blockIdx.x] = c;
}

// Block = 256 transition rows.  rank = (valid rows of earlier blocks) + exclusive scan of the valid flags; the
// copies are flat loops over the block's [256][S] / [256][A] sub-matrices (coalesced reads, and coalesced writes
// because ranks of neighbouring rows are consecutive).  The last block to finish bumps the cursor: every block
// reads the cursor before it takes its ticket, so the bump cannot overtake a reader.
template <int ROWS>
__global__ void __launch_bounds__(kAppendThreads)
replay_append_kernel(rloa_replay rb, int n, const float* __restrict__ states, const float* __restrict__ actions,
                     const float* __restrict__ rewards, const float* __restrict__ next_states,
                     const uint8_t* __restrict__ dones, const uint8_t* __restrict__ valid,
                     const int* __restrict__ block_counts, unsigned* __restrict__ ticket, int commit) {
    __shared__ int s_slot[ROWS];
    __shared__ int s_red[2][kAppendThreads / 32];
    __shared__ int s_wsum[ROWS / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int row0 = blockIdx.x * ROWS;
    const long long cur = *reinterpret_cast<const volatile long long*>(rb.cursor);
    // valid rows of the earlier blocks / of all blocks: from the per-block counts of replay_count_kernel, or — for
    // the batch sizes of a training loop (n <= 16384) — counted here straight from the mask, which saves the launch
    int before = 0, total = 0;
    if (valid != nullptr) {
        if (block_counts != nullptr) {
            for (int b = tid; b < (int)gridDim.x; b += kAppendThreads) {
                const int cnt = block_counts[b];
                total += cnt;
                if (b < (int)blockIdx.x) before += cnt;
            }
        } else {
            for (int i = tid; i < n; i += kAppendThreads) {
                const int vv = valid[i] != 0 ? 1 : 0;
                total += vv;
                if (i < row0) before += vv;
            }
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            before += __shfl_xor_sync(0xffffffffu, before, off);
            total += __shfl_xor_sync(0xffffffffu, total, off);
        }
        if (lane == 0) { s_red[0][warp] = before; s_red[1][warp] = total; }
        __syncthreads();
        before = 0; total = 0;
        for (int w = 0; w < kAppendThreads / 32; w++) { before += s_red[0][w]; total += s_red[1][w]; }
    } else {
        before = row0;
        total = n;
    }
    // ranks of this block's rows: the first ROWS threads own one row each
    bool v = false;
    int in_warp = 0;
    if (tid < ROWS) {
        const int i = row0 + tid;
        v = i < n && (valid == nullptr || valid[i] != 0);
        const unsigned bal = __ballot_sync(0xffffffffu, v);
        in_warp = __popc(bal & ((1u << lane) - 1u));
        if (lane == 0) s_wsum[warp] = __popc(bal);
    }
    __syncthreads();
    if (tid < ROWS) {
        int wbase = 0;
        for (int w = 0; w < warp; w++) wbase += s_wsum[w];
        const int rank = before + wbase + in_warp;
        const int slot = v ? (int)((cur + rank) % rb.capacity) : -1;
        s_slot[tid] = slot;
        if (v) {
            rb.rewards[slot] = rewards[row0 + tid];
            rb.dones[slot] = dones != nullptr ? (float)dones[row0 + tid] : 0.f;
        }
    }
    __syncthreads();
    const int S = rb.state_size, A = rb.action_size;
    const int rows = min(ROWS, n - row0);
    // flat copies of the block's [rows][S] / [rows][A] sub-matrices, four independent loads in flight per thread
    for (int e0 = tid; e0 < rows * S; e0 += 4 * kAppendThreads) {
        float a[4], b[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int e = e0 + u * kAppendThreads;
            a[u] = e < rows * S ? states[(size_t)row0 * S + e] : 0.f;
            b[u] = e < rows * S ? next_states[(size_t)row0 * S + e] : 0.f;
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int e = e0 + u * kAppendThreads;
            if (e < rows * S) {
                const int r = e / S, k = e - r * S;
                const int slot = s_slot[r];
                if (slot >= 0) {
                    rb.states[(size_t)slot * S + k] = a[u];
                    rb.next_states[(size_t)slot * S + k] = b[u];
                }
            }
        }
    }
    for (int e = tid; e < rows * A; e += kAppendThreads) {
        const int r = e / A, k = e - r * A;
        const int slot = s_slot[r];
        if (slot >= 0) rb.actions[(size_t)slot * A + k] = actions[(size_t)row0 * A + e];
    }
    if (!commit) return;            // rows only: rloa_replay_commit moves the cursor later
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        const unsigned t = atomicAdd(ticket, 1u);
        if (t == gridDim.x - 1u) {
            *ticket = 0u;
            *reinterpret_cast<long long*>(rb.cursor) = cur + total;
        }
    }
}

// the second half of a deferred append: cursor += number of valid rows (one block)
__global__ void __launch_bounds__(256) replay_commit_kernel(rloa_replay rb, int n, const uint8_t* __restrict__ valid) {
    __shared__ int s_red[8];
    int cnt = 0;
    if (valid != nullptr) {
        for (int i = threadIdx.x; i < n; i += 256) cnt += valid[i] != 0 ? 1 : 0;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, off);
        if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = cnt;
        __syncthreads();
        cnt = 0;
        for (int w = 0; w < 8; w++) cnt += s_red[w];
    } else {
        cnt = n;
    }
    if (threadIdx.x == 0) *reinterpret_cast<long long*>(rb.cursor) += cnt;
}

__global__ void __launch_bounds__(256)
replay_sample_kernel(rloa_replay rb, int batch, unsigned long long seed, unsigned long long draw0,
                     const unsigned long long* __restrict__ draw_offset, float* __restrict__ states, float* __restrict__ actions, float* __restrict__ rewards,
                     float* __restrict__ next_states, float* __restrict__ dones, int* __restrict__ indices) {
    const int warp = (blockIdx.x * 256 + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= batch) return;
    const unsigned long long draw = draw0 + (draw_offset != nullptr ? *draw_offset : 0ull);
    if (*rb.cursor <= 0) return;
    const size_t slot = replay_sample_slot(rb, warp, seed, draw);
    const int S = rb.state_size, A = rb.action_size;
    for (int k = lane; k < S; k += 32) {
        states[(size_t)warp * S + k] = rb.states[slot * S + k];
        next_states[(size_t)warp * S + k] = rb.next_states[slot * S + k];
    }
    for (int k = lane; k < A; k += 32) actions[(size_t)warp * A + k] = rb.actions[slot * A + k];
    if (lane == 0) {
        rewards[warp] = rb.rewards[slot];
        if (dones != nullptr) dones[warp] = rb.dones[slot];
        if (indices != nullptr) indices[warp] = (int)slot;
    }
}

}  // namespace rloa

using namespace rloa;

static int check_rb(const rloa_replay* rb, const char* who) {
    if (rb == nullptr || rb->capacity < 1 || rb->state_size < 1 || rb->action_size < 1 || !rb->states || !rb->actions ||
        !rb->rewards || !rb->next_states || !rb->dones || !rb->cursor) {
        set_error("%s: incomplete replay descriptor", who);
        return RLOA_ERR_INVALID;
    }
    return RLOA_OK;
}

static int replay_append_impl(const rloa_replay* rb, int32_t n, const float* states, const float* actions,
                              const float* rewards, const float* next_states, const uint8_t* dones,
                              const uint8_t* valid, void* stream, int commit) {
    int rc = check_rb(rb, "rloa_replay_append");
    if (rc != RLOA_OK) return rc;
    RLOA_REQUIRE(n >= 1 && states && actions && rewards && next_states, "rloa_replay_append: null argument");
    RLOA_REQUIRE(n <= rb->capacity, "rloa_replay_append: more transitions than the ring holds in one call");
    cudaStream_t st = as_stream(stream);
    const int nblocks = (n + kAppendRows - 1) / kAppendRows;
    RLOA_REQUIRE(rb->scratch != nullptr, "rloa_replay_append: rb->scratch is required");
    const bool count_pass = valid != nullptr && n > 16384;
    if (count_pass) {
        replay_count_kernel<<<nblocks, kAppendRows, 0, st>>>(valid, n, rb->scratch + 1);
        RLOA_LAUNCHED();
    }
    // scratch[0] = ticket counter (zero at rest), scratch[1..] = per-block valid counts (large appends only)
    if (!commit && !count_pass) {
        // rows-only copy beside the learn kernel (whose CTAs need EMPTY SMs): 256 rows per block, a quarter of the blocks
        replay_append_kernel<256><<<(n + 255) / 256, kAppendThreads, 0, st>>>(*rb, n, states, actions, rewards, next_states, dones, valid,
                                                                              nullptr, reinterpret_cast<unsigned*>(rb->scratch), 0);
    } else {
        replay_append_kernel<kAppendRows><<<nblocks, kAppendThreads, 0, st>>>(*rb, n, states, actions, rewards, next_states, dones, valid,
                                                                               count_pass ? rb->scratch + 1 : nullptr,
                                                                               reinterpret_cast<unsigned*>(rb->scratch), commit);
    }
    RLOA_LAUNCHED();
    return RLOA_OK;
}

extern "C" int rloa_replay_append(const rloa_replay* rb, int32_t n, const float* states, const float* actions,
                                  const float* rewards, const float* next_states, const uint8_t* dones,
                                  const uint8_t* valid, void* stream) {
    return replay_append_impl(rb, n, states, actions, rewards, next_states, dones, valid, stream, 1);
}

extern "C" int rloa_replay_append_rows(const rloa_replay* rb, int32_t n, const float* states, const float* actions,
                                       const float* rewards, const float* next_states, const uint8_t* dones,
                                       const uint8_t* valid, void* stream) {
    return replay_append_impl(rb, n, states, actions, rewards, next_states, dones, valid, stream, 0);
}

extern "C" int rloa_replay_commit(const rloa_replay* rb, int32_t n, const uint8_t* valid, void* stream) {
    int rc = check_rb(rb, "rloa_replay_commit");
    if (rc != RLOA_OK) return rc;
    RLOA_REQUIRE(n >= 1 && n <= rb->capacity, "rloa_replay_commit: bad row count");
    replay_commit_kernel<<<1, 256, 0, as_stream(stream)>>>(*rb, n, valid);
    RLOA_LAUNCHED();
    return RLOA_OK;
}

extern "C" int rloa_replay_sample(const rloa_replay* rb, int32_t batch, uint64_t seed, uint64_t draw,
                                  const uint64_t* draw_offset, float* states, float* actions, float* rewards,
                                  float* next_states, float* dones, int32_t* indices, void* stream) {
    int rc = check_rb(rb, "rloa_replay_sample");
    if (rc != RLOA_OK) return rc;
    RLOA_REQUIRE(batch >= 1 && states && actions && rewards && next_states, "rloa_replay_sample: null argument");
    const int blocks = (batch * 32 + 255) / 256;
    replay_sample_kernel<<<blocks, 256, 0, as_stream(stream)>>>(*rb, batch, seed, draw,
                                                               reinterpret_cast<const unsigned long long*>(draw_offset),
                                                               states, actions, rewards, next_states, dones, indices);
    RLOA_LAUNCHED();
    return RLOA_OK;
}
