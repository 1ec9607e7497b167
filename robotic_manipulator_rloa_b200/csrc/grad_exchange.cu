// librloa_b200: gradient exchange over NVLink peer memory, fused with the global-norm pass of the optimiser
// (include/rloa_b200.h, rloa_xchg_* and rloa_naf_learn_apply_xchg).
//
// The data-parallel NAF update has ONE exchange step: the mean of the flat main-net gradient
// (79,644 fp32 = 318,576 B) over the ranks, between backward and clip-norm + Adam.  The reference has nothing to
// replace here (single process, naf_components/naf_algorithm.py:208-210); the baseline is an NCCL all-reduce
// between two kernels.  This file does the exchange inside the optimiser's own kernels instead:
//
//   every rank owns one cudaMalloc'ed block  [ grad n | ready u64 x world | done u64 x world ]  exported with CUDA
//   IPC and mapped by every peer (NVLink 5 / NVSwitch P2P); flags are PUSHED: rank r stores its flag into slot r of
//   every peer's block, so waiting ranks poll their own HBM instead of issuing NVLink reads;
//   publish kernel: waits until every peer has finished READING my block for the previous step (done >= t - 1,
//       normally long true), copies the local gradient into it, and the last block releases ready = t;
//   reduce + Adam kernel (ONE launch, 80 co-resident blocks): waits for every peer's ready >= t, then each thread
//       sums its elements over the ranks IN RANK ORDER through the mapped peer pointers (the sum is bit-identical
//       on every rank, so clip norm and Adam update are too and the parameters never need a broadcast), the blocks
//       meet at a device-wide barrier with their squared-norm partials, release done = t, and apply clip + Adam +
//       the soft target update to the elements they summed.
// No host synchronisation, CUDA-graph capturable; a peer that never arrives trips a ~2 s device-side timeout that
// is reported through rloa_xchg_status instead of hanging the GPU.
#include <cstring>
#include <new>

#include "common.cuh"
#include "grad_exchange.cuh"
#include "optim.cuh"

namespace rloa {

constexpr int kXchgBlocks = kNormBlocks;   // one partial per block feeds adam_coefficients; all blocks co-resident (80 <= 148 SMs)

// thread 0 of the block spins until slot r of the LOCAL flag array is >= want for every peer r != rank
__device__ __forceinline__ bool wait_peers(const unsigned long long* local_flags, int world, int rank,
                                           unsigned long long want, int* __restrict__ status) {
    __shared__ int s_ok;
    if (threadIdx.x == 0) {
        int ok = 1;
        const long long t0 = clock64();
        for (int r = 0; r < world && ok; r++) {
            if (r == rank) continue;
            while (ld_sys(local_flags + r) < want) {
                if (clock64() - t0 > kXchgTimeoutCycles) { ok = 0; break; }     // tight spin: the flag sits in local HBM / L2
            }
        }
        if (!ok) atomicExch(status, 1);
        __threadfence_system();
        s_ok = ok;
    }
    __syncthreads();
    return s_ok != 0;
}

__global__ void __launch_bounds__(256)
xchg_publish_kernel(XchgPeers P, ReduceArgs r, float* __restrict__ local_grad, float* __restrict__ my_grad, int n,
                    int64_t* __restrict__ step_ptr, unsigned* __restrict__ ticket,
                    int* __restrict__ status) {
    const unsigned long long t = (unsigned long long)(*step_ptr) + 1ull;      // this update's number
    wait_peers(P.done[P.rank], P.world, P.rank, t - 1ull, status);
    // copy the local gradient into the exchange block; segments whose split-K partials were left unsummed by
    // rloa_naf_learn_step_xchg are summed here (fixed order) and written back to the local gradient as well
    const bool deferred = r.s[0].n > 0;
    const int n4 = ((reinterpret_cast<uintptr_t>(local_grad) & 15u) == 0) ? (n >> 2) : 0;
    for (int i = blockIdx.x * 256 + threadIdx.x; i < n4; i += gridDim.x * 256) {
        float4 g = reinterpret_cast<const float4*>(local_grad)[i];
        if (deferred) {
            g.x = splitk_element(r, 4 * i, g.x); g.y = splitk_element(r, 4 * i + 1, g.y);
            g.z = splitk_element(r, 4 * i + 2, g.z); g.w = splitk_element(r, 4 * i + 3, g.w);
            reinterpret_cast<float4*>(local_grad)[i] = g;
        }
        reinterpret_cast<float4*>(my_grad)[i] = g;
    }
    for (int i = 4 * n4 + blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256) {
        float g = local_grad[i];
        if (deferred) {
            g = splitk_element(r, i, g);
            local_grad[i] = g;
        }
        my_grad[i] = g;
    }
    // one system-scope fence per block, after the CTA barrier: fences are cumulative, so the block's stores (ordered
    // before thread 0 by bar.sync) are visible system-wide before the ticket / the ready flag (20k per-thread
    // fence.sc.sys would cost more than the copy itself)
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();
        const unsigned k = atomicAdd(ticket, 1u);
        if (k == gridDim.x - 1u) {                 // every block has read *step_ptr and stored its slice
            *ticket = 0u;
            *step_ptr += 1;                        // optimizer.step() counter (read by the kernels that follow)
            __threadfence_system();
            for (int r = 0; r < P.world; r++) st_sys(P.ready[r] + P.rank, t);     // push: slot `rank` in every block
        }
    }
}

__global__ void __launch_bounds__(256)
xchg_reduce_adam_kernel(XchgPeers P, float* __restrict__ sum_out, ParamTable pt, float* __restrict__ m, float* __restrict__ v,
                        rloa_naf_hyper hp, float* __restrict__ sq_partial, const int64_t* __restrict__ step_ptr,
                        unsigned* __restrict__ barrier, float* __restrict__ grad_norm_out, int* __restrict__ status) {
    __shared__ float red[256];
    __shared__ AdamCoef s_c;
    __shared__ int s_bad;
    const int n = pt.offset[14];
    const unsigned long long t = (unsigned long long)(*step_ptr);             // already incremented by publish
    wait_peers(P.ready[P.rank], P.world, P.rank, t, status);
    float s = 0.f;
    const int n4 = n >> 2;
    for (int i = blockIdx.x * 256 + threadIdx.x; i < n4; i += gridDim.x * 256) {
        float4 q[kXchgMaxWorld];
#pragma unroll
        for (int r = 0; r < kXchgMaxWorld; r++)                 // every rank's load in flight before the first add
            if (r < P.world) q[r] = ld_peer_f32x4(P.grad[r] + 4 * (size_t)i);
        float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int r = 0; r < kXchgMaxWorld; r++)                 // rank order: same bits on every rank
            if (r < P.world) { g.x += q[r].x; g.y += q[r].y; g.z += q[r].z; g.w += q[r].w; }
        reinterpret_cast<float4*>(sum_out)[i] = g;
        const float x0 = g.x * hp.grad_scale, x1 = g.y * hp.grad_scale, x2 = g.z * hp.grad_scale, x3 = g.w * hp.grad_scale;
        s = fmaf(x0, x0, s); s = fmaf(x1, x1, s); s = fmaf(x2, x2, s); s = fmaf(x3, x3, s);
    }
    for (int i = 4 * n4 + blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256) {
        float g = 0.f;
        for (int r = 0; r < P.world; r++) g += ld_peer_f32(P.grad[r] + i);
        sum_out[i] = g;
        const float x = g * hp.grad_scale;
        s = fmaf(x, x, s);
    }
    red[threadIdx.x] = s;
    __syncthreads();
    for (int off = 128; off > 0; off >>= 1) {
        if (threadIdx.x < off) red[threadIdx.x] += red[threadIdx.x + off];
        __syncthreads();
    }
    // ---- device-wide barrier (arrival count + generation): every block has read its peers and published its
    // partial; the last arriver resets the count, releases done = t to the peers and opens the next generation ----
    if (threadIdx.x == 0) {
        volatile unsigned* gen = barrier + 1;
        const unsigned gen0 = *gen;                             // cannot change before this block has arrived
        sq_partial[blockIdx.x] = red[0];
        __threadfence();
        if (atomicAdd(barrier, 1u) == gridDim.x - 1u) {         // last block: nobody reads the peers any more
            barrier[0] = 0u;
            __threadfence_system();
            for (int r = 0; r < P.world; r++) st_sys(P.done[r] + P.rank, t);
            __threadfence();
            atomicAdd(barrier + 1, 1u);
        } else {
            const long long t0 = clock64();
            while (*gen == gen0) {
                if (clock64() - t0 > kXchgTimeoutCycles) { atomicExch(status, 1); break; }
            }
        }
        __threadfence();
        s_c = adam_coefficients(sq_partial, (long long)t, hp);
        if (blockIdx.x == 0 && grad_norm_out != nullptr) *grad_norm_out = s_c.norm;
        // a wait that gave up (here or in any earlier update: the flag is sticky) means some peer's gradient may be
        // stale: every block of every later launch skips the optimiser, so the ranks stop moving instead of drifting
        // apart silently; rloa_xchg_status reports it
        s_bad = *reinterpret_cast<volatile int*>(status);
    }
    __syncthreads();
    if (s_bad) return;
    const AdamCoef c = s_c;
    // ---- clip + Adam + soft update of the elements this block summed (re-read from L2) ----
    for (int i = blockIdx.x * 256 + threadIdx.x; i < n4; i += gridDim.x * 256) {
        const float4 g = reinterpret_cast<const float4*>(sum_out)[i];
        adam_soft_update_element(pt, 4 * i, g.x, c, m, v, hp);
        adam_soft_update_element(pt, 4 * i + 1, g.y, c, m, v, hp);
        adam_soft_update_element(pt, 4 * i + 2, g.z, c, m, v, hp);
        adam_soft_update_element(pt, 4 * i + 3, g.w, c, m, v, hp);
    }
    for (int i = 4 * n4 + blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256)
        adam_soft_update_element(pt, i, sum_out[i], c, m, v, hp);
}

}  // namespace rloa

using namespace rloa;

struct rloa_xchg {
    int n = 0, rank = 0, world = 1;
    float* block = nullptr;                 // [n grad][pad][ready][done]
    float* sum = nullptr;                   // [n] summed gradient (local)
    unsigned* tickets = nullptr;            // [0] publish ticket, [1] barrier arrivals, [2] barrier generation
    int* status = nullptr;                  // [0] sticky time-out flag, [1] exchanges done by the fused learn kernel (its word tag)
    float* sq_partial = nullptr;            // [kXchgBlocks]
    void* mapped[kXchgMaxWorld] = {};
    XchgPeers peers{};
    bool connected = false;
};

// block layout: xchg_block_floats(n) floats (grad_exchange.cuh: the pulled gradient, then the tagged words of the fused learn
// kernel's two exchange rounds) | ready flags | done flags
static size_t xchg_flag_offset(int n) { return ((xchg_block_floats(n) * sizeof(float)) + 255) & ~(size_t)255; }

extern "C" int rloa_xchg_create(int32_t n_floats, rloa_xchg** out) {
    RLOA_REQUIRE(out != nullptr && n_floats >= 1, "rloa_xchg_create: bad argument");
    rloa_xchg* x = new (std::nothrow) rloa_xchg();
    RLOA_REQUIRE(x != nullptr, "rloa_xchg_create: out of host memory");
    x->n = n_floats;
    const size_t bytes = xchg_flag_offset(n_floats) + 2 * kXchgSlices * kXchgMaxWorld * sizeof(unsigned long long);
    if (cudaMalloc(&x->block, bytes) != cudaSuccess || cudaMalloc(&x->sum, (size_t)n_floats * sizeof(float)) != cudaSuccess ||
        cudaMalloc(&x->tickets, 4 * sizeof(unsigned)) != cudaSuccess || cudaMalloc(&x->status, 2 * sizeof(int)) != cudaSuccess ||
        cudaMalloc(&x->sq_partial, kXchgBlocks * sizeof(float)) != cudaSuccess) {
        set_error("rloa_xchg_create: cudaMalloc failed: %s", cudaGetErrorString(cudaGetLastError()));
        delete x;
        return RLOA_ERR_CUDA;
    }
    cudaMemset(x->block, 0, bytes);
    cudaMemset(x->tickets, 0, 4 * sizeof(unsigned));
    cudaMemset(x->status, 0, 2 * sizeof(int));
    cudaDeviceSynchronize();
    *out = x;
    return RLOA_OK;
}

extern "C" int rloa_xchg_handle(const rloa_xchg* x, uint8_t* handle_out_host) {
    RLOA_REQUIRE(x != nullptr && handle_out_host != nullptr, "rloa_xchg_handle: null argument");
    cudaIpcMemHandle_t h;
    RLOA_CUDA(cudaIpcGetMemHandle(&h, x->block));
    static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
    std::memcpy(handle_out_host, &h, 64);
    return RLOA_OK;
}

extern "C" int rloa_xchg_connect(rloa_xchg* x, int32_t rank, int32_t world, const uint8_t* handles_host) {
    RLOA_REQUIRE(x != nullptr && !x->connected, "rloa_xchg_connect: null or already connected");
    RLOA_REQUIRE(world >= 1 && world <= kXchgMaxWorld && rank >= 0 && rank < world, "rloa_xchg_connect: 1 <= world <= 16");
    RLOA_REQUIRE(world == 1 || handles_host != nullptr, "rloa_xchg_connect: handles missing");
    x->rank = rank;
    x->world = world;
    const size_t fo = xchg_flag_offset(x->n);
    for (int r = 0; r < world; r++) {
        void* base = x->block;
        if (r != rank) {
            cudaIpcMemHandle_t h;
            std::memcpy(&h, handles_host + 64 * (size_t)r, 64);
            if (cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
                set_error("rloa_xchg_connect: cudaIpcOpenMemHandle(rank %d) failed: %s", r, cudaGetErrorString(cudaGetLastError()));
                return RLOA_ERR_CUDA;
            }
            x->mapped[r] = base;
        }
        uint8_t* b = static_cast<uint8_t*>(base);
        x->peers.grad[r] = reinterpret_cast<const float*>(b);
        x->peers.ready[r] = reinterpret_cast<unsigned long long*>(b + fo);
        x->peers.done[r] = reinterpret_cast<unsigned long long*>(b + fo + kXchgSlices * kXchgMaxWorld * sizeof(unsigned long long));
    }
    x->peers.world = world;
    x->peers.rank = rank;
    x->connected = true;
    return RLOA_OK;
}

extern "C" int rloa_xchg_status(const rloa_xchg* x) {
    if (x == nullptr) return RLOA_ERR_INVALID;
    int s = 0;
    if (cudaMemcpy(&s, x->status, sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) return RLOA_ERR_CUDA;
    return s;
}

extern "C" void rloa_xchg_destroy(rloa_xchg* x) {
    if (x == nullptr) return;
    cudaDeviceSynchronize();
    for (int r = 0; r < x->world; r++)
        if (x->mapped[r] != nullptr) cudaIpcCloseMemHandle(x->mapped[r]);
    if (x->block) cudaFree(x->block);
    if (x->sum) cudaFree(x->sum);
    if (x->tickets) cudaFree(x->tickets);
    if (x->status) cudaFree(x->status);
    if (x->sq_partial) cudaFree(x->sq_partial);
    delete x;
}

// used by rloa_naf_learn_apply_xchg (naf.cu): publish, then exchange + clip + Adam + soft update in one kernel
namespace rloa {
void xchg_peers(const rloa_xchg* x, XchgPeers* out, int** status) {
    *out = x->peers;
    *status = x->status;
}
int xchg_connected_world(const rloa_xchg* x) { return (x != nullptr && x->connected) ? x->world : 0; }

int xchg_exchange_adam(rloa_xchg* x, float* local_grad, const ReduceArgs& deferred, const ParamTable& pt, float* m, float* v,
                       int64_t* step_ptr, const rloa_naf_hyper& hp, float* grad_norm, cudaStream_t st) {
    RLOA_REQUIRE(x != nullptr && x->connected, "gradient exchange: rloa_xchg_connect was not called");
    const int n = pt.offset[14];
    RLOA_REQUIRE(n == x->n, "gradient exchange: gradient length does not match the exchange buffer");
    xchg_publish_kernel<<<kXchgBlocks, 256, 0, st>>>(x->peers, deferred, local_grad, x->block, n, step_ptr, x->tickets,
                                                     x->status);
    RLOA_LAUNCHED();
    xchg_reduce_adam_kernel<<<kXchgBlocks, 256, 0, st>>>(x->peers, x->sum, pt, m, v, hp, x->sq_partial, step_ptr,
                                                         x->tickets + 1, grad_norm, x->status);
    RLOA_LAUNCHED();
    return RLOA_OK;
}
}  // namespace rloa
