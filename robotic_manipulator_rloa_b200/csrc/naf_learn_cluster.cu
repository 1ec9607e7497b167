// NAFAgent.learn in ONE launch on the 5th-generation tensor cores (reference naf_components/naf_algorithm.py:180-226:
// target forward, TD target, main forward, MSE loss, backward, clip_grad_norm_, Adam, soft update).
//
// Two thread-block clusters of 8 CTAs — cluster 0 = target network on next_states, cluster 1 = main network on states —
// 128 replay rows per CTA (batch <= 1024), 512 threads per CTA.  Nothing but the weights, the batch and the split-K partials
// of the weight gradients touches global memory: pre-activations live in TMEM (z1 in columns 0-255, z2 in 256-511), the
// bf16 operand tiles in shared memory, and the train-mode BatchNorm couples the rows through DISTRIBUTED SHARED MEMORY:
// every CTA leaves its per-column partial statistics in its own shared memory, one barrier.cluster later every CTA reads
// the eight partials of each column over DSMEM and merges them in rank order (Chan's update), so all eight hold the same
// batch statistics without a global round trip.  Four such exchanges (BN1 / BN2 forward, BN2 / BN1 backward), two more for
// the gradient-norm and the final hand-over, replace nine kernel boundaries of the multi-launch path.
//
// Every contraction runs as tcgen05.mma with fp32 TMEM accumulators:
//   z1 = x W1^T            kind::tf32, observations split hi + lo (2^-22), K = 24
//   z2 = a1 W2^T           bf16 128x256x256                      heads = a2 Wh^T        bf16 128x64x256
//   dWh^T = a2^T dzh       A and B read MN-major (contraction over the 128 batch rows of both tiles, nothing transposed)
//   da2 = dzh Wh           B = the forward head image read MN-major
//   da1 = dz2 W2           B = the forward W2 image read MN-major (TMA-reloaded: its slot held the head tiles meanwhile)
//   dW2 = dz2^T a1         A and B MN-major, 256 x 256 fp32 = all 512 TMEM columns
//   dW1 = dz1^T [x_hi|x_lo] A and B MN-major, bf16 hi / lo split of the observations
// (csrc/umma_probe.cu pins the MN-major views).  z1 is not kept across the backward: its TMEM columns are needed, and
// recomputing it is six K = 8 instructions.  The TD target of a row needs V'(s') of the same row from the other cluster:
// the target CTA of rank r writes y for its 128 rows and raises a flag, the main CTA of rank r waits for it in its head
// phase (the target never waits for the main, so the pair cannot deadlock; the flag returns to 0: graph-replay safe).
// Tail: every CTA dumps its partial weight gradients, one cluster barrier, then each CTA sums an eighth of the flat
// gradient over the 8 partials in rank order, the squared-norm partials meet over DSMEM, and clip + Adam + soft target
// update are applied to the same elements (optim.cuh arithmetic).  Deterministic: no atomics on data, fixed orders.
//
// Numerics: bf16 operands (tf32 for W1), fp32 accumulation and fp32 BatchNorm statistics from the TMEM accumulators; the
// bound against the all-fp32 path is stated in tests/test_naf_learn_cluster_gpu.py.  sm_100a only.
#include <cuda_bf16.h>

#include "common.cuh"
#include "bn_fuse.cuh"
#include "grad_exchange.cuh"
#include "naf_learn_cluster.cuh"
#include "optim.cuh"
#include "replay_sample.cuh"
#include "tc_common.cuh"

namespace rloa {

using namespace tc;

namespace lc {
constexpr int H = 256, KP = 32, NHP = 64, ROWS = 128, THREADS = 512, CL = 8;
// weight image of one network (global memory, written by learn_pack_kernel in shared-memory byte order)
constexpr uint32_t W2_BYTES = H * H * 2;                 // 131072: 4 k-blocks x [256 rows][128 B]
constexpr uint32_t WH_BYTES = NHP * H * 2;               // 32768: 4 k-blocks x [64 rows][128 B]
constexpr uint32_t W1_BYTES = H * KP * 4;                // 32768: tf32 [256 rows][128 B]
constexpr uint32_t VEC_FLOATS = 2 * H + NHP;             // b1 | b2 | bh
constexpr uint32_t VEC_BYTES = VEC_FLOATS * 4;           // 2304
constexpr uint32_t OFF_W2 = 0, OFF_WH = W2_BYTES, OFF_W1 = OFF_WH + WH_BYTES, OFF_VEC = OFF_W1 + W1_BYTES;
constexpr uint32_t IMAGE_BYTES = OFF_VEC + VEC_BYTES;    // 198912
// shared memory map (offsets from the 1024-byte aligned base)
constexpr uint32_t R1 = 0;                               // 64 KB  A-operand tile [4 k-blocks][128 rows][128 B]
constexpr uint32_t R0 = 65536;                           // 128 KB W2 image | head image + scratch | a1 + dz1
constexpr uint32_t S1 = R0 + 131072;                     // 16 KB  dzh tile | reduction scratch + coefficients | x tile
constexpr uint32_t COEF = S1 + 16384;                    // [2 layers][sc | sh | mean | rstd][256] fp32 = 8 KB
constexpr uint32_t XBUF = COEF + 8192;                   // [2][3][256] fp32 = 6 KB: this CTA's partial statistics (DSMEM)
constexpr uint32_t VEC = XBUF + 6144;                    // b1 | b2 | bh
constexpr uint32_t MISC = VEC + VEC_BYTES;               // mbarriers, TMEM slot, small reductions
constexpr uint32_t SMEM_BYTES = MISC + 512 + 1024;       // + alignment slack
static_assert(SMEM_BYTES <= 232448, "shared memory budget of one SM");
constexpr uint32_t A_BLK = ROWS * 128;                   // k-block stride of a 128-row tile
constexpr uint32_t HEAD_SCRATCH = R0 + 32768;            // fp32 [128][65] head pre-activations / gradients (33,280 B)
}  // namespace lc

size_t learn_cluster_image_bytes() { return 2 * (size_t)lc::IMAGE_BYTES; }

// ---------------------------------------------------------------------------------------------------
// weight images of both networks (net 0 = target, 1 = main)
__global__ void __launch_bounds__(256) learn_pack_kernel(rloa_naf_params pt, rloa_naf_params pm, uint8_t* __restrict__ images) {
    using namespace lc;
    const rloa_naf_params& p = blockIdx.y == 0 ? pt : pm;
    uint8_t* image = images + (size_t)blockIdx.y * IMAGE_BYTES;
    const int S = p.state_size, A = p.action_size, NL = A * (A + 1) / 2, NH = A + 1 + NL;
    const int t = blockIdx.x * 256 + threadIdx.x;
    if (t < H * H / 8) {                                  // W2: one 16-byte chunk (8 consecutive k of row n)
        const int n = t >> 5, kc = t & 31;
        const float4 a = *reinterpret_cast<const float4*>(p.w2 + (size_t)n * H + kc * 8);
        const float4 b = *reinterpret_cast<const float4*>(p.w2 + (size_t)n * H + kc * 8 + 4);
        const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
        *reinterpret_cast<uint4*>(image + OFF_W2 + (uint32_t)(kc >> 3) * (H * 128) + sw128_chunk_offset(n, kc & 7)) = pack8_bf16(v);
        return;
    }
    int u = t - H * H / 8;
    if (u < NHP * H / 8) {                                // head rows: [0,A) mu, A value, A+1.. matrix entries, rest zero
        const int o = u >> 5, kc = u & 31;
        const float* src = o < A ? p.w_mu + (size_t)o * H : (o == A ? p.w_v : (o < NH ? p.w_l + (size_t)(o - A - 1) * H : nullptr));
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; i++) v[i] = src != nullptr ? src[kc * 8 + i] : 0.f;
        *reinterpret_cast<uint4*>(image + OFF_WH + (uint32_t)(kc >> 3) * (NHP * 128) + sw128_chunk_offset(o, kc & 7)) = pack8_bf16(v);
        return;
    }
    u -= NHP * H / 8;
    if (u < H * KP / 4) {                                 // W1 as tf32, K padded to 32
        const int j = u >> 3, kc = u & 7;
        float4 v;
        float* vv = reinterpret_cast<float*>(&v);
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int k = kc * 4 + i;
            vv[i] = k < S ? to_tf32(p.w1[(size_t)j * S + k]) : 0.f;
        }
        *reinterpret_cast<float4*>(image + OFF_W1 + sw128_chunk_offset(j, kc)) = v;
        return;
    }
    u -= H * KP / 4;
    float* vec = reinterpret_cast<float*>(image + OFF_VEC);
    if (u < H) {
        vec[u] = p.b1[u];
        vec[H + u] = p.b2[u];
        if (u < NHP) vec[2 * H + u] = u < A ? p.b_mu[u] : (u == A ? p.b_v[0] : (u < NH ? p.b_l[u - A - 1] : 0.f));
    }
}

// ---------------------------------------------------------------------------------------------------
struct LearnClusterArgs {
    const uint8_t* images;                  // [2][IMAGE_BYTES]
    rloa_naf_params p[2];                   // 0 = target, 1 = main
    const float *states, *actions, *rewards, *next_states, *dones;
    int B, S, A, NL, NH;
    rloa_naf_hyper hp;
    float* y;                               // [1024] TD targets, target cluster -> main cluster
    unsigned* yflag;                        // [8], zero at rest
    unsigned* aflag;                        // [8], zero at rest: main CTA r -> target CTA r, "the optimiser coefficients are final"
    float* acoef;                           // [8][8]: clip | bc1 | bc2s | norm | skip (a timed-out exchange) of that hand-over
    float *part_w2, *part_w1, *part_wh;     // [8][256*256], [8][256*S], [8][NH][256]
    float *part_hb, *part_loss;             // [8][64], [8]
    float *grad, *loss, *gnorm;             // outputs
    ParamTable pt;
    float *m, *v;
    int64_t* step;
    int do_adam;
    int off_w1, off_b1, off_bn1w, off_bn1b, off_w2, off_b2, off_bn2w, off_bn2b, off_wmu, off_bmu, off_wv, off_bv, off_wl, off_bl, n_params;
    float* dbg;                             // optional: z1 | z2 | dzh | dz2 | da1 | dz1 as fp32 [1024][256] each (dzh [1024][64])
    long long* prof;                        // optional: [16 CTAs][32] clock64 phase stamps
    // fused ReplayBuffer.sample: row i of the batch is ring slot replay_sample_slot(rb, i, seed, draw + *draw_offset), read
    // straight from the ring (states / actions / rewards / next_states / dones above are unused then)
    // N > 1 ranks: the gradient exchange over NVLink peer memory INSIDE the tail — CTA c owns slice c of the flat gradient
    // on every rank; it STORES its local sums into region (1 + rank) of every rank's exchange block (posted writes through
    // the mapped peer pointers), pushes a ready flag behind them, waits for the peers' flags, and sums the slice over the
    // ranks in rank order from its own block — no NVLink read round trip on the critical path
    int use_xchg;
    XchgPeers xp;
    int* xstatus;
    int use_replay;
    rloa_replay rb;
    unsigned long long rs_seed, rs_draw;
    const unsigned long long* rs_draw_offset;
    // pending rows: the step's transitions, being copied into the ring by a concurrent rloa_replay_append_rows and not
    // committed yet (the ring's cursor still excludes them).  The sampler sees the ring as it will be after the commit; a
    // drawn slot that falls into the pending range is read from these arrays (row = the k-th valid one) instead of the ring.
    const float *pd_states, *pd_actions, *pd_rewards, *pd_next_states;
    const uint8_t *pd_dones, *pd_valid;
    int pd_n;                               // 0: none; at most kPendingMax
};

__device__ __forceinline__ uint32_t cluster_rank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ float ld_dsmem(const float* local, uint32_t rank) {
    uint32_t ra;
    float v;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(smem_u32(local)), "r"(rank));
    asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(ra));
    return v;
}

// mbarrier wait that cannot hang the GPU: a barrier that does not complete within ~2 s aborts the launch (a bug, never a
// legitimate wait: every barrier here is fed by this CTA's own TMA copies or MMAs)
__device__ __forceinline__ void mbar_wait_bounded(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    const long long t0 = clock64();
    uint32_t done = 0;
    while (!done) {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}\n"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
        if (!done && clock64() - t0 > 4000000000ll) __trap();
    }
}

// 32 accumulator columns without the wait: several loads can be in flight before one tmem_ld_wait()
__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// column sums over the 32 rows of a warp: lane l enters with v[i] = (row l, column i), leaves with the sum of column l
__device__ __forceinline__ float colsum32(float (&v)[32], int lane) {
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) {
        const bool up = (lane & s) != 0;
#pragma unroll
        for (int i = 0; i < s; i++) {
            const float send = up ? v[i] : v[i + s];
            const float keep = up ? v[i + s] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
        }
    }
    return v[0];
}

__device__ __forceinline__ void store_chunks32(uint8_t* tile, int r, int c0, const float (&a)[32]) {
    // 32 consecutive columns c0.. of row r -> four 16-byte bf16 chunks of a K-major SWIZZLE_128B tile (128 rows per k-block)
#pragma unroll
    for (int g = 0; g < 4; g++) {
        const float x[8] = {a[8 * g], a[8 * g + 1], a[8 * g + 2], a[8 * g + 3], a[8 * g + 4], a[8 * g + 5], a[8 * g + 6], a[8 * g + 7]};
        const int chunk = (c0 >> 3) + g;
        *reinterpret_cast<uint4*>(tile + (chunk >> 3) * lc::A_BLK + sw128_chunk_offset(r, chunk & 7)) = pack8_bf16(x);
    }
}

// one element of clip + Adam + soft target update: the arithmetic of adam_soft_update_element (optim.cuh) on operands that
// are already in registers
__device__ __forceinline__ void adam_math(float grad_i, const AdamCoef& c, const rloa_naf_hyper& hp, float& m, float& v, float& p,
                                          float& pt) {
    const float gq = grad_i * c.clip;
    m = fmaf(1.f - hp.beta1, gq - m, m);
    v = fmaf(1.f - hp.beta2, gq * gq - v, v);
    const float denom = sqrtf(v) / c.bc2s + hp.eps;
    p = p - (hp.lr / c.bc1) * (m / denom);
    pt = hp.tau * p + (1.f - hp.tau) * pt;
}


// ---- per-thread context of the phase functions below.  They are deliberately NOT inlined: every one is called from two or
// three places (both BatchNorm layers, forward and recompute), and the kernel executes its code exactly once per launch, so
// instruction FETCH (230 KB of straight-line code when everything was inlined) was a first-order cost, doubled by a cold L2.
struct LcCtx {
    uint32_t tlane;                 // TMEM address of this warp's 32 lanes
    int tid, q, r, lane, wq, rank, B, nrows, row, net;
    bool valid;
    float *coef, *xbuf, *scratch, *colaux, *bcoef;
    float* dbg;
};

__device__ __forceinline__ void lc_dbg_rows32(const LcCtx& c, int section, int width, int c0, const float (&a)[32]) {
    if (c.dbg != nullptr && c.net == 1 && c.valid) {
        float* d = c.dbg + (size_t)section * 1024 * 256 + (size_t)c.row * width + c0;
#pragma unroll
        for (int i = 0; i < 32; i++) d[i] = a[i];
    }
}

// train-mode BatchNorm statistics of the pre-activations acc + bias (acc = TMEM columns tcol..) in ONE pass over the
// accumulator: per column sum and sum of squares of (z - K), K = the layer's running mean (a pivot near the batch mean keeps
// the one-pass variance as accurate as the two-pass form); CTA partial (mean, M2) -> cluster exchange over DSMEM -> Chan
// merge in rank order -> coef[layer] = sc | shb | rstd | xo with the linear bias folded in:
//   relu(bn(z)) = max(acc sc + shb, 0),   xhat = acc rstd + xo.   Running statistics are updated by rank 0.
__device__ __forceinline__ void lc_forward_stats(const LcCtx& c, int layer, uint32_t tcol, const float* bias, const float* bn_w,
                                              const float* bn_b, float* run_mean, float* run_var, int64_t* batches) {
    using namespace lc;
    float* xb = c.xbuf + (layer & 1) * 3 * 256;
    float* scratch = c.scratch;
    const int tid = c.tid;
    if (tid < 256) c.colaux[tid] = bias[tid] - run_mean[tid];
    __syncthreads();
#pragma unroll 1
    for (int p = 0; p < 2; p++) {
        const int c0 = c.q * 64 + p * 32;
        uint32_t v[32];
        tmem_ld32(c.tlane + tcol + c0, v);
        float d[32], e[32];
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const float4 k4 = *reinterpret_cast<const float4*>(c.colaux + c0 + 4 * j);
            d[4 * j] = __uint_as_float(v[4 * j]) + k4.x;
            d[4 * j + 1] = __uint_as_float(v[4 * j + 1]) + k4.y;
            d[4 * j + 2] = __uint_as_float(v[4 * j + 2]) + k4.z;
            d[4 * j + 3] = __uint_as_float(v[4 * j + 3]) + k4.w;
        }
#pragma unroll
        for (int i = 0; i < 32; i++) {
            d[i] = c.valid ? d[i] : 0.f;
            e[i] = d[i] * d[i];
        }
        const float s1 = colsum32(d, c.lane), s2 = colsum32(e, c.lane);
        scratch[c.wq * 256 + c0 + c.lane] = s1;
        scratch[1024 + c.wq * 256 + c0 + c.lane] = s2;
    }
    __syncthreads();
    if (tid < 256) {
        const float S1 = (scratch[tid] + scratch[256 + tid]) + (scratch[512 + tid] + scratch[768 + tid]);
        const float S2 = (scratch[1024 + tid] + scratch[1280 + tid]) + (scratch[1536 + tid] + scratch[1792 + tid]);
        const float n = (float)c.nrows;
        const float md = c.nrows > 0 ? S1 / n : 0.f;
        xb[tid] = md + run_mean[tid];                          // CTA mean of z
        xb[256 + tid] = fmaxf(S2 - md * S1, 0.f);              // CTA centred second moment
    }
    cluster_sync_all();
    if (tid < 256) {
        const int B = c.B;
        float pm[CL], pM[CL];
#pragma unroll
        for (int k = 0; k < CL; k++) { pm[k] = ld_dsmem(xb + tid, k); pM[k] = ld_dsmem(xb + 256 + tid, k); }
        float n = 0.f, mean = 0.f, M2 = 0.f;
#pragma unroll
        for (int k = 0; k < CL; k++) {
            const float nb = (float)min(max(B - k * ROWS, 0), ROWS);
            if (nb > 0.f) {
                const float nn = n + nb, delta = pm[k] - mean;
                mean = fmaf(delta, nb / nn, mean);
                M2 = M2 + pM[k] + delta * delta * (n * nb / nn);
                n = nn;
            }
        }
        const float var = M2 / (float)B;
        const float rstd = 1.f / sqrtf(var + kBnEps);
        const float sc = bn_w[tid] * rstd;
        const float sh = fmaf(-mean, sc, bn_b[tid]);
        float* cf = c.coef + layer * 4 * 256;
        cf[tid] = sc;
        cf[256 + tid] = fmaf(bias[tid], sc, sh);
        cf[512 + tid] = rstd;
        cf[768 + tid] = (bias[tid] - mean) * rstd;
        if (c.rank == 0) {
            const float unbiased = B > 1 ? M2 / (float)(B - 1) : var;
            run_mean[tid] = fmaf(kBnMomentum, mean - run_mean[tid], run_mean[tid]);
            run_var[tid] = fmaf(kBnMomentum, unbiased - run_var[tid], run_var[tid]);
            if (tid == 0 && batches != nullptr) *batches += 1;
        }
    }
    __syncthreads();
}

// relu(bn(z)) of TMEM columns tcol.. -> bf16 K-major tile (rows past the batch are zero)
__device__ __forceinline__ void lc_activation_to_tile(const LcCtx& c, int layer, uint32_t tcol, uint8_t* tile) {
    const float* cf = c.coef + layer * 4 * 256;
#pragma unroll 1
    for (int p = 0; p < 2; p++) {
        const int c0 = c.q * 64 + p * 32;
        uint32_t v[32];
        tmem_ld32(c.tlane + tcol + c0, v);
        float a[32];
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const float4 sc = *reinterpret_cast<const float4*>(cf + c0 + 4 * j);
            const float4 sh = *reinterpret_cast<const float4*>(cf + 256 + c0 + 4 * j);
            a[4 * j] = fmaxf(fmaf(__uint_as_float(v[4 * j]), sc.x, sh.x), 0.f);
            a[4 * j + 1] = fmaxf(fmaf(__uint_as_float(v[4 * j + 1]), sc.y, sh.y), 0.f);
            a[4 * j + 2] = fmaxf(fmaf(__uint_as_float(v[4 * j + 2]), sc.z, sh.z), 0.f);
            a[4 * j + 3] = fmaxf(fmaf(__uint_as_float(v[4 * j + 3]), sc.w, sh.w), 0.f);
        }
#pragma unroll
        for (int i = 0; i < 32; i++) a[i] = c.valid ? a[i] : 0.f;
        store_chunks32(tile, c.r, c0, a);
    }
}

// ReLU + BatchNorm backward of one layer in ONE pass over the two accumulators (da in TMEM columns dcol.., the layer's
// pre-activations in zcol..): column sums of g = da [bn(z) > 0] and g xhat over the whole batch through the cluster, then
// the coefficients of dz = k1 (g - c1 - c2 xhat).  gamma / beta gradients are totals: rank 0 writes them; the linear-bias
// gradient under a train-mode BatchNorm is the rounding residue k1 (sum g - B c1).
__device__ __forceinline__ void lc_backward_stats(const LcCtx& c, int layer, uint32_t dcol, uint32_t zcol, const float* bn_w, int xslot,
                                               float* d_w, float* d_b, float* d_lin) {
    using namespace lc;
    const float* cf = c.coef + layer * 4 * 256;
    float* xb = c.xbuf + xslot * 3 * 256;
    float* scratch = c.scratch;
    const int tid = c.tid;
#pragma unroll 1
    for (int p = 0; p < 2; p++) {
        const int c0 = c.q * 64 + p * 32;
        uint32_t vd[32], vz[32];
        tmem_ld32_nowait(c.tlane + zcol + c0, vz);
        tmem_ld32_nowait(c.tlane + dcol + c0, vd);
        tmem_ld_wait();
        float a0[32], a1[32];
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const float4 sc = *reinterpret_cast<const float4*>(cf + c0 + 4 * j);
            const float4 sh = *reinterpret_cast<const float4*>(cf + 256 + c0 + 4 * j);
            const float4 rs = *reinterpret_cast<const float4*>(cf + 512 + c0 + 4 * j);
            const float4 xo = *reinterpret_cast<const float4*>(cf + 768 + c0 + 4 * j);
            const float scv[4] = {sc.x, sc.y, sc.z, sc.w}, shv[4] = {sh.x, sh.y, sh.z, sh.w};
            const float rsv[4] = {rs.x, rs.y, rs.z, rs.w}, xov[4] = {xo.x, xo.y, xo.z, xo.w};
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int i = 4 * j + u;
                const float acc = __uint_as_float(vz[i]);
                const bool on = c.valid && fmaf(acc, scv[u], shv[u]) > 0.f;
                const float gg = on ? __uint_as_float(vd[i]) : 0.f;
                a0[i] = gg;
                a1[i] = gg * fmaf(acc, rsv[u], xov[u]);
            }
        }
        const float s0 = colsum32(a0, c.lane), s1 = colsum32(a1, c.lane);
        scratch[c.wq * 256 + c0 + c.lane] = s0;
        scratch[1024 + c.wq * 256 + c0 + c.lane] = s1;
    }
    __syncthreads();
    if (tid < 256) {
        xb[tid] = (scratch[tid] + scratch[256 + tid]) + (scratch[512 + tid] + scratch[768 + tid]);
        xb[256 + tid] = (scratch[1024 + tid] + scratch[1280 + tid]) + (scratch[1536 + tid] + scratch[1792 + tid]);
    }
    cluster_sync_all();
    if (tid < 256) {
        float p0[CL], p1[CL];
#pragma unroll
        for (int k = 0; k < CL; k++) { p0[k] = ld_dsmem(xb + tid, k); p1[k] = ld_dsmem(xb + 256 + tid, k); }
        float tg = 0.f, tgx = 0.f;
#pragma unroll
        for (int k = 0; k < CL; k++) { tg += p0[k]; tgx += p1[k]; }
        const float k1 = bn_w[tid] * cf[512 + tid], c1 = tg / (float)c.B, c2 = tgx / (float)c.B;
        c.bcoef[tid] = k1;
        c.bcoef[256 + tid] = k1 * c1;
        c.bcoef[512 + tid] = k1 * c2;
        if (c.rank == 0) {
            d_w[tid] = tgx;
            d_b[tid] = tg;
            d_lin[tid] = k1 * (tg - (float)c.B * c1);
        }
    }
    __syncthreads();
}

// dz = k1 g - k1 c1 - k1 c2 xhat of this thread's 64 columns -> bf16 K-major tile `dz_tile`; with act_tile != NULL also
// relu(bn(z)) of the same pre-activations -> `act_tile` (layer 1: a1 is recomputed for dW2 while z1 is in registers anyway)
__device__ __forceinline__ void lc_dz_to_tile(const LcCtx& c, int layer, uint32_t dcol, uint32_t zcol, uint8_t* dz_tile, uint8_t* act_tile,
                                           int dbg_dz, int dbg_da) {
    const float* cf = c.coef + layer * 4 * 256;
    const float* bcoef = c.bcoef;
#pragma unroll 1
    for (int p = 0; p < 2; p++) {
        const int c0 = c.q * 64 + p * 32;
        uint32_t vd[32], vz[32];
        tmem_ld32_nowait(c.tlane + dcol + c0, vd);
        tmem_ld32_nowait(c.tlane + zcol + c0, vz);
        tmem_ld_wait();
        float d[32];
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const float4 sc = *reinterpret_cast<const float4*>(cf + c0 + 4 * j);
            const float4 sh = *reinterpret_cast<const float4*>(cf + 256 + c0 + 4 * j);
            const float4 rs = *reinterpret_cast<const float4*>(cf + 512 + c0 + 4 * j);
            const float4 xo = *reinterpret_cast<const float4*>(cf + 768 + c0 + 4 * j);
            const float4 k1 = *reinterpret_cast<const float4*>(bcoef + c0 + 4 * j);
            const float4 kc1 = *reinterpret_cast<const float4*>(bcoef + 256 + c0 + 4 * j);
            const float4 kc2 = *reinterpret_cast<const float4*>(bcoef + 512 + c0 + 4 * j);
            const float scv[4] = {sc.x, sc.y, sc.z, sc.w}, shv[4] = {sh.x, sh.y, sh.z, sh.w};
            const float rsv[4] = {rs.x, rs.y, rs.z, rs.w}, xov[4] = {xo.x, xo.y, xo.z, xo.w};
            const float k1v[4] = {k1.x, k1.y, k1.z, k1.w}, c1v[4] = {kc1.x, kc1.y, kc1.z, kc1.w}, c2v[4] = {kc2.x, kc2.y, kc2.z, kc2.w};
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int i = 4 * j + u;
                const float acc = __uint_as_float(vz[i]);
                const float gg = fmaf(acc, scv[u], shv[u]) > 0.f ? __uint_as_float(vd[i]) : 0.f;
                const float dzv = fmaf(-c2v[u], fmaf(acc, rsv[u], xov[u]), fmaf(k1v[u], gg, -c1v[u]));
                d[i] = c.valid ? dzv : 0.f;
            }
        }
        store_chunks32(dz_tile, c.r, c0, d);
        if (c.dbg != nullptr) {
            lc_dbg_rows32(c, dbg_dz, 256, c0, d);
            if (dbg_da >= 0) {
                float da[32];
#pragma unroll
                for (int i = 0; i < 32; i++) da[i] = __uint_as_float(vd[i]);
                lc_dbg_rows32(c, dbg_da, 256, c0, da);
            }
        }
        if (act_tile != nullptr) {
            float a[32];
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const float4 sc = *reinterpret_cast<const float4*>(cf + c0 + 4 * j);
                const float4 sh = *reinterpret_cast<const float4*>(cf + 256 + c0 + 4 * j);
                a[4 * j] = fmaxf(fmaf(__uint_as_float(vz[4 * j]), sc.x, sh.x), 0.f);
                a[4 * j + 1] = fmaxf(fmaf(__uint_as_float(vz[4 * j + 1]), sc.y, sh.y), 0.f);
                a[4 * j + 2] = fmaxf(fmaf(__uint_as_float(vz[4 * j + 2]), sc.z, sh.z), 0.f);
                a[4 * j + 3] = fmaxf(fmaf(__uint_as_float(vz[4 * j + 3]), sc.w, sh.w), 0.f);
            }
#pragma unroll
            for (int i = 0; i < 32; i++) a[i] = c.valid ? a[i] : 0.f;
            store_chunks32(act_tile, c.r, c0, a);
        }
    }
}

constexpr int kXchgInKernelWorld = 8;      // ranks the in-kernel exchange keeps in flight per round (one NVSwitch box)
struct TrueT { static constexpr bool value = true; };
struct FalseT { static constexpr bool value = false; };

// clip + Adam + soft target update on one half of CTA `rank`'s share of the parameters: W2 chunks [j0, j1) of its four and
// slots [s0, s1) of its six other elements.  The main CTA takes its gradients from shared memory (t_* from the gather pass);
// the target-cluster CTA of the same rank - idle since its forward pass - takes the other half from the flat gradient in
// global memory and recomputes the flat indices.  Every element is touched by exactly one thread of the grid.
__device__ __forceinline__ void lc_adam_apply(const LearnClusterArgs& g, const AdamCoef& c, int rank, int tid, int j0, int j1, int s0,
                                              int s1, const float4* t_gw, const float* t_gr, const int* t_gi) {
    using namespace lc;
#pragma unroll 2
    for (int j = j0; j < j1; j++) {                            // W2: operands as float4 straight from the tensors
        const int i4 = rank * 2048 + j * THREADS + tid;
        float4 am = reinterpret_cast<const float4*>(g.m + g.off_w2)[i4], av = reinterpret_cast<const float4*>(g.v + g.off_w2)[i4];
        float4 ap = reinterpret_cast<const float4*>(g.pt.main[4])[i4], at = reinterpret_cast<const float4*>(g.pt.target[4])[i4];
        const float4 gw = t_gw != nullptr ? t_gw[j * THREADS + tid] : __ldcg(reinterpret_cast<const float4*>(g.grad + g.off_w2) + i4);
        adam_math(gw.x, c, g.hp, am.x, av.x, ap.x, at.x);
        adam_math(gw.y, c, g.hp, am.y, av.y, ap.y, at.y);
        adam_math(gw.z, c, g.hp, am.z, av.z, ap.z, at.z);
        adam_math(gw.w, c, g.hp, am.w, av.w, ap.w, at.w);
        reinterpret_cast<float4*>(g.m + g.off_w2)[i4] = am;
        reinterpret_cast<float4*>(g.v + g.off_w2)[i4] = av;
        reinterpret_cast<float4*>(g.pt.main[4])[i4] = ap;
        reinterpret_cast<float4*>(g.pt.target[4])[i4] = at;
    }
    const int rest = g.n_params - H * H, per = (rest + CL - 1) / CL;
#pragma unroll 3
    for (int j = s0; j < s1; j++) {
        int i = -1;
        if (t_gi != nullptr) {
            i = t_gi[j * THREADS + tid];
        } else {
            const int e = rank * per + j * THREADS + tid;
            if (j * THREADS + tid < per && e < rest) i = e < g.off_w2 ? e : e + H * H;
        }
        if (i >= 0) {
            int t = 0;
#pragma unroll
            for (int k = 1; k < 14; k++) t += (i >= g.pt.offset[k]) ? 1 : 0;
            const int jj = i - g.pt.offset[t];
            float rm = g.m[i], rv = g.v[i], rp = g.pt.main[t][jj], rt = g.pt.target[t][jj];
            adam_math(t_gr != nullptr ? t_gr[j * THREADS + tid] : __ldcg(g.grad + i), c, g.hp, rm, rv, rp, rt);
            g.m[i] = rm;
            g.v[i] = rv;
            g.pt.main[t][jj] = rp;
            g.pt.target[t][jj] = rt;
        }
    }
}

__global__ void __cluster_dims__(8, 1, 1) __launch_bounds__(lc::THREADS, 1)
naf_learn_cluster_kernel(const __grid_constant__ LearnClusterArgs g) {
    using namespace lc;
    extern __shared__ uint8_t lc_smem_raw[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int r = tid & 127, q = tid >> 7, wq = warp & 3;           // tile row (= TMEM lane), 64-column quarter, lane quarter
    const int net = blockIdx.x >> 3;                                  // 0 = target (scheduled first), 1 = main
    const int rank = (int)cluster_rank();
    const int B = g.B, S = g.S, A = g.A, NH = g.NH;
    const int row0 = rank * ROWS, row = row0 + r;
    const bool valid = row < B;
    const int nrows = min(max(B - row0, 0), ROWS);
    const rloa_naf_params& P = g.p[net];
    const uint8_t* image = g.images + (size_t)net * IMAGE_BYTES;

    const uint32_t raw = smem_u32(lc_smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* sm = lc_smem_raw + (base - raw);
    float* coef = reinterpret_cast<float*>(sm + COEF);               // [layer][sc | shb | rstd | xo][256]
    float* xbuf = reinterpret_cast<float*>(sm + XBUF);               // [2][3][256]
    float* vec = reinterpret_cast<float*>(sm + VEC);
    float* scratch = reinterpret_cast<float*>(sm + S1);              // [2][4][256] partial column sums + [256] + [3][256]
    float* colaux = scratch + 2 * 4 * 256;                           // per-column operand of the current statistics pass
    float* bcoef = scratch + 2 * 4 * 256 + 256;                      // k1 | k1 c1 | k1 c2
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + MISC);         // 0: W1 (+vec), 1: W2, 2: Wh, 3: MMA
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);
    float* red = reinterpret_cast<float*>(bars + 6);                 // [16] + AdamCoef at [20]
    uint32_t mma_phase = 0;

    if (warp == 0) tmem_alloc(tmem_slot, 512);
    if (tid == 0) {
        mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); mbar_init(&bars[2], 1); mbar_init(&bars[3], 1);
        mbar_init_fence();
        mbar_expect_tx(&bars[0], W1_BYTES + VEC_BYTES);
        bulk_g2s(sm + R1, image + OFF_W1, W1_BYTES, &bars[0]);
        bulk_g2s(sm + VEC, image + OFF_VEC, VEC_BYTES, &bars[0]);
        mbar_expect_tx(&bars[1], W2_BYTES);
        bulk_g2s(sm + R0, image + OFF_W2, W2_BYTES, &bars[1]);
    }
    // pending rows (see LearnClusterArgs): valid-row counts of the groups of 32 step rows, exclusive prefix, in the statistics
    // buffers (unused until the first BatchNorm pass): thread t owns group t
    uint32_t* pd_mask = reinterpret_cast<uint32_t*>(sm + COEF);       // [512] valid bits of group t
    int* pd_pre = reinterpret_cast<int*>(sm + COEF + 2048);           // [512] valid rows before group t
    int pd_total = 0;
    if (g.use_replay && g.pd_n > 0) {
        uint32_t m = 0;
        if (tid * 32 < g.pd_n) {
            if (g.pd_valid == nullptr) {
                const int left = g.pd_n - tid * 32;
                m = left >= 32 ? 0xffffffffu : ((1u << left) - 1u);
            } else {
                if (tid * 32 + 32 <= g.pd_n) {            // a whole group: two 16-byte loads (the mask is 16-byte aligned)
                    const uint4 v0 = *reinterpret_cast<const uint4*>(g.pd_valid + tid * 32);
                    const uint4 v1 = *reinterpret_cast<const uint4*>(g.pd_valid + tid * 32 + 16);
                    const uint32_t w[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
#pragma unroll
                    for (int k = 0; k < 8; k++)
#pragma unroll
                        for (int b = 0; b < 4; b++)
                            if (((w[k] >> (8 * b)) & 0xffu) != 0u) m |= 1u << (4 * k + b);
                } else {
                    for (int b = 0; tid * 32 + b < g.pd_n; b++)
                        if (g.pd_valid[tid * 32 + b] != 0) m |= 1u << b;
                }
            }
        }
        int inc = __popc(m);
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const int o = __shfl_up_sync(0xffffffffu, inc, off);
            if (lane >= off) inc += o;
        }
        int* wtot = reinterpret_cast<int*>(sm + COEF + 4096);         // [16] inclusive totals of the warps
        if (lane == 31) wtot[warp] = inc;
        __syncthreads();
        int wbase = 0;
        for (int w = 0; w < THREADS / 32; w++) {
            if (w < warp) wbase += wtot[w];
            pd_total += wtot[w];
        }
        pd_mask[tid] = m;
        pd_pre[tid] = wbase + inc - __popc(m);
        __syncthreads();
    }
    // this thread's 8 observation values (row r, k = 8 q .. 8 q + 7); kept in registers for the z1 recompute and dW1
    float xk[8];
    size_t src_row = (size_t)row;                                     // row of the batch arrays, or ring slot of the sample
    const float* in_actions = g.use_replay ? g.rb.actions : g.actions;
    const float* in_rewards = g.use_replay ? g.rb.rewards : g.rewards;
    const float* in_dones = g.use_replay ? g.rb.dones : g.dones;
    const float* x = g.use_replay ? (net == 0 ? g.rb.next_states : g.rb.states) : (net == 0 ? g.next_states : g.states);
    const uint8_t* in_dones_u8 = nullptr;                             // pending rows keep their done flags as bytes
    if (g.use_replay && valid) {
        const long long cur = *g.rb.cursor;
        src_row = replay_sample_slot_at(g.rb, cur + pd_total, row, g.rs_seed, g.rs_draw + (g.rs_draw_offset != nullptr ? *g.rs_draw_offset : 0ull));
        if (pd_total > 0) {
            const long long cap = g.rb.capacity;
            const long long off = ((long long)src_row - cur % cap + cap) % cap;      // distance from the first pending slot
            if (off < pd_total) {
                int lo = 0, hi = (g.pd_n + 31) / 32 - 1;                              // last group with pd_pre <= off
                while (lo < hi) {
                    const int mid = (lo + hi + 1) >> 1;
                    if (pd_pre[mid] <= (int)off) lo = mid; else hi = mid - 1;
                }
                src_row = (size_t)(lo * 32 + (int)__fns(pd_mask[lo], 0, (int)off - pd_pre[lo] + 1));
                in_actions = g.pd_actions; in_rewards = g.pd_rewards; in_dones = nullptr; in_dones_u8 = g.pd_dones;
                x = net == 0 ? g.pd_next_states : g.pd_states;
            }
        }
    }
    {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const int k = q * 8 + i;
            xk[i] = (k < S && valid) ? x[src_row * S + k] : 0.f;
        }
    }
    auto stage_x_tf32 = [&](uint8_t* xh, uint8_t* xl) {
        float h[8], l[8];
#pragma unroll
        for (int i = 0; i < 8; i++) { h[i] = to_tf32(xk[i]); l[i] = to_tf32(xk[i] - h[i]); }
#pragma unroll
        for (int c = 0; c < 2; c++) {
            const uint32_t off = sw128_chunk_offset(r, q * 2 + c);
            *reinterpret_cast<float4*>(xh + off) = make_float4(h[4 * c], h[4 * c + 1], h[4 * c + 2], h[4 * c + 3]);
            *reinterpret_cast<float4*>(xl + off) = make_float4(l[4 * c], l[4 * c + 1], l[4 * c + 2], l[4 * c + 3]);
        }
    };
    stage_x_tf32(sm + R1 + W1_BYTES, sm + R1 + W1_BYTES + 16384);
    fence_proxy_async();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem = *tmem_slot;
    const uint32_t tlane = tmem + ((uint32_t)(wq * 32) << 16);        // this warp's 32 TMEM lanes
    long long* prof = g.prof != nullptr ? g.prof + blockIdx.x * 32 : nullptr;
    auto stamp = [&](int i) {
        if (prof != nullptr && tid == 0) prof[i] = clock64();
    };
    stamp(0);

    auto mma_wait = [&]() {
        __syncwarp();
        mbar_wait_bounded(&bars[3], mma_phase);
        mma_phase ^= 1u;
        fence_after_sync();
    };
    auto issue_layer1 = [&](uint32_t w1_addr, uint32_t xh_addr, uint32_t xl_addr, uint32_t bar_phase) {
        if (tid == 0) {
            mbar_wait_bounded(&bars[0], bar_phase);
            fence_after_sync();
            constexpr uint32_t idesc = idesc_tf32_f32(128, H);
#pragma unroll
            for (int k = 0; k < 3; k++) {
                mma_tf32(tmem, umma_desc_sw128(xh_addr + k * 32), umma_desc_sw128(w1_addr + k * 32), idesc, k > 0);
                mma_tf32(tmem, umma_desc_sw128(xl_addr + k * 32), umma_desc_sw128(w1_addr + k * 32), idesc, true);
            }
            mma_commit(&bars[3]);
        }
    };

    LcCtx ctx;
    ctx.tlane = tlane; ctx.tid = tid; ctx.q = q; ctx.r = r; ctx.lane = lane; ctx.wq = wq; ctx.rank = rank; ctx.B = B;
    ctx.nrows = nrows; ctx.row = row; ctx.net = net; ctx.valid = valid;
    ctx.coef = coef; ctx.xbuf = xbuf; ctx.scratch = scratch; ctx.colaux = colaux; ctx.bcoef = bcoef; ctx.dbg = g.dbg;
    auto forward_stats = [&](int layer, uint32_t tcol, const float* bias, const float* bn_w, const float* bn_b,
                             float* run_mean, float* run_var, int64_t* batches) {
        lc_forward_stats(ctx, layer, tcol, bias, bn_w, bn_b, run_mean, run_var, batches);
    };
    auto dbg_rows32 = [&](int section, int width, int c0, const float (&a)[32]) { lc_dbg_rows32(ctx, section, width, c0, a); };
    auto dbg_preact = [&](int section, uint32_t tcol, const float* bias) {
        if (g.dbg != nullptr && net == 1) {
#pragma unroll 1
            for (int p = 0; p < 2; p++) {
                uint32_t v[32];
                float a[32];
                tmem_ld32(tlane + tcol + q * 64 + p * 32, v);
#pragma unroll
                for (int i = 0; i < 32; i++) a[i] = __uint_as_float(v[i]) + bias[q * 64 + p * 32 + i];
                dbg_rows32(section, 256, q * 64 + p * 32, a);
            }
        }
    };
    auto activation_to_tile = [&](int layer, uint32_t tcol, uint8_t* tile) { lc_activation_to_tile(ctx, layer, tcol, tile); };

    // =========================== forward ===========================
    issue_layer1(base + R1, base + R1 + W1_BYTES, base + R1 + W1_BYTES + 16384, 0);
    mbar_wait_bounded(&bars[0], 0);                       // biases, for every thread
    mma_wait();
    stamp(1);
    forward_stats(0, 0, vec, P.bn1_w, P.bn1_b, P.bn1_mean, P.bn1_var, P.bn1_batches);
    stamp(2);
    dbg_preact(0, 0, vec);
    activation_to_tile(0, 0, sm + R1);                    // a1 over the dead layer-1 operands
    fence_proxy_async();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    if (tid == 0) {                                       // z2 = a1 W2^T -> TMEM columns 256-511
        mbar_wait_bounded(&bars[1], 0);
        fence_after_sync();
        constexpr uint32_t idesc = idesc_bf16_f32(128, H);
#pragma unroll
        for (int k = 0; k < 16; k++)
            mma_bf16(tmem + 256, umma_desc_sw128(base + R1 + (k >> 2) * A_BLK + (k & 3) * 32),
                     umma_desc_sw128(base + R0 + (k >> 2) * (H * 128) + (k & 3) * 32), idesc, k > 0);
        mma_commit(&bars[3]);
    }
    mma_wait();
    if (tid == 0) {                                       // the W2 slot is free: head image -> its first 32 KB
        mbar_expect_tx(&bars[2], WH_BYTES);
        bulk_g2s(sm + R0, image + OFF_WH, WH_BYTES, &bars[2]);
    }
    stamp(3);
    forward_stats(1, 256, vec + H, P.bn2_w, P.bn2_b, P.bn2_mean, P.bn2_var, P.bn2_batches);
    stamp(4);
    dbg_preact(1, 256, vec + H);
    activation_to_tile(1, 256, sm + R1);                  // a2
    fence_proxy_async();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    if (tid == 0) {                                       // heads = a2 Wh^T -> TMEM columns 0-63
        mbar_wait_bounded(&bars[2], 0);
        fence_after_sync();
        constexpr uint32_t idesc = idesc_bf16_f32(128, NHP);
#pragma unroll
        for (int k = 0; k < 16; k++)
            mma_bf16(tmem, umma_desc_sw128(base + R1 + (k >> 2) * A_BLK + (k & 3) * 32),
                     umma_desc_sw128(base + R0 + (k >> 2) * (NHP * 128) + (k & 3) * 32), idesc, k > 0);
        mma_commit(&bars[3]);
    }
    mma_wait();

    // =========================== head ===========================
    stamp(5);
    // every thread moves 16 of its row's 64 head values TMEM -> shared (zo) and clears the same 16 slots of the gradient row
    float* zo = reinterpret_cast<float*>(sm + HEAD_SCRATCH) + r * 65;
    float* dzr = reinterpret_cast<float*>(sm + HEAD_SCRATCH + 33280) + r * 65;
    float* hpart = reinterpret_cast<float*>(sm + HEAD_SCRATCH + 2 * 33280);     // [128][4] advantage partials, then [8][64]
    const float* bh = vec + 2 * H;
    {
        uint32_t v[32];
        tmem_ld32(tlane + (q >> 1) * 32, v);               // two threads share a 32-column load, each keeps 16
#pragma unroll
        for (int i = 0; i < 16; i++) {
            const int o = q * 16 + i;
            zo[o] = __uint_as_float((q & 1) ? v[16 + i] : v[i]) + bh[o];
            dzr[o] = 0.f;
        }
    }
    __syncthreads();
    if (net == 0) {
        // ---- target network: y = r + gamma V'(s') for this CTA's rows, then hand over to the main CTA of the same rank
        if (q == 0 && valid) {
            float vv = zo[A];
            if (g.hp.use_done_mask && in_dones != nullptr) vv *= (1.f - in_dones[src_row]);
            if (g.hp.use_done_mask && in_dones_u8 != nullptr) vv *= (1.f - (float)in_dones_u8[src_row]);
            g.y[row] = fmaf(g.hp.gamma, vv, in_rewards[src_row]);
        }
        __threadfence();
        __syncthreads();
        if (tid == 0) {
            unsigned one = 1u;
            asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(g.yflag + rank), "r"(one) : "memory");
        }
        if (g.do_adam) {
            // the optimiser tail is shared: this CTA stays for the second half of main CTA `rank`'s elements
            if (tid == 0) {
                unsigned f = 0u;
                const long long t0 = clock64();
                do {
                    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(f) : "l"(g.aflag + rank) : "memory");
                    // the main cluster's backward pass lies between (60 us; seconds under compute-sanitizer): a generous bound,
                    // its own waits abort within 2 s if anything is stuck
                    if (f == 0u && clock64() - t0 > 120000000000ll) __trap();
                } while (f == 0u);
                g.aflag[rank] = 0u;                              // consumed: zero at rest
            }
            __syncthreads();
            stamp(20);
            const float* ac = g.acoef + rank * 8;
            AdamCoef c;
            c.clip = __ldcg(ac); c.bc1 = __ldcg(ac + 1); c.bc2s = __ldcg(ac + 2); c.norm = __ldcg(ac + 3);
            if (__ldcg(ac + 4) == 0.f) lc_adam_apply(g, c, rank, tid, 2, 4, 3, 6, nullptr, nullptr, nullptr);
        }
    } else {
        // ---- main network: TD error, loss, gradients of the head pre-activations; thread (r, q) owns actions k = q, q + 4, q + 8
        float mu_[3], t_[3], P_[3], df_[3];
        float advp = 0.f;
#pragma unroll
        for (int u = 0; u < 3; u++) {
            const int k = q + 4 * u;
            mu_[u] = t_[u] = P_[u] = df_[u] = 0.f;
            if (k < A && valid) {
                mu_[u] = tanhf(zo[k]);
                t_[u] = tanhf(zo[A + 1 + (k * (k + 3)) / 2]);
                P_[u] = expf(2.f * t_[u]);
                float act = in_actions[src_row * A + k];
                if (g.hp.trunc_action) act = truncf(act);
                df_[u] = act - mu_[u];
                advp = fmaf(-0.5f * P_[u] * df_[u], df_[u], advp);
            }
        }
        hpart[r * 4 + q] = advp;
        if (tid == 0) {
            unsigned f = 0u;
            const long long t0 = clock64();
            do {
                asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(f) : "l"(g.yflag + rank) : "memory");
                if (f == 0u && clock64() - t0 > 4000000000ll) __trap();      // the target cluster never ran: abort, do not hang
            } while (f == 0u);
        }
        __syncthreads();
        stamp(6);
        float err2 = 0.f;
        if (valid) {
            const float adv = (hpart[r * 4] + hpart[r * 4 + 1]) + (hpart[r * 4 + 2] + hpart[r * 4 + 3]);
            const float err = (zo[A] + adv) - __ldcg(g.y + row);
            const float dq = 2.f * err / (float)B;                 // MSE over the batch (naf_algorithm.py:204)
            if (q == 0) {
                err2 = err * err;
                dzr[A] = dq;
            }
#pragma unroll
            for (int u = 0; u < 3; u++) {
                const int k = q + 4 * u;
                if (k < A) {
                    dzr[k] = dq * P_[u] * df_[u] * (1.f - mu_[u] * mu_[u]);
                    dzr[A + 1 + (k * (k + 3)) / 2] = dq * (-P_[u] * df_[u] * df_[u]) * (1.f - t_[u] * t_[u]);
                }
            }
        }
        __syncthreads();
        {   // bf16 tile [128 rows][64 columns] (one k-block) for dWh^T and da2: 16 columns = 2 chunks per thread
#pragma unroll
            for (int c = 0; c < 2; c++) {
                float x[8];
#pragma unroll
                for (int i = 0; i < 8; i++) x[i] = dzr[q * 16 + c * 8 + i];
                *reinterpret_cast<uint4*>(sm + S1 + sw128_chunk_offset(r, q * 2 + c)) = pack8_bf16(x);
            }
            if (g.dbg != nullptr && valid) {
                float* d = g.dbg + (size_t)2 * 1024 * 256 + (size_t)row * 64 + q * 16;
#pragma unroll
                for (int i = 0; i < 16; i++) d[i] = dzr[q * 16 + i];
            }
        }
        // loss partial of this CTA (fixed order: lanes, then warps 0-3); head bias gradients: column sums of dzh over the rows
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) err2 += __shfl_xor_sync(0xffffffffu, err2, off);
        if (q == 0 && lane == 0) red[wq] = err2;
        {
            const int o = tid & 63, part = tid >> 6;                // 8 parts of 16 rows
            const float* ds = reinterpret_cast<const float*>(sm + HEAD_SCRATCH + 33280);
            float sacc = 0.f;
#pragma unroll
            for (int rr = 0; rr < 16; rr++) sacc += ds[(part * 16 + rr) * 65 + o];
            __syncthreads();                                        // the advantage partials in hpart are dead
            hpart[part * 64 + o] = sacc;
        }
        fence_proxy_async();
        __syncthreads();
        if (tid == 0) {
            g.part_loss[rank] = (red[0] + red[1]) + (red[2] + red[3]);
            g.yflag[rank] = 0u;                                  // consumed: zero at rest
        }
        if (tid < 64) {
            float sacc = 0.f;
#pragma unroll
            for (int part = 0; part < 8; part++) sacc += hpart[part * 64 + tid];
            g.part_hb[rank * 64 + tid] = sacc;
        }
        fence_proxy_async();     // the generic accesses of the head scratch precede the TMA write that reuses its slot
        fence_before_sync();
        __syncthreads();
        fence_after_sync();

        // =========================== backward ===========================
        stamp(7);
        if (tid == 0) {          // dWh^T [256 f][64 o] = a2^T dzh: both tiles MN-major, contraction over the 128 rows
            constexpr uint32_t idesc = idesc_bf16_f32(128, NHP) | kIdescAMajorMN | kIdescBMajorMN;
#pragma unroll
            for (int half = 0; half < 2; half++)
#pragma unroll
                for (int k = 0; k < 8; k++)
                    mma_bf16(tmem + 64 + half * 64, umma_desc_sw128_mn(base + R1 + half * 2 * A_BLK + k * 2048, A_BLK),
                             umma_desc_sw128_mn(base + S1 + k * 2048, A_BLK), idesc, k > 0);
            mma_commit(&bars[3]);
        }
        mma_wait();
        if (q < 2) {             // partial head-weight gradient of this CTA: part_wh[rank][o][f], f = 128 q + r
            float* dst = g.part_wh + (size_t)rank * NH * H + (q * 128 + r);
#pragma unroll 1
            for (int p = 0; p < 2; p++) {
                uint32_t v[32];
                tmem_ld32(tlane + 64 + q * 64 + p * 32, v);
#pragma unroll
                for (int i = 0; i < 32; i++)
                    if (p * 32 + i < NH) dst[(size_t)(p * 32 + i) * H] = __uint_as_float(v[i]);
            }
        }
        fence_before_sync();
        __syncthreads();
        fence_after_sync();
        stamp(8);
        if (tid == 0) {          // da2 [128][256] = dzh Wh: B = the head image read MN-major (contraction over its 64 rows)
            constexpr uint32_t idesc = idesc_bf16_f32(128, H) | kIdescBMajorMN;
#pragma unroll
            for (int k = 0; k < 4; k++)
                mma_bf16(tmem, umma_desc_sw128(base + S1 + k * 32), umma_desc_sw128_mn(base + R0 + k * 2048, NHP * 128), idesc, k > 0);
            mma_commit(&bars[3]);
        }
        mma_wait();
        if (tid == 0) {          // head image and scratch are dead: W2 image back into its slot for da1
            mbar_expect_tx(&bars[1], W2_BYTES);
            bulk_g2s(sm + R0, image + OFF_W2, W2_BYTES, &bars[1]);
        }

        stamp(9);
        lc_backward_stats(ctx, 1, 0, 256, P.bn2_w, 0, g.grad + g.off_bn2w, g.grad + g.off_bn2b, g.grad + g.off_b2);
        stamp(10);
        lc_dz_to_tile(ctx, 1, 0, 256, sm + R1, nullptr, 3, -1);      // dz2 -> bf16 tile in R1 (a2 is dead: dWh^T finished)
        fence_proxy_async();
        fence_before_sync();
        __syncthreads();
        fence_after_sync();
        stamp(11);
        if (tid == 0) {          // da1 [128][256] = dz2 W2: B = the forward W2 image read MN-major -> TMEM columns 256-511
            mbar_wait_bounded(&bars[1], 1);
            fence_after_sync();
            constexpr uint32_t idesc = idesc_bf16_f32(128, H) | kIdescBMajorMN;
#pragma unroll
            for (int k = 0; k < 16; k++)
                mma_bf16(tmem + 256, umma_desc_sw128(base + R1 + (k >> 2) * A_BLK + (k & 3) * 32),
                         umma_desc_sw128_mn(base + R0 + k * 2048, H * 128), idesc, k > 0);
            mma_commit(&bars[3]);
        }
        mma_wait();
        if (tid == 0) {          // W2 is dead: layer-1 operands back (W1 image by TMA, observations from registers)
            mbar_expect_tx(&bars[0], W1_BYTES);
            bulk_g2s(sm + R0, image + OFF_W1, W1_BYTES, &bars[0]);
        }
        stage_x_tf32(sm + R0 + W1_BYTES, sm + R0 + W1_BYTES + 16384);
        fence_proxy_async();
        fence_before_sync();
        __syncthreads();
        fence_after_sync();
        stamp(12);
        issue_layer1(base + R0, base + R0 + W1_BYTES, base + R0 + W1_BYTES + 16384, 1);      // z1 again -> TMEM columns 0-255
        mma_wait();
        stamp(13);
        lc_backward_stats(ctx, 0, 256, 0, P.bn1_w, 1, g.grad + g.off_bn1w, g.grad + g.off_bn1b, g.grad + g.off_b1);
        stamp(14);
        // dz1 -> R0 + 64 KB, a1 -> R0 (the layer-1 operands there are dead: z1 is in TMEM)
        lc_dz_to_tile(ctx, 0, 256, 0, sm + R0 + 65536, sm + R0, 5, 4);
        __syncthreads();         // the backward coefficients in S1 are dead: the observation tile takes their place
        {   // [x_hi | x_lo] as bf16, 64 columns: hi(k) at column k, lo(k) at column 32 + k
            float h[8], l[8];
#pragma unroll
            for (int i = 0; i < 8; i++) {
                h[i] = __bfloat162float(__float2bfloat16_rn(xk[i]));
                l[i] = xk[i] - h[i];
            }
            *reinterpret_cast<uint4*>(sm + S1 + sw128_chunk_offset(r, q)) = pack8_bf16(h);
            *reinterpret_cast<uint4*>(sm + S1 + sw128_chunk_offset(r, 4 + q)) = pack8_bf16(l);
        }
        fence_proxy_async();
        fence_before_sync();
        __syncthreads();
        fence_after_sync();
        stamp(15);
        if (tid == 0) {          // dW1 [256 h][hi 32 | lo 32] = dz1^T [x_hi | x_lo] -> TMEM columns 0-127
            constexpr uint32_t idesc = idesc_bf16_f32(128, 64) | kIdescAMajorMN | kIdescBMajorMN;
#pragma unroll
            for (int half = 0; half < 2; half++)
#pragma unroll
                for (int k = 0; k < 8; k++)
                    mma_bf16(tmem + half * 64, umma_desc_sw128_mn(base + R0 + 65536 + half * 2 * A_BLK + k * 2048, A_BLK),
                             umma_desc_sw128_mn(base + S1 + k * 2048, A_BLK), idesc, k > 0);
            mma_commit(&bars[3]);
        }
        mma_wait();
        if (q < 2) {
            float* dst = g.part_w1 + (size_t)rank * H * S + (size_t)(q * 128 + r) * S;
            uint32_t v0[32], v1[32];
            tmem_ld32_nowait(tlane + q * 64, v0);
            tmem_ld32_nowait(tlane + q * 64 + 32, v1);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 24; i++)
                if (i < S) dst[i] = __uint_as_float(v0[i]) + __uint_as_float(v1[i]);
        }
        fence_before_sync();
        __syncthreads();
        fence_after_sync();
        stamp(16);
        if (tid == 0) {          // dW2 [256 n][256 k] = dz2^T a1 -> all 512 TMEM columns (n half h in columns 256 h ..)
            constexpr uint32_t idesc = idesc_bf16_f32(128, H) | kIdescAMajorMN | kIdescBMajorMN;
#pragma unroll
            for (int half = 0; half < 2; half++)
#pragma unroll
                for (int k = 0; k < 8; k++)
                    mma_bf16(tmem + half * 256, umma_desc_sw128_mn(base + R1 + half * 2 * A_BLK + k * 2048, A_BLK),
                             umma_desc_sw128_mn(base + R0 + k * 2048, A_BLK), idesc, k > 0);
            mma_commit(&bars[3]);
        }
        mma_wait();
        {   // partial dW2 of this CTA, staged through shared memory (every operand tile is dead) so that the global stores
            // are 512 contiguous bytes per warp instruction: fp32 [128 rows][260] per half (pitch 260: conflict-free STS.128)
            float* stg = reinterpret_cast<float*>(sm + R1);
            float* dst = g.part_w2 + (size_t)rank * H * H;
#pragma unroll 1
            for (int half = 0; half < 2; half++) {
#pragma unroll 1
                for (int p = 0; p < 2; p++) {
                    const int c0 = q * 64 + p * 32;
                    uint32_t v[32];
                    tmem_ld32(tlane + half * 256 + c0, v);
                    float4* o = reinterpret_cast<float4*>(stg + r * 260 + c0);
#pragma unroll
                    for (int j = 0; j < 8; j++)
                        o[j] = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]),
                                           __uint_as_float(v[4 * j + 3]));
                }
                __syncthreads();
#pragma unroll
                for (int rr = 0; rr < 8; rr++) {           // warp w stores rows w, w + 16, ...: 2 x 512 B per row
                    const int srow = warp + 16 * rr;
                    const float4 lo = *reinterpret_cast<const float4*>(stg + srow * 260 + lane * 4);
                    const float4 hi = *reinterpret_cast<const float4*>(stg + srow * 260 + 128 + lane * 4);
                    float4* o = reinterpret_cast<float4*>(dst + (size_t)(half * 128 + srow) * H);
                    o[lane] = lo;
                    o[32 + lane] = hi;
                }
                __syncthreads();
            }
        }

        // =========================== optimiser tail ===========================
        stamp(17);
        __threadfence();
        cluster_sync_all();                                    // every CTA's partials (and rank 0's totals) are visible
        stamp(18);
        const int n = g.n_params;
        const float gs = g.hp.grad_scale;
        // The tail's loops are rolled (their bodies are fetched once and reused) and hand their per-thread results from the
        // gather pass to the optimiser pass through shared memory — every operand tile is dead by now.
        float4* t_gw = reinterpret_cast<float4*>(sm + R1);                    // [4][512] summed dW2 elements of this thread
        float* t_gr = reinterpret_cast<float*>(sm + R1 + 32768);              // [6][512] summed other elements
        int* t_gi = reinterpret_cast<int*>(sm + R1 + 32768 + 12288);          // [6][512] their flat indices (-1: none)
        // N > 1: local sums go to this rank's exchange block (peers read them); before the first write every peer must have
        // finished reading the block for the previous update (normally long true)
        const bool xch = g.use_xchg != 0;
        const int me = g.xp.rank, world = g.xp.world;
        const unsigned long long tstep = xch ? (unsigned long long)(*g.step) + 1ull : 0ull;
        __shared__ int s_xbad;
        // N > 1: reduce-scatter + all-gather over NVLink in tagged words (grad_exchange.cuh), per CTA slice.  Every element has
        // an OWNER rank (contiguous index ranges of the slice).  Round 1: this rank's local sums are STORED into the owner's
        // block (posted NVLink writes); the owner adds the `world` contributions in rank order - one result, so every rank
        // applies the same bits.  Round 2: the owner stores the sum into every rank's block; all ranks read the reduced
        // gradient from their own.  A rank moves 2 (world - 1) / world of a gradient through NVLink, and nothing waits for a
        // fence or a flag: the receiver polls the data words for the update's tag.
        // the tag counts this handle's exchanges, not optimiser steps: a restored training state may rewind the step counter,
        // and a repeated number would let last run's words pass for this update's
        const unsigned tag = xch ? (unsigned)g.xstatus[1] + 1u : 0u;
        const size_t np = (size_t)g.n_params;
        auto ll1 = [&](int owner, int src) {      // round-1 region `src` in rank `owner`'s block
            return reinterpret_cast<unsigned long long*>(const_cast<float*>(g.xp.grad[owner]) + xchg_ll1_offset(g.n_params)) + (size_t)src * np;
        };
        auto ll2 = [&](int dst) {
            return reinterpret_cast<unsigned long long*>(const_cast<float*>(g.xp.grad[dst]) + xchg_ll2_offset(g.n_params));
        };
        bool xbad = false;
        const long long xt0 = clock64();
        float sq = 0.f;
        {   // (a) this CTA's eighth of dW2 (16384 float4 in all, 2048 per CTA, 4 per thread): two rounds of 16 loads in flight
            const float4* pw = reinterpret_cast<const float4*>(g.part_w2);
#pragma unroll 1
            for (int jj = 0; jj < 2; jj++) {
                float4 pv[2][CL];
#pragma unroll
                for (int u = 0; u < 2; u++)
#pragma unroll
                    for (int c = 0; c < CL; c++) pv[u][c] = __ldcg(pw + (size_t)c * (H * H / 4) + rank * 2048 + (2 * jj + u) * 512 + tid);
#pragma unroll
                for (int u = 0; u < 2; u++) {
                    float4 s4 = pv[u][0];
#pragma unroll
                    for (int c = 1; c < CL; c++) { s4.x += pv[u][c].x; s4.y += pv[u][c].y; s4.z += pv[u][c].z; s4.w += pv[u][c].w; }
                    t_gw[(2 * jj + u) * THREADS + tid] = s4;
                    if (xch) {
                        const int idx = (2 * jj + u) * 512 + tid;
                        const int owner = (idx * world) >> 11;
                        if (owner != me) st_ll4(ll1(owner, me) + g.off_w2 + 4 * (rank * 2048 + idx), s4, tag);
                    } else {
                        reinterpret_cast<float4*>(g.grad + g.off_w2)[rank * 2048 + (2 * jj + u) * 512 + tid] = s4;
                    }
                    const float x0 = s4.x * gs, x1 = s4.y * gs, x2 = s4.z * gs, x3 = s4.w * gs;
                    sq = fmaf(x0, x0, sq); sq = fmaf(x1, x1, sq); sq = fmaf(x2, x2, sq); sq = fmaf(x3, x3, sq);
                }
            }
        }
        // (b) the remaining n - 65536 elements, compacted (W2 skipped): element e of that list, CTA c takes e in [c m, (c+1) m)
        const int rest = n - H * H, per = (rest + CL - 1) / CL;
        constexpr int kSlots = 6;                              // per <= 2727 (A = 9, S = 24) < 6 x 512
#pragma unroll 1
        for (int grp = 0; grp < 2; grp++) {                    // three elements x 8 partials in flight per round
            float pvs[3][CL];
            int gi3[3];
#pragma unroll
            for (int u = 0; u < 3; u++) {
                const int j = 3 * grp + u;
                const int e = rank * per + j * THREADS + tid;
                gi3[u] = -1;
#pragma unroll
                for (int c = 0; c < CL; c++) pvs[u][c] = 0.f;
                if (j * THREADS + tid < per && e < rest) {
                    const int i = e < g.off_w2 ? e : e + H * H;               // flat index
                    gi3[u] = i;
                    if (i < g.off_b1) {                                       // dW1: 8 partials
#pragma unroll
                        for (int c = 0; c < CL; c++) pvs[u][c] = __ldcg(g.part_w1 + (size_t)c * H * S + i);
                    } else if (i < g.off_wmu) {                               // b1, bn1, b2, bn2: totals left by rank 0
                        pvs[u][0] = __ldcg(g.grad + i);
                    } else {                                                  // heads
                        int o = -1, f = 0;
                        if (i < g.off_bmu) { o = (i - g.off_wmu) / H; f = (i - g.off_wmu) - o * H; }
                        else if (i < g.off_wv) { o = i - g.off_bmu; f = -1; }
                        else if (i < g.off_bv) { o = A; f = i - g.off_wv; }
                        else if (i < g.off_wl) { o = A; f = -1; }
                        else if (i < g.off_bl) { const int oo = (i - g.off_wl) / H; o = A + 1 + oo; f = (i - g.off_wl) - oo * H; }
                        else { o = A + 1 + (i - g.off_bl); f = -1; }
#pragma unroll
                        for (int c = 0; c < CL; c++)
                            pvs[u][c] = f >= 0 ? __ldcg(g.part_wh + (size_t)c * NH * H + (size_t)o * H + f) : __ldcg(g.part_hb + c * 64 + o);
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < 3; u++) {
                const int j = 3 * grp + u;
                float ssum = 0.f;
#pragma unroll
                for (int c = 0; c < CL; c++) ssum += pvs[u][c];
                t_gr[j * THREADS + tid] = ssum;
                t_gi[j * THREADS + tid] = gi3[u];
                if (gi3[u] >= 0) {
                    if (xch) {
                        const int owner = ((j * THREADS + tid) * world) / (kSlots * THREADS);
                        if (owner != me) st_ll(ll1(owner, me) + gi3[u], ssum, tag);
                    } else {
                        g.grad[gi3[u]] = ssum;
                    }
                    const float x = ssum * gs;
                    sq = fmaf(x, x, sq);
                }
            }
        }
        if (xch) {
            __syncthreads();                                   // t_gw / t_gr / t_gi of this thread's elements are its own: no hazard;
                                                               // the barrier only keeps the phases of the timeline apart
            // round 1 (owner): the peers' words of the owned elements, polled; sum in rank order; round 2: the sum to everyone
#pragma unroll 1
            for (int j = 0; j < 4; j++) {
                const int idx = j * THREADS + tid;
                if (((idx * world) >> 11) != me) continue;
                const size_t e = (size_t)g.off_w2 + 4 * (size_t)(rank * 2048 + idx);
                float4 s4 = make_float4(0.f, 0.f, 0.f, 0.f);
                for (int rr = 0; rr < world; rr++) {
                    float4 v = t_gw[j * THREADS + tid];
                    if (rr != me)
                        while (!ld_ll4(ll1(me, rr) + e, tag, v))
                            if (clock64() - xt0 > kXchgTimeoutCycles) { xbad = true; break; }
                    s4.x += v.x; s4.y += v.y; s4.z += v.z; s4.w += v.w;
                }
                for (int rr = 0; rr < world; rr++) st_ll4(ll2(rr) + e, s4, tag);
            }
#pragma unroll 1
            for (int j = 0; j < kSlots; j++) {
                const int i = t_gi[j * THREADS + tid];
                if (i < 0 || ((j * THREADS + tid) * world) / (kSlots * THREADS) != me) continue;
                float ssum = 0.f;
                for (int rr = 0; rr < world; rr++) {
                    float v = t_gr[j * THREADS + tid];
                    if (rr != me)
                        while (!ld_ll(ll1(me, rr) + i, tag, v))
                            if (clock64() - xt0 > kXchgTimeoutCycles) { xbad = true; break; }
                    ssum += v;
                }
                for (int rr = 0; rr < world; rr++) st_ll(ll2(rr) + i, ssum, tag);
            }
            // round 2 (everyone): the reduced elements of this thread from its own block
            sq = 0.f;
            const unsigned long long* l2 = ll2(me);
#pragma unroll 1
            for (int j = 0; j < 4; j++) {
                float4 s4;
                while (!ld_ll4(l2 + g.off_w2 + 4 * (size_t)(rank * 2048 + j * THREADS + tid), tag, s4))
                    if (clock64() - xt0 > kXchgTimeoutCycles) { xbad = true; break; }
                t_gw[j * THREADS + tid] = s4;
                reinterpret_cast<float4*>(g.grad + g.off_w2)[rank * 2048 + j * THREADS + tid] = s4;
                const float x0 = s4.x * gs, x1 = s4.y * gs, x2 = s4.z * gs, x3 = s4.w * gs;
                sq = fmaf(x0, x0, sq); sq = fmaf(x1, x1, sq); sq = fmaf(x2, x2, sq); sq = fmaf(x3, x3, sq);
            }
#pragma unroll 1
            for (int j = 0; j < kSlots; j++) {
                const int i = t_gi[j * THREADS + tid];
                if (i < 0) continue;
                float v;
                while (!ld_ll(l2 + i, tag, v))
                    if (clock64() - xt0 > kXchgTimeoutCycles) { xbad = true; break; }
                t_gr[j * THREADS + tid] = v;
                g.grad[i] = v;
                const float x = v * gs;
                sq = fmaf(x, x, sq);
            }
            if (xbad) atomicExch(g.xstatus, 1);
            // keep the flag protocol of the whole-gradient kernels (slice 0) in step, should the caller switch paths
            if (rank == 0 && tid < world && tid != me) {
                st_sys(g.xp.ready[tid] + me, tstep);
                st_sys(g.xp.done[tid] + me, tstep);
            }
            __syncthreads();
        }
        stamp(19);
        // squared norm: warp tree, warps in order, CTAs in rank order over DSMEM
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, off);
        if (lane == 0) red[warp] = sq;
        __syncthreads();
        float* normslot = xbuf;                                 // DSMEM-visible slot (the statistics buffers are dead)
        if (tid == 0) {
            float t = 0.f;
            for (int w = 0; w < THREADS / 32; w++) t += red[w];
            normslot[0] = t;
        }
        cluster_sync_all();
        stamp(20);
        if (g.do_adam) {
            if (tid == 0) {
                float t = 0.f;
                for (int c = 0; c < CL; c++) t += ld_dsmem(normslot, c);
                const long long step = (long long)(*g.step) + 1;            // optimizer.step() count of this update
                AdamCoef c;
                c.norm = sqrtf(t);
                c.clip = fminf(g.hp.clip_norm / (c.norm + 1e-6f), 1.f) * g.hp.grad_scale;
                c.bc1 = 1.f - powf(g.hp.beta1, (float)step);
                c.bc2s = sqrtf(1.f - powf(g.hp.beta2, (float)step));
                *reinterpret_cast<AdamCoef*>(red + 20) = c;
                if (rank == 0) {
                    if (g.gnorm != nullptr) *g.gnorm = c.norm;
                    if (g.loss != nullptr) {
                        float l = 0.f;
                        for (int cc = 0; cc < CL; cc++) l += __ldcg(g.part_loss + cc);
                        *g.loss = l / (float)B;
                    }
                }
            }
            if (tid == 0) {
                s_xbad = xch ? *reinterpret_cast<volatile int*>(g.xstatus) : 0;
                // hand the second half of this CTA's elements to the target-cluster CTA of the same rank (idle since its
                // forward pass): coefficients + "skip" behind a release flag; the gradients it needs are in g.grad
                const AdamCoef cc = *reinterpret_cast<const AdamCoef*>(red + 20);
                float* ac = g.acoef + rank * 8;
                ac[0] = cc.clip; ac[1] = cc.bc1; ac[2] = cc.bc2s; ac[3] = cc.norm; ac[4] = s_xbad ? 1.f : 0.f;
                __threadfence();
                unsigned one = 1u;
                asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(g.aflag + rank), "r"(one) : "memory");
            }
            __syncthreads();
            const AdamCoef c = *reinterpret_cast<const AdamCoef*>(red + 20);
            // a wait that gave up (now or in an earlier update: the flag is sticky) means a peer's gradient may be stale:
            // nobody applies the update, so the ranks stop instead of drifting apart; rloa_xchg_status reports it
            if (!s_xbad) lc_adam_apply(g, c, rank, tid, 0, 2, 0, 3, t_gw, t_gr, t_gi);
        } else if (rank == 0 && tid == 0 && g.loss != nullptr) {
            float l = 0.f;
            for (int cc = 0; cc < CL; cc++) l += __ldcg(g.part_loss + cc);
            *g.loss = l / (float)B;
        }
    }
    stamp(21);
    // a CTA's shared memory must outlive every DSMEM read of it; the step counter moves after everybody has read it
    cluster_sync_all();
    if (net == 1 && g.do_adam && rank == 0 && tid == 0) *g.step += 1;      // also after a timed-out exchange: the flags count updates
    if (net == 1 && g.use_xchg && rank == 0 && tid == 0) g.xstatus[1] += 1;
    fence_before_sync();
    __syncthreads();
    stamp(22);
    if (warp == 0) tmem_dealloc(tmem, 512);
}

// ---------------------------------------------------------------------------------------------------
bool learn_cluster_supported(int S, int A, int H, int B) {
    const int NH = A + 1 + A * (A + 1) / 2;
    return H == lc::H && S >= 1 && S <= 24 && A >= 1 && A <= 9 && NH <= lc::NHP && B >= 2 && B <= lc::CL * lc::ROWS;
}

int learn_cluster_prepare(LearnCluster* lcw, int S, int A) {
    if (lcw->images != nullptr) return RLOA_OK;
    const int NH = A + 1 + A * (A + 1) / 2;
    RLOA_CUDA(cudaFuncSetAttribute(naf_learn_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lc::SMEM_BYTES));
    const size_t floats = 1024 + (size_t)lc::CL * (lc::H * lc::H + lc::H * S + (size_t)NH * lc::H + 64 + 1) + 16 + 8 + 64;
    RLOA_CUDA(cudaMalloc(&lcw->images, learn_cluster_image_bytes()));
    if (cudaMalloc(&lcw->block, floats * sizeof(float)) != cudaSuccess) {
        cudaFree(lcw->images);
        lcw->images = nullptr;
        set_error("learn cluster: cudaMalloc failed: %s", cudaGetErrorString(cudaGetLastError()));
        return RLOA_ERR_CUDA;
    }
    RLOA_CUDA(cudaMemset(lcw->block, 0, floats * sizeof(float)));
    float* p = lcw->block;
    lcw->y = p; p += 1024;
    lcw->part_w2 = p; p += (size_t)lc::CL * lc::H * lc::H;
    lcw->part_w1 = p; p += (size_t)lc::CL * lc::H * S;
    lcw->part_wh = p; p += (size_t)lc::CL * NH * lc::H;
    lcw->part_hb = p; p += lc::CL * 64;
    lcw->part_loss = p; p += lc::CL;
    lcw->yflag = reinterpret_cast<unsigned*>(p); p += 16;
    lcw->aflag = reinterpret_cast<unsigned*>(p); p += 8;
    lcw->acoef = p;
    RLOA_CUDA(cudaEventCreateWithFlags(&lcw->pack_done, cudaEventDisableTiming));
    RLOA_CUDA(cudaEventCreateWithFlags(&lcw->pack_fork, cudaEventDisableTiming));
    return RLOA_OK;
}

void learn_cluster_free(LearnCluster* lcw) {
    if (lcw->pack_done) cudaEventDestroy(lcw->pack_done);
    if (lcw->pack_fork) cudaEventDestroy(lcw->pack_fork);
    if (lcw->images) cudaFree(lcw->images);
    if (lcw->block) cudaFree(lcw->block);
    *lcw = LearnCluster{};
}

int learn_cluster_pack(LearnCluster* lcw, const rloa_naf_params* mn, const rloa_naf_params* tg, cudaStream_t st) {
    RLOA_REQUIRE(lcw->images != nullptr, "learn cluster: not prepared");
    const int pack_threads = lc::H * lc::H / 8 + lc::NHP * lc::H / 8 + lc::H * lc::KP / 4 + lc::H;
    learn_pack_kernel<<<dim3((pack_threads + 255) / 256, 2), 256, 0, st>>>(*tg, *mn, static_cast<uint8_t*>(lcw->images));
    RLOA_LAUNCHED();
    return RLOA_OK;
}

int learn_cluster_step(LearnCluster* lcw, const rloa_naf_params* mn, const rloa_naf_params* tg, const rloa_adam_state* adam,
                       const float* states, const float* actions, const float* rewards, const float* next_states,
                       const float* dones, int B, const rloa_naf_hyper* hp, const ParamTable& pt, const int* flat_offsets,
                       float* grad, float* loss, float* gnorm, int do_adam, const LearnClusterReplay* replay, const rloa_xchg* xchg,
                       cudaStream_t st) {
    RLOA_REQUIRE(lcw->images != nullptr, "learn cluster: not prepared");
    const int S = mn->state_size, A = mn->action_size;
    if (lcw->prepacked) {            // rloa_naf_learn_prepack already wrote the images on the side stream: join it
        RLOA_CUDA(cudaStreamWaitEvent(st, lcw->pack_done, 0));
        lcw->prepacked = false;
    } else {
        const int rc = learn_cluster_pack(lcw, mn, tg, st);
        if (rc != RLOA_OK) return rc;
    }
    LearnClusterArgs a{};
    a.images = static_cast<const uint8_t*>(lcw->images);
    a.p[0] = *tg; a.p[1] = *mn;
    a.states = states; a.actions = actions; a.rewards = rewards; a.next_states = next_states; a.dones = dones;
    a.B = B; a.S = S; a.A = A; a.NL = A * (A + 1) / 2; a.NH = A + 1 + a.NL;
    a.hp = *hp;
    a.y = lcw->y; a.yflag = lcw->yflag; a.aflag = lcw->aflag; a.acoef = lcw->acoef;
    a.part_w2 = lcw->part_w2; a.part_w1 = lcw->part_w1; a.part_wh = lcw->part_wh; a.part_hb = lcw->part_hb; a.part_loss = lcw->part_loss;
    a.grad = grad; a.loss = loss; a.gnorm = gnorm;
    a.pt = pt;
    a.m = adam != nullptr ? adam->m : nullptr; a.v = adam != nullptr ? adam->v : nullptr; a.step = adam != nullptr ? adam->step : nullptr;
    a.do_adam = do_adam;
    a.off_w1 = flat_offsets[0]; a.off_b1 = flat_offsets[1]; a.off_bn1w = flat_offsets[2]; a.off_bn1b = flat_offsets[3];
    a.off_w2 = flat_offsets[4]; a.off_b2 = flat_offsets[5]; a.off_bn2w = flat_offsets[6]; a.off_bn2b = flat_offsets[7];
    a.off_wmu = flat_offsets[8]; a.off_bmu = flat_offsets[9]; a.off_wv = flat_offsets[10]; a.off_bv = flat_offsets[11];
    a.off_wl = flat_offsets[12]; a.off_bl = flat_offsets[13]; a.n_params = flat_offsets[14];
    a.dbg = lcw->dbg;
    a.prof = lcw->prof;
    if (xchg != nullptr) {
        RLOA_REQUIRE(do_adam != 0 && xchg_connected_world(xchg) >= 1 && xchg_connected_world(xchg) <= kXchgInKernelWorld,
                     "learn cluster: the in-kernel exchange needs a connected handle of at most 8 ranks and the optimiser");
        a.use_xchg = 1;
        xchg_peers(xchg, &a.xp, &a.xstatus);
    }
    if (replay != nullptr) {
        a.use_replay = 1;
        a.rb = *replay->rb;
        a.rs_seed = replay->seed; a.rs_draw = replay->draw;
        a.rs_draw_offset = reinterpret_cast<const unsigned long long*>(replay->draw_offset);
        if (replay->pd_n > 0) {
            RLOA_REQUIRE(replay->pd_n <= kLearnClusterMaxPending && replay->pd_states && replay->pd_actions && replay->pd_rewards &&
                             replay->pd_next_states && (reinterpret_cast<uintptr_t>(replay->pd_valid) & 15u) == 0,
                         "learn cluster: bad pending-row arguments");
            a.pd_states = replay->pd_states; a.pd_actions = replay->pd_actions; a.pd_rewards = replay->pd_rewards;
            a.pd_next_states = replay->pd_next_states; a.pd_dones = replay->pd_dones; a.pd_valid = replay->pd_valid;
            a.pd_n = replay->pd_n;
        }
    }
    naf_learn_cluster_kernel<<<2 * lc::CL, lc::THREADS, lc::SMEM_BYTES, st>>>(a);
    RLOA_LAUNCHED();
    return RLOA_OK;
}

}  // namespace rloa
