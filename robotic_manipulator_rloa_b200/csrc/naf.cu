// librloa_b200: NAF network kernels and C ABI (include/rloa_b200.h).
//
// Replaces the torch ops under NAF.forward (reference naf_components/naf_neural_network.py:56-123),
// NAFAgent.act / learn / soft_update (naf_components/naf_algorithm.py:158-226).
// Reference quirks reproduced on purpose: P = L o L^T is elementwise, so only
// the diagonal of L matters and P_kk = exp(2 tanh z_kk); replayed actions are truncated like .long();
// the TD target has no (1 - done) factor; both nets run train-mode BatchNorm inside learn().
//
// fp32 CUDA-core path: tiled SGEMMs with fused BN+ReLU prologues, column-owned BatchNorm statistics
// (two-pass, no atomics), one fused head kernel (mu, V, diag L, advantage, TD error, head gradients),
// deterministic split-K weight gradients, fused clip-norm + Adam + soft target update.
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <new>

#include "common.cuh"
#include "bn_fuse.cuh"
#include "naf_trunk_tc.cuh"
#include "naf_learn_cluster.cuh"
#include "grad_exchange.cuh"
#include "optim.cuh"
#include "philox.cuh"

namespace rloa {


// ------------------------------------------------------------------------------------------------
// generic tiled SGEMM  C[i][j] = sum_k Aop(i,k) Bop(k,j)
//   NT: Aop = pro(A[i][k]),  Bop = W[j][k]      (forward linear, + bias[j])
//   NN: Aop = A[i][k],       Bop = W[k][j]      (input gradient)
//   TN: Aop = A[k][i],       Bop = pro(B[k][j]) (weight gradient, split over k = batch)
// pro(x, f) = relu(x * scale[f] + shift[f]) re-creates the BN+ReLU activation on the fly.
// ------------------------------------------------------------------------------------------------
enum { kNT = 0, kNN = 1, kTN = 2 };
constexpr int BM = 64, BN = 64, BK = 16, kGemmThreads = 256, kPad = 4;

struct GemmArgs {
    const float* A; int lda;
    const float* B; int ldb;
    float* C; int ldc;
    const float* bias;
    const float* pro_scale; const float* pro_shift;
    int M, N, K;
    int ksplit;                 // TN: k-range length per blockIdx.z (0 = no split)
};
struct GemmBatch {
    GemmArgs a[2];
    BnFuse bn[2];               // NT layout only: BatchNorm statistics of the output fused into the epilogue
};

// Epilogue of an NT tile whose output feeds a train-mode BatchNorm (bn_fuse.cuh): v[i][j] = the 4 x 4 outputs of this
// thread (rows i0 + 4 ty + i, columns j0 + 4 tx + j).  red: [16][BM + kPad] shared scratch, cm: [BN] shared scratch.
__device__ __forceinline__ void gemm_bn_epilogue(const BnFuse& f, float (*red)[BM + kPad], float* cm, const float (&v)[4][4],
                                                 int i0, int j0, int M, int H, int tx, int ty, int tid) {
    __shared__ unsigned s_last;
    const int rows = min(BM, M - i0);
    // pass 1: tile mean per column
#pragma unroll
    for (int j = 0; j < 4; j++) {
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 4; i++) s += (ty * 4 + i < rows) ? v[i][j] : 0.f;
        red[ty][tx * 4 + j] = s;
    }
    __syncthreads();
    if (tid < BN) {
        float s = 0.f;
#pragma unroll
        for (int r = 0; r < 16; r++) s += red[r][tid];
        cm[tid] = s / (float)rows;
    }
    __syncthreads();
    // pass 2: centred second moment
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const float m = cm[tx * 4 + j];
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const float d = (ty * 4 + i < rows) ? v[i][j] - m : 0.f;
            s = fmaf(d, d, s);
        }
        red[ty][tx * 4 + j] = s;
    }
    __syncthreads();
    const int chunks = gridDim.y;
    if (tid < BN && j0 + tid < H) {
        float s = 0.f;
#pragma unroll
        for (int r = 0; r < 16; r++) s += red[r][tid];
        float* P = f.part + (size_t)blockIdx.y * 2 * H;
        P[j0 + tid] = cm[tid];
        P[H + j0 + tid] = s;
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        const unsigned t = atomicAdd(&f.ticket[blockIdx.x], 1u);
        s_last = (t == (unsigned)chunks - 1u) ? 1u : 0u;
        if (s_last) f.ticket[blockIdx.x] = 0u;
    }
    __syncthreads();
    if (s_last && tid < BN && j0 + tid < H) {
        __threadfence();
        bn_finalize_column(f.bn, f.part, chunks, BM, M, H, j0 + tid, blockIdx.x == 0 && tid == 0);
    }
}

template <int LAYOUT, bool PRO>
__global__ void __launch_bounds__(kGemmThreads) gemm_kernel(GemmBatch batch) {
    const GemmArgs& g = (LAYOUT == kTN) ? batch.a[0] : batch.a[blockIdx.z];
    __shared__ float As[BK][BM + kPad];
    __shared__ float Bs[BK][BN + kPad];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int i0 = blockIdx.y * BM, j0 = blockIdx.x * BN;
    int k_begin = 0, k_end = g.K;
    float* C = g.C;
    if (LAYOUT == kTN && g.ksplit > 0) {
        k_begin = blockIdx.z * g.ksplit;
        k_end = min(g.K, k_begin + g.ksplit);
        C += (size_t)blockIdx.z * g.M * g.ldc;
    }
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = 0.f;
    for (int k0 = k_begin; k0 < k_end; k0 += BK) {
#pragma unroll
        for (int r = 0; r < 4; r++) {
            const int e = tid + r * kGemmThreads;
            {   // A tile
                int i, k;
                if (LAYOUT == kTN) { i = e & (BM - 1); k = e / BM; }
                else { k = e & (BK - 1); i = e / BK; }
                const int gi = i0 + i, gk = k0 + k;
                float v = 0.f;
                if (gi < g.M && gk < k_end) {
                    if (LAYOUT == kTN) v = g.A[(size_t)gk * g.lda + gi];
                    else {
                        v = g.A[(size_t)gi * g.lda + gk];
                        if (PRO && LAYOUT == kNT) v = fmaxf(fmaf(v, g.pro_scale[gk], g.pro_shift[gk]), 0.f);
                    }
                }
                As[k][i] = v;
            }
            {   // B tile
                int j, k;
                if (LAYOUT == kNT) { k = e & (BK - 1); j = e / BK; }
                else { j = e & (BN - 1); k = e / BN; }
                const int gj = j0 + j, gk = k0 + k;
                float v = 0.f;
                if (gj < g.N && gk < k_end) {
                    if (LAYOUT == kNT) v = g.B[(size_t)gj * g.ldb + gk];
                    else {
                        v = g.B[(size_t)gk * g.ldb + gj];
                        if (PRO && LAYOUT == kTN) v = fmaxf(fmaf(v, g.pro_scale[gj], g.pro_shift[gj]), 0.f);
                    }
                }
                Bs[k][j] = v;
            }
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < BK; k++) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; i++) a[i] = As[k][ty * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; j++) b[j] = Bs[k][tx * 4 + j];
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int j = 0; j < 4; j++) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int gi = i0 + ty * 4 + i;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int gj = j0 + tx * 4 + j;
            if (LAYOUT == kNT && g.bias != nullptr && gj < g.N) acc[i][j] += g.bias[gj];
            if (gi < g.M && gj < g.N) C[(size_t)gi * g.ldc + gj] = acc[i][j];
        }
    }
    if (LAYOUT == kNT && batch.bn[blockIdx.z].enabled)
        gemm_bn_epilogue(batch.bn[blockIdx.z], As, &Bs[0][0], acc, i0, j0, g.M, g.N, tx, ty, tid);
}

// Vectorised, double-buffered version of the same tile (used whenever lda, ldb, K-range and N are multiples of
// 4 and the operands are 16-byte aligned — every contraction of the network except the two that touch the
// S-wide input).  One float4 global load per operand per thread per k-step, issued a full tile ahead of the FMAs
// that consume it (registers -> the other shared-memory buffer), float4 shared-memory reads in the inner loop.
__device__ __forceinline__ float4 ldg4(const float* p) { return *reinterpret_cast<const float4*>(p); }

template <int LAYOUT, bool PRO>
__global__ void __launch_bounds__(kGemmThreads) gemm_vec_kernel(GemmBatch batch) {
    const GemmArgs& g = (LAYOUT == kTN) ? batch.a[0] : batch.a[blockIdx.z];
    __shared__ __align__(16) float As[2][BK][BM + kPad];
    __shared__ __align__(16) float Bs[2][BK][BN + kPad];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int i0 = blockIdx.y * BM, j0 = blockIdx.x * BN;
    int k_begin = 0, k_end = g.K;
    float* C = g.C;
    if (LAYOUT == kTN && g.ksplit > 0) {
        k_begin = blockIdx.z * g.ksplit;
        k_end = min(g.K, k_begin + g.ksplit);
        C += (size_t)blockIdx.z * g.M * g.ldc;
    }
    // operand fetch coordinates: "row-of-4-k" operands (A of NT/NN, B of NT) use (r = tid / 4, kq = tid % 4);
    // "row-of-4-mn" operands (A of TN, B of NN/TN) use (k = tid / 16, q = tid % 16)
    const int r4 = tid >> 2, kq = tid & 3, k16 = tid >> 4, q16 = tid & 15;
    float4 ra, rb;
    auto fetch = [&](int k0) {
        ra = make_float4(0.f, 0.f, 0.f, 0.f);
        rb = ra;
        if (LAYOUT == kTN) {
            const int gk = k0 + k16;
            if (gk < k_end) {
                if (i0 + 4 * q16 < g.M) ra = ldg4(g.A + (size_t)gk * g.lda + i0 + 4 * q16);
                if (j0 + 4 * q16 < g.N) {
                    rb = ldg4(g.B + (size_t)gk * g.ldb + j0 + 4 * q16);
                    if (PRO) {
                        const float4 sc = ldg4(g.pro_scale + j0 + 4 * q16), sh = ldg4(g.pro_shift + j0 + 4 * q16);
                        rb.x = fmaxf(fmaf(rb.x, sc.x, sh.x), 0.f); rb.y = fmaxf(fmaf(rb.y, sc.y, sh.y), 0.f);
                        rb.z = fmaxf(fmaf(rb.z, sc.z, sh.z), 0.f); rb.w = fmaxf(fmaf(rb.w, sc.w, sh.w), 0.f);
                    }
                }
            }
        } else {
            const int gk = k0 + 4 * kq;
            if (i0 + r4 < g.M && gk < k_end) {
                ra = ldg4(g.A + (size_t)(i0 + r4) * g.lda + gk);
                if (PRO) {
                    const float4 sc = ldg4(g.pro_scale + gk), sh = ldg4(g.pro_shift + gk);
                    ra.x = fmaxf(fmaf(ra.x, sc.x, sh.x), 0.f); ra.y = fmaxf(fmaf(ra.y, sc.y, sh.y), 0.f);
                    ra.z = fmaxf(fmaf(ra.z, sc.z, sh.z), 0.f); ra.w = fmaxf(fmaf(ra.w, sc.w, sh.w), 0.f);
                }
            }
            if (LAYOUT == kNT) {
                if (j0 + r4 < g.N && gk < k_end) rb = ldg4(g.B + (size_t)(j0 + r4) * g.ldb + gk);
            } else {
                const int gk2 = k0 + k16;
                if (gk2 < k_end && j0 + 4 * q16 < g.N) rb = ldg4(g.B + (size_t)gk2 * g.ldb + j0 + 4 * q16);
            }
        }
    };
    auto stash = [&](int buf) {
        if (LAYOUT == kTN) {
            *reinterpret_cast<float4*>(&As[buf][k16][4 * q16]) = ra;
            *reinterpret_cast<float4*>(&Bs[buf][k16][4 * q16]) = rb;
        } else {
            As[buf][4 * kq][r4] = ra.x; As[buf][4 * kq + 1][r4] = ra.y; As[buf][4 * kq + 2][r4] = ra.z; As[buf][4 * kq + 3][r4] = ra.w;
            if (LAYOUT == kNT) {
                Bs[buf][4 * kq][r4] = rb.x; Bs[buf][4 * kq + 1][r4] = rb.y; Bs[buf][4 * kq + 2][r4] = rb.z; Bs[buf][4 * kq + 3][r4] = rb.w;
            } else {
                *reinterpret_cast<float4*>(&Bs[buf][k16][4 * q16]) = rb;
            }
        }
    };
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = 0.f;
    fetch(k_begin);
    stash(0);
    __syncthreads();
    int buf = 0;
    for (int k0 = k_begin; k0 < k_end; k0 += BK) {
        const bool more = k0 + BK < k_end;
        if (more) fetch(k0 + BK);
#pragma unroll
        for (int k = 0; k < BK; k++) {
            const float4 a = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
            const float4 b = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
            const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int j = 0; j < 4; j++) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        if (more) stash(buf ^ 1);
        __syncthreads();
        buf ^= 1;
    }
    const int gj = j0 + tx * 4;
    if (LAYOUT == kNT && g.bias != nullptr && gj < g.N) {
        const float4 bb = ldg4(g.bias + gj);
#pragma unroll
        for (int i = 0; i < 4; i++) { acc[i][0] += bb.x; acc[i][1] += bb.y; acc[i][2] += bb.z; acc[i][3] += bb.w; }
    }
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int gi = i0 + ty * 4 + i;
        if (gi >= g.M || gj >= g.N) continue;
        *reinterpret_cast<float4*>(C + (size_t)gi * g.ldc + gj) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
    }
    if (LAYOUT == kNT && batch.bn[blockIdx.z].enabled)
        gemm_bn_epilogue(batch.bn[blockIdx.z], As[0], &Bs[0][0][0], acc, i0, j0, g.M, g.N, tx, ty, tid);
}

static bool gemm_vec_ok(const GemmArgs& g, int layout) {
    auto al = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; };
    bool ok = g.lda % 4 == 0 && g.ldb % 4 == 0 && g.ldc % 4 == 0 && g.N % 4 == 0 && al(g.A) && al(g.B) && al(g.C) &&
              (g.bias == nullptr || al(g.bias)) && (g.pro_scale == nullptr || (al(g.pro_scale) && al(g.pro_shift)));
    if (layout == kTN) ok = ok && g.M % 4 == 0;
    else ok = ok && g.K % 4 == 0;
    return ok;
}

// launches the vectorised kernel when every operand of the batch qualifies, the scalar one otherwise
template <int LAYOUT, bool PRO>
static void launch_gemm(const GemmBatch& gb, int nbatch, dim3 grid, cudaStream_t st) {
    bool vec = true;
    for (int n = 0; n < nbatch; n++) vec = vec && gemm_vec_ok(gb.a[n], LAYOUT);
    if (vec) gemm_vec_kernel<LAYOUT, PRO><<<grid, kGemmThreads, 0, st>>>(gb);
    else gemm_kernel<LAYOUT, PRO><<<grid, kGemmThreads, 0, st>>>(gb);
}

// ------------------------------------------------------------------------------------------------
// BatchNorm1d statistics, one block per 32 features, all rows (two-pass: mean, then centred variance).
// train: scale = w / sqrt(var_b + eps), shift = b - mean_b scale; running stats updated with momentum 0.1
// and the unbiased variance; num_batches_tracked += 1.   eval: the same from the running statistics.
// ------------------------------------------------------------------------------------------------
struct BnBatch {
    BnArgs a[2];
};

__device__ __forceinline__ float block_colsum_32x8(float v, float (*red)[33]) {
    // 256 threads = 8 row lanes x 32 columns; returns the column total to every thread of the column
    const int c = threadIdx.x & 31, r = threadIdx.x >> 5;
    red[r][c] = v;
    __syncthreads();
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 8; k++) s += red[k][c];
    __syncthreads();
    return s;
}

__global__ void __launch_bounds__(256) bn_stats_kernel(BnBatch batch, int B, int H, int train) {
    const BnArgs& a = batch.a[blockIdx.y];
    __shared__ float red[8][33];
    const int c = threadIdx.x & 31, r = threadIdx.x >> 5;
    const int f = blockIdx.x * 32 + c;
    if (f >= H) return;         // H is a multiple of 32 (checked on the host): whole block stays converged
    float mean, var;
    if (train) {
        float s = 0.f;
        for (int row = r; row < B; row += 8) s += a.z[(size_t)row * H + f];
        mean = block_colsum_32x8(s, red) / (float)B;
        float m2 = 0.f;
        for (int row = r; row < B; row += 8) {
            const float d = a.z[(size_t)row * H + f] - mean;
            m2 = fmaf(d, d, m2);
        }
        m2 = block_colsum_32x8(m2, red);
        var = m2 / (float)B;
        if (r == 0) {
            const float unbiased = B > 1 ? m2 / (float)(B - 1) : var;
            a.run_mean[f] = fmaf(kBnMomentum, mean - a.run_mean[f], a.run_mean[f]);
            a.run_var[f] = fmaf(kBnMomentum, unbiased - a.run_var[f], a.run_var[f]);
            if (blockIdx.x == 0 && c == 0 && a.batches != nullptr) *a.batches += 1;
        }
    } else {
        mean = a.run_mean[f];
        var = a.run_var[f];
    }
    if (r == 0) {
        const float rstd = 1.f / sqrtf(var + kBnEps);
        const float sc = a.w[f] * rstd;
        a.scale[f] = sc;
        a.shift[f] = fmaf(-mean, sc, a.b[f]);
        a.mean[f] = mean;
        a.rstd[f] = rstd;
    }
}

// Train-mode statistics at full-chip parallelism: grid (H/32, R row chunks, nets).  Every block computes the
// (count, mean, M2) of its row chunk two-pass from registers; the last block to arrive for a column group
// merges the R partials in a fixed order with Chan's pairwise update (deterministic, as accurate as the
// two-pass form) and finalises exactly like bn_stats_kernel.  `ticket` counters return to 0 (graph replay).
constexpr int kBnRowsPerThread = 8;            // 8 row lanes x 8 rows = 64 rows per block
__global__ void __launch_bounds__(256)
bn_stats_split_kernel(BnBatch batch, int B, int H, float* __restrict__ part, unsigned* __restrict__ ticket) {
    const BnArgs& a = batch.a[blockIdx.z];
    __shared__ float red[8][33];
    __shared__ unsigned s_last;
    const int c = threadIdx.x & 31, r = threadIdx.x >> 5;
    const int f = blockIdx.x * 32 + c;
    const int R = gridDim.y;
    const int row0 = blockIdx.y * (8 * kBnRowsPerThread);
    const int nrows = min(8 * kBnRowsPerThread, B - row0);
    float x[kBnRowsPerThread];
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < kBnRowsPerThread; k++) {
        const int row = row0 + r + 8 * k;
        x[k] = row < B ? a.z[(size_t)row * H + f] : 0.f;
        s += x[k];
    }
    const float mean_c = block_colsum_32x8(s, red) / (float)nrows;
    float m2 = 0.f;
#pragma unroll
    for (int k = 0; k < kBnRowsPerThread; k++) {
        const int row = row0 + r + 8 * k;
        const float d = row < B ? x[k] - mean_c : 0.f;
        m2 = fmaf(d, d, m2);
    }
    m2 = block_colsum_32x8(m2, red);
    float* P = part + ((size_t)blockIdx.z * R + blockIdx.y) * 2 * H;     // [net][chunk][mean | M2][H]
    if (r == 0) {
        P[f] = mean_c;
        P[H + f] = m2;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned t = atomicAdd(&ticket[blockIdx.z * gridDim.x + blockIdx.x], 1u);
        s_last = (t == (unsigned)R - 1u) ? 1u : 0u;
        if (s_last) ticket[blockIdx.z * gridDim.x + blockIdx.x] = 0u;
    }
    __syncthreads();
    if (!s_last || r != 0) return;
    __threadfence();
    bn_finalize_column(a, part + (size_t)blockIdx.z * R * 2 * H, R, 8 * kBnRowsPerThread, B, H, f, blockIdx.x == 0 && c == 0);
}

// ------------------------------------------------------------------------------------------------
// BN + ReLU backward, column owned: g = da * [z scale + shift > 0]; dgamma = sum g xhat; dbeta = sum g;
// dz = w rstd (g - dbeta/B - xhat dgamma/B) written in place over da; dbias(linear) = sum dz.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
bn_relu_backward_kernel(float* __restrict__ da, const float* __restrict__ z, const float* __restrict__ w,
                        const float* __restrict__ scale, const float* __restrict__ shift, const float* __restrict__ mean,
                        const float* __restrict__ rstd, int B, int H, float* __restrict__ d_w, float* __restrict__ d_b,
                        float* __restrict__ d_lin_bias) {
    __shared__ float red[8][33];
    const int c = threadIdx.x & 31, r = threadIdx.x >> 5;
    const int f = blockIdx.x * 32 + c;
    if (f >= H) return;
    const float sc = scale[f], sh = shift[f], mu = mean[f], rs = rstd[f];
    float sg = 0.f, sgx = 0.f;
    for (int row = r; row < B; row += 8) {
        const size_t idx = (size_t)row * H + f;
        const float zz = z[idx];
        const float g = fmaf(zz, sc, sh) > 0.f ? da[idx] : 0.f;
        sg += g;
        sgx = fmaf(g, (zz - mu) * rs, sgx);
    }
    sg = block_colsum_32x8(sg, red);
    sgx = block_colsum_32x8(sgx, red);
    const float k1 = w[f] * rs, c1 = sg / (float)B, c2 = sgx / (float)B;
    float sdz = 0.f;
    for (int row = r; row < B; row += 8) {
        const size_t idx = (size_t)row * H + f;
        const float zz = z[idx];
        const float g = fmaf(zz, sc, sh) > 0.f ? da[idx] : 0.f;
        const float dz = k1 * (g - c1 - (zz - mu) * rs * c2);
        da[idx] = dz;
        sdz += dz;
    }
    sdz = block_colsum_32x8(sdz, red);
    if (r == 0) {
        d_w[f] = sgx;
        d_b[f] = sg;
        d_lin_bias[f] = sdz;
    }
}

// The same backward at full-chip parallelism.  Pass 1 (grid (H/32, R)): per-chunk column sums of g, g xhat and
// (z - mean); the last block of a column group adds the R partials in a fixed order and leaves the per-column
// coefficients of dz = k1 g - cB - cC (z - mean):  k1 = w rstd, cB = k1 sum(g)/B, cC = k1 rstd sum(g xhat)/B.
// Pass 2 (elementwise, float4): da <- dz in place.
struct BnBwdCoef {
    float *k1, *cB, *cC;        // [H] each
};
__global__ void __launch_bounds__(256)
bn_bwd_reduce_kernel(const float* __restrict__ da, const float* __restrict__ z, const float* __restrict__ w,
                     const float* __restrict__ scale, const float* __restrict__ shift, const float* __restrict__ mean,
                     const float* __restrict__ rstd, int B, int H, float* __restrict__ part, unsigned* __restrict__ ticket,
                     BnBwdCoef coef, float* __restrict__ d_w, float* __restrict__ d_b, float* __restrict__ d_lin_bias) {
    __shared__ float red[8][33];
    __shared__ unsigned s_last;
    const int c = threadIdx.x & 31, r = threadIdx.x >> 5;
    const int f = blockIdx.x * 32 + c;
    const int R = gridDim.y;
    const int row0 = blockIdx.y * (8 * kBnRowsPerThread);
    const float sc = scale[f], sh = shift[f], mu = mean[f], rs = rstd[f];
    // every operand is requested before the first use and the ReLU mask is a select, not a guarded load: one memory
    // round trip for the block instead of two dependent ones per row
    float zv[kBnRowsPerThread], dv_[kBnRowsPerThread];
#pragma unroll
    for (int k = 0; k < kBnRowsPerThread; k++) {
        const int row = row0 + r + 8 * k;
        const size_t idx = (size_t)(row < B ? row : 0) * H + f;
        zv[k] = z[idx];
        dv_[k] = da[idx];
    }
    float sg = 0.f, sgx = 0.f, sz = 0.f;
#pragma unroll
    for (int k = 0; k < kBnRowsPerThread; k++) {
        if (row0 + r + 8 * k < B) {
            const float zz = zv[k];
            const float g = fmaf(zz, sc, sh) > 0.f ? dv_[k] : 0.f;
            sg += g;
            sgx = fmaf(g, (zz - mu) * rs, sgx);
            sz += zz - mu;
        }
    }
    sg = block_colsum_32x8(sg, red);
    sgx = block_colsum_32x8(sgx, red);
    sz = block_colsum_32x8(sz, red);
    float* P = part + (size_t)blockIdx.y * 3 * H;
    if (r == 0) {
        P[f] = sg;
        P[H + f] = sgx;
        P[2 * H + f] = sz;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned t = atomicAdd(&ticket[blockIdx.x], 1u);
        s_last = (t == (unsigned)R - 1u) ? 1u : 0u;
        if (s_last) ticket[blockIdx.x] = 0u;
    }
    __syncthreads();
    if (!s_last || r != 0) return;
    __threadfence();
    // partials requested in batches of 16 chunks before they are added (in chunk order): the loads overlap
    float tg = 0.f, tgx = 0.f, tz = 0.f;
    for (int k0 = 0; k0 < R; k0 += 16) {
        float a0[16], a1[16], a2[16];
#pragma unroll
        for (int k = 0; k < 16; k++) {
            const bool ok = k0 + k < R;
            a0[k] = ok ? __ldcg(part + (size_t)(k0 + k) * 3 * H + f) : 0.f;
            a1[k] = ok ? __ldcg(part + (size_t)(k0 + k) * 3 * H + H + f) : 0.f;
            a2[k] = ok ? __ldcg(part + (size_t)(k0 + k) * 3 * H + 2 * H + f) : 0.f;
        }
#pragma unroll
        for (int k = 0; k < 16; k++) { tg += a0[k]; tgx += a1[k]; tz += a2[k]; }
    }
    const float k1 = w[f] * rs, c1 = tg / (float)B, c2 = tgx / (float)B;
    coef.k1[f] = k1;
    coef.cB[f] = k1 * c1;
    coef.cC[f] = k1 * c2 * rs;
    d_w[f] = tgx;
    d_b[f] = tg;
    // sum over rows of dz: analytically 0, numerically the rounding residue (as in the reference's autograd)
    d_lin_bias[f] = k1 * (tg - (float)B * c1) - k1 * c2 * rs * tz;
}

__global__ void __launch_bounds__(256)
bn_bwd_apply_kernel(float* __restrict__ da, const float* __restrict__ z, const float* __restrict__ scale,
                    const float* __restrict__ shift, const float* __restrict__ mean, BnBwdCoef coef, int B, int H) {
    const size_t n4 = (size_t)B * H / 4;
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n4; i += (size_t)gridDim.x * 256) {
        const int f = (int)((i * 4) % H);
        const float4 zz = reinterpret_cast<const float4*>(z)[i];
        float4 g = reinterpret_cast<float4*>(da)[i];
        const float4 sc = *reinterpret_cast<const float4*>(scale + f), sh = *reinterpret_cast<const float4*>(shift + f);
        const float4 mu = *reinterpret_cast<const float4*>(mean + f), k1 = *reinterpret_cast<const float4*>(coef.k1 + f);
        const float4 cB = *reinterpret_cast<const float4*>(coef.cB + f), cC = *reinterpret_cast<const float4*>(coef.cC + f);
        g.x = fmaf(zz.x, sc.x, sh.x) > 0.f ? g.x : 0.f;
        g.y = fmaf(zz.y, sc.y, sh.y) > 0.f ? g.y : 0.f;
        g.z = fmaf(zz.z, sc.z, sh.z) > 0.f ? g.z : 0.f;
        g.w = fmaf(zz.w, sc.w, sh.w) > 0.f ? g.w : 0.f;
        float4 o;
        o.x = fmaf(k1.x, g.x, -cB.x) - cC.x * (zz.x - mu.x);
        o.y = fmaf(k1.y, g.y, -cB.y) - cC.y * (zz.y - mu.y);
        o.z = fmaf(k1.z, g.z, -cB.z) - cC.z * (zz.z - mu.z);
        o.w = fmaf(k1.w, g.w, -cB.w) - cC.w * (zz.w - mu.w);
        reinterpret_cast<float4*>(da)[i] = o;
    }
}

// Layer-1 BatchNorm backward fused with the input-layer weight gradient: dz1 = k1 g - cB - cC (z1 - mean) is only
// ever consumed by dW1 = dz1^T x (and by the bias gradient, which bn_bwd_reduce_kernel already produced), so it is
// never written: block = 32 batch rows, thread = hidden column h, accumulates dW1[h][0..S) over the block's rows with
// the x tile broadcast from shared memory, and leaves one split-K partial per block (summed in chunk order by the tail).
constexpr int kDw1Rows = 64, kDw1MaxS = 24, kDw1Cols = 128;
__global__ void __launch_bounds__(256)
bn_bwd_apply_dw1_kernel(const float* __restrict__ da, const float* __restrict__ z, const float* __restrict__ scale,
                        const float* __restrict__ shift, const float* __restrict__ mean, BnBwdCoef coef,
                        const float* __restrict__ x, int B, int H, int S, float* __restrict__ part) {
    // block = (64-row chunk, 128 columns); thread = (column h, half of the chunk's rows): 32 rows each, all 64 operand
    // loads in flight at once, then the two halves are added through shared memory
    __shared__ __align__(16) float xs[kDw1Rows][kDw1MaxS];
    __shared__ float half_acc[kDw1Cols][kDw1MaxS + 1];
    const int row0 = blockIdx.x * kDw1Rows, rows = min(kDw1Rows, B - row0);
    const int hl = threadIdx.x & (kDw1Cols - 1), rh = threadIdx.x >> 7;
    const int h = blockIdx.y * kDw1Cols + hl;
    constexpr int kHalf = kDw1Rows / 2;
    float zz[kHalf], gg[kHalf];
#pragma unroll
    for (int u = 0; u < kHalf; u++) {
        const int r = rh * kHalf + u;
        const size_t idx = (size_t)(row0 + (r < rows ? r : 0)) * H + (h < H ? h : 0);
        zz[u] = z[idx];
        gg[u] = da[idx];
    }
    for (int e = threadIdx.x; e < kDw1Rows * kDw1MaxS; e += 256) {
        const int r = e / kDw1MaxS, k = e - r * kDw1MaxS;
        xs[r][k] = (r < rows && k < S) ? x[(size_t)(row0 + r) * S + k] : 0.f;
    }
    __syncthreads();
    float acc[kDw1MaxS];
#pragma unroll
    for (int k = 0; k < kDw1MaxS; k++) acc[k] = 0.f;
    if (h < H) {
        const float sc = scale[h], sh = shift[h], mu = mean[h], k1 = coef.k1[h], cB = coef.cB[h], cC = coef.cC[h];
#pragma unroll
        for (int u = 0; u < kHalf; u++) {
            const int r = rh * kHalf + u;
            const float g = fmaf(zz[u], sc, sh) > 0.f ? gg[u] : 0.f;
            const float dz = r < rows ? fmaf(k1, g, -cB) - cC * (zz[u] - mu) : 0.f;
#pragma unroll
            for (int k4 = 0; k4 < kDw1MaxS / 4; k4++) {
                const float4 xv = *reinterpret_cast<const float4*>(&xs[r][4 * k4]);
                acc[4 * k4] = fmaf(dz, xv.x, acc[4 * k4]);
                acc[4 * k4 + 1] = fmaf(dz, xv.y, acc[4 * k4 + 1]);
                acc[4 * k4 + 2] = fmaf(dz, xv.z, acc[4 * k4 + 2]);
                acc[4 * k4 + 3] = fmaf(dz, xv.w, acc[4 * k4 + 3]);
            }
        }
    }
    if (rh == 1) {
#pragma unroll
        for (int k = 0; k < kDw1MaxS; k++) half_acc[hl][k] = acc[k];
    }
    __syncthreads();
    if (rh == 0 && h < H) {
        float* out = part + (size_t)blockIdx.x * H * S + (size_t)h * S;
#pragma unroll
        for (int k = 0; k < kDw1MaxS; k++)
            if (k < S) out[k] = acc[k] + half_acc[hl][k];       // first half of the rows, then the second: fixed order
    }
}

// ------------------------------------------------------------------------------------------------
// Philox4x32-10 (philox.cuh) + Box-Muller for the exploration noise (replaces MultivariateNormal.sample())
// ------------------------------------------------------------------------------------------------
// ------------------------------------------------------------------------------------------------
// Fused head: a2 = relu(bn2(z2)); [mu | V | L-entries] = a2 Wh^T + bh; mu = tanh, l_kk = tanh;
// P_kk = exp(2 l_kk); Q = V - 1/2 sum_k P_kk (u_k - mu_k)^2; TD error; gradients of the head
// pre-activations; or (act mode) the clamped noisy action.   One warp per row, lane = head output.
// ------------------------------------------------------------------------------------------------
struct HeadArgs {
    const float* z2; const float *scale, *shift;    // [B][H], [H]
    const float *w_mu, *b_mu, *w_v, *b_v, *w_l, *b_l;
    int B, H, A, NL;
    const float* action;         // [B][A] or null
    int trunc_action;
    // outputs (any may be null)
    float *mu, *pdiag, *q, *v;
    // TD-target production (target net): y = reward + gamma * V * (use_done ? 1 - done : 1)
    const float *reward, *done; float gamma; int use_done; float* y_out;
    // training (main net): y_in -> loss partial + dZh
    const float* y_in; float* dzh; float* loss_part;
    // fused learn: TD target of the SAME row from the target net's trunk output (row-local dependency), and the
    // gradient with respect to a2, da2 = dZh Wh, written straight from the head weights held in shared memory
    const float* t_z2; const float *t_scale, *t_shift, *t_wv, *t_bv;
    float* da_out;
    // act mode
    float* act_out; unsigned long long seed, step; const unsigned long long* step_offset; float noise_scale;
};

constexpr int kHeadWarps = 8;

__global__ void __launch_bounds__(kHeadWarps * 32) naf_head_kernel(HeadArgs h) {
    extern __shared__ float hs[];
    const int H = h.H, A = h.A, NL = h.NL, NH = A + 1 + NL, HS = H + 1;
    float* Wh = hs;                               // [NH][H+1]
    float* bh = Wh + NH * HS;                     // [NH]
    float* arow = bh + ((NH + 31) & ~31);         // [warps][H]
    float* twv = arow + kHeadWarps * H;           // [H] target value-head weights (fused learn only)
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (h.t_z2 != nullptr)
        for (int k = tid; k < H; k += blockDim.x) twv[k] = h.t_wv[k];
    // head weights -> shared memory: the three tensors are copied as flat float4 streams, 8 loads in flight per thread
    // (a per-element loop with the tensor select inside cost one dependent L2 round trip per iteration)
    {
        const int H4 = H >> 2;                       // H is a multiple of 32
        auto stage = [&](const float* __restrict__ src, int rows, int first_row) {
            const int n4 = rows * H4;
            for (int i0 = tid; i0 < n4; i0 += 8 * blockDim.x) {
                float4 v[8];
#pragma unroll
                for (int u = 0; u < 8; u++) {
                    const int i = i0 + u * blockDim.x;
                    v[u] = i < n4 ? reinterpret_cast<const float4*>(src)[i] : make_float4(0.f, 0.f, 0.f, 0.f);
                }
#pragma unroll
                for (int u = 0; u < 8; u++) {
                    const int i = i0 + u * blockDim.x;
                    if (i < n4) {
                        const int o = i / H4, k = 4 * (i - o * H4);
                        float* d = Wh + (first_row + o) * HS + k;
                        d[0] = v[u].x; d[1] = v[u].y; d[2] = v[u].z; d[3] = v[u].w;
                    }
                }
            }
        };
        stage(h.w_mu, A, 0);
        stage(h.w_v, 1, A);
        stage(h.w_l, NL, A + 1);
    }
    for (int o = tid; o < NH; o += blockDim.x) bh[o] = o < A ? h.b_mu[o] : (o == A ? h.b_v[0] : h.b_l[o - A - 1]);
    __syncthreads();
    float* ar = arow + warp * H;
    float loss_acc = 0.f;
    const int row_stride = gridDim.x * kHeadWarps;
    for (int row = blockIdx.x * kHeadWarps + warp; row < h.B; row += row_stride) {
        float y_row = 0.f;
        if (h.t_z2 != nullptr) {                  // y = r + gamma V_target(s') for this row (naf_algorithm.py:196-199)
            float acc = 0.f;
            for (int k = lane; k < H; k += 32)
                acc = fmaf(fmaxf(fmaf(h.t_z2[(size_t)row * H + k], h.t_scale[k], h.t_shift[k]), 0.f), twv[k], acc);
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
            float vv = acc + h.t_bv[0];
            if (h.use_done && h.done) vv *= (1.f - h.done[row]);
            y_row = fmaf(h.gamma, vv, h.reward[row]);
        } else if (h.y_in != nullptr) {
            y_row = h.y_in[row];
        }
        for (int k = lane; k < H; k += 32) ar[k] = fmaxf(fmaf(h.z2[(size_t)row * H + k], h.scale[k], h.shift[k]), 0.f);
        __syncwarp();
        float zo[2] = {0.f, 0.f};                 // head pre-activations of outputs lane, lane + 32
#pragma unroll
        for (int s = 0; s < 2; s++) {
            const int o = lane + 32 * s;
            if (o < NH) {
                const float* wr = Wh + o * HS;
                float acc0 = 0.f, acc1 = 0.f;
                for (int k = 0; k < H; k += 2) {
                    acc0 = fmaf(ar[k], wr[k], acc0);
                    acc1 = fmaf(ar[k + 1], wr[k + 1], acc1);
                }
                zo[s] = acc0 + acc1 + bh[o];
            }
        }
        __syncwarp();
        // lane k < A gathers its mu pre-activation (own slot 0) and the diagonal entry k(k+3)/2 of L
        const int didx = A + 1 + (lane * (lane + 3)) / 2;
        const int dsrc = (lane < A) ? (didx & 31) : lane;
        const float d0 = __shfl_sync(0xffffffffu, zo[0], dsrc), d1 = __shfl_sync(0xffffffffu, zo[1], dsrc);
        const float vval = __shfl_sync(0xffffffffu, zo[0], A & 31);     // A < 32
        float mu = 0.f, t = 0.f, P = 0.f, diff = 0.f, adv = 0.f;
        if (lane < A) {
            mu = tanhf(zo[0]);
            t = tanhf(didx < 32 ? d0 : d1);
            P = expf(2.f * t);
            if (h.action != nullptr) {
                float u = h.action[(size_t)row * A + lane];
                if (h.trunc_action) u = truncf(u);
                diff = u - mu;
                adv = -0.5f * P * diff * diff;
            }
            if (h.mu) h.mu[(size_t)row * A + lane] = mu;
            if (h.pdiag) h.pdiag[(size_t)row * A + lane] = P;
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) adv += __shfl_xor_sync(0xffffffffu, adv, off);
        const float q = adv + vval;
        if (lane == 0) {
            if (h.v) h.v[row] = vval;
            if (h.q && h.action) h.q[row] = q;
            if (h.y_out) {
                float vv = vval;
                if (h.use_done && h.done) vv *= (1.f - h.done[row]);
                h.y_out[row] = fmaf(h.gamma, vv, h.reward[row]);
            }
        }
        if (h.dzh != nullptr) {
            // MSE over the batch: dLoss/dQ = 2 (Q - y) / B
            const float err = q - y_row;
            if (lane == 0) loss_acc = fmaf(err, err, loss_acc);
            const float dq = 2.f * err / (float)h.B;
            const float g_mu = dq * P * diff * (1.f - mu * mu);         // lanes < A
            const float g_l = dq * (-P * diff * diff) * (1.f - t * t);  // lanes < A, belongs to output didx
            float gs[2] = {0.f, 0.f};
            // route g_l from lane k to the lane/slot that owns output didx(k)
#pragma unroll
            for (int s = 0; s < 2; s++) {
                const int o = lane + 32 * s;
                float g = 0.f;
                // which k (if any) has didx(k) == o ?  k(k+3)/2 = o - A - 1
                int ksrc = -1;
                const int e = o - A - 1;
                if (e >= 0) {
                    for (int k = 0; k < A; k++) if ((k * (k + 3)) / 2 == e) ksrc = k;
                }
                const float gl = __shfl_sync(0xffffffffu, g_l, ksrc < 0 ? 0 : ksrc);
                if (o < A) g = g_mu;
                else if (o == A) g = dq;
                else if (ksrc >= 0) g = gl;
                if (o < NH) h.dzh[(size_t)row * NH + o] = g;
                gs[s] = o < NH ? g : 0.f;
            }
            if (h.da_out != nullptr) {            // da2[row][f] = sum_o dZh[row][o] Wh[o][f], f = lane + 32 j (H <= 256)
                float acc[8];
#pragma unroll
                for (int j = 0; j < 8; j++) acc[j] = 0.f;
                for (int o = 0; o < NH; o++) {
                    const float go = __shfl_sync(0xffffffffu, o < 32 ? gs[0] : gs[1], o & 31);
                    const float* wr = Wh + o * HS + lane;
#pragma unroll
                    for (int j = 0; j < 8; j++)
                        if (32 * j < H) acc[j] = fmaf(go, wr[32 * j], acc[j]);
                }
#pragma unroll
                for (int j = 0; j < 8; j++)
                    if (32 * j < H) h.da_out[(size_t)row * H + lane + 32 * j] = acc[j];
            }
        }
        if (h.act_out != nullptr && lane < A) {
            // action ~ N(mu, P^-1) with diagonal P: std_k = exp(-l_kk); clamp to [-1, 1]
            const unsigned long long stp = h.step + (h.step_offset != nullptr ? *h.step_offset : 0ull);
            uint32_t c[4] = {(uint32_t)row, (uint32_t)lane, (uint32_t)stp, (uint32_t)(stp >> 32)};
            philox4x32_10(c, (uint32_t)h.seed, (uint32_t)(h.seed >> 32));
            const float r = sqrtf(-2.f * logf(u01(c[0])));
            const float eps = r * cospif(2.f * u01(c[1]));
            const float a = fmaf(h.noise_scale * expf(-t), eps, mu);
            h.act_out[(size_t)row * A + lane] = fminf(fmaxf(a, -1.f), 1.f);
        }
        __syncwarp();
    }
    if (h.loss_part != nullptr) {
        __shared__ float lp[kHeadWarps];
        if (lane == 0) lp[warp] = loss_acc;
        __syncthreads();
        if (tid == 0) {
            float s = 0.f;
            for (int w = 0; w < kHeadWarps; w++) s += lp[w];
            h.loss_part[blockIdx.x] = s;
        }
    }
}

// column sums of dZh [B][NH] -> head bias gradients, and the loss scalar from the per-block partials
__global__ void __launch_bounds__(256)
head_bias_grad_kernel(const float* __restrict__ dzh, int B, int NH, int A, float* __restrict__ g_bmu,
                      float* __restrict__ g_bv, float* __restrict__ g_bl, const float* __restrict__ loss_part, int nparts,
                      float* __restrict__ loss) {
    __shared__ float red[256];
    const int o = blockIdx.x;
    float s = 0.f;
    if (o < NH) {
        for (int row = threadIdx.x; row < B; row += 256) s += dzh[(size_t)row * NH + o];
    } else {
        for (int i = threadIdx.x; i < nparts; i += 256) s += loss_part[i];
    }
    red[threadIdx.x] = s;
    __syncthreads();
    for (int off = 128; off > 0; off >>= 1) {
        if (threadIdx.x < off) red[threadIdx.x] += red[threadIdx.x + off];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const float t = red[0];
        if (o < A) g_bmu[o] = t;
        else if (o == A) g_bv[0] = t;
        else if (o < NH) g_bl[o - A - 1] = t;
        else if (loss != nullptr) *loss = t / (float)B;
    }
}

__global__ void __launch_bounds__(256) splitk_reduce_kernel(ReduceArgs r) {
    const ReduceSeg& s = r.s[blockIdx.y];
    for (int i = blockIdx.x * 256 + threadIdx.x; i < s.n; i += gridDim.x * 256) {
        float t = 0.f;
        for (int k = 0; k < s.nsplit; k++) t += s.part[k * s.pstride + i];
        s.out[i] = t;
    }
}

// ------------------------------------------------------------------------------------------------
// optimiser: global grad norm, then clip + Adam + soft target update in one vectorised pass
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) grad_sqsum_kernel(const float* __restrict__ g, int n, float scale,
                                                         float* __restrict__ partial, int64_t* __restrict__ step_ptr) {
    __shared__ float red[256];
    if (blockIdx.x == 0 && threadIdx.x == 0 && step_ptr != nullptr) *step_ptr += 1;   // optimizer.step() counter
    // same element-to-thread mapping and accumulation order as xchg_reduce_kernel (grad_exchange.cu), so the
    // single-rank and the exchanged update produce the same bits
    float s = 0.f;
    const int n4 = ((reinterpret_cast<uintptr_t>(g) & 15u) == 0) ? (n >> 2) : 0;
    for (int i = blockIdx.x * 256 + threadIdx.x; i < n4; i += gridDim.x * 256) {
        const float4 v = reinterpret_cast<const float4*>(g)[i];
        const float x0 = v.x * scale, x1 = v.y * scale, x2 = v.z * scale, x3 = v.w * scale;
        s = fmaf(x0, x0, s); s = fmaf(x1, x1, s); s = fmaf(x2, x2, s); s = fmaf(x3, x3, s);
    }
    for (int i = 4 * n4 + blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256) {
        const float x = g[i] * scale;
        s = fmaf(x, x, s);
    }
    red[threadIdx.x] = s;
    __syncthreads();
    for (int off = 128; off > 0; off >>= 1) {
        if (threadIdx.x < off) red[threadIdx.x] += red[threadIdx.x + off];
        __syncthreads();
    }
    if (threadIdx.x == 0) partial[blockIdx.x] = red[0];
}

__global__ void __launch_bounds__(256)
adam_soft_update_kernel(ParamTable pt, const float* __restrict__ grad, const float* __restrict__ sq_partial,
                        float* __restrict__ m, float* __restrict__ v, int64_t* __restrict__ step_ptr, rloa_naf_hyper hp,
                        float* __restrict__ grad_norm_out) {
    __shared__ AdamCoef s_c;
    if (threadIdx.x == 0) {
        s_c = adam_coefficients(sq_partial, *step_ptr, hp);      // step already incremented by grad_sqsum_kernel
        if (blockIdx.x == 0 && grad_norm_out != nullptr) *grad_norm_out = s_c.norm;
    }
    __syncthreads();
    const AdamCoef c = s_c;
    const int n = pt.offset[14];
    for (int i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256)
        adam_soft_update_element(pt, i, grad[i], c, m, v, hp);
}

// The single-GPU tail of NAFAgent.learn in ONE launch (kNormBlocks co-resident blocks): fixed-order sum of the
// split-K partials into the flat gradient, squared-norm partials, a device-wide barrier, then clip + Adam + soft
// update of the same elements.  Element-to-thread mapping, summation order and arithmetic are those of
// splitk_reduce_kernel + grad_sqsum_kernel + adam_soft_update_kernel, so the result is bit-identical to the three.
__global__ void __launch_bounds__(256)
splitk_adam_kernel(ReduceArgs r, float* __restrict__ grad, ParamTable pt, float* __restrict__ m, float* __restrict__ v,
                   int64_t* __restrict__ step_ptr, rloa_naf_hyper hp, float* __restrict__ sq_partial,
                   unsigned* __restrict__ barrier, float* __restrict__ grad_norm_out) {
    __shared__ float red[256];
    __shared__ AdamCoef s_c;
    const int n = pt.offset[14];
    const int n4 = ((reinterpret_cast<uintptr_t>(grad) & 15u) == 0) ? (n >> 2) : 0;
    float s = 0.f;
    for (int i = blockIdx.x * 256 + threadIdx.x; i < n4; i += gridDim.x * 256) {
        float4 g = reinterpret_cast<float4*>(grad)[i];
        g.x = splitk_element(r, 4 * i, g.x); g.y = splitk_element(r, 4 * i + 1, g.y);
        g.z = splitk_element(r, 4 * i + 2, g.z); g.w = splitk_element(r, 4 * i + 3, g.w);
        reinterpret_cast<float4*>(grad)[i] = g;
        const float x0 = g.x * hp.grad_scale, x1 = g.y * hp.grad_scale, x2 = g.z * hp.grad_scale, x3 = g.w * hp.grad_scale;
        s = fmaf(x0, x0, s); s = fmaf(x1, x1, s); s = fmaf(x2, x2, s); s = fmaf(x3, x3, s);
    }
    for (int i = 4 * n4 + blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256) {
        const float g = splitk_element(r, i, grad[i]);
        grad[i] = g;
        const float x = g * hp.grad_scale;
        s = fmaf(x, x, s);
    }
    red[threadIdx.x] = s;
    __syncthreads();
    for (int off = 128; off > 0; off >>= 1) {
        if (threadIdx.x < off) red[threadIdx.x] += red[threadIdx.x + off];
        __syncthreads();
    }
    if (threadIdx.x == 0) {       // device-wide barrier: arrival count + generation (all blocks are co-resident)
        volatile unsigned* gen = barrier + 1;
        const unsigned gen0 = *gen;
        sq_partial[blockIdx.x] = red[0];
        if (blockIdx.x == 0) *step_ptr += 1;                     // optimizer.step() counter
        __threadfence();
        if (atomicAdd(barrier, 1u) == gridDim.x - 1u) {
            barrier[0] = 0u;
            __threadfence();
            atomicAdd(barrier + 1, 1u);
        } else {
            while (*gen == gen0) { }
        }
        __threadfence();
        s_c = adam_coefficients(sq_partial, *reinterpret_cast<volatile int64_t*>(step_ptr), hp);
        if (blockIdx.x == 0 && grad_norm_out != nullptr) *grad_norm_out = s_c.norm;
    }
    __syncthreads();
    const AdamCoef c = s_c;
    for (int i = blockIdx.x * 256 + threadIdx.x; i < n4; i += gridDim.x * 256) {
        const float4 g = reinterpret_cast<const float4*>(grad)[i];
        adam_soft_update_element(pt, 4 * i, g.x, c, m, v, hp);
        adam_soft_update_element(pt, 4 * i + 1, g.y, c, m, v, hp);
        adam_soft_update_element(pt, 4 * i + 2, g.z, c, m, v, hp);
        adam_soft_update_element(pt, 4 * i + 3, g.w, c, m, v, hp);
    }
    for (int i = 4 * n4 + blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256)
        adam_soft_update_element(pt, i, grad[i], c, m, v, hp);
}

__global__ void __launch_bounds__(256) soft_update_kernel(ParamTable pt, float tau) {
    const int n = pt.offset[14];
    for (int i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256) {
        int t = 0;
#pragma unroll
        for (int k = 1; k < 14; k++) t += (i >= pt.offset[k]) ? 1 : 0;
        const int j = i - pt.offset[t];
        pt.target[t][j] = tau * pt.main[t][j] + (1.f - tau) * pt.target[t][j];
    }
}

// [mu | V | L] head weights packed as one [NH][H] matrix for the input-gradient GEMM
__global__ void __launch_bounds__(256) pack_heads_kernel(const float* __restrict__ w_mu, const float* __restrict__ w_v,
                                                         const float* __restrict__ w_l, int A, int NL, int H,
                                                         float* __restrict__ out) {
    const int n = (A + 1 + NL) * H;
    for (int i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256) {
        const int o = i / H, k = i - o * H;
        out[i] = o < A ? w_mu[o * H + k] : (o == A ? w_v[k] : w_l[(o - A - 1) * H + k]);
    }
}

}  // namespace rloa

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
using namespace rloa;

struct rloa_naf_ws {
    int S, A, H, NL, NH, max_batch;
    int trunk_mode = 0;
    float* block = nullptr;
    // per net (0 = main / single, 1 = target)
    float *z1[2], *z2[2];
    float *scale[2][2], *shift[2][2], *mean[2][2], *rstd[2][2];   // [net][layer]
    float *dzh, *da, *y, *loss_part, *splitk, *sq_partial, *wh_pack;
    float *v_tmp;
    float *bn_part;              // [2 nets][R][2][H] chunk statistics / [R][3][H] backward partial sums
    float *w1_part;              // [<= 64 chunks of 64 rows][H][S] partials of dW1 (bn_bwd_apply_dw1_kernel)
    float *bwd_coef;             // [3][H] k1, cB, cC of the BatchNorm backward
    unsigned* tickets;           // [2 * H / 32] last-block-done counters (always return to 0) + [2] device barrier
    size_t splitk_floats;
    int n_loss_parts;
    cudaStream_t side[2] = {nullptr, nullptr};   // weight-gradient contractions run beside the critical path
    cudaEvent_t fork_ev[2] = {nullptr, nullptr}, join_ev[2] = {nullptr, nullptr};
    TrunkTC tc;                  // tcgen05 trunk state
    LearnCluster lc;             // fused learn on two thread-block clusters (naf_learn_cluster.cu); used in trunk mode 1
    bool use_cluster = true;     // RLOA_LEARN_CLUSTER=0 keeps the multi-launch tcgen05 path (A/B measurements)
};

static int naf_nparams(int S, int A, int H) {
    const int NL = A * (A + 1) / 2;
    return H * S + H + H + H + H * H + H + H + H + A * H + A + H + 1 + NL * H + NL;
}

// flat gradient layout = nn.Module.parameters() order (14 tensors)
struct FlatLayout {
    int w1, b1, bn1w, bn1b, w2, b2, bn2w, bn2b, wmu, bmu, wv, bv, wl, bl, total;
};
static FlatLayout flat_layout(int S, int A, int H) {
    const int NL = A * (A + 1) / 2;
    FlatLayout f;
    int o = 0;
    f.w1 = o; o += H * S;
    f.b1 = o; o += H;
    f.bn1w = o; o += H;
    f.bn1b = o; o += H;
    f.w2 = o; o += H * H;
    f.b2 = o; o += H;
    f.bn2w = o; o += H;
    f.bn2b = o; o += H;
    f.wmu = o; o += A * H;
    f.bmu = o; o += A;
    f.wv = o; o += H;
    f.bv = o; o += 1;
    f.wl = o; o += NL * H;
    f.bl = o; o += NL;
    f.total = o;
    return f;
}

extern "C" int rloa_naf_num_params(int32_t S, int32_t A, int32_t H) {
    RLOA_REQUIRE(S >= 1 && A >= 1 && H >= 1, "rloa_naf_num_params: sizes must be positive");
    return naf_nparams(S, A, H);
}

static int splitk_len(int B) { return 128; }
// dW1 partials: one per 64 batch rows, kept for batches up to 4096 rows
static int dw1_chunks(int B) { const int c = (B + kDw1Rows - 1) / kDw1Rows; return c < 64 ? c : 64; }
static int bn_chunks(int B) { return (B + 8 * kBnRowsPerThread - 1) / (8 * kBnRowsPerThread); }
constexpr int kBnSplitMinBatch = 128;    // below this the single-block-per-column kernels are already short
static int splitk_count(int B) { return (B + splitk_len(B) - 1) / splitk_len(B); }

extern "C" void rloa_naf_ws_destroy(rloa_naf_ws* ws);

extern "C" int rloa_naf_ws_create(int32_t S, int32_t A, int32_t H, int32_t max_batch, rloa_naf_ws** out) {
    RLOA_REQUIRE(out != nullptr, "rloa_naf_ws_create: null out");
    RLOA_REQUIRE(S >= 1 && S <= 4096, "rloa_naf_ws_create: state_size out of range");
    // the fused head kernel computes outputs `lane` and `lane + 32`: A + 1 + A (A + 1) / 2 <= 64  <=>  A <= 9
    RLOA_REQUIRE(A >= 1 && A + 1 + A * (A + 1) / 2 <= 64, "rloa_naf_ws_create: 1 <= action_size <= 9 supported (head outputs <= 64)");
    RLOA_REQUIRE(H >= 32 && H % 32 == 0 && H <= 1024, "rloa_naf_ws_create: hidden size must be a multiple of 32 (<= 1024)");
    RLOA_REQUIRE(max_batch >= 1, "rloa_naf_ws_create: max_batch >= 1 required");
    rloa_naf_ws* ws = new (std::nothrow) rloa_naf_ws();
    RLOA_REQUIRE(ws != nullptr, "rloa_naf_ws_create: out of host memory");
    ws->S = S; ws->A = A; ws->H = H; ws->NL = A * (A + 1) / 2; ws->NH = A + 1 + ws->NL; ws->max_batch = max_batch;
    const size_t BH = (size_t)max_batch * H;
    const int nsplit = splitk_count(max_batch);
    ws->splitk_floats = (size_t)nsplit * ((size_t)H * H + (size_t)H * S + (size_t)ws->NH * H);
    ws->n_loss_parts = 148;
    size_t total = 4 * BH            // z1[2], z2[2]
                   + 16 * (size_t)H  // scale/shift/mean/rstd [2][2]
                   + (size_t)max_batch * ws->NH + BH + 2 * (size_t)max_batch + ws->n_loss_parts + ws->splitk_floats +
                   kNormBlocks + (size_t)ws->NH * H +
                   (size_t)bn_chunks(max_batch) * 4 * H + 3 * (size_t)H + 2 * (size_t)(H / 32) + 2 +
                   (size_t)dw1_chunks(max_batch) * H * S;
    if (cudaMalloc(&ws->block, total * sizeof(float)) != cudaSuccess) {
        set_error("rloa_naf_ws_create: cudaMalloc of %zu bytes failed: %s", total * sizeof(float),
                  cudaGetErrorString(cudaGetLastError()));
        delete ws;
        return RLOA_ERR_CUDA;
    }
    float* p = ws->block;
    for (int n = 0; n < 2; n++) { ws->z1[n] = p; p += BH; ws->z2[n] = p; p += BH; }
    for (int n = 0; n < 2; n++)
        for (int l = 0; l < 2; l++) {
            ws->scale[n][l] = p; p += H; ws->shift[n][l] = p; p += H; ws->mean[n][l] = p; p += H; ws->rstd[n][l] = p; p += H;
        }
    ws->dzh = p; p += (size_t)max_batch * ws->NH;
    ws->da = p; p += BH;
    ws->y = p; p += max_batch;
    ws->v_tmp = p; p += max_batch;
    ws->loss_part = p; p += ws->n_loss_parts;
    ws->splitk = p; p += ws->splitk_floats;
    ws->sq_partial = p; p += kNormBlocks;
    ws->wh_pack = p; p += (size_t)ws->NH * H;
    ws->bn_part = p; p += (size_t)bn_chunks(max_batch) * 4 * H;
    ws->bwd_coef = p; p += 3 * (size_t)H;
    ws->w1_part = p; p += (size_t)dw1_chunks(max_batch) * H * S;
    ws->tickets = reinterpret_cast<unsigned*>(p); p += 2 * (size_t)(H / 32) + 2;     // + arrival count, generation
    cudaMemset(ws->tickets, 0, (2 * (size_t)(H / 32) + 2) * sizeof(unsigned));
    const int hb = (ws->NH * (H + 1) + ((ws->NH + 31) & ~31) + kHeadWarps * H + H) * (int)sizeof(float);
    if (hb > 48 * 1024) cudaFuncSetAttribute(naf_head_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, hb);
    trunk_tc_init(&ws->tc);
    {
        const char* e = getenv("RLOA_LEARN_CLUSTER");
        ws->use_cluster = !(e != nullptr && e[0] == '0');
    }
    for (int i = 0; i < 2; i++) {
        if (cudaStreamCreateWithFlags(&ws->side[i], cudaStreamNonBlocking) != cudaSuccess ||
            cudaEventCreateWithFlags(&ws->fork_ev[i], cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&ws->join_ev[i], cudaEventDisableTiming) != cudaSuccess) {
            set_error("rloa_naf_ws_create: stream / event creation failed: %s", cudaGetErrorString(cudaGetLastError()));
            rloa_naf_ws_destroy(ws);
            return RLOA_ERR_CUDA;
        }
    }
    *out = ws;
    return RLOA_OK;
}

extern "C" void rloa_naf_ws_destroy(rloa_naf_ws* ws) {
    if (ws == nullptr) return;
    trunk_tc_free(&ws->tc);
    learn_cluster_free(&ws->lc);
    for (int i = 0; i < 2; i++) {
        if (ws->side[i]) cudaStreamDestroy(ws->side[i]);
        if (ws->fork_ev[i]) cudaEventDestroy(ws->fork_ev[i]);
        if (ws->join_ev[i]) cudaEventDestroy(ws->join_ev[i]);
    }
    if (ws->block) cudaFree(ws->block);
    delete ws;
}

extern "C" int rloa_naf_ws_set_trunk(rloa_naf_ws* ws, int32_t mode) {
    RLOA_REQUIRE(ws != nullptr, "rloa_naf_ws_set_trunk: null workspace");
    RLOA_REQUIRE(mode == 0 || mode == 1, "rloa_naf_ws_set_trunk: mode must be 0 (fp32) or 1 (tcgen05)");
    if (mode == 1) {
        RLOA_REQUIRE(ws->H == 256, "rloa_naf_ws_set_trunk: the tcgen05 trunk is built for hidden = 256");
        int rc = trunk_tc_prepare(&ws->tc, ws->max_batch, ws->H);
        if (rc != RLOA_OK) return rc;
        if (policy_tc_supported(ws->S, ws->A, ws->H)) {
            rc = policy_tc_prepare(&ws->tc);
            if (rc != RLOA_OK) return rc;
        }
        if (ws->use_cluster && learn_cluster_supported(ws->S, ws->A, ws->H, 2)) {
            rc = learn_cluster_prepare(&ws->lc, ws->S, ws->A);
            if (rc != RLOA_OK) return rc;
        }
    }
    ws->trunk_mode = mode;
    return RLOA_OK;
}

extern "C" int rloa_naf_ws_set_debug(rloa_naf_ws* ws, float* buffer, int64_t* stamps) {
    RLOA_REQUIRE(ws != nullptr, "rloa_naf_ws_set_debug: null workspace");
    ws->lc.dbg = buffer;
    ws->lc.prof = reinterpret_cast<long long*>(stamps);
    return RLOA_OK;
}

static size_t head_smem_bytes(const rloa_naf_ws* ws) {
    return (size_t)(ws->NH * (ws->H + 1) + ((ws->NH + 31) & ~31) + kHeadWarps * ws->H + ws->H) * sizeof(float);
}

static int check_params(const rloa_naf_ws* ws, const rloa_naf_params* p, const char* who) {
    if (p == nullptr || p->state_size != ws->S || p->action_size != ws->A || p->hidden != ws->H) {
        set_error("%s: parameter block does not match the workspace (S=%d A=%d H=%d)", who, ws->S, ws->A, ws->H);
        return RLOA_ERR_INVALID;
    }
    return RLOA_OK;
}

// trunk forward for `nets` networks at once: z1 = x W1^T + b1, BN1 coefficients, z2 = relu(bn1(z1)) W2^T + b2,
// BN2 coefficients.  x[n] are the inputs, P[n] the parameter blocks.
static int trunk_forward(rloa_naf_ws* ws, int nets, const rloa_naf_params* const* P, const float* const* x, int B,
                         int train, cudaStream_t st) {
    const int H = ws->H, S = ws->S;
    // train-mode statistics ride in the epilogue of the producing kernel (bn_fuse.cuh) once the batch is large
    // enough to be split over row tiles; eval mode (and tiny batches) use the stand-alone column kernel
    const bool fuse = train && B >= kBnSplitMinBatch;
    const int chunks = bn_chunks(B);
    GemmBatch gb{};
    BnBatch bb{};
    for (int n = 0; n < nets; n++) {
        gb.a[n] = GemmArgs{x[n], S, P[n]->w1, S, ws->z1[n], H, P[n]->b1, nullptr, nullptr, B, H, S, 0};
        bb.a[n] = BnArgs{ws->z1[n], P[n]->bn1_w, P[n]->bn1_b, P[n]->bn1_mean, P[n]->bn1_var, P[n]->bn1_batches,
                         ws->scale[n][0], ws->shift[n][0], ws->mean[n][0], ws->rstd[n][0]};
        gb.bn[n] = BnFuse{bb.a[n], ws->bn_part + (size_t)n * chunks * 2 * H, ws->tickets + n * (H / 32), fuse ? 1 : 0};
    }
    dim3 grid((H + BN - 1) / BN, (B + BM - 1) / BM, nets);
    if (ws->trunk_mode == 1 && trunk_tc_layer1_supported(S, H) && B >= 128) {    // tcgen05 kind::tf32 (naf_trunk_tc.cu)
        const float *xp[2], *w1p[2], *b1p[2];
        float* z1p[2];
        for (int n = 0; n < nets; n++) { xp[n] = x[n]; w1p[n] = P[n]->w1; b1p[n] = P[n]->b1; z1p[n] = ws->z1[n]; }
        const int rc = trunk_tc_layer1(&ws->tc, nets, xp, w1p, b1p, z1p, B, S, H, fuse ? gb.bn : nullptr, st);
        if (rc != RLOA_OK) return rc;
    } else {
        launch_gemm<kNT, false>(gb, nets, grid, st);
        RLOA_LAUNCHED();
    }
    if (!fuse) {
        bn_stats_kernel<<<dim3(H / 32, nets), 256, 0, st>>>(bb, B, H, train);
        RLOA_LAUNCHED();
    }
    for (int n = 0; n < nets; n++) {
        bb.a[n] = BnArgs{ws->z2[n], P[n]->bn2_w, P[n]->bn2_b, P[n]->bn2_mean, P[n]->bn2_var, P[n]->bn2_batches,
                         ws->scale[n][1], ws->shift[n][1], ws->mean[n][1], ws->rstd[n][1]};
        gb.bn[n].bn = bb.a[n];
    }
    if (ws->trunk_mode == 1) {
        const float *z1p[2], *scp[2], *shp[2], *w2p[2], *b2p[2];
        float* z2p[2];
        for (int n = 0; n < nets; n++) {
            z1p[n] = ws->z1[n]; scp[n] = ws->scale[n][0]; shp[n] = ws->shift[n][0];
            w2p[n] = P[n]->w2; b2p[n] = P[n]->b2; z2p[n] = ws->z2[n];
        }
        const int rc = trunk_tc_layer2(&ws->tc, nets, z1p, scp, shp, w2p, b2p, z2p, B, H, fuse ? gb.bn : nullptr, st);
        if (rc != RLOA_OK) return rc;
    } else {
        for (int n = 0; n < nets; n++)
            gb.a[n] = GemmArgs{ws->z1[n], H, P[n]->w2, H, ws->z2[n], H, P[n]->b2, ws->scale[n][0], ws->shift[n][0], B, H, H, 0};
        launch_gemm<kNT, true>(gb, nets, grid, st);
        RLOA_LAUNCHED();
    }
    if (!fuse) {
        bn_stats_kernel<<<dim3(H / 32, nets), 256, 0, st>>>(bb, B, H, train);
        RLOA_LAUNCHED();
    }
    return RLOA_OK;
}

extern "C" int rloa_naf_hidden_layer(rloa_naf_ws* ws, const float* z1, const float* scale, const float* shift,
                                     const float* w2, const float* b2, float* z2, int32_t batch, void* stream) {
    RLOA_REQUIRE(ws && z1 && scale && shift && w2 && b2 && z2, "rloa_naf_hidden_layer: null argument");
    RLOA_REQUIRE(batch >= 1, "rloa_naf_hidden_layer: batch >= 1 required");
    cudaStream_t st = as_stream(stream);
    const int H = ws->H;
    if (ws->trunk_mode == 1) return trunk_tc_layer2(&ws->tc, 1, &z1, &scale, &shift, &w2, &b2, &z2, batch, H, nullptr, st);
    GemmBatch gb{};
    gb.a[0] = GemmArgs{z1, H, w2, H, z2, H, b2, scale, shift, batch, H, H, 0};
    launch_gemm<kNT, true>(gb, 1, dim3((H + BN - 1) / BN, (batch + BM - 1) / BM, 1), st);
    RLOA_LAUNCHED();
    return RLOA_OK;
}

static HeadArgs head_base(const rloa_naf_ws* ws, int net, const rloa_naf_params* p, int B) {
    HeadArgs h{};
    h.z2 = ws->z2[net]; h.scale = ws->scale[net][1]; h.shift = ws->shift[net][1];
    h.w_mu = p->w_mu; h.b_mu = p->b_mu; h.w_v = p->w_v; h.b_v = p->b_v; h.w_l = p->w_l; h.b_l = p->b_l;
    h.B = B; h.H = ws->H; h.A = ws->A; h.NL = ws->NL;
    return h;
}

static int head_blocks(int B) {
    const int need = (B + kHeadWarps - 1) / kHeadWarps;
    return need < 148 ? need : 148;
}

extern "C" int rloa_naf_forward(rloa_naf_ws* ws, const rloa_naf_params* p, const float* states, const float* action,
                                int32_t batch, int32_t train_mode, int32_t trunc_action, float* mu, float* pdiag,
                                float* q, float* v, void* stream) {
    RLOA_REQUIRE(ws != nullptr && states != nullptr, "rloa_naf_forward: null argument");
    RLOA_REQUIRE(batch >= 1 && batch <= ws->max_batch, "rloa_naf_forward: batch exceeds the workspace");
    RLOA_REQUIRE(!(train_mode && batch < 2), "rloa_naf_forward: train-mode BatchNorm needs more than 1 row");
    int rc = check_params(ws, p, "rloa_naf_forward");
    if (rc != RLOA_OK) return rc;
    cudaStream_t st = as_stream(stream);
    const rloa_naf_params* P[1] = {p};
    const float* X[1] = {states};
    rc = trunk_forward(ws, 1, P, X, batch, train_mode, st);
    if (rc != RLOA_OK) return rc;
    HeadArgs h = head_base(ws, 0, p, batch);
    h.action = action; h.trunc_action = trunc_action;
    h.mu = mu; h.pdiag = pdiag; h.q = q; h.v = v;
    naf_head_kernel<<<head_blocks(batch), kHeadWarps * 32, head_smem_bytes(ws), st>>>(h);
    RLOA_LAUNCHED();
    return RLOA_OK;
}

extern "C" int rloa_naf_act(rloa_naf_ws* ws, const rloa_naf_params* p, const float* states, int32_t batch,
                            uint64_t seed, uint64_t step, const uint64_t* step_offset, float noise_scale,
                            float* actions, void* stream) {
    RLOA_REQUIRE(ws != nullptr && states != nullptr && actions != nullptr, "rloa_naf_act: null argument");
    RLOA_REQUIRE(batch >= 1 && batch <= ws->max_batch, "rloa_naf_act: batch exceeds the workspace");
    int rc = check_params(ws, p, "rloa_naf_act");
    if (rc != RLOA_OK) return rc;
    cudaStream_t st = as_stream(stream);
    if (ws->trunk_mode == 1 && ws->tc.policy_image != nullptr)      // one fused tcgen05 launch (naf_policy_tc.cu)
        return policy_tc_act(&ws->tc, p, states, batch, seed, step, step_offset, noise_scale, actions, st);
    const rloa_naf_params* P[1] = {p};
    const float* X[1] = {states};
    rc = trunk_forward(ws, 1, P, X, batch, 0, st);      // qnetwork_main.eval() (naf_algorithm.py:170)
    if (rc != RLOA_OK) return rc;
    HeadArgs h = head_base(ws, 0, p, batch);
    h.act_out = actions; h.seed = seed; h.step = step; h.noise_scale = noise_scale;
    h.step_offset = reinterpret_cast<const unsigned long long*>(step_offset);
    naf_head_kernel<<<head_blocks(batch), kHeadWarps * 32, head_smem_bytes(ws), st>>>(h);
    RLOA_LAUNCHED();
    return RLOA_OK;
}

// ReLU + BatchNorm backward of the main net's layer `layer` (0 / 1): da -> dz in place, gamma / beta / linear-bias grads
static int bn_relu_backward(rloa_naf_ws* ws, float* da, const float* z, const float* bn_w, int layer, int B, float* d_w,
                            float* d_b, float* d_lin_bias, cudaStream_t st) {
    const int H = ws->H;
    if (B < kBnSplitMinBatch) {
        bn_relu_backward_kernel<<<H / 32, 256, 0, st>>>(da, z, bn_w, ws->scale[0][layer], ws->shift[0][layer],
                                                        ws->mean[0][layer], ws->rstd[0][layer], B, H, d_w, d_b, d_lin_bias);
        RLOA_LAUNCHED();
        return RLOA_OK;
    }
    BnBwdCoef coef{ws->bwd_coef, ws->bwd_coef + H, ws->bwd_coef + 2 * H};
    bn_bwd_reduce_kernel<<<dim3(H / 32, bn_chunks(B)), 256, 0, st>>>(da, z, bn_w, ws->scale[0][layer], ws->shift[0][layer],
                                                                     ws->mean[0][layer], ws->rstd[0][layer], B, H, ws->bn_part,
                                                                     ws->tickets, coef, d_w, d_b, d_lin_bias);
    RLOA_LAUNCHED();
    const int blocks = (int)(((size_t)B * H / 4 + 255) / 256);
    bn_bwd_apply_kernel<<<blocks < 1184 ? blocks : 1184, 256, 0, st>>>(da, z, ws->scale[0][layer], ws->shift[0][layer],
                                                                       ws->mean[0][layer], coef, B, H);
    RLOA_LAUNCHED();
    return RLOA_OK;
}

static void fill_param_table(const rloa_naf_params* mn, const rloa_naf_params* tg, ParamTable* pt);

// trunk mode 1 and a batch of at most 1024 rows: the whole of learn() in one cluster launch (naf_learn_cluster.cu)
static bool cluster_learn_applies(const rloa_naf_ws* ws, int batch) {
    return ws->trunk_mode == 1 && ws->use_cluster && ws->lc.images != nullptr && learn_cluster_supported(ws->S, ws->A, ws->H, batch);
}
static int cluster_learn(rloa_naf_ws* ws, const rloa_naf_params* mn, const rloa_naf_params* tg, const rloa_adam_state* adam,
                         const float* states, const float* actions, const float* rewards, const float* next_states,
                         const float* dones, int32_t batch, const rloa_naf_hyper* hp, float* grad, float* loss, float* gnorm,
                         int do_adam, void* stream, const LearnClusterReplay* replay = nullptr, const rloa_xchg* xchg = nullptr) {
    RLOA_REQUIRE(ws && hp && grad && (replay != nullptr || (states && actions && rewards && next_states)), "rloa_naf_learn: null argument");
    RLOA_REQUIRE(!(hp->use_done_mask && dones == nullptr && replay == nullptr), "rloa_naf_learn: use_done_mask needs dones");
    int rc = check_params(ws, mn, "rloa_naf_learn(main)");
    if (rc != RLOA_OK) return rc;
    rc = check_params(ws, tg, "rloa_naf_learn(target)");
    if (rc != RLOA_OK) return rc;
    const FlatLayout fl = flat_layout(ws->S, ws->A, ws->H);
    const int offs[15] = {fl.w1, fl.b1, fl.bn1w, fl.bn1b, fl.w2, fl.b2, fl.bn2w, fl.bn2b, fl.wmu, fl.bmu, fl.wv, fl.bv, fl.wl, fl.bl, fl.total};
    ParamTable pt;
    fill_param_table(mn, tg, &pt);
    return learn_cluster_step(&ws->lc, mn, tg, adam, states, actions, rewards, next_states, dones, batch, hp, pt, offs, grad, loss,
                              gnorm, do_adam, replay, xchg, as_stream(stream));
}

// forward of both nets, loss, backward.  With defer != NULL the final split-K reduction is NOT launched: its
// description is returned instead, for the fused tail kernel of rloa_naf_learn_step.
static int learn_grads_impl(rloa_naf_ws* ws, const rloa_naf_params* mn, const rloa_naf_params* tg,
                            const float* states, const float* actions, const float* rewards,
                            const float* next_states, const float* dones, int32_t batch,
                            const rloa_naf_hyper* hp, float* grad, float* loss, ReduceArgs* defer, void* stream) {
    RLOA_REQUIRE(ws && states && actions && rewards && next_states && hp && grad, "rloa_naf_learn_grads: null argument");
    ws->lc.prepacked = false;        // a pending rloa_naf_learn_prepack belongs to the fused path; this path changes the weights
    RLOA_REQUIRE(batch >= 2 && batch <= ws->max_batch, "rloa_naf_learn_grads: 2 <= batch <= workspace max_batch required");
    RLOA_REQUIRE(!(hp->use_done_mask && dones == nullptr), "rloa_naf_learn_grads: use_done_mask needs dones");
    int rc = check_params(ws, mn, "rloa_naf_learn_grads(main)");
    if (rc != RLOA_OK) return rc;
    rc = check_params(ws, tg, "rloa_naf_learn_grads(target)");
    if (rc != RLOA_OK) return rc;
    cudaStream_t st = as_stream(stream);
    const int B = batch, H = ws->H, S = ws->S, A = ws->A, NH = ws->NH, NL = ws->NL;
    const FlatLayout fl = flat_layout(S, A, H);
    // forward of both nets with train-mode BatchNorm (naf_algorithm.py:194-202):
    // net 0 = main on states, net 1 = target on next_states
    const rloa_naf_params* P[2] = {mn, tg};
    const float* X[2] = {states, next_states};
    rc = trunk_forward(ws, 2, P, X, B, 1, st);
    if (rc != RLOA_OK) return rc;
    const size_t hsm = head_smem_bytes(ws);
    const int hblocks = head_blocks(B);
    const bool fused = H <= 256;
    if (!fused) {   // target head: y = r + gamma V'(s')   (naf_algorithm.py:199)
        HeadArgs h = head_base(ws, 1, tg, B);
        h.reward = rewards; h.done = dones; h.gamma = hp->gamma; h.use_done = hp->use_done_mask; h.y_out = ws->y;
        naf_head_kernel<<<hblocks, kHeadWarps * 32, hsm, st>>>(h);
        RLOA_LAUNCHED();
    }
    {   // main head: Q(s, a), squared TD error, gradients of the head pre-activations; when fused, also the TD
        // target of the same row from the target trunk and da2 = dZh Wh
        HeadArgs h = head_base(ws, 0, mn, B);
        h.action = actions; h.trunc_action = hp->trunc_action; h.dzh = ws->dzh; h.loss_part = ws->loss_part;
        if (fused) {
            h.t_z2 = ws->z2[1]; h.t_scale = ws->scale[1][1]; h.t_shift = ws->shift[1][1]; h.t_wv = tg->w_v; h.t_bv = tg->b_v;
            h.reward = rewards; h.done = dones; h.gamma = hp->gamma; h.use_done = hp->use_done_mask;
            h.da_out = ws->da;
        } else {
            h.y_in = ws->y;
        }
        naf_head_kernel<<<hblocks, kHeadWarps * 32, hsm, st>>>(h);
        RLOA_LAUNCHED();
    }
    // fork: the head bias / weight gradients only feed the optimiser, so they run on a side stream beside the
    // backward chain (in a captured CUDA graph this becomes a parallel branch) and join before the split-K sum
    RLOA_CUDA(cudaEventRecord(ws->fork_ev[0], st));
    RLOA_CUDA(cudaStreamWaitEvent(ws->side[0], ws->fork_ev[0], 0));
    head_bias_grad_kernel<<<NH + 1, 256, 0, ws->side[0]>>>(ws->dzh, B, NH, A, grad + fl.bmu, grad + fl.bv, grad + fl.bl,
                                                           ws->loss_part, hblocks, loss);
    RLOA_LAUNCHED();
    if (!fused) {
        pack_heads_kernel<<<(NH * H + 255) / 256, 256, 0, st>>>(mn->w_mu, mn->w_v, mn->w_l, A, NL, H, ws->wh_pack);
        RLOA_LAUNCHED();
    }
    const int klen = splitk_len(B), nsplit = splitk_count(B);
    float* part_w2 = ws->splitk;
    float* part_w1 = part_w2 + (size_t)nsplit * H * H;
    float* part_wh = part_w1 + (size_t)nsplit * H * S;
    GemmBatch gb{};
    // dWh [NH][H] = dZh^T a2, a2 = relu(bn2(z2)) re-created in the prologue
    gb.a[0] = GemmArgs{ws->dzh, NH, ws->z2[0], H, part_wh, H, nullptr, ws->scale[0][1], ws->shift[0][1], NH, H, B, klen};
    launch_gemm<kTN, true>(gb, 1, dim3((H + BN - 1) / BN, (NH + BM - 1) / BM, nsplit), ws->side[0]);
    RLOA_LAUNCHED();
    {   // the head-weight partials are summed on the side stream too (fixed split order), off the critical path
        ReduceArgs rh{};
        rh.s[0] = ReduceSeg{part_wh, grad + fl.wmu, A * H, nsplit, (size_t)NH * H, fl.wmu};
        rh.s[1] = ReduceSeg{part_wh + (size_t)A * H, grad + fl.wv, H, nsplit, (size_t)NH * H, fl.wv};
        rh.s[2] = ReduceSeg{part_wh + (size_t)(A + 1) * H, grad + fl.wl, NL * H, nsplit, (size_t)NH * H, fl.wl};
        splitk_reduce_kernel<<<dim3(16, 3), 256, 0, ws->side[0]>>>(rh);
        RLOA_LAUNCHED();
    }
    RLOA_CUDA(cudaEventRecord(ws->join_ev[0], ws->side[0]));
    if (!fused) {   // da2 [B][H] = dZh Wh
        gb.a[0] = GemmArgs{ws->dzh, NH, ws->wh_pack, H, ws->da, H, nullptr, nullptr, nullptr, B, H, NH, 0};
        launch_gemm<kNN, false>(gb, 1, dim3((H + BN - 1) / BN, (B + BM - 1) / BM, 1), st);
        RLOA_LAUNCHED();
    }
    // through ReLU + BN2: da -> dz2 in place; bn2 weight/bias and hidden_layer.bias gradients
    rc = bn_relu_backward(ws, ws->da, ws->z2[0], mn->bn2_w, 1, B, grad + fl.bn2w, grad + fl.bn2b, grad + fl.b2, st);
    if (rc != RLOA_OK) return rc;
    // dW2 [H][H] = dz2^T a1, a1 = relu(bn1(z1))
    gb.a[0] = GemmArgs{ws->da, H, ws->z1[0], H, part_w2, H, nullptr, ws->scale[0][0], ws->shift[0][0], H, H, B, klen};
    RLOA_CUDA(cudaEventRecord(ws->fork_ev[1], st));
    RLOA_CUDA(cudaStreamWaitEvent(ws->side[1], ws->fork_ev[1], 0));
    launch_gemm<kTN, true>(gb, 1, dim3((H + BN - 1) / BN, (H + BM - 1) / BM, nsplit), ws->side[1]);
    RLOA_LAUNCHED();
    {
        ReduceArgs r2{};
        r2.s[0] = ReduceSeg{part_w2, grad + fl.w2, H * H, nsplit, (size_t)H * H, fl.w2};
        splitk_reduce_kernel<<<dim3(64, 1), 256, 0, ws->side[1]>>>(r2);
        RLOA_LAUNCHED();
    }
    RLOA_CUDA(cudaEventRecord(ws->join_ev[1], ws->side[1]));
    // da1 [B][H] = dz2 W2 -> reuse z2[1] (the target's z2 is no longer needed) as the output buffer
    float* da1 = ws->z2[1];
    if (ws->trunk_mode == 1) {          // tcgen05: bf16 operands, fp32 accumulate (the bound is stated in the tests)
        rc = trunk_tc_input_grad(&ws->tc, ws->da, mn->w2, da1, B, H, st);
        if (rc != RLOA_OK) return rc;
    } else {
        gb.a[0] = GemmArgs{ws->da, H, mn->w2, H, da1, H, nullptr, nullptr, nullptr, B, H, H, 0};
        launch_gemm<kNN, false>(gb, 1, dim3((H + BN - 1) / BN, (B + BM - 1) / BM, 1), st);
        RLOA_LAUNCHED();
    }
    int w1_split = nsplit;
    if (B >= kBnSplitMinBatch && B <= 4096 && S <= kDw1MaxS) {
        // layer-1 BatchNorm backward: column sums -> coefficients, then dz1 is formed on the fly inside the dW1
        // contraction (it has no other consumer) and never written
        BnBwdCoef coef{ws->bwd_coef, ws->bwd_coef + H, ws->bwd_coef + 2 * H};
        bn_bwd_reduce_kernel<<<dim3(H / 32, bn_chunks(B)), 256, 0, st>>>(da1, ws->z1[0], mn->bn1_w, ws->scale[0][0],
                                                                         ws->shift[0][0], ws->mean[0][0], ws->rstd[0][0], B, H,
                                                                         ws->bn_part, ws->tickets, coef, grad + fl.bn1w,
                                                                         grad + fl.bn1b, grad + fl.b1);
        RLOA_LAUNCHED();
        w1_split = (B + kDw1Rows - 1) / kDw1Rows;
        part_w1 = ws->w1_part;
        bn_bwd_apply_dw1_kernel<<<dim3(w1_split, (H + kDw1Cols - 1) / kDw1Cols), 256, 0, st>>>(da1, ws->z1[0], ws->scale[0][0], ws->shift[0][0], ws->mean[0][0],
                                                         coef, states, B, H, S, part_w1);
        RLOA_LAUNCHED();
    } else {
        rc = bn_relu_backward(ws, da1, ws->z1[0], mn->bn1_w, 0, B, grad + fl.bn1w, grad + fl.bn1b, grad + fl.b1, st);
        if (rc != RLOA_OK) return rc;
        // dW1 [H][S] = dz1^T x
        gb.a[0] = GemmArgs{da1, H, states, S, part_w1, S, nullptr, nullptr, nullptr, H, S, B, klen};
        launch_gemm<kTN, false>(gb, 1, dim3((S + BN - 1) / BN, (H + BM - 1) / BM, nsplit), st);
        RLOA_LAUNCHED();
    }
    // join the side branches, then the fixed-order reduction of the split-K partials into the flat gradient
    RLOA_CUDA(cudaStreamWaitEvent(st, ws->join_ev[0], 0));
    RLOA_CUDA(cudaStreamWaitEvent(st, ws->join_ev[1], 0));
    // only the input-layer partials are left for the main stream (the others were summed on the side streams)
    ReduceArgs ra{};
    ra.s[0] = ReduceSeg{part_w1, grad + fl.w1, H * S, w1_split, (size_t)H * S, fl.w1};
    if (defer != nullptr) {
        *defer = ra;
        return RLOA_OK;
    }
    splitk_reduce_kernel<<<dim3(16, 1), 256, 0, st>>>(ra);
    RLOA_LAUNCHED();
    return RLOA_OK;
}

extern "C" int rloa_naf_learn_grads(rloa_naf_ws* ws, const rloa_naf_params* mn, const rloa_naf_params* tg,
                                    const float* states, const float* actions, const float* rewards,
                                    const float* next_states, const float* dones, int32_t batch,
                                    const rloa_naf_hyper* hp, float* grad, float* loss, void* stream) {
    if (ws != nullptr && cluster_learn_applies(ws, batch))
        return cluster_learn(ws, mn, tg, nullptr, states, actions, rewards, next_states, dones, batch, hp, grad, loss, nullptr, 0, stream);
    return learn_grads_impl(ws, mn, tg, states, actions, rewards, next_states, dones, batch, hp, grad, loss, nullptr, stream);
}

static void fill_param_table(const rloa_naf_params* mn, const rloa_naf_params* tg, ParamTable* pt) {
    const int S = mn->state_size, A = mn->action_size, H = mn->hidden, NL = A * (A + 1) / 2;
    float* m[14] = {mn->w1, mn->b1, mn->bn1_w, mn->bn1_b, mn->w2, mn->b2, mn->bn2_w, mn->bn2_b, mn->w_mu, mn->b_mu,
                    mn->w_v, mn->b_v, mn->w_l, mn->b_l};
    float* t[14] = {tg->w1, tg->b1, tg->bn1_w, tg->bn1_b, tg->w2, tg->b2, tg->bn2_w, tg->bn2_b, tg->w_mu, tg->b_mu,
                    tg->w_v, tg->b_v, tg->w_l, tg->b_l};
    const int sz[14] = {H * S, H, H, H, H * H, H, H, H, A * H, A, H, 1, NL * H, NL};
    int o = 0;
    for (int i = 0; i < 14; i++) {
        pt->main[i] = m[i];
        pt->target[i] = t[i];
        pt->offset[i] = o;
        o += sz[i];
    }
    pt->offset[14] = o;
}

extern "C" int rloa_naf_learn_apply(rloa_naf_ws* ws, const rloa_naf_params* mn, const rloa_naf_params* tg,
                                    const rloa_adam_state* adam, const rloa_naf_hyper* hp, float* grad,
                                    float* grad_norm, void* stream) {
    RLOA_REQUIRE(ws && adam && hp && grad && adam->m && adam->v && adam->step, "rloa_naf_learn_apply: null argument");
    int rc = check_params(ws, mn, "rloa_naf_learn_apply(main)");
    if (rc != RLOA_OK) return rc;
    rc = check_params(ws, tg, "rloa_naf_learn_apply(target)");
    if (rc != RLOA_OK) return rc;
    cudaStream_t st = as_stream(stream);
    ParamTable pt;
    fill_param_table(mn, tg, &pt);
    const int n = pt.offset[14];
    grad_sqsum_kernel<<<kNormBlocks, 256, 0, st>>>(grad, n, hp->grad_scale, ws->sq_partial, adam->step);
    RLOA_LAUNCHED();
    adam_soft_update_kernel<<<(n + 1023) / 1024, 256, 0, st>>>(pt, grad, ws->sq_partial, adam->m, adam->v, adam->step, *hp,
                                                             grad_norm);
    RLOA_LAUNCHED();
    return RLOA_OK;
}

extern "C" int rloa_naf_learn_step(rloa_naf_ws* ws, const rloa_naf_params* mn, const rloa_naf_params* tg,
                                   const rloa_adam_state* adam, const float* states, const float* actions,
                                   const float* rewards, const float* next_states, const float* dones, int32_t batch,
                                   const rloa_naf_hyper* hp, float* grad, float* loss, float* grad_norm, void* stream) {
    RLOA_REQUIRE(adam && adam->m && adam->v && adam->step, "rloa_naf_learn_step: null optimiser state");
    if (ws != nullptr && cluster_learn_applies(ws, batch))
        return cluster_learn(ws, mn, tg, adam, states, actions, rewards, next_states, dones, batch, hp, grad, loss, grad_norm, 1, stream);
    ReduceArgs ra{};
    int rc = learn_grads_impl(ws, mn, tg, states, actions, rewards, next_states, dones, batch, hp, grad, loss, &ra, stream);
    if (rc != RLOA_OK) return rc;
    ParamTable pt;
    fill_param_table(mn, tg, &pt);
    splitk_adam_kernel<<<kNormBlocks, 256, 0, as_stream(stream)>>>(ra, grad, pt, adam->m, adam->v, adam->step, *hp,
                                                                  ws->sq_partial, ws->tickets + 2 * (ws->H / 32), grad_norm);
    RLOA_LAUNCHED();
    return RLOA_OK;
}

struct rloa_xchg;
namespace rloa {
int xchg_exchange_adam(rloa_xchg* x, float* local_grad, const ReduceArgs& deferred, const ParamTable& pt, float* m, float* v,
                       int64_t* step_ptr, const rloa_naf_hyper& hp, float* grad_norm, cudaStream_t st);
}

extern "C" int rloa_naf_learn_apply_xchg(rloa_naf_ws* ws, const rloa_naf_params* mn, const rloa_naf_params* tg,
                                         const rloa_adam_state* adam, const rloa_naf_hyper* hp, rloa_xchg* xchg,
                                         float* grad, float* grad_norm, void* stream) {
    RLOA_REQUIRE(ws && adam && hp && grad && xchg && adam->m && adam->v && adam->step, "rloa_naf_learn_apply_xchg: null argument");
    int rc = check_params(ws, mn, "rloa_naf_learn_apply_xchg(main)");
    if (rc != RLOA_OK) return rc;
    rc = check_params(ws, tg, "rloa_naf_learn_apply_xchg(target)");
    if (rc != RLOA_OK) return rc;
    ParamTable pt;
    fill_param_table(mn, tg, &pt);
    return xchg_exchange_adam(xchg, grad, ReduceArgs{}, pt, adam->m, adam->v, adam->step, *hp, grad_norm, as_stream(stream));
}

extern "C" int rloa_naf_learn_step_xchg(rloa_naf_ws* ws, const rloa_naf_params* mn, const rloa_naf_params* tg,
                                        const rloa_adam_state* adam, rloa_xchg* xchg, const float* states,
                                        const float* actions, const float* rewards, const float* next_states,
                                        const float* dones, int32_t batch, const rloa_naf_hyper* hp, float* grad, float* loss,
                                        float* grad_norm, void* stream) {
    RLOA_REQUIRE(adam && xchg && adam->m && adam->v && adam->step, "rloa_naf_learn_step_xchg: null argument");
    if (ws != nullptr && cluster_learn_applies(ws, batch)) {
        if (xchg_connected_world(xchg) <= kLearnClusterMaxWorld)       // the exchange runs inside the cluster kernel's tail
            return cluster_learn(ws, mn, tg, adam, states, actions, rewards, next_states, dones, batch, hp, grad, loss, grad_norm, 1,
                                 stream, nullptr, xchg);
        int rcc = cluster_learn(ws, mn, tg, nullptr, states, actions, rewards, next_states, dones, batch, hp, grad, loss, nullptr, 0, stream);
        if (rcc != RLOA_OK) return rcc;
        ParamTable ptc;
        fill_param_table(mn, tg, &ptc);
        return xchg_exchange_adam(xchg, grad, ReduceArgs{}, ptc, adam->m, adam->v, adam->step, *hp, grad_norm, as_stream(stream));
    }
    ReduceArgs ra{};
    int rc = learn_grads_impl(ws, mn, tg, states, actions, rewards, next_states, dones, batch, hp, grad, loss, &ra, stream);
    if (rc != RLOA_OK) return rc;
    ParamTable pt;
    fill_param_table(mn, tg, &pt);
    return xchg_exchange_adam(xchg, grad, ra, pt, adam->m, adam->v, adam->step, *hp, grad_norm, as_stream(stream));
}

extern "C" int rloa_naf_learn_fused_supported(const rloa_naf_ws* ws, int32_t batch) {
    return (ws != nullptr && cluster_learn_applies(ws, batch)) ? 1 : 0;
}

extern "C" int rloa_naf_learn_prepack(rloa_naf_ws* ws, const rloa_naf_params* mn, const rloa_naf_params* tg, void* stream) {
    RLOA_REQUIRE(ws != nullptr && mn != nullptr && tg != nullptr, "rloa_naf_learn_prepack: null argument");
    if (!(ws->trunk_mode == 1 && ws->use_cluster && ws->lc.images != nullptr)) return RLOA_OK;      // nothing to prepare on the other paths
    int rc = check_params(ws, mn, "rloa_naf_learn_prepack(main)");
    if (rc != RLOA_OK) return rc;
    rc = check_params(ws, tg, "rloa_naf_learn_prepack(target)");
    if (rc != RLOA_OK) return rc;
    cudaStream_t st = as_stream(stream);
    RLOA_CUDA(cudaEventRecord(ws->lc.pack_fork, st));
    RLOA_CUDA(cudaStreamWaitEvent(ws->side[0], ws->lc.pack_fork, 0));
    rc = learn_cluster_pack(&ws->lc, mn, tg, ws->side[0]);
    if (rc != RLOA_OK) return rc;
    RLOA_CUDA(cudaEventRecord(ws->lc.pack_done, ws->side[0]));
    ws->lc.prepacked = true;
    return RLOA_OK;
}

extern "C" int rloa_naf_learn_step_replay(rloa_naf_ws* ws, const rloa_naf_params* mn, const rloa_naf_params* tg,
                                          const rloa_adam_state* adam, rloa_xchg* xchg, const rloa_replay* rb, uint64_t seed,
                                          uint64_t draw, const uint64_t* draw_offset, int32_t batch, const rloa_naf_hyper* hp,
                                          float* grad, float* loss, float* grad_norm, void* stream) {
    RLOA_REQUIRE(ws && adam && adam->m && adam->v && adam->step && rb && hp, "rloa_naf_learn_step_replay: null argument");
    RLOA_REQUIRE(cluster_learn_applies(ws, batch),
                 "rloa_naf_learn_step_replay: needs trunk mode 1 and a batch the fused kernel supports (rloa_naf_learn_fused_supported); "
                 "use rloa_replay_sample + rloa_naf_learn_step otherwise");
    RLOA_REQUIRE(rb->state_size == ws->S && rb->action_size == ws->A, "rloa_naf_learn_step_replay: replay row layout does not match the network");
    const LearnClusterReplay rp{rb, seed, draw, draw_offset};
    if (xchg == nullptr || xchg_connected_world(xchg) <= kLearnClusterMaxWorld)
        return cluster_learn(ws, mn, tg, adam, nullptr, nullptr, nullptr, nullptr, nullptr, batch, hp, grad, loss, grad_norm, 1, stream, &rp, xchg);
    int rc = cluster_learn(ws, mn, tg, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, batch, hp, grad, loss, nullptr, 0, stream, &rp);
    if (rc != RLOA_OK) return rc;
    ParamTable pt;
    fill_param_table(mn, tg, &pt);
    return xchg_exchange_adam(xchg, grad, ReduceArgs{}, pt, adam->m, adam->v, adam->step, *hp, grad_norm, as_stream(stream));
}

extern "C" int rloa_naf_learn_step_pending(rloa_naf_ws* ws, const rloa_naf_params* mn, const rloa_naf_params* tg,
                                           const rloa_adam_state* adam, rloa_xchg* xchg, const rloa_replay* rb, uint64_t seed,
                                           uint64_t draw, const uint64_t* draw_offset, int32_t batch, const rloa_naf_hyper* hp,
                                           int32_t n_pending, const float* states, const float* actions, const float* rewards,
                                           const float* next_states, const uint8_t* dones, const uint8_t* valid, float* grad,
                                           float* loss, float* grad_norm, void* stream) {
    RLOA_REQUIRE(ws && adam && adam->m && adam->v && adam->step && rb && hp, "rloa_naf_learn_step_pending: null argument");
    RLOA_REQUIRE(cluster_learn_applies(ws, batch) && (xchg == nullptr || xchg_connected_world(xchg) <= kLearnClusterMaxWorld),
                 "rloa_naf_learn_step_pending: needs the one-kernel update (rloa_naf_learn_fused_supported, at most 8 ranks)");
    RLOA_REQUIRE(rb->state_size == ws->S && rb->action_size == ws->A, "rloa_naf_learn_step_pending: replay row layout does not match the network");
    RLOA_REQUIRE(n_pending >= 1 && n_pending <= kLearnClusterMaxPending && n_pending <= rb->capacity && states && actions && rewards &&
                     next_states, "rloa_naf_learn_step_pending: bad pending rows");
    LearnClusterReplay rp{rb, seed, draw, draw_offset};
    rp.pd_states = states; rp.pd_actions = actions; rp.pd_rewards = rewards; rp.pd_next_states = next_states;
    rp.pd_dones = dones; rp.pd_valid = valid; rp.pd_n = n_pending;
    return cluster_learn(ws, mn, tg, adam, nullptr, nullptr, nullptr, nullptr, nullptr, batch, hp, grad, loss, grad_norm, 1, stream, &rp, xchg);
}

extern "C" int rloa_naf_soft_update(const rloa_naf_params* mn, const rloa_naf_params* tg, float tau, void* stream) {
    RLOA_REQUIRE(mn && tg, "rloa_naf_soft_update: null argument");
    RLOA_REQUIRE(mn->state_size == tg->state_size && mn->action_size == tg->action_size && mn->hidden == tg->hidden,
                 "rloa_naf_soft_update: parameter blocks differ in shape");
    ParamTable pt;
    fill_param_table(mn, tg, &pt);
    const int n = pt.offset[14];
    soft_update_kernel<<<(n + 1023) / 1024, 256, 0, as_stream(stream)>>>(pt, tau);
    RLOA_LAUNCHED();
    return RLOA_OK;
}
