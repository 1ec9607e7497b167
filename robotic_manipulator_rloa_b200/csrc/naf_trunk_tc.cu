// tcgen05 tensor-core trunk: the hidden_layer contraction of the NAF network on the 5th-generation
// tensor cores (reference naf_components/naf_neural_network.py:77-78, `hidden_layer(relu(bn1(.)))`).
//
//   z2[m][n] = sum_k bf16(relu(z1[m][k] * scale[k] + shift[k])) * bf16(W2[n][k]) + b2[n]      (fp32 accumulate)
//
// One CTA owns a 128 x 64 output tile (UMMA M = 128, N = 64, cta_group::1) and the full K = 256:
//   1. all 4 warps read the fp32 operands from global (rows are contiguous: coalesced 1 KB per warp load),
//      apply the BatchNorm + ReLU prologue, round to bf16 and store 16-byte chunks into shared memory in the
//      canonical K-major SWIZZLE_128B layout (row pitch 128 B = 64 bf16, 8-row atoms of 1 KB, chunk index
//      XOR row % 8) — the layout TMA would have produced, written by hand because the prologue has to touch
//      every element anyway;
//   2. fence.proxy.async makes those generic-proxy stores visible to the tensor core's async proxy;
//   3. one elected thread issues 16 tcgen05.mma (kind::f16, 128 x 64 x 16 each) walking K through the shared
//      memory descriptors; the accumulator lives in 64 TMEM columns; tcgen05.commit arrives on an mbarrier;
//   4. every warp reads its 32 TMEM lanes (tcgen05.ld 32x32b), adds the bias and writes fp32 rows.
// W2 is converted from the fp32 master weights inside the kernel (it changes every update), so there is no
// bf16 shadow copy to keep coherent.  sm_100a only.
#include "common.cuh"
#include "naf_trunk_tc.cuh"

#include "tc_common.cuh"

namespace rloa {

constexpr int kTcM = 128;                     // rows per CTA (UMMA M)
constexpr int kTcN = 64;                      // output features per CTA (UMMA N)
constexpr int kTcK = 256;                     // contraction length (hidden)
constexpr int kTcThreads = 256;              // 8 warps stage the operands; warps 0-3 read the accumulator back
constexpr int kTcKBlocks = kTcK / 64;         // 128-byte swizzle rows hold 64 bf16
constexpr uint32_t kABlockBytes = kTcM * 128; // one k-block of A: 128 rows x 128 B
constexpr uint32_t kBBlockBytes = kTcN * 128;
constexpr uint32_t kTcSmemBytes = kTcKBlocks * (kABlockBytes + kBBlockBytes) + 1024 /*alignment slack*/ + 64;
constexpr uint32_t kTmemCols = 64;

constexpr uint32_t kIdesc = tc::idesc_bf16_f32(kTcM, kTcN);
using tc::smem_u32;
using tc::umma_desc_sw128;

struct TcNet {
    const float* z1;            // [B][256]
    const float *scale, *shift; // [256] BatchNorm1 folded coefficients
    const float* w2;            // [256][256] hidden_layer.weight (out, in)
    const float* b2;            // [256]
    float* z2;                  // [B][256]
};
struct TcBatch {
    TcNet n[2];
    BnFuse bn[2];               // train-mode BatchNorm statistics of z2, fused into the epilogue (bn_fuse.cuh)
};

// 8 consecutive fp32 -> one 16-byte chunk of bf16
template <bool PRO>
__device__ __forceinline__ uint4 pack8(const float4 a, const float4 b, const float* sc, const float* sh) {
    float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    if (PRO) {
#pragma unroll
        for (int i = 0; i < 8; i++) v[i] = fmaxf(fmaf(v[i], sc[i], sh[i]), 0.f);
    }
    __nv_bfloat162 p0 = __floats2bfloat162_rn(v[0], v[1]), p1 = __floats2bfloat162_rn(v[2], v[3]);
    __nv_bfloat162 p2 = __floats2bfloat162_rn(v[4], v[5]), p3 = __floats2bfloat162_rn(v[6], v[7]);
    uint4 r;
    r.x = *reinterpret_cast<uint32_t*>(&p0); r.y = *reinterpret_cast<uint32_t*>(&p1);
    r.z = *reinterpret_cast<uint32_t*>(&p2); r.w = *reinterpret_cast<uint32_t*>(&p3);
    return r;
}

// MODE 0 (forward):  C = bf16(relu(A * scale + shift)) @ bf16(W)^T + bias,  W given as [n][k]  (+ fused BN statistics)
// MODE 1 (backward): C = bf16(A) @ bf16(W),                                 W given as [k][n]: the input gradient
//                    da1 = dz2 W2 of reference naf_algorithm.py:208 (autograd of hidden_layer), no bias
// MODE 2 (layer 1):  C = (tf32_hi(A) + tf32_lo(A)) @ tf32(W)^T + bias with A = the S-wide observations (S <= 24),
//                    W = input_layer.weight [n][S]  (+ fused BN statistics).  kind::tf32, 2 x 3 MMAs of K = 8: the
//                    inputs keep 2^-22 relative precision through the hi / lo split, the weights are rounded to tf32.
//                    Field use: z1 = observations, w2 = W1, b2 = b1, z2 = the layer-1 pre-activations.
template <int MODE>
__global__ void __launch_bounds__(kTcThreads) trunk_tc_layer2_kernel(const __grid_constant__ TcBatch batch, int B, int S) {
    extern __shared__ uint8_t tc_smem_raw[];
    const TcNet& g = batch.n[blockIdx.z];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int m0 = blockIdx.x * kTcM, n0 = blockIdx.y * kTcN;

    const uint32_t raw = smem_u32(tc_smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;              // SWIZZLE_128B atoms need 1024-byte alignment
    uint8_t* sm = tc_smem_raw + (base - raw);
    const uint32_t a_base = base, b_base = base + kTcKBlocks * kABlockBytes;
    uint8_t* sm_a = sm;
    uint8_t* sm_b = sm + kTcKBlocks * kABlockBytes;
    uint64_t* bar = reinterpret_cast<uint64_t*>(sm + kTcKBlocks * (kABlockBytes + kBBlockBytes));
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"(kTmemCols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }

    if (MODE == 2) {
        // ---- layer-1 operands (tf32, K-major, one 128-byte swizzle row = 32 fp32 per matrix row): x_hi | x_lo | W1 ----
        uint8_t *sm_xh = sm_a, *sm_xl = sm_a + kABlockBytes;
        {
            const int r = tid & 127, h = tid >> 7;             // row, which 16 of the 32 k
            const int row = m0 + r;
            float x[16];
#pragma unroll
            for (int i = 0; i < 16; i++) {
                const int k = h * 16 + i;
                x[i] = (k < S && row < B) ? g.z1[(size_t)row * S + k] : 0.f;
            }
#pragma unroll
            for (int c = 0; c < 4; c++) {
                float hi[4], lo[4];
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    hi[i] = tc::to_tf32(x[4 * c + i]);
                    lo[i] = tc::to_tf32(x[4 * c + i] - hi[i]);
                }
                const uint32_t off = tc::sw128_chunk_offset(r, h * 4 + c);
                *reinterpret_cast<float4*>(sm_xh + off) = make_float4(hi[0], hi[1], hi[2], hi[3]);
                *reinterpret_cast<float4*>(sm_xl + off) = make_float4(lo[0], lo[1], lo[2], lo[3]);
            }
        }
        {
            const int n = tid & 63, q = tid >> 6;              // weight row, which 8 of the 32 k
            float w[8];
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const int k = q * 8 + i;
                w[i] = k < S ? tc::to_tf32(g.w2[(size_t)(n0 + n) * S + k]) : 0.f;
            }
#pragma unroll
            for (int c = 0; c < 2; c++)
                *reinterpret_cast<float4*>(sm_b + tc::sw128_chunk_offset(n, q * 2 + c)) =
                    make_float4(w[4 * c], w[4 * c + 1], w[4 * c + 2], w[4 * c + 3]);
        }
    } else {
    // ---- operands -> shared memory (bf16, K-major, 128-byte swizzle) ----
    // lane owns the 16-byte chunk (8 bf16) k = 8 lane .. 8 lane + 7 of a row: k-block lane / 8, chunk lane % 8
    const int kb = lane >> 3, ch = lane & 7;
    float sc[8], sh[8];
#pragma unroll
    for (int i = 0; i < 8; i++) {
        sc[i] = MODE == 0 ? g.scale[lane * 8 + i] : 1.f;
        sh[i] = MODE == 0 ? g.shift[lane * 8 + i] : 0.f;
    }
    // 8 rows per warp per batch, every load of a batch issued before the first conversion (memory-level
    // parallelism: the kernel is one L2 round trip per batch, not per row)
    constexpr int kWarps = kTcThreads / 32, kBatch = 8;
#pragma unroll 1
    for (int r0 = warp; r0 < kTcM; r0 += kWarps * kBatch) {
        float4 lo[kBatch], hi[kBatch];
#pragma unroll
        for (int b = 0; b < kBatch; b++) {
            const int row = m0 + r0 + b * kWarps;
            lo[b] = hi[b] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (row < B) {
                const float4* src = reinterpret_cast<const float4*>(g.z1 + (size_t)row * kTcK + lane * 8);
                lo[b] = src[0];
                hi[b] = src[1];
            }
        }
#pragma unroll
        for (int b = 0; b < kBatch; b++) {
            const int r = r0 + b * kWarps;
            uint4 packed = make_uint4(0u, 0u, 0u, 0u);
            if (m0 + r < B) packed = pack8<MODE == 0>(lo[b], hi[b], sc, sh);
            const uint32_t off = kb * kABlockBytes + (r >> 3) * 1024 + (r & 7) * 128 + ((ch ^ (r & 7)) << 4);
            *reinterpret_cast<uint4*>(sm_a + off) = packed;
        }
    }
    if (MODE == 0) {
        float4 lo[kTcN / kWarps], hi[kTcN / kWarps];
#pragma unroll
        for (int b = 0; b < kTcN / kWarps; b++) {
            const float4* src = reinterpret_cast<const float4*>(g.w2 + (size_t)(n0 + warp + b * kWarps) * kTcK + lane * 8);
            lo[b] = src[0];
            hi[b] = src[1];
        }
#pragma unroll
        for (int b = 0; b < kTcN / kWarps; b++) {
            const int r = warp + b * kWarps;
            const uint32_t off = kb * kBBlockBytes + (r >> 3) * 1024 + (r & 7) * 128 + ((ch ^ (r & 7)) << 4);
            *reinterpret_cast<uint4*>(sm_b + off) = pack8<false>(lo[b], hi[b], nullptr, nullptr);
        }
    } else {
        // W is [k][n]: the UMMA B operand wants (n, k) K-major, so the slice W[0..255][n0..n0+63] is transposed while
        // it is staged — half a warp reads one k row (64 floats), each lane scatters its 4 n values as bf16
        const int nq = (lane & 15) * 4;
#pragma unroll 1
        for (int k0 = warp * 2 + (lane >> 4); k0 < kTcK; k0 += 8 * 16) {
            float4 w[8];
#pragma unroll
            for (int u = 0; u < 8; u++) w[u] = *reinterpret_cast<const float4*>(g.w2 + (size_t)(k0 + 16 * u) * kTcK + n0 + nq);
#pragma unroll
            for (int u = 0; u < 8; u++) {
                const int k = k0 + 16 * u;
                const float vals[4] = {w[u].x, w[u].y, w[u].z, w[u].w};
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    const int n = nq + c;
                    const uint32_t off = (k >> 6) * kBBlockBytes + (n >> 3) * 1024 + (n & 7) * 128 +
                                         ((((k & 63) >> 3) ^ (n & 7)) << 4) + (k & 7) * 2;
                    *reinterpret_cast<__nv_bfloat16*>(sm_b + off) = __float2bfloat16_rn(vals[c]);
                }
            }
        }
    }
    }
    // generic-proxy writes -> visible to the async proxy (tensor core operand fetch)
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tmem_slot;

    // ---- one thread issues the 16 MMAs of the tile, then commits to the mbarrier ----
    if (tid == 0 && MODE == 2) {
        constexpr uint32_t idesc = tc::idesc_tf32_f32(kTcM, kTcN);
#pragma unroll
        for (int k = 0; k < 3; k++) {                          // K = 24 in steps of 8 (32 bytes inside the swizzle row)
            tc::mma_tf32(tmem, umma_desc_sw128(a_base + k * 32), umma_desc_sw128(b_base + k * 32), idesc, k > 0);
            tc::mma_tf32(tmem, umma_desc_sw128(a_base + kABlockBytes + k * 32), umma_desc_sw128(b_base + k * 32), idesc, true);
        }
        tc::mma_commit(bar);
    }
    if (tid == 0 && MODE != 2) {
#pragma unroll
        for (int k = 0; k < kTcK / 16; k++) {
            // k-block k / 4 (64 bf16 = one 128-byte swizzle row), 32-byte step k % 4 inside the swizzle atom
            const uint64_t adesc = umma_desc_sw128(a_base + (k >> 2) * kABlockBytes + (k & 3) * 32);
            const uint64_t bdesc = umma_desc_sw128(b_base + (k >> 2) * kBBlockBytes + (k & 3) * 32);
            const uint32_t accumulate = k > 0 ? 1u : 0u;
            asm volatile(
                "{\n\t"
                ".reg .pred p;\n\t"
                "setp.ne.b32 p, %4, 0;\n\t"
                "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
                "}\n" ::"r"(tmem),
                "l"(adesc), "l"(bdesc), "r"(kIdesc), "r"(accumulate)
                : "memory");
        }
        // commit: the mbarrier completes when every MMA issued above has finished (implies before_thread_sync)
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                     : "memory");
    }
    __syncwarp();
    {   // wait for phase 0 of the mbarrier
        uint32_t done = 0;
        const uint32_t bar_addr = smem_u32(bar);
        while (!done) {
            asm volatile(
                "{\n\t"
                ".reg .pred p;\n\t"
                "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                "selp.u32 %0, 1, 0, p;\n\t"
                "}\n"
                : "=r"(done)
                : "r"(bar_addr), "r"(0u)
                : "memory");
        }
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    // ---- epilogue: TMEM lane = tile row; warp w < 4 owns lanes 32 w .. 32 w + 31, columns 0-31 and 32-63 go to
    // warps w and w + 4 (a warp may only touch the TMEM lanes 32 (w % 4) .. + 31) ----
    const int row = m0 + (warp & 3) * 32 + lane;
    float* dst = g.z2 + (size_t)row * kTcK + n0;
    {
        const int half = warp >> 2;
        uint32_t v[32];
        const uint32_t taddr = tmem + ((uint32_t)((warp & 3) * 32) << 16) + half * 32;
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
              "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
              "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
              "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
            : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        const bool fuse = MODE != 1 && batch.bn[blockIdx.z].enabled != 0;
        float* tile = reinterpret_cast<float*>(sm_a);          // [128][65] fp32: the operand tiles are dead after the MMAs
        const int trow = (warp & 3) * 32 + lane;
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
            const float4 bias = MODE != 1 ? *reinterpret_cast<const float4*>(g.b2 + n0 + half * 32 + j)
                                          : make_float4(0.f, 0.f, 0.f, 0.f);
            float4 o;
            o.x = __uint_as_float(v[j]) + bias.x;
            o.y = __uint_as_float(v[j + 1]) + bias.y;
            o.z = __uint_as_float(v[j + 2]) + bias.z;
            o.w = __uint_as_float(v[j + 3]) + bias.w;
            if (row < B) *reinterpret_cast<float4*>(dst + half * 32 + j) = o;
            if (fuse) {
                float* tr = tile + trow * 65 + half * 32 + j;
                tr[0] = o.x; tr[1] = o.y; tr[2] = o.z; tr[3] = o.w;
            }
        }
    }
    if (MODE != 1 && batch.bn[blockIdx.z].enabled) {
        // ---- BatchNorm statistics of this 128 x 64 tile: column mean and centred second moment over its valid rows,
        // then the last row tile of the column group merges all partials and finalises (bn_fuse.cuh) ----
        const BnFuse& f = batch.bn[blockIdx.z];
        const float* tile = reinterpret_cast<const float*>(sm_a);
        float* red = reinterpret_cast<float*>(sm_a) + 128 * 65;     // [4][64] + [64]
        __shared__ unsigned s_last;
        const int rows = min(kTcM, B - m0), c = tid & 63, grp = tid >> 6;
        __syncthreads();
        float s = 0.f;
        for (int r = grp * 32; r < grp * 32 + 32; r++) s += r < rows ? tile[r * 65 + c] : 0.f;
        red[grp * 64 + c] = s;
        __syncthreads();
        if (tid < 64) red[256 + tid] = (red[tid] + red[64 + tid] + red[128 + tid] + red[192 + tid]) / (float)rows;
        __syncthreads();
        const float mean_c = red[256 + c];
        s = 0.f;
        for (int r = grp * 32; r < grp * 32 + 32; r++) {
            const float d = r < rows ? tile[r * 65 + c] - mean_c : 0.f;
            s = fmaf(d, d, s);
        }
        __syncthreads();
        red[grp * 64 + c] = s;
        __syncthreads();
        const int chunks = gridDim.x;
        if (tid < 64) {
            float* P = f.part + (size_t)blockIdx.x * 2 * kTcK;
            P[n0 + tid] = mean_c;
            P[kTcK + n0 + tid] = red[tid] + red[64 + tid] + red[128 + tid] + red[192 + tid];
        }
        __threadfence();
        __syncthreads();
        if (tid == 0) {
            const unsigned t = atomicAdd(&f.ticket[blockIdx.y], 1u);
            s_last = (t == (unsigned)chunks - 1u) ? 1u : 0u;
            if (s_last) f.ticket[blockIdx.y] = 0u;
        }
        __syncthreads();
        if (s_last && tid < 64) {
            __threadfence();
            bn_finalize_column(f.bn, f.part, chunks, kTcM, B, kTcK, n0 + tid, blockIdx.y == 0 && tid == 0);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kTmemCols));
    }
}

void trunk_tc_init(TrunkTC* t) { *t = TrunkTC{}; }
void trunk_tc_free(TrunkTC* t) {
    if (t->policy_image != nullptr) cudaFree(t->policy_image);
    *t = TrunkTC{};
}

int trunk_tc_prepare(TrunkTC* t, int max_batch, int H) {
    RLOA_REQUIRE(H == kTcK, "tcgen05 trunk: hidden size must be 256");
    int dev = 0, major = 0;
    RLOA_CUDA(cudaGetDevice(&dev));
    RLOA_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
    RLOA_REQUIRE(major == 10, "tcgen05 trunk: needs an sm_100 device");
    RLOA_CUDA(cudaFuncSetAttribute(trunk_tc_layer2_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTcSmemBytes));
    RLOA_CUDA(cudaFuncSetAttribute(trunk_tc_layer2_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTcSmemBytes));
    RLOA_CUDA(cudaFuncSetAttribute(trunk_tc_layer2_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTcSmemBytes));
    t->max_batch = max_batch;
    t->ready = true;
    return RLOA_OK;
}

int trunk_tc_layer2(TrunkTC* t, int nets, const float* const* z1, const float* const* scale, const float* const* shift,
                    const float* const* w2, const float* const* b2, float* const* z2, int B, int H, const BnFuse* bn,
                    cudaStream_t st) {
    RLOA_REQUIRE(t->ready, "tcgen05 trunk: rloa_naf_ws_set_trunk(1) was not called");
    RLOA_REQUIRE(H == kTcK && nets >= 1 && nets <= 2, "tcgen05 trunk: hidden = 256 and 1..2 networks per launch");
    TcBatch tb{};
    for (int n = 0; n < nets; n++) {
        tb.n[n] = TcNet{z1[n], scale[n], shift[n], w2[n], b2[n], z2[n]};
        if (bn != nullptr) tb.bn[n] = bn[n];
    }
    dim3 grid((B + kTcM - 1) / kTcM, kTcK / kTcN, nets);
    trunk_tc_layer2_kernel<0><<<grid, kTcThreads, kTcSmemBytes, st>>>(tb, B, 0);
    RLOA_LAUNCHED();
    return RLOA_OK;
}

bool trunk_tc_layer1_supported(int S, int H) { return H == kTcK && S >= 1 && S <= 24; }

int trunk_tc_layer1(TrunkTC* t, int nets, const float* const* x, const float* const* w1, const float* const* b1,
                    float* const* z1, int B, int S, int H, const BnFuse* bn, cudaStream_t st) {
    RLOA_REQUIRE(t->ready, "tcgen05 trunk: rloa_naf_ws_set_trunk(1) was not called");
    RLOA_REQUIRE(trunk_tc_layer1_supported(S, H) && nets >= 1 && nets <= 2, "tcgen05 layer 1: hidden = 256, S <= 24, 1..2 networks");
    TcBatch tb{};
    for (int n = 0; n < nets; n++) {
        tb.n[n] = TcNet{x[n], nullptr, nullptr, w1[n], b1[n], z1[n]};
        if (bn != nullptr) tb.bn[n] = bn[n];
    }
    dim3 grid((B + kTcM - 1) / kTcM, kTcK / kTcN, nets);
    trunk_tc_layer2_kernel<2><<<grid, kTcThreads, kTcSmemBytes, st>>>(tb, B, S);
    RLOA_LAUNCHED();
    return RLOA_OK;
}

int trunk_tc_input_grad(TrunkTC* t, const float* dz, const float* w, float* da, int B, int H, cudaStream_t st) {
    RLOA_REQUIRE(t->ready, "tcgen05 trunk: rloa_naf_ws_set_trunk(1) was not called");
    RLOA_REQUIRE(H == kTcK, "tcgen05 trunk: hidden = 256");
    TcBatch tb{};
    tb.n[0] = TcNet{dz, nullptr, nullptr, w, nullptr, da};
    dim3 grid((B + kTcM - 1) / kTcM, kTcK / kTcN, 1);
    trunk_tc_layer2_kernel<1><<<grid, kTcThreads, kTcSmemBytes, st>>>(tb, B, 0);
    RLOA_LAUNCHED();
    return RLOA_OK;
}

}  // namespace rloa
