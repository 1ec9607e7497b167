// Fused tcgen05 policy kernel: NAFAgent.act for 128 arms per CTA in ONE launch
// (reference naf_components/naf_algorithm.py:158-178 = eval-mode NAF.forward, naf_neural_network.py:76-87,
// 95-102, 119-121: trunk, mu head, diagonal of L, noise, clamp).
//
//   policy_pack_kernel   folds the eval-mode BatchNorms into per-column (scale, shift) vectors and writes the
//                        "weight image": hidden_layer.weight and the 2A head rows that act() needs (mu, diag L)
//                        as bf16, input_layer.weight as tf32, all in the exact shared-memory byte order of a K-major
//                        SWIZZLE_128B UMMA operand.
//   policy_act_tc_kernel per 128-row tile:
//     0. one thread issues the 1-D TMA bulk copies (cp.async.bulk, mbarrier complete_tx): W1 image 32 KB + vectors,
//        W2 image 128 KB, head image 16 KB;
//     1. layer 1 on the tensor core as kind::tf32 (K = S <= 24 -> three 128x256x8 steps).  The observation carries raw
//        joint angles, so it is NOT rounded to tf32: every thread splits its values into x = hi + lo (both tf32,
//        |x - hi - lo| <= 2^-22 |x|) and the MMA runs twice into the same accumulator; only the weights see the 10-bit
//        mantissa.  (The first version did this layer in fp32 on the CUDA cores: 50 % of the kernel's time in ncu.)
//     2. epilogue 1: tcgen05.ld, relu(bn1(.)) -> bf16 -> A operand tile, hand-swizzled (it overwrites the layer-1
//        operands, which are dead);
//     3. layer 2: 16 x tcgen05.mma 128x256x16 (bf16 -> fp32 TMEM, the same 256 columns), commit -> mbarrier;
//     4. epilogue 2: relu(bn2(.)) -> bf16 -> the same A tile;
//     5. heads: 16 x tcgen05.mma 128x32x16 into TMEM columns 0-31;
//     6. epilogue: mu = tanh, l_kk = tanh, action = clamp(mu + noise_scale exp(-l_kk) eps, -1, 1) with the same
//        Philox4x32-10 keying as the fp32 head kernel (seed, step + *step_offset, row, k).
// Shared memory: A tile 64 KB (first: W1 32 KB | x_hi 16 KB | x_lo 16 KB) + W2 128 KB + heads 16 KB + vectors 4.2 KB.
#include "common.cuh"
#include "naf_trunk_tc.cuh"
#include "philox.cuh"
#include "tc_common.cuh"

namespace rloa {

using namespace tc;

constexpr int kPolH = 256;                     // hidden width the kernel is built for
constexpr int kPolSP = 24;                     // largest state width (KUKA 21, Panda 23): three tf32 k-steps of 8
constexpr int kPolKP = 32;                     // K of the layer-1 operand tiles: one 128-byte swizzle row of fp32
constexpr int kPolNH = 32;                     // head rows in the image: [0,A) mu, [A,2A) diag L, rest zero
constexpr int kPolThreads = 512;               // 16 warps: thread = (tile row, quarter of the 256 hidden columns)
constexpr uint32_t kPolW2Bytes = kPolH * kPolH * 2;            // 131072
constexpr uint32_t kPolWhBytes = kPolNH * kPolH * 2;           // 16384
constexpr uint32_t kPolVecFloats = 4 * kPolH + kPolNH;         // sc1 sh1 sc2 sh2 bh
constexpr uint32_t kPolVecBytes = kPolVecFloats * 4;           // 4224
constexpr uint32_t kPolW1Bytes = kPolH * kPolKP * 4;           // 32768 (tf32 [256 rows][128 B], swizzled)
constexpr uint32_t kPolXBytes = 128 * kPolKP * 4;              // 16384 per observation tile (hi, lo)
constexpr uint32_t kPolOffWh = kPolW2Bytes;
constexpr uint32_t kPolOffW1 = kPolOffWh + kPolWhBytes;        // W1 and the vectors are contiguous: one bulk copy
constexpr uint32_t kPolOffVec = kPolOffW1 + kPolW1Bytes;
constexpr uint32_t kPolImageBytes = kPolOffVec + kPolVecBytes;
constexpr uint32_t kPolABytes = 128 * kPolH * 2;               // 65536
constexpr uint32_t kPolSmemBytes = kPolABytes + kPolW2Bytes + kPolWhBytes + kPolVecBytes + 64 + 1024;
static_assert(kPolW1Bytes + 2 * kPolXBytes <= kPolABytes, "layer-1 operands live in the A tile");
constexpr uint32_t kPolTmemCols = 256;

size_t policy_image_bytes() { return kPolImageBytes; }

// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) policy_pack_kernel(rloa_naf_params p, uint8_t* __restrict__ image) {
    const int S = p.state_size, A = p.action_size;
    const int t = blockIdx.x * 256 + threadIdx.x;
    // (a) W2: one thread per 16-byte chunk (8 consecutive k of one output row n)
    if (t < kPolH * kPolH / 8) {
        const int n = t >> 5, kc = t & 31;             // 32 chunks per row
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; i++) v[i] = p.w2[(size_t)n * kPolH + kc * 8 + i];
        const uint32_t off = (uint32_t)(kc >> 3) * (kPolH * 128) + sw128_chunk_offset(n, kc & 7);
        *reinterpret_cast<uint4*>(image + off) = pack8_bf16(v);
        return;
    }
    int u = t - kPolH * kPolH / 8;
    // (b) head rows: [0,A) action_values, [A,2A) the diagonal entries k(k+3)/2 of matrix_entries
    if (u < kPolNH * kPolH / 8) {
        const int n = u >> 5, kc = u & 31;
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const int k = kc * 8 + i;
            float w = 0.f;
            if (n < A) w = p.w_mu[(size_t)n * kPolH + k];
            else if (n < 2 * A) {
                const int d = n - A;
                w = p.w_l[(size_t)((d * (d + 3)) / 2) * kPolH + k];
            }
            v[i] = w;
        }
        const uint32_t off = kPolOffWh + (uint32_t)(kc >> 3) * (kPolNH * 128) + sw128_chunk_offset(n, kc & 7);
        *reinterpret_cast<uint4*>(image + off) = pack8_bf16(v);
        return;
    }
    u -= kPolNH * kPolH / 8;
    // (c) W1 as tf32: one thread per 16-byte chunk (4 consecutive k of output row j), K padded to 32 with zeros
    if (u < kPolH * kPolKP / 4) {
        const int j = u >> 3, kc = u & 7;
        float4 v;
        float* vv = reinterpret_cast<float*>(&v);
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int k = kc * 4 + i;
            vv[i] = k < S ? to_tf32(p.w1[(size_t)j * S + k]) : 0.f;
        }
        *reinterpret_cast<float4*>(image + kPolOffW1 + sw128_chunk_offset(j, kc)) = v;
        return;
    }
    u -= kPolH * kPolKP / 4;
    // (d) eval-mode BatchNorm folded with the linear bias: a = relu(acc * sc + sh)
    float* vec = reinterpret_cast<float*>(image + kPolOffVec);
    if (u < kPolH) {
        const float sc = p.bn1_w[u] / sqrtf(p.bn1_var[u] + 1e-5f);
        vec[u] = sc;
        vec[kPolH + u] = fmaf(p.b1[u] - p.bn1_mean[u], sc, p.bn1_b[u]);
        const float sc2 = p.bn2_w[u] / sqrtf(p.bn2_var[u] + 1e-5f);
        vec[2 * kPolH + u] = sc2;
        vec[3 * kPolH + u] = fmaf(p.b2[u] - p.bn2_mean[u], sc2, p.bn2_b[u]);
        if (u < kPolNH) {
            float b = 0.f;
            if (u < A) b = p.b_mu[u];
            else if (u < 2 * A) b = p.b_l[((u - A) * (u - A + 3)) / 2];
            vec[4 * kPolH + u] = b;
        }
    }
}

// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kPolThreads, 1)
policy_act_tc_kernel(const uint8_t* __restrict__ image, const float* __restrict__ states, int n_rows, int S, int A,
                     unsigned long long seed, unsigned long long step0, const unsigned long long* __restrict__ step_offset,
                     float noise_scale, float* __restrict__ actions) {
    extern __shared__ uint8_t pol_smem_raw[];
    const int tid = threadIdx.x, warp = tid >> 5;
    const int r = tid & 127, half = tid >> 7;            // tile row, which 64 of the 256 hidden columns (0..3)
    const int row = blockIdx.x * 128 + r;

    const uint32_t raw = smem_u32(pol_smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* sm = pol_smem_raw + (base - raw);
    uint8_t* sm_a = sm;                                  // [4 k-blocks][128 rows][128 B]
    uint8_t* sm_w2 = sm + kPolABytes;                    // [4 k-blocks][256 rows][128 B]
    uint8_t* sm_r = sm_w2 + kPolW2Bytes;                 // head image [4][32 rows][128 B]
    float* sm_vec = reinterpret_cast<float*>(sm_r + kPolWhBytes);
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm_r + kPolWhBytes + kPolVecBytes);   // w2, w1 + vectors, wh, mma
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);
    const uint32_t a_base = base, w2_base = base + kPolABytes, r_base = w2_base + kPolW2Bytes;
    // layer-1 operands inside the A tile (dead before epilogue 1 writes the tile): W1 | x_hi | x_lo
    uint8_t *sm_xh = sm_a + kPolW1Bytes, *sm_xl = sm_xh + kPolXBytes;
    const uint32_t xh_base = a_base + kPolW1Bytes, xl_base = xh_base + kPolXBytes;

    if (warp == 0) tmem_alloc(tmem_slot, kPolTmemCols);
    if (tid == 0) {
        mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); mbar_init(&bars[2], 1); mbar_init(&bars[3], 1);
        mbar_init_fence();
        mbar_expect_tx(&bars[1], kPolW1Bytes + kPolVecBytes);
        bulk_g2s(sm_a, image + kPolOffW1, kPolW1Bytes, &bars[1]);
        bulk_g2s(sm_vec, image + kPolOffVec, kPolVecBytes, &bars[1]);
        mbar_expect_tx(&bars[0], kPolW2Bytes);
        bulk_g2s(sm_w2, image, kPolW2Bytes, &bars[0]);
        mbar_expect_tx(&bars[2], kPolWhBytes);
        bulk_g2s(sm_r, image + kPolOffWh, kPolWhBytes, &bars[2]);
    }
    // this thread's 8 observation values (row r, k = 8 half .. 8 half + 7) as tf32 hi / lo pairs, swizzled
    {
        float xh[8], xl[8];
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const int k = half * 8 + i;
            const float x = (k < S && row < n_rows) ? states[(size_t)row * S + k] : 0.f;
            xh[i] = to_tf32(x);
            xl[i] = to_tf32(x - xh[i]);
        }
#pragma unroll
        for (int c = 0; c < 2; c++) {
            const uint32_t off = sw128_chunk_offset(r, half * 2 + c);
            *reinterpret_cast<float4*>(sm_xh + off) = make_float4(xh[4 * c], xh[4 * c + 1], xh[4 * c + 2], xh[4 * c + 3]);
            *reinterpret_cast<float4*>(sm_xl + off) = make_float4(xl[4 * c], xl[4 * c + 1], xl[4 * c + 2], xl[4 * c + 3]);
        }
    }
    fence_proxy_async();                                 // observation tiles -> async proxy
    fence_before_sync();
    __syncthreads();                                     // barrier inits + TMEM address visible
    fence_after_sync();
    const uint32_t tmem = *tmem_slot;

    // relu(acc * scale + shift) of this thread's 64 accumulator columns -> bf16 -> the A tile (TMEM lane = row =
    // 32 (warp % 4) + lane; warp / 4 picks the 64-column quarter)
    auto bn_relu_to_a = [&](const float* sc, const float* sh) {
#pragma unroll 1
        for (int q = 0; q < 2; q++) {
            const int c0 = half * 64 + q * 32;
            uint32_t v[32];
            tmem_ld32(tmem + ((uint32_t)((warp & 3) * 32) << 16) + c0, v);
#pragma unroll
            for (int g = 0; g < 4; g++) {
                float a[8];
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    const int j = c0 + g * 8 + i;
                    a[i] = fmaxf(fmaf(__uint_as_float(v[g * 8 + i]), sc[j], sh[j]), 0.f);
                }
                const int chunk = (c0 >> 3) + g;
                *reinterpret_cast<uint4*>(sm_a + (chunk >> 3) * (128 * 128) + sw128_chunk_offset(r, chunk & 7)) = pack8_bf16(a);
            }
        }
    };

    // ---- layer 1 on the tensor core: D = (x_hi + x_lo) W1^T, tf32 operands, fp32 accumulate ----
    if (tid == 0) {
        mbar_wait(&bars[1], 0);                          // W1 image + vectors landed
        fence_after_sync();
        constexpr uint32_t idesc = idesc_tf32_f32(128, kPolH);
#pragma unroll
        for (int k = 0; k < kPolSP / 8; k++) {
            mma_tf32(tmem, umma_desc_sw128(xh_base + k * 32), umma_desc_sw128(a_base + k * 32), idesc, k > 0);
            mma_tf32(tmem, umma_desc_sw128(xl_base + k * 32), umma_desc_sw128(a_base + k * 32), idesc, true);
        }
        mma_commit(&bars[3]);
    }
    __syncwarp();
    mbar_wait(&bars[1], 0);                              // the vectors, for every thread
    mbar_wait(&bars[3], 0);
    fence_after_sync();
    bn_relu_to_a(sm_vec, sm_vec + kPolH);                // epilogue 1: relu(bn1(.)) over the dead layer-1 operands
    fence_proxy_async();
    fence_before_sync();
    __syncthreads();

    // ---- layer 2 on the tensor core ----
    if (tid == 0) {
        mbar_wait(&bars[0], 0);                          // W2 image landed
        fence_after_sync();
        constexpr uint32_t idesc = idesc_bf16_f32(128, kPolH);
#pragma unroll
        for (int k = 0; k < kPolH / 16; k++)
            mma_bf16(tmem, umma_desc_sw128(a_base + (k >> 2) * (128 * 128) + (k & 3) * 32),
                     umma_desc_sw128(w2_base + (k >> 2) * (kPolH * 128) + (k & 3) * 32), idesc, k > 0);
        mma_commit(&bars[3]);
    }
    __syncwarp();
    mbar_wait(&bars[3], 1);
    fence_after_sync();
    bn_relu_to_a(sm_vec + 2 * kPolH, sm_vec + 3 * kPolH);    // epilogue 2: relu(bn2(.)) -> the A tile again
    fence_proxy_async();
    fence_before_sync();
    __syncthreads();

    // ---- heads on the tensor core: [mu | diag L] = a2 Wh^T, TMEM columns 0-31 ----
    if (tid == 0) {
        mbar_wait(&bars[2], 0);
        fence_after_sync();
        constexpr uint32_t idesc = idesc_bf16_f32(128, kPolNH);
#pragma unroll
        for (int k = 0; k < kPolH / 16; k++)
            mma_bf16(tmem, umma_desc_sw128(a_base + (k >> 2) * (128 * 128) + (k & 3) * 32),
                     umma_desc_sw128(r_base + (k >> 2) * (kPolNH * 128) + (k & 3) * 32), idesc, k > 0);
        mma_commit(&bars[3]);
    }
    __syncwarp();
    if (warp < 4) {
        mbar_wait(&bars[3], 0);                          // third completion of the MMA barrier
        fence_after_sync();
        uint32_t v[32];
        tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16), v);
        if (row < n_rows) {
            const float* bh = sm_vec + 4 * kPolH;
            const unsigned long long stp = step0 + (step_offset != nullptr ? *step_offset : 0ull);
            for (int k = 0; k < A; k++) {
                float zmu = 0.f, zl = 0.f;
#pragma unroll
                for (int i = 0; i < 32; i++) {           // register array: select by compile-time index
                    if (i == k) zmu = __uint_as_float(v[i]);
                    if (i == A + k) zl = __uint_as_float(v[i]);
                }
                const float mu = tanhf(zmu + bh[k]);
                const float t = tanhf(zl + bh[A + k]);
                uint32_t c[4] = {(uint32_t)row, (uint32_t)k, (uint32_t)stp, (uint32_t)(stp >> 32)};
                philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
                const float rr = sqrtf(-2.f * logf(u01(c[0])));
                const float eps = rr * cospif(2.f * u01(c[1]));
                const float a = fmaf(noise_scale * expf(-t), eps, mu);
                actions[(size_t)row * A + k] = fminf(fmaxf(a, -1.f), 1.f);
            }
        }
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, kPolTmemCols);
}

// ---------------------------------------------------------------------------------------------------
bool policy_tc_supported(int S, int A, int H) { return H == kPolH && S >= 1 && S <= kPolSP && A >= 1 && 2 * A <= kPolNH; }

int policy_tc_prepare(TrunkTC* t) {
    if (t->policy_image != nullptr) return RLOA_OK;
    RLOA_CUDA(cudaFuncSetAttribute(policy_act_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPolSmemBytes));
    RLOA_CUDA(cudaMalloc(&t->policy_image, kPolImageBytes));
    return RLOA_OK;
}

int policy_tc_act(TrunkTC* t, const rloa_naf_params* p, const float* states, int batch, uint64_t seed, uint64_t step,
                  const uint64_t* step_offset, float noise_scale, float* actions, cudaStream_t st) {
    RLOA_REQUIRE(t->policy_image != nullptr, "tcgen05 policy: rloa_naf_ws_set_trunk(1) was not called");
    uint8_t* image = static_cast<uint8_t*>(t->policy_image);
    const int pack_threads = kPolH * kPolH / 8 + kPolNH * kPolH / 8 + kPolH * kPolKP / 4 + kPolH;
    policy_pack_kernel<<<(pack_threads + 255) / 256, 256, 0, st>>>(*p, image);
    RLOA_LAUNCHED();
    policy_act_tc_kernel<<<(batch + 127) / 128, kPolThreads, kPolSmemBytes, st>>>(
        image, states, batch, p->state_size, p->action_size, seed, step,
        reinterpret_cast<const unsigned long long*>(step_offset), noise_scale, actions);
    RLOA_LAUNCHED();
    return RLOA_OK;
}

}  // namespace rloa
