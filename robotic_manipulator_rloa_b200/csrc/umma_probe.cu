// Self-test of the tcgen05 operand views the fused learn kernel (naf_learn_cluster.cu) relies on: one bf16 tile stored
// ONCE in the canonical K-major SWIZZLE_128B layout ([row][col], 64 columns = one 128-byte swizzle row per k-block) is
// consumed both as a K-major operand (contraction over its columns) and — without being transposed or re-staged — as an
// MN-major operand (contraction over its ROWS), by flipping the major bit of the instruction descriptor and describing
// the same bytes with LBO = the k-block stride and SBO = 1024 (the 8-row atom).  The backward of a linear layer needs
// exactly that: da = dz W contracts over W's rows, dW = dz^T a contracts over the batch rows of both tiles
// (reference naf_algorithm.py:208, autograd of naf_neural_network.py:76-87).
//
//   mode 0:  C[128][256] = A[128][256] . B[256][256]^T        A, B K-major                      (forward shape)
//   mode 1:  C[128][256] = A[128][256] . B[256][256]          B read as MN-major                (input gradient)
//   mode 2:  C[256][256] = A[128][256]^T . B[128][256]        A and B read as MN-major          (weight gradient)
//   mode 3:  C[256][64]  = A[128][256]^T . B[128][64]         as mode 2 with a 64-wide B tile   (dW1 / head gradients)
// Inputs fp32 row-major, rounded to bf16 while staged; output fp32 row-major.  sm_100a only; exposed as rloa_umma_probe
// so tests/test_umma_probe_gpu.py can pin the layouts against a plain matrix product.
#include "common.cuh"
#include "tc_common.cuh"

namespace rloa {

constexpr int kProbeThreads = 256;

// fp32 [rows][cols] (leading dimension ld) -> bf16 K-major SWIZZLE_128B k-blocks of 64 columns; block stride rows * 128 B
__device__ void probe_stage(uint8_t* sm, const float* __restrict__ src, int rows, int cols, int ld) {
    const int chunks_per_row = cols / 8;
    for (int i = threadIdx.x; i < rows * chunks_per_row; i += blockDim.x) {
        const int r = i / chunks_per_row, ch = i - r * chunks_per_row;
        const float4 a = *reinterpret_cast<const float4*>(src + (size_t)r * ld + ch * 8);
        const float4 b = *reinterpret_cast<const float4*>(src + (size_t)r * ld + ch * 8 + 4);
        const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
        const uint32_t off = (uint32_t)(ch >> 3) * (uint32_t)rows * 128u + tc::sw128_chunk_offset(r, ch & 7);
        *reinterpret_cast<uint4*>(sm + off) = tc::pack8_bf16(v);
    }
}

__global__ void __launch_bounds__(kProbeThreads) umma_probe_kernel(int mode, const float* __restrict__ A,
                                                                   const float* __restrict__ B, float* __restrict__ C) {
    extern __shared__ uint8_t probe_smem_raw[];
    const uint32_t raw = tc::smem_u32(probe_smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* sm = probe_smem_raw + (base - raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // A tile: 128 rows x 256 cols (4 k-blocks of 16 KB); B: 256 x 256 (4 x 32 KB), 128 x 256 (4 x 16 KB) or 128 x 64 (16 KB)
    uint8_t* sm_a = sm;
    uint8_t* sm_b = sm + 65536;
    uint64_t* bar = reinterpret_cast<uint64_t*>(sm + 65536 + 131072);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
    if (warp == 0) tc::tmem_alloc(tmem_slot, 512);
    if (tid == 0) { tc::mbar_init(bar, 1); tc::mbar_init_fence(); }
    const int b_rows = mode <= 1 ? 256 : 128, b_cols = mode == 3 ? 64 : 256;
    probe_stage(sm_a, A, 128, 256, 256);
    probe_stage(sm_b, B, b_rows, b_cols, b_cols);
    tc::fence_proxy_async();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = *tmem_slot;
    const uint32_t a_base = base, b_base = base + 65536;
    const uint32_t a_blk = 128 * 128, b_blk = (uint32_t)b_rows * 128;
    if (tid == 0) {
        if (mode == 0) {
            const uint32_t idesc = tc::idesc_bf16_f32(128, 256);
            for (int k = 0; k < 16; k++)
                tc::mma_bf16(tmem, tc::umma_desc_sw128(a_base + (k >> 2) * a_blk + (k & 3) * 32),
                             tc::umma_desc_sw128(b_base + (k >> 2) * b_blk + (k & 3) * 32), idesc, k > 0);
        } else if (mode == 1) {
            // contraction over B's 256 rows: 16 steps of 16 rows = 2 atoms of 1024 B; the 256 output columns are 4 MN atoms
            // one k-block (LBO = b_blk) apart
            const uint32_t idesc = tc::idesc_bf16_f32(128, 256) | tc::kIdescBMajorMN;
            for (int k = 0; k < 16; k++)
                tc::mma_bf16(tmem, tc::umma_desc_sw128(a_base + (k >> 2) * a_blk + (k & 3) * 32),
                             tc::umma_desc_sw128_mn(b_base + k * 2048, b_blk), idesc, k > 0);
        } else {
            // contraction over the 128 batch rows of both tiles: 8 steps of 16 rows; output rows = A's 256 columns in two
            // M = 128 halves (2 MN atoms each, the second half starts 2 LBO further)
            const int N = mode == 2 ? 256 : 64;
            const uint32_t idesc = tc::idesc_bf16_f32(128, N) | tc::kIdescAMajorMN | tc::kIdescBMajorMN;
            for (int half = 0; half < 2; half++)
                for (int k = 0; k < 8; k++)
                    tc::mma_bf16(tmem + half * N, tc::umma_desc_sw128_mn(a_base + half * 2 * a_blk + k * 2048, a_blk),
                                 tc::umma_desc_sw128_mn(b_base + k * 2048, b_blk), idesc, k > 0);
        }
        tc::mma_commit(bar);
    }
    __syncwarp();
    tc::mbar_wait(bar, 0);
    tc::fence_after_sync();
    // epilogue: warp w reads TMEM lanes 32 (w % 4) .. + 31; the column range is split between warps w and w + 4
    const int n_out = mode == 3 ? 64 : 256, halves = mode >= 2 ? 2 : 1;
    for (int half = 0; half < halves; half++) {
        const int row = half * 128 + (warp & 3) * 32 + lane;
        for (int c0 = (warp >> 2) * 32; c0 < n_out; c0 += 64) {
            uint32_t v[32];
            tc::tmem_ld32(tmem + ((uint32_t)((warp & 3) * 32) << 16) + half * n_out + c0, v);
#pragma unroll
            for (int j = 0; j < 32; j += 4)
                *reinterpret_cast<float4*>(C + (size_t)row * n_out + c0 + j) =
                    make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, 512);
}

}  // namespace rloa

using namespace rloa;

extern "C" int rloa_umma_probe(int32_t mode, const float* a, const float* b, float* c, void* stream) {
    RLOA_REQUIRE(mode >= 0 && mode <= 3 && a && b && c, "rloa_umma_probe: bad argument");
    int dev = 0, major = 0;
    RLOA_CUDA(cudaGetDevice(&dev));
    RLOA_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
    RLOA_REQUIRE(major == 10, "rloa_umma_probe: needs an sm_100 device");
    const int smem = 65536 + 131072 + 1024 + 64;
    RLOA_CUDA(cudaFuncSetAttribute(umma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    umma_probe_kernel<<<1, kProbeThreads, smem, as_stream(stream)>>>(mode, a, b, c);
    RLOA_LAUNCHED();
    return RLOA_OK;
}
