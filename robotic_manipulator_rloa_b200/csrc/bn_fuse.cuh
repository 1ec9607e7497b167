// Train-mode BatchNorm1d statistics fused into the kernel that PRODUCES the pre-activations (naf.cu SGEMMs,
// naf_trunk_tc.cu): every output tile leaves a per-column (mean, M2) partial over its rows, and the last tile of a
// column group to finish merges the partials in a fixed order with Chan's update (deterministic; as accurate as a
// two-pass mean / variance) and finalises exactly like nn.BatchNorm1d in training mode
// (reference naf_components/naf_neural_network.py:76-78): running statistics with momentum 0.1 and the unbiased
// variance, num_batches_tracked += 1, and the folded per-column scale / shift the consumers apply.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace rloa {

constexpr float kBnEps = 1e-5f;        // nn.BatchNorm1d defaults
constexpr float kBnMomentum = 0.1f;

struct BnArgs {
    const float* z;             // [B][H]
    const float *w, *b;
    float *run_mean, *run_var;
    int64_t* batches;
    float *scale, *shift, *mean, *rstd;
};

struct BnFuse {
    BnArgs bn;
    float* part;                // [chunks][2][H]: per-chunk column mean | M2
    unsigned* ticket;           // one counter per column tile, zero at rest
    int enabled;
};

// column f from `chunks` partials of `chunk_rows` rows each (the last one shorter)
__device__ __forceinline__ void bn_finalize_column(const BnArgs& a, const float* part, int chunks, int chunk_rows, int B,
                                                   int H, int f, bool bump_batches) {
    // all partials are requested before the (serial) merge starts: one L2 round trip instead of one per chunk
    constexpr int kPrefetch = 32;
    float pm[kPrefetch], pM[kPrefetch];
#pragma unroll
    for (int k = 0; k < kPrefetch; k++) {
        pm[k] = k < chunks ? __ldcg(part + (size_t)k * 2 * H + f) : 0.f;
        pM[k] = k < chunks ? __ldcg(part + (size_t)k * 2 * H + H + f) : 0.f;
    }
    float n = 0.f, mean = 0.f, M2 = 0.f;
    auto merge = [&](int k, float mb, float Mb) {
        const float nb = (float)min(chunk_rows, B - k * chunk_rows);
        const float nn = n + nb, delta = mb - mean;
        mean = fmaf(delta, nb / nn, mean);
        M2 = M2 + Mb + delta * delta * (n * nb / nn);
        n = nn;
    };
#pragma unroll
    for (int k = 0; k < kPrefetch; k++)
        if (k < chunks) merge(k, pm[k], pM[k]);
    for (int k = kPrefetch; k < chunks; k++)
        merge(k, __ldcg(part + (size_t)k * 2 * H + f), __ldcg(part + (size_t)k * 2 * H + H + f));
    const float var = M2 / (float)B;
    const float unbiased = B > 1 ? M2 / (float)(B - 1) : var;
    a.run_mean[f] = fmaf(kBnMomentum, mean - a.run_mean[f], a.run_mean[f]);
    a.run_var[f] = fmaf(kBnMomentum, unbiased - a.run_var[f], a.run_var[f]);
    if (bump_batches && a.batches != nullptr) *a.batches += 1;
    const float rstd = 1.f / sqrtf(var + kBnEps);
    const float sc = a.w[f] * rstd;
    a.scale[f] = sc;
    a.shift[f] = fmaf(-mean, sc, a.b[f]);
    a.mean[f] = mean;
    a.rstd[f] = rstd;
}

}  // namespace rloa
