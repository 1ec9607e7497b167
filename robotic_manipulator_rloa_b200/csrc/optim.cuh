// Optimiser pieces shared by naf.cu (clip + Adam + soft update) and grad_exchange.cu (the same fused with the NVLink
// gradient exchange): the parameter table, the clip / bias-correction coefficients, and the per-element update of
// reference naf_components/naf_algorithm.py:209-226 (clip_grad_norm_, optim.Adam.step, soft_update).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "rloa_b200.h"

namespace rloa {

constexpr int kNormBlocks = 80;    // blocks (= partial sums) of the gradient-norm pass; 80 x 256 threads cover 79,644 / 4 float4

struct ParamTable {                // the 14 parameter tensors in nn.Module.parameters() order
    float* main[14];
    float* target[14];
    int offset[15];
};

struct AdamCoef {
    float clip;                    // min(clip_norm / (norm + 1e-6), 1) * grad_scale
    float bc1, bc2s;               // 1 - beta1^t, sqrt(1 - beta2^t)
    float norm;
};

// sq_partial: kNormBlocks partial sums of (grad * grad_scale)^2 in a fixed order; step = optimizer.step() count (>= 1)
__device__ __forceinline__ AdamCoef adam_coefficients(const float* sq_partial, long long step, const rloa_naf_hyper& hp) {
    float t = 0.f;
    for (int i = 0; i < kNormBlocks; i++) t += sq_partial[i];
    AdamCoef c;
    c.norm = sqrtf(t);
    // clip_grad_norm_: coef = max_norm / (norm + 1e-6), clamped to 1
    c.clip = fminf(hp.clip_norm / (c.norm + 1e-6f), 1.f) * hp.grad_scale;
    c.bc1 = 1.f - powf(hp.beta1, (float)step);
    c.bc2s = sqrtf(1.f - powf(hp.beta2, (float)step));
    return c;
}

// element i of the flat gradient: Adam moments, main parameter, then the soft target update with the fresh value
__device__ __forceinline__ void adam_soft_update_element(const ParamTable& pt, int i, float grad_i, const AdamCoef& c,
                                                         float* __restrict__ m, float* __restrict__ v,
                                                         const rloa_naf_hyper& hp) {
    int t = 0;
#pragma unroll
    for (int k = 1; k < 14; k++) t += (i >= pt.offset[k]) ? 1 : 0;
    const int j = i - pt.offset[t];
    const float g = grad_i * c.clip;
    const float mi = fmaf(1.f - hp.beta1, g - m[i], m[i]);          // m = b1 m + (1-b1) g
    const float vi = fmaf(1.f - hp.beta2, g * g - v[i], v[i]);      // v = b2 v + (1-b2) g^2
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) / c.bc2s + hp.eps;
    const float p = pt.main[t][j] - (hp.lr / c.bc1) * (mi / denom);
    pt.main[t][j] = p;
    // soft update with the freshly updated main parameter (naf_algorithm.py:213, 225-226)
    pt.target[t][j] = hp.tau * p + (1.f - hp.tau) * pt.target[t][j];
}

// sums the split-K partials of up to 3 weight gradients in a fixed order (deterministic)
struct ReduceSeg {
    const float* part; float* out; int n; int nsplit; size_t pstride;
    int off;                    // offset of `out` in the flat gradient
};
struct ReduceArgs {
    ReduceSeg s[5];
};

// element e of the flat gradient: the fixed-order sum of its split-K partials when e lies in a deferred segment,
// `current` otherwise
__device__ __forceinline__ float splitk_element(const ReduceArgs& r, int e, float current) {
#pragma unroll
    for (int q = 0; q < 5; q++) {
        const ReduceSeg& sg = r.s[q];
        const int j = e - sg.off;
        if (j >= 0 && j < sg.n) {
            float t = 0.f;
            int k = 0;
            for (; k + 16 <= sg.nsplit; k += 16) {        // 16 partials in flight, added in split order
                float x[16];
#pragma unroll
                for (int u = 0; u < 16; u++) x[u] = sg.part[(k + u) * sg.pstride + j];
#pragma unroll
                for (int u = 0; u < 16; u++) t += x[u];
            }
            for (; k < sg.nsplit; k++) t += sg.part[k * sg.pstride + j];
            return t;
        }
    }
    return current;
}


}  // namespace rloa
