// tcgen05 tensor-core trunk (hidden_layer 256x256 contraction) — declared here, defined in naf_trunk_tc.cu
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace rloa {

struct TrunkTC {
    void* a_bf16 = nullptr;     // [max_batch][256] activations, bf16
    void* w_bf16 = nullptr;     // [256][256] weights, bf16
    int max_batch = 0;
    bool ready = false;
};

void trunk_tc_init(TrunkTC* t);
void trunk_tc_free(TrunkTC* t);
int trunk_tc_prepare(TrunkTC* t, int max_batch, int H);
// z2[B][H] = relu(z1 * scale + shift) @ w2^T + b2 on the tensor cores
int trunk_tc_layer2(TrunkTC* t, const float* z1, const float* scale, const float* shift, const float* w2,
                    const float* b2, float* z2, int B, int H, cudaStream_t st);

}  // namespace rloa
