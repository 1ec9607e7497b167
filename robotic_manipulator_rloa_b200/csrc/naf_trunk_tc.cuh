// tcgen05 tensor-core trunk (hidden_layer 256x256 contraction) — declared here, defined in naf_trunk_tc.cu
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "bn_fuse.cuh"
#include "rloa_b200.h"

namespace rloa {

struct TrunkTC {
    int max_batch = 0;
    bool ready = false;
    void* policy_image = nullptr;   // device: bf16 weight image of the fused policy kernel (naf_policy_tc.cu)
};

void trunk_tc_init(TrunkTC* t);
void trunk_tc_free(TrunkTC* t);
// checks the device, opts the kernel into its dynamic shared memory size
int trunk_tc_prepare(TrunkTC* t, int max_batch, int H);
// for each of `nets` networks: z2[B][H] = bf16(relu(z1 * scale + shift)) @ bf16(w2)^T + b2, fp32 accumulation in TMEM;
// bn (one BnFuse per net, or NULL): train-mode BatchNorm statistics of z2 fused into the epilogue (128-row chunks)
int trunk_tc_layer2(TrunkTC* t, int nets, const float* const* z1, const float* const* scale, const float* const* shift,
                    const float* const* w2, const float* const* b2, float* const* z2, int B, int H, const BnFuse* bn,
                    cudaStream_t st);

// layer 1 on the tensor core (kind::tf32, observations split hi + lo): z1[B][H] = x[B][S] @ tf32(w1[H][S])^T + b1, with the
// train-mode BatchNorm statistics of z1 fused like in trunk_tc_layer2
bool trunk_tc_layer1_supported(int S, int H);
int trunk_tc_layer1(TrunkTC* t, int nets, const float* const* x, const float* const* w1, const float* const* b1,
                    float* const* z1, int B, int S, int H, const BnFuse* bn, cudaStream_t st);

// input gradient of the hidden layer on the tensor core: da[B][H] = bf16(dz[B][H]) @ bf16(w[H][H]) (w = [out][in])
int trunk_tc_input_grad(TrunkTC* t, const float* dz, const float* w, float* da, int B, int H, cudaStream_t st);

// fused eval-mode policy (naf_policy_tc.cu): supported shapes, one-time setup, and NAFAgent.act for `batch` rows
bool policy_tc_supported(int S, int A, int H);
int policy_tc_prepare(TrunkTC* t);
int policy_tc_act(TrunkTC* t, const rloa_naf_params* p, const float* states, int batch, uint64_t seed, uint64_t step,
                  const uint64_t* step_offset, float noise_scale, float* actions, cudaStream_t st);

}  // namespace rloa
