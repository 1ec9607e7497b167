// Fused NAFAgent.learn on two thread-block clusters (naf_learn_cluster.cu) — declarations for naf.cu
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "optim.cuh"
#include "rloa_b200.h"

struct rloa_xchg;

namespace rloa {

struct LearnCluster {
    void* images = nullptr;          // bf16 / tf32 weight images of both networks (shared-memory byte order)
    float* block = nullptr;          // y | partial weight gradients | head-bias / loss partials | hand-over flags
    float *y = nullptr, *part_w2 = nullptr, *part_w1 = nullptr, *part_wh = nullptr, *part_hb = nullptr, *part_loss = nullptr;
    unsigned* yflag = nullptr;
    unsigned* aflag = nullptr;       // optimiser hand-over main CTA -> target CTA (flags, coefficients)
    float* acoef = nullptr;
    float* dbg = nullptr;            // optional intermediate dump (tests)
    long long* prof = nullptr;       // optional clock64 phase stamps
    bool prepacked = false;          // rloa_naf_learn_prepack wrote the images on a side stream; the next learn joins pack_done
    cudaEvent_t pack_done = nullptr, pack_fork = nullptr;
};

struct LearnClusterReplay {          // fused ReplayBuffer.sample: the kernel reads its rows straight from the ring
    const rloa_replay* rb;
    uint64_t seed, draw;
    const uint64_t* draw_offset;
    // optional pending rows (appended to the ring concurrently, not committed yet): the kernel samples the ring as it will be
    const float *pd_states = nullptr, *pd_actions = nullptr, *pd_rewards = nullptr, *pd_next_states = nullptr;
    const uint8_t *pd_dones = nullptr, *pd_valid = nullptr;
    int pd_n = 0;
};
constexpr int kLearnClusterMaxPending = 16384;      // 32 rows per thread of a CTA

bool learn_cluster_supported(int S, int A, int H, int B);
constexpr int kLearnClusterMaxWorld = 8;     // the in-kernel gradient exchange; larger worlds use the separate exchange kernels
int learn_cluster_prepare(LearnCluster* lc, int S, int A);
void learn_cluster_free(LearnCluster* lc);
// flat_offsets: the 14 segment offsets of the flat gradient in nn.Module.parameters() order + the total
int learn_cluster_step(LearnCluster* lc, const rloa_naf_params* mn, const rloa_naf_params* tg, const rloa_adam_state* adam,
                       const float* states, const float* actions, const float* rewards, const float* next_states,
                       const float* dones, int B, const rloa_naf_hyper* hp, const ParamTable& pt, const int* flat_offsets,
                       float* grad, float* loss, float* gnorm, int do_adam, const LearnClusterReplay* replay, const rloa_xchg* xchg,
                       cudaStream_t st);
// the weight images of both networks (what learn_cluster_step does first unless a prepack is pending)
int learn_cluster_pack(LearnCluster* lc, const rloa_naf_params* mn, const rloa_naf_params* tg, cudaStream_t st);

}  // namespace rloa
