// placeholder: replay entry points (filled in next)
#include "common.cuh"
using namespace rloa;
extern "C" int rloa_replay_append(const rloa_replay*, int32_t, const float*, const float*, const float*, const float*, const uint8_t*, const uint8_t*, void*) { return fail(RLOA_ERR_INVALID, "rloa_replay_append: not implemented yet"); }
extern "C" int rloa_replay_sample(const rloa_replay*, int32_t, uint64_t, uint64_t, float*, float*, float*, float*, float*, int32_t*, void*) { return fail(RLOA_ERR_INVALID, "rloa_replay_sample: not implemented yet"); }
