// tcgen05 tensor-core trunk — placeholder until the UMMA kernel lands
#include "common.cuh"
#include "naf_trunk_tc.cuh"
namespace rloa {
void trunk_tc_init(TrunkTC* t) { *t = TrunkTC{}; }
void trunk_tc_free(TrunkTC* t) {
    if (t->a_bf16) cudaFree(t->a_bf16);
    if (t->w_bf16) cudaFree(t->w_bf16);
    *t = TrunkTC{};
}
int trunk_tc_prepare(TrunkTC*, int, int) { return fail(RLOA_ERR_INVALID, "tcgen05 trunk: not built yet"); }
int trunk_tc_layer2(TrunkTC*, const float*, const float*, const float*, const float*, const float*, float*, int, int,
                    cudaStream_t) { return fail(RLOA_ERR_INVALID, "tcgen05 trunk: not built yet"); }
}  // namespace rloa
