// placeholder: NAF entry points (filled in next)
#include "common.cuh"
using namespace rloa;
struct rloa_naf_ws { int dummy; };
#define NOTIMPL(name) return fail(RLOA_ERR_INVALID, name ": not implemented yet")
extern "C" int rloa_naf_ws_create(int32_t, int32_t, int32_t, int32_t, rloa_naf_ws**) { NOTIMPL("rloa_naf_ws_create"); }
extern "C" void rloa_naf_ws_destroy(rloa_naf_ws*) {}
extern "C" int rloa_naf_ws_set_trunk(rloa_naf_ws*, int32_t) { NOTIMPL("rloa_naf_ws_set_trunk"); }
extern "C" int rloa_naf_forward(rloa_naf_ws*, const rloa_naf_params*, const float*, const float*, int32_t, int32_t, int32_t, float*, float*, float*, float*, void*) { NOTIMPL("rloa_naf_forward"); }
extern "C" int rloa_naf_act(rloa_naf_ws*, const rloa_naf_params*, const float*, int32_t, uint64_t, uint64_t, float, float*, void*) { NOTIMPL("rloa_naf_act"); }
extern "C" int rloa_naf_num_params(int32_t, int32_t, int32_t) { NOTIMPL("rloa_naf_num_params"); }
extern "C" int rloa_naf_learn_grads(rloa_naf_ws*, const rloa_naf_params*, const rloa_naf_params*, const float*, const float*, const float*, const float*, const float*, int32_t, const rloa_naf_hyper*, float*, float*, void*) { NOTIMPL("rloa_naf_learn_grads"); }
extern "C" int rloa_naf_learn_apply(rloa_naf_ws*, const rloa_naf_params*, const rloa_naf_params*, const rloa_adam_state*, const rloa_naf_hyper*, float*, float*, void*) { NOTIMPL("rloa_naf_learn_apply"); }
extern "C" int rloa_naf_soft_update(const rloa_naf_params*, const rloa_naf_params*, float, void*) { NOTIMPL("rloa_naf_soft_update"); }
