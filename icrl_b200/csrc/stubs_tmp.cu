// TEMPORARY: entry points not implemented yet fail loudly.
#include "common.cuh"
extern "C" {
int icrl_cn_train(const icrl_cn_desc*, const icrl_cn_train_cfg*, const void*, int32_t, const float*, int64_t, const int32_t*, int32_t, const void*, int32_t, const float*, int64_t, float*, float*, int64_t*, icrl_cn_train_metrics*, void*) { icrl::set_error("icrl_cn_train not implemented"); return ICRL_EUNSUPPORTED; }
}
