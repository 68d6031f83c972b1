// TEMPORARY: entry points not implemented yet fail loudly.
#include "common.cuh"
extern "C" {
int icrl_cn_train(const icrl_cn_desc*, const icrl_cn_train_cfg*, const void*, int32_t, const float*, int64_t, const int32_t*, int32_t, const void*, int32_t, const float*, int64_t, float*, float*, int64_t*, icrl_cn_train_metrics*, void*) { icrl::set_error("icrl_cn_train not implemented"); return ICRL_EUNSUPPORTED; }
int64_t icrl_ppo_param_count(const icrl_ppo_cfg*) { return -1; }
int icrl_ppo_train(const icrl_ppo_cfg*, const icrl_ppo_data*, float*, float*, float*, int64_t, float*, int32_t*, void*) { icrl::set_error("icrl_ppo_train not implemented"); return ICRL_EUNSUPPORTED; }
int icrl_policy_forward(const icrl_ppo_cfg*, const float*, const float*, int64_t, float*, float*, float*, void*) { icrl::set_error("not implemented"); return ICRL_EUNSUPPORTED; }
int icrl_dual_update(float*, const float*, int64_t, double, double, int64_t, double, void*) { icrl::set_error("not implemented"); return ICRL_EUNSUPPORTED; }
}
