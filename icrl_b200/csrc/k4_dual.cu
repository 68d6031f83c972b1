// K4 (dual step) -- Lagrange multiplier update, replaces DualVariable.update_parameter + Nu.clamp
// (stable_baselines3/common/dual_variable.py:9-57) and the np.mean(orig_costs) that feeds it (ppo_lag.py:303-306).
// One CTA: float64 tree reduction of the rollout's original costs, then a scalar softplus / Adam(eps 1e-8) / clamp
// step on log_nu.  state = {log_nu, exp_avg, exp_avg_sq, last_loss, nu_after, mean_cost}.
#include "common.cuh"

namespace icrl {

__global__ void __launch_bounds__(1024) dual_update_kernel(float* __restrict__ state, const float* __restrict__ costs,
                                                           long long n, float alpha, double lr, long long step_before,
                                                           float clamp_min) {
    __shared__ double red[32];
    double s = 0.0;
    for (long long i = threadIdx.x; i < n; i += blockDim.x) s += (double)costs[i];
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double tot = 0.0;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) tot += red[i];
        const float mean_cost = (float)(tot / (double)n);          // np.mean(float32 array) -> float32
        const float diff = mean_cost - alpha;                      // float32 - python float -> float32
        float log_nu = state[0], m = state[1], v = state[2];
        const float ex = expf(log_nu);
        const float nu = log_nu > 20.f ? log_nu : log1pf(ex);      // F.softplus(beta=1, threshold=20)
        const float loss = -nu * diff;
        const float dsp = log_nu > 20.f ? 1.f : ex / (ex + 1.f);   // softplus backward: z / (z + 1), z = exp(x)
        const float g = -diff * dsp;
        const double t = (double)(step_before + 1);
        const double bc1 = 1.0 - pow(0.9, t), bc2 = 1.0 - pow(0.999, t);
        m = fmaf(0.1f, g - m, m);
        v = fmaf(0.001f * g, g, v * 0.999f);
        const float denom = sqrtf(v) / (float)sqrt(bc2) + 1e-8f;
        log_nu = fmaf((float)(-(lr / bc1)), m / denom, log_nu);
        log_nu = fmaxf(log_nu, clamp_min);                         // Nu.clamp (dual_variable.py:27-29)
        state[0] = log_nu; state[1] = m; state[2] = v; state[3] = loss;
        const float ex2 = expf(log_nu);
        state[4] = log_nu > 20.f ? log_nu : log1pf(ex2);
        state[5] = mean_cost;
    }
}

}  // namespace icrl

extern "C" int icrl_dual_update(float* state, const float* orig_costs, int64_t n, double alpha, double lr,
                                int64_t adam_step_before, double clamp_min_log_nu, void* stream) {
    ICRL_CHECK_ARG(state && orig_costs && n > 0, "icrl_dual_update: NULL pointer or empty cost array");
    icrl::dual_update_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(state, orig_costs, (long long)n, (float)alpha, lr,
                                                                   (long long)adam_step_before, (float)clamp_min_log_nu);
    ICRL_LAUNCH_CHECK();
    return 0;
}
