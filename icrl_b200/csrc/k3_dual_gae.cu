// K3 -- dual reward/cost GAE as a chunked, segmented reverse scan.
// Replaces RolloutBufferWithCost._compute_returns_and_advantage x2
// (stable_baselines3/common/buffers.py:493-552), a Python loop over n_steps.
//
// A_t = delta_t + c_t * A_{t+1} is an affine recurrence; `done` flags make it segmented (c_t = 0).
// A CTA owns a group of adjacent env columns (32: every [T,E] access is a coalesced 128-byte row segment; 8: four
// time sub-chunks share a warp) and its threads split the time axis into chunks:
//   phase 1  each thread folds its chunk into an affine map (M, B):  A_lo = B + M * A_hi      (float64)
//   phase 2  chunk maps are combined back-to-front through shared memory -> carry-in per chunk
//   phase 3  the chunk is replayed with the reference's exact operation order and dtypes:
//            delta in float32 (rounded per op, no FMA contraction), carry in float64, one rounding to
//            float32 per stored advantage, returns = adv_f32 + value_f32.
// Reward and cost scans run in the same thread (two independent dependency chains).
// Roofline: 5 float reads + 4 float writes = 36 B per transition; HBM-bound.  Phase 3 re-reads the chunk,
// which is served by L1/L2 for reference-sized buffers.
#include "common.cuh"

namespace icrl {

struct GaeArgs {
    const float *r, *vr, *c, *vc, *dones, *last_vr, *last_vc;
    const uint8_t* last_dones;
    float *adv_r, *ret_r, *adv_c, *ret_c;
    int T, E;
    float g_r, gl_r, g_c, gl_c;   // float32(gamma), float32(gamma*lambda) -- numpy's weak-scalar casts
};

struct Step {
    double delta;  // exact widening of the float32 delta (float64-computed for t = T-1)
    double coef;   // c_t
};

// delta_t and c_t for one signal at time t, exactly as numpy evaluates buffers.py:528-537.
__device__ __forceinline__ Step gae_step(const float* __restrict__ rew, const float* __restrict__ val, float next_val,
                                         float alive_f32, bool is_last, double alive_last, float g32, float gl32,
                                         int64_t idx) {
    Step s;
    const float r = rew[idx], v = val[idx];
    if (!is_last) {
        const float t1 = __fmul_rn(g32, next_val);
        const float t2 = __fmul_rn(t1, alive_f32);
        const float t3 = __fadd_rn(r, t2);
        s.delta = (double)__fsub_rn(t3, v);
        s.coef = (double)__fmul_rn(gl32, alive_f32);
    } else {
        // next_non_terminal comes from a bool array -> float64 from here on
        const double t2 = __dmul_rn((double)__fmul_rn(g32, next_val), alive_last);
        s.delta = __dsub_rn(__dadd_rn((double)r, t2), (double)v);
        s.coef = 0.0;   // multiplies the initial carry 0
    }
    return s;
}

constexpr int UB = 8;   // time steps whose loads are batched

// both signals' (delta, coef) at time t for column col
__device__ __forceinline__ void load_steps(const GaeArgs& a, int t, int col, float lvr, float lvc, double alive_last,
                                           Step& sr, Step& sc) {
    const int64_t idx = (int64_t)t * a.E + col;
    const bool last = (t == a.T - 1);
    const float alive = last ? 0.f : __fsub_rn(1.0f, a.dones[idx + a.E]);
    const float nvr = last ? lvr : a.vr[idx + a.E];
    const float nvc = last ? lvc : a.vc[idx + a.E];
    sr = gae_step(a.r, a.vr, nvr, alive, last, alive_last, a.g_r, a.gl_r, idx);
    sc = gae_step(a.c, a.vc, nvc, alive, last, alive_last, a.g_c, a.gl_c, idx);
}

// CG = env columns per CTA (32: a warp row is one 128-byte segment; 8: four time sub-chunks share a warp, each lane group
// reading a 32-byte sector -- 4x more chunks and 4x more CTAs for narrow / mid-sized buffers).  NW warps; the time axis is
// split into NCH = NW * (32 / CG) chunks.
template <int NW, int CG>
__global__ void __launch_bounds__(NW * 32) dual_gae_kernel(const GaeArgs a) {
    constexpr int SUB = 32 / CG, NCH = NW * SUB;
    __shared__ double sM[2][NCH][CG], sB[2][NCH][CG];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int cl = lane % CG;                       // column within the CTA's group
    const int ch = w * SUB + lane / CG;             // this thread's time chunk
    const int col = blockIdx.x * CG + cl;
    const bool active = col < a.E;
    const int T = a.T, E = a.E;
    const int Lc = (T + NCH - 1) / NCH;
    const int lo = min(T, ch * Lc), hi = min(T, lo + Lc);

    double alive_last = 0.0;
    float lvr = 0.f, lvc = 0.f;
    if (active) {
        alive_last = a.last_dones[col] ? 0.0 : 1.0;
        lvr = a.last_vr[col];
        lvc = a.last_vc[col];
    }

    // ---- phase 1: fold the chunk.  Loads of UB consecutive steps are issued together (the recurrence itself is
    // serial, the memory traffic must not be), then folded in order.
    double Mr = 1.0, Br = 0.0, Mc = 1.0, Bc = 0.0;
    if (active) {
        for (int t1 = hi - 1; t1 >= lo; t1 -= UB) {
            Step sr[UB], sc[UB];
#pragma unroll
            for (int u = 0; u < UB; ++u) {
                const int t = t1 - u;
                if (t >= lo) load_steps(a, t, col, lvr, lvc, alive_last, sr[u], sc[u]);
            }
#pragma unroll
            for (int u = 0; u < UB; ++u) {
                if (t1 - u >= lo) {
                    Br = sr[u].delta + sr[u].coef * Br;
                    Mr = sr[u].coef * Mr;
                    Bc = sc[u].delta + sc[u].coef * Bc;
                    Mc = sc[u].coef * Mc;
                }
            }
        }
    }
    sM[0][ch][cl] = Mr; sB[0][ch][cl] = Br;
    sM[1][ch][cl] = Mc; sB[1][ch][cl] = Bc;
    __syncthreads();

    // ---- phase 2: carry-in of this chunk = composition of all later chunks applied to A_T = 0
    double carry_r = 0.0, carry_c = 0.0;
    for (int cc = NCH - 1; cc > ch; --cc) {
        carry_r = sB[0][cc][cl] + sM[0][cc][cl] * carry_r;
        carry_c = sB[1][cc][cl] + sM[1][cc][cl] * carry_c;
    }

    // ---- phase 3: replay with the reference's rounding, store
    if (!active) return;
    for (int t1 = hi - 1; t1 >= lo; t1 -= UB) {
        Step sr[UB], sc[UB];
        float vr[UB], vc[UB];
#pragma unroll
        for (int u = 0; u < UB; ++u) {
            const int t = t1 - u;
            if (t >= lo) {
                load_steps(a, t, col, lvr, lvc, alive_last, sr[u], sc[u]);
                vr[u] = a.vr[(int64_t)t * E + col];
                vc[u] = a.vc[(int64_t)t * E + col];
            }
        }
#pragma unroll
        for (int u = 0; u < UB; ++u) {
            const int t = t1 - u;
            if (t >= lo) {
                const int64_t idx = (int64_t)t * E + col;
                carry_r = __dadd_rn(sr[u].delta, __dmul_rn(sr[u].coef, carry_r));
                carry_c = __dadd_rn(sc[u].delta, __dmul_rn(sc[u].coef, carry_c));
                const float ar = (float)carry_r, ac = (float)carry_c;
                a.adv_r[idx] = ar;
                a.adv_c[idx] = ac;
                a.ret_r[idx] = __fadd_rn(ar, vr[u]);
                a.ret_c[idx] = __fadd_rn(ac, vc[u]);
            }
        }
    }
}

int dual_gae_device(const GaeArgs& a, cudaStream_t st) {
    if (a.T <= 0 || a.E <= 0) return 0;
    // Few columns: narrow column groups (8 per CTA) and many time chunks, so that short buffers still spread over
    // many threads / CTAs (the reference's 2048 x 5 buffer is ONE CTA with 64 chunks of 32 steps).  Many columns:
    // 32-column groups (full 128-byte rows), the grid alone fills the GPU.
    const int g8 = (a.E + 7) / 8, g32 = (a.E + 31) / 32;
    if (g32 >= 4 * sm_count()) {
        if (a.T >= 256) dual_gae_kernel<8, 32><<<g32, 8 * 32, 0, st>>>(a);
        else dual_gae_kernel<1, 32><<<g32, 32, 0, st>>>(a);
    } else if (a.T >= 1024) {
        dual_gae_kernel<16, 8><<<g8, 16 * 32, 0, st>>>(a);
    } else if (a.T >= 128) {
        dual_gae_kernel<4, 8><<<g8, 4 * 32, 0, st>>>(a);
    } else {
        dual_gae_kernel<1, 8><<<g8, 32, 0, st>>>(a);
    }
    ICRL_LAUNCH_CHECK();
    return 0;
}

static int fill_args(GaeArgs& a, const float* rewards, const float* reward_values, const float* costs,
                     const float* cost_values, const float* dones, const float* reward_last_value,
                     const float* cost_last_value, const uint8_t* last_dones, int32_t T, int32_t E, double rg, double rl,
                     double cg, double cl, float* ra, float* rr, float* ca, float* cr) {
    ICRL_CHECK_ARG(T >= 0 && E >= 0, "T/E negative");
    ICRL_CHECK_ARG((int64_t)T * E == 0 || (rewards && reward_values && costs && cost_values && dones && reward_last_value &&
                                           cost_last_value && last_dones && ra && rr && ca && cr),
                   "NULL pointer passed to dual_gae");
    a.r = rewards; a.vr = reward_values; a.c = costs; a.vc = cost_values; a.dones = dones;
    a.last_vr = reward_last_value; a.last_vc = cost_last_value; a.last_dones = last_dones;
    a.adv_r = ra; a.ret_r = rr; a.adv_c = ca; a.ret_c = cr;
    a.T = T; a.E = E;
    // numpy: python-float gamma is cast to float32 when it meets a float32 array; gamma*gae_lambda is a
    // python (float64) product first, then cast.
    a.g_r = (float)rg; a.gl_r = (float)(rg * rl);
    a.g_c = (float)cg; a.gl_c = (float)(cg * cl);
    return 0;
}

}  // namespace icrl

extern "C" {

int icrl_dual_gae(const float* rewards, const float* reward_values, const float* costs, const float* cost_values,
                  const float* dones, const float* reward_last_value, const float* cost_last_value,
                  const uint8_t* last_dones, int32_t T, int32_t E, double reward_gamma, double reward_gae_lambda,
                  double cost_gamma, double cost_gae_lambda, float* reward_advantages, float* reward_returns,
                  float* cost_advantages, float* cost_returns, void* stream) {
    icrl::GaeArgs a;
    int rc = icrl::fill_args(a, rewards, reward_values, costs, cost_values, dones, reward_last_value, cost_last_value,
                             last_dones, T, E, reward_gamma, reward_gae_lambda, cost_gamma, cost_gae_lambda,
                             reward_advantages, reward_returns, cost_advantages, cost_returns);
    if (rc) return rc;
    return icrl::dual_gae_device(a, (cudaStream_t)stream);
}

int icrl_dual_gae_host(const float* rewards, const float* reward_values, const float* costs, const float* cost_values,
                       const float* dones, const float* reward_last_value, const float* cost_last_value,
                       const uint8_t* last_dones, int32_t T, int32_t E, double reward_gamma, double reward_gae_lambda,
                       double cost_gamma, double cost_gae_lambda, float* reward_advantages, float* reward_returns,
                       float* cost_advantages, float* cost_returns, void* stream) {
    icrl::GaeArgs a;
    int rc = icrl::fill_args(a, rewards, reward_values, costs, cost_values, dones, reward_last_value, cost_last_value,
                             last_dones, T, E, reward_gamma, reward_gae_lambda, cost_gamma, cost_gae_lambda,
                             reward_advantages, reward_returns, cost_advantages, cost_returns);
    if (rc) return rc;
    const size_t n = (size_t)T * E;
    if (n == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    // one staging buffer: 5 inputs [T,E] + 2 [E] + last_dones | 4 outputs [T,E]
    const size_t in_bytes = (5 * n + 2 * (size_t)E) * 4 + (size_t)E, out_bytes = 4 * n * 4;
    void *din, *dout;
    if ((rc = icrl::device_scratch(icrl::SLOT_IN0, in_bytes, &din))) return rc;
    if ((rc = icrl::device_scratch(icrl::SLOT_OUT0, out_bytes, &dout))) return rc;
    float* f = (float*)din;
    const float* hin[5] = {rewards, reward_values, costs, cost_values, dones};
    const float** dev_in[5] = {&a.r, &a.vr, &a.c, &a.vc, &a.dones};
    for (int i = 0; i < 5; ++i) {
        ICRL_CUDA(cudaMemcpyAsync(f + i * n, hin[i], n * 4, cudaMemcpyHostToDevice, st));
        *dev_in[i] = f + i * n;
    }
    ICRL_CUDA(cudaMemcpyAsync(f + 5 * n, reward_last_value, (size_t)E * 4, cudaMemcpyHostToDevice, st));
    ICRL_CUDA(cudaMemcpyAsync(f + 5 * n + E, cost_last_value, (size_t)E * 4, cudaMemcpyHostToDevice, st));
    ICRL_CUDA(cudaMemcpyAsync(f + 5 * n + 2 * (size_t)E, last_dones, (size_t)E, cudaMemcpyHostToDevice, st));
    a.last_vr = f + 5 * n;
    a.last_vc = f + 5 * n + E;
    a.last_dones = (const uint8_t*)(f + 5 * n + 2 * (size_t)E);
    float* o = (float*)dout;
    a.adv_r = o; a.ret_r = o + n; a.adv_c = o + 2 * n; a.ret_c = o + 3 * n;
    if ((rc = icrl::dual_gae_device(a, st))) return rc;
    float* hout[4] = {reward_advantages, reward_returns, cost_advantages, cost_returns};
    for (int i = 0; i < 4; ++i) ICRL_CUDA(cudaMemcpyAsync(hout[i], o + i * n, n * 4, cudaMemcpyDeviceToHost, st));
    ICRL_CUDA(cudaStreamSynchronize(st));
    return 0;
}

}  // extern "C"
