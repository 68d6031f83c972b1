// K3 -- dual reward/cost GAE as a single-pass, register-resident segmented reverse scan.
// Replaces RolloutBufferWithCost._compute_returns_and_advantage x2
// (stable_baselines3/common/buffers.py:493-552), a Python loop over n_steps.
//
// A_t = delta_t + c_t * A_{t+1} is an affine recurrence; `done` flags make it segmented (c_t = 0).
// Every transition is read ONCE and written once (36 B, the algorithmic minimum):
//   * a thread owns 8 consecutive steps of one env column and keeps their raw float32 inputs in registers (43 independent
//     loads issued up front: the memory system sees all of them at once, the recurrence none);
//   * it folds its 8 steps into an affine map (M, B):  A_lo = B + M * A_hi  (float64), per signal;
//   * lane groups of a warp = consecutive 8-step blocks of the same columns -> warp-shuffle scan of the maps, warp aggregates
//     meet in shared memory, a CTA window (NT / CG blocks of 8 steps) hands its carry to the next window (earlier in time);
//   * the 8 steps are then replayed FROM REGISTERS with the reference's exact operation order and dtypes: delta in float32
//     (rounded per op, no FMA contraction), carry in float64, one rounding to float32 per stored advantage,
//     returns = adv_f32 + value_f32.
// Reward and cost scans run in the same thread (two independent dependency chains).  CG adjacent columns per CTA: 8 (every
// row access of a lane group is one full 32-byte sector) for wide buffers, 4 for narrow ones (twice the time slots per CTA).
#include <stdlib.h>

#include <stdlib.h>
#include <string.h>

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

constexpr int BS = 8;   // steps per thread

struct Map2 { double Mr, Br, Mc, Bc; };     // affine maps of the two signals

__device__ __forceinline__ double shfl_up_d(double v, int delta) { return __shfl_up_sync(0xffffffffu, v, delta); }

// NT threads, CG adjacent env columns per CTA; a window = (NT / CG) blocks of BS steps.
template <int NT, int CG>
__global__ void __launch_bounds__(NT, NT <= 128 ? 4 : NT <= 256 ? 2 : 1) dual_gae_kernel(const GaeArgs a) {
    constexpr int LG = 32 / CG, NW = NT / 32, SLOTS = NW * LG;
    __shared__ double sAgg[NW][CG][4];
    __shared__ double sCarry[2][CG][2];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int cl = lane % CG, lg = lane / CG;        // column within the group, time slot within the warp
    const int slot = w * LG + lg;                    // slot 0 = the LATEST block of a window
    const int col = blockIdx.x * CG + cl;
    const bool active_col = col < a.E;
    const int T = a.T, E = a.E;
    const int nb = (T + BS - 1) / BS, nwin = (nb + SLOTS - 1) / SLOTS;

    double alive_last = 0.0;
    float lvr = 0.f, lvc = 0.f;
    if (active_col) {
        alive_last = a.last_dones[col] ? 0.0 : 1.0;
        lvr = a.last_vr[col];
        lvc = a.last_vc[col];
    }
    if (threadIdx.x < CG * 2) sCarry[0][threadIdx.x >> 1][threadIdx.x & 1] = 0.0;   // A_T = 0
    __syncthreads();

    for (int k = 0; k < nwin; ++k) {
        const int tb = nb - 1 - k * SLOTS - slot;
        const bool act = active_col && tb >= 0;
        const int t0 = tb * BS;
        // ---- raw inputs of the block, all loads independent.  vr / vc carry one extra row (the next state's value), al[u] is
        // next_non_terminal after step t0 + u (float32 path); the step that ends the buffer uses the float64 path below.
        float r[BS], c[BS], vr[BS + 1], vc[BS + 1], al[BS];
        if (act) {
#pragma unroll
            for (int u = 0; u <= BS; ++u) {
                const int t = t0 + u;
                const int64_t idx = (int64_t)t * E + col;
                const bool in = t < T;
                vr[u] = in ? a.vr[idx] : lvr;
                vc[u] = in ? a.vc[idx] : lvc;
                if (u < BS) { r[u] = in ? a.r[idx] : 0.f; c[u] = in ? a.c[idx] : 0.f; }
                if (u > 0) al[u - 1] = in ? __fsub_rn(1.0f, a.dones[idx]) : 0.f;
            }
        }
        // (delta, coef) of step u exactly as numpy evaluates buffers.py:528-537
        auto step = [&](int u, float rew, float val, float nval, float g32, float gl32) {
            Step s;
            if (t0 + u != T - 1) {
                const float t1 = __fmul_rn(g32, nval);
                const float t2 = __fmul_rn(t1, al[u]);
                const float t3 = __fadd_rn(rew, t2);
                s.delta = (double)__fsub_rn(t3, val);
                s.coef = (double)__fmul_rn(gl32, al[u]);
            } else {       // next_non_terminal comes from a bool array -> float64 from here on
                const double t2 = __dmul_rn((double)__fmul_rn(g32, nval), alive_last);
                s.delta = __dsub_rn(__dadd_rn((double)rew, t2), (double)val);
                s.coef = 0.0;   // multiplies the initial carry 0
            }
            return s;
        };
        // ---- fold the block into one affine map per signal
        Map2 m = {1.0, 0.0, 1.0, 0.0};
        if (act) {
#pragma unroll
            for (int u = BS - 1; u >= 0; --u) {
                if (t0 + u < T) {
                    const Step sr = step(u, r[u], vr[u], vr[u + 1], a.g_r, a.gl_r);
                    const Step sc = step(u, c[u], vc[u], vc[u + 1], a.g_c, a.gl_c);
                    m.Br = sr.delta + sr.coef * m.Br; m.Mr = sr.coef * m.Mr;
                    m.Bc = sc.delta + sc.coef * m.Bc; m.Mc = sc.coef * m.Mc;
                }
            }
        }
        // ---- inclusive scan over the warp's time slots (slot order: later blocks first).  mine o theirs.
        Map2 inc = m;
#pragma unroll
        for (int off = 1; off < LG; off <<= 1) {
            const double pMr = shfl_up_d(inc.Mr, off * CG), pBr = shfl_up_d(inc.Br, off * CG);
            const double pMc = shfl_up_d(inc.Mc, off * CG), pBc = shfl_up_d(inc.Bc, off * CG);
            if (lg >= off) {
                inc.Br = inc.Br + inc.Mr * pBr; inc.Mr = inc.Mr * pMr;
                inc.Bc = inc.Bc + inc.Mc * pBc; inc.Mc = inc.Mc * pMc;
            }
        }
        Map2 exc;       // composition of the warp's earlier slots (identity for the first)
        exc.Mr = shfl_up_d(inc.Mr, CG); exc.Br = shfl_up_d(inc.Br, CG);
        exc.Mc = shfl_up_d(inc.Mc, CG); exc.Bc = shfl_up_d(inc.Bc, CG);
        if (lg == 0) exc = Map2{1.0, 0.0, 1.0, 0.0};
        if (lg == LG - 1) { sAgg[w][cl][0] = inc.Mr; sAgg[w][cl][1] = inc.Br; sAgg[w][cl][2] = inc.Mc; sAgg[w][cl][3] = inc.Bc; }
        __syncthreads();
        // ---- carry into this warp: the window's carry pushed through the earlier warps' aggregates
        double xr = sCarry[k & 1][cl][0], xc = sCarry[k & 1][cl][1];
        for (int ww = 0; ww < w; ++ww) {
            xr = sAgg[ww][cl][1] + sAgg[ww][cl][0] * xr;
            xc = sAgg[ww][cl][3] + sAgg[ww][cl][2] * xc;
        }
        if (w == NW - 1 && lg == LG - 1) {        // the window's carry-out (into the next, earlier window)
            sCarry[(k + 1) & 1][cl][0] = inc.Br + inc.Mr * xr;
            sCarry[(k + 1) & 1][cl][1] = inc.Bc + inc.Mc * xc;
        }
        double carry_r = exc.Br + exc.Mr * xr, carry_c = exc.Bc + exc.Mc * xc;
        // ---- replay from registers with the reference's rounding, store
        if (act) {
#pragma unroll
            for (int u = BS - 1; u >= 0; --u) {
                if (t0 + u < T) {
                    const int64_t idx = (int64_t)(t0 + u) * E + col;
                    const Step sr = step(u, r[u], vr[u], vr[u + 1], a.g_r, a.gl_r);
                    const Step sc = step(u, c[u], vc[u], vc[u + 1], a.g_c, a.gl_c);
                    carry_r = __dadd_rn(sr.delta, __dmul_rn(sr.coef, carry_r));
                    carry_c = __dadd_rn(sc.delta, __dmul_rn(sc.coef, carry_c));
                    const float ar = (float)carry_r, ac = (float)carry_c;
                    a.adv_r[idx] = ar;
                    a.adv_c[idx] = ac;
                    a.ret_r[idx] = __fadd_rn(ar, vr[u]);
                    a.ret_c[idx] = __fadd_rn(ac, vc[u]);
                }
            }
        }
        __syncthreads();   // the next window's carry is in place; sAgg may be rewritten
    }
}

int dual_gae_device(const GaeArgs& a, cudaStream_t st) {
    if (a.T <= 0 || a.E <= 0) return 0;
    // launch shape by column count (measured, profiles/SUMMARY_r02.md): the reference's 2048 x 5 rollout: 4-column groups and
    // 1024-step windows so that two CTAs cover the time axis quickly; < 1024 columns: 4-column groups (twice the CTAs);
    // wide buffers: 8-column groups (a lane group's row access is one full 32-byte sector), 256-step windows, two CTAs per SM,
    // and from 4096 columns on 128-thread CTAs, four per SM (load and replay phases of different CTAs overlap better)
    static const int forced_nt = getenv("ICRL_K3_NT") ? atoi(getenv("ICRL_K3_NT")) : 0;     // A/B switches (profiling)
    static const int forced_cg = getenv("ICRL_K3_CG") ? atoi(getenv("ICRL_K3_CG")) : 0;
    const bool many_cols = a.E >= 4096;      // enough column groups to fill every SM four times with small CTAs
    if (a.E < 64) dual_gae_kernel<512, 4><<<(a.E + 3) / 4, 512, 0, st>>>(a);
    else if (forced_cg == 4 || (forced_cg == 0 && a.E < 1024)) dual_gae_kernel<256, 4><<<(a.E + 3) / 4, 256, 0, st>>>(a);
    else if (forced_nt == 128 || (forced_nt == 0 && many_cols)) dual_gae_kernel<128, 8><<<(a.E + 7) / 8, 128, 0, st>>>(a);
    else dual_gae_kernel<256, 8><<<(a.E + 7) / 8, 256, 0, st>>>(a);
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
    // rollout-sized buffers (2048 x 5 in every shipped config): the kernel reads every input once and writes every output
    // once, so it works straight on a host-mapped pinned block (UVA) -- one launch + one synchronisation instead of eight
    // staged H2D copies, the launch, four D2H copies and the synchronisation
    const bool zero_copy = in_bytes + out_bytes <= (1u << 20) && getenv("ICRL_K3_NO_ZEROCOPY") == nullptr;
    if (zero_copy) {
        const size_t in_al = (in_bytes + 15) / 16 * 16;
        void* pin;
        if ((rc = icrl::pinned_scratch(icrl::SLOT_WORK0, in_al + out_bytes, &pin))) return rc;
        din = pin;
        dout = static_cast<unsigned char*>(pin) + in_al;
    } else {
        if ((rc = icrl::device_scratch(icrl::SLOT_IN0, in_bytes, &din))) return rc;
        if ((rc = icrl::device_scratch(icrl::SLOT_OUT0, out_bytes, &dout))) return rc;
    }
    float* f = (float*)din;
    if (zero_copy) {
        const float* hin0[5] = {rewards, reward_values, costs, cost_values, dones};
        for (int i = 0; i < 5; ++i) memcpy(f + i * n, hin0[i], n * 4);
        memcpy(f + 5 * n, reward_last_value, (size_t)E * 4);
        memcpy(f + 5 * n + E, cost_last_value, (size_t)E * 4);
        memcpy(f + 5 * n + 2 * (size_t)E, last_dones, (size_t)E);
        a.r = f; a.vr = f + n; a.c = f + 2 * n; a.vc = f + 3 * n; a.dones = f + 4 * n;
        a.last_vr = f + 5 * n;
        a.last_vc = f + 5 * n + E;
        a.last_dones = (const uint8_t*)(f + 5 * n + 2 * (size_t)E);
        float* o = (float*)dout;
        a.adv_r = o; a.ret_r = o + n; a.adv_c = o + 2 * n; a.ret_c = o + 3 * n;
        if ((rc = icrl::dual_gae_device(a, st))) return rc;
        ICRL_CUDA(cudaStreamSynchronize(st));
        float* hout0[4] = {reward_advantages, reward_returns, cost_advantages, cost_returns};
        for (int i = 0; i < 4; ++i) memcpy(hout0[i], o + i * n, n * 4);
        return 0;
    }
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
