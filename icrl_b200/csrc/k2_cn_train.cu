// K2 -- importance-sampling-weighted constraint-net training (full batch), replaces ConstraintNet.train +
// compute_is_weights + th.optim.Adam.step (icrl/constraint_net.py:137-256).
//
// Per backward iteration, four launches on one stream and NO host synchronisation (the early-stop decision is a
// device flag every later kernel checks first):
//   A  cn_forward (K1 kernel, out_kind = prediction) over the nominal rows           -> current_preds
//   B  cn_is_kernel     IS ratios, per-episode products (float64 log-sum), normalised weights, both KLs,
//                       early-stop test (constraint_net.py:231-256, 173-177)
//   C  cn_grad_kernel   forward + backward over nominal and expert tiles; weight gradients are contractions over
//                       the rows of a tile (dW_l = dH_l^T A_{l-1}) accumulated in shared memory, one partial
//                       gradient per CTA (deterministic reduction, no float atomics)
//   D  cn_adam_kernel   reduce partials, loss / prediction statistics, Adam (torch single-tensor rule, eps 1e-5)
// Quirk A: in per-step IS mode the reference's [N,1,1] x [N,1] broadcast makes the nominal loss
// mean(w) * mean(log(p+eps)); we reproduce that value and gradient in O(N).
#include <math.h>

#include "cn_common.cuh"

namespace icrl {

int cn_forward_device(const CnPlan& plan, const void* obs, int obs_is_f64, const float* acs, int64_t n_rows, float* out,
                      int out_kind, cudaStream_t st);

struct CnCtrl {          // device-resident control / result block
    int stopped;         // set by B when a KL threshold is exceeded
    int early_stop_itr;
    int steps_taken;
    int pad;
    float mean_w;        // mean(is_weights) -- the scalar the per-step broadcast collapses to
    float is_mean, is_max, is_min, kl_old_new, kl_new_old;
    float m[14];         // loss / prediction metrics of the last completed iteration (icrl_cn_train_metrics order)
};

enum { ST_NLOG = 0, ST_NWLOG, ST_N1MP, ST_NSUM, ST_NMAX, ST_NMIN, ST_ELOG, ST_E1MP, ST_ESUM, ST_EMAX, ST_EMIN, ST_COUNT = 12 };

// ------------------------------------------------------------------------------------------------ B
__global__ void __launch_bounds__(1024) cn_is_kernel(CnCtrl* ctrl, const float* __restrict__ p_old,
                                                     const float* __restrict__ p_new, const int* __restrict__ offsets,
                                                     int n_episodes, long long n, float eps, int use_is, int per_step,
                                                     float tkon, float tkno, int itr, float* __restrict__ w,
                                                     float* __restrict__ prod) {
    __shared__ double red[32];
    __shared__ float redf[2][32];
    __shared__ double bcast[4];
    if (ctrl->stopped) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = blockDim.x >> 5;
    auto bsum = [&](double v) {
        v = warp_sum(v);
        __syncthreads();
        if (lane == 0) red[warp] = v;
        __syncthreads();
        double t = 0.0;
        for (int i = 0; i < nwarp; ++i) t += red[i];
        return t;
    };
    if (!use_is) {
        for (long long i = tid; i < n; i += blockDim.x) w[i] = 1.f;
        if (tid == 0) { ctrl->mean_w = 1.f; ctrl->is_mean = 1.f; ctrl->is_max = 1.f; ctrl->is_min = 1.f; }
        return;
    }
    // mean ratio
    double s = 0.0;
    for (long long i = tid; i < n; i += blockDim.x) s += (double)((p_new[i] + eps) / (p_old[i] + eps));
    const float mean_ratio = (float)(bsum(s) / (double)n);
    // per-episode products: prod_j = exp(sum log ratio) accumulated in float64 (overflow -> inf, underflow -> 0, as fp32 prod)
    for (int j = warp; j < n_episodes; j += nwarp) {
        double ls = 0.0;
        for (int i = offsets[j] + lane; i < offsets[j + 1]; i += 32) ls += log((double)((p_new[i] + eps) / (p_old[i] + eps)));
        ls = warp_sum(ls);
        if (lane == 0) prod[j] = (float)exp(ls);
    }
    __syncthreads();
    double sp = 0.0;
    for (int j = tid; j < n_episodes; j += blockDim.x) sp += (double)prod[j];
    const float sum_prod = (float)bsum(sp);
    const float prod_mean = sum_prod / (float)n_episodes;
    double k1 = 0.0, k2 = 0.0;
    for (int j = tid; j < n_episodes; j += blockDim.x) {
        const float pj = prod[j], lg = logf(pj + eps);
        k1 += (double)(-lg);
        k2 += (double)((pj - prod_mean) * lg / (prod_mean + eps));
    }
    const float kl_old_new = (float)(bsum(k1) / n_episodes), kl_new_old = (float)(bsum(k2) / n_episodes);
    // weights + their statistics
    double ws = 0.0;
    float wmax = -INFINITY, wmin = INFINITY;
    bool wnan = false;
    if (per_step) {
        for (long long i = tid; i < n; i += blockDim.x) {
            const float wi = ((p_new[i] + eps) / (p_old[i] + eps)) / mean_ratio;
            w[i] = wi; ws += (double)wi; wmax = fmaxf(wmax, wi); wmin = fminf(wmin, wi); wnan |= (wi != wi);
        }
    } else {
        for (int j = warp; j < n_episodes; j += nwarp) {
            const float wj = (float)n_episodes * prod[j] / (sum_prod + eps);
            for (int i = offsets[j] + lane; i < offsets[j + 1]; i += 32) w[i] = wj;
            if (lane == 0) {
                ws += (double)wj * (double)(offsets[j + 1] - offsets[j]);
                wmax = fmaxf(wmax, wj); wmin = fminf(wmin, wj); wnan |= (wj != wj);
            }
        }
    }
    const float mean_w = (float)(bsum(ws) / (double)n);
    wmax = warp_max(wmax); wmin = warp_min(wmin);
    const unsigned anynan = __ballot_sync(0xffffffffu, wnan);
    if (lane == 0) { redf[0][warp] = anynan ? NAN : wmax; redf[1][warp] = anynan ? NAN : wmin; }
    __syncthreads();
    if (tid == 0) {
        float mx = -INFINITY, mn = INFINITY;
        bool nn = false;
        for (int i = 0; i < nwarp; ++i) {
            nn |= (redf[0][i] != redf[0][i]);
            mx = fmaxf(mx, redf[0][i]); mn = fminf(mn, redf[1][i]);
        }
        ctrl->mean_w = mean_w; ctrl->is_mean = mean_w;
        ctrl->is_max = nn ? NAN : mx; ctrl->is_min = nn ? NAN : mn;   // torch.max / min propagate NaN
        ctrl->kl_old_new = kl_old_new; ctrl->kl_new_old = kl_new_old;
        // constraint_net.py:174-177 (NaN / -inf never compare greater)
        if ((tkon != -1.f && kl_old_new > tkon) || (tkno != -1.f && kl_new_old > tkno)) {
            ctrl->stopped = 1;
            ctrl->early_stop_itr = itr;
        }
    }
    (void)bcast;
}

// ------------------------------------------------------------------------------------------------ C
struct GradSmem {
    int bar, mean, rstd, low, high, w[ICRL_MAX_HIDDEN], wb[ICRL_MAX_HIDDEN], b[ICRL_MAX_HIDDEN], wout, raw, acs, xp, hp, dz, g,
        total;
    int LD, KP0;
};

__host__ __device__ inline GradSmem grad_smem_layout(const CnPlan& p, int HP, int TILE, int obs_elem, int n_params) {
    GradSmem s;
    s.LD = TILE + 4;
    s.KP0 = align_up(p.n_select, 4);
    int off = 0;
    s.bar = off; off += 16;
    s.mean = off; off += p.has_norm ? p.obs_dim * 8 : 0;
    s.rstd = off; off += p.has_norm ? p.obs_dim * 8 : 0;
    s.low = off; off += p.has_clip_acs ? p.acs_dim * 4 : 0;
    s.high = off; off += p.has_clip_acs ? p.acs_dim * 4 : 0;
    off = align_up(off, 16);
    for (int l = 0; l < ICRL_MAX_HIDDEN; ++l) {
        s.w[l] = off;                                   // forward copy, k-major [K][HP]
        if (l < p.n_hidden) off += (l == 0 ? p.n_select : HP) * HP * 4;
        s.wb[l] = off;                                  // backward copy, [j][HP] (row j = output unit, cols = inputs), l >= 1
        if (l < p.n_hidden && l >= 1) off += HP * HP * 4;
        s.b[l] = off;
        if (l < p.n_hidden) off += HP * 4;
    }
    s.wout = off; off += (HP + 4) * 4;
    off = align_up(off, 16);
    const int raw_bytes = align_up(TILE * p.obs_dim * obs_elem, 16);
    const int dh_bytes = p.n_hidden * HP * s.LD * 4;    // dH planes alias the raw obs tile (dead after input preparation)
    s.raw = off; off += raw_bytes > dh_bytes ? raw_bytes : dh_bytes;
    s.acs = off; off += align_up(TILE * p.acs_w * 4, 16);
    s.xp = off; off += s.KP0 * s.LD * 4;                // prepared inputs, k-major planes [KP0][LD]
    s.hp = off; off += p.n_hidden * HP * s.LD * 4;      // activations per layer [l][HP][LD]
    s.dz = off; off += s.LD * 4;
    s.g = off; off += align_up(n_params, 4) * 4;        // per-CTA gradient accumulator, flat parameter order
    s.total = off;
    return s;
}

template <typename ObsT, int HP>
__global__ void __launch_bounds__(128) cn_grad_kernel(const __grid_constant__ CnPlan plan, const CnCtrl* __restrict__ ctrl,
                                                      const ObsT* __restrict__ nobs, const float* __restrict__ nacs,
                                                      long long n_nom, const ObsT* __restrict__ eobs,
                                                      const float* __restrict__ eacs, long long n_exp,
                                                      const float* __restrict__ w, int per_step, int gail, float eps,
                                                      float reg, int n_params, int tma_ok, float* __restrict__ part_grad,
                                                      float* __restrict__ part_stats) {
    extern __shared__ __align__(16) unsigned char smem[];
    if (ctrl->stopped) return;
    const int TILE = blockDim.x, tid = threadIdx.x;
    const GradSmem L = grad_smem_layout(plan, HP, TILE, sizeof(ObsT), n_params);
    const int LD = L.LD, KP0 = L.KP0, NH = plan.n_hidden;
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + L.bar);
    float* XP = reinterpret_cast<float*>(smem + L.xp);
    float* HPL = reinterpret_cast<float*>(smem + L.hp);
    float* DH = reinterpret_cast<float*>(smem + L.raw);
    float* DZ = reinterpret_cast<float*>(smem + L.dz);
    float* G = reinterpret_cast<float*>(smem + L.g);
    float* WO = reinterpret_cast<float*>(smem + L.wout);
    if (tid == 0) { mbar_init(bar, 1); fence_mbar_init(); }
    // weights: forward (k-major) copies via the shared loader layout, plus [j][k] copies for the backward pass
    {
        CnSmem F;   // reuse cn_load_weights with our offsets
        F.bar = L.bar; F.mean = L.mean; F.rstd = L.rstd; F.low = L.low; F.high = L.high; F.wout = L.wout;
        for (int l = 0; l < ICRL_MAX_HIDDEN; ++l) { F.w[l] = L.w[l]; F.b[l] = L.b[l]; }
        F.obs = L.raw; F.acs = L.acs; F.h = L.hp; F.total = L.total;
        cn_load_weights<HP>(plan, F, smem);
        const float* src = plan.params;
        int in_dim = plan.n_select;
        for (int l = 0; l < NH; ++l) {
            const int out_dim = plan.hidden[l];
            if (l >= 1) {
                float* WB = reinterpret_cast<float*>(smem + L.wb[l]);
                for (int i = tid; i < HP * HP; i += TILE) {
                    const int j = i / HP, k = i - j * HP;
                    WB[i] = (j < out_dim && k < in_dim) ? src[j * in_dim + k] : 0.f;
                }
            }
            src += out_dim * in_dim + out_dim;
            in_dim = out_dim;
        }
    }
    for (int i = tid; i < align_up(n_params, 4); i += TILE) G[i] = 0.f;
    __syncthreads();

    const float mean_w = ctrl->mean_w;
    const long long tiles_nom = (n_nom + TILE - 1) / TILE, tiles_exp = (n_exp + TILE - 1) / TILE;
    const float inv_nn = 1.f / (float)n_nom, inv_ne = 1.f / (float)n_exp;
    float st[ST_COUNT];
#pragma unroll
    for (int i = 0; i < ST_COUNT; ++i) st[i] = 0.f;
    st[ST_NMAX] = -INFINITY; st[ST_NMIN] = INFINITY; st[ST_EMAX] = -INFINITY; st[ST_EMIN] = INFINITY;
    uint32_t phase = 0;

    CnSmem S;   // staging view for cn_stage_tile / cn_input
    S.bar = L.bar; S.mean = L.mean; S.rstd = L.rstd; S.low = L.low; S.high = L.high; S.obs = L.raw; S.acs = L.acs;

    for (long long tile = blockIdx.x; tile < tiles_nom + tiles_exp; tile += gridDim.x) {
        const bool is_nom = tile < tiles_nom;
        const long long t_in = is_nom ? tile : tile - tiles_nom;
        const long long n_set = is_nom ? n_nom : n_exp;
        const long long row0 = t_in * TILE;
        const int rows = (int)min((long long)TILE, n_set - row0);
        cn_stage_tile<ObsT>(plan, S, smem, is_nom ? nobs : eobs, is_nom ? nacs : eacs, row0, rows, tma_ok && rows == TILE,
                            phase);
        const int r = tid;
        const bool active = r < rows;
        // ---- prepared inputs -> k-major planes
        for (int k = 0; k < KP0; ++k) XP[k * LD + r] = (active && k < plan.n_select) ? cn_input<ObsT>(plan, S, smem, r, k) : 0.f;
        __syncthreads();   // raw tile is dead from here on (DH aliases it)

        // ---- forward, keeping every layer's activations
        float acc[HP];
        {
            const float* B = reinterpret_cast<const float*>(smem + L.b[0]);
#pragma unroll
            for (int j = 0; j < HP; ++j) acc[j] = B[j];
            const float* W = reinterpret_cast<const float*>(smem + L.w[0]);
#pragma unroll 2
            for (int k = 0; k < plan.n_select; ++k) cn_fma_row<HP>(acc, XP[k * LD + r], W + k * HP);
        }
        for (int l = 1; l < NH; ++l) {
            float* Hp = HPL + (l - 1) * HP * LD;
#pragma unroll
            for (int j = 0; j < HP; ++j) Hp[j * LD + r] = fmaxf(acc[j], 0.f);
            const float* B = reinterpret_cast<const float*>(smem + L.b[l]);
#pragma unroll
            for (int j = 0; j < HP; ++j) acc[j] = B[j];
            const float* W = reinterpret_cast<const float*>(smem + L.w[l]);
            const int kin = plan.hidden[l - 1];
#pragma unroll 4
            for (int k = 0; k < kin; ++k) cn_fma_row<HP>(acc, Hp[k * LD + r], W + k * HP);
        }
        float z = WO[HP];
        {
            float* Hp = HPL + (NH - 1) * HP * LD;
#pragma unroll
            for (int j = 0; j < HP; ++j) {
                const float a = fmaxf(acc[j], 0.f);
                Hp[j * LD + r] = a;
                z = fmaf(a, WO[j], z);
            }
        }
        // ---- loss terms and dL/dz for this row
        const float pr = sigmoidf_ref(z);
        float dz = 0.f;
        if (active) {
            float dLdp;
            if (is_nom) {
                const float wi = per_step ? mean_w : w[row0 + r];
                if (gail) {
                    const float l1 = fmaxf(logf(1.f - pr), -100.f);            // nn.BCELoss clamps log at -100
                    st[ST_NLOG] += -l1; st[ST_NWLOG] += -l1;
                    dLdp = (l1 > -100.f ? 1.f / (1.f - pr) : 0.f) * inv_nn;
                } else {
                    const float lg = logf(pr + eps);
                    st[ST_NLOG] += lg; st[ST_NWLOG] += wi * lg;
                    dLdp = wi * inv_nn / (pr + eps) - reg * inv_nn;
                }
                st[ST_N1MP] += 1.f - pr; st[ST_NSUM] += pr;
                st[ST_NMAX] = fmaxf(st[ST_NMAX], pr); st[ST_NMIN] = fminf(st[ST_NMIN], pr);
            } else {
                if (gail) {
                    const float l1 = fmaxf(logf(pr), -100.f);
                    st[ST_ELOG] += -l1;
                    dLdp = (l1 > -100.f ? -1.f / pr : 0.f) * inv_ne;
                } else {
                    st[ST_ELOG] += logf(pr + eps);
                    dLdp = -inv_ne / (pr + eps) - reg * inv_ne;
                }
                st[ST_E1MP] += 1.f - pr; st[ST_ESUM] += pr;
                st[ST_EMAX] = fmaxf(st[ST_EMAX], pr); st[ST_EMIN] = fminf(st[ST_EMIN], pr);
            }
            dz = dLdp * pr * (1.f - pr);
        }
        DZ[r] = dz;
        // ---- backward through the hidden layers (thread-per-row); dH planes [l][HP][LD]
        {
            float* Hp = HPL + (NH - 1) * HP * LD;
            float* Dp = DH + (NH - 1) * HP * LD;
#pragma unroll
            for (int j = 0; j < HP; ++j) Dp[j * LD + r] = (Hp[j * LD + r] > 0.f) ? dz * WO[j] : 0.f;
        }
        for (int l = NH - 1; l >= 1; --l) {
            const float* Dp = DH + l * HP * LD;
            const float* WB = reinterpret_cast<const float*>(smem + L.wb[l]);
            float dacc[HP];
#pragma unroll
            for (int k = 0; k < HP; ++k) dacc[k] = 0.f;
            const int jout = plan.hidden[l];
#pragma unroll 4
            for (int j = 0; j < jout; ++j) cn_fma_row<HP>(dacc, Dp[j * LD + r], WB + j * HP);
            const float* Hq = HPL + (l - 1) * HP * LD;
            float* Dq = DH + (l - 1) * HP * LD;
#pragma unroll
            for (int k = 0; k < HP; ++k) Dq[k * LD + r] = (Hq[k * LD + r] > 0.f) ? dacc[k] : 0.f;
        }
        __syncthreads();

        // ---- weight gradients: contractions over the tile's rows, accumulated into the CTA's flat gradient
        {
            int goff = 0, in_dim = plan.n_select;
            for (int l = 0; l < NH; ++l) {
                const int out_dim = plan.hidden[l];
                const float* Dp = DH + l * HP * LD;
                const float* Ap = (l == 0) ? XP : HPL + (l - 1) * HP * LD;
                const int ktiles = (in_dim + 3) / 4, jtiles = (out_dim + 3) / 4;
                for (int t = tid; t < ktiles * jtiles; t += TILE) {
                    const int tj = t % jtiles, tk = t / jtiles;
                    float a4[4][4];
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int j = 0; j < 4; ++j) a4[i][j] = 0.f;
                    for (int rr = 0; rr < TILE; rr += 4) {
                        float4 d[4], x[4];
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            d[i] = *reinterpret_cast<const float4*>(Dp + (4 * tj + i) * LD + rr);
                            x[i] = *reinterpret_cast<const float4*>(Ap + (4 * tk + i) * LD + rr);
                        }
#pragma unroll
                        for (int i = 0; i < 4; ++i)
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                a4[i][j] = fmaf(d[i].x, x[j].x, a4[i][j]);
                                a4[i][j] = fmaf(d[i].y, x[j].y, a4[i][j]);
                                a4[i][j] = fmaf(d[i].z, x[j].z, a4[i][j]);
                                a4[i][j] = fmaf(d[i].w, x[j].w, a4[i][j]);
                            }
                    }
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const int jo = 4 * tj + i, ki = 4 * tk + j;
                            if (jo < out_dim && ki < in_dim) G[goff + jo * in_dim + ki] += a4[i][j];
                        }
                }
                for (int j = tid; j < out_dim; j += TILE) {
                    float s = 0.f;
                    for (int rr = 0; rr < TILE; rr += 4) {
                        const float4 d = *reinterpret_cast<const float4*>(Dp + j * LD + rr);
                        s += (d.x + d.y) + (d.z + d.w);
                    }
                    G[goff + out_dim * in_dim + j] += s;
                }
                goff += out_dim * in_dim + out_dim;
                in_dim = out_dim;
            }
            // output layer: dw_out[j] = sum_r dz[r] * h_last[j][r], db_out = sum_r dz[r]
            const float* Hp = HPL + (NH - 1) * HP * LD;
            for (int j = tid; j <= in_dim; j += TILE) {
                float s = 0.f;
                if (j < in_dim) {
                    for (int rr = 0; rr < TILE; rr += 4) {
                        const float4 d = *reinterpret_cast<const float4*>(DZ + rr);
                        const float4 h = *reinterpret_cast<const float4*>(Hp + j * LD + rr);
                        s = fmaf(d.x, h.x, s); s = fmaf(d.y, h.y, s); s = fmaf(d.z, h.z, s); s = fmaf(d.w, h.w, s);
                    }
                } else {
                    for (int rr = 0; rr < TILE; ++rr) s += DZ[rr];
                }
                G[goff + j] += s;
            }
        }
        __syncthreads();   // planes are rewritten by the next tile
    }

    // ---- per-CTA partial gradient and statistics
    for (int i = tid; i < n_params; i += TILE) part_grad[(size_t)blockIdx.x * n_params + i] = G[i];
    __shared__ float sred[ST_COUNT][4];
    const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
    for (int i = 0; i < ST_COUNT; ++i) {
        float v = st[i];
        if (i == ST_NMAX || i == ST_EMAX) v = warp_max(v);
        else if (i == ST_NMIN || i == ST_EMIN) v = warp_min(v);
        else v = warp_sum(v);
        if (lane == 0) sred[i][warp] = v;
    }
    __syncthreads();
    if (tid < ST_COUNT) {
        const int nw = (TILE + 31) / 32;
        float v = sred[tid][0];
        for (int i = 1; i < nw; ++i) {
            if (tid == ST_NMAX || tid == ST_EMAX) v = fmaxf(v, sred[tid][i]);
            else if (tid == ST_NMIN || tid == ST_EMIN) v = fminf(v, sred[tid][i]);
            else v += sred[tid][i];
        }
        part_stats[(size_t)blockIdx.x * ST_COUNT + tid] = v;
    }
}

// ------------------------------------------------------------------------------------------------ D
__global__ void __launch_bounds__(1024) cn_adam_kernel(CnCtrl* ctrl, float* __restrict__ params, float* __restrict__ adam_m,
                                                       float* __restrict__ adam_v, const float* __restrict__ part_grad,
                                                       const float* __restrict__ part_stats, int n_parts, int n_params,
                                                       long long n_nom, long long n_exp, int per_step, int gail,
                                                       float reg, double lr, double beta1, double beta2, double adam_eps,
                                                       long long step_before) {
    if (ctrl->stopped) return;
    const int tid = threadIdx.x;
    const long long step = step_before + ctrl->steps_taken + 1;
    __syncthreads();   // everyone has read steps_taken before thread 0 advances it
    const double bc1 = 1.0 - pow(beta1, (double)step), bc2 = 1.0 - pow(beta2, (double)step);
    const float omb1 = (float)(1.0 - beta1), b2 = (float)beta2, omb2 = (float)(1.0 - beta2);
    const float bc2s = (float)sqrt(bc2), epsf = (float)adam_eps, nss = (float)(-(lr / bc1));
    for (int i = tid; i < n_params; i += blockDim.x) {
        float g = 0.f;
        for (int b = 0; b < n_parts; ++b) g += part_grad[(size_t)b * n_params + i];
        float m = adam_m[i], v = adam_v[i];
        m = fmaf(omb1, g - m, m);
        v = fmaf(omb2 * g, g, v * b2);
        const float denom = sqrtf(v) / bc2s + epsf;
        params[i] = fmaf(nss, m / denom, params[i]);
        adam_m[i] = m; adam_v[i] = v;
    }
    if (tid == 0) {
        double s[ST_COUNT];
        for (int i = 0; i < ST_COUNT; ++i) s[i] = 0.0;
        float nmax = -INFINITY, nmin = INFINITY, emax = -INFINITY, emin = INFINITY;
        for (int b = 0; b < n_parts; ++b) {
            const float* ps = part_stats + (size_t)b * ST_COUNT;
            for (int i = 0; i < ST_COUNT; ++i) s[i] += (double)ps[i];
            nmax = fmaxf(nmax, ps[ST_NMAX]); nmin = fminf(nmin, ps[ST_NMIN]);
            emax = fmaxf(emax, ps[ST_EMAX]); emin = fminf(emin, ps[ST_EMIN]);
        }
        const double nn = (double)n_nom, ne = (double)n_exp;
        const float expert_loss = (float)(s[ST_ELOG] / ne), unweighted = (float)(s[ST_NLOG] / nn);
        float nominal_loss, reg_loss, loss;
        if (gail) {
            nominal_loss = unweighted; reg_loss = 0.f;
            loss = nominal_loss + expert_loss;
        } else {
            nominal_loss = per_step ? ctrl->mean_w * unweighted : (float)(s[ST_NWLOG] / nn);
            reg_loss = reg * ((float)(s[ST_E1MP] / ne) + (float)(s[ST_N1MP] / nn));
            loss = (-expert_loss + nominal_loss) + reg_loss;
        }
        float* m = ctrl->m;
        m[0] = loss; m[1] = expert_loss; m[2] = unweighted; m[3] = nominal_loss; m[4] = reg_loss;
        m[8] = nmax; m[9] = nmin; m[10] = (float)(s[ST_NSUM] / nn);
        m[11] = emax; m[12] = emin; m[13] = (float)(s[ST_ESUM] / ne);
        ctrl->steps_taken += 1;
    }
}

template <typename ObsT, int HP>
static int launch_grad(const CnPlan& plan, const CnCtrl* ctrl, const void* nobs, const float* nacs, int64_t n_nom,
                       const void* eobs, const float* eacs, int64_t n_exp, const float* w, const icrl_cn_train_cfg& cfg,
                       int n_params, float** part_grad, float** part_stats, int* n_parts, cudaStream_t st) {
    auto kern = cn_grad_kernel<ObsT, HP>;
    int tile = 128;
    GradSmem L = grad_smem_layout(plan, HP, tile, sizeof(ObsT), n_params);
    while (L.total > 220 * 1024 && tile > 32) {
        tile /= 2;
        L = grad_smem_layout(plan, HP, tile, sizeof(ObsT), n_params);
    }
    if (L.total > 220 * 1024) {
        set_error("constraint net too large for the training kernel's shared memory (%d bytes)", L.total);
        return ICRL_EUNSUPPORTED;
    }
    ICRL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L.total));
    int per_sm = 1;
    ICRL_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, tile, L.total));
    if (per_sm < 1) per_sm = 1;
    const int64_t tiles = (n_nom + tile - 1) / tile + (n_exp + tile - 1) / tile;
    const int64_t cap = (int64_t)per_sm * sm_count();
    const int grid = (int)(tiles < cap ? tiles : cap);
    int rc;
    void *pg, *ps;
    if ((rc = device_scratch(SLOT_WORK0, (size_t)grid * n_params * 4, &pg))) return rc;
    if ((rc = device_scratch(SLOT_WORK1, (size_t)grid * ST_COUNT * 4, &ps))) return rc;
    *part_grad = (float*)pg; *part_stats = (float*)ps; *n_parts = grid;
    const int tma_ok = ((reinterpret_cast<uintptr_t>(nobs) | reinterpret_cast<uintptr_t>(nacs) |
                         reinterpret_cast<uintptr_t>(eobs) | reinterpret_cast<uintptr_t>(eacs)) & 15u) == 0;
    kern<<<grid, tile, L.total, st>>>(plan, ctrl, (const ObsT*)nobs, nacs, n_nom, (const ObsT*)eobs, eacs, n_exp, w,
                                      cfg.per_step_is && cfg.importance_sampling, cfg.train_gail_lambda, cfg.eps,
                                      cfg.regularizer_coeff, n_params, tma_ok, *part_grad, *part_stats);
    ICRL_LAUNCH_CHECK();
    return 0;
}

template <typename ObsT>
static int dispatch_grad(const CnPlan& plan, const CnCtrl* ctrl, const void* nobs, const float* nacs, int64_t n_nom,
                         const void* eobs, const float* eacs, int64_t n_exp, const float* w, const icrl_cn_train_cfg& cfg,
                         int n_params, float** pg, float** ps, int* np, cudaStream_t st) {
#define ICRL_GRAD_CASE(W) \
    case W: return launch_grad<ObsT, W>(plan, ctrl, nobs, nacs, n_nom, eobs, eacs, n_exp, w, cfg, n_params, pg, ps, np, st)
    switch (cn_padded_width(plan)) {
        ICRL_GRAD_CASE(8); ICRL_GRAD_CASE(16); ICRL_GRAD_CASE(24); ICRL_GRAD_CASE(32);
        ICRL_GRAD_CASE(40); ICRL_GRAD_CASE(48); ICRL_GRAD_CASE(64);
    }
#undef ICRL_GRAD_CASE
    set_error("unsupported constraint-net width");
    return ICRL_EUNSUPPORTED;
}

}  // namespace icrl

extern "C" int icrl_cn_train(const icrl_cn_desc* d, const icrl_cn_train_cfg* cfg, const void* nominal_obs,
                             int32_t nominal_obs_is_f64, const float* nominal_acs, int64_t n_nominal,
                             const int32_t* episode_offsets, int32_t n_episodes, const void* expert_obs,
                             int32_t expert_obs_is_f64, const float* expert_acs, int64_t n_expert, float* adam_m,
                             float* adam_v, int64_t* adam_step, icrl_cn_train_metrics* metrics, void* stream) {
    using namespace icrl;
    CnPlan plan;
    int rc = make_plan(d, &plan);
    if (rc) return rc;
    ICRL_CHECK_ARG(cfg && metrics && adam_m && adam_v && adam_step, "NULL pointer passed to icrl_cn_train");
    ICRL_CHECK_ARG(n_nominal > 0 && n_expert > 0 && nominal_obs && nominal_acs && expert_obs && expert_acs,
                   "empty nominal or expert batch");
    ICRL_CHECK_ARG(nominal_obs_is_f64 == expert_obs_is_f64, "nominal and expert observations must share a dtype");
    ICRL_CHECK_ARG(!cfg->importance_sampling || (episode_offsets && n_episodes > 0), "importance sampling needs episodes");
    ICRL_CHECK_ARG(cfg->iterations >= 0, "iterations < 0");
    cudaStream_t st = (cudaStream_t)stream;
    const int n_params = (int)cn_param_count(plan);
    void *wbuf, *misc;
    // work buffer: start_preds | cur_preds | w  (each n_nominal), then prod[n_episodes]; misc: CnCtrl
    if ((rc = device_scratch(SLOT_WORK2, ((size_t)3 * n_nominal + (size_t)(n_episodes > 0 ? n_episodes : 1)) * 4, &wbuf)))
        return rc;
    if ((rc = device_scratch(SLOT_WORK3, sizeof(CnCtrl), &misc))) return rc;
    float* start_preds = (float*)wbuf;
    float* cur_preds = start_preds + n_nominal;
    float* w = cur_preds + n_nominal;
    float* prod = w + n_nominal;
    CnCtrl* ctrl = (CnCtrl*)misc;
    ICRL_CUDA(cudaMemsetAsync(ctrl, 0, sizeof(CnCtrl), st));
    CnCtrl init = {};
    init.early_stop_itr = cfg->iterations;
    init.mean_w = 1.f;
    ICRL_CUDA(cudaMemcpyAsync(ctrl, &init, sizeof(CnCtrl), cudaMemcpyHostToDevice, st));

    if (cfg->importance_sampling) {   // start_preds (constraint_net.py:159-162)
        if ((rc = cn_forward_device(plan, nominal_obs, nominal_obs_is_f64, nominal_acs, n_nominal, start_preds, 1, st)))
            return rc;
    }
    float *part_grad = nullptr, *part_stats = nullptr;
    int n_parts = 0;
    for (int itr = 0; itr < cfg->iterations; ++itr) {
        if (cfg->importance_sampling) {
            if (itr == 0) {
                ICRL_CUDA(cudaMemcpyAsync(cur_preds, start_preds, (size_t)n_nominal * 4, cudaMemcpyDeviceToDevice, st));
            } else if ((rc = cn_forward_device(plan, nominal_obs, nominal_obs_is_f64, nominal_acs, n_nominal, cur_preds, 1,
                                               st))) {
                return rc;   // (a stopped run still pays this forward; it is tiny and keeps the host free of syncs)
            }
        }
        cn_is_kernel<<<1, 1024, 0, st>>>(ctrl, start_preds, cur_preds, episode_offsets, n_episodes, (long long)n_nominal,
                                         cfg->eps, cfg->importance_sampling, cfg->per_step_is, cfg->target_kl_old_new,
                                         cfg->target_kl_new_old, itr, w, prod);
        ICRL_LAUNCH_CHECK();
        rc = nominal_obs_is_f64
                 ? dispatch_grad<double>(plan, ctrl, nominal_obs, nominal_acs, n_nominal, expert_obs, expert_acs, n_expert, w,
                                         *cfg, n_params, &part_grad, &part_stats, &n_parts, st)
                 : dispatch_grad<float>(plan, ctrl, nominal_obs, nominal_acs, n_nominal, expert_obs, expert_acs, n_expert, w,
                                        *cfg, n_params, &part_grad, &part_stats, &n_parts, st);
        if (rc) return rc;
        cn_adam_kernel<<<1, 1024, 0, st>>>(ctrl, const_cast<float*>(plan.params), adam_m, adam_v, part_grad, part_stats,
                                           n_parts, n_params, (long long)n_nominal, (long long)n_expert,
                                           cfg->per_step_is && cfg->importance_sampling, cfg->train_gail_lambda,
                                           cfg->regularizer_coeff, cfg->lr, cfg->adam_beta1, cfg->adam_beta2, cfg->adam_eps,
                                           (long long)*adam_step);
        ICRL_LAUNCH_CHECK();
    }
    CnCtrl out;
    ICRL_CUDA(cudaMemcpyAsync(&out, ctrl, sizeof(CnCtrl), cudaMemcpyDeviceToHost, st));
    ICRL_CUDA(cudaStreamSynchronize(st));
    *adam_step += out.steps_taken;
    metrics->cn_loss = out.m[0]; metrics->expert_loss = out.m[1]; metrics->unweighted_nominal_loss = out.m[2];
    metrics->nominal_loss = out.m[3]; metrics->regularizer_loss = out.m[4];
    metrics->is_mean = out.is_mean; metrics->is_max = out.is_max; metrics->is_min = out.is_min;
    metrics->nominal_preds_max = out.m[8]; metrics->nominal_preds_min = out.m[9]; metrics->nominal_preds_mean = out.m[10];
    metrics->expert_preds_max = out.m[11]; metrics->expert_preds_min = out.m[12]; metrics->expert_preds_mean = out.m[13];
    metrics->kl_old_new = out.kl_old_new; metrics->kl_new_old = out.kl_new_old;
    metrics->early_stop_itr = out.early_stop_itr; metrics->steps_taken = out.steps_taken;
    return 0;
}
