// K1 -- fused constraint-net forward over a whole buffer of transitions (cost relabel).
// Replaces ConstraintNet.cost_function -> prepare_data -> nn.Sequential (icrl/constraint_net.py:121-130,
// 258-299, 101-119), called per env step by VecCostWrapper.step_wait (vec_cost_wrapper.py:51-66).
//
// Roofline: reads len(select_dim)*4 B and writes 4 B per transition (Ant: 488 B); layer 0 costs
// 2*n_select*h1 flops per row, so LGW/HC shapes are HBM-bound and Ant (121->40->40->1, 27 flop/B) is
// FP32-FMA-bound on CUDA cores.  Persistent grid: (resident CTAs per SM) x (SM count) CTAs loop over tiles.
#include <stdlib.h>
#include <string.h>

#include "cn_common.cuh"

namespace icrl {

// cleared around launches whose inputs live in host-mapped pinned memory (the small-batch host path): plain loads, no bulk copies
static thread_local bool g_k1_allow_tma = true;

int cn_padded_width(const CnPlan& p) {
    int m = 1;
    for (int l = 0; l < p.n_hidden; ++l) m = p.hidden[l] > m ? p.hidden[l] : m;
    const int widths[] = {8, 16, 24, 32, 40, 48, 64};
    for (int w : widths)
        if (m <= w) return w;
    return -1;
}

int64_t cn_param_count(const CnPlan& p) {
    int64_t n = 0;
    int in_dim = p.n_select;
    for (int l = 0; l < p.n_hidden; ++l) {
        n += (int64_t)p.hidden[l] * in_dim + p.hidden[l];
        in_dim = p.hidden[l];
    }
    return n + in_dim + 1;
}

int make_plan(const icrl_cn_desc* d, CnPlan* p) {
    ICRL_CHECK_ARG(d != nullptr, "cn desc is NULL");
    ICRL_CHECK_ARG(d->obs_dim > 0 && d->acs_dim > 0, "obs_dim/acs_dim must be positive");
    ICRL_CHECK_ARG(d->n_select > 0 && d->n_select <= ICRL_MAX_SELECT, "n_select %d out of range (1..%d)", d->n_select,
                   ICRL_MAX_SELECT);
    ICRL_CHECK_ARG(d->n_hidden >= 1 && d->n_hidden <= ICRL_MAX_HIDDEN, "n_hidden %d out of range (1..%d)", d->n_hidden,
                   ICRL_MAX_HIDDEN);
    ICRL_CHECK_ARG(d->params != nullptr, "params is NULL");
    ICRL_CHECK_ARG(!d->has_norm || (d->obs_mean && d->obs_rstd), "has_norm set but obs_mean/obs_rstd NULL");
    ICRL_CHECK_ARG(!d->has_clip_acs || (d->acs_low && d->acs_high), "has_clip_acs set but acs_low/acs_high NULL");
    p->obs_dim = d->obs_dim;
    p->acs_dim = d->acs_dim;
    p->is_discrete = d->is_discrete;
    p->acs_w = d->is_discrete ? 1 : d->acs_dim;
    p->n_select = d->n_select;
    p->n_hidden = d->n_hidden;
    for (int l = 0; l < ICRL_MAX_HIDDEN; ++l) {
        p->hidden[l] = l < d->n_hidden ? d->hidden[l] : 0;
        if (l < d->n_hidden)
            ICRL_CHECK_ARG(d->hidden[l] >= 1 && d->hidden[l] <= ICRL_CN_MAX_WIDTH, "hidden[%d]=%d out of range (1..%d)", l,
                           d->hidden[l], ICRL_CN_MAX_WIDTH);
    }
    for (int i = 0; i < d->n_select; ++i) {
        ICRL_CHECK_ARG(d->select[i] >= 0 && d->select[i] < d->obs_dim + d->acs_dim, "select[%d]=%d out of range", i,
                       d->select[i]);
        p->sel[i] = d->select[i];
    }
    p->has_norm = d->has_norm;
    p->has_clip_obs = d->has_clip_obs;
    p->has_clip_acs = d->has_clip_acs && !d->is_discrete;
    p->clip_obs = d->clip_obs;
    p->params = d->params;
    p->mean = d->obs_mean;
    p->rstd = d->obs_rstd;
    p->low = d->acs_low;
    p->high = d->acs_high;
    return 0;
}

template <typename ObsT, int HP>
__global__ void __launch_bounds__(128) cn_forward_kernel(const __grid_constant__ CnPlan plan, const ObsT* __restrict__ obs,
                                                         const float* __restrict__ acs, int64_t n_rows,
                                                         float* __restrict__ out, int out_kind, int tma_ok) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int TILE = blockDim.x;
    const CnSmem L = cn_smem_layout(plan, HP, TILE, sizeof(ObsT), 0);
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + L.bar);
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        fence_mbar_init();
    }
    cn_load_weights<HP>(plan, L, smem);
    __syncthreads();

    const int64_t n_tiles = (n_rows + TILE - 1) / TILE;
    uint32_t phase = 0;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t row0 = tile * TILE;
        const int rows = (int)min((int64_t)TILE, n_rows - row0);
        cn_stage_tile<ObsT>(plan, L, smem, obs, acs, row0, rows, tma_ok && rows == TILE, phase);
        const int r = threadIdx.x;
        if (r < rows) {
            const float z = cn_forward_row<ObsT, HP>(plan, L, smem, r, TILE, false);
            const float pr = sigmoidf_ref(z);
            out[row0 + r] = out_kind == 0 ? 1.0f - pr : pr;
        }
        __syncthreads();   // everyone is done with the staged tile before it is overwritten
    }
}

// ---------------------------------------------------------------- tiny batches: one CTA per row, one thread per hidden unit
// The per-environment-step cost call hands over [n_envs, .] rows (5 in every shipped config).  With a thread per ROW those few
// threads walk the whole MLP serially (AntWall: 6 480 dependent-issue FMAs, ~35 us for 5 rows); here thread j owns hidden unit j
// and reads its weight row straight from the flat parameter vector (L2 resident), the inputs and activations of the row pass
// through shared memory.  Same fma order per unit (bias, then k ascending) and a serial output dot: bit-identical results.
template <typename ObsT>
__device__ __forceinline__ float cn_input_global(const CnPlan& p, const ObsT* __restrict__ orow, const float* __restrict__ arow, int k) {
    const int s = p.sel[k];
    if (s < p.obs_dim) {
        const ObsT o = orow[s];
        if (p.has_norm) {
            double t = ((double)o - p.mean[s]) * p.rstd[s];
            if (p.has_clip_obs) t = fmin(fmax(t, -p.clip_obs), p.clip_obs);
            return (float)t;
        }
        if (sizeof(ObsT) == 8) {
            double t = (double)o;
            if (p.has_clip_obs) t = fmin(fmax(t, -p.clip_obs), p.clip_obs);
            return (float)t;
        }
        float t = (float)o;
        if (p.has_clip_obs) t = fminf(fmaxf(t, -(float)p.clip_obs), (float)p.clip_obs);
        return t;
    }
    const int j = s - p.obs_dim;
    if (p.is_discrete) return ((int)arow[0] == j) ? 1.f : 0.f;
    float a = arow[j];
    if (p.has_clip_acs) a = fminf(fmaxf(a, p.low[j]), p.high[j]);
    return a;
}

template <typename ObsT>
__global__ void __launch_bounds__(ICRL_CN_MAX_WIDTH) cn_forward_small_kernel(const __grid_constant__ CnPlan plan,
                                                                              const ObsT* __restrict__ obs,
                                                                              const float* __restrict__ acs,
                                                                              float* __restrict__ out, int out_kind) {
    __shared__ float X[ICRL_MAX_SELECT];
    __shared__ float Hs[2][ICRL_CN_MAX_WIDTH];
    const int row = blockIdx.x, j = threadIdx.x;
    const ObsT* orow = obs + (size_t)row * plan.obs_dim;
    const float* arow = acs + (size_t)row * plan.acs_w;
    for (int k = j; k < plan.n_select; k += blockDim.x) X[k] = cn_input_global<ObsT>(plan, orow, arow, k);
    __syncthreads();
    const float* src = plan.params;
    int in_dim = plan.n_select;
    const float* in = X;
    for (int l = 0; l < plan.n_hidden; ++l) {
        const int out_dim = plan.hidden[l];
        if (j < out_dim) {
            const float* wrow = src + (size_t)j * in_dim;
            float acc = src[(size_t)out_dim * in_dim + j];
#pragma unroll 4
            for (int k = 0; k < in_dim; ++k) acc = fmaf(in[k], wrow[k], acc);
            Hs[l & 1][j] = fmaxf(acc, 0.f);
        }
        __syncthreads();
        src += (size_t)out_dim * in_dim + out_dim;
        in_dim = out_dim;
        in = Hs[l & 1];
    }
    if (j == 0) {
        float z = src[in_dim];
        for (int k = 0; k < in_dim; ++k) z = fmaf(in[k], src[k], z);
        const float pr = sigmoidf_ref(z);
        out[row] = out_kind == 0 ? 1.0f - pr : pr;
    }
}

// ---------------------------------------------------------------- packed-FP32 kernel (the default)
// Measured on B200 (tools/micro/ffma2_bench.cu, profiles/ffma2_bench_r02.txt): the thread-per-row loop above is bound by the
// RETURN BANDWIDTH OF BROADCAST LDS.128 (one weight quad per ~2.2 cycles per SM: every weight is fetched once per row) and
// by 3-register FFMA issue, at 45 % of the FP32 peak whatever the FMA flavour.  Here a thread owns 2 RP rows and keeps them as
// the two halves of 64-bit registers: one `fma.rn.f32x2` (SASS FFMA2, the weight as a broadcast scalar operand) updates one
// hidden unit of a row PAIR, so every weight read from shared memory feeds 2 RP rows and the FMA instruction count halves.
// Per input k and row pair: HP FFMA2 + HP/4 LDS.128 (shared by all RP pairs) + the two inputs.  The arithmetic per row is
// the same IEEE fma sequence as before (bias + sum over k in order): results are bit-identical to the scalar kernel.
// Tiles of 2 RP blockDim rows are staged by TMA bulk copies, double buffered when shared memory allows (the next tile's
// copy is in flight while this one is computed).  Hidden activations of layers >= 1 pass through thread-private shared
// memory slots only to avoid dynamic register indexing (no barrier involved).
__device__ __forceinline__ void ffma2_bcast(float2& acc, const float2 x, const float w) {
    unsigned long long a, xx, ww;
    asm("mov.b64 %0, {%1, %2};" : "=l"(a) : "f"(acc.x), "f"(acc.y));
    asm("mov.b64 %0, {%1, %2};" : "=l"(xx) : "f"(x.x), "f"(x.y));
    asm("mov.b64 %0, {%1, %1};" : "=l"(ww) : "f"(w));                   // folded into the FFMA2 operand by ptxas
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(a) : "l"(xx), "l"(ww));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(acc.x), "=f"(acc.y) : "l"(a));
}

struct CnPairSmem {
    int bar, mean, rstd, low, high, w[ICRL_MAX_HIDDEN], b[ICRL_MAX_HIDDEN], wout, stage, stage_bytes, acs_off, h, total;
};
__host__ __device__ inline CnPairSmem cn_pair_layout(const CnPlan& p, int HP, int NT, int RP, int obs_elem, int nbuf) {
    CnPairSmem s;
    const int TILE = 2 * RP * NT;
    int off = 0;
    s.bar = off; off += 16;
    s.mean = off; off += p.has_norm ? p.obs_dim * 8 : 0;
    s.rstd = off; off += p.has_norm ? p.obs_dim * 8 : 0;
    s.low = off; off += p.has_clip_acs ? p.acs_dim * 4 : 0;
    s.high = off; off += p.has_clip_acs ? p.acs_dim * 4 : 0;
    off = align_up(off, 16);
    for (int l = 0; l < ICRL_MAX_HIDDEN; ++l) {
        s.w[l] = off;
        if (l < p.n_hidden) off += (l == 0 ? align_up(p.n_select, 8) : HP) * HP * 4;
        s.b[l] = off;
        if (l < p.n_hidden) off += HP * 4;
    }
    s.wout = off; off += (HP + 4) * 4;
    off = align_up(off, 16);
    s.acs_off = align_up(TILE * p.obs_dim * obs_elem, 16);
    s.stage_bytes = s.acs_off + align_up(TILE * p.acs_w * 4, 16);
    s.stage = off; off += nbuf * s.stage_bytes;
    s.h = off; off += (p.n_hidden > 1) ? RP * HP * NT * 8 : 0;
    s.total = off;
    return s;
}

template <typename ObsT, int HP, int RP>
__global__ void __launch_bounds__(128) cn_forward_pair_kernel(const __grid_constant__ CnPlan plan, const ObsT* __restrict__ obs,
                                                              const float* __restrict__ acs, int64_t n_rows,
                                                              float* __restrict__ out, int out_kind, int tma_ok, int nbuf) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int NT = blockDim.x, TILE = 2 * RP * NT, tid = threadIdx.x;
    const CnPairSmem L = cn_pair_layout(plan, HP, NT, RP, sizeof(ObsT), nbuf);
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + L.bar);
    CnSmem V;                      // the common helpers' view: weights + the current staging buffer
    V.bar = L.bar; V.mean = L.mean; V.rstd = L.rstd; V.low = L.low; V.high = L.high; V.wout = L.wout;
    for (int l = 0; l < ICRL_MAX_HIDDEN; ++l) { V.w[l] = L.w[l]; V.b[l] = L.b[l]; }
    V.obs = L.stage; V.acs = L.stage + L.acs_off; V.h = L.h; V.total = L.total;
    if (tid == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        fence_mbar_init();
    }
    cn_load_weights<HP>(plan, V, smem);
    __syncthreads();

    bool fast_in = sizeof(ObsT) == 4 && !plan.has_norm && !plan.is_discrete && plan.n_select == plan.obs_dim + plan.acs_dim;
    for (int k = 0; fast_in && k < plan.n_select; ++k) fast_in = plan.sel[k] == k;
    const float co = (float)plan.clip_obs;

    const int64_t n_tiles = (n_rows + TILE - 1) / TILE;
    auto is_full = [&](int64_t tile) { return tile * TILE + TILE <= n_rows; };
    auto issue = [&](int64_t tile, int buf) {          // one thread: both bulk copies of a FULL tile onto bar[buf]
        const uint32_t ob = (uint32_t)TILE * plan.obs_dim * sizeof(ObsT), ab = (uint32_t)TILE * plan.acs_w * 4u;
        unsigned char* base = smem + L.stage + buf * L.stage_bytes;
        fence_proxy_async();
        mbar_expect_tx(&bar[buf], ob + ab);
        bulk_g2s(base, obs + tile * TILE * plan.obs_dim, ob, &bar[buf]);
        bulk_g2s(base + L.acs_off, acs + tile * TILE * plan.acs_w, ab, &bar[buf]);
    };
    uint32_t phases = 0;
    int it = 0;
    if (nbuf == 2 && tma_ok && tid == 0 && blockIdx.x < n_tiles && is_full(blockIdx.x)) issue(blockIdx.x, 0);
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
        const int buf = nbuf == 2 ? (it & 1) : 0;
        const int64_t row0 = tile * TILE;
        const int rows = (int)min((int64_t)TILE, n_rows - row0);
        V.obs = L.stage + buf * L.stage_bytes;
        V.acs = V.obs + L.acs_off;
        if (nbuf == 2) {
            const int64_t nxt = tile + gridDim.x;
            if (tma_ok && tid == 0 && nxt < n_tiles && is_full(nxt)) issue(nxt, buf ^ 1);
        }
        if (tma_ok && rows == TILE) {
            if (nbuf == 1 && tid == 0) issue(tile, 0);
            mbar_wait(&bar[buf], (phases >> buf) & 1u);
            phases ^= 1u << buf;
        } else {
            ObsT* so = reinterpret_cast<ObsT*>(smem + V.obs);
            float* sa = reinterpret_cast<float*>(smem + V.acs);
            const ObsT* go = obs + row0 * plan.obs_dim;
            const float* ga = acs + row0 * plan.acs_w;
            for (int i = tid; i < rows * plan.obs_dim; i += NT) so[i] = go[i];
            for (int i = tid; i < rows * plan.acs_w; i += NT) sa[i] = ga[i];
            __syncthreads();
        }
        // rows of this thread: pair p = (tid + 2p NT, tid + (2p + 1) NT).  Rows beyond `rows` read stale staging data; finite or
        // not, it only reaches their own (discarded) outputs.
        float2 acc[RP][HP];
        {
            const float* B = reinterpret_cast<const float*>(smem + V.b[0]);
#pragma unroll
            for (int j = 0; j < HP; ++j) {
                const float b = B[j];
#pragma unroll
                for (int p = 0; p < RP; ++p) acc[p][j] = make_float2(b, b);
            }
            const float* W = reinterpret_cast<const float*>(smem + V.w[0]);
            auto fma_row = [&](const float2 (&x)[RP], const float* wrow) {
                const float4* w4 = reinterpret_cast<const float4*>(wrow);
#pragma unroll
                for (int j = 0; j < HP / 4; ++j) {
                    const float4 w = w4[j];
#pragma unroll
                    for (int p = 0; p < RP; ++p) {
                        ffma2_bcast(acc[p][4 * j + 0], x[p], w.x);
                        ffma2_bcast(acc[p][4 * j + 1], x[p], w.y);
                        ffma2_bcast(acc[p][4 * j + 2], x[p], w.z);
                        ffma2_bcast(acc[p][4 * j + 3], x[p], w.w);
                    }
                }
            };
            if (fast_in) {
                // every dimension selected in order, float32 observations, no normalisation, continuous actions (the
                // HalfCheetah / Ant command lines): clip and go, no per-input indirection
                const float* so = reinterpret_cast<const float*>(smem + V.obs);
                const float* sa = reinterpret_cast<const float*>(smem + V.acs);
                const float* lo = reinterpret_cast<const float*>(smem + V.low);
                const float* hi = reinterpret_cast<const float*>(smem + V.high);
#pragma unroll 2
                for (int k = 0; k < plan.obs_dim; ++k) {
                    float2 x[RP];
#pragma unroll
                    for (int p = 0; p < RP; ++p) {
                        x[p].x = so[(tid + 2 * p * NT) * plan.obs_dim + k];
                        x[p].y = so[(tid + (2 * p + 1) * NT) * plan.obs_dim + k];
                        if (plan.has_clip_obs) {
                            x[p].x = fminf(fmaxf(x[p].x, -co), co);
                            x[p].y = fminf(fmaxf(x[p].y, -co), co);
                        }
                    }
                    fma_row(x, W + k * HP);
                }
#pragma unroll 2
                for (int k = 0; k < plan.acs_dim; ++k) {
                    float2 x[RP];
#pragma unroll
                    for (int p = 0; p < RP; ++p) {
                        x[p].x = sa[(tid + 2 * p * NT) * plan.acs_dim + k];
                        x[p].y = sa[(tid + (2 * p + 1) * NT) * plan.acs_dim + k];
                        if (plan.has_clip_acs) {
                            x[p].x = fminf(fmaxf(x[p].x, lo[k]), hi[k]);
                            x[p].y = fminf(fmaxf(x[p].y, lo[k]), hi[k]);
                        }
                    }
                    fma_row(x, W + (plan.obs_dim + k) * HP);
                }
            } else {
#pragma unroll 2
                for (int k = 0; k < plan.n_select; ++k) {
                    float2 x[RP];
#pragma unroll
                    for (int p = 0; p < RP; ++p) {
                        x[p].x = cn_input<ObsT>(plan, V, smem, tid + 2 * p * NT, k);
                        x[p].y = cn_input<ObsT>(plan, V, smem, tid + (2 * p + 1) * NT, k);
                    }
                    fma_row(x, W + k * HP);
                }
            }
            float2* Hs = reinterpret_cast<float2*>(smem + L.h);
            for (int l = 1; l < plan.n_hidden; ++l) {
#pragma unroll
                for (int j = 0; j < HP; ++j)
#pragma unroll
                    for (int p = 0; p < RP; ++p)
                        Hs[(p * HP + j) * NT + tid] = make_float2(fmaxf(acc[p][j].x, 0.f), fmaxf(acc[p][j].y, 0.f));
                const float* Bl = reinterpret_cast<const float*>(smem + V.b[l]);
#pragma unroll
                for (int j = 0; j < HP; ++j) {
                    const float b = Bl[j];
#pragma unroll
                    for (int p = 0; p < RP; ++p) acc[p][j] = make_float2(b, b);
                }
                const float* Wl = reinterpret_cast<const float*>(smem + V.w[l]);
                const int kin = plan.hidden[l - 1];
#pragma unroll 2
                for (int k = 0; k < kin; ++k) {
                    float2 x[RP];
#pragma unroll
                    for (int p = 0; p < RP; ++p) x[p] = Hs[(p * HP + k) * NT + tid];
                    fma_row(x, Wl + k * HP);
                }
            }
        }
        const float* WO = reinterpret_cast<const float*>(smem + V.wout);
#pragma unroll
        for (int p = 0; p < RP; ++p) {
            float2 z = make_float2(WO[HP], WO[HP]);
#pragma unroll
            for (int j = 0; j < HP; ++j)
                ffma2_bcast(z, make_float2(fmaxf(acc[p][j].x, 0.f), fmaxf(acc[p][j].y, 0.f)), WO[j]);
            const int r0 = tid + 2 * p * NT, r1 = r0 + NT;
            if (r0 < rows) {
                const float pr = sigmoidf_ref(z.x);
                out[row0 + r0] = out_kind == 0 ? 1.0f - pr : pr;
            }
            if (r1 < rows) {
                const float pr = sigmoidf_ref(z.y);
                out[row0 + r1] = out_kind == 0 ? 1.0f - pr : pr;
            }
        }
        __syncthreads();   // everyone is done with this staging buffer before a later copy overwrites it
    }
}

// Tensor-core variant (one or two hidden layers): a warp owns 32 rows of the 128-row tile (two m16 tiles) and runs every
// layer as 3xTF32 mma.sync tiles.  Inputs are prepared once (select / normalise / clip in cn_input, thread = row so that
// the per-input branches stay warp-uniform) into a padded shared-memory tile that feeds the A fragments.  Layer 1
// consumes the layer-0 accumulators from registers: with the K index permuted (k' = 2t -> slot t, 2t+1 -> slot t+4, the same
// permutation on the weight rows) a C fragment is an A fragment.  The output layer is an fp32 dot over the C fragments
// plus a 4-lane shuffle reduction.  ~18 warp-instructions per row instead of 40: the HalfCheetah shape moves from
// FFMA-issue bound towards the HBM bound, the Ant shape from FP32 to tensor throughput.
template <typename ObsT, int HP>
__global__ void __launch_bounds__(128) cn_forward_mma_kernel(const __grid_constant__ CnPlan plan, const ObsT* __restrict__ obs,
                                                             const float* __restrict__ acs, int64_t n_rows,
                                                             float* __restrict__ out, int out_kind, int tma_ok) {
    extern __shared__ __align__(16) unsigned char smem[];
    constexpr int TILE = 128, NTH = HP / 8;
    const int k0pad = align_up(plan.n_select, 8);
    const int KLD = k0pad + 4;                       // prepared-input row stride: == 4 (mod 8) -> conflict-free fragment loads
    const CnSmem L = cn_smem_layout(plan, HP, TILE, sizeof(ObsT), -1);
    float* XS = reinterpret_cast<float*>(smem + L.total);      // [TILE][KLD] prepared inputs, appended to the common layout
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + L.bar);
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        fence_mbar_init();
    }
    cn_load_weights<HP>(plan, L, smem);
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const float* W0 = reinterpret_cast<const float*>(smem + L.w[0]);
    const float* B0 = reinterpret_cast<const float*>(smem + L.b[0]);
    const float* W1 = reinterpret_cast<const float*>(smem + L.w[1]);
    const float* B1 = reinterpret_cast<const float*>(smem + L.b[1]);
    const float* WO = reinterpret_cast<const float*>(smem + L.wout);

    bool fast_prep = sizeof(ObsT) == 4 && !plan.has_norm && !plan.is_discrete && plan.n_select == plan.obs_dim + plan.acs_dim;
    for (int k = 0; fast_prep && k < plan.n_select; ++k) fast_prep = plan.sel[k] == k;

    const int64_t n_tiles = (n_rows + TILE - 1) / TILE;
    uint32_t phase = 0;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t row0 = tile * TILE;
        const int rows = (int)min((int64_t)TILE, n_rows - row0);
        cn_stage_tile<ObsT>(plan, L, smem, obs, acs, row0, rows, tma_ok && rows == TILE, phase);
        float acc[2][NTH][4];
#pragma unroll
        for (int m = 0; m < 2; ++m)
#pragma unroll
            for (int nt = 0; nt < NTH; ++nt) {
                const float2 b = *reinterpret_cast<const float2*>(B0 + 8 * nt + 2 * t);
                acc[m][nt][0] = b.x; acc[m][nt][1] = b.y; acc[m][nt][2] = b.x; acc[m][nt][3] = b.y;
            }
        // ---- input preparation: thread = row, k uniform across the warp (uniform branches and constant-bank reads in
        // cn_input).  A warp prepares exactly the 32 rows it consumes, so a warp-level barrier is enough.  Rows beyond `rows`
        // hold stale staging data: finite or not, they only reach their own (discarded) output rows.
        {
            float* xr = XS + threadIdx.x * KLD;
            if (fast_prep) {
                // all dimensions selected in order, float32 observations, no normalisation, continuous actions (the
                // HalfCheetah / Ant command lines): clip and copy, no per-input indirection
                const float* so = reinterpret_cast<const float*>(smem + L.obs) + threadIdx.x * plan.obs_dim;
                const float* sa = reinterpret_cast<const float*>(smem + L.acs) + threadIdx.x * plan.acs_dim;
                const float co = (float)plan.clip_obs;
                if (plan.has_clip_obs) {
#pragma unroll 4
                    for (int k = 0; k < plan.obs_dim; ++k) xr[k] = fminf(fmaxf(so[k], -co), co);
                } else {
#pragma unroll 4
                    for (int k = 0; k < plan.obs_dim; ++k) xr[k] = so[k];
                }
                if (plan.has_clip_acs) {
                    const float* lo = reinterpret_cast<const float*>(smem + L.low);
                    const float* hi = reinterpret_cast<const float*>(smem + L.high);
#pragma unroll 2
                    for (int j = 0; j < plan.acs_dim; ++j) xr[plan.obs_dim + j] = fminf(fmaxf(sa[j], lo[j]), hi[j]);
                } else {
#pragma unroll 2
                    for (int j = 0; j < plan.acs_dim; ++j) xr[plan.obs_dim + j] = sa[j];
                }
            } else {
                for (int k = 0; k < plan.n_select; ++k) xr[k] = cn_input<ObsT>(plan, L, smem, threadIdx.x, k);
            }
            for (int k = plan.n_select; k < k0pad; ++k) xr[k] = 0.f;
        }
        __syncwarp();
        // ---- layer 0
        for (int k0 = 0; k0 < k0pad; k0 += 8) {
            uint32_t ahi[2][4], alo[2][4];
#pragma unroll
            for (int m = 0; m < 2; ++m) {
                const float* xa = XS + (32 * warp + 16 * m + g) * KLD + k0 + t;
                split_tf32(xa[0], ahi[m][0], alo[m][0]);
                split_tf32(xa[8 * KLD], ahi[m][1], alo[m][1]);
                split_tf32(xa[4], ahi[m][2], alo[m][2]);
                split_tf32(xa[8 * KLD + 4], ahi[m][3], alo[m][3]);
            }
#pragma unroll
            for (int nt = 0; nt < NTH; ++nt) {
                uint32_t bhi[2], blo[2];
                split_tf32(W0[(k0 + t) * HP + 8 * nt + g], bhi[0], blo[0]);
                split_tf32(W0[(k0 + t + 4) * HP + 8 * nt + g], bhi[1], blo[1]);
#pragma unroll
                for (int m = 0; m < 2; ++m) {
                    mma_tf32(acc[m][nt], alo[m], bhi);
                    mma_tf32(acc[m][nt], ahi[m], blo);
                    mma_tf32(acc[m][nt], ahi[m], bhi);
                }
            }
        }
        // ---- layer 1 (when present): the ReLU'd layer-0 tiles are the A fragments (permuted K), weights read with the same
        // row permutation
        if (plan.n_hidden == 2) {
            float acc2[2][NTH][4];
#pragma unroll
            for (int m = 0; m < 2; ++m)
#pragma unroll
                for (int nt = 0; nt < NTH; ++nt) {
                    const float2 b = *reinterpret_cast<const float2*>(B1 + 8 * nt + 2 * t);
                    acc2[m][nt][0] = b.x; acc2[m][nt][1] = b.y; acc2[m][nt][2] = b.x; acc2[m][nt][3] = b.y;
                }
#pragma unroll
            for (int kt = 0; kt < NTH; ++kt) {
                uint32_t ahi[2][4], alo[2][4];
#pragma unroll
                for (int m = 0; m < 2; ++m) {
                    split_tf32(fmaxf(acc[m][kt][0], 0.f), ahi[m][0], alo[m][0]);
                    split_tf32(fmaxf(acc[m][kt][2], 0.f), ahi[m][1], alo[m][1]);
                    split_tf32(fmaxf(acc[m][kt][1], 0.f), ahi[m][2], alo[m][2]);
                    split_tf32(fmaxf(acc[m][kt][3], 0.f), ahi[m][3], alo[m][3]);
                }
#pragma unroll
                for (int nt = 0; nt < NTH; ++nt) {
                    uint32_t bhi[2], blo[2];
                    split_tf32(W1[(8 * kt + 2 * t) * HP + 8 * nt + g], bhi[0], blo[0]);
                    split_tf32(W1[(8 * kt + 2 * t + 1) * HP + 8 * nt + g], bhi[1], blo[1]);
#pragma unroll
                    for (int m = 0; m < 2; ++m) {
                        mma_tf32(acc2[m][nt], alo[m], bhi);
                        mma_tf32(acc2[m][nt], ahi[m], blo);
                        mma_tf32(acc2[m][nt], ahi[m], bhi);
                    }
                }
            }
#pragma unroll
            for (int m = 0; m < 2; ++m)
#pragma unroll
                for (int nt = 0; nt < NTH; ++nt)
#pragma unroll
                    for (int c = 0; c < 4; ++c) acc[m][nt][c] = acc2[m][nt][c];
        }
        // ---- output layer: z = b_out + sum_j relu(h_j) w_j; each thread holds columns 8 nt + 2t (+1) of rows g and g + 8
#pragma unroll
        for (int m = 0; m < 2; ++m) {
            float za = 0.f, zb = 0.f;
#pragma unroll
            for (int nt = 0; nt < NTH; ++nt) {
                const float2 w = *reinterpret_cast<const float2*>(WO + 8 * nt + 2 * t);
                za = fmaf(fmaxf(acc[m][nt][0], 0.f), w.x, za);
                za = fmaf(fmaxf(acc[m][nt][1], 0.f), w.y, za);
                zb = fmaf(fmaxf(acc[m][nt][2], 0.f), w.x, zb);
                zb = fmaf(fmaxf(acc[m][nt][3], 0.f), w.y, zb);
            }
            za += __shfl_xor_sync(0xffffffffu, za, 1); za += __shfl_xor_sync(0xffffffffu, za, 2);
            zb += __shfl_xor_sync(0xffffffffu, zb, 1); zb += __shfl_xor_sync(0xffffffffu, zb, 2);
            if (t == 0) {
                const int r = 32 * warp + 16 * m + g;
                if (r < rows) {
                    const float pr = sigmoidf_ref(za + WO[HP]);
                    out[row0 + r] = out_kind == 0 ? 1.0f - pr : pr;
                }
                if (r + 8 < rows) {
                    const float pr = sigmoidf_ref(zb + WO[HP]);
                    out[row0 + r + 8] = out_kind == 0 ? 1.0f - pr : pr;
                }
            }
        }
        __syncthreads();   // everyone is done with the staged tile before it is overwritten
    }
}

static int env_int(const char* name, int dflt) {
    const char* v = getenv(name);
    return v ? atoi(v) : dflt;
}

template <typename ObsT, int HP, int RP>
static int launch_pair_cfg(const CnPlan& plan, const void* obs, const float* acs, int64_t n_rows, float* out, int out_kind,
                           cudaStream_t st, int nt, int nbuf) {
    auto kern = cn_forward_pair_kernel<ObsT, HP, RP>;
    const CnPairSmem L = cn_pair_layout(plan, HP, nt, RP, sizeof(ObsT), nbuf);
    ICRL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L.total));
    int per_sm = 1;
    ICRL_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, nt, L.total));
    if (per_sm < 1) per_sm = 1;
    const int tile = 2 * RP * nt;
    const int64_t n_tiles = (n_rows + tile - 1) / tile;
    const int grid = (int)((n_tiles < (int64_t)per_sm * sm_count()) ? n_tiles : (int64_t)per_sm * sm_count());
    const int tma_ok = g_k1_allow_tma && ((reinterpret_cast<uintptr_t>(obs) | reinterpret_cast<uintptr_t>(acs)) & 15u) == 0 &&
                       (tile * plan.obs_dim * sizeof(ObsT)) % 16 == 0 && (tile * plan.acs_w * 4) % 16 == 0;
    kern<<<grid, nt, L.total, st>>>(plan, static_cast<const ObsT*>(obs), acs, n_rows, out, out_kind, tma_ok, nbuf);
    ICRL_LAUNCH_CHECK();
    return 0;
}

// Launch shape of the packed kernel: rows per thread (2 RP), threads per CTA and single / double buffering are chosen so
// that (a) small buffers (a rollout) still spread over the SMs, (b) at least two CTAs fit an SM, double buffered if possible.
template <typename ObsT, int HP>
static int launch_forward_pair(const CnPlan& plan, const void* obs, const float* acs, int64_t n_rows, float* out,
                               int out_kind, cudaStream_t st) {
    const int env_rp = env_int("ICRL_K1_RP", 0), env_nt = env_int("ICRL_K1_NT", 0), env_nbuf = env_int("ICRL_K1_NBUF", 0);
    const int sms = sm_count();
    int rp = 1;      // four rows per thread (RP 2) halves the weight reads again but measured slower: 260 us against 186 (HalfCheetah)
    if (env_rp) rp = (env_rp >= 2 && HP <= 32) ? 2 : 1;
    int nt = 128;
    while (nt > 32 && (n_rows + 2 * rp * nt - 1) / (2 * rp * nt) < 2 * (int64_t)sms) nt /= 2;
    const int budget = 113 * 1024;                   // two CTAs per SM
    int nbuf = 2;
    for (;;) {
        if (cn_pair_layout(plan, HP, nt, rp, sizeof(ObsT), 2).total <= budget) { nbuf = 2; break; }
        if (cn_pair_layout(plan, HP, nt, rp, sizeof(ObsT), 1).total <= budget) { nbuf = 1; break; }
        if (nt > 32) { nt /= 2; continue; }
        if (rp > 1) { rp = 1; continue; }
        nbuf = 1;
        break;
    }
    if (env_nt) nt = env_nt;
    if (env_nbuf) nbuf = env_nbuf;
    if (cn_pair_layout(plan, HP, nt, rp, sizeof(ObsT), nbuf).total > 227 * 1024) {
        set_error("constraint net too large for shared memory (%d bytes)", cn_pair_layout(plan, HP, nt, rp, sizeof(ObsT), nbuf).total);
        return ICRL_EUNSUPPORTED;
    }
    if constexpr (HP <= 32) {
        if (rp == 2) return launch_pair_cfg<ObsT, HP, 2>(plan, obs, acs, n_rows, out, out_kind, st, nt, nbuf);
    }
    return launch_pair_cfg<ObsT, HP, 1>(plan, obs, acs, n_rows, out, out_kind, st, nt, nbuf);
}

template <typename ObsT, int HP>
static int launch_forward(const CnPlan& plan, const void* obs, const float* acs, int64_t n_rows, float* out, int out_kind,
                          cudaStream_t st) {
    const bool force_ffma = getenv("ICRL_K1_FFMA") != nullptr;       // A/B switches for profiling and tests (read per call)
    const bool force_mma = getenv("ICRL_K1_MMA") != nullptr;
    if (plan.n_hidden <= 2 && force_mma && !force_ffma) {
        bool use_mma = false;
        auto mk = cn_forward_mma_kernel<ObsT, HP>;
        CnSmem Lm = cn_smem_layout(plan, HP, 128, sizeof(ObsT), -1);
        Lm.total += 128 * (align_up(plan.n_select, 8) + 4) * 4;          // + the prepared-input tile
        if (Lm.total <= 220 * 1024) {
            ICRL_CUDA(cudaFuncSetAttribute(mk, cudaFuncAttributeMaxDynamicSharedMemorySize, Lm.total));
            int per_sm = 1;
            ICRL_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, mk, 128, Lm.total));
            if (per_sm < 1) per_sm = 1;
            // Measured on B200 (4.19 M rows): HalfCheetah 202-208 us against 209 us for the FFMA kernel, Ant 2.76 ms against
            // 2.26 ms, LapGrid 111 us against 66 us.  ncu: 126 M warp instructions against 169 M, issue slots 65 % busy,
            // tensor pipe 29 % -- the MLP arithmetic is no longer what bounds this kernel (tile staging, preparation and
            // integer address work are), so the FFMA kernel stays the default and this one is opt-in (ICRL_K1_MMA=1).
            use_mma = force_mma;
            (void)per_sm;
        }
        if (use_mma) {
            int per_sm = 1;
            ICRL_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, mk, 128, Lm.total));
            if (per_sm < 1) per_sm = 1;
            const int64_t n_tiles = (n_rows + 127) / 128;
            const int grid = (int)((n_tiles < (int64_t)per_sm * sm_count()) ? n_tiles : (int64_t)per_sm * sm_count());
            const int tma_ok = g_k1_allow_tma && ((reinterpret_cast<uintptr_t>(obs) | reinterpret_cast<uintptr_t>(acs)) & 15u) == 0;
            mk<<<grid, 128, Lm.total, st>>>(plan, static_cast<const ObsT*>(obs), acs, n_rows, out, out_kind, tma_ok);
            ICRL_LAUNCH_CHECK();
            return 0;
        }
    }
    const bool force_v1 = getenv("ICRL_K1_V1") != nullptr;             // the scalar thread-per-row kernel (A/B runs)
    // packed kernel where it wins (measured on B200, 4.19 M rows: HalfCheetah 186 us against 203, PointCircle 324 against 516,
    // LapGrid 59 against 60); the wide Ant net (HP 40, 484 B of staging per row) keeps the scalar kernel (2.25 ms against 2.62)
    const bool force_pair = getenv("ICRL_K1_PAIR") != nullptr;
    // (rollout-sized buffers -- fewer rows than 128 per SM -- keep the scalar kernel too: 12.3 us against 14.3 for 10 240 rows)
    if (!force_v1 && !force_ffma && ((HP <= 32 && n_rows >= (int64_t)128 * sm_count()) || force_pair))
        return launch_forward_pair<ObsT, HP>(plan, obs, acs, n_rows, out, out_kind, st);
    auto kern = cn_forward_kernel<ObsT, HP>;
    // pick the largest tile (== block size) whose shared memory fits
    int tile = 128;
    CnSmem L = cn_smem_layout(plan, HP, tile, sizeof(ObsT), 0);
    while (L.total > 220 * 1024 && tile > 32) {
        tile /= 2;
        L = cn_smem_layout(plan, HP, tile, sizeof(ObsT), 0);
    }
    if (L.total > 220 * 1024) {
        set_error("constraint net too large for shared memory (%d bytes)", L.total);
        return ICRL_EUNSUPPORTED;
    }
    ICRL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L.total));
    int per_sm = 1;
    ICRL_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, tile, L.total));
    if (per_sm < 1) per_sm = 1;
    const int64_t n_tiles = (n_rows + tile - 1) / tile;
    const int grid = (int)((n_tiles < (int64_t)per_sm * sm_count()) ? n_tiles : (int64_t)per_sm * sm_count());
    const int tma_ok = g_k1_allow_tma && ((reinterpret_cast<uintptr_t>(obs) | reinterpret_cast<uintptr_t>(acs)) & 15u) == 0;
    kern<<<grid, tile, L.total, st>>>(plan, static_cast<const ObsT*>(obs), acs, n_rows, out, out_kind, tma_ok);
    ICRL_LAUNCH_CHECK();
    return 0;
}

template <typename ObsT>
static int dispatch_width(const CnPlan& plan, const void* obs, const float* acs, int64_t n, float* out, int kind,
                          cudaStream_t st) {
    switch (cn_padded_width(plan)) {
        case 8: return launch_forward<ObsT, 8>(plan, obs, acs, n, out, kind, st);
        case 16: return launch_forward<ObsT, 16>(plan, obs, acs, n, out, kind, st);
        case 24: return launch_forward<ObsT, 24>(plan, obs, acs, n, out, kind, st);
        case 32: return launch_forward<ObsT, 32>(plan, obs, acs, n, out, kind, st);
        case 40: return launch_forward<ObsT, 40>(plan, obs, acs, n, out, kind, st);
        case 48: return launch_forward<ObsT, 48>(plan, obs, acs, n, out, kind, st);
        case 64: return launch_forward<ObsT, 64>(plan, obs, acs, n, out, kind, st);
    }
    set_error("unsupported constraint-net width");
    return ICRL_EUNSUPPORTED;
}

int cn_forward_device(const CnPlan& plan, const void* obs, int obs_is_f64, const float* acs, int64_t n_rows, float* out,
                      int out_kind, cudaStream_t st) {
    if (n_rows == 0) return 0;
    if (n_rows <= 16 && getenv("ICRL_K1_NO_SMALL") == nullptr) {          // a cost-wrapper call: one CTA per row, thread per unit
        int width = 32;
        for (int l = 0; l < plan.n_hidden; ++l) width = plan.hidden[l] > width ? plan.hidden[l] : width;
        width = (width + 31) / 32 * 32;
        if (obs_is_f64)
            cn_forward_small_kernel<double><<<(int)n_rows, width, 0, st>>>(plan, static_cast<const double*>(obs), acs, out, out_kind);
        else
            cn_forward_small_kernel<float><<<(int)n_rows, width, 0, st>>>(plan, static_cast<const float*>(obs), acs, out, out_kind);
        ICRL_LAUNCH_CHECK();
        return 0;
    }
    return obs_is_f64 ? dispatch_width<double>(plan, obs, acs, n_rows, out, out_kind, st)
                      : dispatch_width<float>(plan, obs, acs, n_rows, out, out_kind, st);
}

}  // namespace icrl

extern "C" {

int64_t icrl_cn_param_count(const icrl_cn_desc* d) {
    icrl::CnPlan p;
    if (icrl::make_plan(d, &p) != 0) return -1;
    return icrl::cn_param_count(p);
}

int icrl_cn_forward(const icrl_cn_desc* d, const void* obs, int32_t obs_is_f64, const float* acs, int64_t n_rows,
                    float* out, int32_t out_kind, void* stream) {
    icrl::CnPlan p;
    int rc = icrl::make_plan(d, &p);
    if (rc) return rc;
    ICRL_CHECK_ARG(n_rows >= 0, "n_rows < 0");
    ICRL_CHECK_ARG(n_rows == 0 || (obs && acs && out), "NULL data pointer");
    ICRL_CHECK_ARG(out_kind == 0 || out_kind == 1, "out_kind must be 0 (cost) or 1 (prediction)");
    return icrl::cn_forward_device(p, obs, obs_is_f64, acs, n_rows, out, out_kind, (cudaStream_t)stream);
}

int icrl_cn_forward_host(const icrl_cn_desc* d, const void* obs, int32_t obs_is_f64, const float* acs, int64_t n_rows,
                         float* out, int32_t out_kind, void* stream) {
    icrl::CnPlan p;
    int rc = icrl::make_plan(d, &p);
    if (rc) return rc;
    ICRL_CHECK_ARG(n_rows >= 0, "n_rows < 0");
    if (n_rows == 0) return 0;
    ICRL_CHECK_ARG(obs && acs && out, "NULL data pointer");
    ICRL_CHECK_ARG(out_kind == 0 || out_kind == 1, "out_kind must be 0 (cost) or 1 (prediction)");
    cudaStream_t st = (cudaStream_t)stream;
    const size_t ob = (size_t)n_rows * p.obs_dim * (obs_is_f64 ? 8 : 4), ab = (size_t)n_rows * p.acs_w * 4,
                 cb = (size_t)n_rows * 4;
    if (ob + ab + cb <= 64 * 1024 && getenv("ICRL_K1_NO_ZEROCOPY") == nullptr) {
        // Small batches -- the per-environment-step calls of VecCostWrapper.step_wait ([n_envs, .] rows, 2048 per rollout): three
        // staged copies + a synchronisation cost ~45 us per call on the host.  Instead the kernel reads the rows from, and
        // writes the costs to, one host-mapped pinned block (UVA: the device addresses it directly over PCIe): one launch, one
        // synchronisation.
        const size_t o_acs = (ob + 15) / 16 * 16, o_out = o_acs + (ab + 15) / 16 * 16;
        void* pin;
        if ((rc = icrl::pinned_scratch(icrl::SLOT_IN0, o_out + cb, &pin))) return rc;
        unsigned char* base = static_cast<unsigned char*>(pin);
        memcpy(base, obs, ob);
        memcpy(base + o_acs, acs, ab);
        icrl::g_k1_allow_tma = false;
        rc = icrl::cn_forward_device(p, base, obs_is_f64, reinterpret_cast<const float*>(base + o_acs), n_rows,
                                     reinterpret_cast<float*>(base + o_out), out_kind, st);
        icrl::g_k1_allow_tma = true;
        if (rc) return rc;
        ICRL_CUDA(cudaStreamSynchronize(st));
        memcpy(out, base + o_out, cb);
        return 0;
    }
    void *dobs, *dacs, *dout;
    if ((rc = icrl::device_scratch(icrl::SLOT_IN0, ob, &dobs))) return rc;
    if ((rc = icrl::device_scratch(icrl::SLOT_IN1, ab, &dacs))) return rc;
    if ((rc = icrl::device_scratch(icrl::SLOT_OUT0, cb, &dout))) return rc;
    ICRL_CUDA(cudaMemcpyAsync(dobs, obs, ob, cudaMemcpyHostToDevice, st));
    ICRL_CUDA(cudaMemcpyAsync(dacs, acs, ab, cudaMemcpyHostToDevice, st));
    rc = icrl::cn_forward_device(p, dobs, obs_is_f64, (const float*)dacs, n_rows, (float*)dout, out_kind, st);
    if (rc) return rc;
    ICRL_CUDA(cudaMemcpyAsync(out, dout, cb, cudaMemcpyDeviceToHost, st));
    ICRL_CUDA(cudaStreamSynchronize(st));
    return 0;
}

}  // extern "C"
