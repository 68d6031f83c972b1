// K1 -- fused constraint-net forward over a whole buffer of transitions (cost relabel).
// Replaces ConstraintNet.cost_function -> prepare_data -> nn.Sequential (icrl/constraint_net.py:121-130,
// 258-299, 101-119), called per env step by VecCostWrapper.step_wait (vec_cost_wrapper.py:51-66).
//
// Roofline: reads len(select_dim)*4 B and writes 4 B per transition (Ant: 488 B); layer 0 costs
// 2*n_select*h1 flops per row, so LGW/HC shapes are HBM-bound and Ant (121->40->40->1, 27 flop/B) is
// FP32-FMA-bound on CUDA cores.  Persistent grid: (resident CTAs per SM) x (SM count) CTAs loop over tiles.
#include <stdlib.h>

#include "cn_common.cuh"

namespace icrl {

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

template <typename ObsT, int HP>
static int launch_forward(const CnPlan& plan, const void* obs, const float* acs, int64_t n_rows, float* out, int out_kind,
                          cudaStream_t st) {
    static const bool force_ffma = getenv("ICRL_K1_FFMA") != nullptr;       // A/B switches for profiling
    static const bool force_mma = getenv("ICRL_K1_MMA") != nullptr;
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
            const int tma_ok = ((reinterpret_cast<uintptr_t>(obs) | reinterpret_cast<uintptr_t>(acs)) & 15u) == 0;
            mk<<<grid, 128, Lm.total, st>>>(plan, static_cast<const ObsT*>(obs), acs, n_rows, out, out_kind, tma_ok);
            ICRL_LAUNCH_CHECK();
            return 0;
        }
    }
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
    const int tma_ok = ((reinterpret_cast<uintptr_t>(obs) | reinterpret_cast<uintptr_t>(acs)) & 15u) == 0;
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
