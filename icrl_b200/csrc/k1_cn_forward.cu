// K1 -- fused constraint-net forward over a whole buffer of transitions (cost relabel).
// Replaces ConstraintNet.cost_function -> prepare_data -> nn.Sequential (icrl/constraint_net.py:121-130,
// 258-299, 101-119), called per env step by VecCostWrapper.step_wait (vec_cost_wrapper.py:51-66).
//
// Roofline: reads len(select_dim)*4 B and writes 4 B per transition (Ant: 488 B); layer 0 costs
// 2*n_select*h1 flops per row, so LGW/HC shapes are HBM-bound and Ant (121->40->40->1, 27 flop/B) is
// FP32-FMA-bound on CUDA cores.  Persistent grid: (resident CTAs per SM) x (SM count) CTAs loop over tiles.
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

template <typename ObsT, int HP>
static int launch_forward(const CnPlan& plan, const void* obs, const float* acs, int64_t n_rows, float* out, int out_kind,
                          cudaStream_t st) {
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
