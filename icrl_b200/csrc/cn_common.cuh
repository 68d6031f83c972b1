// Constraint-net building blocks shared by K1 (forward / cost relabel) and K2 (train).
//
// Work decomposition: one thread per row of a TILE-row tile.  The raw obs / acs rows of a tile are one
// contiguous chunk of global memory each, so a full tile is staged with two 1-D bulk async copies (TMA,
// cp.async.bulk -> UBLKCP) into shared memory and completion is signalled on an mbarrier; ragged tails and
// unaligned bases fall back to a cooperative coalesced copy.  Row r of the tile lives at stride obs_dim
// words (conflict-free whenever obs_dim is odd; at worst a small-way conflict on the x read that feeds HP
// FMAs).  Input preparation (normalise in float64, clip, one-hot, select; constraint_net.py:258-299) happens
// on the way from shared memory into the register that feeds layer 0.  All weights sit in shared memory,
// transposed to k-major [K][HP] so that the HP outputs of one input k are read with broadcast LDS.128.
#pragma once
#include "common.cuh"

namespace icrl {

struct CnPlan {
    int obs_dim, acs_dim, acs_w, is_discrete;
    int n_select, n_hidden, hidden[ICRL_MAX_HIDDEN];
    int has_norm, has_clip_obs, has_clip_acs;
    double clip_obs;
    const float* params;
    const double* mean;
    const double* rstd;
    const float* low;
    const float* high;
    int sel[ICRL_MAX_SELECT];
};

int make_plan(const icrl_cn_desc* d, CnPlan* p);   // validates; returns 0 or ICRL_E*
int cn_padded_width(const CnPlan& p);              // HP: smallest supported width >= max(hidden)
int64_t cn_param_count(const CnPlan& p);

#ifdef __CUDACC__

// shared-memory carve-up for a (plan, HP, TILE, obs element size); identical on host and device.
struct CnSmem {
    // byte offsets from the (16-byte aligned) dynamic smem base
    int bar, mean, rstd, low, high, w[ICRL_MAX_HIDDEN], b[ICRL_MAX_HIDDEN], wout, obs, acs, h, total;
};

__host__ __device__ inline int align_up(int x, int a) { return (x + a - 1) / a * a; }

__host__ __device__ inline CnSmem cn_smem_layout(const CnPlan& p, int HP, int TILE, int obs_elem, int extra_h_planes) {
    CnSmem s;
    int off = 0;
    s.bar = off; off += 16;
    s.mean = off; off += p.has_norm ? p.obs_dim * 8 : 0;
    s.rstd = off; off += p.has_norm ? p.obs_dim * 8 : 0;
    s.low = off; off += p.has_clip_acs ? p.acs_dim * 4 : 0;
    s.high = off; off += p.has_clip_acs ? p.acs_dim * 4 : 0;
    off = align_up(off, 16);
    for (int l = 0; l < ICRL_MAX_HIDDEN; ++l) {
        s.w[l] = off;
        if (l < p.n_hidden) off += (l == 0 ? align_up(p.n_select, 8) : HP) * HP * 4;   // zero rows pad layer 0 to whole k-steps
        s.b[l] = off;
        if (l < p.n_hidden) off += HP * 4;
    }
    s.wout = off; off += (HP + 4) * 4;                  // w_out[HP], b_out
    off = align_up(off, 16);
    s.obs = off; off += align_up(TILE * p.obs_dim * obs_elem, 16);
    s.acs = off; off += align_up(TILE * p.acs_w * 4, 16);
    s.h = off; off += (1 + extra_h_planes) * HP * TILE * 4;
    s.total = off;
    return s;
}

// Load + transpose the net's weights into shared memory (zero padded to HP).  All threads of the block.
template <int HP>
__device__ __forceinline__ void cn_load_weights(const CnPlan& p, const CnSmem& L, unsigned char* smem) {
    const float* src = p.params;
    int in_dim = p.n_select;
    for (int l = 0; l < p.n_hidden; ++l) {
        const int out_dim = p.hidden[l];
        const int kpad = (l == 0) ? align_up(p.n_select, 8) : HP;
        float* W = reinterpret_cast<float*>(smem + L.w[l]);
        float* B = reinterpret_cast<float*>(smem + L.b[l]);
        for (int i = threadIdx.x; i < kpad * HP; i += blockDim.x) {
            const int k = i / HP, j = i - k * HP;
            W[i] = (k < in_dim && j < out_dim) ? src[j * in_dim + k] : 0.f;
        }
        for (int j = threadIdx.x; j < HP; j += blockDim.x) B[j] = (j < out_dim) ? src[out_dim * in_dim + j] : 0.f;
        src += out_dim * in_dim + out_dim;
        in_dim = out_dim;
    }
    float* WO = reinterpret_cast<float*>(smem + L.wout);
    for (int j = threadIdx.x; j < HP + 1; j += blockDim.x) WO[j] = (j < in_dim) ? src[j] : (j == HP ? src[in_dim] : 0.f);
    if (p.has_norm) {
        double* M = reinterpret_cast<double*>(smem + L.mean);
        double* R = reinterpret_cast<double*>(smem + L.rstd);
        for (int i = threadIdx.x; i < p.obs_dim; i += blockDim.x) { M[i] = p.mean[i]; R[i] = p.rstd[i]; }
    }
    if (p.has_clip_acs) {
        float* lo = reinterpret_cast<float*>(smem + L.low);
        float* hi = reinterpret_cast<float*>(smem + L.high);
        for (int i = threadIdx.x; i < p.acs_dim; i += blockDim.x) { lo[i] = p.low[i]; hi[i] = p.high[i]; }
    }
}

// Stage rows [row0, row0+rows) of obs / acs into shared memory.  `use_tma` must be block-uniform.
template <typename ObsT>
__device__ __forceinline__ void cn_stage_tile(const CnPlan& p, const CnSmem& L, unsigned char* smem, const ObsT* obs,
                                              const float* acs, int64_t row0, int rows, bool use_tma,
                                              uint32_t& phase) {
    ObsT* so = reinterpret_cast<ObsT*>(smem + L.obs);
    float* sa = reinterpret_cast<float*>(smem + L.acs);
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + L.bar);
    if (use_tma) {
        const uint32_t ob = (uint32_t)rows * p.obs_dim * sizeof(ObsT), ab = (uint32_t)rows * p.acs_w * 4u;
        if (threadIdx.x == 0) {
            mbar_expect_tx(bar, ob + ab);
            bulk_g2s(so, obs + row0 * p.obs_dim, ob, bar);
            bulk_g2s(sa, acs + row0 * p.acs_w, ab, bar);
        }
        mbar_wait(bar, phase);
        phase ^= 1u;
    } else {
        const ObsT* go = obs + row0 * p.obs_dim;
        const float* ga = acs + row0 * p.acs_w;
        for (int i = threadIdx.x; i < rows * p.obs_dim; i += blockDim.x) so[i] = go[i];
        for (int i = threadIdx.x; i < rows * p.acs_w; i += blockDim.x) sa[i] = ga[i];
        __syncthreads();
    }
}

// x_k for row r of the staged tile: prepare_data() for one selected input (constraint_net.py:258-299).
template <typename ObsT>
__device__ __forceinline__ float cn_input(const CnPlan& p, const CnSmem& L, const unsigned char* smem, int r, int k) {
    const int s = p.sel[k];
    if (s < p.obs_dim) {
        const ObsT o = reinterpret_cast<const ObsT*>(smem + L.obs)[r * p.obs_dim + s];
        if (p.has_norm) {
            // float64 like numpy: (obs - mean) / sqrt(var + eps), then clip, then one rounding to float32
            double t = ((double)o - reinterpret_cast<const double*>(smem + L.mean)[s]) *
                       reinterpret_cast<const double*>(smem + L.rstd)[s];
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
    const float* sa = reinterpret_cast<const float*>(smem + L.acs);
    if (p.is_discrete) return ((int)sa[r * p.acs_w] == j) ? 1.f : 0.f;
    float a = sa[r * p.acs_w + j];
    if (p.has_clip_acs)
        a = fminf(fmaxf(a, reinterpret_cast<const float*>(smem + L.low)[j]), reinterpret_cast<const float*>(smem + L.high)[j]);
    return a;
}

// acc[0..HP) = bias + sum_k in_k * W[k][0..HP)   with W k-major in shared memory (broadcast float4 reads)
template <int HP>
__device__ __forceinline__ void cn_fma_row(float (&acc)[HP], float xv, const float* __restrict__ wrow) {
    const float4* w4 = reinterpret_cast<const float4*>(wrow);
#pragma unroll
    for (int j = 0; j < HP / 4; ++j) {
        const float4 w = w4[j];
        acc[4 * j + 0] = fmaf(xv, w.x, acc[4 * j + 0]);
        acc[4 * j + 1] = fmaf(xv, w.y, acc[4 * j + 1]);
        acc[4 * j + 2] = fmaf(xv, w.z, acc[4 * j + 2]);
        acc[4 * j + 3] = fmaf(xv, w.w, acc[4 * j + 3]);
    }
}

// Forward for the thread's row.  Post-ReLU activations of layer l are left in plane min(l, planes-1) of the h
// staging area ([plane][HP][TILE], column r) -- K2 keeps every layer's plane for the backward pass, K1 reuses
// plane 0.  Returns z (pre-sigmoid logit).
template <typename ObsT, int HP>
__device__ __forceinline__ float cn_forward_row(const CnPlan& p, const CnSmem& L, unsigned char* smem, int r, int TILE,
                                                bool keep_planes) {
    float acc[HP];
    {
        const float* B = reinterpret_cast<const float*>(smem + L.b[0]);
#pragma unroll
        for (int j = 0; j < HP; ++j) acc[j] = B[j];
        const float* W = reinterpret_cast<const float*>(smem + L.w[0]);
#pragma unroll 2
        for (int k = 0; k < p.n_select; ++k) cn_fma_row<HP>(acc, cn_input<ObsT>(p, L, smem, r, k), W + k * HP);
    }
    float* H = reinterpret_cast<float*>(smem + L.h);
    for (int l = 1; l < p.n_hidden; ++l) {
        float* Hp = H + (keep_planes ? (l - 1) * HP * TILE : 0);
#pragma unroll
        for (int j = 0; j < HP; ++j) Hp[j * TILE + r] = fmaxf(acc[j], 0.f);
        const float* B = reinterpret_cast<const float*>(smem + L.b[l]);
#pragma unroll
        for (int j = 0; j < HP; ++j) acc[j] = B[j];
        const float* W = reinterpret_cast<const float*>(smem + L.w[l]);
        const int kin = p.hidden[l - 1];
#pragma unroll 4
        for (int k = 0; k < kin; ++k) cn_fma_row<HP>(acc, Hp[k * TILE + r], W + k * HP);
    }
    const float* WO = reinterpret_cast<const float*>(smem + L.wout);
    float z = WO[HP];
    if (keep_planes) {
        float* Hp = H + (p.n_hidden - 1) * HP * TILE;
#pragma unroll
        for (int j = 0; j < HP; ++j) {
            const float a = fmaxf(acc[j], 0.f);
            Hp[j * TILE + r] = a;
            z = fmaf(a, WO[j], z);
        }
    } else {
#pragma unroll
        for (int j = 0; j < HP; ++j) z = fmaf(fmaxf(acc[j], 0.f), WO[j], z);
    }
    return z;
}

__device__ __forceinline__ float sigmoidf_ref(float z) { return 1.0f / (1.0f + expf(-z)); }

#endif  // __CUDACC__
}  // namespace icrl
