// K4 -- the whole PPOLagrangian.train() epoch/minibatch loop as ONE persistent thread-block-cluster launch.
// Replaces stable_baselines3/ppo_lag/ppo_lag.py:198-297 (+ policies.py:752-767, distributions.py, clip_grad_norm_,
// Adam): per minibatch gather -> 3 tanh MLP forward -> losses -> backward -> global-norm clip -> Adam, for every
// minibatch of every epoch, with the per-epoch target_kl early stop decided on the device.
//
// Why this shape.  The reference runs 1 600 *dependent* optimiser steps per rollout on 64-128 rows each: the path
// is latency-bound, not HBM-bound (65 KB of algorithmic traffic per step).  So:
//   * one launch; no host round trip between steps (the host only supplies numpy's permutations up front);
//   * the three trunks (pi / vf / cvf) are independent networks (torch_layers.py:129-254), so each gets its own CTA
//     of a 3-CTA cluster -- model parallel with NO activation exchange.  The only coupling is clip_grad_norm_'s
//     global norm: one float per CTA per step, exchanged through distributed shared memory + a cluster barrier;
//   * weights live in shared memory for the whole launch (k-major copies for the forward GEMMs, row-major W2 for
//     the backward), Adam moments and the gradient live in REGISTERS: every thread owns fixed 4x4 tiles of W1/W2
//     plus a few scalars for all 1 600 steps, so gradients are never materialised in memory;
//   * the small GEMMs (64 x D x 64, 64 x 64 x 64) are FP32 FFMA with 4x4 register tiles and float4 broadcast
//     shared-memory operand reads (fp32 tolerances of the north star rule out single-pass TF32/BF16 tensor cores).
// Rollout data stay in the buffer's time-major layout; env-major minibatch indices are translated here.
#include <math.h>

#include "common.cuh"

namespace icrl {

constexpr int H = 64;          // padded hidden width (both layers)
constexpr int RB = 64;         // rows per chunk
constexpr int NTH = 256;       // threads per CTA
constexpr int AMAX = 16;       // max action dims / discrete actions
constexpr int WA_LD = 68;      // leading dim of the action head weight in smem (bank-conflict-free float4 rows)
constexpr float LOG_SQRT_2PI = 0.91893853320467274178f;
constexpr float HALF_LOG_2PI_PLUS_HALF = 1.4189385332046727418f;

struct PpoArgs {
    int D, DP, A, is_discrete, h0, h1;
    int T, E, N, B, n_epochs, steps_per_epoch, max_steps;
    int has_target_kl, has_clip_vf_r, has_clip_vf_c;
    float clip_range, clip_vf_r, clip_vf_c, ent_coef, vf_coef_r, vf_coef_c, max_grad_norm, nu;
    double target_kl, lr, beta1, beta2, adam_eps;
    long long step_before;
    // flat parameter offsets (reference parameters() order)
    int off_logstd, off_w1[3], off_b1[3], off_w2[3], off_b2[3], off_hw[3], off_hb[3];
    const float *obs, *act, *old_logp, *old_vr, *adv_r, *ret_r, *old_vc, *adv_c, *ret_c;
    const int* perm;
    const int* poff;          // perm translated to time-major element offsets (prologue kernel)
    const float* advstats;    // [steps][4]: mean(adv_r), std(adv_r) (unbiased), mean(adv_c) per minibatch
    const float* nu_dev;
    float *params, *adam_m, *adam_v, *stats;
    int* result;
};

// ---------------------------------------------------------------- cluster primitives (raw PTX)
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t cluster_nctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void st_remote_f32(float* local_ptr, uint32_t rank, float v) {
    uint32_t a = smem_u32(local_ptr), ra;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(a), "r"(rank));
    asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(ra), "f"(v) : "memory");
}

__device__ __forceinline__ void cp_async_4(void* dst_smem, const void* src_gmem) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst_smem)), "l"(src_gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// ---------------------------------------------------------------- block reductions (NTH threads)
__device__ __forceinline__ float block_sum(float v, float* scratch) {
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
    __syncthreads();
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < NTH / 32; ++i) t += scratch[i];
    return t;
}
__device__ __forceinline__ double block_sum(double v, double* scratch) {
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
#pragma unroll
    for (int i = 0; i < NTH / 32; ++i) t += scratch[i];
    return t;
}

// ---------------------------------------------------------------- shared memory carve-up (float offsets)
struct PpoSmem {
    int w1t, w2t, w2, b1, b2, hw, hb, logstd, x, h1, h2, dh, rowf, dmean, act, mu, rowoff, scratch, xch, total_bytes;
};
__host__ __device__ inline PpoSmem ppo_smem_layout(int DP) {
    PpoSmem s;
    int o = 0;
    s.w1t = o; o += DP * H;
    s.w2t = o; o += H * H;
    s.w2 = o; o += H * H;
    s.b1 = o; o += H;
    s.b2 = o; o += H;
    s.hw = o; o += AMAX * WA_LD;
    s.hb = o; o += AMAX;
    s.logstd = o; o += AMAX;
    s.x = o; o += 2 * RB * DP;      // double buffered (cp.async prefetch of the next chunk)
    s.h1 = o; o += RB * H;
    s.h2 = o; o += RB * H;
    s.dh = o; o += RB * H;
    s.rowf = o; o += 2 * RB * 8;      // per-row scalars: 0 old_logp, 1 adv_r~, 2 adv_c~, 3 target return, 4 old value, 5 g/dV
    s.dmean = o; o += RB * AMAX;
    s.act = o; o += 2 * RB * AMAX;
    s.mu = o; o += RB * AMAX;        // action-head outputs (means / logits)
    s.rowoff = o; o += 3 * RB;       // int: 3-deep ring of chunk row offsets t*E+e (-1 = padding)
    s.scratch = o; o += 64;         // 32 floats / 16 doubles of reduction scratch (8-byte aligned: o is even)
    s.xch = o; o += 2 * 4 * 2;      // [parity][rank][{sumsq, stop}]
    s.total_bytes = o * 4;
    return s;
}

// 64x64 += A[64 x K] * Bt[K x 64]  (A row-major lda, Bt k-major ld 64); thread tile rows 4ty.., cols 4tx..
// KC > 0: compile-time K (fully unrolled so operand loads run ahead of the FMAs); KC == 0: runtime K.
template <int KC>
__device__ __forceinline__ void gemm_tile_4x4(float (&acc)[4][4], const float* __restrict__ A, int lda,
                                              const float* __restrict__ Bt, int K, int ty, int tx) {
    const float* a0 = A + (4 * ty) * lda;
    const float* b0 = Bt + 4 * tx;
    const int kend = KC > 0 ? KC : K;
#pragma unroll (KC > 0 ? KC / 4 : 2)
    for (int k = 0; k < kend; k += 4) {
        float4 a[4], w[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] = *reinterpret_cast<const float4*>(a0 + i * lda + k);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) w[kk] = *reinterpret_cast<const float4*>(b0 + (k + kk) * H);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float av = kk == 0 ? a[i].x : kk == 1 ? a[i].y : kk == 2 ? a[i].z : a[i].w;
                acc[i][0] = fmaf(av, w[kk].x, acc[i][0]);
                acc[i][1] = fmaf(av, w[kk].y, acc[i][1]);
                acc[i][2] = fmaf(av, w[kk].z, acc[i][2]);
                acc[i][3] = fmaf(av, w[kk].w, acc[i][3]);
            }
        }
    }
}

// acc[jj][kk] += sum_r L[r][4tj+jj] * R[r][4tk+kk]   (both row-major; reduction over the chunk's rows)
__device__ __forceinline__ void outer_tile_4x4(float (&acc)[4][4], const float* __restrict__ L, int ldl,
                                               const float* __restrict__ R, int ldr, int tj, int tk) {
    const float* l0 = L + 4 * tj;
    const float* r0 = R + 4 * tk;
#pragma unroll 8
    for (int r = 0; r < RB; ++r) {      // padding rows of the chunk hold zero gradients, so the trip count is fixed
        const float4 d = *reinterpret_cast<const float4*>(l0 + r * ldl);
        const float4 h = *reinterpret_cast<const float4*>(r0 + r * ldr);
        acc[0][0] = fmaf(d.x, h.x, acc[0][0]); acc[0][1] = fmaf(d.x, h.y, acc[0][1]);
        acc[0][2] = fmaf(d.x, h.z, acc[0][2]); acc[0][3] = fmaf(d.x, h.w, acc[0][3]);
        acc[1][0] = fmaf(d.y, h.x, acc[1][0]); acc[1][1] = fmaf(d.y, h.y, acc[1][1]);
        acc[1][2] = fmaf(d.y, h.z, acc[1][2]); acc[1][3] = fmaf(d.y, h.w, acc[1][3]);
        acc[2][0] = fmaf(d.z, h.x, acc[2][0]); acc[2][1] = fmaf(d.z, h.y, acc[2][1]);
        acc[2][2] = fmaf(d.z, h.z, acc[2][2]); acc[2][3] = fmaf(d.z, h.w, acc[2][3]);
        acc[3][0] = fmaf(d.w, h.x, acc[3][0]); acc[3][1] = fmaf(d.w, h.y, acc[3][1]);
        acc[3][2] = fmaf(d.w, h.z, acc[3][2]); acc[3][3] = fmaf(d.w, h.w, acc[3][3]);
    }
}

// one Adam update in torch's single-tensor form (torch/optim/adam.py); returns the new parameter
struct AdamConsts {
    float one_minus_b1, b2, one_minus_b2, inv_bc2_sqrt, eps, neg_step_size;
};
__device__ __forceinline__ float adam_update(float p, float g, float& m, float& v, const AdamConsts& c) {
    m = fmaf(c.one_minus_b1, g - m, m);                        // exp_avg.lerp_(grad, 1 - beta1)
    v = fmaf(c.one_minus_b2 * g, g, v * c.b2);                 // exp_avg_sq.mul_(beta2).addcmul_(g, g, 1 - beta2)
    // (sqrt(v) / sqrt(bc2)).add_(eps); param.addcdiv_(m, denom, value=-step_size).  The two divisions are done as a
    // multiply by the precomputed reciprocal and a 2-ulp fast division: <= 3e-7 relative on an lr-sized update.
    const float denom = fmaf(sqrtf(v), c.inv_bc2_sqrt, c.eps);
    return fmaf(c.neg_step_size, __fdividef(m, denom), p);
}

// ---------------------------------------------------------------- prologue: one CTA per optimiser step
// (i) numpy's env-major minibatch indices (row = e*T + t, buffers.py:52-65) -> time-major element offsets t*E + e,
// (ii) the minibatch statistics of ppo_lag.py:218-222: mean and unbiased std of the reward advantages, mean of the cost
// advantages (float64 accumulation like ATen's CPU reductions).  They depend only on data, never on parameters, so all
// 1 600 of them are computed in parallel here instead of inside the dependent step chain.
__global__ void __launch_bounds__(128) ppo_prologue_kernel(const int* __restrict__ perm, int* __restrict__ poff,
                                                           float* __restrict__ advstats, const float* __restrict__ adv_r,
                                                           const float* __restrict__ adv_c, int T, int E, int N, int B,
                                                           int spe) {
    __shared__ double red[2][4];
    const int step = blockIdx.x, epoch = step / spe, mb = step - epoch * spe;
    const int Bn = min(B, N - mb * B);
    const size_t base = (size_t)epoch * N + (size_t)mb * B;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double sr = 0.0, sc = 0.0;
    for (int i = tid; i < Bn; i += blockDim.x) {
        const int row = perm[base + i], t = row % T, e = row / T, o = t * E + e;
        poff[base + i] = o;
        sr += (double)adv_r[o];
        sc += (double)adv_c[o];
    }
    sr = warp_sum(sr); sc = warp_sum(sc);
    if (lane == 0) { red[0][warp] = sr; red[1][warp] = sc; }
    __syncthreads();
    const double mr = (red[0][0] + red[0][1] + red[0][2] + red[0][3]) / Bn;
    const double mc = (red[1][0] + red[1][1] + red[1][2] + red[1][3]) / Bn;
    __syncthreads();
    double ssq = 0.0;
    for (int i = tid; i < Bn; i += blockDim.x) {
        const int row = perm[base + i], t = row % T, e = row / T;
        const double dv = (double)adv_r[t * E + e] - mr;
        ssq += dv * dv;
    }
    ssq = warp_sum(ssq);
    if (lane == 0) red[0][warp] = ssq;
    __syncthreads();
    if (tid == 0) {
        const double tot = red[0][0] + red[0][1] + red[0][2] + red[0][3];
        advstats[step * 4 + 0] = (float)mr;
        advstats[step * 4 + 1] = (float)sqrt(tot / (double)(Bn - 1));
        advstats[step * 4 + 2] = (float)mc;
        advstats[step * 4 + 3] = 0.f;
    }
}

template <int NT1>
__global__ void __launch_bounds__(NTH, 1) ppo_train_kernel(const __grid_constant__ PpoArgs a) {
    extern __shared__ __align__(16) float sm[];
    const float nu = a.nu_dev ? *a.nu_dev : a.nu;
    const PpoSmem L = ppo_smem_layout(a.DP);
    const int tid = threadIdx.x;
    const int role = (int)cluster_ctarank();        // 0 pi, 1 vf, 2 cvf, >=3 idle (only joins the barriers)
    const int ncta = (int)cluster_nctarank();
    const bool working = role < 3;
    const int trunk = working ? role : 0;
    const int D = a.D, DP = a.DP;
    const int AOUT = (role == 0) ? a.A : 1;
    const bool has_logstd = (role == 0) && !a.is_discrete;

    float* W1t = sm + L.w1t; float* W2t = sm + L.w2t; float* W2 = sm + L.w2;
    float* B1 = sm + L.b1; float* B2 = sm + L.b2; float* HW = sm + L.hw; float* HB = sm + L.hb;
    float* LOGSTD = sm + L.logstd;
    float* X = sm + L.x; float* H1 = sm + L.h1; float* H2 = sm + L.h2; float* DH = sm + L.dh;
    float* ROWF = sm + L.rowf; float* DMEAN = sm + L.dmean; float* ACT = sm + L.act; float* MU = sm + L.mu;
    int* IDX = reinterpret_cast<int*>(sm + L.rowoff);
    float* scratch = sm + L.scratch;
    float* XCH = sm + L.xch;

    // ---- thread -> parameter ownership (fixed for the whole launch)
    const int tj2 = tid >> 4, tk2 = tid & 15;                 // W2 tile: rows j = 4*tj2.., cols k = 4*tk2..
    const int n_w1_tiles = 16 * (DP / 4);
    const int hd = tid >> 4, hk4 = tid & 15;                  // head weight: row hd, cols 4*hk4..
    // scalar slot: b1 | b2 | head bias | log_std
    int s_kind = -1, s_idx = 0;
    if (tid < 64) { s_kind = 0; s_idx = tid; }
    else if (tid < 128) { s_kind = 1; s_idx = tid - 64; }
    else if (tid < 128 + AOUT) { s_kind = 2; s_idx = tid - 128; }
    else if (tid >= 160 && tid < 160 + a.A && has_logstd) { s_kind = 3; s_idx = tid - 160; }

    auto flat_w2 = [&](int j, int k) { return (j < a.h1 && k < a.h0) ? a.off_w2[trunk] + j * a.h0 + k : -1; };
    auto flat_w1 = [&](int j, int k) { return (j < a.h0 && k < D) ? a.off_w1[trunk] + j * D + k : -1; };
    auto flat_hw = [&](int d, int k) { return (d < AOUT && k < a.h1) ? a.off_hw[trunk] + d * a.h1 + k : -1; };
    auto flat_scalar = [&]() {
        switch (s_kind) {
            case 0: return s_idx < a.h0 ? a.off_b1[trunk] + s_idx : -1;
            case 1: return s_idx < a.h1 ? a.off_b2[trunk] + s_idx : -1;
            case 2: return a.off_hb[trunk] + s_idx;
            case 3: return a.off_logstd + s_idx;
        }
        return -1;
    };

    // Adam moments in registers
    float m_w2[4][4], v_w2[4][4], m_w1[NT1][4][4], v_w1[NT1][4][4], m_hw[4], v_hw[4], m_s = 0.f, v_s = 0.f;
    // ---- load parameters into shared memory (zero padded) and moments into registers
    for (int i = tid; i < DP * H; i += NTH) W1t[i] = 0.f;
    for (int i = tid; i < H * H; i += NTH) { W2t[i] = 0.f; W2[i] = 0.f; }
    for (int i = tid; i < AMAX * WA_LD; i += NTH) HW[i] = 0.f;
    if (tid < H) { B1[tid] = 0.f; B2[tid] = 0.f; }
    if (tid < AMAX) { HB[tid] = 0.f; LOGSTD[tid] = 0.f; }
    if (tid < 16) XCH[tid] = 0.f;
    for (int i = tid; i < 2 * RB * DP; i += NTH) X[i] = 0.f;     // padding columns k in [D, DP) stay zero forever
    for (int i = tid; i < 2 * RB * 8; i += NTH) ROWF[i] = 0.f;
    for (int i = tid; i < 2 * RB * AMAX; i += NTH) ACT[i] = 0.f;
    __syncthreads();
    if (working) {
#pragma unroll
        for (int jj = 0; jj < 4; ++jj)
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                const int j = 4 * tj2 + jj, k = 4 * tk2 + kk, f = flat_w2(j, k);
                m_w2[jj][kk] = f >= 0 ? a.adam_m[f] : 0.f;
                v_w2[jj][kk] = f >= 0 ? a.adam_v[f] : 0.f;
                if (f >= 0) { const float w = a.params[f]; W2[j * H + k] = w; W2t[k * H + j] = w; }
            }
#pragma unroll
        for (int n = 0; n < NT1; ++n) {
            const int t = tid + NTH * n, tk1 = t >> 4, tj1 = t & 15;
#pragma unroll
            for (int jj = 0; jj < 4; ++jj)
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
                    const int j = 4 * tj1 + jj, k = 4 * tk1 + kk;
                    const int f = (t < n_w1_tiles) ? flat_w1(j, k) : -1;
                    m_w1[n][jj][kk] = f >= 0 ? a.adam_m[f] : 0.f;
                    v_w1[n][jj][kk] = f >= 0 ? a.adam_v[f] : 0.f;
                    if (f >= 0) W1t[k * H + j] = a.params[f];
                }
        }
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
            const int f = flat_hw(hd, 4 * hk4 + kk);
            m_hw[kk] = f >= 0 ? a.adam_m[f] : 0.f;
            v_hw[kk] = f >= 0 ? a.adam_v[f] : 0.f;
            if (f >= 0) HW[hd * WA_LD + 4 * hk4 + kk] = a.params[f];
        }
        {
            const int f = flat_scalar();
            if (f >= 0) {
                m_s = a.adam_m[f]; v_s = a.adam_v[f];
                const float p = a.params[f];
                if (s_kind == 0) B1[s_idx] = p; else if (s_kind == 1) B2[s_idx] = p;
                else if (s_kind == 2) HB[s_idx] = p; else LOGSTD[s_idx] = p;
            }
        }
    }
    __syncthreads();
    cluster_sync_all();   // every CTA has zeroed its exchange slots before any peer writes into them

    // bias-correction powers, advanced multiplicatively each step (double)
    double b1_pow = pow(a.beta1, (double)a.step_before), b2_pow = pow(a.beta2, (double)a.step_before);
    const int ty = tid >> 4, tx = tid & 15;     // forward tile: rows 4ty.., cols 4tx..
    const int hr = tid >> 2, hq = tid & 3;      // head mapping: row hr, quarter hq
    const int warp = tid >> 5, lane = tid & 31;
    const int aw = a.is_discrete ? 1 : a.A;

    // ---- chunk pipeline: the (epoch, minibatch, 64-row chunk) sequence is known up front (the prologue kernel turned
    // numpy's permutations into time-major element offsets), so chunk q+1's rows are gathered with cp.async while chunk q
    // is being computed, and chunk q+2's offsets are fetched one stage earlier still.  No global latency is exposed.
    struct Cursor { int epoch, mb, c0; };
    auto cur_valid = [&](const Cursor& c) { return c.epoch < a.n_epochs; };
    auto cur_bn = [&](const Cursor& c) { return min(a.B, a.N - c.mb * a.B); };
    auto cur_rows = [&](const Cursor& c) { return min(RB, cur_bn(c) - c.c0); };
    auto cur_next = [&](Cursor c) {
        c.c0 += RB;
        if (c.c0 >= cur_bn(c)) { c.c0 = 0; if (++c.mb >= a.steps_per_epoch) { c.mb = 0; ++c.epoch; } }
        return c;
    };
    auto prefetch_idx = [&](const Cursor& c, int slot) {          // offsets of the chunk's rows -> IDX[slot] (-1 = padding)
        if (tid < RB) {
            int* dst = IDX + slot * RB + tid;
            if (cur_valid(c) && tid < cur_rows(c))
                cp_async_4(dst, a.poff + (size_t)c.epoch * a.N + (size_t)c.mb * a.B + c.c0 + tid);
            else
                *dst = -1;
        }
    };
    auto prefetch_data = [&](int slot, int buf) {                 // gather the rows named by IDX[slot] into buffer `buf`
        const int* offs = IDX + slot * RB;
        float* Xb = X + buf * RB * DP;
        for (int r = warp; r < RB; r += NTH / 32) {
            const int o = offs[r];
            if (o >= 0) {
                const float* src = a.obs + (size_t)o * D;
                for (int k = lane; k < D; k += 32) cp_async_4(Xb + r * DP + k, src + k);
            } else {
                for (int k = lane; k < D; k += 32) Xb[r * DP + k] = 0.f;
            }
        }
        float* Rb = ROWF + buf * RB * 8;
        if (tid < RB) {
            const int o = offs[tid];
            if (o >= 0) {
                if (role == 0) {
                    cp_async_4(Rb + tid * 8 + 0, a.old_logp + o);
                    cp_async_4(Rb + tid * 8 + 1, a.adv_r + o);
                    cp_async_4(Rb + tid * 8 + 2, a.adv_c + o);
                } else if (role == 1) {
                    cp_async_4(Rb + tid * 8 + 3, a.ret_r + o);
                    cp_async_4(Rb + tid * 8 + 4, a.old_vr + o);
                } else {
                    cp_async_4(Rb + tid * 8 + 3, a.ret_c + o);
                    cp_async_4(Rb + tid * 8 + 4, a.old_vc + o);
                }
            } else {
#pragma unroll
                for (int i = 0; i < 5; ++i) Rb[tid * 8 + i] = 0.f;
            }
        }
        if (role == 0) {
            float* Ab = ACT + buf * RB * AMAX;
            const int r = tid >> 2, o = offs[r];
            for (int d = tid & 3; d < aw; d += 4) {
                if (o >= 0) cp_async_4(Ab + r * AMAX + d, a.act + (size_t)o * aw + d);
                else Ab[r * AMAX + d] = 0.f;
            }
        }
    };

    Cursor cur = {0, 0, 0};
    int q = 0;                                   // running chunk counter (buffer / slot parity)
    if (working) {
        prefetch_idx(cur, 0);
        prefetch_idx(cur_next(cur), 1);
        cp_async_commit();
        cp_async_wait_all();
        __syncthreads();
        prefetch_data(0, 0);
        cp_async_commit();
    }

    int step = 0, early_stop_epoch = a.n_epochs;
    double epoch_kl_sum = 0.0;
    bool stop_all = false;

    for (int epoch = 0; epoch < a.n_epochs && !stop_all; ++epoch) {
        epoch_kl_sum = 0.0;
        int epoch_steps = 0;
        for (int mb = 0; mb < a.steps_per_epoch && !stop_all; ++mb, ++step) {
            const int parity = step & 1;
            const int Bn = min(a.B, a.N - mb * a.B);
            const float invB = 1.0f / (float)Bn;

            // gradient accumulators (registers)
            float g_w2[4][4], g_w1[NT1][4][4], g_hw[4], g_s = 0.f;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                g_hw[i] = 0.f;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    g_w2[i][j] = 0.f;
#pragma unroll
                    for (int n = 0; n < NT1; ++n) g_w1[n][i][j] = 0.f;
                }
            }
            // loss partial sums (thread-local, one merged block reduction at the end of the step)
            float s_a = 0.f, s_b = 0.f, s_c = 0.f, s_d = 0.f, s_e = 0.f;
            // minibatch statistics of the advantages (ppo_lag.py:218-222) from the prologue kernel's table
            const float adv_mean_r = a.advstats[step * 4 + 0], adv_std_r = a.advstats[step * 4 + 1],
                        adv_mean_c = a.advstats[step * 4 + 2];

            if (working) {
                for (int c0 = 0; c0 < Bn; c0 += RB, ++q) {
                    const int rows = min(RB, Bn - c0);
                    const int buf = q & 1;
                    const float* Xc = X + buf * RB * DP;
                    float* Rc = ROWF + buf * RB * 8;
                    const float* Ac = ACT + buf * RB * AMAX;
                    cp_async_wait_all();
                    __syncthreads();          // chunk q landed (and everybody is done with chunk q-1's buffers)
                    {
                        const Cursor nxt = cur_next(cur);
                        if (cur_valid(nxt)) prefetch_data((q + 1) % 3, buf ^ 1);
                        prefetch_idx(cur_next(nxt), (q + 2) % 3);
                        cp_async_commit();
                        cur = nxt;
                    }

                    // ---- forward layer 1: H1 = tanh(X W1^T + b1)
                    {
                        float acc[4][4];
#pragma unroll
                        for (int i = 0; i < 4; ++i)
#pragma unroll
                            for (int j = 0; j < 4; ++j) acc[i][j] = B1[4 * tx + j];
                        gemm_tile_4x4<0>(acc, Xc, DP, W1t, DP, ty, tx);
#pragma unroll
                        for (int i = 0; i < 4; ++i)
                            *reinterpret_cast<float4*>(H1 + (4 * ty + i) * H + 4 * tx) =
                                make_float4(tanhf(acc[i][0]), tanhf(acc[i][1]), tanhf(acc[i][2]), tanhf(acc[i][3]));
                    }
                    __syncthreads();
                    // ---- forward layer 2: H2 = tanh(H1 W2^T + b2)
                    {
                        float acc[4][4];
#pragma unroll
                        for (int i = 0; i < 4; ++i)
#pragma unroll
                            for (int j = 0; j < 4; ++j) acc[i][j] = B2[4 * tx + j];
                        gemm_tile_4x4<H>(acc, H1, H, W2t, H, ty, tx);
#pragma unroll
                        for (int i = 0; i < 4; ++i)
                            *reinterpret_cast<float4*>(H2 + (4 * ty + i) * H + 4 * tx) =
                                make_float4(tanhf(acc[i][0]), tanhf(acc[i][1]), tanhf(acc[i][2]), tanhf(acc[i][3]));
                    }
                    __syncthreads();

                    // ---- heads + losses + d(loss)/d(head output).  4 threads per row (hr, hq).
                    if (role == 0) {
                        // action head: outputs d = hq, hq+4, hq+8, hq+12
                        float out[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const int d = hq + 4 * u;
                            float acc = 0.f;
                            if (d < a.A) {
                                acc = HB[d];
                                const float4* hrow = reinterpret_cast<const float4*>(H2 + hr * H);
                                const float4* wrow = reinterpret_cast<const float4*>(HW + d * WA_LD);
#pragma unroll
                                for (int k = 0; k < H / 4; ++k) {
                                    const float4 h = hrow[k], w = wrow[k];
                                    acc = fmaf(h.x, w.x, acc); acc = fmaf(h.y, w.y, acc);
                                    acc = fmaf(h.z, w.z, acc); acc = fmaf(h.w, w.w, acc);
                                }
                            }
                            out[u] = acc;
                            MU[hr * AMAX + d] = acc;
                        }
                        const bool valid = hr < rows;
                        float logp = 0.f, ent = 0.f;
                        float dcoef[4] = {0.f, 0.f, 0.f, 0.f};   // d logp / d out[u]
                        float dent[4] = {0.f, 0.f, 0.f, 0.f};    // d entropy / d out[u] (discrete only)
                        if (!a.is_discrete) {
                            float lp = 0.f, en = 0.f;
#pragma unroll
                            for (int u = 0; u < 4; ++u) {
                                const int d = hq + 4 * u;
                                if (d < a.A) {
                                    const float sigma = expf(LOGSTD[d]);
                                    const float var = sigma * sigma, log_scale = logf(sigma);
                                    const float diff = Ac[hr * AMAX + d] - out[u];
                                    lp += -(diff * diff) / (2.f * var) - log_scale - LOG_SQRT_2PI;
                                    en += HALF_LOG_2PI_PLUS_HALF + log_scale;
                                    dcoef[u] = diff / var;
                                }
                            }
                            lp += __shfl_xor_sync(0xffffffffu, lp, 1); lp += __shfl_xor_sync(0xffffffffu, lp, 2);
                            en += __shfl_xor_sync(0xffffffffu, en, 1); en += __shfl_xor_sync(0xffffffffu, en, 2);
                            logp = lp; ent = en;
                        } else {
                            float mx = -INFINITY;
#pragma unroll
                            for (int u = 0; u < 4; ++u) if (hq + 4 * u < a.A) mx = fmaxf(mx, out[u]);
                            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
                            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
                            float se = 0.f;
#pragma unroll
                            for (int u = 0; u < 4; ++u) if (hq + 4 * u < a.A) se += expf(out[u] - mx);
                            se += __shfl_xor_sync(0xffffffffu, se, 1); se += __shfl_xor_sync(0xffffffffu, se, 2);
                            const float lse = mx + logf(se);
                            const int ai = (int)Ac[hr * AMAX + 0];
                            float lp = 0.f, en = 0.f, pr[4], lg[4];
#pragma unroll
                            for (int u = 0; u < 4; ++u) {
                                const int d = hq + 4 * u;
                                pr[u] = 0.f; lg[u] = 0.f;
                                if (d < a.A) {
                                    lg[u] = out[u] - lse;
                                    pr[u] = expf(lg[u]);
                                    en -= lg[u] * pr[u];
                                    if (d == ai) lp = lg[u];
                                }
                            }
                            lp += __shfl_xor_sync(0xffffffffu, lp, 1); lp += __shfl_xor_sync(0xffffffffu, lp, 2);
                            en += __shfl_xor_sync(0xffffffffu, en, 1); en += __shfl_xor_sync(0xffffffffu, en, 2);
                            logp = lp; ent = en;
#pragma unroll
                            for (int u = 0; u < 4; ++u) {
                                const int d = hq + 4 * u;
                                if (d < a.A) {
                                    dcoef[u] = (d == ai ? 1.f : 0.f) - pr[u];
                                    dent[u] = -pr[u] * (lg[u] + ent);
                                }
                            }
                        }
                        // surrogate (ppo_lag.py:218-236); advantages normalised here with the minibatch statistics
                        const float old_lp = Rc[hr * 8 + 0];
                        const float A_r = (Rc[hr * 8 + 1] - adv_mean_r) / (adv_std_r + 1e-8f);
                        const float A_c = Rc[hr * 8 + 2] - adv_mean_c;
                        const float ratio = expf(logp - old_lp);
                        const float lo = 1.f - a.clip_range, hi = 1.f + a.clip_range;
                        const float clipped = fminf(fmaxf(ratio, lo), hi);
                        const float pl1 = A_r * ratio, pl2 = A_r * clipped;
                        const bool inrange = (ratio >= lo) && (ratio <= hi);
                        float wgt;   // d min(pl1, pl2) / d ratio divided by A_r (torch.min splits ties evenly)
                        if (pl1 < pl2) wgt = 1.f;
                        else if (pl1 > pl2) wgt = inrange ? 1.f : 0.f;
                        else wgt = 0.5f + (inrange ? 0.5f : 0.f);
                        const float inv1pnu = 1.f / (1.f + nu);
                        float g = 0.f;
                        if (valid) {
                            g = ratio * (-A_r * wgt + nu * A_c) * invB * inv1pnu;      // dL/dlogp
                            if (hq == 0) {
                                s_a += fminf(pl1, pl2);                                   // sum min(pl1, pl2)
                                s_b += A_c * ratio;                                       // sum cost_adv * ratio
                                s_c += (fabsf(ratio - 1.f) > a.clip_range) ? 1.f : 0.f;   // clip count
                                s_d += old_lp - logp;                                     // approx kl numerator
                                s_e += ent;                                               // entropy sum
                            }
                        }
                        const float ge = valid ? -a.ent_coef * invB : 0.f;               // d(ent_coef*entropy_loss)/d ent
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const int d = hq + 4 * u;
                            if (d < AMAX) DMEAN[hr * AMAX + d] = (d < a.A) ? (g * dcoef[u] + ge * dent[u]) : 0.f;
                        }
                        if (hq == 0) Rc[hr * 8 + 5] = g;
                    } else {
                        // value head: partial dot over k in [16hq, 16hq+16)
                        float acc = 0.f;
                        const float4* hrow = reinterpret_cast<const float4*>(H2 + hr * H + 16 * hq);
                        const float4* wrow = reinterpret_cast<const float4*>(HW + 16 * hq);
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const float4 h = hrow[k], w = wrow[k];
                            acc = fmaf(h.x, w.x, acc); acc = fmaf(h.y, w.y, acc);
                            acc = fmaf(h.z, w.z, acc); acc = fmaf(h.w, w.w, acc);
                        }
                        acc += __shfl_xor_sync(0xffffffffu, acc, 1);
                        acc += __shfl_xor_sync(0xffffffffu, acc, 2);
                        const float V = acc + HB[0];
                        const float target = Rc[hr * 8 + 3], oldv = Rc[hr * 8 + 4];
                        const bool clipvf = (role == 1) ? a.has_clip_vf_r : a.has_clip_vf_c;
                        const float cr = (role == 1) ? a.clip_vf_r : a.clip_vf_c;
                        float Vp = V, pass = 1.f;
                        if (clipvf) {
                            const float dv = V - oldv;
                            Vp = oldv + fminf(fmaxf(dv, -cr), cr);
                            pass = (dv >= -cr && dv <= cr) ? 1.f : 0.f;
                        }
                        const float coef = (role == 1) ? a.vf_coef_r : a.vf_coef_c;
                        const float err = Vp - target;
                        float dV = 0.f;
                        if (hr < rows) {
                            dV = coef * 2.f * err * invB * pass;
                            if (hq == 0) s_a += err * err;
                        }
                        if (hq == 0) { DMEAN[hr * AMAX + 0] = dV; }
                    }
                    __syncthreads();

                    // ---- head weight / bias / log_std gradients (rows beyond `rows` carry zero dmean: loops run over RB)
                    if (hd < AOUT) {
                        float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
#pragma unroll 8
                        for (int r = 0; r < RB; ++r) {
                            const float dm = DMEAN[r * AMAX + hd];
                            const float4 h = *reinterpret_cast<const float4*>(H2 + r * H + 4 * hk4);
                            acc0 = fmaf(dm, h.x, acc0); acc1 = fmaf(dm, h.y, acc1);
                            acc2 = fmaf(dm, h.z, acc2); acc3 = fmaf(dm, h.w, acc3);
                        }
                        g_hw[0] += acc0; g_hw[1] += acc1; g_hw[2] += acc2; g_hw[3] += acc3;
                    }
                    if (s_kind == 2) {
                        float acc = 0.f;
#pragma unroll 8
                        for (int r = 0; r < RB; ++r) acc += DMEAN[r * AMAX + s_idx];
                        g_s += acc;
                    } else if (s_kind == 3) {
                        // d logp / d log_std_d = diff^2/var - 1 ; entropy: d(-mean H)/d log_std = -1
                        const float sigma = expf(LOGSTD[s_idx]);
                        const float inv_var = 1.f / (sigma * sigma);
                        float acc = 0.f;
#pragma unroll 8
                        for (int r = 0; r < RB; ++r) {
                            const float diff = Ac[r * AMAX + s_idx] - MU[r * AMAX + s_idx];
                            acc = fmaf(Rc[r * 8 + 5], diff * diff * inv_var - 1.f, acc);
                        }
                        g_s += acc - a.ent_coef * (float)rows * invB;
                    }
                    // ---- dH2pre[r][k] = (sum_d dmean[r][d] * HW[d][k]) * (1 - H2^2)   (thread: row hr, k in [16hq,16hq+16))
                    {
                        float dloc[16];
#pragma unroll
                        for (int k = 0; k < 16; ++k) dloc[k] = 0.f;
                        for (int d = 0; d < AOUT; ++d) {
                            const float dm = DMEAN[hr * AMAX + d];
                            const float4* wrow = reinterpret_cast<const float4*>(HW + d * WA_LD + 16 * hq);
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                const float4 w = wrow[k];
                                dloc[4 * k + 0] = fmaf(dm, w.x, dloc[4 * k + 0]);
                                dloc[4 * k + 1] = fmaf(dm, w.y, dloc[4 * k + 1]);
                                dloc[4 * k + 2] = fmaf(dm, w.z, dloc[4 * k + 2]);
                                dloc[4 * k + 3] = fmaf(dm, w.w, dloc[4 * k + 3]);
                            }
                        }
                        const float4* hrow = reinterpret_cast<const float4*>(H2 + hr * H + 16 * hq);
                        float4* drow = reinterpret_cast<float4*>(DH + hr * H + 16 * hq);
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const float4 h = hrow[k];
                            drow[k] = make_float4(dloc[4 * k + 0] * (1.f - h.x * h.x), dloc[4 * k + 1] * (1.f - h.y * h.y),
                                                  dloc[4 * k + 2] * (1.f - h.z * h.z), dloc[4 * k + 3] * (1.f - h.w * h.w));
                        }
                    }
                    __syncthreads();

                    // ---- dW2 += dH2pre^T H1 ; db2 ; dH1pre = (dH2pre W2) * (1 - H1^2) -> written over H2
                    outer_tile_4x4(g_w2, DH, H, H1, H, tj2, tk2);
                    if (s_kind == 1) {
                        float acc = 0.f;
#pragma unroll 8
                        for (int r = 0; r < RB; ++r) acc += DH[r * H + s_idx];
                        g_s += acc;
                    }
                    {
                        float acc[4][4];
#pragma unroll
                        for (int i = 0; i < 4; ++i)
#pragma unroll
                            for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
                        gemm_tile_4x4<H>(acc, DH, H, W2, H, ty, tx);   // sum_j dH2[r][j] * W2[j][k]  (W2 row-major == "k-major" in j)
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const float4 h = *reinterpret_cast<const float4*>(H1 + (4 * ty + i) * H + 4 * tx);
                            *reinterpret_cast<float4*>(H2 + (4 * ty + i) * H + 4 * tx) =
                                make_float4(acc[i][0] * (1.f - h.x * h.x), acc[i][1] * (1.f - h.y * h.y),
                                            acc[i][2] * (1.f - h.z * h.z), acc[i][3] * (1.f - h.w * h.w));
                        }
                    }
                    __syncthreads();
                    // ---- dW1 += dH1pre^T X ; db1
#pragma unroll
                    for (int n = 0; n < NT1; ++n) {
                        const int t = tid + NTH * n;
                        if (t < n_w1_tiles) outer_tile_4x4(g_w1[n], H2, H, Xc, DP, t & 15, t >> 4);
                    }
                    if (s_kind == 0) {
                        float acc = 0.f;
#pragma unroll 8
                        for (int r = 0; r < RB; ++r) acc += H2[r * H + s_idx];
                        g_s += acc;
                    }
                }  // chunks
            }      // working

            // ---- one merged block reduction: loss partial sums + local sum of squared gradients
            float ss = 0.f;
            if (working) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    ss = fmaf(g_hw[i], g_hw[i], ss);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        ss = fmaf(g_w2[i][j], g_w2[i][j], ss);
#pragma unroll
                        for (int n = 0; n < NT1; ++n) ss = fmaf(g_w1[n][i][j], g_w1[n][i][j], ss);
                    }
                }
                ss = fmaf(g_s, g_s, ss);
            }
            float red[6] = {s_a, s_b, s_c, s_d, s_e, ss};
#pragma unroll
            for (int i = 0; i < 6; ++i) red[i] = warp_sum(red[i]);
            if (lane == 0) {
#pragma unroll
                for (int i = 0; i < 6; ++i) scratch[warp * 8 + i] = red[i];
            }
            __syncthreads();
#pragma unroll
            for (int i = 0; i < 6; ++i) {
                float t = 0.f;
#pragma unroll
                for (int wv = 0; wv < NTH / 32; ++wv) t += scratch[wv * 8 + i];
                red[i] = t;
            }
            ss = red[5];
            const size_t so = (size_t)step * ICRL_PPO_STATS_PER_STEP;
            float kl_step = 0.f;
            if (role == 0) {
                float pl = -(red[0] * invB);
                pl = pl + nu * (red[1] * invB);
                pl = pl / (1.f + nu);
                kl_step = red[3] * invB;
                if (tid == 0) {
                    a.stats[so + 0] = pl;
                    a.stats[so + 1] = red[2] * invB;
                    a.stats[so + 4] = -(red[4] * invB);
                    a.stats[so + 5] = kl_step;
                }
            } else if (working) {
                if (tid == 0) a.stats[so + (role == 1 ? 2 : 3)] = red[0] * invB;
            }

            // ---- global gradient norm: local sum of squares -> DSMEM exchange -> cluster barrier
            // epoch-level KL early stop is decided by the pi CTA right here (it has this step's KL) and rides along
            float stop_flag = 0.f;
            if (role == 0) {
                epoch_kl_sum += (double)kl_step;
                ++epoch_steps;
                const bool last_of_epoch = (mb == a.steps_per_epoch - 1);
                if (last_of_epoch && a.has_target_kl && (epoch_kl_sum / epoch_steps) > 1.5 * a.target_kl) stop_flag = 1.f;
                if (a.max_steps > 0 && step + 1 >= a.max_steps) stop_flag = 2.f;
            }
            if (tid < ncta && working) {
                st_remote_f32(XCH + (parity * 4 + role) * 2 + 0, (uint32_t)tid, ss);
                st_remote_f32(XCH + (parity * 4 + role) * 2 + 1, (uint32_t)tid, stop_flag);
            }
            cluster_sync_all();
            const float total_ss = XCH[(parity * 4 + 0) * 2] + XCH[(parity * 4 + 1) * 2] + XCH[(parity * 4 + 2) * 2];
            const float stop_rx = XCH[(parity * 4 + 0) * 2 + 1];
            const float total_norm = sqrtf(total_ss);
            const float clip_coef = fminf(a.max_grad_norm / (total_norm + 1e-6f), 1.0f);
            if (role == 0 && tid == 0) {
                a.stats[so + 7] = total_norm;
                a.stats[so + 6] = 0.f;   // total loss is assembled on the host from the parts (needs all three CTAs)
            }

            // ---- Adam (each thread updates the parameters it owns; smem copies refreshed in place)
            if (working) {
                b1_pow *= a.beta1; b2_pow *= a.beta2;
                AdamConsts ac;
                ac.one_minus_b1 = (float)(1.0 - a.beta1);
                ac.b2 = (float)a.beta2;
                ac.one_minus_b2 = (float)(1.0 - a.beta2);
                ac.inv_bc2_sqrt = (float)(1.0 / sqrt(1.0 - b2_pow));
                ac.eps = (float)a.adam_eps;
                ac.neg_step_size = (float)(-(a.lr / (1.0 - b1_pow)));
#pragma unroll
                for (int jj = 0; jj < 4; ++jj)
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) {
                        const int j = 4 * tj2 + jj, k = 4 * tk2 + kk;
                        if (j < a.h1 && k < a.h0) {
                            const float p = adam_update(W2[j * H + k], g_w2[jj][kk] * clip_coef, m_w2[jj][kk], v_w2[jj][kk], ac);
                            W2[j * H + k] = p; W2t[k * H + j] = p;
                        }
                    }
#pragma unroll
                for (int n = 0; n < NT1; ++n) {
                    const int t = tid + NTH * n, tk1 = t >> 4, tj1 = t & 15;
                    if (t < n_w1_tiles) {
#pragma unroll
                        for (int jj = 0; jj < 4; ++jj)
#pragma unroll
                            for (int kk = 0; kk < 4; ++kk) {
                                const int j = 4 * tj1 + jj, k = 4 * tk1 + kk;
                                if (j < a.h0 && k < D)
                                    W1t[k * H + j] = adam_update(W1t[k * H + j], g_w1[n][jj][kk] * clip_coef, m_w1[n][jj][kk],
                                                                 v_w1[n][jj][kk], ac);
                            }
                    }
                }
                if (hd < AOUT) {
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) {
                        const int k = 4 * hk4 + kk;
                        if (k < a.h1)
                            HW[hd * WA_LD + k] = adam_update(HW[hd * WA_LD + k], g_hw[kk] * clip_coef, m_hw[kk], v_hw[kk], ac);
                    }
                }
                if (s_kind >= 0 && flat_scalar() >= 0) {
                    float* slot = s_kind == 0 ? &B1[s_idx] : s_kind == 1 ? &B2[s_idx] : s_kind == 2 ? &HB[s_idx] : &LOGSTD[s_idx];
                    *slot = adam_update(*slot, g_s * clip_coef, m_s, v_s, ac);
                }
            }
            if (stop_rx == 1.f) { early_stop_epoch = epoch; stop_all = true; }
            if (stop_rx == 2.f) { stop_all = true; }
            // (the next chunk's leading __syncthreads orders these shared-memory weight updates before their first use)
        }  // minibatches
    }      // epochs
    cp_async_wait_all();
    __syncthreads();

    // ---- write back parameters and moments
    if (working) {
#pragma unroll
        for (int jj = 0; jj < 4; ++jj)
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                const int j = 4 * tj2 + jj, k = 4 * tk2 + kk, f = flat_w2(j, k);
                if (f >= 0) { a.params[f] = W2[j * H + k]; a.adam_m[f] = m_w2[jj][kk]; a.adam_v[f] = v_w2[jj][kk]; }
            }
#pragma unroll
        for (int n = 0; n < NT1; ++n) {
            const int t = tid + NTH * n, tk1 = t >> 4, tj1 = t & 15;
#pragma unroll
            for (int jj = 0; jj < 4; ++jj)
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
                    const int j = 4 * tj1 + jj, k = 4 * tk1 + kk;
                    const int f = (t < n_w1_tiles) ? flat_w1(j, k) : -1;
                    if (f >= 0) { a.params[f] = W1t[k * H + j]; a.adam_m[f] = m_w1[n][jj][kk]; a.adam_v[f] = v_w1[n][jj][kk]; }
                }
        }
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
            const int f = flat_hw(hd, 4 * hk4 + kk);
            if (f >= 0) { a.params[f] = HW[hd * WA_LD + 4 * hk4 + kk]; a.adam_m[f] = m_hw[kk]; a.adam_v[f] = v_hw[kk]; }
        }
        const int f = flat_scalar();
        if (f >= 0) {
            const float* slot = s_kind == 0 ? &B1[s_idx] : s_kind == 1 ? &B2[s_idx] : s_kind == 2 ? &HB[s_idx] : &LOGSTD[s_idx];
            a.params[f] = *slot; a.adam_m[f] = m_s; a.adam_v[f] = v_s;
        }
    }
    if (role == 0 && tid == 0) {
        a.result[0] = early_stop_epoch;
        a.result[1] = step;
        a.result[2] = 0;
        a.result[3] = 0;
    }
    cluster_sync_all();   // nobody exits while a peer may still address its shared memory
}


// ---------------------------------------------------------------- policy forward (rollout / evaluation side)
// grid = (row chunks, 3 trunks).  Each CTA keeps its trunk in shared memory and walks 64-row chunks.
__global__ void __launch_bounds__(NTH) policy_forward_kernel(const __grid_constant__ PpoArgs a, const float* __restrict__ obs,
                                                             long long n, float* __restrict__ head,
                                                             float* __restrict__ values, float* __restrict__ cost_values) {
    extern __shared__ __align__(16) float sm[];
    const PpoSmem L = ppo_smem_layout(a.DP);
    const int tid = threadIdx.x, trunk = blockIdx.y, D = a.D, DP = a.DP;
    const int AOUT = trunk == 0 ? a.A : 1;
    float* W1t = sm + L.w1t; float* W2t = sm + L.w2t; float* B1 = sm + L.b1; float* B2 = sm + L.b2;
    float* HW = sm + L.hw; float* HB = sm + L.hb; float* X = sm + L.x; float* H1 = sm + L.h1; float* H2 = sm + L.h2;
    for (int i = tid; i < DP * H; i += NTH) {
        const int k = i / H, j = i - k * H;
        W1t[i] = (k < D && j < a.h0) ? a.params[a.off_w1[trunk] + j * D + k] : 0.f;
    }
    for (int i = tid; i < H * H; i += NTH) {
        const int k = i / H, j = i - k * H;
        W2t[i] = (k < a.h0 && j < a.h1) ? a.params[a.off_w2[trunk] + j * a.h0 + k] : 0.f;
    }
    for (int i = tid; i < AMAX * WA_LD; i += NTH) {
        const int d = i / WA_LD, k = i - d * WA_LD;
        HW[i] = (d < AOUT && k < a.h1) ? a.params[a.off_hw[trunk] + d * a.h1 + k] : 0.f;
    }
    if (tid < H) {
        B1[tid] = tid < a.h0 ? a.params[a.off_b1[trunk] + tid] : 0.f;
        B2[tid] = tid < a.h1 ? a.params[a.off_b2[trunk] + tid] : 0.f;
    }
    if (tid < AMAX) HB[tid] = tid < AOUT ? a.params[a.off_hb[trunk] + tid] : 0.f;
    const int ty = tid >> 4, tx = tid & 15, hr = tid >> 2, hq = tid & 3;
    const long long n_chunks = (n + RB - 1) / RB;
    for (long long c = blockIdx.x; c < n_chunks; c += gridDim.x) {
        const long long row0 = c * RB;
        const int rows = (int)min((long long)RB, n - row0);
        __syncthreads();
        for (int i = tid; i < RB * DP; i += NTH) {
            const int r = i / DP, k = i - r * DP;
            X[i] = (r < rows && k < D) ? obs[(row0 + r) * D + k] : 0.f;
        }
        __syncthreads();
        float acc[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = B1[4 * tx + j];
        gemm_tile_4x4<0>(acc, X, DP, W1t, DP, ty, tx);
#pragma unroll
        for (int i = 0; i < 4; ++i)
            *reinterpret_cast<float4*>(H1 + (4 * ty + i) * H + 4 * tx) =
                make_float4(tanhf(acc[i][0]), tanhf(acc[i][1]), tanhf(acc[i][2]), tanhf(acc[i][3]));
        __syncthreads();
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = B2[4 * tx + j];
        gemm_tile_4x4<H>(acc, H1, H, W2t, H, ty, tx);
#pragma unroll
        for (int i = 0; i < 4; ++i)
            *reinterpret_cast<float4*>(H2 + (4 * ty + i) * H + 4 * tx) =
                make_float4(tanhf(acc[i][0]), tanhf(acc[i][1]), tanhf(acc[i][2]), tanhf(acc[i][3]));
        __syncthreads();
        for (int u = 0; u < 4; ++u) {
            const int d = hq + 4 * u;
            if (d < AOUT && hr < rows) {
                float o = HB[d];
                const float4* hrow = reinterpret_cast<const float4*>(H2 + hr * H);
                const float4* wrow = reinterpret_cast<const float4*>(HW + d * WA_LD);
#pragma unroll
                for (int k = 0; k < H / 4; ++k) {
                    const float4 h = hrow[k], w = wrow[k];
                    o = fmaf(h.x, w.x, o); o = fmaf(h.y, w.y, o); o = fmaf(h.z, w.z, o); o = fmaf(h.w, w.w, o);
                }
                if (trunk == 0) head[(row0 + hr) * a.A + d] = o;
                else if (trunk == 1) values[row0 + hr] = o;
                else cost_values[row0 + hr] = o;
            }
        }
    }
}

// ---------------------------------------------------------------- host side
static int fill_offsets(PpoArgs& a) {
    int o = 0;
    a.off_logstd = a.is_discrete ? -1 : 0;
    if (!a.is_discrete) o += a.A;
    for (int t = 0; t < 3; ++t) {
        a.off_w1[t] = o; o += a.h0 * a.D;
        a.off_b1[t] = o; o += a.h0;
        a.off_w2[t] = o; o += a.h1 * a.h0;
        a.off_b2[t] = o; o += a.h1;
    }
    const int outs[3] = {a.A, 1, 1};
    for (int t = 0; t < 3; ++t) {
        a.off_hw[t] = o; o += outs[t] * a.h1;
        a.off_hb[t] = o; o += outs[t];
    }
    return o;
}

static int make_args(const icrl_ppo_cfg* c, PpoArgs& a) {
    ICRL_CHECK_ARG(c != nullptr, "ppo cfg is NULL");
    ICRL_CHECK_ARG(c->obs_dim >= 1, "obs_dim must be >= 1");
    ICRL_CHECK_ARG(c->act_dim >= 1 && c->act_dim <= AMAX, "act_dim %d out of range (1..%d)", c->act_dim, AMAX);
    ICRL_CHECK_ARG(c->hidden[0] >= 1 && c->hidden[0] <= H && c->hidden[1] >= 1 && c->hidden[1] <= H,
                   "policy hidden sizes (%d, %d) must be in 1..%d (two hidden layers per trunk)", c->hidden[0],
                   c->hidden[1], H);
    a.D = c->obs_dim; a.DP = (c->obs_dim + 3) / 4 * 4; a.A = c->act_dim; a.is_discrete = c->is_discrete;
    a.h0 = c->hidden[0]; a.h1 = c->hidden[1];
    a.T = c->T; a.E = c->E; a.N = c->T * c->E;
    a.B = c->batch_size > 0 ? c->batch_size : a.N;
    if (a.B > a.N && a.N > 0) a.B = a.N;
    a.n_epochs = c->n_epochs;
    a.steps_per_epoch = a.N > 0 ? (a.N + a.B - 1) / a.B : 0;
    a.max_steps = c->max_steps;
    a.has_target_kl = c->has_target_kl; a.has_clip_vf_r = c->has_clip_vf_reward; a.has_clip_vf_c = c->has_clip_vf_cost;
    a.clip_range = c->clip_range; a.clip_vf_r = c->clip_range_reward_vf; a.clip_vf_c = c->clip_range_cost_vf;
    a.ent_coef = c->ent_coef; a.vf_coef_r = c->reward_vf_coef; a.vf_coef_c = c->cost_vf_coef;
    a.max_grad_norm = c->max_grad_norm; a.nu = c->nu; a.target_kl = c->target_kl;
    a.lr = c->lr; a.beta1 = c->adam_beta1; a.beta2 = c->adam_beta2; a.adam_eps = c->adam_eps;
    fill_offsets(a);
    return 0;
}

template <int NT1>
static int launch_ppo(const PpoArgs& a, cudaStream_t st) {
    auto kern = ppo_train_kernel<NT1>;
    const PpoSmem L = ppo_smem_layout(a.DP);
    if (L.total_bytes > 227 * 1024) {
        set_error("obs_dim %d needs %d bytes of shared memory (> 227 KB)", a.D, L.total_bytes);
        return ICRL_EUNSUPPORTED;
    }
    ICRL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L.total_bytes));
    cudaError_t err = cudaErrorUnknown;
    // one CTA per trunk; clusters of 3 are legal, but fall back to 4 (one idle CTA) should a driver refuse
    for (int cluster = 3; cluster <= 4; ++cluster) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(cluster);
        cfg.blockDim = dim3(NTH);
        cfg.dynamicSmemBytes = L.total_bytes;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = cluster;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        err = cudaLaunchKernelEx(&cfg, kern, a);
        if (err == cudaSuccess) break;
        (void)cudaGetLastError();
    }
    if (err != cudaSuccess) {
        set_error("ppo_train_kernel launch failed: %s", cudaGetErrorString(err));
        return (int)err;
    }
    count_launch();
    return 0;
}

}  // namespace icrl

extern "C" {

int64_t icrl_ppo_param_count(const icrl_ppo_cfg* cfg) {
    icrl::PpoArgs a = {};
    if (icrl::make_args(cfg, a)) return -1;
    return icrl::fill_offsets(a);
}

int icrl_ppo_train(const icrl_ppo_cfg* cfg, const icrl_ppo_data* data, float* params, float* adam_m, float* adam_v,
                   int64_t adam_step_before, float* step_stats, int32_t* result, void* stream) {
    icrl::PpoArgs a = {};
    int rc = icrl::make_args(cfg, a);
    if (rc) return rc;
    ICRL_CHECK_ARG(data && params && adam_m && adam_v && step_stats && result, "NULL pointer passed to icrl_ppo_train");
    ICRL_CHECK_ARG(a.N > 0 && a.n_epochs > 0, "empty rollout buffer or n_epochs <= 0");
    ICRL_CHECK_ARG(data->observations && data->actions && data->old_log_prob && data->reward_advantages &&
                       data->reward_returns && data->cost_advantages && data->cost_returns && data->perm,
                   "NULL rollout array");
    ICRL_CHECK_ARG((!a.has_clip_vf_r || data->old_reward_values) && (!a.has_clip_vf_c || data->old_cost_values),
                   "value clipping enabled but old values are NULL");
    a.obs = data->observations; a.act = data->actions; a.old_logp = data->old_log_prob;
    a.old_vr = data->old_reward_values ? data->old_reward_values : data->reward_returns;
    a.adv_r = data->reward_advantages; a.ret_r = data->reward_returns;
    a.old_vc = data->old_cost_values ? data->old_cost_values : data->cost_returns;
    a.adv_c = data->cost_advantages; a.ret_c = data->cost_returns;
    a.perm = data->perm;
    a.nu_dev = data->nu_device;
    a.params = params; a.adam_m = adam_m; a.adam_v = adam_v; a.stats = step_stats; a.result = result;
    a.step_before = adam_step_before;
    {
        const int total_steps = a.n_epochs * a.steps_per_epoch;
        void *poff, *advstats;
        if ((rc = icrl::device_scratch(icrl::SLOT_PPO0, (size_t)a.n_epochs * a.N * sizeof(int), &poff))) return rc;
        if ((rc = icrl::device_scratch(icrl::SLOT_PPO1, (size_t)total_steps * 4 * sizeof(float), &advstats))) return rc;
        icrl::ppo_prologue_kernel<<<total_steps, 128, 0, (cudaStream_t)stream>>>(
            a.perm, (int*)poff, (float*)advstats, a.adv_r, a.adv_c, a.T, a.E, a.N, a.B, a.steps_per_epoch);
        ICRL_LAUNCH_CHECK();
        a.poff = (const int*)poff;
        a.advstats = (const float*)advstats;
    }
    const int n_tiles = 16 * (a.DP / 4);
    const int nt1 = (n_tiles + icrl::NTH - 1) / icrl::NTH;
    switch (nt1) {
        case 1: return icrl::launch_ppo<1>(a, (cudaStream_t)stream);
        case 2: return icrl::launch_ppo<2>(a, (cudaStream_t)stream);
        case 3: return icrl::launch_ppo<3>(a, (cudaStream_t)stream);
    }
    icrl::set_error("obs_dim %d too large for the PPO kernel (max 192)", a.D);
    return ICRL_EUNSUPPORTED;
}

int icrl_policy_forward(const icrl_ppo_cfg* cfg, const float* params, const float* obs, int64_t n, float* head,
                        float* values, float* cost_values, void* stream) {
    icrl::PpoArgs a = {};
    icrl_ppo_cfg c = *cfg;
    if (c.T <= 0) c.T = 1;
    if (c.E <= 0) c.E = 1;
    int rc = icrl::make_args(&c, a);
    if (rc) return rc;
    if (n == 0) return 0;
    ICRL_CHECK_ARG(params && obs && head && values && cost_values && n > 0, "NULL pointer passed to icrl_policy_forward");
    a.params = const_cast<float*>(params);
    const icrl::PpoSmem L = icrl::ppo_smem_layout(a.DP);
    ICRL_CUDA(cudaFuncSetAttribute(icrl::policy_forward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, L.total_bytes));
    const int64_t chunks = (n + icrl::RB - 1) / icrl::RB;
    const int gx = (int)(chunks < icrl::sm_count() ? chunks : icrl::sm_count());
    icrl::policy_forward_kernel<<<dim3(gx, 3), icrl::NTH, L.total_bytes, (cudaStream_t)stream>>>(a, obs, n, head, values,
                                                                                                cost_values);
    ICRL_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
