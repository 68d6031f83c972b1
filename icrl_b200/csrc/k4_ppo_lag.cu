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
#include <stdlib.h>

#include "common.cuh"

namespace icrl {

constexpr int H = 64;          // padded hidden width (both layers)
constexpr int RB = 64;         // rows per chunk
constexpr int NTH = 256;       // threads per CTA
constexpr int AMAX = 16;       // max action dims / discrete actions
constexpr int LDH = 68;        // row stride of the activation tiles in the train kernel (bank-conflict-free float4 rows)
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
    const float *xs, *as, *ss; // minibatch-ordered streams built by the prologue: obs [P][DP], actions [P][AP], scalars [P][8]
    int AP;                    // padded action width of the stream (multiple of 4)
    const float* advstats;    // [steps][8]: mean(adv_r), std(adv_r) (unbiased), mean(adv_c), 1/sqrt(1-b2^t), -lr/(1-b1^t)
    const float* nu_dev;
    float *params, *adam_m, *adam_v, *stats;
    int* result;
    unsigned long long* timing;   // optional [3 roles][16 phases] cycle accumulators (profiling aid)
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
__device__ __forceinline__ void cp_async_8(void* dst_smem, const void* src_gmem) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst_smem)), "l"(src_gmem) : "memory");
}
__device__ __forceinline__ void cp_async_16(void* dst_smem, const void* src_gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst_smem)), "l"(src_gmem) : "memory");
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
    int w1t, w2t, w2, b1, b2, hw, hb, logstd, sig, x, h1, h2, dh, rowf, dmean, act, mu, rowoff, scratch, xch, total_bytes;
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
    s.sig = o; o += 2 * AMAX;        // per action dim: 1/sigma^2, log(sigma) -- refreshed by the thread that updates log_std
    s.x = o; o += 2 * RB * DP;      // double buffered (cp.async prefetch of the next chunk)
    s.h1 = o; o += RB * LDH;
    s.h2 = o; o += RB * LDH;
    s.dh = o; o += RB * LDH;
    s.rowf = o; o += 2 * RB * 8;      // per-row scalars: 0 old_logp, 1 adv_r~, 2 adv_c~, 3 target return, 4 old value, 5 g/dV
    s.dmean = o; o += RB * AMAX;
    s.act = o; o += 2 * RB * AMAX;
    s.mu = o; o += RB * AMAX;        // action-head outputs (means / logits)
    s.rowoff = o; o += 4;            // two 8-byte mbarriers (one per chunk buffer)
    s.scratch = o; o += 64;         // 32 floats / 16 doubles of reduction scratch (8-byte aligned: o is even)
    s.xch = o; o += 2 * 4 * 2;      // [parity][rank][{sumsq, stop}]
    s.total_bytes = o * 4;
    return s;
}

// 64x64 += A[64 x K] * Bt[K x 64]  (A row-major lda, Bt k-major ld 64); thread tile rows 4ty.., cols 4tx..
// KC > 0: compile-time K (fully unrolled so operand loads run ahead of the FMAs); KC == 0: runtime K.
// SWZ: Bt's float4 column slots are XOR-swizzled with (k >> 2) & 15 (the W2t copy: lets the Adam phase write the
// transposed tile with conflict-free 128-bit stores while these row reads stay conflict-free).
template <int KC, bool SWZ = false>
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
        for (int kk = 0; kk < 4; ++kk)
            w[kk] = SWZ ? *reinterpret_cast<const float4*>(Bt + (k + kk) * H + 4 * (tx ^ ((k >> 2) & 15)))
                        : *reinterpret_cast<const float4*>(b0 + (k + kk) * H);
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

// (unused: measured no gain, kept for reference) tanh(x) = 1 - 2 / (1 + e^{2x}) on the SFU (ex2.approx + rcp.approx): ~7 instructions instead of tanhf's ~30, absolute
// error <= 2e-7 (about 2 ulp of 1.0), saturates correctly to +-1 for large |x|.
__device__ __forceinline__ float tanh_fast(float x) {
    const float e = __expf(2.f * x);
    return 1.f - __fdividef(2.f, 1.f + e);
}

// physical column of element (k, j) in the swizzled transposed copy W2t[k][.]
__device__ __forceinline__ int w2t_col(int k, int j) { return (((j >> 2) ^ ((k >> 2) & 15)) << 2) | (j & 3); }

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

// ---------------------------------------------------------------- prologue kernels (massively parallel, HBM-bound)
// The minibatch schedule is known before the first optimiser step (numpy's permutations come from the host), so the
// random-access gather is taken OUT of the dependent step chain: one pass builds minibatch-ordered streams
//     xs[p][DP] obs (zero padded), as[p][AP] actions, ss[p][8] = {old_logp, adv_r, adv_c, ret_r, old_vr, ret_c, old_vc, -}
// for p = epoch*N + position, translating numpy's env-major index (row = e*T + t, buffers.py:52-65) to the buffer's
// time-major storage on the fly.  The persistent kernel then fetches each 64-row chunk with TMA bulk copies.
__global__ void __launch_bounds__(256) ppo_gather_kernel(const __grid_constant__ PpoArgs a, float* __restrict__ xs,
                                                         float* __restrict__ as, float* __restrict__ ss,
                                                         long long n_rows, long long n_alloc) {
    const int lane = threadIdx.x & 31;
    const int aw = a.is_discrete ? 1 : a.A;
    for (long long p = (long long)blockIdx.x * 8 + (threadIdx.x >> 5); p < n_alloc; p += (long long)gridDim.x * 8) {
        int o = -1;
        if (p < n_rows) {
            const int row = a.perm[p], t = row % a.T, e = row / a.T;
            o = t * a.E + e;
        }
        for (int k = lane; k < a.DP; k += 32) xs[p * a.DP + k] = (o >= 0 && k < a.D) ? a.obs[(size_t)o * a.D + k] : 0.f;
        if (lane < a.AP) as[p * a.AP + lane] = (o >= 0 && lane < aw) ? a.act[(size_t)o * aw + lane] : 0.f;
        if (lane < 8) {
            float v = 0.f;
            if (o >= 0) {
                switch (lane) {
                    case 0: v = a.old_logp[o]; break;
                    case 1: v = a.adv_r[o]; break;
                    case 2: v = a.adv_c[o]; break;
                    case 3: v = a.ret_r[o]; break;
                    case 4: v = a.old_vr[o]; break;
                    case 5: v = a.ret_c[o]; break;
                    case 6: v = a.old_vc[o]; break;
                }
            }
            ss[p * 8 + lane] = v;
        }
    }
}

// One CTA per optimiser step: the minibatch statistics of ppo_lag.py:218-222 (mean and unbiased std of the reward
// advantages, mean of the cost advantages; float64 accumulation like ATen's CPU reductions) and that step's Adam bias
// corrections.  They depend only on data and the step number, never on parameters.
__global__ void __launch_bounds__(128) ppo_stats_kernel(const float* __restrict__ ss, float* __restrict__ advstats, int N,
                                                        int B, int spe, double beta1, double beta2, double lr,
                                                        long long step_before) {
    __shared__ double red[2][4];
    const int step = blockIdx.x, epoch = step / spe, mb = step - epoch * spe;
    const int Bn = min(B, N - mb * B);
    const size_t base = (size_t)epoch * N + (size_t)mb * B;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double sr = 0.0, sc = 0.0;
    for (int i = tid; i < Bn; i += blockDim.x) {
        sr += (double)ss[(base + i) * 8 + 1];
        sc += (double)ss[(base + i) * 8 + 2];
    }
    sr = warp_sum(sr); sc = warp_sum(sc);
    if (lane == 0) { red[0][warp] = sr; red[1][warp] = sc; }
    __syncthreads();
    const double mr = (red[0][0] + red[0][1] + red[0][2] + red[0][3]) / Bn;
    const double mc = (red[1][0] + red[1][1] + red[1][2] + red[1][3]) / Bn;
    __syncthreads();
    double ssq = 0.0;
    for (int i = tid; i < Bn; i += blockDim.x) {
        const double dv = (double)ss[(base + i) * 8 + 1] - mr;
        ssq += dv * dv;
    }
    ssq = warp_sum(ssq);
    if (lane == 0) red[0][warp] = ssq;
    __syncthreads();
    if (tid == 0) {
        const double tot = red[0][0] + red[0][1] + red[0][2] + red[0][3];
        advstats[step * 8 + 0] = (float)mr;
        advstats[step * 8 + 1] = (float)sqrt(tot / (double)(Bn - 1));
        advstats[step * 8 + 2] = (float)mc;
        // Adam bias corrections of optimiser step t (torch/optim/adam.py: 1 - beta ** step, python float64 pow)
        const double t = (double)(step_before + step + 1);
        advstats[step * 8 + 3] = (float)(1.0 / sqrt(1.0 - pow(beta2, t)));
        advstats[step * 8 + 4] = (float)(-(lr / (1.0 - pow(beta1, t))));
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
    float* LOGSTD = sm + L.logstd; float* SIG = sm + L.sig;
    float* X = sm + L.x; float* H1 = sm + L.h1; float* H2 = sm + L.h2; float* DH = sm + L.dh;
    float* ROWF = sm + L.rowf; float* DMEAN = sm + L.dmean; float* ACT = sm + L.act; float* MU = sm + L.mu;
    uint64_t* BAR = reinterpret_cast<uint64_t*>(sm + L.rowoff);
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
    __syncthreads();
    if (working) {
#pragma unroll
        for (int jj = 0; jj < 4; ++jj)
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                const int j = 4 * tj2 + jj, k = 4 * tk2 + kk, f = flat_w2(j, k);
                m_w2[jj][kk] = f >= 0 ? a.adam_m[f] : 0.f;
                v_w2[jj][kk] = f >= 0 ? a.adam_v[f] : 0.f;
                if (f >= 0) { const float w = a.params[f]; W2[j * H + k] = w; W2t[k * H + w2t_col(k, j)] = w; }
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
    if (tid < AMAX) {
        const float sigma = expf(LOGSTD[tid]);
        SIG[tid] = 1.f / (sigma * sigma);
        SIG[AMAX + tid] = logf(sigma);
    }
    __syncthreads();
    cluster_sync_all();   // every CTA has zeroed its exchange slots before any peer writes into them

    const int ty = tid >> 4, tx = tid & 15;     // forward tile: rows 4ty.., cols 4tx..
    const int hr = tid >> 2, hq = tid & 3;      // head mapping: row hr, quarter hq
    const int warp = tid >> 5, lane = tid & 31;
    const int aw = a.is_discrete ? 1 : a.A;
    // widest cp.async the obs rows allow (row stride D floats, 16-byte aligned base)
    const int xvec = ((reinterpret_cast<uintptr_t>(a.obs) & 15) == 0) ? ((D % 4 == 0) ? 4 : (D % 2 == 0) ? 2 : 1) : 1;

    // ---- chunk pipeline: chunk q+1 of the minibatch-ordered streams is fetched by TMA bulk copies (one elected thread,
    // completion on an mbarrier) while chunk q is being computed.  No global latency is exposed to the step chain.
    struct Cursor { int epoch, mb, c0; };
    auto cur_valid = [&](const Cursor& c) { return c.epoch < a.n_epochs; };
    auto cur_bn = [&](const Cursor& c) { return min(a.B, a.N - c.mb * a.B); };
    auto cur_next = [&](Cursor c) {
        c.c0 += RB;
        if (c.c0 >= cur_bn(c)) { c.c0 = 0; if (++c.mb >= a.steps_per_epoch) { c.mb = 0; ++c.epoch; } }
        return c;
    };
    const int AP = a.AP;
    auto fetch_chunk = [&](const Cursor& c, int buf) {   // called by thread 0 only
        const size_t p = (size_t)c.epoch * a.N + (size_t)c.mb * a.B + c.c0;   // streams carry RB rows of tail padding
        const uint32_t xb = RB * DP * 4, sb = RB * 8 * 4, ab = (role == 0) ? RB * AP * 4 : 0;
        fence_proxy_async();   // earlier generic-proxy accesses to this buffer are ordered before the async-proxy writes
        mbar_expect_tx(&BAR[buf], xb + sb + ab);
        bulk_g2s(X + buf * RB * DP, a.xs + p * DP, xb, &BAR[buf]);
        bulk_g2s(ROWF + buf * RB * 8, a.ss + p * 8, sb, &BAR[buf]);
        if (role == 0) bulk_g2s(ACT + buf * RB * AMAX, a.as + p * AP, ab, &BAR[buf]);
    };

    Cursor cur = {0, 0, 0};
    int q = 0;                                   // running chunk counter (buffer parity / mbarrier phase)
    if (tid == 0) {
        mbar_init(&BAR[0], 1);
        mbar_init(&BAR[1], 1);
        fence_mbar_init();
    }
    __syncthreads();
    if (working && tid == 0) fetch_chunk(cur, 0);

    int step = 0, early_stop_epoch = a.n_epochs;
    double epoch_kl_sum = 0.0;
    bool stop_all = false;
    const bool timed = (a.timing != nullptr) && tid == 0 && working;
    long long tmark = clock64();
    unsigned long long tacc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) tacc[i] = 0;
#define ICRL_MARK(i)                                   \
    if (timed) {                                       \
        const long long now__ = clock64();             \
        tacc[i] += (unsigned long long)(now__ - tmark); \
        tmark = now__;                                 \
    }

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
            const float adv_mean_r = a.advstats[step * 8 + 0], adv_std_r = a.advstats[step * 8 + 1],
                        adv_mean_c = a.advstats[step * 8 + 2];
            const float adam_inv_bc2_sqrt = a.advstats[step * 8 + 3], adam_neg_step = a.advstats[step * 8 + 4];

            if (working) {
                for (int c0 = 0; c0 < Bn; c0 += RB, ++q) {
                    const int rows = min(RB, Bn - c0);
                    const int buf = q & 1;
                    const float* Xc = X + buf * RB * DP;
                    float* Rc = ROWF + buf * RB * 8;
                    const float* Ac = ACT + buf * RB * AMAX;   // rows at stride AP (as the stream stores them)
                    mbar_wait(&BAR[buf], (uint32_t)((q >> 1) & 1));
                    __syncthreads();          // chunk q landed, and everybody is done with chunk q-1's buffers
                    ICRL_MARK(0)
                    {
                        const Cursor nxt = cur_next(cur);
                        if (tid == 0 && cur_valid(nxt)) fetch_chunk(nxt, buf ^ 1);
                        cur = nxt;
                    }
                    ICRL_MARK(1)
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
                            *reinterpret_cast<float4*>(H1 + (4 * ty + i) * LDH + 4 * tx) =
                                make_float4(tanhf(acc[i][0]), tanhf(acc[i][1]), tanhf(acc[i][2]), tanhf(acc[i][3]));
                    }
                    __syncthreads();
                    ICRL_MARK(2)
                    // ---- forward layer 2: H2 = tanh(H1 W2^T + b2)
                    {
                        float acc[4][4];
#pragma unroll
                        for (int i = 0; i < 4; ++i)
#pragma unroll
                            for (int j = 0; j < 4; ++j) acc[i][j] = B2[4 * tx + j];
                        gemm_tile_4x4<H, true>(acc, H1, LDH, W2t, H, ty, tx);
#pragma unroll
                        for (int i = 0; i < 4; ++i)
                            *reinterpret_cast<float4*>(H2 + (4 * ty + i) * LDH + 4 * tx) =
                                make_float4(tanhf(acc[i][0]), tanhf(acc[i][1]), tanhf(acc[i][2]), tanhf(acc[i][3]));
                    }
                    __syncthreads();

                    ICRL_MARK(3)
                    // ---- heads + losses + d(loss)/d(head output).  4 threads per row (hr, hq).
                    if (role == 0) {
                        // action head: outputs d = hq, hq+4, hq+8, hq+12
                        float out[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const int d = hq + 4 * u;
                            float acc = 0.f;
                            if (d < a.A) {
                                const float4* hrow = reinterpret_cast<const float4*>(H2 + hr * LDH);
                                const float4* wrow = reinterpret_cast<const float4*>(HW + d * WA_LD);
                                float p0 = HB[d], p1 = 0.f, p2 = 0.f, p3 = 0.f;     // 4 independent chains
#pragma unroll
                                for (int k = 0; k < H / 4; ++k) {
                                    const float4 h = hrow[k], w = wrow[k];
                                    p0 = fmaf(h.x, w.x, p0); p1 = fmaf(h.y, w.y, p1);
                                    p2 = fmaf(h.z, w.z, p2); p3 = fmaf(h.w, w.w, p3);
                                }
                                acc = (p0 + p1) + (p2 + p3);
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
                                    const float inv_var = SIG[d], log_scale = SIG[AMAX + d];
                                    const float diff = Ac[hr * AP + d] - out[u];
                                    lp += -(diff * diff) * (0.5f * inv_var) - log_scale - LOG_SQRT_2PI;
                                    en += HALF_LOG_2PI_PLUS_HALF + log_scale;
                                    dcoef[u] = diff * inv_var;
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
                            const int ai = (int)Ac[hr * AP + 0];
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
                        if (hq == 0) Rc[hr * 8 + 7] = g;
                    } else {
                        // value head: partial dot over k in [16hq, 16hq+16)
                        float acc = 0.f;
                        const float4* hrow = reinterpret_cast<const float4*>(H2 + hr * LDH + 16 * hq);
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
                        const float target = Rc[hr * 8 + (role == 1 ? 3 : 5)], oldv = Rc[hr * 8 + (role == 1 ? 4 : 6)];
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

                    ICRL_MARK(4)
                    // ---- head weight / bias / log_std gradients (rows beyond `rows` carry zero dmean: loops run over RB)
                    if (hd < AOUT) {
                        float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
#pragma unroll 8
                        for (int r = 0; r < RB; ++r) {
                            const float dm = DMEAN[r * AMAX + hd];
                            const float4 h = *reinterpret_cast<const float4*>(H2 + r * LDH + 4 * hk4);
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
                        const float inv_var = SIG[s_idx];
                        float acc = 0.f;
#pragma unroll 8
                        for (int r = 0; r < RB; ++r) {
                            const float diff = Ac[r * AP + s_idx] - MU[r * AMAX + s_idx];
                            acc = fmaf(Rc[r * 8 + 7], diff * diff * inv_var - 1.f, acc);
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
                        const float4* hrow = reinterpret_cast<const float4*>(H2 + hr * LDH + 16 * hq);
                        float4* drow = reinterpret_cast<float4*>(DH + hr * LDH + 16 * hq);
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const float4 h = hrow[k];
                            drow[k] = make_float4(dloc[4 * k + 0] * (1.f - h.x * h.x), dloc[4 * k + 1] * (1.f - h.y * h.y),
                                                  dloc[4 * k + 2] * (1.f - h.z * h.z), dloc[4 * k + 3] * (1.f - h.w * h.w));
                        }
                    }
                    __syncthreads();

                    ICRL_MARK(5)
                    // ---- dW2 += dH2pre^T H1 ; db2 ; dH1pre = (dH2pre W2) * (1 - H1^2) -> written over H2
                    outer_tile_4x4(g_w2, DH, LDH, H1, LDH, tj2, tk2);
                    if (s_kind == 1) {
                        float acc = 0.f;
#pragma unroll 8
                        for (int r = 0; r < RB; ++r) acc += DH[r * LDH + s_idx];
                        g_s += acc;
                    }
                    {
                        float acc[4][4];
#pragma unroll
                        for (int i = 0; i < 4; ++i)
#pragma unroll
                            for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
                        gemm_tile_4x4<H>(acc, DH, LDH, W2, H, ty, tx);   // sum_j dH2[r][j] * W2[j][k]  (W2 row-major == "k-major" in j)
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const float4 h = *reinterpret_cast<const float4*>(H1 + (4 * ty + i) * LDH + 4 * tx);
                            *reinterpret_cast<float4*>(H2 + (4 * ty + i) * LDH + 4 * tx) =
                                make_float4(acc[i][0] * (1.f - h.x * h.x), acc[i][1] * (1.f - h.y * h.y),
                                            acc[i][2] * (1.f - h.z * h.z), acc[i][3] * (1.f - h.w * h.w));
                        }
                    }
                    __syncthreads();
                    ICRL_MARK(6)
                    // ---- dW1 += dH1pre^T X ; db1
#pragma unroll
                    for (int n = 0; n < NT1; ++n) {
                        const int t = tid + NTH * n;
                        if (t < n_w1_tiles) outer_tile_4x4(g_w1[n], H2, LDH, Xc, DP, t & 15, t >> 4);
                    }
                    if (s_kind == 0) {
                        float acc = 0.f;
#pragma unroll 8
                        for (int r = 0; r < RB; ++r) acc += H2[r * LDH + s_idx];
                        g_s += acc;
                    }
                    ICRL_MARK(7)
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

            ICRL_MARK(8)
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
            ICRL_MARK(9)
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
                AdamConsts ac;
                ac.one_minus_b1 = (float)(1.0 - a.beta1);
                ac.b2 = (float)a.beta2;
                ac.one_minus_b2 = (float)(1.0 - a.beta2);
                ac.inv_bc2_sqrt = adam_inv_bc2_sqrt;
                ac.eps = (float)a.adam_eps;
                ac.neg_step_size = adam_neg_step;
                {
                    // W2 tile (rows 4tj2.., cols 4tk2..): 128-bit reads/writes of the row-major copy, 128-bit swizzled
                    // writes of the transposed copy.  Padding entries (j >= h1 or k >= h0) keep g = m = v = 0 and stay 0.
                    float pw[4][4];
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj) {
                        const float4 w4 = *reinterpret_cast<const float4*>(W2 + (4 * tj2 + jj) * H + 4 * tk2);
                        pw[jj][0] = adam_update(w4.x, g_w2[jj][0] * clip_coef, m_w2[jj][0], v_w2[jj][0], ac);
                        pw[jj][1] = adam_update(w4.y, g_w2[jj][1] * clip_coef, m_w2[jj][1], v_w2[jj][1], ac);
                        pw[jj][2] = adam_update(w4.z, g_w2[jj][2] * clip_coef, m_w2[jj][2], v_w2[jj][2], ac);
                        pw[jj][3] = adam_update(w4.w, g_w2[jj][3] * clip_coef, m_w2[jj][3], v_w2[jj][3], ac);
                        *reinterpret_cast<float4*>(W2 + (4 * tj2 + jj) * H + 4 * tk2) =
                            make_float4(pw[jj][0], pw[jj][1], pw[jj][2], pw[jj][3]);
                    }
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk)
                        *reinterpret_cast<float4*>(W2t + (4 * tk2 + kk) * H + 4 * (tj2 ^ tk2)) =
                            make_float4(pw[0][kk], pw[1][kk], pw[2][kk], pw[3][kk]);
                }
#pragma unroll
                for (int n = 0; n < NT1; ++n) {
                    const int t = tid + NTH * n, tk1 = t >> 4, tj1 = t & 15;
                    if (t < n_w1_tiles) {
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk) {
                            float4* wp = reinterpret_cast<float4*>(W1t + (4 * tk1 + kk) * H + 4 * tj1);
                            const float4 w4 = *wp;
                            *wp = make_float4(adam_update(w4.x, g_w1[n][0][kk] * clip_coef, m_w1[n][0][kk], v_w1[n][0][kk], ac),
                                              adam_update(w4.y, g_w1[n][1][kk] * clip_coef, m_w1[n][1][kk], v_w1[n][1][kk], ac),
                                              adam_update(w4.z, g_w1[n][2][kk] * clip_coef, m_w1[n][2][kk], v_w1[n][2][kk], ac),
                                              adam_update(w4.w, g_w1[n][3][kk] * clip_coef, m_w1[n][3][kk], v_w1[n][3][kk], ac));
                        }
                    }
                }
                if (hd < AOUT) {
                    float4* wp = reinterpret_cast<float4*>(HW + hd * WA_LD + 4 * hk4);
                    const float4 w4 = *wp;
                    *wp = make_float4(adam_update(w4.x, g_hw[0] * clip_coef, m_hw[0], v_hw[0], ac),
                                      adam_update(w4.y, g_hw[1] * clip_coef, m_hw[1], v_hw[1], ac),
                                      adam_update(w4.z, g_hw[2] * clip_coef, m_hw[2], v_hw[2], ac),
                                      adam_update(w4.w, g_hw[3] * clip_coef, m_hw[3], v_hw[3], ac));
                }
                if (s_kind >= 0 && flat_scalar() >= 0) {
                    float* slot = s_kind == 0 ? &B1[s_idx] : s_kind == 1 ? &B2[s_idx] : s_kind == 2 ? &HB[s_idx] : &LOGSTD[s_idx];
                    const float pnew = adam_update(*slot, g_s * clip_coef, m_s, v_s, ac);
                    *slot = pnew;
                    if (s_kind == 3) {
                        const float sigma = expf(pnew);
                        SIG[s_idx] = 1.f / (sigma * sigma);
                        SIG[AMAX + s_idx] = logf(sigma);
                    }
                }
            }
            ICRL_MARK(10)
            if (stop_rx == 1.f) { early_stop_epoch = epoch; stop_all = true; }
            if (stop_rx == 2.f) { stop_all = true; }
            // (the next chunk's leading __syncthreads orders these shared-memory weight updates before their first use)
        }  // minibatches
    }      // epochs
    if (working && cur_valid(cur)) mbar_wait(&BAR[q & 1], (uint32_t)((q >> 1) & 1));   // drain the in-flight prefetch
    __syncthreads();
    if (timed) {
        for (int i = 0; i < 16; ++i) a.timing[role * 16 + i] = tacc[i];
    }
#undef ICRL_MARK

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
        const long long n_rows = (long long)a.n_epochs * a.N, n_alloc = n_rows + icrl::RB;
        const int aw = a.is_discrete ? 1 : a.A;
        a.AP = (aw + 3) / 4 * 4;
        void *xs, *as, *ss, *advstats;
        if ((rc = icrl::device_scratch(icrl::SLOT_PPO0, (size_t)n_alloc * a.DP * 4, &xs))) return rc;
        if ((rc = icrl::device_scratch(icrl::SLOT_PPO1, (size_t)total_steps * 8 * sizeof(float), &advstats))) return rc;
        if ((rc = icrl::device_scratch(icrl::SLOT_PPO2, (size_t)n_alloc * a.AP * 4, &as))) return rc;
        if ((rc = icrl::device_scratch(icrl::SLOT_PPO3, (size_t)n_alloc * 8 * 4, &ss))) return rc;
        const long long blocks = (n_alloc + 7) / 8;
        const int grid = (int)(blocks < 8LL * icrl::sm_count() ? blocks : 8LL * icrl::sm_count());
        icrl::ppo_gather_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(a, (float*)xs, (float*)as, (float*)ss, n_rows, n_alloc);
        ICRL_LAUNCH_CHECK();
        icrl::ppo_stats_kernel<<<total_steps, 128, 0, (cudaStream_t)stream>>>((const float*)ss, (float*)advstats, a.N, a.B,
                                                                             a.steps_per_epoch, a.beta1, a.beta2, a.lr,
                                                                             a.step_before);
        ICRL_LAUNCH_CHECK();
        a.xs = (const float*)xs; a.as = (const float*)as; a.ss = (const float*)ss;
        a.advstats = (const float*)advstats;
    }
    static unsigned long long* timing_dev = nullptr;
    const bool want_timing = getenv("ICRL_PPO_TIMING") != nullptr;
    if (want_timing && !timing_dev) cudaMalloc(&timing_dev, 48 * sizeof(unsigned long long));
    a.timing = want_timing ? timing_dev : nullptr;
    const int n_tiles = 16 * (a.DP / 4);
    const int nt1 = (n_tiles + icrl::NTH - 1) / icrl::NTH;
    switch (nt1) {
        case 1: rc = icrl::launch_ppo<1>(a, (cudaStream_t)stream); break;
        case 2: rc = icrl::launch_ppo<2>(a, (cudaStream_t)stream); break;
        case 3: rc = icrl::launch_ppo<3>(a, (cudaStream_t)stream); break;
        default:
            icrl::set_error("obs_dim %d too large for the PPO kernel (max 192)", a.D);
            return ICRL_EUNSUPPORTED;
    }
    if (rc == 0 && want_timing) {   // profiling aid: per-phase cycles of thread 0 of each trunk CTA (synchronises!)
        unsigned long long h[48];
        cudaStreamSynchronize((cudaStream_t)stream);
        cudaMemcpy(h, timing_dev, sizeof(h), cudaMemcpyDeviceToHost);
        const char* names[11] = {"wait+sync", "prefetch", "L1", "L2", "head", "headgrad+dH2", "dW2+dH1", "dW1", "reduce+stats",
                                 "xchg+cluster", "adam"};
        const double steps = (double)(a.max_steps > 0 ? a.max_steps : a.n_epochs * a.steps_per_epoch);
        for (int r = 0; r < 3; ++r) {
            fprintf(stderr, "[ppo timing] role %d cycles/step:", r);
            double tot = 0;
            for (int i = 0; i < 11; ++i) { fprintf(stderr, " %s=%.0f", names[i], h[r * 16 + i] / steps); tot += h[r * 16 + i] / steps; }
            fprintf(stderr, " total=%.0f\n", tot);
        }
    }
    return rc;
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
