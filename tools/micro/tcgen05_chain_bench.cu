// Microbenchmark: K4's dependent forward chain on the 5th-generation tensor cores (tcgen05 + TMEM) against the mma.sync path the
// kernel ships with.  Question: does a chain of tiny DEPENDENT GEMMs
//     H1^T[64 x 32] = tanh(W1[64 x K1] . X^T[K1 x 32] + b1)      K1 = 24 (HalfCheetah) or 120 (Ant)
//     H2^T[64 x 32] = tanh(W2[64 x 64] . H1^T[64 x 32] + b2)
// get cheaper when each layer is  3 x tcgen05.mma.kind::tf32 (3xTF32 split, M = 64 hidden units, N = 32 rows) per k-step ->
// tcgen05.commit -> mbarrier -> tcgen05.ld -> tanh / split epilogue -> st.shared planes -> fence.proxy.async -> next layer,
// compared with 8 warps of mma.sync.m16n8k8 with register accumulators (what k4_ppo_lag.cu does)?
// One CTA, 256 threads, operands resident in shared memory (as in the kernel).  Prints cycles per chained pair of layers for
// both paths and checks both against a float64 CPU evaluation.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tcgen05_chain_bench tcgen05_chain_bench.cu && ./tcgen05_chain_bench
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include <cuda_runtime.h>

constexpr int H = 64;        // hidden units (M of the transposed GEMMs)
constexpr int NB = 32;       // batch rows per CTA (N)
constexpr int NTHR = 256;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
    hi = __uint_as_float(__float_as_uint(x) & 0xffffe000u);
    lo = x - hi;
}

// ------------------------------------------------------------------------------------------------ tcgen05 path
// K-major canonical layout without swizzle (cute::UMMA::LayoutType::SWIZZLE_NONE): element (row r, k) of an operand with R rows at
//   byte (k / 4) * LBO + r * 16 + (k % 4) * 4,   LBO = R * 16 + 16 (one 16-byte pad per K chunk: conflict-free epilogue stores),
// i.e. 8-row core matrices of 8 x 16 bytes, SBO = 128 bytes between 8-row groups, LBO between the two K chunks of an instruction.
__host__ __device__ constexpr int lbo_bytes(int rows) { return rows * 16 + 16; }
__host__ __device__ constexpr int plane_bytes(int rows, int K) { return (K / 4) * lbo_bytes(rows); }
__device__ __forceinline__ int plane_off(int rows, int r, int k) { return ((k >> 2) * lbo_bytes(rows) + r * 16 + (k & 3) * 4) >> 2; }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, int rows) {
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)(lbo_bytes(rows) >> 4) << 16) | ((uint64_t)(128 >> 4) << 32) |
           (1ull << 46);                       // version 1 (Blackwell), base offset 0, SWIZZLE_NONE
}
// instruction descriptor: D = F32, A = B = TF32, both K-major, N >> 3 at bit 17, M >> 4 at bit 24
constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NB >> 3) << 17) | ((uint32_t)(H >> 4) << 24);

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(IDESC), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}\n" ::"r"(
            smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// 16 consecutive accumulator columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

struct Result {
    long long cyc_total, cyc_issue1, cyc_wait1, cyc_epi1, cyc_issue2, cyc_wait2, cyc_epi2;
};

// shared memory (floats): W1hi W1lo [H x K1] | W2hi W2lo [H x 64] | Xhi Xlo [NB x K1] | H1hi H1lo [NB x 64] | H2 [NB x 64] | b1 b2
// SPLIT: the three products of the 3xTF32 split accumulate into three separate TMEM tiles (independent tensor-pipe chains,
// summed in the epilogue) instead of one tile (one dependent chain of 3 * K / 8 instructions).
template <int K1, bool SPLIT>
__global__ void __launch_bounds__(NTHR, 1) chain_tcgen05(const float* __restrict__ W1, const float* __restrict__ b1,
                                                         const float* __restrict__ W2, const float* __restrict__ b2,
                                                         const float* __restrict__ X, float* __restrict__ H2out, int iters,
                                                         Result* res) {
    extern __shared__ __align__(1024) unsigned char smem[];
    constexpr int W1B = plane_bytes(H, K1), W2B = plane_bytes(H, H), XB = plane_bytes(NB, K1), HB = plane_bytes(NB, H);
    float* W1hi = reinterpret_cast<float*>(smem);
    float* W1lo = reinterpret_cast<float*>(smem + W1B);
    float* W2hi = reinterpret_cast<float*>(smem + 2 * W1B);
    float* W2lo = reinterpret_cast<float*>(smem + 2 * W1B + W2B);
    float* Xhi = reinterpret_cast<float*>(smem + 2 * W1B + 2 * W2B);
    float* Xlo = reinterpret_cast<float*>(smem + 2 * W1B + 2 * W2B + XB);
    float* H1hi = reinterpret_cast<float*>(smem + 2 * W1B + 2 * W2B + 2 * XB);
    float* H1lo = reinterpret_cast<float*>(smem + 2 * W1B + 2 * W2B + 2 * XB + HB);
    float* H2s = reinterpret_cast<float*>(smem + 2 * W1B + 2 * W2B + 2 * XB + 2 * HB);         // [NB][H + 4]
    float* B1 = H2s + NB * (H + 4);
    float* B2 = B1 + H;
    __shared__ uint64_t bar[2];
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    for (int i = tid; i < H * K1; i += NTHR) {
        const int m = i / K1, k = i - m * K1;
        float hi, lo;
        split_tf32(W1[i], hi, lo);
        W1hi[plane_off(H, m, k)] = hi; W1lo[plane_off(H, m, k)] = lo;
    }
    for (int i = tid; i < H * H; i += NTHR) {
        const int m = i / H, k = i - m * H;
        float hi, lo;
        split_tf32(W2[i], hi, lo);
        W2hi[plane_off(H, m, k)] = hi; W2lo[plane_off(H, m, k)] = lo;
    }
    for (int i = tid; i < NB * K1; i += NTHR) {
        const int r = i / K1, k = i - r * K1;
        float hi, lo;
        split_tf32(X[i], hi, lo);
        Xhi[plane_off(NB, r, k)] = hi; Xlo[plane_off(NB, r, k)] = lo;
    }
    if (tid < H) { B1[tid] = b1[tid]; B2[tid] = b2[tid]; }
    if (tid == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {      // 256 TMEM columns: two layers x up to three 32-column accumulators
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(&tmem_base_s)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_s;
    const uint32_t acc1 = tmem_base, acc2 = tmem_base + 96;
    constexpr uint32_t S1 = SPLIT ? 32u : 0u, S2 = SPLIT ? 64u : 0u;      // column offsets of the lo*hi / hi*lo tiles
    // epilogue mapping (M = 64, one CTA): hidden unit m = 16 q + i lives in TMEM lane 32 q + i (q = warp % 4, i = lane < 16);
    // warps 0-3 take batch rows 0..15, warps 4-7 rows 16..31 of that lane
    const int q = warp & 3, unit = 16 * q + lane, col0 = (warp >> 2) * 16;
    const bool epi = lane < 16;
    const uint32_t lane_addr = ((uint32_t)(32 * q) << 16);

    const uint64_t d_w1hi = make_desc(smem_u32(W1hi), H), d_w1lo = make_desc(smem_u32(W1lo), H);
    const uint64_t d_w2hi = make_desc(smem_u32(W2hi), H), d_w2lo = make_desc(smem_u32(W2lo), H);
    const uint64_t d_xhi = make_desc(smem_u32(Xhi), NB), d_xlo = make_desc(smem_u32(Xlo), NB);
    const uint64_t d_h1hi = make_desc(smem_u32(H1hi), NB), d_h1lo = make_desc(smem_u32(H1lo), NB);
    long long t_total = 0, t_i1 = 0, t_w1 = 0, t_e1 = 0, t_i2 = 0, t_w2 = 0, t_e2 = 0;
    for (int it = -8; it < iters; ++it) {          // 8 warm-up rounds
        const uint32_t ph = (uint32_t)(it + 8) & 1u;
        const long long t0 = clock64();
        if (tid == 0) {
#pragma unroll
            for (int k0 = 0; k0 < K1; k0 += 8) {      // descriptors: loop-invariant bases + a compile-time start-address step
                const uint64_t sa = (uint64_t)(((k0 >> 2) * lbo_bytes(H)) >> 4), sb = (uint64_t)(((k0 >> 2) * lbo_bytes(NB)) >> 4);
                umma_tf32(acc1 + S1, d_w1lo + sa, d_xhi + sb, k0 > 0);
                umma_tf32(acc1 + S2, d_w1hi + sa, d_xlo + sb, SPLIT ? (k0 > 0) : 1u);
                umma_tf32(acc1, d_w1hi + sa, d_xhi + sb, SPLIT ? (k0 > 0) : 1u);
            }
            umma_commit(&bar[0]);
        }
        const long long t1 = clock64();
        mbar_wait(&bar[0], ph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const long long t2 = clock64();
        {
            float v[16];
            tmem_ld16(acc1 + lane_addr + col0, v);
            if (SPLIT) {
                float u1[16], u2[16];
                tmem_ld16(acc1 + S1 + lane_addr + col0, u1);
                tmem_ld16(acc1 + S2 + lane_addr + col0, u2);
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] += u1[j] + u2[j];
            }
            if (epi) {
                const float b = B1[unit];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    float hi, lo;
                    split_tf32(tanhf(v[j] + b), hi, lo);
                    const int o = plane_off(NB, col0 + j, unit);
                    H1hi[o] = hi; H1lo[o] = lo;
                }
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        const long long t3 = clock64();
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
            for (int k0 = 0; k0 < H; k0 += 8) {
                const uint64_t sa = (uint64_t)(((k0 >> 2) * lbo_bytes(H)) >> 4), sb = (uint64_t)(((k0 >> 2) * lbo_bytes(NB)) >> 4);
                umma_tf32(acc2 + S1, d_w2lo + sa, d_h1hi + sb, k0 > 0);
                umma_tf32(acc2 + S2, d_w2hi + sa, d_h1lo + sb, SPLIT ? (k0 > 0) : 1u);
                umma_tf32(acc2, d_w2hi + sa, d_h1hi + sb, SPLIT ? (k0 > 0) : 1u);
            }
            umma_commit(&bar[1]);
        }
        const long long t4 = clock64();
        mbar_wait(&bar[1], ph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const long long t5 = clock64();
        {
            float v[16];
            tmem_ld16(acc2 + lane_addr + col0, v);
            if (SPLIT) {
                float u1[16], u2[16];
                tmem_ld16(acc2 + S1 + lane_addr + col0, u1);
                tmem_ld16(acc2 + S2 + lane_addr + col0, u2);
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] += u1[j] + u2[j];
            }
            if (epi) {
                const float b = B2[unit];
#pragma unroll
                for (int j = 0; j < 16; ++j) H2s[(col0 + j) * (H + 4) + unit] = tanhf(v[j] + b);
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        const long long t6 = clock64();
        if (it >= 0) {
            t_total += t6 - t0; t_i1 += t1 - t0; t_w1 += t2 - t1; t_e1 += t3 - t2; t_i2 += t4 - t3; t_w2 += t5 - t4; t_e2 += t6 - t5;
        }
    }
    for (int i = tid; i < NB * H; i += NTHR) H2out[i] = H2s[(i / H) * (H + 4) + (i % H)];
    if (tid == 0) *res = Result{t_total, t_i1, t_w1, t_e1, t_i2, t_w2, t_e2};
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem_base) : "memory");
}

// ------------------------------------------------------------------------------------------------ mma.sync path (as K4 ships)
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void split_u(float x, uint32_t& hi, uint32_t& lo) {
    hi = __float_as_uint(x) & 0xffffe000u;
    lo = __float_as_uint(x - __uint_as_float(hi));
}
// C[16 x 16 per warp] += A[rows x K] . W[cols x K]^T, 3xTF32, operands row-major in shared memory with strides == 4 (mod 32)
template <int K>
__device__ __forceinline__ void warp_gemm(float (&c)[2][4], const float* A, int lda, const float* W, int ldw, int g, int t) {
#pragma unroll
    for (int k0 = 0; k0 < K; k0 += 8) {
        uint32_t ahi[4], alo[4];
        const float* ap = A + g * lda + k0 + t;
        split_u(ap[0], ahi[0], alo[0]);
        split_u(ap[8 * lda], ahi[1], alo[1]);
        split_u(ap[4], ahi[2], alo[2]);
        split_u(ap[8 * lda + 4], ahi[3], alo[3]);
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            uint32_t bhi[2], blo[2];
            const float* bp = W + (8 * i + g) * ldw + k0 + t;
            split_u(bp[0], bhi[0], blo[0]);
            split_u(bp[4], bhi[1], blo[1]);
            mma_tf32(c[i], alo, bhi);
            mma_tf32(c[i], ahi, blo);
            mma_tf32(c[i], ahi, bhi);
        }
    }
}

template <int K1>
__global__ void __launch_bounds__(NTHR, 1) chain_mma_sync(const float* __restrict__ W1, const float* __restrict__ b1,
                                                          const float* __restrict__ W2, const float* __restrict__ b2,
                                                          const float* __restrict__ X, float* __restrict__ H2out, int iters,
                                                          Result* res) {
    constexpr int LDX = (K1 / 32) * 32 + 4 + ((K1 % 32) > 4 ? 32 : 0), LDH = 68;
    extern __shared__ __align__(1024) unsigned char smem[];
    float* sW1 = reinterpret_cast<float*>(smem);     // [H][LDX]
    float* sW2 = sW1 + H * LDX;                      // [H][LDH]
    float* sX = sW2 + H * LDH;                       // [NB][LDX]
    float* sH1 = sX + NB * LDX;                      // [NB][LDH]
    float* sH2 = sH1 + NB * LDH;
    float* B1 = sH2 + NB * LDH;
    float* B2 = B1 + H;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    for (int i = tid; i < H * K1; i += NTHR) sW1[(i / K1) * LDX + i % K1] = W1[i];
    for (int i = tid; i < H * H; i += NTHR) sW2[(i / H) * LDH + i % H] = W2[i];
    for (int i = tid; i < NB * K1; i += NTHR) sX[(i / K1) * LDX + i % K1] = X[i];
    if (tid < H) { B1[tid] = b1[tid]; B2[tid] = b2[tid]; }
    __syncthreads();
    const int mt = warp & 1, ng = warp >> 1;          // warp tile: rows 16 mt.., hidden units 16 ng..
    long long t_total = 0, t_l1 = 0, t_l2 = 0;
    for (int it = -8; it < iters; ++it) {
        const long long t0 = clock64();
        float acc[2][4];
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const float2 b = *reinterpret_cast<const float2*>(B1 + 16 * ng + 8 * i + 2 * t);
            acc[i][0] = b.x; acc[i][1] = b.y; acc[i][2] = b.x; acc[i][3] = b.y;
        }
        warp_gemm<K1>(acc, sX + 16 * mt * LDX, LDX, sW1 + 16 * ng * LDX, LDX, g, t);
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            float* p = sH1 + (16 * mt + g) * LDH + 16 * ng + 8 * i + 2 * t;
            *reinterpret_cast<float2*>(p) = make_float2(tanhf(acc[i][0]), tanhf(acc[i][1]));
            *reinterpret_cast<float2*>(p + 8 * LDH) = make_float2(tanhf(acc[i][2]), tanhf(acc[i][3]));
        }
        __syncthreads();
        const long long t1 = clock64();
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const float2 b = *reinterpret_cast<const float2*>(B2 + 16 * ng + 8 * i + 2 * t);
            acc[i][0] = b.x; acc[i][1] = b.y; acc[i][2] = b.x; acc[i][3] = b.y;
        }
        warp_gemm<H>(acc, sH1 + 16 * mt * LDH, LDH, sW2 + 16 * ng * LDH, LDH, g, t);
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            float* p = sH2 + (16 * mt + g) * LDH + 16 * ng + 8 * i + 2 * t;
            *reinterpret_cast<float2*>(p) = make_float2(tanhf(acc[i][0]), tanhf(acc[i][1]));
            *reinterpret_cast<float2*>(p + 8 * LDH) = make_float2(tanhf(acc[i][2]), tanhf(acc[i][3]));
        }
        __syncthreads();
        const long long t2 = clock64();
        if (it >= 0) { t_total += t2 - t0; t_l1 += t1 - t0; t_l2 += t2 - t1; }
    }
    for (int i = tid; i < NB * H; i += NTHR) H2out[i] = sH2[(i / H) * LDH + i % H];
    if (tid == 0) *res = Result{t_total, t_l1, 0, 0, t_l2, 0, 0};
}

// ------------------------------------------------------------------------------------------------ host
template <int K1>
static int run(const char* name) {
    std::vector<float> W1(H * K1), b1(H), W2(H * H), b2(H), X(NB * K1);
    srand(7);
    auto rnd = [] { return (float)rand() / RAND_MAX * 2.f - 1.f; };
    for (auto& v : W1) v = rnd() * 0.4f;
    for (auto& v : W2) v = rnd() * 0.3f;
    for (auto& v : b1) v = rnd() * 0.1f;
    for (auto& v : b2) v = rnd() * 0.1f;
    for (auto& v : X) v = rnd() * 2.f;
    std::vector<double> ref(NB * H);
    for (int r = 0; r < NB; ++r) {
        double h1[H];
        for (int m = 0; m < H; ++m) {
            double s = b1[m];
            for (int k = 0; k < K1; ++k) s += (double)W1[m * K1 + k] * X[r * K1 + k];
            h1[m] = tanh(s);
        }
        for (int m = 0; m < H; ++m) {
            double s = b2[m];
            for (int k = 0; k < H; ++k) s += (double)W2[m * H + k] * h1[k];
            ref[r * H + m] = tanh(s);
        }
    }
    float *dW1, *db1, *dW2, *db2, *dX, *dH2;
    Result* dres;
    cudaMalloc(&dW1, W1.size() * 4); cudaMalloc(&db1, H * 4); cudaMalloc(&dW2, W2.size() * 4); cudaMalloc(&db2, H * 4);
    cudaMalloc(&dX, X.size() * 4); cudaMalloc(&dH2, NB * H * 4); cudaMalloc(&dres, sizeof(Result));
    cudaMemcpy(dW1, W1.data(), W1.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(db1, b1.data(), H * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dW2, W2.data(), W2.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(db2, b2.data(), H * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dX, X.data(), X.size() * 4, cudaMemcpyHostToDevice);
    const int iters = 2000;
    std::vector<float> out(NB * H);
    Result r;
    auto check = [&](const char* what) {
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("%s %s: CUDA error %s\n", name, what, cudaGetErrorString(e)); return 1e30; }
        cudaMemcpy(out.data(), dH2, out.size() * 4, cudaMemcpyDeviceToHost);
        cudaMemcpy(&r, dres, sizeof(r), cudaMemcpyDeviceToHost);
        double worst = 0;
        for (int i = 0; i < NB * H; ++i) worst = fmax(worst, fabs(out[i] - ref[i]));
        return worst;
    };
    for (int split = 0; split < 2; ++split) {
        constexpr int bytes = 2 * plane_bytes(H, K1) + 2 * plane_bytes(H, H) + 2 * plane_bytes(NB, K1) + 2 * plane_bytes(NB, H) +
                              (NB * (H + 4) + 2 * H) * 4;
        cudaFuncSetAttribute(chain_tcgen05<K1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
        cudaFuncSetAttribute(chain_tcgen05<K1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
        cudaMemset(dH2, 0, NB * H * 4);
        if (split) chain_tcgen05<K1, true><<<1, NTHR, bytes>>>(dW1, db1, dW2, db2, dX, dH2, iters, dres);
        else chain_tcgen05<K1, false><<<1, NTHR, bytes>>>(dW1, db1, dW2, db2, dX, dH2, iters, dres);
        const double err = check("tcgen05");
        printf("%s tcgen05 (3 x kind::tf32, M=64 N=32, %s): max |err| vs float64 %.2e | cycles per L1+L2 chain %.0f = "
               "L1 issue %.0f + commit wait %.0f + ld/tanh/split/store %.0f | L2 issue %.0f + commit wait %.0f + ld/tanh/store %.0f\n",
               name, split ? "three TMEM accumulators per layer (independent chains)" : "one TMEM accumulator per layer", err,
               (double)r.cyc_total / iters, (double)r.cyc_issue1 / iters, (double)r.cyc_wait1 / iters,
               (double)r.cyc_epi1 / iters, (double)r.cyc_issue2 / iters, (double)r.cyc_wait2 / iters, (double)r.cyc_epi2 / iters);
    }
    {
        constexpr int LDX = (K1 / 32) * 32 + 4 + ((K1 % 32) > 4 ? 32 : 0);
        constexpr int bytes = (H * LDX + H * 68 + NB * LDX + 2 * NB * 68 + 2 * H) * 4;
        cudaFuncSetAttribute(chain_mma_sync<K1>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
        cudaMemset(dH2, 0, NB * H * 4);
        chain_mma_sync<K1><<<1, NTHR, bytes>>>(dW1, db1, dW2, db2, dX, dH2, iters, dres);
        const double err = check("mma.sync");
        printf("%s mma.sync (3xTF32 m16n8k8, register accumulators, 8 warps): max |err| vs float64 %.2e | cycles per L1+L2 chain %.0f = "
               "L1 %.0f + L2 %.0f\n", name, err, (double)r.cyc_total / iters, (double)r.cyc_issue1 / iters, (double)r.cyc_issue2 / iters);
    }
    return 0;
}

int main() {
    run<24>("K1=24 (HalfCheetah)");
    run<120>("K1=120 (AntWall)");
    return 0;
}
