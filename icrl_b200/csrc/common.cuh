// Shared device/host helpers for the icrl_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/icrl_b200.h"

namespace icrl {

// ---------------------------------------------------------------- host-side error plumbing
void set_error(const char* fmt, ...);
void count_launch(int n = 1);

#define ICRL_CHECK_ARG(cond, ...)                 \
    do {                                          \
        if (!(cond)) {                            \
            ::icrl::set_error(__VA_ARGS__);       \
            return ICRL_EINVAL;                   \
        }                                         \
    } while (0)

#define ICRL_CUDA(call)                                                                         \
    do {                                                                                        \
        cudaError_t err__ = (call);                                                             \
        if (err__ != cudaSuccess) {                                                             \
            ::icrl::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(err__), __FILE__, __LINE__); \
            return (int)err__;                                                                  \
        }                                                                                       \
    } while (0)

#define ICRL_LAUNCH_CHECK()                                                                     \
    do {                                                                                        \
        cudaError_t err__ = cudaGetLastError();                                                 \
        if (err__ != cudaSuccess) {                                                             \
            ::icrl::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(err__), __FILE__, __LINE__); \
            return (int)err__;                                                                  \
        }                                                                                       \
        ::icrl::count_launch();                                                                 \
    } while (0)

int sm_count();

// grow-only device / pinned-host scratch, one slot per purpose (not thread safe: the learner is single threaded,
// as the reference's is -- SURVEY §8b "Threading").
enum Slot { SLOT_IN0 = 0, SLOT_IN1, SLOT_OUT0, SLOT_WORK0, SLOT_WORK1, SLOT_WORK2, SLOT_WORK3, SLOT_WORK4, SLOT_WORK5, SLOT_PPO0, SLOT_PPO1, SLOT_PPO2, SLOT_PPO3, SLOT_PPO4, SLOT_PPO5, SLOT_K5, SLOT_COUNT };
int device_scratch(Slot s, size_t bytes, void** ptr);
int pinned_scratch(Slot s, size_t bytes, void** ptr);

// ---------------------------------------------------------------- device helpers
#ifdef __CUDACC__

constexpr int kWarp = 32;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- mbarrier + 1-D bulk async copy (TMA, `cp.async.bulk` -> SASS UBLKCP): global -> shared::cta
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// bounded wait: false if the phase did not complete within `max_cycles` (the caller records the failure and moves on, so a
// lost peer ends the launch with an error code instead of hanging the GPU)
__device__ __forceinline__ bool mbar_wait_bounded(uint64_t* bar, uint32_t parity, long long max_cycles) {
    const long long t0 = clock64();
    for (;;) {
        uint32_t done;
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
        if (done) return true;
        if (clock64() - t0 > max_cycles) return false;
    }
}
// bytes must be a multiple of 16; src and dst 16-byte aligned.
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// ---- tensor-core helpers (mma.sync m16n8k8, TF32 operands, fp32 accumulate) shared by K1 and K4
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
    // hi = x with the 13 low mantissa bits cleared (one LOP3; cvt.rna.tf32 expands to a multi-instruction sequence and
    // was 29% of all issued instructions), lo = x - hi exactly.  x = hi + lo holds exactly, hi is a valid TF32 value and
    // the tensor core keeps 11 significant bits of lo (<= 2^-10 |x|): ~2^-20 relative per product, fp32-class.
    hi = __float_as_uint(x) & 0xffffe000u;
    lo = __float_as_uint(x - __uint_as_float(hi));
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile(
        "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// ---- warp-level 3xTF32 GEMM on mma.sync m16n8k8 (shared by K4's five GEMMs per chunk and K2's weight gradients)
// One warp: C[16 x (8 per n-tile)] += A[16 x K] * B[K x .] with 3xTF32.
//   A element (m, k) at A[m * sam + k * sak]   (A already points at the warp's first row / column)
//   B element (k, n) at B[k * sbk + n * sbn]
// n-tile i (i < NT, skipped when ncol0 + i * nstep >= nmax) covers columns ncol0 + i * nstep .. +7.
// Fragment layout (PTX ISA, m16n8k8 .tf32): g = lane >> 2, t = lane & 3;
//   a0 (g, t) a1 (g+8, t) a2 (g, t+4) a3 (g+8, t+4);  b0 (k = t, n = g) b1 (k = t+4, n = g);
//   c0 (g, 2t) c1 (g, 2t+1) c2 (g+8, 2t) c3 (g+8, 2t+1)
// GUARD: skip n-tiles at or beyond nmax (a per-tile branch: it keeps ptxas from interleaving the independent MMA chains of
// different n-tiles, so it is only instantiated where a tile can really fall outside the operand)
// KPERM: the summation index of a k-step is permuted (slot t <-> k0 + 2t, slot t + 4 <-> k0 + 2t + 1, on BOTH operands, which
// leaves the product unchanged).  For operands whose K runs along shared-memory ROWS of stride == 4 (mod 32) this turns the
// 2-way bank conflicts of the natural order (bank 4t + g) into conflict-free loads (banks 8t + g and 8t + 4 + g).
template <int NT, int KC = 0, bool GUARD = false, bool KPERM = false>
__device__ __forceinline__ void warp_gemm_3xtf32(float (&c)[NT][4], const float* __restrict__ A, int sam, int sak,
                                                 const float* __restrict__ B, int sbk, int sbn, int K, int ncol0, int nstep,
                                                 int nmax, int g, int t) {
    const int kend = KC > 0 ? KC : K;   // KC > 0: compile-time K, fully unrolled so fragment loads run ahead of the MMAs
#pragma unroll (KC > 0 ? KC / 8 : 2)
    for (int k0 = 0; k0 < kend; k0 += 8) {
        uint32_t ahi[4], alo[4];
        {
            const float* ap = A + (k0 + (KPERM ? 2 * t : t)) * sak + g * sam;
            const int ks = KPERM ? sak : 4 * sak;          // distance between the thread's two K slots
            split_tf32(ap[0], ahi[0], alo[0]);
            split_tf32(ap[8 * sam], ahi[1], alo[1]);
            split_tf32(ap[ks], ahi[2], alo[2]);
            split_tf32(ap[ks + 8 * sam], ahi[3], alo[3]);
        }
#pragma unroll
        for (int i = 0; i < NT; ++i) {
            const int n0 = ncol0 + i * nstep;
            if (!GUARD || n0 < nmax) {
                uint32_t bhi[2], blo[2];
                const float* bp = B + (k0 + (KPERM ? 2 * t : t)) * sbk + (n0 + g) * sbn;
                split_tf32(bp[0], bhi[0], blo[0]);
                split_tf32(bp[KPERM ? sbk : 4 * sbk], bhi[1], blo[1]);
                mma_tf32(c[i], alo, bhi);
                mma_tf32(c[i], ahi, blo);
                mma_tf32(c[i], ahi, bhi);
            }
        }
    }
}

// streaming (read-once) global loads that do not pollute L1
__device__ __forceinline__ float ld_stream(const float* p) {
    float v;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}

#endif  // __CUDACC__

}  // namespace icrl
