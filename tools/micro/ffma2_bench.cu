// Microbenchmark: packed FP32 FMA (fma.rn.f32x2 -> SASS FFMA2) against scalar FFMA on one SM, in the shape K1 / K2 use it:
// one thread = one row, acc[HP] += x_k * W[k][0..HP) with the k-major weight row read from shared memory by broadcast LDS.128.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/micro/ffma2_bench tools/micro/ffma2_bench.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void fma2(float2& acc, float x, float2 w) {
    unsigned long long a = *reinterpret_cast<unsigned long long*>(&acc), ww = *reinterpret_cast<unsigned long long*>(&w), xx;
    asm("mov.b64 %0, {%1, %1};" : "=l"(xx) : "f"(x));                 // ptxas folds the broadcast into the FFMA2 operand
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(a) : "l"(xx), "l"(ww));
    acc = *reinterpret_cast<float2*>(&a);
}

constexpr int HP = 24, K = 24;

template <bool PACKED>
__global__ void k_row(const float* __restrict__ x, const float* __restrict__ w, float* out, int iters, long long* cyc) {
    __shared__ __align__(16) float W[K * HP];
    __shared__ float X[K * 32];
    for (int i = threadIdx.x; i < K * HP; i += blockDim.x) W[i] = w[i];
    for (int i = threadIdx.x; i < K * 32; i += blockDim.x) X[i] = x[i];
    __syncthreads();
    float acc[HP];
    float2 acc2[HP / 2];
    for (int j = 0; j < HP; ++j) acc[j] = 0.f;
    for (int j = 0; j < HP / 2; ++j) acc2[j] = make_float2(0.f, 0.f);
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll 4
        for (int k = 0; k < K; ++k) {
            const float xv = fminf(fmaxf(X[k * 32 + (threadIdx.x & 31)], -20.f), 20.f);
            const float4* w4 = reinterpret_cast<const float4*>(W + k * HP);
#pragma unroll
            for (int j = 0; j < HP / 4; ++j) {
                const float4 ww = w4[j];
                if (PACKED) {
                    fma2(acc2[2 * j], xv, make_float2(ww.x, ww.y));
                    fma2(acc2[2 * j + 1], xv, make_float2(ww.z, ww.w));
                } else {
                    acc[4 * j + 0] = fmaf(xv, ww.x, acc[4 * j + 0]);
                    acc[4 * j + 1] = fmaf(xv, ww.y, acc[4 * j + 1]);
                    acc[4 * j + 2] = fmaf(xv, ww.z, acc[4 * j + 2]);
                    acc[4 * j + 3] = fmaf(xv, ww.w, acc[4 * j + 3]);
                }
            }
        }
    }
    const long long t1 = clock64();
    float z = 0.f;
    for (int j = 0; j < HP; ++j) z += acc[j];
    for (int j = 0; j < HP / 2; ++j) z += acc2[j].x + acc2[j].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = z;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// register-only peaks: 8 independent chains per thread
template <int MODE>   // 0: FFMA a = a * x + y;  1: FFMA2 (three 64-bit operands);  2: FFMA2 with a broadcast scalar operand
__global__ void k_peak(float* out, int iters, long long* cyc) {
    float a[16], x = 1.0001f + 1e-7f * threadIdx.x, y = 0.9999f;
    for (int i = 0; i < 16; ++i) a[i] = (float)i;
    unsigned long long xx, yy;
    asm("mov.b64 %0, {%1, %2};" : "=l"(xx) : "f"(x), "f"(y));
    asm("mov.b64 %0, {%1, %2};" : "=l"(yy) : "f"(y), "f"(x));
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) {
#pragma unroll
            for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], x, y);
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                unsigned long long v;
                asm("mov.b64 %0, {%1, %2};" : "=l"(v) : "f"(a[2 * i]), "f"(a[2 * i + 1]));
                if (MODE == 1) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(v) : "l"(xx), "l"(yy));
                else {
                    unsigned long long xb;
                    asm("mov.b64 %0, {%1, %1};" : "=l"(xb) : "f"(x));
                    asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(v) : "l"(xb), "l"(yy));
                }
                asm("mov.b64 {%0, %1}, %2;" : "=f"(a[2 * i]), "=f"(a[2 * i + 1]) : "l"(v));
            }
        }
    }
    const long long t1 = clock64();
    float s = 0.f;
    for (int i = 0; i < 16; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}


// register-only row loops: 24 hidden units, weights held in registers (no shared-memory traffic), NLDS extra broadcast
// LDS.128 per 24 units mixed in (their results are consumed by a cheap xor so that they are not dead)
template <int MODE, int NLDS>   // 0: FFMA acc[j] += x * w[j];  1: FFMA2 {acc_r0, acc_r1}[j] += {x_r0, x_r1} * w[j] (broadcast w);  2: FFMA2 {acc[j], acc[j+1]} += {x, x} * {w[j], w[j+1]}
__global__ void k_reg(const float* __restrict__ wg, float* out, int iters, long long* cyc) {
    __shared__ __align__(16) float S[64 * 4];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) S[i] = 0.f;
    __syncthreads();
    float w[24], acc[24];
    float2 acc2[24];
    for (int j = 0; j < 24; ++j) { w[j] = wg[j]; acc[j] = 0.f; acc2[j] = make_float2(0.f, 0.f); }
    float x = 1.0f + 1e-6f * threadIdx.x;
    float2 x2 = make_float2(x, x * 0.5f);
    unsigned sink = 0;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int l = 0; l < NLDS; ++l) {
            const uint4 v = *reinterpret_cast<const uint4*>(S + 4 * ((it + l) & 63));
            sink ^= v.x ^ v.y ^ v.z ^ v.w;
        }
        if (MODE == 0) {
#pragma unroll
            for (int j = 0; j < 24; ++j) acc[j] = fmaf(x, w[j], acc[j]);
        } else if (MODE == 1) {
#pragma unroll
            for (int j = 0; j < 24; ++j) {
                unsigned long long a, xx, ww;
                asm("mov.b64 %0, {%1, %2};" : "=l"(a) : "f"(acc2[j].x), "f"(acc2[j].y));
                asm("mov.b64 %0, {%1, %2};" : "=l"(xx) : "f"(x2.x), "f"(x2.y));
                asm("mov.b64 %0, {%1, %1};" : "=l"(ww) : "f"(w[j]));
                asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(a) : "l"(xx), "l"(ww));
                asm("mov.b64 {%0, %1}, %2;" : "=f"(acc2[j].x), "=f"(acc2[j].y) : "l"(a));
            }
        } else {
#pragma unroll
            for (int j = 0; j < 12; ++j) {
                unsigned long long a, xx, ww;
                asm("mov.b64 %0, {%1, %2};" : "=l"(a) : "f"(acc2[j].x), "f"(acc2[j].y));
                asm("mov.b64 %0, {%1, %1};" : "=l"(xx) : "f"(x));
                asm("mov.b64 %0, {%1, %2};" : "=l"(ww) : "f"(w[2 * j]), "f"(w[2 * j + 1]));
                asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(a) : "l"(xx), "l"(ww));
                asm("mov.b64 {%0, %1}, %2;" : "=f"(acc2[j].x), "=f"(acc2[j].y) : "l"(a));
            }
        }
    }
    const long long t1 = clock64();
    float z = __uint_as_float(sink & 1u);
    for (int j = 0; j < 24; ++j) z += acc[j] + acc2[j].x + acc2[j].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = z;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

int main() {
    float *out, *x, *w;
    long long* cyc;
    cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 1024); cudaMalloc(&x, K * 512 * 4); cudaMalloc(&w, K * HP * 4);
    cudaMemset(x, 0, K * 512 * 4); cudaMemset(w, 0, K * HP * 4);
    const int iters = 500;
    long long h;
    for (int warps : {4, 8, 16, 32}) {
        k_row<false><<<1, 32 * warps>>>(x, w, out, iters, cyc); cudaDeviceSynchronize(); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("row loop FFMA   warps=%2d: %.1f cycles per k per warp-pass, %.1f FMA/cycle/SM\n", warps, (double)h / (iters * K), (double)HP * 32 * warps * K * iters / h);
        k_row<true><<<1, 32 * warps>>>(x, w, out, iters, cyc); cudaDeviceSynchronize(); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("row loop FFMA2  warps=%2d: %.1f cycles per k per warp-pass, %.1f FMA/cycle/SM\n", warps, (double)h / (iters * K), (double)HP * 32 * warps * K * iters / h);
    }
    for (int warps : {4, 8, 16, 32}) {
        k_peak<0><<<1, 32 * warps>>>(out, 4000, cyc); cudaDeviceSynchronize(); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("peak FFMA            warps=%2d: %.1f FMA/cycle/SM\n", warps, 16.0 * 32 * warps * 4000 / h);
        k_peak<1><<<1, 32 * warps>>>(out, 4000, cyc); cudaDeviceSynchronize(); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("peak FFMA2 (3 x 64b) warps=%2d: %.1f FMA/cycle/SM\n", warps, 16.0 * 32 * warps * 4000 / h);
        k_peak<2><<<1, 32 * warps>>>(out, 4000, cyc); cudaDeviceSynchronize(); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("peak FFMA2 (bcast)   warps=%2d: %.1f FMA/cycle/SM\n", warps, 16.0 * 32 * warps * 4000 / h);
    }

    {
        const int it2 = 4000;
#define RUNREG(MODE, NLDS, FMAS, NAME)                                                                                      \
    for (int warps : {8, 16, 32}) {                                                                                         \
        k_reg<MODE, NLDS><<<1, 32 * warps>>>(w, out, it2, cyc); cudaDeviceSynchronize(); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); \
        printf("reg loop %-28s +%d LDS.128 warps=%2d: %.1f FMA/cycle/SM\n", NAME, NLDS, warps, (double)FMAS * 32 * warps * it2 / h);  \
    }
        RUNREG(0, 0, 24, "FFMA acc+=x*w[j]") RUNREG(1, 0, 48, "FFMA2 2 rows, bcast w[j]") RUNREG(2, 0, 24, "FFMA2 unit pairs, bcast x")
        RUNREG(0, 3, 24, "FFMA acc+=x*w[j]") RUNREG(0, 6, 24, "FFMA acc+=x*w[j]") RUNREG(1, 3, 48, "FFMA2 2 rows, bcast w[j]") RUNREG(1, 6, 48, "FFMA2 2 rows, bcast w[j]") RUNREG(2, 6, 24, "FFMA2 unit pairs, bcast x")
    }
    printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
