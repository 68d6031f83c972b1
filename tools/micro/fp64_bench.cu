// Microbenchmark: FP64 pipe on one SM -- dependent-chain latency and issue rate of DFMA / DADD / DMUL, and of the
// float <-> double conversions (F2F), which K3 and K5 lean on for bit-exact numpy float64 arithmetic.
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP, int OP>
__global__ void k(double* out, int iters, long long* cyc, double seed) {
    double x[ILP];
    float f[ILP];
    for (int i = 0; i < ILP; ++i) { x[i] = seed + i + threadIdx.x * 1e-3; f[i] = (float)x[i]; }
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            if (OP == 0) x[i] = __fma_rn(x[i], 1.0000001, 1e-9);
            if (OP == 1) x[i] = __dadd_rn(x[i], 1e-9);
            if (OP == 2) x[i] = __dmul_rn(x[i], 1.0000001);
            if (OP == 3) { f[i] = (float)x[i]; x[i] = (double)f[i] + 0.0; x[i] = __longlong_as_double(__double_as_longlong(x[i]) ^ 1); }   // 2 F2F
        }
    }
    const long long t1 = clock64();
    double s = 0;
    for (int i = 0; i < ILP; ++i) s += x[i] + f[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
int main() {
    double* out; long long* cyc; cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 1024);
    const int iters = 2000;
    long long h;
    const char* names[4] = {"DFMA", "DADD", "DMUL", "F2F.F32.F64 + F2F.F64.F32"};
#define RUN(ILP, WARPS, OP) k<ILP, OP><<<1, 32 * WARPS>>>(out, iters, cyc, 1.5); cudaDeviceSynchronize(); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); \
    printf("%-28s ILP=%d warps=%2d: %7.2f cycles per op per warp, SM rate %6.2f lanes/cycle\n", names[OP], ILP, WARPS, (double)h / (iters * ILP), 32.0 * ILP * WARPS * iters / h);
    RUN(1, 1, 0) RUN(8, 1, 0) RUN(8, 4, 0) RUN(8, 16, 0) RUN(1, 1, 1) RUN(8, 16, 1) RUN(1, 1, 2) RUN(8, 16, 2) RUN(1, 1, 3) RUN(8, 16, 3)
    printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
