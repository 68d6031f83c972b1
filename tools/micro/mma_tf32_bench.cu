// Microbenchmark: mma.sync.m16n8k8 tf32 issue rate / latency on one SM (8 warps), vs FFMA.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void mma_tf32(float (&c)[4], const unsigned (&a)[4], const unsigned (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
template <int ILP>
__global__ void k_mma(float* out, int iters, long long* cyc) {
    float c[ILP][4];
    unsigned a[4] = {0x3f800000u + threadIdx.x, 0x3f800000u, 0x3f900000u, 0x3fa00000u}, b[2] = {0x3f800000u, 0x3f700000u};
    for (int i = 0; i < ILP; ++i) for (int j = 0; j < 4; ++j) c[i][j] = 0.f;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) mma_tf32(c[i], a, b);
    }
    long long t1 = clock64();
    float s = 0.f;
    for (int i = 0; i < ILP; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
int main() {
    float* out; long long* cyc; cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 1024);
    const int iters = 2000;
    long long h;
#define RUN(ILP, WARPS) k_mma<ILP><<<1, 32 * WARPS>>>(out, iters, cyc); cudaDeviceSynchronize(); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); \
    printf("ILP=%d warps=%d: %.2f cycles per mma per warp, SM rate %.1f tf32-FMA/cycle\n", ILP, WARPS, (double)h / (iters * ILP), 1024.0 * ILP * WARPS * iters / h);
    RUN(1, 1) RUN(4, 1) RUN(8, 1) RUN(1, 8) RUN(4, 8) RUN(8, 8) RUN(4, 16) RUN(12, 8)
    printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
