// Measured arithmetic peaks of THIS device, for the compute side of the rooflines bench.py reports ("report
// fraction of min(HBM, FP32 or TC) roofline and state which bound applies").  MEASURED_PEAKS.json (driver-written) holds
// the HBM copy bandwidth and the dense bf16 cuBLAS throughput only; the two numbers the learner kernels are actually bound by
// are measured here with the same instructions they issue:
//   [0] FP32 SIMT: FFMA, 8 independent chains per thread, every SM full          -> TFLOP/s  (K1 / K2 arithmetic)
//   [1] mma.sync.m16n8k8 TF32 (fp32 accumulate), 8 independent accumulators/warp  -> TFLOP/s  (K4's GEMMs issue three of these
//       per fp32-class product: the effective "3xTF32" peak is a third of it)
#include "common.cuh"

namespace icrl {

__global__ void __launch_bounds__(256) peak_ffma_kernel(float* out, int iters) {
    float a[8], x = 1.0001f + 1e-7f * threadIdx.x, y = 0.9999f;
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = (float)i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = fmaf(a[i], x, y);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256) peak_mma_tf32_kernel(float* out, int iters) {
    float c[8][4];
    uint32_t a[4] = {0x3f800000u + threadIdx.x, 0x3f800000u, 0x3f900000u, 0x3fa00000u}, b[2] = {0x3f800000u, 0x3f700000u};
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) c[i][j] = 0.f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) mma_tf32(c[i], a, b);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) s += c[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

}  // namespace icrl

extern "C" int icrl_measure_peaks(double* tflops2, void* stream) {
    using namespace icrl;
    ICRL_CHECK_ARG(tflops2 != nullptr, "icrl_measure_peaks: NULL output");
    cudaStream_t st = (cudaStream_t)stream;
    const int grid = 8 * sm_count(), iters = 4096;
    void* out;
    if (int rc = device_scratch(SLOT_OUT0, (size_t)grid * 256 * 4, &out)) return rc;
    cudaEvent_t e0, e1;
    ICRL_CUDA(cudaEventCreate(&e0));
    ICRL_CUDA(cudaEventCreate(&e1));
    for (int which = 0; which < 2; ++which) {
        float best = 1e30f;
        for (int rep = 0; rep < 4; ++rep) {       // first repetition warms up
            ICRL_CUDA(cudaEventRecord(e0, st));
            if (which == 0) peak_ffma_kernel<<<grid, 256, 0, st>>>((float*)out, iters);
            else peak_mma_tf32_kernel<<<grid, 256, 0, st>>>((float*)out, iters);
            ICRL_LAUNCH_CHECK();
            ICRL_CUDA(cudaEventRecord(e1, st));
            ICRL_CUDA(cudaEventSynchronize(e1));
            float ms = 0.f;
            ICRL_CUDA(cudaEventElapsedTime(&ms, e0, e1));
            if (rep > 0 && ms < best) best = ms;
        }
        const double threads = (double)grid * 256;
        const double flops = which == 0 ? threads * iters * 64.0 * 2.0                       // 64 FFMA per thread per iteration
                                        : (threads / 32.0) * iters * 8.0 * (16.0 * 8 * 8 * 2);  // 8 mma (16x8x8) per warp per iteration
        tflops2[which] = flops / (best * 1e-3) / 1e12;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return 0;
}
