// microbench.cu -- register-resident FP64 peak probes used as roofline denominators.
//
// MEASURED_PEAKS.json carries HBM and bf16 numbers but no FP64 figure, so
// bench.py measures the FP64 tensor-core (DMMA) and FP64 FMA ceilings of the very GPU it runs on:
// no memory traffic, 8 independent accumulator chains per warp/thread, all SMs busy.
#include "../../include/dft_b200_ext.h"
#include "dmma.cuh"
#include <cuda_runtime.h>

namespace {

__global__ void __launch_bounds__(256) dmma_peak_kernel(int iters, double* out) {
    double c[8][2];
#pragma unroll
    for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = 0.0;
    double a = 1.0 + 1e-9 * threadIdx.x, b = 1.0 - 1e-9 * threadIdx.x;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) dmma::mma8x8x4(c[i], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
    if (s == 123.456) out[0] = s;  // keep the chain alive
}

__global__ void __launch_bounds__(256) dfma_peak_kernel(int iters, double* out) {
    double c[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) c[i] = 1e-3 * i;
    const double a = 1.0 + 1e-9 * threadIdx.x, b = 1e-9 * threadIdx.x;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) c[i] = fma(c[i], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += c[i];
    if (s == 123.456) out[0] = s;
}

// DMMA throughput with exactly `warps_per_sm` resident warps per SM (one CTA per SM, the rest of the SM kept
// free by a large dynamic shared-memory request): how many warps per SM sub-partition does the FP64 tensor
// pipe need?  (It shapes the kernels: a lone warp per sub-partition reaches only about half the peak.)
double run_dmma_warps(int warps_per_sm, int iters) {
    int dev = 0, nsm = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
    double* d = nullptr;
    if (cudaMalloc(&d, 8) != cudaSuccess) return -1.0;
    const int threads = warps_per_sm * 32;
    const int smem = 200 * 1024;  // one CTA per SM
    cudaFuncSetAttribute(dmma_peak_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    dmma_peak_kernel<<<nsm, threads, smem>>>(iters / 4 + 1, d);
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        dmma_peak_kernel<<<nsm, threads, smem>>>(iters, d);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d);
    if (cudaGetLastError() != cudaSuccess) return -1.0;
    return (double)nsm * warps_per_sm * (double)iters * 8.0 * 512.0 / (best * 1e-3) / 1e12;
}

template <typename K>
double run_peak(K kernel, int iters, double flops_per_thread_iter) {
    int dev = 0, nsm = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
    double* d = nullptr;
    if (cudaMalloc(&d, 8) != cudaSuccess) return -1.0;
    const int blocks = nsm * 8, threads = 256;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    kernel<<<blocks, threads>>>(iters / 4 + 1, d);  // warm-up
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        kernel<<<blocks, threads>>>(iters, d);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d);
    if (cudaGetLastError() != cudaSuccess) return -1.0;
    const double flops = (double)blocks * threads * (double)iters * flops_per_thread_iter;
    return flops / (best * 1e-3) / 1e12;
}

}  // namespace

extern "C" double DFT_MicrobenchDMMA(int iters) {
    if (iters <= 0) iters = 4096;
    // 8 DMMA.8x8x4 per warp-iteration = 8 * 512 flop per 32 threads
    return run_peak(dmma_peak_kernel, iters, 8.0 * 512.0 / 32.0);
}

extern "C" double DFT_MicrobenchDMMAWarps(int warps_per_sm, int iters) {
    if (iters <= 0) iters = 4096;
    if (warps_per_sm < 1 || warps_per_sm > 8) return -1.0;
    return run_dmma_warps(warps_per_sm, iters);
}

extern "C" double DFT_MicrobenchDFMA(int iters) {
    if (iters <= 0) iters = 4096;
    return run_peak(dfma_peak_kernel, iters, 8.0 * 2.0);
}
