// fp64 throughput of the device, measured: what bench.py divides by.  MEASURED_PEAKS.json carries no fp64 entry, so the
// roofline denominator of the fp64-bound kernels (Schur complement: DMMA m8n8k4; DLT: DFMA) is measured in the same run
// that reports the fractions, at the clocks of that run.
#include <cuda_runtime.h>
#include <stdio.h>
#include "../../include/argus_b200.h"

namespace {

// mma.sync m8n8k4 f64: 8 x 8 x 4 = 256 FMA per warp instruction; 8 independent accumulator pairs per warp
__global__ void __launch_bounds__(256) k_peak_dmma(double *out, int iters, double a, double b)
{
    double c[8][2];
#pragma unroll
    for (int i = 0; i < 8; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; }
    const double fa = a + threadIdx.x * 1e-9, fb = b;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(fa), "d"(fb));
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256) k_peak_dfma(double *out, int iters, double a, double b)
{
    double acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = fma(acc[i], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += acc[i];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

}  // namespace

extern "C" int argus_fp64_peak(int device, double seconds, double *dmma_tflops, double *dfma_tflops)
{
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) return ARGUS_E_CUDA;
    if (cudaSetDevice(device) != cudaSuccess) return ARGUS_E_CUDA;
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    const int blocks = sms * 4, threads = 256;                 // 32 warps per SM
    double *out = nullptr;
    if (cudaMalloc(&out, (size_t)blocks * threads * sizeof(double)) != cudaSuccess) return ARGUS_E_CUDA;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    if (seconds <= 0.0) seconds = 0.2;
    for (int which = 0; which < 2; ++which) {
        const int iters = 20000;
        // per launch: blocks * (threads / 32) warps * iters * 8 instructions * 256 FMA (DMMA), or blocks * threads * iters * 8 FMA (DFMA)
        const double flop = which == 0 ? (double)blocks * (threads / 32) * iters * 8.0 * 256.0 * 2.0 : (double)blocks * threads * iters * 8.0 * 2.0;
        auto launch = [&] { if (which == 0) k_peak_dmma<<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); else k_peak_dfma<<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); };
        launch(); cudaDeviceSynchronize();
        // sustained: launches back to back for about `seconds`
        float ms1 = 0.f;
        cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms1, e0, e1);
        int reps = (int)(seconds * 1e3 / (ms1 > 1e-3f ? ms1 : 1e-3f)); if (reps < 2) reps = 2; if (reps > 2000) reps = 2000;
        float ms = 0.f;
        cudaEventRecord(e0);
        for (int r = 0; r < reps; ++r) launch();
        cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
        const double tf = flop * reps / (ms * 1e-3) / 1e12;
        if (which == 0) { if (dmma_tflops) *dmma_tflops = tf; } else { if (dfma_tflops) *dfma_tflops = tf; }
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(out);
    return cudaGetLastError() == cudaSuccess ? ARGUS_OK : ARGUS_E_CUDA;
}
