// Do DMMA (mma.sync m8n8k4 f64) and DFMA share one datapath on B200?  Three kernels: DMMA only, DFMA only, both interleaved
// in every warp (independent chains).  If the mixed kernel's time is the SUM of the two, they share the fp64 units.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_mix fp64_mix.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int ND, int NF>
__global__ void __launch_bounds__(256) k(double *out, int iters, double a, double b)
{
    double c[8][2], f[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; f[i] = threadIdx.x * 1e-3 + i; }
    const double fa = a + threadIdx.x * 1e-9, fb = b;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (i < ND) asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(fa), "d"(fb));
#pragma unroll
            for (int q = 0; q < NF; ++q) f[(i + q) & 7] = fma(f[(i + q) & 7], a, b);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1] + f[i];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int ND, int NF> float run(double *out, int blocks, int iters)
{
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<ND, NF><<<blocks, 256>>>(out, iters, 1.0000001, 1e-9); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) { cudaEventRecord(e0); k<ND, NF><<<blocks, 256>>>(out, iters, 1.0000001, 1e-9); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms; }
    return best;
}
int main()
{
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int blocks = sms * 4, iters = 20000;
    double *out; cudaMalloc(&out, (size_t)blocks * 256 * 8);
    // per iteration and warp: ND x 8... the loop body runs 8 times: 8*[ND>0] DMMA (256 FMA each per warp) and 8*NF DFMA per thread
    const double dmma_flop = (double)blocks * 8 * iters * 8 * 256 * 2, dfma_unit = (double)blocks * 256 * iters * 8 * 2;
    float t_d = run<8, 0>(out, blocks, iters);
    float t_f = run<0, 8>(out, blocks, iters);       // 8 DFMA per body pass per thread = 256 FMA per warp pass, same FMA count as one DMMA
    float t_m = run<8, 8>(out, blocks, iters);
    printf("DMMA only : %.3f ms  %.2f TFLOP/s\n", t_d, dmma_flop / (t_d * 1e-3) / 1e12);
    printf("DFMA only : %.3f ms  %.2f TFLOP/s\n", t_f, dfma_unit * 8 / (t_f * 1e-3) / 1e12);
    printf("both      : %.3f ms  %.2f TFLOP/s  (sum of the two alone: %.3f ms)\n", t_m, (dmma_flop + dfma_unit * 8) / (t_m * 1e-3) / 1e12, t_d + t_f);
    return 0;
}
