// fp64 pipe micro-benchmark for B200 (sm_100a): DFMA vector rate vs DMMA (mma.sync m8n8k4 / m16n8k8 f64) rate.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peak fp64_peak.cu
// Used to fix the fp64 roofline denominator quoted in DESIGN.md / bench.py (MEASURED_PEAKS.json has no fp64 entry).
#include <cstdio>
#include <cuda_runtime.h>

#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1;}}while(0)

template<int ILP>
__global__ void dfma_kernel(double *out, int iters, double a, double b)
{
    double acc[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) acc[i] = fma(acc[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// mma.sync m8n8k4 f64: 8x8x4 = 256 FMA per warp instruction
template<int NACC>
__global__ void dmma884_kernel(double *out, int iters, double a, double b)
{
    double c[NACC][2];
#pragma unroll
    for (int i = 0; i < NACC; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; }
    double fa = a + threadIdx.x * 1e-9, fb = b;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) {
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(fa), "d"(fb));
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// mma.sync m16n8k8 f64 (sm_90+): 16x8x8 = 1024 FMA per warp instruction
template<int NACC>
__global__ void dmma1688_kernel(double *out, int iters, double a, double b)
{
    double c[NACC][4];
#pragma unroll
    for (int i = 0; i < NACC; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; c[i][2] = 1; c[i][3] = 2; }
    double fa0 = a + threadIdx.x * 1e-9, fa1 = a, fa2 = a * 0.5, fa3 = a * 0.25, fb0 = b, fb1 = b * 0.5;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) {
            asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                         : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                         : "d"(fa0), "d"(fa1), "d"(fa2), "d"(fa3), "d"(fb0), "d"(fb1));
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template<typename F>
static float time_it(F launch, int reps)
{
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    launch(); launch(); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    return best;
}

int main()
{
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    int sms = prop.multiProcessorCount;
    printf("device %s sms %d clock %d kHz\n", prop.name, sms, prop.clockRate);
    double *out; CK(cudaMalloc(&out, sizeof(double) * 148 * 64 * 1024));
    const int iters = 4096;
    for (int wps = 4; wps <= 32; wps *= 2) {            // warps per SM (one CTA per SM x occupancy 1..)
        int threads = 32 * wps > 1024 ? 1024 : 32 * wps;
        int ctas = sms * ((32 * wps + threads - 1) / threads);
        {
            float ms = time_it([&]{ dfma_kernel<8><<<ctas, threads>>>(out, iters, 1.0000001, 1e-9); }, 5);
            double fl = 2.0 * 8 * iters * (double)ctas * threads;
            printf("DFMA   ilp8  warps/SM %2d : %8.3f ms  %8.2f TFLOP/s\n", wps, ms, fl / ms * 1e-9);
        }
        {
            float ms = time_it([&]{ dmma884_kernel<8><<<ctas, threads>>>(out, iters, 1.0000001, 1e-9); }, 5);
            double fl = 2.0 * 256 * 8 * iters * (double)ctas * (threads / 32);
            printf("DMMA884  x8  warps/SM %2d : %8.3f ms  %8.2f TFLOP/s\n", wps, ms, fl / ms * 1e-9);
        }
        {
            float ms = time_it([&]{ dmma1688_kernel<4><<<ctas, threads>>>(out, iters, 1.0000001, 1e-9); }, 5);
            double fl = 2.0 * 1024 * 4 * iters * (double)ctas * (threads / 32);
            printf("DMMA1688 x4  warps/SM %2d : %8.3f ms  %8.2f TFLOP/s\n", wps, ms, fl / ms * 1e-9);
        }
    }
    // sustained (about 2 s of DFMA) to see the clock under load
    {
        int threads = 512, ctas = sms * 2;
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        for (int r = 0; r < 200; ++r) dfma_kernel<8><<<ctas, threads>>>(out, iters * 4, 1.0000001, 1e-9);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double fl = 200.0 * 2.0 * 8 * iters * 4 * (double)ctas * threads;
        printf("DFMA sustained: %8.1f ms  %8.2f TFLOP/s\n", ms, fl / ms * 1e-9);
        cudaEventRecord(e0);
        for (int r = 0; r < 200; ++r) dmma884_kernel<8><<<ctas, threads>>>(out, iters * 4, 1.0000001, 1e-9);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        fl = 200.0 * 2.0 * 256 * 8 * iters * 4 * (double)ctas * (threads / 32);
        printf("DMMA884 sustained: %8.1f ms  %8.2f TFLOP/s\n", ms, fl / ms * 1e-9);
    }
    CK(cudaDeviceSynchronize());
    return 0;
}
