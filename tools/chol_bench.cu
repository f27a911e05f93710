// Stand-alone timing / phase trace of the reduced-system solvers (development tool, not part of the library):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -DARGUS_CHOL_TRACE -I argus_gui_b200/csrc -I include tools/chol_bench.cu -o gpurun_out/chol_bench
//   gpurun_out/chol_bench 128 256 512
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "argus_b200.h"
#include "ba_chol_coop.cuh"
#include "ba_chol_cluster.cuh"
using namespace argus;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

int main(int argc, char **argv)
{
    CK(cudaFuncSetAttribute(k_chol_cluster, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)chol_cluster_smem()));
    int sms = 0; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    for (int ai = 1; ai < argc; ++ai) {
        const int N = atoi(argv[ai]);
        std::vector<double> A((size_t)N * N), S((size_t)N * N), E(N), x(N);
        srand(N);
        for (auto &v : A) v = rand() / (double)RAND_MAX - 0.5;
        for (int i = 0; i < N; ++i)
            for (int j = 0; j < N; ++j) { double s = (i == j) ? 0.05 * N : 0.0; for (int k = 0; k < N; ++k) s += A[(size_t)i * N + k] * A[(size_t)j * N + k]; S[(size_t)i * N + j] = s; }
        for (auto &v : E) v = rand() / (double)RAND_MAX - 0.5;
        double *dS0, *dS, *dE, *dx, *dw; int *dflags; long long *dtr;
        const int nblk = (N + 31) / 32;
        CK(cudaMalloc(&dS0, S.size() * 8)); CK(cudaMalloc(&dS, S.size() * 8)); CK(cudaMalloc(&dE, N * 8)); CK(cudaMalloc(&dx, N * 8));
        CK(cudaMalloc(&dw, ((size_t)nblk * 1056 + N + 64) * 8)); CK(cudaMalloc(&dflags, 16)); CK(cudaMalloc(&dtr, 4096 * 8));
        CK(cudaMemcpy(dS0, S.data(), S.size() * 8, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dE, E.data(), N * 8, cudaMemcpyHostToDevice));
        CK(cudaMemset(dflags, 0, 16));
        cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        for (int which = 0; which < 2; ++which) {
            if (which == 0 && N > kClMaxN) continue;
            float tot = 0.f; const int reps = 20;
            for (int r = 0; r < reps + 3; ++r) {
                CK(cudaMemcpyAsync(dS, dS0, S.size() * 8, cudaMemcpyDeviceToDevice, 0)); CK(cudaMemsetAsync(dflags, 0, 16, 0));
#ifdef ARGUS_CHOL_TRACE
                { long long *p = (r == reps + 2) ? dtr : nullptr; int z = 0; CK(cudaMemcpyToSymbol(g_cl_trace, &p, sizeof(p))); CK(cudaMemcpyToSymbol(g_cl_trace_n, &z, sizeof(z))); }
#endif
                CK(cudaEventRecord(e0, 0));
                if (which == 0) k_chol_cluster<<<kClCtas, kClThreads, chol_cluster_smem(), 0>>>(dS, dE, dx, N, dw, dflags + 1, dflags + 2, dflags);
                else {
                    double *Sp = dS, *xp = dx, *wp = dw; const double *Ep = dE; int Nn = N; int *stt = dflags + 1, *fl = dflags + 2; const int *done = dflags;
                    void *args[] = {&Sp, &Ep, &xp, &Nn, &wp, &stt, &fl, &done};
                    int grid = std::min(sms, std::max(2, (nblk - 1) * nblk / 2));
                    CK(cudaLaunchCooperativeKernel((void *)k_chol_coop, dim3(grid), dim3(kCoopThreads), args, 0, 0));
                }
                CK(cudaEventRecord(e1, 0)); CK(cudaEventSynchronize(e1));
                float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (r >= 3) tot += ms;
            }
            CK(cudaMemcpy(x.data(), dx, N * 8, cudaMemcpyDeviceToHost));
            int st[4]; CK(cudaMemcpy(st, dflags, 16, cudaMemcpyDeviceToHost));
            double rn = 0, en = 0;
            for (int i = 0; i < N; ++i) { double s = -E[i]; for (int j = 0; j < N; ++j) s += S[(size_t)i * N + j] * x[j]; rn += s * s; en += E[i] * E[i]; }
            printf("N=%d %s: %.1f us per solve, status %d, residual %.2e\n", N, which == 0 ? "cluster" : "coop   ", 1e3 * tot / 20, st[1], sqrt(rn / en));
#ifdef ARGUS_CHOL_TRACE
            if (which == 0) {
                int n = 0; CK(cudaMemcpyFromSymbol(&n, g_cl_trace_n, sizeof(n)));
                std::vector<long long> t(n); CK(cudaMemcpy(t.data(), dtr, n * 8, cudaMemcpyDeviceToHost));
                printf("  trace (clocks since start, delta):");
                for (int i = 0; i < n; ++i) printf(" %lld(+%lld)", t[i] - t[0], i ? t[i] - t[i - 1] : 0);
                printf("\n");
            }
#endif
        }
        cudaFree(dS0); cudaFree(dS); cudaFree(dE); cudaFree(dx); cudaFree(dw); cudaFree(dflags); cudaFree(dtr);
    }
    return 0;
}
