// Per-camera 11-parameter DLT fit: the least-squares problem of tools.solve_dlt
// (/root/reference/argus_gui/tools.py:220-280) and OutlierWindow.getCoefficients (argus_gui/sbaDriver.py:1027-1075):
//   rows  [X Y Z 1 0 0 0 0 -uX -uY -uZ | u]  and  [0 0 0 0 X Y Z 1 -vX -vY -vZ | v]  for every point the camera sees,
//   L = lstsq(A, B), then the reprojection distance of every used point, their sum / mean / population std (the 3-sigma
//   outlier rule of sbaDriver.py:1077-1079 is applied by the host on the returned distances).
// The reference solves through LAPACK's SVD; forming A^T A would square the condition number (the -uX columns are ~1e3 times the
// X columns), so the fit is a QR factorisation computed as a reduction tree (TSQR): every thread streams its rows into a private
// 11 x 12 triangle [R | Q^T b] with Givens rotations (registers), triangles are merged pairwise in shared memory, block results by
// a last single-CTA kernel which back-substitutes.  Row order and tree shape depend only on N: results are reproducible.
#include <cuda_runtime.h>
#include <stdio.h>
#include <math.h>
#include "../../include/argus_b200.h"

namespace {

#define FIT_CUDA_CHECK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
    fprintf(stderr, "argus_b200: CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
    return ARGUS_E_CUDA; } } while (0)

constexpr int kT = 128;          // threads per block (one triangle each)
constexpr int kTri = 77;         // packed upper triangle of the 11 x 12 system: row j holds columns j..11
constexpr int kMaxBlocks = 128;  // the final kernel merges at most kT block triangles

__host__ __device__ constexpr int tri_at(int j, int k) { return j * 12 - (j * (j - 1)) / 2 + (k - j); }

// rotate the row r (columns j0..11 significant) into the triangle R (registers; everything unrolled)
__device__ __forceinline__ void givens_row(double (&R)[kTri], double (&r)[12])
{
#pragma unroll
    for (int j = 0; j < 11; ++j) {
        const double b = r[j];
        if (b != 0.0) {
            const double a = R[tri_at(j, j)];
            const double h = sqrt(a * a + b * b);
            const double c = a / h, s = b / h;
#pragma unroll
            for (int k = j; k < 12; ++k) {
                const double x = R[tri_at(j, k)], y = r[k];
                R[tri_at(j, k)] = c * x + s * y;
                r[k] = c * y - s * x;
            }
        }
    }
}

// merge triangle B into triangle A (both in shared memory, runtime loops)
__device__ void merge_tri(double *A, const double *B)
{
    double r[12];
    for (int i = 0; i < 11; ++i) {
        for (int k = 0; k < 12; ++k) r[k] = (k >= i) ? B[tri_at(i, k)] : 0.0;
        for (int j = i; j < 11; ++j) {
            const double b = r[j];
            if (b == 0.0) continue;
            const double a = A[tri_at(j, j)];
            const double h = sqrt(a * a + b * b);
            const double c = a / h, s = b / h;
            for (int k = j; k < 12; ++k) {
                const double x = A[tri_at(j, k)], y = r[k];
                A[tri_at(j, k)] = c * x + s * y;
                r[k] = c * y - s * x;
            }
        }
    }
}

__device__ __forceinline__ bool row_used(double u, double v, int flags) { return !(isnan(u) || ((flags & 1) && isnan(v))); }

__device__ void block_tree(double *sh, double (&R)[kTri])
{
    double *mine = sh + (size_t)threadIdx.x * kTri;
#pragma unroll
    for (int t = 0; t < kTri; ++t) mine[t] = R[t];
    __syncthreads();
    for (int stride = kT / 2; stride > 0; stride >>= 1) {
        if ((int)threadIdx.x < stride) merge_tri(mine, mine + (size_t)stride * kTri);
        __syncthreads();
    }
}

__global__ void __launch_bounds__(kT) k_fit_qr(const double *__restrict__ xyz, const double *__restrict__ uv, int64_t N, int flags,
                                               double *__restrict__ part /* grid x 77 */, long long *__restrict__ used)
{
    extern __shared__ double sh[];
    double R[kTri];
#pragma unroll
    for (int t = 0; t < kTri; ++t) R[t] = 0.0;
    long long cnt = 0;
    for (int64_t i = (int64_t)blockIdx.x * kT + threadIdx.x; i < N; i += (int64_t)gridDim.x * kT) {
        const double u = uv[2 * i], v = uv[2 * i + 1];
        if (!row_used(u, v, flags)) continue;
        const double X = xyz[3 * i], Y = xyz[3 * i + 1], Z = xyz[3 * i + 2];
        double r0[12] = {X, Y, Z, 1.0, 0.0, 0.0, 0.0, 0.0, -u * X, -u * Y, -u * Z, u};
        givens_row(R, r0);
        double r1[12] = {0.0, 0.0, 0.0, 0.0, X, Y, Z, 1.0, -v * X, -v * Y, -v * Z, v};
        givens_row(R, r1);
        ++cnt;
    }
    block_tree(sh, R);
    for (int t = threadIdx.x; t < kTri; t += kT) part[(size_t)blockIdx.x * kTri + t] = sh[t];
    // number of used points (integers: order-independent)
    __shared__ long long scnt;
    if (threadIdx.x == 0) scnt = 0;
    __syncthreads();
    atomicAdd((unsigned long long *)&scnt, (unsigned long long)cnt);
    __syncthreads();
    if (threadIdx.x == 0) atomicAdd((unsigned long long *)used, (unsigned long long)scnt);
}

// merge the block triangles and back-substitute: L (11)
__global__ void __launch_bounds__(kT) k_fit_solve(const double *__restrict__ part, int nparts, double *__restrict__ L, int *__restrict__ status)
{
    extern __shared__ double sh[];
    double R[kTri];
#pragma unroll
    for (int t = 0; t < kTri; ++t) R[t] = ((int)threadIdx.x < nparts) ? part[(size_t)threadIdx.x * kTri + t] : 0.0;
    block_tree(sh, R);
    if (threadIdx.x == 0) {
        double x[11], dmax = 0.0;
        int ok = 1;
        for (int j = 0; j < 11; ++j) dmax = fmax(dmax, fabs(sh[tri_at(j, j)]));
        for (int j = 10; j >= 0; --j) {
            double s = sh[tri_at(j, 11)];
            for (int k = j + 1; k < 11; ++k) s -= sh[tri_at(j, k)] * x[k];
            const double d = sh[tri_at(j, j)];
            if (!(fabs(d) > 1e-13 * dmax)) ok = 0;          // numerically rank deficient (the reference's lstsq would return a minimum-norm answer)
            x[j] = s / d;
        }
        for (int j = 0; j < 11; ++j) L[j] = x[j];
        *status = ok;
    }
}

__device__ __forceinline__ double bsum(double v, double *sh8)
{
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sh8[threadIdx.x >> 5] = v;
    __syncthreads();
    double r = 0.0;
    if (threadIdx.x == 0) for (int q = 0; q < kT / 32; ++q) r += sh8[q];
    return r;
}

// reprojection distance per used point (sbaDriver.py:1058-1067), NaN for skipped points; pass 0: sum, pass 1: sum (e - mean)^2
__global__ void __launch_bounds__(kT) k_fit_errors(const double *__restrict__ xyz, const double *__restrict__ uv, int64_t N, int flags,
                                                   const double *__restrict__ L, double *__restrict__ errors, int pass, const double *__restrict__ mean,
                                                   double *__restrict__ part)
{
    __shared__ double sh8[8];
    double l[11];
    for (int j = 0; j < 11; ++j) l[j] = L[j];
    const double mu = pass ? *mean : 0.0;
    double acc = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * kT + threadIdx.x; i < N; i += (int64_t)gridDim.x * kT) {
        const double u = uv[2 * i], v = uv[2 * i + 1];
        double e = __longlong_as_double(0x7ff8000000000000LL);
        if (row_used(u, v, flags)) {
            const double X = xyz[3 * i], Y = xyz[3 * i + 1], Z = xyz[3 * i + 2];
            const double den = l[8] * X + l[9] * Y + l[10] * Z + 1.0;
            const double ru = (l[0] * X + l[1] * Y + l[2] * Z + l[3]) / den, rv = (l[4] * X + l[5] * Y + l[6] * Z + l[7]) / den;
            const double du = ru - u, dv = rv - v;
            e = sqrt(du * du + dv * dv);
            acc += pass ? (e - mu) * (e - mu) : e;
        }
        if (!pass && errors) errors[i] = e;
    }
    const double r = bsum(acc, sh8);
    if (threadIdx.x == 0) part[blockIdx.x] = r;
}

__global__ void k_fit_stats(const double *__restrict__ part, int nparts, const long long *__restrict__ used, int pass, double *__restrict__ stats)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    double s = 0.0;
    for (int q = 0; q < nparts; ++q) s += part[q];
    const double n = (double)*used;
    if (!pass) { stats[0] = n; stats[1] = s; stats[2] = s / n; }
    else stats[3] = sqrt(s / n);
}

struct FitScratch {
    void *p[8] = {0}; int n = 0;
    template <typename T> cudaError_t alloc(T **out, size_t cnt) { cudaError_t e = cudaMalloc(out, (cnt ? cnt : 1) * sizeof(T)); if (e == cudaSuccess) p[n++] = *out; return e; }
    ~FitScratch() { for (int i = 0; i < n; ++i) cudaFree(p[i]); }
};

}  // namespace

extern "C" int argus_dlt_fit(const double *xyz, const double *uv, int64_t N, int flags, double *L, double *errors, double *stats)
{
    if (!xyz || !uv || !L || !stats || N < 6) return ARGUS_E_ARG;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) {
        fprintf(stderr, "argus_b200: no CUDA device available (there is no CPU fallback)\n");
        return ARGUS_E_CUDA;
    }
    const int grid = (int)((N + kT - 1) / kT < kMaxBlocks ? (N + kT - 1) / kT : kMaxBlocks);
    const size_t smem = (size_t)kT * kTri * sizeof(double);
    FIT_CUDA_CHECK(cudaFuncSetAttribute(k_fit_qr, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    FIT_CUDA_CHECK(cudaFuncSetAttribute(k_fit_solve, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    FitScratch sc;
    double *d_xyz, *d_uv, *d_part, *d_small, *d_err; long long *d_used; int *d_status;
    FIT_CUDA_CHECK(sc.alloc(&d_xyz, (size_t)3 * N)); FIT_CUDA_CHECK(sc.alloc(&d_uv, (size_t)2 * N));
    FIT_CUDA_CHECK(sc.alloc(&d_part, (size_t)kMaxBlocks * kTri)); FIT_CUDA_CHECK(sc.alloc(&d_small, 32));
    FIT_CUDA_CHECK(sc.alloc(&d_err, (size_t)N)); FIT_CUDA_CHECK(sc.alloc(&d_used, 1)); FIT_CUDA_CHECK(sc.alloc(&d_status, 1));
    FIT_CUDA_CHECK(cudaMemcpy(d_xyz, xyz, (size_t)3 * N * sizeof(double), cudaMemcpyHostToDevice));
    FIT_CUDA_CHECK(cudaMemcpy(d_uv, uv, (size_t)2 * N * sizeof(double), cudaMemcpyHostToDevice));
    FIT_CUDA_CHECK(cudaMemset(d_used, 0, sizeof(long long)));
    FIT_CUDA_CHECK(cudaStreamSynchronize(cudaStreamLegacy));
    double *d_L = d_small, *d_stats = d_small + 16;
    k_fit_qr<<<grid, kT, smem>>>(d_xyz, d_uv, N, flags, d_part, d_used);
    k_fit_solve<<<1, kT, smem>>>(d_part, grid, d_L, d_status);
    k_fit_errors<<<grid, kT>>>(d_xyz, d_uv, N, flags, d_L, d_err, 0, nullptr, d_part);
    k_fit_stats<<<1, 32>>>(d_part, grid, d_used, 0, d_stats);
    k_fit_errors<<<grid, kT>>>(d_xyz, d_uv, N, flags, d_L, nullptr, 1, d_stats + 2, d_part);
    k_fit_stats<<<1, 32>>>(d_part, grid, d_used, 1, d_stats);
    FIT_CUDA_CHECK(cudaGetLastError());
    int status = 0;
    FIT_CUDA_CHECK(cudaMemcpy(&status, d_status, sizeof(int), cudaMemcpyDeviceToHost));
    FIT_CUDA_CHECK(cudaMemcpy(L, d_L, 11 * sizeof(double), cudaMemcpyDeviceToHost));
    FIT_CUDA_CHECK(cudaMemcpy(stats, d_stats, 4 * sizeof(double), cudaMemcpyDeviceToHost));
    if (errors) FIT_CUDA_CHECK(cudaMemcpy(errors, d_err, (size_t)N * sizeof(double), cudaMemcpyDeviceToHost));
    return status ? ARGUS_OK : ARGUS_E_SBA;      // rank-deficient system (fewer than 6 points in general position)
}
