// Two-view initialisation of the bundle adjustment: what Triangulator.triangulate does for one camera
// paired with the reference camera (/root/reference/argus_gui/triangulate.py:27-212) as ONE call:
//   normalize()          :40-60    cv2.undistortPoints(src float32, K, dist, P=None) of both pixel sets          -> k_pair_undistort
//   getFundamentalMat()  :209-210  cv2.findFundamentalMat(p2, p1, FM_8POINT): Hartley-normalised 8-point over ALL points
//                                  (OpenCV 4.x calib3d/fundam.cpp run8Point: centroids, sqrt(2)/mean distance, the 9x9 moment
//                                  matrix sum r r^T, its smallest eigenvector, rank-2 projection, de-normalisation)
//                                  -> k_pair_centroid / k_pair_scale / k_pair_moments + a 9x9 Jacobi eigen-solve on the host
//   SVD of F, the four R candidates, det filter :69-93   (3x3, host)
//   4 x cv2.triangulatePoints per point + the sign vote :96-160   -> k_pair_vote (null vector of the 4x4 system through a
//                                  Jacobi eigen-solve of A^T A, one thread per point, all four hypotheses, integer votes)
//   decision logic, baseline sign, final cloud :162-207   -> host + k_pair_emit
// The per-point work (undistortion, moments, 4 SVD-equivalent solves) is what costs seconds in the reference's Python loops.
#include <cuda_runtime.h>
#include <stdio.h>
#include <math.h>
#include <vector>
#include "../../include/argus_b200.h"

namespace {

#define PAIR_CUDA_CHECK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
    fprintf(stderr, "argus_b200: CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
    return ARGUS_E_CUDA; } } while (0)

constexpr int kBlock = 256;
constexpr int kMaxGrid = 148 * 8;

// cyclic Jacobi eigen-solver for a symmetric N x N matrix: A is destroyed (diagonal = eigenvalues), V = eigenvectors (columns)
template <int N>
__host__ __device__ inline void jacobi_eig(double (*A)[N], double (*V)[N])
{
    for (int i = 0; i < N; ++i) for (int j = 0; j < N; ++j) V[i][j] = (i == j) ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 60; ++sweep) {
        double off = 0.0, dia = 0.0;
        for (int i = 0; i < N; ++i) { dia += A[i][i] * A[i][i]; for (int j = i + 1; j < N; ++j) off += A[i][j] * A[i][j]; }
        if (off <= 1e-60 || off <= 1e-34 * dia) break;
        for (int p = 0; p < N - 1; ++p)
            for (int q = p + 1; q < N; ++q) {
                const double apq = A[p][q];
                if (apq == 0.0) continue;
                const double theta = (A[q][q] - A[p][p]) / (2.0 * apq);
                const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < N; ++k) {               // A <- J^T A J
                    const double akp = A[k][p], akq = A[k][q];
                    A[k][p] = c * akp - s * akq; A[k][q] = s * akp + c * akq;
                }
                for (int k = 0; k < N; ++k) {
                    const double apk = A[p][k], aqk = A[q][k];
                    A[p][k] = c * apk - s * aqk; A[q][k] = s * apk + c * aqk;
                }
                for (int k = 0; k < N; ++k) {
                    const double vkp = V[k][p], vkq = V[k][q];
                    V[k][p] = c * vkp - s * vkq; V[k][q] = s * vkp + c * vkq;
                }
            }
    }
}

// cv2.undistortPoints(src(float32), K, dist, P=None): 5 fixed-point iterations, normalised coordinates, float32 out
// (the same arithmetic as undistort_f32 in dlt_kernels.cu without the re-projection)
__device__ __forceinline__ void undistort_norm_f32(double pu, double pv, const double *__restrict__ pr, float &ox, float &oy)
{
    const double fx = pr[0], cx = pr[1], cy = pr[2];
    const double k1 = pr[3], k2 = pr[4], p1 = pr[5], p2 = pr[6], k3 = pr[7];
    const double u = (double)(float)pu, v = (double)(float)pv;
    const double ifx = 1.0 / fx;
    double x = __dmul_rn(__dadd_rn(u, -cx), ifx), y = __dmul_rn(__dadd_rn(v, -cy), ifx);
    const double x0 = x, y0 = y;
#pragma unroll 1
    for (int j = 0; j < 5; ++j) {
        const double r2 = __dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y));
        const double den = __dadd_rn(1.0, __dmul_rn(__dadd_rn(__dmul_rn(__dadd_rn(__dmul_rn(k3, r2), k2), r2), k1), r2));
        const double icdist = 1.0 / den;
        if (icdist < 0.0) { x = x0; y = y0; break; }
        const double dX = __dadd_rn(__dmul_rn(__dmul_rn(__dmul_rn(2.0, p1), x), y),
                                    __dmul_rn(p2, __dadd_rn(r2, __dmul_rn(__dmul_rn(2.0, x), x))));
        const double dY = __dadd_rn(__dmul_rn(p1, __dadd_rn(r2, __dmul_rn(__dmul_rn(2.0, y), y))),
                                    __dmul_rn(__dmul_rn(__dmul_rn(2.0, p2), x), y));
        x = __dmul_rn(__dadd_rn(x0, -dX), icdist);
        y = __dmul_rn(__dadd_rn(y0, -dY), icdist);
    }
    ox = (float)x; oy = (float)y;
}

__device__ __forceinline__ double blk_sum(double v, double *sh)
{
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) sh[w] = v;
    __syncthreads();
    double r = 0.0;
    if (threadIdx.x == 0) for (int q = 0; q < kBlock / 32; ++q) r += sh[q];
    return r;       // valid in thread 0
}

// x1, x2 (float2 per point) + per-block sums of the four coordinates
__global__ void __launch_bounds__(kBlock) k_pair_undistort(const double *__restrict__ p1, const double *__restrict__ p2, int64_t N,
                                                           const double *__restrict__ prof /* 16 */, float2 *__restrict__ x1,
                                                           float2 *__restrict__ x2, double *__restrict__ part /* grid x 4 */)
{
    __shared__ double sh[8];
    double s[4] = {0, 0, 0, 0};
    for (int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x; i < N; i += (int64_t)gridDim.x * kBlock) {
        float a, b, c, d;
        undistort_norm_f32(p1[2 * i], p1[2 * i + 1], prof, a, b);
        undistort_norm_f32(p2[2 * i], p2[2 * i + 1], prof + 8, c, d);
        x1[i] = make_float2(a, b); x2[i] = make_float2(c, d);
        s[0] += (double)a; s[1] += (double)b; s[2] += (double)c; s[3] += (double)d;
    }
    for (int q = 0; q < 4; ++q) { const double r = blk_sum(s[q], sh); if (threadIdx.x == 0) part[4 * blockIdx.x + q] = r; }
}

// mean distance to the centroid of each set (run8Point's scale): per-block sums
__global__ void __launch_bounds__(kBlock) k_pair_scale(const float2 *__restrict__ x1, const float2 *__restrict__ x2, int64_t N,
                                                       const double *__restrict__ cen /* 4 */, double *__restrict__ part /* grid x 2 */)
{
    __shared__ double sh[8];
    double s0 = 0.0, s1 = 0.0;
    const double c0 = cen[0], c1 = cen[1], c2 = cen[2], c3 = cen[3];
    for (int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x; i < N; i += (int64_t)gridDim.x * kBlock) {
        const float2 a = x1[i], b = x2[i];
        const double ax = (double)a.x - c0, ay = (double)a.y - c1, bx = (double)b.x - c2, by = (double)b.y - c3;
        s0 += sqrt(ax * ax + ay * ay); s1 += sqrt(bx * bx + by * by);
    }
    double r = blk_sum(s0, sh); if (threadIdx.x == 0) part[2 * blockIdx.x] = r;
    r = blk_sum(s1, sh); if (threadIdx.x == 0) part[2 * blockIdx.x + 1] = r;
}

// sum over points of r r^T, r = (x2 x1, x2 y1, x2, y2 x1, y2 y1, y2, x1, y1, 1) in Hartley-normalised coordinates, where
// "1" = the FIRST argument of findFundamentalMat (the reference camera's points, triangulate.py:210) - upper triangle, 45 sums
__global__ void __launch_bounds__(kBlock) k_pair_moments(const float2 *__restrict__ m1, const float2 *__restrict__ m2, int64_t N,
                                                         const double *__restrict__ nrm /* c1x c1y s1 c2x c2y s2 */,
                                                         double *__restrict__ part /* grid x 45 */)
{
    __shared__ double sh[8];
    double acc[45];
#pragma unroll
    for (int t = 0; t < 45; ++t) acc[t] = 0.0;
    const double c1x = nrm[0], c1y = nrm[1], s1 = nrm[2], c2x = nrm[3], c2y = nrm[4], s2 = nrm[5];
    for (int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x; i < N; i += (int64_t)gridDim.x * kBlock) {
        const float2 a = m1[i], b = m2[i];
        const double x1 = ((double)a.x - c1x) * s1, y1 = ((double)a.y - c1y) * s1;
        const double x2 = ((double)b.x - c2x) * s2, y2 = ((double)b.y - c2y) * s2;
        const double r[9] = {x2 * x1, x2 * y1, x2, y2 * x1, y2 * y1, y2, x1, y1, 1.0};
        int t = 0;
#pragma unroll
        for (int p = 0; p < 9; ++p)
#pragma unroll
            for (int q = p; q < 9; ++q) acc[t++] += r[p] * r[q];
    }
#pragma unroll 1
    for (int t = 0; t < 45; ++t) { const double r = blk_sum(acc[t], sh); if (threadIdx.x == 0) part[45 * blockIdx.x + t] = r; }
}

__global__ void k_pair_reduce(const double *__restrict__ part, int nparts, int len, double *__restrict__ out)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= len) return;
    double s = 0.0;
    for (int q = 0; q < nparts; ++q) s += part[(size_t)q * len + t];
    out[t] = s;
}

// null vector of the 4x4 system of cv2.triangulatePoints(Pa, Pb, xa, xb) (rows x P[2] - P[0], y P[2] - P[1] per view; OpenCV
// takes the last right singular vector), de-homogenised as the reference does (tmp * (1 / tmp[3]), triangulate.py:137-138)
__device__ __forceinline__ void two_view(const double *__restrict__ Pa, const double *__restrict__ Pb, double xa, double ya, double xb,
                                         double yb, double *X)
{
    double R[4][4], G[4][4], V[4][4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        R[0][k] = xa * Pa[8 + k] - Pa[k]; R[1][k] = ya * Pa[8 + k] - Pa[4 + k];
        R[2][k] = xb * Pb[8 + k] - Pb[k]; R[3][k] = yb * Pb[8 + k] - Pb[4 + k];
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) G[i][j] = R[0][i] * R[0][j] + R[1][i] * R[1][j] + R[2][i] * R[2][j] + R[3][i] * R[3][j];
    jacobi_eig<4>(G, V);
    int best = 0;
    for (int k = 1; k < 4; ++k) if (G[k][k] < G[best][best]) best = k;
    const double iw = 1.0 / V[3][best];
    X[0] = V[0][best] * iw; X[1] = V[1][best] * iw; X[2] = V[2][best] * iw;
}

// the four hypotheses of one pair: (P_r, P1) and (P1, Pinv_r) for r = 0, 1; votes[h] += sign(z) (triangulate.py:147-155)
//   Pm: 4 matrices of 12: P_0, Pinv_0, P_1, Pinv_1;  P1 = [I | 0]
__global__ void __launch_bounds__(128) k_pair_vote(const float2 *__restrict__ x1, const float2 *__restrict__ x2, int64_t N,
                                                   const double *__restrict__ Pm, long long *__restrict__ votes)
{
    __shared__ double sP[48];
    __shared__ int sv[4];
    if (threadIdx.x < 48) sP[threadIdx.x] = Pm[threadIdx.x];
    if (threadIdx.x < 4) sv[threadIdx.x] = 0;
    __syncthreads();
    const double P1[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};
    int v[4] = {0, 0, 0, 0};
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
        const float2 a = x1[i], b = x2[i];
        double X[3];
#pragma unroll 1
        for (int h = 0; h < 4; ++h) {
            if (h & 1) two_view(P1, sP + 12 * h, (double)a.x, (double)a.y, (double)b.x, (double)b.y, X);
            else two_view(sP + 12 * h, P1, (double)a.x, (double)a.y, (double)b.x, (double)b.y, X);
            v[h] += (X[2] > 0.0) ? 1 : -1;
        }
    }
    for (int h = 0; h < 4; ++h) atomicAdd(&sv[h], v[h]);            // integers: order-independent
    __syncthreads();
    if (threadIdx.x < 4) atomicAdd((unsigned long long *)&votes[threadIdx.x], (unsigned long long)(long long)sv[threadIdx.x]);
}

// the chosen hypothesis, scaled by sgn (= -sign(votes), triangulate.py:207)
__global__ void __launch_bounds__(128) k_pair_emit(const float2 *__restrict__ x1, const float2 *__restrict__ x2, int64_t N,
                                                   const double *__restrict__ Pm, int h, double sgn, double *__restrict__ xyz)
{
    __shared__ double sP[12];
    if (threadIdx.x < 12) sP[threadIdx.x] = Pm[12 * h + threadIdx.x];
    __syncthreads();
    const double P1[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
        const float2 a = x1[i], b = x2[i];
        double X[3];
        if (h & 1) two_view(P1, sP, (double)a.x, (double)a.y, (double)b.x, (double)b.y, X);
        else two_view(sP, P1, (double)a.x, (double)a.y, (double)b.x, (double)b.y, X);
        xyz[3 * i] = X[0] * sgn; xyz[3 * i + 1] = X[1] * sgn; xyz[3 * i + 2] = X[2] * sgn;
    }
}

// ---- small host linear algebra -------------------------------------------------------------------------------------
struct M3 { double a[3][3]; };
M3 mul(const M3 &x, const M3 &y) { M3 r; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { r.a[i][j] = 0; for (int k = 0; k < 3; ++k) r.a[i][j] += x.a[i][k] * y.a[k][j]; } return r; }
M3 tr(const M3 &x) { M3 r; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) r.a[i][j] = x.a[j][i]; return r; }
double det(const M3 &x)
{
    return x.a[0][0] * (x.a[1][1] * x.a[2][2] - x.a[1][2] * x.a[2][1]) - x.a[0][1] * (x.a[1][0] * x.a[2][2] - x.a[1][2] * x.a[2][0]) +
           x.a[0][2] * (x.a[1][0] * x.a[2][1] - x.a[1][1] * x.a[2][0]);
}
M3 inv(const M3 &x)
{
    const double d = det(x); M3 r;
    r.a[0][0] = (x.a[1][1] * x.a[2][2] - x.a[1][2] * x.a[2][1]) / d; r.a[0][1] = (x.a[0][2] * x.a[2][1] - x.a[0][1] * x.a[2][2]) / d;
    r.a[0][2] = (x.a[0][1] * x.a[1][2] - x.a[0][2] * x.a[1][1]) / d; r.a[1][0] = (x.a[1][2] * x.a[2][0] - x.a[1][0] * x.a[2][2]) / d;
    r.a[1][1] = (x.a[0][0] * x.a[2][2] - x.a[0][2] * x.a[2][0]) / d; r.a[1][2] = (x.a[0][2] * x.a[1][0] - x.a[0][0] * x.a[1][2]) / d;
    r.a[2][0] = (x.a[1][0] * x.a[2][1] - x.a[1][1] * x.a[2][0]) / d; r.a[2][1] = (x.a[0][1] * x.a[2][0] - x.a[0][0] * x.a[2][1]) / d;
    r.a[2][2] = (x.a[0][0] * x.a[1][1] - x.a[0][1] * x.a[1][0]) / d;
    return r;
}

// thin SVD of a 3x3 matrix through the eigen-decomposition of F^T F: singular values descending, V columns, U columns
// (u_i = F v_i / s_i for the two leading ones, u_3 = u_1 x u_2: F has rank 2 here, so its third pair is a free choice)
void svd3_rank2(const M3 &F, M3 &U, double s[3], M3 &V)
{
    double G[3][3], E[3][3];
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { G[i][j] = 0; for (int k = 0; k < 3; ++k) G[i][j] += F.a[k][i] * F.a[k][j]; }
    jacobi_eig<3>(G, E);
    int ord[3] = {0, 1, 2};
    for (int i = 0; i < 3; ++i) for (int j = i + 1; j < 3; ++j) if (G[ord[j]][ord[j]] > G[ord[i]][ord[i]]) { int t = ord[i]; ord[i] = ord[j]; ord[j] = t; }
    for (int c = 0; c < 3; ++c) { s[c] = sqrt(fmax(G[ord[c]][ord[c]], 0.0)); for (int r = 0; r < 3; ++r) V.a[r][c] = E[r][ord[c]]; }
    for (int c = 0; c < 2; ++c)
        for (int r = 0; r < 3; ++r) { double t = 0; for (int k = 0; k < 3; ++k) t += F.a[r][k] * V.a[k][c]; U.a[r][c] = t / s[c]; }
    U.a[0][2] = U.a[1][0] * U.a[2][1] - U.a[2][0] * U.a[1][1];
    U.a[1][2] = U.a[2][0] * U.a[0][1] - U.a[0][0] * U.a[2][1];
    U.a[2][2] = U.a[0][0] * U.a[1][1] - U.a[1][0] * U.a[0][1];
}

struct DevScratch {
    void *p[8] = {0}; int n = 0;
    template <typename T> cudaError_t alloc(T **out, size_t cnt) { cudaError_t e = cudaMalloc(out, (cnt ? cnt : 1) * sizeof(T)); if (e == cudaSuccess) p[n++] = *out; return e; }
    ~DevScratch() { for (int i = 0; i < n; ++i) cudaFree(p[i]); }
};

}  // namespace

extern "C" int argus_pair_triangulate(const double *p1, const double *p2, int64_t N, const double *prof1, const double *prof2,
                                      double *xyz, double *R_out, double *t_out, double *F_out)
{
    if (!p1 || !p2 || !prof1 || !prof2 || !xyz || !R_out || !t_out || N < 8) return ARGUS_E_ARG;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) {
        fprintf(stderr, "argus_b200: no CUDA device available (there is no CPU fallback)\n");
        return ARGUS_E_CUDA;
    }
    const int grid = (int)((N + kBlock - 1) / kBlock < kMaxGrid ? (N + kBlock - 1) / kBlock : kMaxGrid);
    DevScratch sc;
    double *d_p1, *d_p2, *d_prof, *d_part, *d_small, *d_xyz, *d_P; float2 *d_x1, *d_x2; long long *d_votes;
    PAIR_CUDA_CHECK(sc.alloc(&d_p1, (size_t)2 * N)); PAIR_CUDA_CHECK(sc.alloc(&d_p2, (size_t)2 * N));
    PAIR_CUDA_CHECK(sc.alloc(&d_x1, (size_t)N)); PAIR_CUDA_CHECK(sc.alloc(&d_x2, (size_t)N));
    PAIR_CUDA_CHECK(sc.alloc(&d_part, (size_t)grid * 45)); PAIR_CUDA_CHECK(sc.alloc(&d_small, 128 + 48));
    PAIR_CUDA_CHECK(sc.alloc(&d_xyz, (size_t)3 * N)); PAIR_CUDA_CHECK(sc.alloc(&d_votes, 4));
    d_prof = d_small + 64; d_P = d_small + 128;
    double hprof[16];
    for (int q = 0; q < 8; ++q) { hprof[q] = prof1[q]; hprof[8 + q] = prof2[q]; }
    PAIR_CUDA_CHECK(cudaMemcpy(d_p1, p1, (size_t)2 * N * sizeof(double), cudaMemcpyHostToDevice));
    PAIR_CUDA_CHECK(cudaMemcpy(d_p2, p2, (size_t)2 * N * sizeof(double), cudaMemcpyHostToDevice));
    PAIR_CUDA_CHECK(cudaMemcpy(d_prof, hprof, sizeof(hprof), cudaMemcpyHostToDevice));
    PAIR_CUDA_CHECK(cudaStreamSynchronize(cudaStreamLegacy));

    // ---- normalise, centroids, scales, moments ----
    double h[64];
    k_pair_undistort<<<grid, kBlock>>>(d_p1, d_p2, N, d_prof, d_x1, d_x2, d_part);
    k_pair_reduce<<<1, 64>>>(d_part, grid, 4, d_small);
    PAIR_CUDA_CHECK(cudaMemcpy(h, d_small, 4 * sizeof(double), cudaMemcpyDeviceToHost));
    const double tcount = 1.0 / (double)N;
    double cen[4] = {h[0] * tcount, h[1] * tcount, h[2] * tcount, h[3] * tcount};     // camera-k x, y; reference-camera x, y
    PAIR_CUDA_CHECK(cudaMemcpy(d_small, cen, sizeof(cen), cudaMemcpyHostToDevice));
    k_pair_scale<<<grid, kBlock>>>(d_x1, d_x2, N, d_small, d_part);
    k_pair_reduce<<<1, 64>>>(d_part, grid, 2, d_small + 8);
    PAIR_CUDA_CHECK(cudaMemcpy(h, d_small + 8, 2 * sizeof(double), cudaMemcpyDeviceToHost));
    double sA = h[0] * tcount, sB = h[1] * tcount;
    if (sA < 1.1920929e-07 || sB < 1.1920929e-07) return ARGUS_E_ARG;               // FLT_EPSILON: degenerate point set
    sA = sqrt(2.0) / sA; sB = sqrt(2.0) / sB;
    // findFundamentalMat(points1 = p2 (reference camera), points2 = p1 (camera k)): run8Point's "m1" is the reference camera
    double nrm[6] = {cen[2], cen[3], sB, cen[0], cen[1], sA};
    PAIR_CUDA_CHECK(cudaMemcpy(d_small, nrm, sizeof(nrm), cudaMemcpyHostToDevice));
    k_pair_moments<<<grid, kBlock>>>(d_x2, d_x1, N, d_small, d_part);
    k_pair_reduce<<<1, 64>>>(d_part, grid, 45, d_small + 16);
    PAIR_CUDA_CHECK(cudaMemcpy(h, d_small + 16, 45 * sizeof(double), cudaMemcpyDeviceToHost));
    PAIR_CUDA_CHECK(cudaGetLastError());

    // ---- 8-point: smallest eigenvector of the 9x9 moment matrix, rank 2, de-normalise (run8Point) ----
    double A9[9][9], V9[9][9];
    { int t = 0; for (int p = 0; p < 9; ++p) for (int q = p; q < 9; ++q) { A9[p][q] = h[t]; A9[q][p] = h[t]; ++t; } }
    jacobi_eig<9>(A9, V9);
    int small = 0, nz = 0;
    for (int k = 0; k < 9; ++k) { if (A9[k][k] < A9[small][small]) small = k; if (fabs(A9[k][k]) >= 2.220446049250313e-16) ++nz; }
    if (nz < 8) return ARGUS_E_ARG;                                                   // run8Point: fewer than 8 independent equations
    M3 F0;
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) F0.a[r][c] = V9[3 * r + c][small];
    {   // rank 2: remove the component along the smallest right singular vector (= U diag(w1, w2, 0) V^T)
        M3 U, V; double s[3];
        svd3_rank2(F0, U, s, V);
        double fv[3];
        for (int r = 0; r < 3; ++r) fv[r] = F0.a[r][0] * V.a[0][2] + F0.a[r][1] * V.a[1][2] + F0.a[r][2] * V.a[2][2];
        for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) F0.a[r][c] -= fv[r] * V.a[c][2];
    }
    M3 T1 = {{{nrm[2], 0, -nrm[2] * nrm[0]}, {0, nrm[2], -nrm[2] * nrm[1]}, {0, 0, 1}}};
    M3 T2 = {{{nrm[5], 0, -nrm[5] * nrm[3]}, {0, nrm[5], -nrm[5] * nrm[4]}, {0, 0, 1}}};
    M3 F = mul(mul(tr(T2), F0), T1);
    if (fabs(F.a[2][2]) > 1.1920929e-07) { const double k = 1.0 / F.a[2][2]; for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) F.a[r][c] *= k; }
    if (F_out) for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) F_out[3 * r + c] = F.a[r][c];

    // ---- decomposition: candidates U W^T V^T, U W V^T, and with -V^T; keep the proper rotations (triangulate.py:69-93) ----
    M3 U, V; double sv[3];
    svd3_rank2(F, U, sv, V);
    const M3 W = {{{0, -1, 0}, {1, 0, 0}, {0, 0, 1}}};
    const M3 Vt = tr(V);
    M3 nVt = Vt; for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) nVt.a[r][c] = -Vt.a[r][c];
    M3 cands[4] = {mul(mul(U, tr(W)), Vt), mul(mul(U, W), Vt), mul(mul(U, tr(W)), nVt), mul(mul(U, W), nVt)};
    M3 Rs2[2] = {cands[0], cands[1]};
    int cnt = 0;
    for (int q = 0; q < 4; ++q) if (nearbyint(det(cands[q])) != -1.0) { if (cnt < 2) Rs2[cnt] = cands[q]; ++cnt; }
    const double t[3] = {U.a[0][2], U.a[1][2], U.a[2][2]};
    double Pm[48];
    for (int r = 0; r < 2; ++r) {
        const M3 &R = Rs2[r];
        const M3 Ri = inv(R);
        double *P = Pm + 24 * r, *Pi = Pm + 24 * r + 12;
        for (int a = 0; a < 3; ++a) {
            for (int b = 0; b < 3; ++b) { P[4 * a + b] = R.a[a][b]; Pi[4 * a + b] = Ri.a[a][b]; }
            P[4 * a + 3] = t[a];
            Pi[4 * a + 3] = -(t[0] * R.a[0][a] + t[1] * R.a[1][a] + t[2] * R.a[2][a]);      // (-t^T R)_a
        }
    }
    PAIR_CUDA_CHECK(cudaMemcpy(d_P, Pm, sizeof(Pm), cudaMemcpyHostToDevice));
    PAIR_CUDA_CHECK(cudaMemset(d_votes, 0, 4 * sizeof(long long)));
    const int grid2 = (int)((N + 127) / 128 < kMaxGrid ? (N + 127) / 128 : kMaxGrid);
    k_pair_vote<<<grid2, 128>>>(d_x1, d_x2, N, d_P, d_votes);
    long long pl[4];
    PAIR_CUDA_CHECK(cudaMemcpy(pl, d_votes, sizeof(pl), cudaMemcpyDeviceToHost));
    // ---- decision logic (triangulate.py:162-207) ----
    auto ab = [](long long v) { return v < 0 ? -v : v; };
    int zdx;
    if (ab(pl[0]) > ab(pl[2])) zdx = 0;
    else if (ab(pl[0]) < ab(pl[2])) zdx = 2;
    else zdx = (ab(pl[0] + pl[1]) > ab(pl[2] + pl[3])) ? 0 : 2;
    const M3 &Rsel = Rs2[zdx / 2];
    const double tsign = (pl[zdx] < 0) ? 1.0 : -1.0;
    const double osign = (pl[zdx] > 0) ? -1.0 : ((pl[zdx] < 0) ? 1.0 : 0.0);               // -1 * sign(plusses)
    k_pair_emit<<<grid2, 128>>>(d_x1, d_x2, N, d_P, zdx, osign, d_xyz);
    PAIR_CUDA_CHECK(cudaGetLastError());
    PAIR_CUDA_CHECK(cudaMemcpy(xyz, d_xyz, (size_t)3 * N * sizeof(double), cudaMemcpyDeviceToHost));
    for (int r = 0; r < 3; ++r) { for (int c = 0; c < 3; ++c) R_out[3 * r + c] = Rsel.a[r][c]; t_out[r] = tsign * t[r]; }
    return ARGUS_OK;
}
