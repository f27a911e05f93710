// Batched N-camera DLT triangulation and reprojection error (fp64, one thread per (frame, track)).
//
// Replaces the per-frame Python loops of
//   uv_to_xyz         /root/reference/argus_gui/tools.py:296-337
//   get_repo_errors   /root/reference/argus_gui/tools.py:360-423
//   undistort_pts     /root/reference/argus_gui/tools.py:129-148  (pinhole branch: cv2.undistortPoints(src f32, K, dist, P=K))
//   reconstruct_uv    /root/reference/argus_gui/tools.py:64-76
//
// Faithfulness notes:
//  * undistortion goes through float32 exactly as the reference does: the pixel is rounded to float32, OpenCV's
//    5 fixed-point iterations run in double WITHOUT fused multiply-add (x86 baseline build), the result is
//    re-projected with P = K and rounded to float32.  The device code uses __dmul_rn/__dadd_rn so nvcc cannot
//    contract the expressions.
//  * uv_to_xyz solves the (n+1)-row system the reference actually builds (rows k and k+1 are overwritten in the
//    loop at tools.py:317-328): the u-equation of every visible camera plus the v-equation of the LAST visible one.
//    flags bit 0 selects the textbook 2n-row system instead.
//  * exact zeros become NaN (tools.py:336, 422).
// HBM-bound: 88 B (triangulate) / 96 B (error) per (frame, track) at m = 4; no shared memory needed, the camera
// tables (m x 19 doubles) live in constant-cached global loads.
#include <cuda_runtime.h>
#include <curand_kernel.h>
#include <stdio.h>
#include <math.h>
#include "../../include/argus_b200.h"

namespace {

#define DLT_CUDA_CHECK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
    fprintf(stderr, "argus_b200: CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
    return ARGUS_E_CUDA; } } while (0)

constexpr int kMaxCams = ARGUS_BA_MAX_CAMS;

struct DltCams {            // by-value kernel argument would be 64*19*8 = 9.7 KB > 4 KB limit -> device memory
    const double *prof;     // m x 8: f, cx, cy, k1, k2, p1, p2, k3
    const double *dlt;      // m x 11
};

// cv2.undistortPoints(src(float32), K, dist=[k1,k2,p1,p2,k3], P=K), default criteria (5 iterations), float32 out.
__device__ __forceinline__ void undistort_f32(double pu, double pv, const double *__restrict__ pr, float &ou, float &ov)
{
    const double fx = pr[0], cx = pr[1], cy = pr[2];
    const double k1 = pr[3], k2 = pr[4], p1 = pr[5], p2 = pr[6], k3 = pr[7];
    const double u = (double)(float)pu, v = (double)(float)pv;    // src is a float32 array
    const double ifx = 1.0 / fx;
    double x = __dmul_rn(__dadd_rn(u, -cx), ifx), y = __dmul_rn(__dadd_rn(v, -cy), ifx);
    const double x0 = x, y0 = y;
#pragma unroll 1
    for (int j = 0; j < 5; ++j) {
        const double r2 = __dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y));
        // icdist = (1 + ((0*r2 + 0)*r2 + 0)*r2) / (1 + ((k3*r2 + k2)*r2 + k1)*r2)
        const double den = __dadd_rn(1.0, __dmul_rn(__dadd_rn(__dmul_rn(__dadd_rn(__dmul_rn(k3, r2), k2), r2), k1), r2));
        const double icdist = 1.0 / den;
        if (icdist < 0.0) { x = x0; y = y0; break; }
        // deltaX = 2*p1*x*y + p2*(r2 + 2*x*x);  deltaY = p1*(r2 + 2*y*y) + 2*p2*x*y   (left-to-right as in C)
        const double dX = __dadd_rn(__dmul_rn(__dmul_rn(__dmul_rn(2.0, p1), x), y),
                                    __dmul_rn(p2, __dadd_rn(r2, __dmul_rn(__dmul_rn(2.0, x), x))));
        const double dY = __dadd_rn(__dmul_rn(p1, __dadd_rn(r2, __dmul_rn(__dmul_rn(2.0, y), y))),
                                    __dmul_rn(__dmul_rn(__dmul_rn(2.0, p2), x), y));
        x = __dmul_rn(__dadd_rn(x0, -dX), icdist);
        y = __dmul_rn(__dadd_rn(y0, -dY), icdist);
    }
    // RR = P = K: xx = fx*x + 0*y + cx, yy = 0*x + fx*y + cy, ww = 1/(0*x + 0*y + 1)
    const double xx = __dadd_rn(__dadd_rn(__dmul_rn(fx, x), __dmul_rn(0.0, y)), cx);
    const double yy = __dadd_rn(__dadd_rn(__dmul_rn(0.0, x), __dmul_rn(fx, y)), cy);
    ou = (float)xx; ov = (float)yy;
}

__global__ void k_undistort(const double *__restrict__ pts, int64_t N, const double *__restrict__ prof8, float *__restrict__ out)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    float a, b;
    undistort_f32(pts[2 * i], pts[2 * i + 1], prof8, a, b);
    out[2 * i] = a; out[2 * i + 1] = b;
}

// accumulate one equation row (a . X = b) into the 3x3 normal equations
__device__ __forceinline__ void ne_add(double a0, double a1, double a2, double b, double *M, double *r)
{
    M[0] += a0 * a0; M[1] += a0 * a1; M[2] += a0 * a2; M[3] += a1 * a1; M[4] += a1 * a2; M[5] += a2 * a2;
    r[0] += a0 * b; r[1] += a1 * b; r[2] += a2 * b;
}

// triangulate one (frame, track); returns false when fewer than two cameras see the point
__device__ __forceinline__ bool triangulate_one(const double *__restrict__ row, int m, DltCams cams, int flags, double *X)
{
    double M[6] = {0, 0, 0, 0, 0, 0}, r[3] = {0, 0, 0};
    int nvis = 0;
    float lv = 0.f; int lj = -1;
    for (int j = 0; j < m; ++j) {
        const double2 uv = *reinterpret_cast<const double2 *>(row + 2 * j);
        if (isnan(uv.x) || isnan(uv.y)) continue;
        float fu, fv;
        undistort_f32(uv.x, uv.y, cams.prof + 8 * j, fu, fv);
        const double *L = cams.dlt + 11 * j;
        const double du = (double)fu;
        ne_add(du * L[8] - L[0], du * L[9] - L[1], du * L[10] - L[2], L[3] - du, M, r);
        if (flags & 1) {
            const double dv = (double)fv;
            ne_add(dv * L[8] - L[4], dv * L[9] - L[5], dv * L[10] - L[6], L[7] - dv, M, r);
        }
        lv = fv; lj = j; ++nvis;
    }
    if (nvis < 2) return false;
    if (!(flags & 1)) {     // the reference's system: only the last visible camera contributes its v-equation
        const double *L = cams.dlt + 11 * lj;
        const double dv = (double)lv;
        ne_add(dv * L[8] - L[4], dv * L[9] - L[5], dv * L[10] - L[6], L[7] - dv, M, r);
    }
    // 3x3 Cholesky solve of the normal equations (the systems are well conditioned: cond(A) ~ 4 on camera rings)
    const double l00 = sqrt(M[0]);
    const double l10 = M[1] / l00, l20 = M[2] / l00;
    const double l11 = sqrt(M[3] - l10 * l10);
    const double l21 = (M[4] - l20 * l10) / l11;
    const double l22 = sqrt(M[5] - l20 * l20 - l21 * l21);
    const double y0 = r[0] / l00;
    const double y1 = (r[1] - l10 * y0) / l11;
    const double y2 = (r[2] - l20 * y0 - l21 * y1) / l22;
    X[2] = y2 / l22;
    X[1] = (y1 - l21 * X[2]) / l11;
    X[0] = (y0 - l10 * X[1] - l20 * X[2]) / l00;
    return true;
}

// pts: N x (2 m k); xyz: N x (3 k); thread t -> (frame = t / k, track = t % k)
__global__ void __launch_bounds__(256) k_dlt_triangulate(const double *__restrict__ pts, int64_t N, int m, int k, DltCams cams,
                                                         double *__restrict__ xyz, int flags)
{
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= N * k) return;
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    double X[3];
    const bool ok = triangulate_one(pts + t * 2 * m, m, cams, flags, X);
    double *o = xyz + 3 * t;
#pragma unroll
    for (int q = 0; q < 3; ++q) o[q] = (ok && X[q] != 0.0) ? X[q] : nan;
}

__device__ __forceinline__ double reproj_one(const double *__restrict__ row, const double *__restrict__ X, int m, DltCams cams)
{
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    if (isnan(X[0]) || isnan(X[1]) || isnan(X[2])) return nan;
    double eps = 0.0; int nc = 0;
    for (int j = 0; j < m; ++j) {
        const double2 uv = *reinterpret_cast<const double2 *>(row + 2 * j);
        if (isnan(uv.x)) continue;          // only u is tested (tools.py:386)
        float fu, fv;
        undistort_f32(uv.x, uv.y, cams.prof + 8 * j, fu, fv);
        const double *L = cams.dlt + 11 * j;
        const double den = L[8] * X[0] + L[9] * X[1] + L[10] * X[2] + 1.0;
        const double ru = (L[0] * X[0] + L[1] * X[1] + L[2] * X[2] + L[3]) / den;
        const double rv = (L[4] * X[0] + L[5] * X[1] + L[6] * X[2] + L[7]) / den;
        const double d0 = (double)fu - ru, d1 = (double)fv - rv;
        eps += d0 * d0 + d1 * d1;
        ++nc;
    }
    double e;
    if (nc == 2) e = sqrt(eps / 2.0);
    else e = sqrt(eps / (double)(2 * nc - 3));      // nc < 2 -> negative dof -> NaN (or -0 -> NaN below)
    return (e == 0.0) ? nan : e;
}

// xyz: N x 3k, pts: N x 2mk, err: k x N
// (ldN = row length of err: the whole job's frame count when a chunk of frames is processed)
__global__ void __launch_bounds__(256) k_dlt_reproj(const double *__restrict__ xyz, const double *__restrict__ pts, int64_t N, int m, int k,
                                                    DltCams cams, double *__restrict__ err, int64_t ldN)
{
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= N * k) return;
    const int64_t frame = t / k; const int track = (int)(t % k);
    const double X[3] = {xyz[3 * t], xyz[3 * t + 1], xyz[3 * t + 2]};
    err[(int64_t)track * ldN + frame] = reproj_one(pts + t * 2 * m, X, m, cams);
}

// Fused triangulation + reprojection error (argus_dlt_reconstruct): every pixel is undistorted once instead of twice.
// Identical results to k_dlt_triangulate followed by k_dlt_reproj (same per-camera arithmetic in the same order).
template <int MMAX>
__global__ void __launch_bounds__(256) k_dlt_fused(const double *__restrict__ pts, int64_t N, int m, int k, DltCams cams,
                                                   double *__restrict__ xyz, double *__restrict__ err, int flags, int64_t ldN)
{
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= N * k) return;
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    const double *row = pts + t * 2 * m;
    float fu[MMAX], fv[MMAX];
    unsigned long long seen_u = 0ull, seen_uv = 0ull;
    double M[6] = {0, 0, 0, 0, 0, 0}, r[3] = {0, 0, 0};
    int nvis = 0, lj = -1;
#pragma unroll
    for (int j = 0; j < MMAX; ++j) {
        if (j < m) {
            const double2 uv = *reinterpret_cast<const double2 *>(row + 2 * j);
            fu[j] = 0.f; fv[j] = 0.f;
            if (!isnan(uv.x)) {
                seen_u |= 1ull << j;
                undistort_f32(uv.x, uv.y, cams.prof + 8 * j, fu[j], fv[j]);
                if (!isnan(uv.y)) {
                    seen_uv |= 1ull << j;
                    const double *L = cams.dlt + 11 * j;
                    const double du = (double)fu[j];
                    ne_add(du * L[8] - L[0], du * L[9] - L[1], du * L[10] - L[2], L[3] - du, M, r);
                    if (flags & 1) {
                        const double dv = (double)fv[j];
                        ne_add(dv * L[8] - L[4], dv * L[9] - L[5], dv * L[10] - L[6], L[7] - dv, M, r);
                    }
                    lj = j; ++nvis;
                }
            }
        }
    }
    double X[3] = {nan, nan, nan};
    if (nvis >= 2) {
        if (!(flags & 1)) {
            float lv = 0.f;
#pragma unroll
            for (int j = 0; j < MMAX; ++j) if (j == lj) lv = fv[j];
            const double *L = cams.dlt + 11 * lj;
            const double dv = (double)lv;
            ne_add(dv * L[8] - L[4], dv * L[9] - L[5], dv * L[10] - L[6], L[7] - dv, M, r);
        }
        const double l00 = sqrt(M[0]);
        const double l10 = M[1] / l00, l20 = M[2] / l00;
        const double l11 = sqrt(M[3] - l10 * l10);
        const double l21 = (M[4] - l20 * l10) / l11;
        const double l22 = sqrt(M[5] - l20 * l20 - l21 * l21);
        const double y0 = r[0] / l00;
        const double y1 = (r[1] - l10 * y0) / l11;
        const double y2 = (r[2] - l20 * y0 - l21 * y1) / l22;
        X[2] = y2 / l22;
        X[1] = (y1 - l21 * X[2]) / l11;
        X[0] = (y0 - l10 * X[1] - l20 * X[2]) / l00;
#pragma unroll
        for (int q = 0; q < 3; ++q) if (X[q] == 0.0) X[q] = nan;
    }
    double *o = xyz + 3 * t;
    o[0] = X[0]; o[1] = X[1]; o[2] = X[2];
    double e = nan;
    if (!(isnan(X[0]) || isnan(X[1]) || isnan(X[2]))) {
        double eps = 0.0; int nc = 0;
#pragma unroll
        for (int j = 0; j < MMAX; ++j) {
            if (j < m && ((seen_u >> j) & 1ull)) {
                const double *L = cams.dlt + 11 * j;
                const double den = L[8] * X[0] + L[9] * X[1] + L[10] * X[2] + 1.0;
                const double ru = (L[0] * X[0] + L[1] * X[1] + L[2] * X[2] + L[3]) / den;
                const double rv = (L[4] * X[0] + L[5] * X[1] + L[6] * X[2] + L[7]) / den;
                const double d0 = (double)fu[j] - ru, d1 = (double)fv[j] - rv;
                eps += d0 * d0 + d1 * d1;
                ++nc;
            }
        }
        e = (nc == 2) ? sqrt(eps / 2.0) : sqrt(eps / (double)(2 * nc - 3));
        if (e == 0.0) e = nan;
    }
    const int64_t frame = t / k; const int track = (int)(t % k);
    err[(int64_t)track * ldN + frame] = e;
}

// Bootstrap confidence intervals of the triangulated points (bootstrapXYZs, /root/reference/argus_gui/tools.py:448-544, the
// 250-iteration loop at :513-527): every iteration perturbs the pixels of a (frame, track) with N(0, sigma^2),
// sigma = rmse * sqrt(2) / (finite pixel coordinates / 2), re-triangulates, and the sample standard deviation (ddof = 1,
// NaN results ignored) over the iterations is returned.  One thread per (frame, track) runs all iterations (Welford).
// noise != nullptr: explicit standard-normal numbers in the reference's draw order [track][iteration][frame][2m] (exact
// parity tests); otherwise Philox (curand) seeded per thread.
__global__ void __launch_bounds__(128) k_dlt_bootstrap(const double *__restrict__ pts, int64_t N, int m, int k, DltCams cams,
                                                       const double *__restrict__ rmses, int bs_iter, const double *__restrict__ noise,
                                                       unsigned long long seed, double *__restrict__ sd)
{
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= N * k) return;
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    const int64_t frame = t / k; const int track = (int)(t % k);
    const double *row = pts + t * 2 * m;
    int nfin = 0;
    for (int q = 0; q < 2 * m; ++q) nfin += isfinite(row[q]) ? 1 : 0;
    const double sigma = rmses[frame * k + track] * 1.4142135623730951 / (0.5 * (double)nfin);
    curandStatePhilox4_32_10_t st;
    if (!noise) curand_init(seed, (unsigned long long)t, 0ull, &st);
    double mean[3] = {0, 0, 0}, m2[3] = {0, 0, 0}; int cnt[3] = {0, 0, 0};
    for (int it = 0; it < bs_iter; ++it) {
        double M[6] = {0, 0, 0, 0, 0, 0}, r[3] = {0, 0, 0};
        int nvis = 0, lj = -1; float lv = 0.f;
        const double *nz = noise ? noise + (((int64_t)track * bs_iter + it) * N + frame) * 2 * m : nullptr;
        for (int j = 0; j < m; ++j) {
            double n0, n1;
            if (nz) { n0 = nz[2 * j]; n1 = nz[2 * j + 1]; }
            else { const double2 g = curand_normal2_double(&st); n0 = g.x; n1 = g.y; }
            const double pu = n0 * sigma + row[2 * j], pv = n1 * sigma + row[2 * j + 1];
            if (isnan(pu) || isnan(pv)) continue;
            float fu, fv;
            undistort_f32(pu, pv, cams.prof + 8 * j, fu, fv);
            const double *L = cams.dlt + 11 * j;
            const double du = (double)fu;
            ne_add(du * L[8] - L[0], du * L[9] - L[1], du * L[10] - L[2], L[3] - du, M, r);
            lv = fv; lj = j; ++nvis;
        }
        if (nvis < 2) continue;
        {
            const double *L = cams.dlt + 11 * lj;
            const double dv = (double)lv;
            ne_add(dv * L[8] - L[4], dv * L[9] - L[5], dv * L[10] - L[6], L[7] - dv, M, r);
        }
        const double l00 = sqrt(M[0]);
        const double l10 = M[1] / l00, l20 = M[2] / l00;
        const double l11 = sqrt(M[3] - l10 * l10);
        const double l21 = (M[4] - l20 * l10) / l11;
        const double l22 = sqrt(M[5] - l20 * l20 - l21 * l21);
        const double y0 = r[0] / l00;
        const double y1 = (r[1] - l10 * y0) / l11;
        const double y2 = (r[2] - l20 * y0 - l21 * y1) / l22;
        double X[3];
        X[2] = y2 / l22;
        X[1] = (y1 - l21 * X[2]) / l11;
        X[0] = (y0 - l10 * X[1] - l20 * X[2]) / l00;
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            if (X[q] == 0.0 || isnan(X[q])) continue;           // zeros are NaN in the reference (tools.py:336) and nanstd skips them
            cnt[q] += 1;
            const double d = X[q] - mean[q];
            mean[q] += d / (double)cnt[q];
            m2[q] += d * (X[q] - mean[q]);
        }
    }
#pragma unroll
    for (int q = 0; q < 3; ++q) {
        double v = (cnt[q] > 1) ? sqrt(m2[q] / (double)(cnt[q] - 1)) : nan;
        if (v == 0.0) v = nan;
        sd[3 * t + q] = v;
    }
}

int launch_tri(const double *d_pts, int64_t N, int m, int k, const double *d_prof, const double *d_dlt, double *d_xyz, int flags, cudaStream_t s)
{
    const int64_t T = N * k;
    if (T == 0) return ARGUS_OK;
    DltCams cams{d_prof, d_dlt};
    k_dlt_triangulate<<<(unsigned)((T + 255) / 256), 256, 0, s>>>(d_pts, N, m, k, cams, d_xyz, flags);
    DLT_CUDA_CHECK(cudaGetLastError());
    return ARGUS_OK;
}
int launch_err(const double *d_xyz, const double *d_pts, int64_t N, int m, int k, const double *d_prof, const double *d_dlt, double *d_err, cudaStream_t s,
               int64_t ldN = -1)
{
    if (ldN < 0) ldN = N;
    const int64_t T = N * k;
    if (T == 0) return ARGUS_OK;
    DltCams cams{d_prof, d_dlt};
    k_dlt_reproj<<<(unsigned)((T + 255) / 256), 256, 0, s>>>(d_xyz, d_pts, N, m, k, cams, d_err, ldN);
    DLT_CUDA_CHECK(cudaGetLastError());
    return ARGUS_OK;
}

int launch_fused(const double *d_pts, int64_t N, int m, int k, const double *d_prof, const double *d_dlt, double *d_xyz, double *d_err,
                 int flags, cudaStream_t s, int64_t ldN = -1)
{
    if (ldN < 0) ldN = N;
    const int64_t T = N * k;
    if (T == 0) return ARGUS_OK;
    DltCams cams{d_prof, d_dlt};
    const unsigned grid = (unsigned)((T + 255) / 256);
    if (m <= 4) k_dlt_fused<4><<<grid, 256, 0, s>>>(d_pts, N, m, k, cams, d_xyz, d_err, flags, ldN);
    else if (m <= 8) k_dlt_fused<8><<<grid, 256, 0, s>>>(d_pts, N, m, k, cams, d_xyz, d_err, flags, ldN);
    else if (m <= 16) k_dlt_fused<16><<<grid, 256, 0, s>>>(d_pts, N, m, k, cams, d_xyz, d_err, flags, ldN);
    else {      // many cameras: the two-kernel path (no per-thread camera arrays)
        int rc = launch_tri(d_pts, N, m, k, d_prof, d_dlt, d_xyz, flags, s);
        if (rc) return rc;
        return launch_err(d_xyz, d_pts, N, m, k, d_prof, d_dlt, d_err, s, ldN);
    }
    DLT_CUDA_CHECK(cudaGetLastError());
    return ARGUS_OK;
}

bool have_device()
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) {
        fprintf(stderr, "argus_b200: no CUDA device available (there is no CPU fallback)\n");
        return false;
    }
    return true;
}

struct Scratch {
    void *p[8] = {0}; int n = 0;
    template <typename T> cudaError_t alloc(T **out, size_t cnt) { cudaError_t e = cudaMalloc(out, (cnt ? cnt : 1) * sizeof(T)); if (e == cudaSuccess) p[n++] = *out; return e; }
    ~Scratch() { for (int i = 0; i < n; ++i) cudaFree(p[i]); }
};

// Host-pointer calls move the job through the device in chunks of frames on three streams, so that the upload of chunk c+1,
// the kernel of chunk c and the download of chunk c-1 overlap (PCIe is full duplex; the kernel is far shorter than either copy).
struct ChunkPipe {
    static constexpr int kStreams = 3;
    cudaStream_t s[kStreams] = {nullptr, nullptr, nullptr};
    cudaError_t init() { for (auto &q : s) { cudaError_t e = cudaStreamCreateWithFlags(&q, cudaStreamNonBlocking); if (e != cudaSuccess) return e; } return cudaSuccess; }
    ~ChunkPipe() { for (auto &q : s) if (q) { cudaStreamSynchronize(q); cudaStreamDestroy(q); } }
    cudaError_t finish() { cudaError_t r = cudaSuccess; for (auto &q : s) { cudaError_t e = cudaStreamSynchronize(q); if (e != cudaSuccess) r = e; } return r; }
    static int64_t chunk_frames(int64_t N, size_t bytes_per_frame) {
        int64_t c = (int64_t)((size_t)(24u << 20) / (bytes_per_frame ? bytes_per_frame : 1));
        if (c < 1024) c = 1024;
        return c < N ? c : (N > 0 ? N : 1);
    }
};

// err is k x N on both sides: a chunk of frames is a k-row strided piece
cudaError_t copy_err_chunk(double *err, const double *d_err, int64_t N, int k, int64_t f0, int64_t nf, cudaStream_t s)
{
    return cudaMemcpy2DAsync(err + f0, (size_t)N * sizeof(double), d_err + f0, (size_t)N * sizeof(double), (size_t)nf * sizeof(double), (size_t)k,
                             cudaMemcpyDeviceToHost, s);
}

}  // namespace

extern "C" {

int argus_dlt_triangulate_dev(const double *d_pts, int64_t N, int m, int k, const double *d_prof, const double *d_dlt,
                              double *d_xyz, int flags, void *stream)
{
    if (!d_pts || !d_prof || !d_dlt || !d_xyz || m < 1 || m > kMaxCams || k < 1 || N < 0) return ARGUS_E_ARG;
    return launch_tri(d_pts, N, m, k, d_prof, d_dlt, d_xyz, flags, (cudaStream_t)stream);
}

int argus_dlt_reproj_error_dev(const double *d_xyz, const double *d_pts, int64_t N, int m, int k, const double *d_prof,
                               const double *d_dlt, double *d_err, void *stream)
{
    if (!d_xyz || !d_pts || !d_prof || !d_dlt || !d_err || m < 1 || m > kMaxCams || k < 1 || N < 0) return ARGUS_E_ARG;
    return launch_err(d_xyz, d_pts, N, m, k, d_prof, d_dlt, d_err, (cudaStream_t)stream);
}

int argus_dlt_reconstruct_dev(const double *d_pts, int64_t N, int m, int k, const double *d_prof, const double *d_dlt, double *d_xyz,
                              double *d_err, int flags, void *stream)
{
    if (!d_pts || !d_prof || !d_dlt || !d_xyz || !d_err || m < 1 || m > kMaxCams || k < 1 || N < 0) return ARGUS_E_ARG;
    return launch_fused(d_pts, N, m, k, d_prof, d_dlt, d_xyz, d_err, flags, (cudaStream_t)stream);
}

int argus_dlt_triangulate(const double *pts, int64_t N, int m, int k, const double *prof, const double *dlt, double *xyz, int flags)
{
    if (!pts || !prof || !dlt || !xyz || m < 1 || m > kMaxCams || k < 1 || N < 0) return ARGUS_E_ARG;
    if (!have_device()) return ARGUS_E_CUDA;
    Scratch sc; double *d_pts, *d_prof, *d_dlt, *d_xyz;
    DLT_CUDA_CHECK(sc.alloc(&d_pts, (size_t)N * 2 * m * k)); DLT_CUDA_CHECK(sc.alloc(&d_prof, (size_t)m * 8));
    DLT_CUDA_CHECK(sc.alloc(&d_dlt, (size_t)m * 11)); DLT_CUDA_CHECK(sc.alloc(&d_xyz, (size_t)N * 3 * k));
    DLT_CUDA_CHECK(cudaMemcpy(d_prof, prof, (size_t)m * 8 * sizeof(double), cudaMemcpyHostToDevice));
    DLT_CUDA_CHECK(cudaMemcpy(d_dlt, dlt, (size_t)m * 11 * sizeof(double), cudaMemcpyHostToDevice));
    DLT_CUDA_CHECK(cudaStreamSynchronize(cudaStreamLegacy));      // the pipe's streams do not order against the legacy stream
    ChunkPipe pipe; DLT_CUDA_CHECK(pipe.init());
    const size_t wp = (size_t)2 * m * k, wx = (size_t)3 * k;
    const int64_t cf = ChunkPipe::chunk_frames(N, wp * sizeof(double));
    int q = 0;
    for (int64_t f0 = 0; f0 < N; f0 += cf, q = (q + 1) % ChunkPipe::kStreams) {
        const int64_t nf = (N - f0 < cf) ? N - f0 : cf;
        cudaStream_t s = pipe.s[q];
        DLT_CUDA_CHECK(cudaMemcpyAsync(d_pts + f0 * wp, pts + f0 * wp, nf * wp * sizeof(double), cudaMemcpyHostToDevice, s));
        int rc = launch_tri(d_pts + f0 * wp, nf, m, k, d_prof, d_dlt, d_xyz + f0 * wx, flags, s);
        if (rc) return rc;
        DLT_CUDA_CHECK(cudaMemcpyAsync(xyz + f0 * wx, d_xyz + f0 * wx, nf * wx * sizeof(double), cudaMemcpyDeviceToHost, s));
    }
    DLT_CUDA_CHECK(pipe.finish());
    return ARGUS_OK;
}

int argus_dlt_reproj_error(const double *xyz, const double *pts, int64_t N, int m, int k, const double *prof, const double *dlt, double *err)
{
    if (!xyz || !pts || !prof || !dlt || !err || m < 1 || m > kMaxCams || k < 1 || N < 0) return ARGUS_E_ARG;
    if (!have_device()) return ARGUS_E_CUDA;
    Scratch sc; double *d_pts, *d_prof, *d_dlt, *d_xyz, *d_err;
    DLT_CUDA_CHECK(sc.alloc(&d_pts, (size_t)N * 2 * m * k)); DLT_CUDA_CHECK(sc.alloc(&d_prof, (size_t)m * 8));
    DLT_CUDA_CHECK(sc.alloc(&d_dlt, (size_t)m * 11)); DLT_CUDA_CHECK(sc.alloc(&d_xyz, (size_t)N * 3 * k));
    DLT_CUDA_CHECK(sc.alloc(&d_err, (size_t)N * k));
    DLT_CUDA_CHECK(cudaMemcpy(d_prof, prof, (size_t)m * 8 * sizeof(double), cudaMemcpyHostToDevice));
    DLT_CUDA_CHECK(cudaMemcpy(d_dlt, dlt, (size_t)m * 11 * sizeof(double), cudaMemcpyHostToDevice));
    DLT_CUDA_CHECK(cudaStreamSynchronize(cudaStreamLegacy));      // the pipe's streams do not order against the legacy stream
    ChunkPipe pipe; DLT_CUDA_CHECK(pipe.init());
    const size_t wp = (size_t)2 * m * k, wx = (size_t)3 * k;
    const int64_t cf = ChunkPipe::chunk_frames(N, (wp + wx) * sizeof(double));
    int q = 0;
    for (int64_t f0 = 0; f0 < N; f0 += cf, q = (q + 1) % ChunkPipe::kStreams) {
        const int64_t nf = (N - f0 < cf) ? N - f0 : cf;
        cudaStream_t s = pipe.s[q];
        DLT_CUDA_CHECK(cudaMemcpyAsync(d_pts + f0 * wp, pts + f0 * wp, nf * wp * sizeof(double), cudaMemcpyHostToDevice, s));
        DLT_CUDA_CHECK(cudaMemcpyAsync(d_xyz + f0 * wx, xyz + f0 * wx, nf * wx * sizeof(double), cudaMemcpyHostToDevice, s));
        int rc = launch_err(d_xyz + f0 * wx, d_pts + f0 * wp, nf, m, k, d_prof, d_dlt, d_err + f0, s, N);
        if (rc) return rc;
        DLT_CUDA_CHECK(copy_err_chunk(err, d_err, N, k, f0, nf, s));
    }
    DLT_CUDA_CHECK(pipe.finish());
    return ARGUS_OK;
}

int argus_dlt_reconstruct(const double *pts, int64_t N, int m, int k, const double *prof, const double *dlt, double *xyz, double *err, int flags)
{
    if (!pts || !prof || !dlt || !xyz || !err || m < 1 || m > kMaxCams || k < 1 || N < 0) return ARGUS_E_ARG;
    if (!have_device()) return ARGUS_E_CUDA;
    Scratch sc; double *d_pts, *d_prof, *d_dlt, *d_xyz, *d_err;
    DLT_CUDA_CHECK(sc.alloc(&d_pts, (size_t)N * 2 * m * k)); DLT_CUDA_CHECK(sc.alloc(&d_prof, (size_t)m * 8));
    DLT_CUDA_CHECK(sc.alloc(&d_dlt, (size_t)m * 11)); DLT_CUDA_CHECK(sc.alloc(&d_xyz, (size_t)N * 3 * k));
    DLT_CUDA_CHECK(sc.alloc(&d_err, (size_t)N * k));
    DLT_CUDA_CHECK(cudaMemcpy(d_prof, prof, (size_t)m * 8 * sizeof(double), cudaMemcpyHostToDevice));
    DLT_CUDA_CHECK(cudaMemcpy(d_dlt, dlt, (size_t)m * 11 * sizeof(double), cudaMemcpyHostToDevice));
    DLT_CUDA_CHECK(cudaStreamSynchronize(cudaStreamLegacy));      // the pipe's streams do not order against the legacy stream
    ChunkPipe pipe; DLT_CUDA_CHECK(pipe.init());
    const size_t wp = (size_t)2 * m * k, wx = (size_t)3 * k;
    const int64_t cf = ChunkPipe::chunk_frames(N, wp * sizeof(double));
    int q = 0;
    for (int64_t f0 = 0; f0 < N; f0 += cf, q = (q + 1) % ChunkPipe::kStreams) {
        const int64_t nf = (N - f0 < cf) ? N - f0 : cf;
        cudaStream_t s = pipe.s[q];
        DLT_CUDA_CHECK(cudaMemcpyAsync(d_pts + f0 * wp, pts + f0 * wp, nf * wp * sizeof(double), cudaMemcpyHostToDevice, s));
        int rc = launch_fused(d_pts + f0 * wp, nf, m, k, d_prof, d_dlt, d_xyz + f0 * wx, d_err + f0, flags, s, N);
        if (rc) return rc;
        DLT_CUDA_CHECK(cudaMemcpyAsync(xyz + f0 * wx, d_xyz + f0 * wx, nf * wx * sizeof(double), cudaMemcpyDeviceToHost, s));
        DLT_CUDA_CHECK(copy_err_chunk(err, d_err, N, k, f0, nf, s));
    }
    DLT_CUDA_CHECK(pipe.finish());
    return ARGUS_OK;
}

int argus_dlt_bootstrap(const double *pts, int64_t N, int m, int k, const double *prof, const double *dlt, const double *rmses, int bs_iter,
                        const double *noise, uint64_t seed, double *sd)
{
    if (!pts || !prof || !dlt || !rmses || !sd || m < 1 || m > kMaxCams || k < 1 || N < 0 || bs_iter < 1) return ARGUS_E_ARG;
    if (!have_device()) return ARGUS_E_CUDA;
    if (N == 0) return ARGUS_OK;
    Scratch sc; double *d_pts, *d_prof, *d_dlt, *d_rm, *d_sd, *d_noise = nullptr;
    const size_t npx = (size_t)N * 2 * m * k;
    DLT_CUDA_CHECK(sc.alloc(&d_pts, npx)); DLT_CUDA_CHECK(sc.alloc(&d_prof, (size_t)m * 8)); DLT_CUDA_CHECK(sc.alloc(&d_dlt, (size_t)m * 11));
    DLT_CUDA_CHECK(sc.alloc(&d_rm, (size_t)N * k)); DLT_CUDA_CHECK(sc.alloc(&d_sd, (size_t)N * 3 * k));
    DLT_CUDA_CHECK(cudaMemcpy(d_pts, pts, npx * sizeof(double), cudaMemcpyHostToDevice));
    DLT_CUDA_CHECK(cudaMemcpy(d_prof, prof, (size_t)m * 8 * sizeof(double), cudaMemcpyHostToDevice));
    DLT_CUDA_CHECK(cudaMemcpy(d_dlt, dlt, (size_t)m * 11 * sizeof(double), cudaMemcpyHostToDevice));
    DLT_CUDA_CHECK(cudaMemcpy(d_rm, rmses, (size_t)N * k * sizeof(double), cudaMemcpyHostToDevice));
    if (noise) {
        DLT_CUDA_CHECK(sc.alloc(&d_noise, npx * (size_t)bs_iter));
        DLT_CUDA_CHECK(cudaMemcpy(d_noise, noise, npx * (size_t)bs_iter * sizeof(double), cudaMemcpyHostToDevice));
    }
    DltCams cams{d_prof, d_dlt};
    const int64_t T = N * k;
    k_dlt_bootstrap<<<(unsigned)((T + 127) / 128), 128>>>(d_pts, N, m, k, cams, d_rm, bs_iter, d_noise, (unsigned long long)seed, d_sd);
    DLT_CUDA_CHECK(cudaGetLastError());
    DLT_CUDA_CHECK(cudaMemcpy(sd, d_sd, (size_t)N * 3 * k * sizeof(double), cudaMemcpyDeviceToHost));
    return ARGUS_OK;
}

int argus_undistort_points(const double *pts, int64_t N, const double *prof8, float *out)
{
    if (!pts || !prof8 || !out || N < 0) return ARGUS_E_ARG;
    if (!have_device()) return ARGUS_E_CUDA;
    if (N == 0) return ARGUS_OK;
    Scratch sc; double *d_pts, *d_prof; float *d_out;
    DLT_CUDA_CHECK(sc.alloc(&d_pts, (size_t)N * 2)); DLT_CUDA_CHECK(sc.alloc(&d_prof, 8)); DLT_CUDA_CHECK(sc.alloc(&d_out, (size_t)N * 2));
    DLT_CUDA_CHECK(cudaMemcpy(d_pts, pts, (size_t)N * 2 * sizeof(double), cudaMemcpyHostToDevice));
    DLT_CUDA_CHECK(cudaMemcpy(d_prof, prof8, 8 * sizeof(double), cudaMemcpyHostToDevice));
    k_undistort<<<(unsigned)((N + 255) / 256), 256>>>(d_pts, N, d_prof, d_out);
    DLT_CUDA_CHECK(cudaGetLastError());
    DLT_CUDA_CHECK(cudaMemcpy(out, d_out, (size_t)N * 2 * sizeof(float), cudaMemcpyDeviceToHost));
    return ARGUS_OK;
}

}  // extern "C"
