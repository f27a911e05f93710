// Motion-only and structure-only Levenberg-Marquardt: the reference's sba_mot_levmar_x
// (sba-1.6p/sba_levmar.c:1437-1980) and sba_str_levmar_x (:1987-2540).  Both are block-diagonal: with the points frozen
// every camera has its own 16x16 system (U_j + mu I) da_j = ea_j (:1838-1852), with the cameras frozen every point its own
// 3x3 system (V_i + mu I) db_i = eb_i (:2383-2397) - no Schur complement.  U_j / ea_j / V_i / eb_i and the statistics come
// from the tile lineariser (k_lin_tile, mode 0) run on a problem whose other half is marked fixed; the trial evaluation
// reuses k_residual / k_backsub_recompute.  Only the small solves below are specific to these drivers.
#pragma once
#include "ba_common.cuh"

namespace argus {

enum { kModeMotStr = 0, kModeMot = 1, kModeStr = 2 };

// scalars of one linearisation (:1774-1812 / :2319-2357): scal = {cost, ||J^T e||_inf, ||p||^2, max diagonal}, each over the
// moving half only (p is the camera vector for motion-only, the point vector for structure-only)
__global__ void k_bd_finalize(int mode, const double *__restrict__ Ured, ParamBufs pbuf, const LmState *__restrict__ st, int m, int mcon,
                              const double *__restrict__ sum2, const double *__restrict__ max2, double *__restrict__ scal)
{
    if (threadIdx.x != 0 || blockIdx.x != 0 || st->done) return;
    const double *cam = pb_cam(pbuf, st->cur_cam);
    double inf = 0.0, dmax = 2.2250738585072014e-308 /* DBL_MIN */, psq = 0.0;
    if (mode == kModeMot) {
        for (int j = mcon; j < m; ++j) {
            const double *Uj = Ured + j * kU;
            for (int r = 0; r < 16; ++r) { inf = fmax(inf, fabs(Uj[136 + r])); dmax = fmax(dmax, Uj[tri16(r, r)]); }
        }
        for (int t = 0; t < 16 * m; ++t) psq += cam[t] * cam[t];
    } else {
        inf = max2[0]; dmax = fmax(dmax, max2[1]); psq = sum2[1];
    }
    scal[0] = sum2[0]; scal[1] = inf; scal[2] = psq; scal[3] = dmax;
}

// one thread per camera: Cholesky solve of (U_j + mu_diag I) da_j = ea_j (sba_Axb_Chol, sba_lapack.c:373-479: dpotrf 'U' on a
// copy, then dpotrs), trial cameras, ||da||^2 and dL = sum da (mu da + ea) (:1858-1895).  mu_diag is the damping that sits
// on the diagonal (it accumulates over rejected attempts in sba-1.6 compatibility mode), mu the current damping.
//   out2 = {||da||^2, dL}; flags[1] = 1 when every system was positive definite.
__global__ void __launch_bounds__(64) k_mot_solve(int m, int mcon, const double *__restrict__ Ured, ParamBufs pbuf,
                                                  const LmState *__restrict__ st, double *__restrict__ out2, int *__restrict__ flags)
{
    if (st->done) return;
    const double mu_diag = st->mu_u, mu = st->mu;
    const double *cam = pb_cam(pbuf, st->cur_cam);
    double *cam_new = pb_cam(pbuf, st->cur_cam ^ 1);
    __shared__ double s_d2[ARGUS_BA_MAX_CAMS], s_dl[ARGUS_BA_MAX_CAMS];
    __shared__ int s_ok[ARGUS_BA_MAX_CAMS];
    const int j = threadIdx.x;
    if (j < m) {
        double d2 = 0.0, dl = 0.0; int ok = 1;
        if (j < mcon) {
            for (int r = 0; r < 16; ++r) cam_new[16 * j + r] = cam[16 * j + r];
        } else {
            double R[136], y[16];
            const double *Uj = Ured + j * kU;
            for (int t = 0; t < 136; ++t) R[t] = Uj[t];
            for (int r = 0; r < 16; ++r) R[tri16(r, r)] += mu_diag;
            for (int c = 0; c < 16 && ok; ++c) {              // R^T R = A, R upper, column by column (dpotf2 order)
                double ajj = R[tri16(c, c)];
                for (int k = 0; k < c; ++k) ajj -= R[tri16(k, c)] * R[tri16(k, c)];
                if (!(ajj > 0.0)) { ok = 0; break; }
                ajj = sqrt(ajj);
                R[tri16(c, c)] = ajj;
                for (int cc = c + 1; cc < 16; ++cc) {
                    double s = R[tri16(c, cc)];
                    for (int k = 0; k < c; ++k) s -= R[tri16(k, c)] * R[tri16(k, cc)];
                    R[tri16(c, cc)] = s / ajj;
                }
            }
            if (ok) {
                for (int i = 0; i < 16; ++i) {                // R^T y = ea
                    double s = Uj[136 + i];
                    for (int k = 0; k < i; ++k) s -= R[tri16(k, i)] * y[k];
                    y[i] = s / R[tri16(i, i)];
                }
                for (int i = 15; i >= 0; --i) {               // R da = y
                    double s = y[i];
                    for (int k = i + 1; k < 16; ++k) s -= R[tri16(i, k)] * y[k];
                    y[i] = s / R[tri16(i, i)];
                }
                for (int r = 0; r < 16; ++r) {
                    const double d = y[r];
                    cam_new[16 * j + r] = cam[16 * j + r] + d;
                    d2 += d * d; dl += d * (mu * d + Uj[136 + r]);
                }
            } else {
                for (int r = 0; r < 16; ++r) cam_new[16 * j + r] = cam[16 * j + r];
            }
        }
        s_d2[j] = d2; s_dl[j] = dl; s_ok[j] = ok;
    }
    __syncthreads();
    if (threadIdx.x == 0) {                                   // fixed order: cameras ascending
        double a = 0.0, b = 0.0; int ok = 1;
        for (int q = 0; q < m; ++q) { a += s_d2[q]; b += s_dl[q]; ok &= s_ok[q]; }
        out2[0] = a; out2[1] = b; flags[1] = ok;
    }
}

// one thread per point: inverse of V_i + mu_diag I through LDL^T; a pivot that is not positive is what makes the reference's
// 3x3 dpotrf fail (-> the attempt is rejected, :2398).  Vinv feeds k_backsub_recompute with da = 0: db_i = Vinv_i eb_i.
__global__ void __launch_bounds__(256) k_str_invert(int64_t n, int64_t ncon, const LmState *__restrict__ st, const double *__restrict__ V,
                                                    double *__restrict__ Vinv, int *__restrict__ flags)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || i < ncon || st->done) return;
    const double mu_diag = st->mu_u;
    double vv[6], li[3], di[3];
    for (int t = 0; t < 6; ++t) vv[t] = V[6 * i + t];
    vv[0] += mu_diag; vv[3] += mu_diag; vv[5] += mu_diag;
    if (!sym3_ldl_inv(vv, li, di) || !(di[0] > 0.0) || !(di[1] > 0.0) || !(di[2] > 0.0)) { atomicOr(flags, 1); return; }
    double *Ni = Vinv + 6 * i;
    Ni[0] = di[0] + li[0] * li[0] * di[1] + li[1] * li[1] * di[2];
    Ni[1] = li[0] * di[1] + li[1] * li[2] * di[2];
    Ni[2] = li[1] * di[2];
    Ni[3] = di[1] + li[2] * li[2] * di[2];
    Ni[4] = li[2] * di[2];
    Ni[5] = di[2];
}

}  // namespace argus
