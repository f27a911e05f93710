// Argus / sba "KDRTS" camera model for one observation, fp64, host+device.
//
// Replaces (from-scratch derivation, not a transcription of the Maple-generated code):
//   calcDistImgProjFullR      sba-1.6p/libsbaprojs/imgproj.c:1206-1270   (projection)
//   calcDistImgProjJacKDRTS   sba-1.6p/libsbaprojs/imgproj.c:1273-1748   (Jacobian, A 2x16 / B 2x3)
//   _MK_QUAT_FRM_VEC + quatMultFast   sbaprojs.c:38-41, 61-93           (local (x) initial rotation)
//
// Camera parameter block (16 doubles, order fixed by sbaprojs.c:1158-1163):
//   a = [fu, u0, v0, ar, s | k1, k2, p1, p2, k3 | v0, v1, v2 (local rotation, vector part) | t0, t1, t2]
// plus a constant unit quaternion q0 (scalar first).  Model:
//   ql = (sqrt(1 - v.v), v),  q = ql (x) q0,  P = R(q) M + t
//   x = P0/P2, y = P1/P2, r2 = x^2 + y^2, rho = 1 + r2 (k1 + r2 (k2 + r2 k3))
//   xd = rho x + 2 p1 x y + p2 (r2 + 2 x^2),  yd = rho y + p1 (r2 + 2 y^2) + 2 p2 x y
//   u = fu xd + s yd + u0,  v = fu ar yd + v0
//
// Jacobian by the chain rule  (u,v) <- (xd,yd) <- (x,y) <- P <- {t, v, M}:
//   G = d(u,v)/dP (2x3);  dP/dt = I;  dP/dM = R;
//   dP/dv_a = w_a x (P - t)  with  w_a = 2 [ (v_a / w) v + w e_a + v x e_a ],  w = sqrt(1 - v.v)
//   (angular velocity of the left perturbation ql -> ql + dql; it does not depend on q0 or on M, so the
//    3x3 matrix Omega = [w_0 w_1 w_2] is a per-camera constant precomputed once per linearisation).
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define ARGUS_HD __host__ __device__ __forceinline__
#else
#define ARGUS_HD static inline
#endif

namespace argus {

// Per-camera constants derived from the 16 parameters + q0; computed once per kernel launch.
struct CamConst {
    double fu, u0, v0, ar, s;
    double k1, k2, p1, p2, k3;
    double R[9];      // rotation of q = ql (x) q0, row-major
    double t[3];
    double Om[9];     // Om[3*r + a] = component r of w_a
};

ARGUS_HD void cam_precompute(const double *a, const double *q0, CamConst &c)
{
    c.fu = a[0]; c.u0 = a[1]; c.v0 = a[2]; c.ar = a[3]; c.s = a[4];
    c.k1 = a[5]; c.k2 = a[6]; c.p1 = a[7]; c.p2 = a[8]; c.k3 = a[9];
    const double vx = a[10], vy = a[11], vz = a[12];
    c.t[0] = a[13]; c.t[1] = a[14]; c.t[2] = a[15];
    const double w = sqrt(1.0 - vx * vx - vy * vy - vz * vz);   // NaN if |v| > 1, as in the reference
    // q = (w, v) (x) q0   (Hamilton product)
    const double b0 = q0[0], b1 = q0[1], b2 = q0[2], b3 = q0[3];
    const double qw = w * b0 - vx * b1 - vy * b2 - vz * b3;
    const double qx = w * b1 + vx * b0 + vy * b3 - vz * b2;
    const double qy = w * b2 - vx * b3 + vy * b0 + vz * b1;
    const double qz = w * b3 + vx * b2 - vy * b1 + vz * b0;
    // rotation as the homogeneous quadratic form q (0,M) q^-1 (what the reference evaluates)
    const double ww = qw * qw, xx = qx * qx, yy = qy * qy, zz = qz * qz;
    c.R[0] = ww + xx - yy - zz;       c.R[1] = 2.0 * (qx * qy - qw * qz); c.R[2] = 2.0 * (qx * qz + qw * qy);
    c.R[3] = 2.0 * (qx * qy + qw * qz); c.R[4] = ww - xx + yy - zz;       c.R[5] = 2.0 * (qy * qz - qw * qx);
    c.R[6] = 2.0 * (qx * qz - qw * qy); c.R[7] = 2.0 * (qy * qz + qw * qx); c.R[8] = ww - xx - yy + zz;
    // Omega: w_a = 2 [ (v_a/w) v + w e_a + v x e_a ]
    const double iw = 1.0 / w;
    const double v[3] = {vx, vy, vz};
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int a_ = 0; a_ < 3; ++a_) {
        const double f = v[a_] * iw;
        double o0 = f * vx, o1 = f * vy, o2 = f * vz;
        // v x e_a
        if (a_ == 0) { o0 += w;  o1 += vz; o2 -= vy; }
        if (a_ == 1) { o0 -= vz; o1 += w;  o2 += vx; }
        if (a_ == 2) { o0 += vy; o1 -= vx; o2 += w; }
        c.Om[0 + a_] = 2.0 * o0; c.Om[3 + a_] = 2.0 * o1; c.Om[6 + a_] = 2.0 * o2;
    }
}

// Predicted pixel of point M in camera c.
ARGUS_HD void cam_project(const CamConst &c, const double M0, const double M1, const double M2, double &u, double &v)
{
    const double P0 = c.R[0] * M0 + c.R[1] * M1 + c.R[2] * M2 + c.t[0];
    const double P1 = c.R[3] * M0 + c.R[4] * M1 + c.R[5] * M2 + c.t[1];
    const double P2 = c.R[6] * M0 + c.R[7] * M1 + c.R[8] * M2 + c.t[2];
    const double iz = 1.0 / P2;
    const double x = P0 * iz, y = P1 * iz;
    const double r2 = x * x + y * y;
    const double rho = 1.0 + r2 * (c.k1 + r2 * (c.k2 + r2 * c.k3));
    const double a1 = 2.0 * x * y, a2 = r2 + 2.0 * x * x, a3 = r2 + 2.0 * y * y;
    const double xd = rho * x + c.p1 * a1 + c.p2 * a2;
    const double yd = rho * y + c.p1 * a3 + c.p2 * a1;
    u = c.fu * xd + c.s * yd + c.u0;
    v = c.fu * c.ar * yd + c.v0;
}

// Projection + Jacobian blocks.  A is 2x16 row-major (A[0..15] = d u / d a, A[16..31] = d v / d a),
// B is 2x3 row-major.  nfix_k trailing intrinsics and nfix_d trailing distortion columns are zeroed
// (sbaprojs.c:1229-1247: columns [5-nccalib,5) and 5+[5-ncdist,5)).
ARGUS_HD void cam_project_jac(const CamConst &c, const double M0, const double M1, const double M2,
                              const int nfix_k, const int nfix_d, double &u, double &v, double *A, double *B)
{
    const double Q0 = c.R[0] * M0 + c.R[1] * M1 + c.R[2] * M2;      // R M  (= P - t)
    const double Q1 = c.R[3] * M0 + c.R[4] * M1 + c.R[5] * M2;
    const double Q2 = c.R[6] * M0 + c.R[7] * M1 + c.R[8] * M2;
    const double P0 = Q0 + c.t[0], P1 = Q1 + c.t[1], P2 = Q2 + c.t[2];
    const double iz = 1.0 / P2;
    const double x = P0 * iz, y = P1 * iz;
    const double xx = x * x, yy = y * y, xy = x * y;
    const double r2 = xx + yy;
    const double r4 = r2 * r2, r6 = r4 * r2;
    const double rho = 1.0 + c.k1 * r2 + c.k2 * r4 + c.k3 * r6;
    const double drho = c.k1 + 2.0 * c.k2 * r2 + 3.0 * c.k3 * r4;   // d rho / d r2
    const double a1 = 2.0 * xy, a2 = r2 + 2.0 * xx, a3 = r2 + 2.0 * yy;
    const double xd = rho * x + c.p1 * a1 + c.p2 * a2;
    const double yd = rho * y + c.p1 * a3 + c.p2 * a1;
    const double fv = c.fu * c.ar;
    u = c.fu * xd + c.s * yd + c.u0;
    v = fv * yd + c.v0;

    // intrinsics
    A[0] = xd;  A[16] = c.ar * yd;      // fu
    A[1] = 1.0; A[17] = 0.0;            // u0
    A[2] = 0.0; A[18] = 1.0;            // v0
    A[3] = 0.0; A[19] = c.fu * yd;      // ar
    A[4] = yd;  A[20] = 0.0;            // s
    // distortion: d(xd,yd)/d kc, then through (fu, s; 0, fu ar)
    {
        const double dx0 = x * r2, dy0 = y * r2;       // k1
        const double dx1 = x * r4, dy1 = y * r4;       // k2
        const double dx4 = x * r6, dy4 = y * r6;       // k3
        A[5] = c.fu * dx0 + c.s * dy0; A[21] = fv * dy0;
        A[6] = c.fu * dx1 + c.s * dy1; A[22] = fv * dy1;
        A[7] = c.fu * a1 + c.s * a3;   A[23] = fv * a3;    // p1
        A[8] = c.fu * a2 + c.s * a1;   A[24] = fv * a1;    // p2
        A[9] = c.fu * dx4 + c.s * dy4; A[25] = fv * dy4;
    }
    // d(xd,yd)/d(x,y)
    const double dxdx = rho + 2.0 * xx * drho + 2.0 * c.p1 * y + 6.0 * c.p2 * x;
    const double dxdy = 2.0 * xy * drho + 2.0 * c.p1 * x + 2.0 * c.p2 * y;
    const double dydx = dxdy;   // 2xy drho + 2 p1 x + 2 p2 y  (the distortion map has a symmetric Jacobian)
    const double dydy = rho + 2.0 * yy * drho + 6.0 * c.p1 * y + 2.0 * c.p2 * x;
    // d(u,v)/d(x,y)
    const double ux = c.fu * dxdx + c.s * dydx, uy = c.fu * dxdy + c.s * dydy;
    const double vx = fv * dydx,                vy = fv * dydy;
    // G = d(u,v)/dP
    const double G00 = ux * iz, G01 = uy * iz, G02 = -(ux * x + uy * y) * iz;
    const double G10 = vx * iz, G11 = vy * iz, G12 = -(vx * x + vy * y) * iz;
    // translation
    A[13] = G00; A[14] = G01; A[15] = G02;
    A[29] = G10; A[30] = G11; A[31] = G12;
    // local rotation: dP/dv_a = w_a x Q
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int a_ = 0; a_ < 3; ++a_) {
        const double o0 = c.Om[a_], o1 = c.Om[3 + a_], o2 = c.Om[6 + a_];
        const double d0 = o1 * Q2 - o2 * Q1, d1 = o2 * Q0 - o0 * Q2, d2 = o0 * Q1 - o1 * Q0;
        A[10 + a_] = G00 * d0 + G01 * d1 + G02 * d2;
        A[26 + a_] = G10 * d0 + G11 * d1 + G12 * d2;
    }
    // structure: B = G R
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int b_ = 0; b_ < 3; ++b_) {
        B[b_]     = G00 * c.R[b_] + G01 * c.R[3 + b_] + G02 * c.R[6 + b_];
        B[3 + b_] = G10 * c.R[b_] + G11 * c.R[3 + b_] + G12 * c.R[6 + b_];
    }
    // fixed parameters (static indices + predicates: a run-time loop bound would push A into local memory)
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int k = 0; k < 5; ++k) {
        if (k >= 5 - nfix_k) { A[k] = 0.0; A[16 + k] = 0.0; }
        if (k >= 5 - nfix_d) { A[5 + k] = 0.0; A[21 + k] = 0.0; }
    }
}

// Projection-free directional derivative for the back-substitution: s = A da (the change of the predicted pixel along the
// camera step da[16]) and B, without forming the 32 entries of A (W_ij^T da_j = B_ij^T (A_ij da_j), sba_levmar.c:1215-1256).
// Fixed parameters: their da entries are exactly zero (zero gradient, mu on the diagonal), the counts are applied all the same.
ARGUS_HD void cam_step_dir(const CamConst &c, const double M0, const double M1, const double M2, const int nfix_k, const int nfix_d,
                           const double *da, double &su, double &sv, double *B)
{
    const double Q0 = c.R[0] * M0 + c.R[1] * M1 + c.R[2] * M2;
    const double Q1 = c.R[3] * M0 + c.R[4] * M1 + c.R[5] * M2;
    const double Q2 = c.R[6] * M0 + c.R[7] * M1 + c.R[8] * M2;
    const double P0 = Q0 + c.t[0], P1 = Q1 + c.t[1], P2 = Q2 + c.t[2];
    const double iz = 1.0 / P2;
    const double x = P0 * iz, y = P1 * iz;
    const double xx = x * x, yy = y * y, xy = x * y;
    const double r2 = xx + yy;
    const double r4 = r2 * r2, r6 = r4 * r2;
    const double rho = 1.0 + c.k1 * r2 + c.k2 * r4 + c.k3 * r6;
    const double drho = c.k1 + 2.0 * c.k2 * r2 + 3.0 * c.k3 * r4;
    const double a1 = 2.0 * xy, a2 = r2 + 2.0 * xx, a3 = r2 + 2.0 * yy;
    const double xd = rho * x + c.p1 * a1 + c.p2 * a2;
    const double yd = rho * y + c.p1 * a3 + c.p2 * a1;
    const double fv = c.fu * c.ar;
    double d[10];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int k = 0; k < 5; ++k) {
        d[k] = (k >= 5 - nfix_k) ? 0.0 : da[k];
        d[5 + k] = (k >= 5 - nfix_d) ? 0.0 : da[5 + k];
    }
    // intrinsics + distortion
    const double rad = r2 * d[5] + r4 * d[6] + r6 * d[9];
    const double dxd = x * rad + a1 * d[7] + a2 * d[8];
    const double dyd = y * rad + a3 * d[7] + a1 * d[8];
    su = xd * d[0] + d[1] + yd * d[4] + c.fu * dxd + c.s * dyd;
    sv = c.ar * yd * d[0] + d[2] + c.fu * yd * d[3] + fv * dyd;
    // pose: dP = dt + (Om dv) x Q
    const double w0 = c.Om[0] * da[10] + c.Om[1] * da[11] + c.Om[2] * da[12];
    const double w1 = c.Om[3] * da[10] + c.Om[4] * da[11] + c.Om[5] * da[12];
    const double w2 = c.Om[6] * da[10] + c.Om[7] * da[11] + c.Om[8] * da[12];
    const double dP0 = da[13] + w1 * Q2 - w2 * Q1, dP1 = da[14] + w2 * Q0 - w0 * Q2, dP2 = da[15] + w0 * Q1 - w1 * Q0;
    const double dxdx = rho + 2.0 * xx * drho + 2.0 * c.p1 * y + 6.0 * c.p2 * x;
    const double dxdy = 2.0 * xy * drho + 2.0 * c.p1 * x + 2.0 * c.p2 * y;
    const double dydy = rho + 2.0 * yy * drho + 6.0 * c.p1 * y + 2.0 * c.p2 * x;
    const double ux = c.fu * dxdx + c.s * dxdy, uy = c.fu * dxdy + c.s * dydy;
    const double vx = fv * dxdy,                vy = fv * dydy;
    const double G00 = ux * iz, G01 = uy * iz, G02 = -(ux * x + uy * y) * iz;
    const double G10 = vx * iz, G11 = vy * iz, G12 = -(vx * x + vy * y) * iz;
    su += G00 * dP0 + G01 * dP1 + G02 * dP2;
    sv += G10 * dP0 + G11 * dP1 + G12 * dP2;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int b_ = 0; b_ < 3; ++b_) {
        B[b_]     = G00 * c.R[b_] + G01 * c.R[3 + b_] + G02 * c.R[6 + b_];
        B[3 + b_] = G10 * c.R[b_] + G11 * c.R[3 + b_] + G12 * c.R[6 + b_];
    }
}

// Symmetric 3x3 factorisation V = L D L^T (unit lower L, no pivoting) of the damped point block.
// in: v = [v00 v01 v02 v11 v12 v22] (upper triangle).  out: li = [l10 l20 l21] of L^-1 (unit lower),
// di = 1/d.  Returns false when a pivot is exactly zero (the reference's "singular V*_i" branch,
// sba_lapack.c:1093-1116).  Replaces the LAPACK Bunch-Kaufman inverse sba_symat_invert_BK
// (sba_lapack.c:1037-1124); (V)^-1 = L^-T D^-1 L^-1.
ARGUS_HD bool sym3_ldl_inv(const double *v, double *li, double *di)
{
    const double d0 = v[0];
    if (d0 == 0.0) return false;
    const double i0 = 1.0 / d0;
    const double l10 = v[1] * i0, l20 = v[2] * i0;
    const double d1 = v[3] - l10 * v[1];
    if (d1 == 0.0) return false;
    const double i1 = 1.0 / d1;
    const double l21 = (v[4] - l20 * v[1]) * i1;
    const double d2 = v[5] - l20 * v[2] - l21 * l21 * d1;
    if (d2 == 0.0) return false;
    di[0] = i0; di[1] = i1; di[2] = 1.0 / d2;
    // L^-1 for unit lower L = [[1,0,0],[l10,1,0],[l20,l21,1]]
    li[0] = -l10;
    li[1] = l10 * l21 - l20;
    li[2] = -l21;
    return true;
}

// Same factorisation with one reciprocal square root per pivot instead of a division and a square root:
// rs[k] = 1/sqrt(|d_k|), di[k] = sign(d_k) rs[k]^2 (= 1/d_k to a few ulp).  Three dependent rsqrt chains are the whole
// critical path of the per-point phase of the tile kernel.
ARGUS_HD bool sym3_ldl_inv_rs(const double *v, double *li, double *di, double *rs)
{
    const double d0 = v[0];
    if (d0 == 0.0) return false;
#if defined(__CUDA_ARCH__)
    const double r0 = rsqrt(fabs(d0));
#else
    const double r0 = 1.0 / sqrt(fabs(d0));
#endif
    const double i0 = (d0 < 0.0) ? -(r0 * r0) : r0 * r0;
    const double l10 = v[1] * i0, l20 = v[2] * i0;
    const double d1 = v[3] - l10 * v[1];
    if (d1 == 0.0) return false;
#if defined(__CUDA_ARCH__)
    const double r1 = rsqrt(fabs(d1));
#else
    const double r1 = 1.0 / sqrt(fabs(d1));
#endif
    const double i1 = (d1 < 0.0) ? -(r1 * r1) : r1 * r1;
    const double l21 = (v[4] - l20 * v[1]) * i1;
    const double d2 = v[5] - l20 * v[2] - l21 * l21 * d1;
    if (d2 == 0.0) return false;
#if defined(__CUDA_ARCH__)
    const double r2 = rsqrt(fabs(d2));
#else
    const double r2 = 1.0 / sqrt(fabs(d2));
#endif
    di[0] = i0; di[1] = i1; di[2] = (d2 < 0.0) ? -(r2 * r2) : r2 * r2;
    rs[0] = r0; rs[1] = r1; rs[2] = r2;
    li[0] = -l10;
    li[1] = l10 * l21 - l20;
    li[2] = -l21;
    return true;
}

}  // namespace argus
