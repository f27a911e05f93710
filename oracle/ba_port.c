/* TEST INFRASTRUCTURE - CPU restatement ("port") of the reference's bundle-adjustment algorithm.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may build, link or
 * call this file.  It is never on the product path.
 *
 * It restates, in plain scalar C and written independently of both the reference sources and the CUDA kernels:
 *
 *   orc_project      img_projsKDRTS_x + calcDistImgProjFullR     sba-1.6p/libsbaprojs/sbaprojs.c:1135-1184,
 *                                                                 imgproj.c:1206-1270
 *   orc_jacobian     img_projsKDRTS_jac_x + calcDistImgProjJacKDRTS + column zeroing
 *                                                                 sbaprojs.c:1193-1250, imgproj.c:1273-1748
 *                    (here by forward-mode automatic differentiation of the projection formula: 19 dual
 *                     components, no hand-derived or Maple-generated derivative code)
 *   orc_motstr       sba_motstr_levmar_x                          sba-1.6p/sba_levmar.c:449-1428
 *                    - U/ea, V/eb, W                              :837-938
 *                    - ||J^T e||_inf, ||p||^2, initial mu         :941-987
 *                    - damping, (V*_i)^-1 into V's lower triangle :992-1016  (sba_symat_invert_BK, sba_lapack.c:1037-1124)
 *                    - Y_ij, S, E                                 :1018-1185
 *                    - Cholesky solve                             :1201       (sba_Axb_Chol, sba_lapack.c:373-479)
 *                    - db_i, step tests, gain ratio, mu update    :1215-1337
 *                    - info[]                                     :1375-1399
 *                    including the two behaviours after a rejected step: the U/V diagonals are
 *                    NOT restored (the restore block at :1339-1353 is #if 0) and V's diagonal holds the diagonal of the
 *                    inverse, because the inverse is written over the lower triangle *including the diagonal*
 *                    (sba_lapack.c:1118-1121).  The same memory layout is used here so both fall out naturally.
 *                    compat = 0 restores the diagonals instead (textbook LM).
 *
 * Parity status: pinned.  tests/test_oracle_ba.py checks this file against (a) the compiled reference (oracle/_ref)
 * on synthetic scenes - projection, Jacobian, and LM trajectories; (b) tests/golden/ba_golden.npz, produced by the
 * compiled reference (tests/golden/make_ba_golden.py); (c) the reference's README known answers on its own demo data.
 *
 * Build: gcc -O2 -fPIC -shared -o oracle/_build/libba_port.so oracle/ba_port.c -lm
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>

#define CNP 16
#define PNP 3
#define ND 19 /* dual components: 16 camera + 3 point */

/* ---------- forward-mode AD scalar ---------- */
typedef struct { double v; double d[ND]; } dual;

static dual d_const(double c) { dual r; r.v = c; memset(r.d, 0, sizeof r.d); return r; }
static dual d_var(double c, int k) { dual r = d_const(c); r.d[k] = 1.0; return r; }
static dual d_add(dual a, dual b) { dual r; r.v = a.v + b.v; for (int k = 0; k < ND; ++k) r.d[k] = a.d[k] + b.d[k]; return r; }
static dual d_sub(dual a, dual b) { dual r; r.v = a.v - b.v; for (int k = 0; k < ND; ++k) r.d[k] = a.d[k] - b.d[k]; return r; }
static dual d_mul(dual a, dual b) { dual r; r.v = a.v * b.v; for (int k = 0; k < ND; ++k) r.d[k] = a.d[k] * b.v + a.v * b.d[k]; return r; }
static dual d_div(dual a, dual b) { dual r; r.v = a.v / b.v; for (int k = 0; k < ND; ++k) r.d[k] = (a.d[k] - r.v * b.d[k]) / b.v; return r; }
static dual d_sqrt(dual a) { dual r; r.v = sqrt(a.v); for (int k = 0; k < ND; ++k) r.d[k] = a.d[k] / (2.0 * r.v); return r; }
static dual d_scale(dual a, double s) { dual r; r.v = a.v * s; for (int k = 0; k < ND; ++k) r.d[k] = a.d[k] * s; return r; }

/* the camera model of SURVEY.md appendix A, generic over the scalar type via two instantiations */
static void project_dual(const double *a, const double *q0, const double *M, dual *u, dual *v)
{
    dual fu = d_var(a[0], 0), u0 = d_var(a[1], 1), v0 = d_var(a[2], 2), ar = d_var(a[3], 3), s = d_var(a[4], 4);
    dual k1 = d_var(a[5], 5), k2 = d_var(a[6], 6), p1 = d_var(a[7], 7), p2 = d_var(a[8], 8), k3 = d_var(a[9], 9);
    dual lv[3] = {d_var(a[10], 10), d_var(a[11], 11), d_var(a[12], 12)};
    dual t[3] = {d_var(a[13], 13), d_var(a[14], 14), d_var(a[15], 15)};
    dual X[3] = {d_var(M[0], 16), d_var(M[1], 17), d_var(M[2], 18)};
    /* ql = (sqrt(1 - |lv|^2), lv) */
    dual w = d_sqrt(d_sub(d_const(1.0), d_add(d_add(d_mul(lv[0], lv[0]), d_mul(lv[1], lv[1])), d_mul(lv[2], lv[2]))));
    dual ql[4] = {w, lv[0], lv[1], lv[2]};
    dual b[4] = {d_const(q0[0]), d_const(q0[1]), d_const(q0[2]), d_const(q0[3])};
    /* q = ql (x) q0 */
    dual q[4];
    q[0] = d_sub(d_sub(d_sub(d_mul(ql[0], b[0]), d_mul(ql[1], b[1])), d_mul(ql[2], b[2])), d_mul(ql[3], b[3]));
    q[1] = d_sub(d_add(d_add(d_mul(ql[0], b[1]), d_mul(ql[1], b[0])), d_mul(ql[2], b[3])), d_mul(ql[3], b[2]));
    q[2] = d_add(d_add(d_sub(d_mul(ql[0], b[2]), d_mul(ql[1], b[3])), d_mul(ql[2], b[0])), d_mul(ql[3], b[1]));
    q[3] = d_add(d_sub(d_add(d_mul(ql[0], b[3]), d_mul(ql[1], b[2])), d_mul(ql[2], b[1])), d_mul(ql[3], b[0]));
    /* P = q (0,X) q^-1 + t  evaluated as two quaternion products */
    dual zero = d_const(0.0);
    dual r[4]; /* r = q (x) (0, X) */
    r[0] = d_sub(d_sub(d_sub(zero, d_mul(q[1], X[0])), d_mul(q[2], X[1])), d_mul(q[3], X[2]));
    r[1] = d_sub(d_add(d_mul(q[0], X[0]), d_mul(q[2], X[2])), d_mul(q[3], X[1]));
    r[2] = d_add(d_sub(d_mul(q[0], X[1]), d_mul(q[1], X[2])), d_mul(q[3], X[0]));
    r[3] = d_sub(d_add(d_mul(q[0], X[2]), d_mul(q[1], X[1])), d_mul(q[2], X[0]));
    /* P = vec( r (x) conj(q) ) + t,  conj(q) = (q0, -q1, -q2, -q3) */
    dual P[3];
    P[0] = d_add(d_sub(d_add(d_sub(zero, d_mul(r[0], q[1])), d_mul(r[1], q[0])), d_mul(r[2], q[3])), d_mul(r[3], q[2]));
    P[1] = d_sub(d_add(d_add(d_sub(zero, d_mul(r[0], q[2])), d_mul(r[1], q[3])), d_mul(r[2], q[0])), d_mul(r[3], q[1]));
    P[2] = d_add(d_add(d_sub(d_sub(zero, d_mul(r[0], q[3])), d_mul(r[1], q[2])), d_mul(r[2], q[1])), d_mul(r[3], q[0]));
    for (int k = 0; k < 3; ++k) P[k] = d_add(P[k], t[k]);
    dual x = d_div(P[0], P[2]), y = d_div(P[1], P[2]);
    dual r2 = d_add(d_mul(x, x), d_mul(y, y));
    dual rho = d_add(d_const(1.0), d_mul(r2, d_add(k1, d_mul(r2, d_add(k2, d_mul(r2, k3))))));
    dual xy = d_mul(x, y);
    dual xd = d_add(d_add(d_mul(rho, x), d_scale(d_mul(p1, xy), 2.0)), d_mul(p2, d_add(r2, d_scale(d_mul(x, x), 2.0))));
    dual yd = d_add(d_add(d_mul(rho, y), d_mul(p1, d_add(r2, d_scale(d_mul(y, y), 2.0)))), d_scale(d_mul(p2, xy), 2.0));
    *u = d_add(d_add(d_mul(fu, xd), d_mul(s, yd)), u0);
    *v = d_add(d_mul(d_mul(fu, ar), yd), v0);
}

/* value-only version (same formula, plain doubles) */
static void project_val(const double *a, const double *q0, const double *M, double *out)
{
    double lv0 = a[10], lv1 = a[11], lv2 = a[12];
    double w = sqrt(1.0 - lv0 * lv0 - lv1 * lv1 - lv2 * lv2);
    double q[4];
    q[0] = w * q0[0] - lv0 * q0[1] - lv1 * q0[2] - lv2 * q0[3];
    q[1] = w * q0[1] + lv0 * q0[0] + lv1 * q0[3] - lv2 * q0[2];
    q[2] = w * q0[2] - lv0 * q0[3] + lv1 * q0[0] + lv2 * q0[1];
    q[3] = w * q0[3] + lv0 * q0[2] - lv1 * q0[1] + lv2 * q0[0];
    double r0 = -q[1] * M[0] - q[2] * M[1] - q[3] * M[2];
    double r1 = q[0] * M[0] + q[2] * M[2] - q[3] * M[1];
    double r2_ = q[0] * M[1] - q[1] * M[2] + q[3] * M[0];
    double r3 = q[0] * M[2] + q[1] * M[1] - q[2] * M[0];
    double P0 = -r0 * q[1] + r1 * q[0] - r2_ * q[3] + r3 * q[2] + a[13];
    double P1 = -r0 * q[2] + r1 * q[3] + r2_ * q[0] - r3 * q[1] + a[14];
    double P2 = -r0 * q[3] - r1 * q[2] + r2_ * q[1] + r3 * q[0] + a[15];
    double x = P0 / P2, y = P1 / P2, rr = x * x + y * y;
    double rho = 1.0 + rr * (a[5] + rr * (a[6] + rr * a[9]));
    double xd = rho * x + 2.0 * a[7] * x * y + a[8] * (rr + 2.0 * x * x);
    double yd = rho * y + a[7] * (rr + 2.0 * y * y) + 2.0 * a[8] * x * y;
    out[0] = a[0] * xd + a[4] * yd + a[1];
    out[1] = a[0] * a[3] * yd + a[2];
}

/* ---------- visibility index: point-major CSR (the reference's CRS, sba_levmar.c:639-650) ---------- */
typedef struct { int64_t n; int m; int64_t nobs; int64_t *ptr; int *cam; } vis_t;

static int vis_build(vis_t *vs, int64_t n, int m, const char *vmask)
{
    vs->n = n; vs->m = m;
    vs->ptr = (int64_t *)malloc((n + 1) * sizeof(int64_t));
    int64_t k = 0;
    for (int64_t i = 0; i < n; ++i) { vs->ptr[i] = k; for (int j = 0; j < m; ++j) if (vmask[i * m + j]) ++k; }
    vs->ptr[n] = k; vs->nobs = k;
    vs->cam = (int *)malloc((k ? k : 1) * sizeof(int));
    k = 0;
    for (int64_t i = 0; i < n; ++i) for (int j = 0; j < m; ++j) if (vmask[i * m + j]) vs->cam[k++] = j;
    return 0;
}
static void vis_free(vis_t *vs) { free(vs->ptr); free(vs->cam); }

void orc_project(int64_t n, int m, const char *vmask, const double *p, const double *rot0, double *hx)
{
    vis_t vs; vis_build(&vs, n, m, vmask);
    const double *pb = p + CNP * m;
    for (int64_t i = 0; i < n; ++i)
        for (int64_t k = vs.ptr[i]; k < vs.ptr[i + 1]; ++k)
            project_val(p + CNP * vs.cam[k], rot0 + 4 * vs.cam[k], pb + PNP * i, hx + 2 * k);
    vis_free(&vs);
}

static void jac_one(const double *a, const double *q0, const double *M, int nccalib, int ncdist, double *A, double *B, double *hx)
{
    dual u, v;
    project_dual(a, q0, M, &u, &v);
    for (int c = 0; c < CNP; ++c) { A[c] = u.d[c]; A[CNP + c] = v.d[c]; }
    for (int c = 0; c < PNP; ++c) { B[c] = u.d[CNP + c]; B[PNP + c] = v.d[CNP + c]; }
    for (int c = 5 - nccalib; c < 5; ++c) { A[c] = 0.0; A[CNP + c] = 0.0; }            /* sbaprojs.c:1229-1237 */
    for (int c = 5 - ncdist; c < 5; ++c) { A[5 + c] = 0.0; A[CNP + 5 + c] = 0.0; }      /* sbaprojs.c:1239-1247 */
    if (hx) { hx[0] = u.v; hx[1] = v.v; }
}

void orc_jacobian(int64_t n, int m, const char *vmask, const double *p, const double *rot0, int nccalib, int ncdist,
                  double *A /* nobs x 32 */, double *B /* nobs x 6 */)
{
    vis_t vs; vis_build(&vs, n, m, vmask);
    const double *pb = p + CNP * m;
    for (int64_t i = 0; i < n; ++i)
        for (int64_t k = vs.ptr[i]; k < vs.ptr[i + 1]; ++k)
            jac_one(p + CNP * vs.cam[k], rot0 + 4 * vs.cam[k], pb + PNP * i, nccalib, ncdist, A + 32 * k, B + 6 * k, NULL);
    vis_free(&vs);
}

/* inverse of the symmetric 3x3 matrix whose UPPER triangle is in V (row-major 3x3) -> LOWER triangle incl. diagonal
 * (what sba_symat_invert_BK leaves behind, sba_lapack.c:1118-1121).  Cofactor formula (valid for indefinite
 * matrices too); returns 0 when the matrix is singular. */
static int sym3_inv_upper_to_lower(double *V)
{
    double a = V[0], b = V[1], c = V[2], d = V[4], e = V[5], f = V[8];
    double c00 = d * f - e * e, c01 = c * e - b * f, c02 = b * e - c * d;
    double det = a * c00 + b * c01 + c * c02;
    if (det == 0.0 || !(det == det)) return 0;
    double id = 1.0 / det;
    V[0] = c00 * id;
    V[3] = c01 * id; V[4] = (a * f - c * c) * id;
    V[6] = c02 * id; V[7] = (b * c - a * e) * id; V[8] = (a * d - b * b) * id;
    return 1;
}

/* dense Cholesky solve, upper factor of a column-major (= row-major, symmetric) matrix, LAPACK dpotf2 semantics:
 * returns 0 when a pivot is not positive. S is overwritten. */
static int chol_solve(double *S, const double *E, double *x, int N)
{
    /* factor S = R^T R with R upper (stored in the upper triangle, row-major) */
    for (int j = 0; j < N; ++j) {
        double ajj = S[j * N + j];
        for (int k = 0; k < j; ++k) ajj -= S[k * N + j] * S[k * N + j];
        if (ajj <= 0.0 || ajj != ajj) return 0;
        ajj = sqrt(ajj);
        S[j * N + j] = ajj;
        for (int c = j + 1; c < N; ++c) {
            double s = S[j * N + c];
            for (int k = 0; k < j; ++k) s -= S[k * N + j] * S[k * N + c];
            S[j * N + c] = s / ajj;
        }
    }
    for (int i = 0; i < N; ++i) {            /* R^T y = E */
        double s = E[i];
        for (int k = 0; k < i; ++k) s -= S[k * N + i] * x[k];
        x[i] = s / S[i * N + i];
    }
    for (int i = N - 1; i >= 0; --i) {       /* R x = y */
        double s = x[i];
        for (int k = i + 1; k < N; ++k) s -= S[i * N + k] * x[k];
        x[i] = s / S[i * N + i];
    }
    return 1;
}

static double eval_cost(const vis_t *vs, const double *p, const double *rot0, const double *x, double *e)
{
    const double *pb = p + CNP * vs->m;
    double sum = 0.0;
    for (int64_t i = 0; i < vs->n; ++i)
        for (int64_t k = vs->ptr[i]; k < vs->ptr[i + 1]; ++k) {
            double h[2];
            project_val(p + CNP * vs->cam[k], rot0 + 4 * vs->cam[k], pb + PNP * i, h);
            e[2 * k] = x[2 * k] - h[0]; e[2 * k + 1] = x[2 * k + 1] - h[1];
            sum += e[2 * k] * e[2 * k] + e[2 * k + 1] * e[2 * k + 1];
        }
    return sum;
}

/* per-camera fixed-parameter counts (EXTENSION used by BASELINE config 5; the reference's are global): when non-NULL,
 * cam_nk[j] / cam_nd[j] replace nccalib / ncdist for camera j */
static const int *g_cam_nk = NULL, *g_cam_nd = NULL;
void orc_set_camera_fix(const int *nk, const int *nd) { g_cam_nk = nk; g_cam_nd = nd; }

int orc_motstr(int64_t n, int64_t ncon, int m, int mcon, const char *vmask, double *p, const double *rot0, const double *x,
               int nccalib, int ncdist, int itmax, int verbose, const double opts[5], double info[10], int compat)
{
    vis_t vs; vis_build(&vs, n, m, vmask);
    const int64_t nvis = vs.nobs, nobs = 2 * nvis, nvars = (int64_t)m * CNP + n * PNP;
    const int mmcon = m - mcon, N = mmcon * CNP;
    if (nobs < nvars) {
        fprintf(stderr, "oracle: cannot solve a problem with fewer measurements [%lld] than unknowns [%lld]\n", (long long)nobs, (long long)nvars);
        vis_free(&vs); return -1;
    }
    if (itmax == 0) { vis_free(&vs); return 0; }
    double *A = (double *)malloc((nvis ? nvis : 1) * 32 * sizeof(double));
    double *B = (double *)malloc((nvis ? nvis : 1) * 6 * sizeof(double));
    double *W = (double *)malloc((nvis ? nvis : 1) * 48 * sizeof(double));
    double *Y = (double *)malloc((size_t)m * 48 * sizeof(double));
    double *U = (double *)calloc((size_t)m * 256, sizeof(double));
    double *V = (double *)calloc((size_t)(n ? n : 1) * 9, sizeof(double));
    double *e = (double *)malloc((nobs ? nobs : 1) * sizeof(double));
    double *enew = (double *)malloc((nobs ? nobs : 1) * sizeof(double));
    double *eab = (double *)calloc(nvars, sizeof(double));
    double *diagUV = (double *)calloc(nvars, sizeof(double));
    double *S = (double *)malloc((size_t)N * N * sizeof(double));
    double *E = (double *)malloc((size_t)m * CNP * sizeof(double));
    double *dp = (double *)calloc(nvars, sizeof(double));
    double *pdp = (double *)malloc(nvars * sizeof(double));
    double *ea = eab, *eb = eab + m * CNP, *dpa = dp, *dpb = dp + m * CNP;
    const double *pb;
    const double tau = fabs(opts[0]), eps1 = fabs(opts[1]), eps2 = fabs(opts[2]), eps2_sq = opts[2] * opts[2],
                 eps3_sq = opts[3] * opts[3], eps4_sq = opts[4] * opts[4];
    double mu = 0.0, eab_inf = 0.0, p_eL2, pdp_eL2, p_L2 = 0.0, dp_L2 = DBL_MAX, dF, dL, init_p_eL2, tmp;
    int nu = 2, nu2, stop = 0, nfev, njev = 0, nlss = 0, itno, retval;

    p_eL2 = eval_cost(&vs, p, rot0, x, e); nfev = 1;
    if (verbose) printf("initial motstr-SBA error %g [%g]\n", p_eL2, p_eL2 / nvis);
    init_p_eL2 = p_eL2;
    if (!isfinite(p_eL2)) stop = 7;

    for (itno = 0; itno < itmax && !stop; ++itno) {
        pb = p + CNP * m;
        for (int64_t i = 0; i < n; ++i)
            for (int64_t k = vs.ptr[i]; k < vs.ptr[i + 1]; ++k)
                jac_one(p + CNP * vs.cam[k], rot0 + 4 * vs.cam[k], pb + PNP * i, g_cam_nk ? g_cam_nk[vs.cam[k]] : nccalib,
                        g_cam_nd ? g_cam_nd[vs.cam[k]] : ncdist, A + 32 * k, B + 6 * k, NULL);
        ++njev;
        memset(U, 0, (size_t)m * 256 * sizeof(double)); memset(ea, 0, (size_t)m * CNP * sizeof(double));
        memset(V, 0, (size_t)n * 9 * sizeof(double)); memset(eb, 0, (size_t)n * PNP * sizeof(double));
        for (int64_t i = 0; i < n; ++i)
            for (int64_t k = vs.ptr[i]; k < vs.ptr[i + 1]; ++k) {
                const int j = vs.cam[k];
                const double *Ak = A + 32 * k, *Bk = B + 6 * k, *ek = e + 2 * k;
                if (j >= mcon) {
                    double *Uj = U + 256 * j;
                    for (int r = 0; r < CNP; ++r) {
                        for (int c = 0; c < CNP; ++c) Uj[r * CNP + c] += Ak[r] * Ak[c] + Ak[CNP + r] * Ak[CNP + c];
                        ea[j * CNP + r] += Ak[r] * ek[0] + Ak[CNP + r] * ek[1];
                    }
                }
                if (i >= ncon) {
                    double *Vi = V + 9 * i;
                    for (int r = 0; r < PNP; ++r) {
                        for (int c = r; c < PNP; ++c) Vi[r * PNP + c] += Bk[r] * Bk[c] + Bk[PNP + r] * Bk[PNP + c];
                        eb[i * PNP + r] += Bk[r] * ek[0] + Bk[PNP + r] * ek[1];
                    }
                }
                if (j >= mcon && i >= ncon) {
                    double *Wk = W + 48 * k;
                    for (int r = 0; r < CNP; ++r)
                        for (int c = 0; c < PNP; ++c) Wk[r * PNP + c] = Ak[r] * Bk[c] + Ak[CNP + r] * Bk[PNP + c];
                }
            }
        p_L2 = eab_inf = 0.0;
        for (int64_t t = 0; t < nvars; ++t) { if (eab_inf < (tmp = fabs(eab[t]))) eab_inf = tmp; p_L2 += p[t] * p[t]; }
        for (int j = mcon; j < m; ++j) for (int r = 0; r < CNP; ++r) diagUV[j * CNP + r] = U[256 * j + r * CNP + r];
        for (int64_t i = ncon; i < n; ++i) for (int r = 0; r < PNP; ++r) diagUV[m * CNP + i * PNP + r] = V[9 * i + r * PNP + r];
        if (eab_inf <= eps1) { dp_L2 = 0.0; stop = 1; break; }
        if (itno == 0) {
            tmp = DBL_MIN;
            for (int64_t t = mcon * CNP; t < m * CNP; ++t) if (diagUV[t] > tmp) tmp = diagUV[t];
            for (int64_t t = m * CNP + ncon * PNP; t < nvars; ++t) if (diagUV[t] > tmp) tmp = diagUV[t];
            mu = tau * tmp;
        }
        while (1) {
            int singular = 0, issolved = 0;
            for (int j = mcon; j < m; ++j) for (int r = 0; r < CNP; ++r) U[256 * j + r * CNP + r] += mu;
            for (int64_t i = ncon; i < n; ++i) {
                double *Vi = V + 9 * i;
                for (int r = 0; r < PNP; ++r) Vi[r * PNP + r] += mu;
                if (!sym3_inv_upper_to_lower(Vi)) { singular = 1; break; }   /* -> moredamping, sba_levmar.c:1008-1014 */
            }
            if (!singular) {
                memset(S, 0, (size_t)N * N * sizeof(double));
                memset(E, 0, (size_t)m * CNP * sizeof(double));
                for (int64_t i = ncon; i < n; ++i) {
                    const double *Vi = V + 9 * i;   /* inverse in the lower triangle */
                    const double n00 = Vi[0], n10 = Vi[3], n11 = Vi[4], n20 = Vi[6], n21 = Vi[7], n22 = Vi[8];
                    const int64_t k0 = vs.ptr[i], k1 = vs.ptr[i + 1];
                    for (int64_t k = k0; k < k1; ++k) {          /* Y_ij = W_ij (V*_i)^-1 */
                        const double *Wk = W + 48 * k; double *Yk = Y + 48 * (k - k0);
                        if (vs.cam[k] < mcon) continue;
                        for (int r = 0; r < CNP; ++r) {
                            const double w0 = Wk[3 * r], w1 = Wk[3 * r + 1], w2 = Wk[3 * r + 2];
                            Yk[3 * r] = w0 * n00 + w1 * n10 + w2 * n20;
                            Yk[3 * r + 1] = w0 * n10 + w1 * n11 + w2 * n21;
                            Yk[3 * r + 2] = w0 * n20 + w1 * n21 + w2 * n22;
                        }
                    }
                    for (int64_t ka = k0; ka < k1; ++ka) {
                        const int j = vs.cam[ka];
                        if (j < mcon) continue;
                        const double *Yk = Y + 48 * (ka - k0);
                        for (int r = 0; r < CNP; ++r)
                            E[j * CNP + r] += Yk[3 * r] * eb[i * PNP] + Yk[3 * r + 1] * eb[i * PNP + 1] + Yk[3 * r + 2] * eb[i * PNP + 2];
                        for (int64_t kb = ka; kb < k1; ++kb) {       /* upper blocks j <= k (cameras ascending) */
                            const int kk = vs.cam[kb];
                            const double *Wk = W + 48 * kb;
                            double *Sb = S + (size_t)(j - mcon) * CNP * N + (size_t)(kk - mcon) * CNP;
                            for (int r = 0; r < CNP; ++r)
                                for (int c = 0; c < CNP; ++c)
                                    Sb[r * N + c] -= Yk[3 * r] * Wk[3 * c] + Yk[3 * r + 1] * Wk[3 * c + 1] + Yk[3 * r + 2] * Wk[3 * c + 2];
                        }
                    }
                }
                for (int j = mcon; j < m; ++j) {
                    double *Sb = S + (size_t)(j - mcon) * CNP * N + (size_t)(j - mcon) * CNP;
                    for (int r = 0; r < CNP; ++r) {
                        for (int c = 0; c < CNP; ++c) Sb[r * N + c] += U[256 * j + r * CNP + c];
                        E[j * CNP + r] = ea[j * CNP + r] - E[j * CNP + r];
                    }
                }
                for (int a_ = 0; a_ < mmcon; ++a_)            /* mirror the strictly-lower blocks */
                    for (int b_ = 0; b_ < a_; ++b_)
                        for (int r = 0; r < CNP; ++r)
                            for (int c = 0; c < CNP; ++c)
                                S[(size_t)(a_ * CNP + r) * N + b_ * CNP + c] = S[(size_t)(b_ * CNP + c) * N + a_ * CNP + r];
                issolved = chol_solve(S, E + mcon * CNP, dpa + mcon * CNP, N);
                ++nlss;
                for (int t = 0; t < mcon * CNP; ++t) dpa[t] = 0.0;
            }
            if (issolved) {
                for (int64_t i = ncon; i < n; ++i) {
                    double t0 = eb[i * PNP], t1 = eb[i * PNP + 1], t2 = eb[i * PNP + 2];
                    for (int64_t k = vs.ptr[i]; k < vs.ptr[i + 1]; ++k) {
                        const int j = vs.cam[k];
                        if (j < mcon) continue;
                        const double *Wk = W + 48 * k, *dj = dpa + j * CNP;
                        for (int r = 0; r < CNP; ++r) { t0 -= Wk[3 * r] * dj[r]; t1 -= Wk[3 * r + 1] * dj[r]; t2 -= Wk[3 * r + 2] * dj[r]; }
                    }
                    const double *Vi = V + 9 * i;
                    dpb[i * PNP] = Vi[0] * t0 + Vi[3] * t1 + Vi[6] * t2;
                    dpb[i * PNP + 1] = Vi[3] * t0 + Vi[4] * t1 + Vi[7] * t2;
                    dpb[i * PNP + 2] = Vi[6] * t0 + Vi[7] * t1 + Vi[8] * t2;
                }
                for (int64_t t = 0; t < ncon * PNP; ++t) dpb[t] = 0.0;
                dp_L2 = 0.0;
                for (int64_t t = 0; t < nvars; ++t) { pdp[t] = p[t] + (tmp = dp[t]); dp_L2 += tmp * tmp; }
                if (dp_L2 <= eps2_sq * p_L2) { stop = 2; break; }
                if (dp_L2 >= (p_L2 + eps2) / 1e-24) {
                    fprintf(stderr, "oracle: the matrix of the augmented normal equations is almost singular\n");
                    retval = -1; goto done;
                }
                pdp_eL2 = eval_cost(&vs, pdp, rot0, x, enew); ++nfev;
                if (!isfinite(pdp_eL2)) { stop = 7; break; }
                dL = 0.0;
                for (int64_t t = 0; t < nvars; ++t) dL += dp[t] * (mu * dp[t] + eab[t]);
                dF = p_eL2 - pdp_eL2;
                if (verbose > 1) printf("\ndamping term %8g, gain ratio %8g, errors %8g / %8g = %g\n", mu, dL != 0.0 ? dF / dL : dF / DBL_EPSILON,
                                        p_eL2 / nvis, pdp_eL2 / nvis, p_eL2 / pdp_eL2);
                if (dL > 0.0 && dF > 0.0) {
                    tmp = (2.0 * dF / dL - 1.0); tmp = 1.0 - tmp * tmp * tmp;
                    mu = mu * ((tmp >= 0.3333333334) ? tmp : 0.3333333334);
                    nu = 2;
                    if (pdp_eL2 - 2.0 * sqrt(p_eL2 * pdp_eL2) < (eps4_sq - 1.0) * p_eL2) stop = 4;
                    memcpy(p, pdp, nvars * sizeof(double));
                    memcpy(e, enew, nobs * sizeof(double));
                    p_eL2 = pdp_eL2;
                    break;
                }
            }
            mu *= nu;
            nu2 = (int)((unsigned)nu << 1);
            if (nu2 <= nu) { fprintf(stderr, "oracle: too many failed attempts to increase the damping factor\n"); stop = 6; break; }
            nu = nu2;
            if (!compat) {       /* textbook LM: restore the diagonals (the block the reference compiles out) */
                for (int j = mcon; j < m; ++j) for (int r = 0; r < CNP; ++r) U[256 * j + r * CNP + r] = diagUV[j * CNP + r];
                for (int64_t i = ncon; i < n; ++i) for (int r = 0; r < PNP; ++r) V[9 * i + r * PNP + r] = diagUV[m * CNP + i * PNP + r];
            }
        }
        if (p_eL2 <= eps3_sq) stop = 5;
    }
    if (itno >= itmax) stop = 3;
    if (info) {
        tmp = DBL_MIN;
        for (int64_t t = mcon * CNP; t < m * CNP; ++t) if (tmp < diagUV[t]) tmp = diagUV[t];
        for (int64_t t = m * CNP + ncon * PNP; t < nvars; ++t) if (tmp < diagUV[t]) tmp = diagUV[t];
        info[0] = init_p_eL2; info[1] = p_eL2; info[2] = eab_inf; info[3] = dp_L2; info[4] = mu / tmp;
        info[5] = itno; info[6] = stop; info[7] = nfev; info[8] = njev; info[9] = nlss;
    }
    retval = (stop != 7) ? itno : -1;
done:
    free(A); free(B); free(W); free(Y); free(U); free(V); free(e); free(enew); free(eab); free(diagUV); free(S); free(E); free(dp); free(pdp);
    vis_free(&vs);
    return retval;
}

/* ---------- motion-only / structure-only LM (sba_mot_levmar_x sba_levmar.c:1437-1980, sba_str_levmar_x :1987-2540) ----------
 * Both are the same block-diagonal LM: which = 0 keeps the points fixed and solves one 16x16 system (U_j + mu I) da_j = ea_j
 * per camera j >= mcon (:1838-1852); which = 1 keeps the cameras fixed and solves one 3x3 system (V_i + mu I) db_i = eb_i per
 * point i >= ncon (:2383-2397).  The systems go through sba_Axb_Chol on a COPY of the block (sba_lapack.c:412-426,
 * iscolmaj = 0), so - unlike the motstr driver - the blocks are not overwritten; the damping is still never taken off the
 * diagonals after a rejected step (the restore loops are #if 0, :1932-1940 / :2477-2485), which `compat` reproduces.
 * p = [16 m | 3 n] as in orc_motstr; only the free half is updated.  nlss counts every block solve (:1851, :2396). */
int orc_blockdiag(int which, int64_t n, int64_t ncon, int m, int mcon, const char *vmask, double *p, const double *rot0,
                  const double *x, int nccalib, int ncdist, int itmax, int verbose, const double opts[5], double info[10], int compat)
{
    vis_t vs; vis_build(&vs, n, m, vmask);
    const int64_t nvis = vs.nobs, nobs = 2 * nvis;
    const int bs = which == 0 ? CNP : PNP;                       /* block size */
    const int64_t nblk = which == 0 ? m : n, nfix = which == 0 ? mcon : ncon, nvars = nblk * bs;
    double *q = which == 0 ? p : p + CNP * m;                    /* the free half of p */
    if (nobs < nvars) {
        fprintf(stderr, "oracle: cannot solve a problem with fewer measurements [%lld] than unknowns [%lld]\n", (long long)nobs, (long long)nvars);
        vis_free(&vs); return -1;
    }
    if (itmax == 0) { vis_free(&vs); return 0; }
    const int64_t ntot = (int64_t)m * CNP + n * PNP;
    double *H = (double *)calloc((size_t)(nblk ? nblk : 1) * bs * bs, sizeof(double));     /* U_j or V_i, full row-major blocks */
    double *g = (double *)calloc((size_t)(nvars ? nvars : 1), sizeof(double));             /* ea or eb */
    double *dq = (double *)calloc((size_t)(nvars ? nvars : 1), sizeof(double));
    double *pdp = (double *)malloc((size_t)ntot * sizeof(double));
    double *e = (double *)malloc((size_t)(nobs ? nobs : 1) * sizeof(double));
    double *hx = (double *)malloc((size_t)(nobs ? nobs : 1) * sizeof(double));
    double *diag = (double *)calloc((size_t)(nvars ? nvars : 1), sizeof(double));
    double blk[CNP * CNP];
    const double tau = fabs(opts[0]), eps1 = fabs(opts[1]), eps2 = fabs(opts[2]), eps2_sq = opts[2] * opts[2],
                 eps3_sq = opts[3] * opts[3], eps4_sq = opts[4] * opts[4];
    double mu = 0.0, g_inf = 0.0, p_eL2, pdp_eL2, p_L2 = 0.0, dp_L2 = DBL_MAX, dF, dL, tmp, init_p_eL2;
    int nu = 2, nu2, stop = 0, nfev, njev = 0, nlss = 0, itno, retval;
    p_eL2 = eval_cost(&vs, p, rot0, x, e); nfev = 1;
    if (verbose) printf("initial %s-SBA error %g [%g]\n", which == 0 ? "mot" : "str", p_eL2, p_eL2 / (double)nvis);
    init_p_eL2 = p_eL2;
    if (!isfinite(p_eL2)) stop = 7;
    for (itno = 0; itno < itmax && !stop; ++itno) {
        memset(H, 0, (size_t)nblk * bs * bs * sizeof(double)); memset(g, 0, (size_t)nvars * sizeof(double));
        for (int64_t i = 0; i < n; ++i)
            for (int64_t k = vs.ptr[i]; k < vs.ptr[i + 1]; ++k) {
                const int j = vs.cam[k];
                const int64_t b = which == 0 ? j : i;
                if (b < nfix) continue;
                double A[32], B[6];
                const int nk = g_cam_nk ? g_cam_nk[j] : nccalib, nd = g_cam_nd ? g_cam_nd[j] : ncdist;
                jac_one(p + CNP * j, rot0 + 4 * j, p + CNP * m + PNP * i, nk, nd, A, B, NULL);
                const double *J = which == 0 ? A : B;            /* 2 x bs, row-major */
                double *Hb = H + b * bs * bs, *gb = g + b * bs;
                for (int r = 0; r < bs; ++r) {
                    for (int c = r; c < bs; ++c) Hb[r * bs + c] += J[r] * J[c] + J[bs + r] * J[bs + c];
                    gb[r] += J[r] * e[2 * k] + J[bs + r] * e[2 * k + 1];
                }
            }
        for (int64_t b = nfix; b < nblk; ++b)                    /* symmetric copy (:1757-1759) and saved diagonal */
            for (int r = 0; r < bs; ++r) {
                for (int c = 0; c < r; ++c) H[b * bs * bs + r * bs + c] = H[b * bs * bs + c * bs + r];
                diag[b * bs + r] = H[b * bs * bs + r * bs + r];
            }
        ++njev;
        p_L2 = 0.0; g_inf = 0.0;
        for (int64_t t = 0; t < nvars; ++t) { if (g_inf < (tmp = fabs(g[t]))) g_inf = tmp; p_L2 += q[t] * q[t]; }
        if (g_inf <= eps1) { dp_L2 = 0.0; stop = 1; break; }
        if (itno == 0) {
            tmp = DBL_MIN;
            for (int64_t t = nfix * bs; t < nvars; ++t) if (diag[t] > tmp) tmp = diag[t];
            mu = tau * tmp;
        }
        while (1) {
            for (int64_t b = nfix; b < nblk; ++b)
                for (int r = 0; r < bs; ++r) H[b * bs * bs + r * bs + r] += mu;
            int64_t nsolved = nfix;
            memset(dq, 0, (size_t)nfix * bs * sizeof(double));
            for (int64_t b = nfix; b < nblk; ++b) {
                memcpy(blk, H + b * bs * bs, (size_t)bs * bs * sizeof(double));
                nsolved += chol_solve(blk, g + b * bs, dq + b * bs, bs);
                ++nlss;
            }
            if (nsolved == nblk) {
                memcpy(pdp, p, (size_t)ntot * sizeof(double));
                double *qn = which == 0 ? pdp : pdp + CNP * m;
                dp_L2 = 0.0;
                for (int64_t t = 0; t < nvars; ++t) { qn[t] = q[t] + (tmp = dq[t]); dp_L2 += tmp * tmp; }
                if (dp_L2 <= eps2_sq * p_L2) { stop = 2; break; }
                if (dp_L2 >= (p_L2 + eps2) / 1e-24) {
                    fprintf(stderr, "oracle: the matrix of the augmented normal equations is almost singular\n");
                    retval = -1; goto done;
                }
                pdp_eL2 = eval_cost(&vs, pdp, rot0, x, hx); ++nfev;
                if (!isfinite(pdp_eL2)) { stop = 7; break; }
                dL = 0.0;
                for (int64_t t = 0; t < nvars; ++t) dL += dq[t] * (mu * dq[t] + g[t]);
                dF = p_eL2 - pdp_eL2;
                if (dL > 0.0 && dF > 0.0) {
                    tmp = 2.0 * dF / dL - 1.0;
                    tmp = 1.0 - tmp * tmp * tmp;
                    mu = mu * ((tmp >= 0.3333333334) ? tmp : 0.3333333334);
                    nu = 2;
                    if (pdp_eL2 - 2.0 * sqrt(p_eL2 * pdp_eL2) < (eps4_sq - 1.0) * p_eL2) stop = 4;
                    memcpy(p, pdp, (size_t)ntot * sizeof(double));
                    memcpy(e, hx, (size_t)nobs * sizeof(double));
                    p_eL2 = pdp_eL2;
                    break;
                }
            }
            mu *= nu;
            nu2 = (int)((unsigned)nu << 1);
            if (nu2 <= nu) { stop = 6; break; }
            nu = nu2;
            if (!compat)                                         /* textbook LM: take the damping off again */
                for (int64_t b = nfix; b < nblk; ++b)
                    for (int r = 0; r < bs; ++r) H[b * bs * bs + r * bs + r] = diag[b * bs + r];
        }
        if (p_eL2 <= eps3_sq) stop = 5;
    }
    if (itno >= itmax) stop = 3;
    if (info) {
        tmp = DBL_MIN;
        for (int64_t t = nfix * bs; t < nvars; ++t) if (tmp < diag[t]) tmp = diag[t];      /* restored diagonals (:1945-1962) */
        info[0] = init_p_eL2; info[1] = p_eL2; info[2] = g_inf; info[3] = dp_L2; info[4] = mu / tmp;
        info[5] = itno; info[6] = stop; info[7] = nfev; info[8] = njev; info[9] = nlss;
    }
    retval = (stop != 7) ? itno : -1;
done:
    free(H); free(g); free(dq); free(pdp); free(e); free(hx); free(diag);
    vis_free(&vs);
    return retval;
}
