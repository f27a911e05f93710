// In-solver wand-length constraint.  EXTENSION - parity unpinned: the reference applies the wand length only AFTER the bundle
// adjustment, as one global scale (argus_gui/sbaDriver.py:703-713) and reports the spread as the "wand score" (:779-781);
// nothing in sba_motstr_levmar_x knows about it.  BASELINE's north star and config 3 ask for the constraint inside the solver.
//
// Model.  A wand pair p = (a, b) of points with known length L adds one measurement  w L  predicted as  w ||X_a - X_b||:
//     r_p = w (L - d),  d = ||X_a - X_b||,  n = (X_a - X_b) / d,  Jacobian rows  g_a = w n  (w.r.t. X_a),  g_b = -w n  (X_b)
// (g = 0 for a fixed point).  It touches point parameters only, so U, ea and W are unchanged and the point block of the pair
// becomes   V*_p = diag(V*_a, V*_b) + g g^T   (6x6, g = [g_a; g_b]),   eb'_i = eb_i + g_i r.
// Sherman-Morrison with h_i = V*_i^-1 g_i, sigma = 1 / (1 + g_a.h_a + g_b.h_b):
//     V*_p^-1 = diag(V*_a^-1, V*_b^-1) - sigma h h^T
// which leaves the whole per-observation pipeline (Z, the Schur kernel, the Gram kernel) untouched and adds, per pair,
//     q_pj = sum_{i in {a,b}} W_ij h_i                 (16-vector per camera that sees a or b)
//     S_jk += sigma q_pj q_pk^T ,    E_j -= beta q_pj ,    beta = sigma (r - g_a.c_a - g_b.c_b),  c_i = V*_i^-1 eb_i
//     db_i = V*_i^-1 t'_i - sigma h_i (h_a.t'_a + h_b.t'_b),   t'_i = eb'_i - sum_j W_ij^T da_j
// and r^2 to every cost, max |eb'| / max diag(V_i + g_i g_i^T) to the statistics of a linearisation.
// Kernels: k_wand_cost (cost at the current point), k_wand_prepare (after the tile lineariser: statistics, h, sigma, beta,
// q), k_wand_schur + k_wand_apply (the rank-one corrections, fixed-order partial sums: no floating-point atomics),
// k_wand_backsub (the two points of every pair; the per-point kernel skips them).  With no pairs none of them is launched and
// the solver is bit-identical to the unconstrained one.
#pragma once
#include "ba_common.cuh"
#include "ba_kernels.cuh"

namespace argus {

constexpr int kWandGridMax = 64;     // blocks of the per-pair kernels (their partial sums are appended to the per-point kernels')
constexpr int kWandThreads = 128;
constexpr int kWandRec = 16;         // doubles per pair: h_a[3], h_b[3], sigma, beta, r, g_a[3], g_b[3], free

struct WandPlan {
    int64_t npairs;
    const int32_t *ia, *ib;          // the two points of every pair (local indices)
    const double *len;               // wand length of every pair
    double w;                        // weight (pixels per length unit)
    double *rec;                     // npairs x kWandRec
    double *Q;                       // npairs x m x 16
};

struct WandGeom { double r, d, g[3]; };
__device__ __forceinline__ WandGeom wand_geom(const double *__restrict__ pts, int a, int b, double len, double w)
{
    WandGeom o;
    const double dx = pts[3 * (int64_t)a] - pts[3 * (int64_t)b], dy = pts[3 * (int64_t)a + 1] - pts[3 * (int64_t)b + 1],
                 dz = pts[3 * (int64_t)a + 2] - pts[3 * (int64_t)b + 2];
    o.d = sqrt(dx * dx + dy * dy + dz * dz);
    const double id = (o.d > 0.0) ? w / o.d : 0.0;        // coincident points: no direction, the pair only contributes its cost
    o.g[0] = dx * id; o.g[1] = dy * id; o.g[2] = dz * id;
    o.r = w * (len - o.d);
    return o;
}

// cameras for which q_pj exists: those that see a free point of the pair
__device__ __forceinline__ uint64_t wand_cams(const BaProblem &pb, const WandPlan &wp, int64_t p)
{
    const int a = wp.ia[p], b = wp.ib[p];
    return ((a >= pb.ncon) ? pb.pt_mask[a] : 0ull) | ((b >= pb.ncon) ? pb.pt_mask[b] : 0ull);
}

// sum of r^2 at the current (trial = 0) or trial points; per-block partials -> part[blockIdx.x]
__global__ void __launch_bounds__(kWandThreads) k_wand_cost(WandPlan wp, ParamBufs pbuf, const LmState *__restrict__ st, double *__restrict__ part)
{
    __shared__ double red[32];
    if (st->done) return;
    const double *pts = pb_pts(pbuf, st->cur_pts);
    double acc = 0.0;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < wp.npairs; p += (int64_t)gridDim.x * blockDim.x) {
        const WandGeom g = wand_geom(pts, wp.ia[p], wp.ib[p], wp.len[p], wp.w);
        acc += g.r * g.r;
    }
    const double s = block_sum(acc, red);
    if (threadIdx.x == 0) part[blockIdx.x] = s;
}

__device__ __forceinline__ void sym3_mul(const double *N, const double *v, double *o)
{
    o[0] = N[0] * v[0] + N[1] * v[1] + N[2] * v[2];
    o[1] = N[1] * v[0] + N[3] * v[1] + N[4] * v[2];
    o[2] = N[2] * v[0] + N[4] * v[1] + N[5] * v[2];
}

// q_pj (+)= W_ij h_i over the observations of point i, W_ij = A_ij^T B_ij re-evaluated; `seen`: cameras already written
__device__ __forceinline__ void wand_accumulate_q(const BaProblem &pb, const CamConst *sc, const double *__restrict__ pts, int i, const double *h,
                                                  double *__restrict__ Qp, uint64_t seen)
{
    const double M0 = pts[3 * (int64_t)i], M1 = pts[3 * (int64_t)i + 1], M2 = pts[3 * (int64_t)i + 2];
    const int k1 = pb.obs_ptr[i + 1];
    for (int k = pb.obs_ptr[i]; k < k1; ++k) {
        const int j = pb.obs_cam[k];
        double *q = Qp + 16 * j;
        if (j < pb.mcon) continue;                        // fixed camera: no block in the reduced system
        double u, v, A[32], B[6];
        const int nfk = pb.cam_fix ? pb.cam_fix[2 * j] : pb.nfix_k, nfd = pb.cam_fix ? pb.cam_fix[2 * j + 1] : pb.nfix_d;
        cam_project_jac(sc[j], M0, M1, M2, nfk, nfd, u, v, A, B);
        const double y0 = B[0] * h[0] + B[1] * h[1] + B[2] * h[2], y1 = B[3] * h[0] + B[4] * h[1] + B[5] * h[2];
        const bool add = (seen >> j) & 1ull;
#pragma unroll
        for (int r = 0; r < 16; ++r) {
            const double t = A[r] * y0 + A[16 + r] * y1;
            q[r] = add ? q[r] + t : t;
        }
    }
}

// after the tile lineariser (which leaves V, eb and, mode 1, V*^-1 of every point): statistics of the pairs
// -> scpart[4 b + {0,1,2}] = {sum r^2, max |eb'|, max diag(V + g g^T)}, and (mode 1) the pair records and q vectors
__global__ void __launch_bounds__(kWandThreads) k_wand_prepare(BaProblem pb, WandPlan wp, ParamBufs pbuf, const LmState *__restrict__ st, int mode,
                                                               const double *__restrict__ V, const double *__restrict__ eb,
                                                               const double *__restrict__ Vinv, double *__restrict__ scpart)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    CamConst *sc = reinterpret_cast<CamConst *>(smem_raw);
    __shared__ double red[32];
    if (st->done) return;
    const double *cam = pb_cam(pbuf, st->cur_cam), *pts = pb_pts(pbuf, st->cur_pts);
    load_cameras(sc, cam, pb.rot0, pb.m);
    double cost = 0.0, ebmax = 0.0, vdmax = -1.79769313486231570e308;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < wp.npairs; p += (int64_t)gridDim.x * blockDim.x) {
        const int a = wp.ia[p], b = wp.ib[p];
        const WandGeom g = wand_geom(pts, a, b, wp.len[p], wp.w);
        cost += g.r * g.r;
        const bool fa = a >= pb.ncon, fb = b >= pb.ncon;
        double ga[3], gb[3];
#pragma unroll
        for (int t = 0; t < 3; ++t) { ga[t] = fa ? g.g[t] : 0.0; gb[t] = fb ? -g.g[t] : 0.0; }
        const double *Va = V + 6 * (int64_t)a, *Vb = V + 6 * (int64_t)b, *ea_ = eb + 3 * (int64_t)a, *eb_ = eb + 3 * (int64_t)b;
        if (fa) {
#pragma unroll
            for (int t = 0; t < 3; ++t) ebmax = fmax(ebmax, fabs(ea_[t] + ga[t] * g.r));
            vdmax = fmax(vdmax, fmax(Va[0] + ga[0] * ga[0], fmax(Va[3] + ga[1] * ga[1], Va[5] + ga[2] * ga[2])));
        }
        if (fb) {
#pragma unroll
            for (int t = 0; t < 3; ++t) ebmax = fmax(ebmax, fabs(eb_[t] + gb[t] * g.r));
            vdmax = fmax(vdmax, fmax(Vb[0] + gb[0] * gb[0], fmax(Vb[3] + gb[1] * gb[1], Vb[5] + gb[2] * gb[2])));
        }
        if (mode == 0) continue;
        double ha[3] = {0, 0, 0}, hb[3] = {0, 0, 0}, ca[3] = {0, 0, 0}, cb[3] = {0, 0, 0};
        if (fa) { sym3_mul(Vinv + 6 * (int64_t)a, ga, ha); sym3_mul(Vinv + 6 * (int64_t)a, ea_, ca); }
        if (fb) { sym3_mul(Vinv + 6 * (int64_t)b, gb, hb); sym3_mul(Vinv + 6 * (int64_t)b, eb_, cb); }
        const double sigma = 1.0 / (1.0 + ga[0] * ha[0] + ga[1] * ha[1] + ga[2] * ha[2] + gb[0] * hb[0] + gb[1] * hb[1] + gb[2] * hb[2]);
        const double beta = sigma * (g.r - (ga[0] * ca[0] + ga[1] * ca[1] + ga[2] * ca[2]) - (gb[0] * cb[0] + gb[1] * cb[1] + gb[2] * cb[2]));
        double *rc = wp.rec + (size_t)p * kWandRec;
        rc[0] = ha[0]; rc[1] = ha[1]; rc[2] = ha[2]; rc[3] = hb[0]; rc[4] = hb[1]; rc[5] = hb[2];
        rc[6] = sigma; rc[7] = beta; rc[8] = g.r;
        rc[9] = ga[0]; rc[10] = ga[1]; rc[11] = ga[2]; rc[12] = gb[0]; rc[13] = gb[1]; rc[14] = gb[2];
        double *Qp = wp.Q + (size_t)p * pb.m * 16;
        if (fa) wand_accumulate_q(pb, sc, pts, a, ha, Qp, 0ull);
        if (fb) wand_accumulate_q(pb, sc, pts, b, hb, Qp, fa ? pb.pt_mask[a] : 0ull);
    }
    const double s0 = block_sum(cost, red), s1 = block_max(ebmax, red), s2 = block_max(vdmax, red);
    if (threadIdx.x == 0) { scpart[4 * blockIdx.x] = s0; scpart[4 * blockIdx.x + 1] = s1; scpart[4 * blockIdx.x + 2] = s2; scpart[4 * blockIdx.x + 3] = 0.0; }
}

// rank-one corrections of the reduced system, per chunk of pairs: grid = (npairs_cam + 1, nchunk), 256 threads.
//   blockIdx.x < npairs_cam: block (j, k), thread (r, c):  part = sum_p sigma_p q_pj[r] q_pk[c]
//   blockIdx.x == npairs_cam: thread (j, r) over all cameras:  part = sum_p beta_p q_pj[r]
// part[chunk][npairs_cam * 256 + 16 m]
constexpr int kWandBatch = 32;
__global__ void __launch_bounds__(256) k_wand_schur(BaProblem pb, WandPlan wp, int mf, const LmState *__restrict__ st, double *__restrict__ part)
{
    __shared__ double sj[kWandBatch][16], sk[kWandBatch][16], ss[kWandBatch];
    __shared__ uint64_t smk[kWandBatch];
    if (st->done) return;
    const int npc = mf * (mf + 1) / 2, tid = threadIdx.x;
    const int64_t per = (wp.npairs + gridDim.y - 1) / gridDim.y;
    const int64_t p0 = per * blockIdx.y, p1 = (p0 + per < wp.npairs) ? p0 + per : wp.npairs;
    double *out = part + (size_t)blockIdx.y * ((size_t)npc * 256 + 16 * (size_t)pb.m);
    if ((int)blockIdx.x < npc) {
        int a = 0, rem = blockIdx.x;
        while (rem >= mf - a) { rem -= mf - a; ++a; }
        const int j = pb.mcon + a, k = pb.mcon + a + rem;
        const int r = tid >> 4, c = tid & 15;
        double acc = 0.0;
        for (int64_t pb0 = p0; pb0 < p1; pb0 += kWandBatch) {
            // 256 threads = 32 pairs x 8 lanes: each lane brings two entries of q_pj and two of q_pk
            const int pp = tid >> 3, l = tid & 7;
            const int64_t p = pb0 + pp;
            double2 vj = make_double2(0.0, 0.0), vk = vj; double sg = 0.0;
            if (p < p1) {
                const uint64_t mk = wand_cams(pb, wp, p);
                if (((mk >> j) & 1ull) && ((mk >> k) & 1ull)) {
                    const double *Qp = wp.Q + (size_t)p * pb.m * 16;
                    vj = *reinterpret_cast<const double2 *>(Qp + 16 * j + 2 * l);
                    vk = *reinterpret_cast<const double2 *>(Qp + 16 * k + 2 * l);
                    sg = wp.rec[(size_t)p * kWandRec + 6];
                }
            }
            sj[pp][2 * l] = vj.x * sg; sj[pp][2 * l + 1] = vj.y * sg; sk[pp][2 * l] = vk.x; sk[pp][2 * l + 1] = vk.y;
            __syncthreads();
#pragma unroll
            for (int q = 0; q < kWandBatch; ++q) acc += sj[q][r] * sk[q][c];
            __syncthreads();
        }
        out[(size_t)blockIdx.x * 256 + tid] = acc;
    } else {
        for (int t0 = 0; t0 < 16 * pb.m; t0 += 256) {
            const int t = t0 + tid, j = t >> 4;
            double acc = 0.0;
            for (int64_t pb0 = p0; pb0 < p1; pb0 += kWandBatch) {
                if (tid < kWandBatch) {
                    const bool in = pb0 + tid < p1;
                    ss[tid] = in ? wp.rec[(size_t)(pb0 + tid) * kWandRec + 7] : 0.0;
                    smk[tid] = in ? wand_cams(pb, wp, pb0 + tid) : 0ull;
                }
                __syncthreads();
                if (j < pb.m && j >= pb.mcon) {
#pragma unroll 4
                    for (int q = 0; q < kWandBatch; ++q)
                        if ((smk[q] >> j) & 1ull) acc += ss[q] * wp.Q[(size_t)(pb0 + q) * pb.m * 16 + t];
                }
                __syncthreads();
            }
            if (j < pb.m) out[(size_t)npc * 256 + t] = acc;
        }
    }
}

// red2 (Schur sums in the fragment order of k_schur_mma, then E per camera) -= the chunk partials, in fixed order
__global__ void k_wand_apply(int npc, int m, int nchunk, const double *__restrict__ part, double *__restrict__ red2, const int *__restrict__ done)
{
    if (*done) return;
    const int64_t len = (int64_t)npc * 256 + 16 * (int64_t)m;
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= len) return;
    double s = 0.0;
    for (int c = 0; c < nchunk; ++c) s += part[(size_t)c * len + t];
    if (t < (int64_t)npc * 256) {
        const int blk = (int)(t >> 8), r = (int)(t & 255) >> 4, cc = (int)(t & 15);
        const int ln = ((r & 7) << 2) | ((cc & 7) >> 1), q = ((r >> 3) * 2 + (cc >> 3)) * 2 + (cc & 1);
        red2[(int64_t)blk * 256 + ln * 8 + q] -= s;
    } else {
        red2[t] -= s;                 // E_j[r] sits right behind the Schur sums, one 16-vector per camera
    }
}

// t_i = eb_i - sum_j B_ij^T (A_ij da_j) of one point (the loop of k_backsub_recompute)
__device__ __forceinline__ void backsub_rhs(const BaProblem &pb, const CamConst *sc, const double *das, int64_t i, double M0, double M1, double M2,
                                            const double *__restrict__ eb, double *t)
{
    t[0] = eb[3 * i]; t[1] = eb[3 * i + 1]; t[2] = eb[3 * i + 2];
    const int k1 = pb.obs_ptr[i + 1];
    for (int k = pb.obs_ptr[i]; k < k1; ++k) {
        const int j = pb.obs_cam[k];
        if (j < pb.mcon) continue;
        double B[6], s0, s1;
        const int nfk = pb.cam_fix ? pb.cam_fix[2 * j] : pb.nfix_k, nfd = pb.cam_fix ? pb.cam_fix[2 * j + 1] : pb.nfix_d;
        cam_step_dir(sc[j], M0, M1, M2, nfk, nfd, das + 17 * j, s0, s1, B);
        t[0] -= B[0] * s0 + B[3] * s1; t[1] -= B[1] * s0 + B[4] * s1; t[2] -= B[2] * s0 + B[5] * s1;
    }
}
__device__ __forceinline__ double trial_cost(const BaProblem &pb, const CamConst *sn, int64_t i, double N0, double N1, double N2)
{
    double cost = 0.0;
    const int k1 = pb.obs_ptr[i + 1];
    for (int k = pb.obs_ptr[i]; k < k1; ++k) {
        double u, v;
        cam_project(sn[pb.obs_cam[k]], N0, N1, N2, u, v);
        const double2 z = pb.obs_uv[k];
        const double e0 = z.x - u, e1 = z.y - v;
        cost += e0 * e0 + e1 * e1;
    }
    return cost;
}

// back-substitution + trial evaluation of the two points of every pair (k_backsub_recompute skips them):
// part[3 b + {0,1,2}] = {trial cost incl. r_new^2, ||db||^2, dL}
__global__ void __launch_bounds__(kWandThreads) k_wand_backsub(BaProblem pb, WandPlan wp, ParamBufs pbuf, const LmState *__restrict__ st,
                                                               const double *__restrict__ da, const double *__restrict__ Vinv,
                                                               const double *__restrict__ eb, double *__restrict__ dbout, double *__restrict__ part)
{
    if (st->done) return;
    const double mu = st->mu;
    const double *__restrict__ cam = pb_cam(pbuf, st->cur_cam);
    const double *__restrict__ cam_new = pb_cam(pbuf, st->cur_cam ^ 1);
    const double *__restrict__ pts = pb_pts(pbuf, st->cur_pts);
    double *__restrict__ pts_new = pb_pts(pbuf, st->cur_pts ^ 1);
    extern __shared__ __align__(16) unsigned char smem_raw[];
    CamConst *sc = reinterpret_cast<CamConst *>(smem_raw);
    CamConst *sn = sc + pb.m;
    double *das = reinterpret_cast<double *>(sn + pb.m);
    __shared__ double red[32];
    for (int t = threadIdx.x; t < 16 * pb.m; t += blockDim.x) das[(t >> 4) * 17 + (t & 15)] = da[t];
    for (int j = threadIdx.x; j < pb.m; j += blockDim.x) { cam_precompute(cam + 16 * j, pb.rot0 + 4 * j, sc[j]); cam_precompute(cam_new + 16 * j, pb.rot0 + 4 * j, sn[j]); }
    __syncthreads();
    double cost = 0.0, dbsq = 0.0, dl = 0.0;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < wp.npairs; p += (int64_t)gridDim.x * blockDim.x) {
        const double *rc = wp.rec + (size_t)p * kWandRec;
        const double sigma = rc[6], r = rc[8];
        const int idx[2] = {wp.ia[p], wp.ib[p]};
        double M[2][3], tt[2][3], d[2][3], ht = 0.0;
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            const int64_t i = idx[s];
            M[s][0] = pts[3 * i]; M[s][1] = pts[3 * i + 1]; M[s][2] = pts[3 * i + 2];
            tt[s][0] = tt[s][1] = tt[s][2] = 0.0; d[s][0] = d[s][1] = d[s][2] = 0.0;
            if (i >= pb.ncon) {
                backsub_rhs(pb, sc, das, i, M[s][0], M[s][1], M[s][2], eb, tt[s]);
#pragma unroll
                for (int t = 0; t < 3; ++t) { tt[s][t] += rc[9 + 3 * s + t] * r; ht += rc[3 * s + t] * tt[s][t]; }
            }
        }
        double Nw[2][3];
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            const int64_t i = idx[s];
            if (i >= pb.ncon) {
                sym3_mul(Vinv + 6 * i, tt[s], d[s]);
#pragma unroll
                for (int t = 0; t < 3; ++t) {
                    d[s][t] -= sigma * rc[3 * s + t] * ht;
                    dbsq += d[s][t] * d[s][t];
                    dl += d[s][t] * (mu * d[s][t] + eb[3 * i + t] + rc[9 + 3 * s + t] * r);
                }
            }
#pragma unroll
            for (int t = 0; t < 3; ++t) {
                Nw[s][t] = M[s][t] + d[s][t];
                pts_new[3 * i + t] = Nw[s][t];
                if (dbout) dbout[3 * i + t] = d[s][t];
            }
            cost += trial_cost(pb, sn, i, Nw[s][0], Nw[s][1], Nw[s][2]);
        }
        const double dx = Nw[0][0] - Nw[1][0], dy = Nw[0][1] - Nw[1][1], dz = Nw[0][2] - Nw[1][2];
        const double rn = wp.w * (wp.len[p] - sqrt(dx * dx + dy * dy + dz * dz));
        cost += rn * rn;
    }
    const double s0 = block_sum(cost, red), s1 = block_sum(dbsq, red), s2 = block_sum(dl, red);
    if (threadIdx.x == 0) { part[3 * blockIdx.x] = s0; part[3 * blockIdx.x + 1] = s1; part[3 * blockIdx.x + 2] = s2; }
}

}  // namespace argus
