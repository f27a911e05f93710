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

constexpr int kWandGridMax = 1024;    // blocks of the per-pair kernels (their partial sums are appended to the per-point kernels')
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

// q_pj = sum_{i in {a,b} free, seen by j} A_ij^T (B_ij h_i): one thread per (pair, free camera), Jacobians re-evaluated
// (every pair costs up to 2 m of them: spread over pairs x cameras they run in parallel)
__global__ void __launch_bounds__(128) k_wand_q(BaProblem pb, WandPlan wp, ParamBufs pbuf, const LmState *__restrict__ st)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    CamConst *sc = reinterpret_cast<CamConst *>(smem_raw);
    if (st->done) return;
    const double *cam = pb_cam(pbuf, st->cur_cam), *pts = pb_pts(pbuf, st->cur_pts);
    load_cameras(sc, cam, pb.rot0, pb.m);
    const int mf = pb.m - pb.mcon;
    const int64_t total = wp.npairs * mf;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t p = t / mf;
        const int j = pb.mcon + (int)(t % mf);
        const int idx[2] = {wp.ia[p], wp.ib[p]};
        const double *rc = wp.rec + (size_t)p * kWandRec;
        double q[16];
#pragma unroll
        for (int r = 0; r < 16; ++r) q[r] = 0.0;
        bool any = false;
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            const int i = idx[s];
            const uint64_t mk = pb.pt_mask[i];
            if (i < pb.ncon || !((mk >> j) & 1ull)) continue;
            any = true;
            const double M0 = pts[3 * (int64_t)i], M1 = pts[3 * (int64_t)i + 1], M2 = pts[3 * (int64_t)i + 2];
            double u, v, A[32], B[6];
            const int nfk = pb.cam_fix ? pb.cam_fix[2 * j] : pb.nfix_k, nfd = pb.cam_fix ? pb.cam_fix[2 * j + 1] : pb.nfix_d;
            cam_project_jac(sc[j], M0, M1, M2, nfk, nfd, u, v, A, B);
            const double h0 = rc[3 * s], h1 = rc[3 * s + 1], h2 = rc[3 * s + 2];
            const double y0 = B[0] * h0 + B[1] * h1 + B[2] * h2, y1 = B[3] * h0 + B[4] * h1 + B[5] * h2;
#pragma unroll
            for (int r = 0; r < 16; ++r) q[r] += A[r] * y0 + A[16 + r] * y1;
        }
        if (any) {
            double *Qp = wp.Q + ((size_t)p * pb.m + j) * 16;
#pragma unroll
            for (int r = 0; r < 16; r += 2) *reinterpret_cast<double2 *>(Qp + r) = make_double2(q[r], q[r + 1]);
        }
    }
}

// after the tile lineariser (which leaves V, eb and, mode 1, V*^-1 of every point): statistics of the pairs
// -> scpart[4 b + {0,1,2}] = {sum r^2, max |eb'|, max diag(V + g g^T)}, and (mode 1) the pair records (k_wand_q follows)
__global__ void __launch_bounds__(kWandThreads) k_wand_prepare(BaProblem pb, WandPlan wp, ParamBufs pbuf, const LmState *__restrict__ st, int mode,
                                                               const double *__restrict__ V, const double *__restrict__ eb,
                                                               const double *__restrict__ Vinv, double *__restrict__ scpart)
{
    __shared__ double red[32];
    if (st->done) return;
    const double *pts = pb_pts(pbuf, st->cur_pts);
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
    }
    const double s0 = block_sum(cost, red), s1 = block_max(ebmax, red), s2 = block_max(vdmax, red);
    if (threadIdx.x == 0) { scpart[4 * blockIdx.x] = s0; scpart[4 * blockIdx.x + 1] = s1; scpart[4 * blockIdx.x + 2] = s2; scpart[4 * blockIdx.x + 3] = 0.0; }
}

// rank-one corrections of the reduced system.  grid = (nsub, nchunk), 256 threads.  A CTA takes a chunk of pairs and a subset of
// kWandBlocks (16) camera-pair blocks: per batch of pairs the q vectors of ALL cameras (zeros where a camera sees neither point)
// are staged in shared memory, plain and scaled by sigma; thread (block of the subset, 4 x 4 micro-tile) accumulates its tile of
// sum_p sigma_p q_pj q_pk^T  in 16 registers (4 vector loads per 16 FMAs; few registers: several CTAs per SM hide the latencies);
// subset 0 also forms  sum_p beta_p q_pj  for E.   part[chunk][npairs_cam * 256 + 16 m]
constexpr int kWandBlocks = 16;
__host__ __device__ __forceinline__ int wand_batch(int m) { return m <= 8 ? 32 : m <= 16 ? 16 : m <= 32 ? 8 : 4; }
__host__ __device__ __forceinline__ size_t wand_schur_smem(int m) { return (size_t)wand_batch(m) * m * 16 * 2 * sizeof(double) + 2 * kWandBlocks * sizeof(int); }
__global__ void __launch_bounds__(256) k_wand_schur(BaProblem pb, WandPlan wp, int mf, const LmState *__restrict__ st, double *__restrict__ part)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    if (st->done) return;
    const int m = pb.m, W = 16 * m, nb = wand_batch(m), tid = threadIdx.x;
    double *sQ = reinterpret_cast<double *>(smem_raw);            // [nb][16 m]
    double *sS = sQ + (size_t)nb * W;                              // the same scaled by sigma
    int *sj = reinterpret_cast<int *>(sS + (size_t)nb * W), *sk = sj + kWandBlocks;
    __shared__ double s_sig[32], s_beta[32];
    __shared__ unsigned long long s_mask[32];
    const int npc = mf * (mf + 1) / 2, sub = blockIdx.x;
    const int b0 = sub * kWandBlocks, nblk = min(kWandBlocks, npc - b0);
    if (tid < nblk) {
        int a = 0, rem = b0 + tid;
        while (rem >= mf - a) { rem -= mf - a; ++a; }
        sj[tid] = (pb.mcon + a) * 16; sk[tid] = (pb.mcon + a + rem) * 16;
    }
    const int64_t per = (wp.npairs + gridDim.y - 1) / gridDim.y;
    const int64_t p0 = per * blockIdx.y, p1 = (p0 + per < wp.npairs) ? p0 + per : wp.npairs;
    const int slot = tid >> 4, tr = ((tid & 15) >> 2) * 4, tc = (tid & 3) * 4;
    double acc[1][16], ae[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
    for (int q = 0; q < 16; ++q) acc[0][q] = 0.0;
    for (int64_t pb0 = p0; pb0 < p1; pb0 += nb) {
        __syncthreads();
        if (tid < nb) {
            const bool in = pb0 + tid < p1;
            s_sig[tid] = in ? wp.rec[(size_t)(pb0 + tid) * kWandRec + 6] : 0.0;
            s_beta[tid] = in ? wp.rec[(size_t)(pb0 + tid) * kWandRec + 7] : 0.0;
            s_mask[tid] = in ? wand_cams(pb, wp, pb0 + tid) : 0ull;
        }
        __syncthreads();
        for (int t = tid; t < nb * W / 2; t += 256) {              // double2 per thread and pass
            const int pp = (2 * t) / W, e = (2 * t) % W, j = e >> 4;
            double2 v = make_double2(0.0, 0.0);
            if (j >= pb.mcon && ((s_mask[pp] >> j) & 1ull)) v = *reinterpret_cast<const double2 *>(wp.Q + (size_t)(pb0 + pp) * W + e);
            *reinterpret_cast<double2 *>(sQ + (size_t)pp * W + e) = v;
            const double sg = s_sig[pp];
            *reinterpret_cast<double2 *>(sS + (size_t)pp * W + e) = make_double2(v.x * sg, v.y * sg);
        }
        __syncthreads();
#pragma unroll
        for (int bb = 0; bb < 1; ++bb) {
            const int b = slot + 16 * bb;
            if (b < nblk) {
                const double *qa = sS + sj[b] + tr, *qb = sQ + sk[b] + tc;
                for (int pp = 0; pp < nb; ++pp) {
                    const double2 a0 = *reinterpret_cast<const double2 *>(qa + (size_t)pp * W), a1 = *reinterpret_cast<const double2 *>(qa + (size_t)pp * W + 2);
                    const double2 c0 = *reinterpret_cast<const double2 *>(qb + (size_t)pp * W), c1 = *reinterpret_cast<const double2 *>(qb + (size_t)pp * W + 2);
                    const double av[4] = {a0.x, a0.y, a1.x, a1.y}, cv[4] = {c0.x, c0.y, c1.x, c1.y};
#pragma unroll
                    for (int x = 0; x < 4; ++x)
#pragma unroll
                        for (int y = 0; y < 4; ++y) acc[bb][4 * x + y] += av[x] * cv[y];
                }
            }
        }
        if (sub == 0) {
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                const int e = g * 256 + tid;
                if (e < W) { double s = ae[g]; for (int pp = 0; pp < nb; ++pp) s += s_beta[pp] * sQ[(size_t)pp * W + e]; ae[g] = s; }
            }
        }
    }
    double *out = part + (size_t)blockIdx.y * ((size_t)npc * 256 + 16 * (size_t)m);
#pragma unroll
    for (int bb = 0; bb < 1; ++bb) {
        const int b = slot + 16 * bb;
        if (b < nblk) {
#pragma unroll
            for (int x = 0; x < 4; ++x)
#pragma unroll
                for (int y = 0; y < 4; ++y) out[(size_t)(b0 + b) * 256 + (tr + x) * 16 + tc + y] = acc[bb][4 * x + y];
        }
    }
    if (sub == 0) {
#pragma unroll
        for (int g = 0; g < 4; ++g) { const int e = g * 256 + tid; if (e < W) out[(size_t)npc * 256 + e] = ae[g]; }
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
