// BA kernels shared by the three LM drivers: residual evaluation, fixed-order reductions, the scalars of one
// linearisation, the single-CTA Cholesky of small reduced systems and the camera update.  fp64 throughout.
// Reference arithmetic replaced: sba-1.6p/sba_levmar.c:781-1300 (see each kernel).
#pragma once
#include <stdio.h>
#include "ba_common.cuh"

namespace argus {

// ------------------------------------------------------------------------------------------------
// shared-memory camera table: every block precomputes R, Omega, intrinsics of all m cameras once.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void load_cameras(CamConst *sc, const double *cam, const double *rot0, int m)
{
    for (int j = threadIdx.x; j < m; j += blockDim.x) cam_precompute(cam + 16 * j, rot0 + 4 * j, sc[j]);
    __syncthreads();
}

// K1: residual cost ||x - hx(p)||^2 (img_projsKDRTS_x + nrmL2xmy: sbaprojs.c:1135-1184, sba_levmar.c:151-205),
// optionally writing hx.  One thread per point; per-block partial sums -> part[blockIdx.x].
// trial_cam: evaluate with the trial cameras (the other buffer of the pair) - the motion-only driver's func(p + dp).
__global__ void __launch_bounds__(256) k_residual(BaProblem pb, ParamBufs pbuf, const LmState *__restrict__ st, int trial_cam,
                                                  double *__restrict__ hx, double *__restrict__ part)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    CamConst *sc = reinterpret_cast<CamConst *>(smem_raw);
    __shared__ double red[32];
    if (st->done) return;
    const double *cam = pb_cam(pbuf, st->cur_cam ^ trial_cam), *pts = pb_pts(pbuf, st->cur_pts);
    load_cameras(sc, cam, pb.rot0, pb.m);
    double acc = 0.0;
    for (int64_t base = (int64_t)blockIdx.x * blockDim.x; base < pb.n; base += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = base + threadIdx.x;
        if (i < pb.n) {
            const double M0 = pts[3 * i], M1 = pts[3 * i + 1], M2 = pts[3 * i + 2];
            const int k1 = pb.obs_ptr[i + 1];
            for (int k = pb.obs_ptr[i]; k < k1; ++k) {
                double u, v;
                cam_project(sc[pb.obs_cam[k]], M0, M1, M2, u, v);
                if (hx) { hx[2 * (int64_t)k] = u; hx[2 * (int64_t)k + 1] = v; }
                const double2 z = pb.obs_uv[k];
                const double e0 = z.x - u, e1 = z.y - v;
                acc += e0 * e0 + e1 * e1;
            }
        }
    }
    const double s = block_sum(acc, red);
    if (threadIdx.x == 0) part[blockIdx.x] = s;
}

// out[i] = sum_c part[c * stride + i] in fixed order c = 0..nparts-1 (deterministic partial-sum reduction)
__global__ void k_reduce_sum(const double *__restrict__ part, int nparts, int64_t stride, int64_t len, double *__restrict__ out,
                             const int *__restrict__ done)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= len || *done) return;
    double s = 0.0;
    for (int c = 0; c < nparts; ++c) s += part[c * stride + i];
    out[i] = s;
}

// After the (optional) all-reduce: host-visible scalars of one linearisation (sba_levmar.c:941-987).
//   Ured: m x 152 (reduced U / ea), sum2 = {cost, sum pts^2}, max2 = {max|eb|, max diag V}
//   scal[0]=cost, [1]=||J^T e||_inf, [2]=||p||^2, [3]=max diag(U,V) over free cameras/points
__global__ void k_lin_finalize(const double *__restrict__ Ured, ParamBufs pbuf, const LmState *__restrict__ st, int m, int mcon,
                               const double *__restrict__ sum2, const double *__restrict__ max2, double *__restrict__ scal)
{
    if (blockIdx.x != 0 || threadIdx.x >= 32 || st->done) return;
    const double *cam = pb_cam(pbuf, st->cur_cam);
    const int lane = threadIdx.x;
    double inf = 0.0, dmax = 2.2250738585072014e-308 /* DBL_MIN, sba_levmar.c:982 */, psq = 0.0;
    for (int t = mcon * 16 + lane; t < m * 16; t += 32) {           // one (camera, parameter) per lane and pass
        const int j = t >> 4, r = t & 15;
        const double *Uj = Ured + j * kU;
        inf = fmax(inf, fabs(Uj[136 + r]));
        dmax = fmax(dmax, Uj[tri16(r, r)]);
    }
    for (int t = lane; t < 16 * m; t += 32) psq += cam[t] * cam[t];
    for (int o = 16; o > 0; o >>= 1) {                               // fixed-order butterflies
        inf = fmax(inf, __shfl_xor_sync(0xffffffffu, inf, o));
        dmax = fmax(dmax, __shfl_xor_sync(0xffffffffu, dmax, o));
        psq += __shfl_xor_sync(0xffffffffu, psq, o);
    }
    if (lane == 0) {
        scal[0] = sum2[0]; scal[1] = fmax(inf, max2[0]); scal[2] = psq + sum2[1]; scal[3] = fmax(dmax, max2[1]);
    }
}

// K5: dense Cholesky solve S da = E (sba_Axb_Chol: dpotrf + dpotrs, sba_lapack.c:373-479) in ONE thread block, for reduced
// systems below the size where the cooperative multi-SM kernel (ba_chol_coop.cuh) pays off.  status[0] = 1 on success, 0 when a
// pivot is not positive (the reference then increases the damping, sba_levmar.c:1201-1210).  S is overwritten.
constexpr int kCholNB = 32;
constexpr int kCholThreads = 512;

// K5, second generation: same contract as k_chol_solve (one CTA, S overwritten by its lower Cholesky factor, status[0] = 1 / 0),
// restructured so that no long dependent chain touches global memory:
//   * the 32x32 diagonal block is factored in shared memory by the whole CTA;
//   * its inverse (lower triangular) is formed by all warps (one column per warp pass, shuffle broadcast) and kept in
//     work[kb/32][32][32]: the panel solve L = A D^-T becomes a small matrix product with full instruction-level
//     parallelism, and both substitutions become per-block triangular matrix-vector products + parallel updates.
// work: (ceil(N/32)) x 1024 doubles of scratch.
__global__ void __launch_bounds__(kCholThreads) k_chol_solve2(double *__restrict__ S, const double *__restrict__ E, double *__restrict__ x, int N,
                                                              double *__restrict__ work, int *__restrict__ status, const int *__restrict__ done)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *P = reinterpret_cast<double *>(smem_raw);          // rows below the panel x 33 (also the running rhs later)
    __shared__ double D[kCholNB][kCholNB + 1];
    __shared__ double Di[kCholNB][kCholNB + 1];
    __shared__ double ys[kCholNB];
    __shared__ int fail;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = blockDim.x >> 5;
    if (*done) return;
    if (tid == 0) fail = 0;
    __syncthreads();
    for (int kb = 0; kb < N; kb += kCholNB) {
        const int nb = min(kCholNB, N - kb);
        for (int t = tid; t < kCholNB * kCholNB; t += blockDim.x) {
            const int r = t >> 5, c = t & 31;
            D[r][c] = (r < nb && c <= r) ? S[(int64_t)(kb + r) * N + kb + c] : ((r == c) ? 1.0 : 0.0);   // identity padding
            Di[r][c] = 0.0;
        }
        __syncthreads();
        // (1) factor the diagonal block with the whole CTA.  The step loop runs the square-root-free recurrence
        //     (D[r][c] -= D[r][j] D[c][j] / d_j, one barrier and one division on the critical path per step; every thread
        //     reads the pivot itself, so the positive-definiteness test is uniform); the columns are scaled by
        //     1/sqrt(d_j) afterwards, which gives exactly the Cholesky factor.
        bool bad = false;
        for (int j = 0; j < nb; ++j) {
            const double d = D[j][j];
            if (!(d > 0.0)) { bad = true; break; }
            const double id = 1.0 / d;
            const int w = nb - 1 - j;
            for (int t = tid; t < w * w; t += blockDim.x) {
                const int r = j + 1 + t / w, c = j + 1 + t % w;
                if (c <= r) D[r][c] -= D[r][j] * D[c][j] * id;
            }
            __syncthreads();
        }
        if (bad) { if (tid == 0) fail = 1; }
        __syncthreads();
        if (fail) break;
        if (tid < kCholNB) ys[tid] = D[tid][tid];            // pivots d_j
        __syncthreads();
        for (int t = tid; t < kCholNB * kCholNB; t += blockDim.x) {
            const int r = t >> 5, c = t & 31;
            if (c < r) D[r][c] *= rsqrt(ys[c]); else if (c == r) D[r][c] = sqrt(ys[c]);
        }
        __syncthreads();
        // (2) inverse of the factor: column jj by warp (jj mod nwarp); lane = row
        for (int jj = warp; jj < kCholNB; jj += nwarp) {
            double xi = 0.0, acc = 0.0;
            const double rdii = 1.0 / D[lane][lane];        // one division per lane, none on the 31-step dependent chain
            if (lane == jj) xi = rdii;
            for (int k = jj; k < kCholNB - 1; ++k) {
                const double xk = __shfl_sync(0xffffffffu, xi, k);
                if (lane > k) acc += D[lane][k] * xk;
                if (lane == k + 1) xi = -acc * rdii;
            }
            Di[lane][jj] = (lane >= jj) ? xi : 0.0;
        }
        __syncthreads();
        for (int t = tid; t < kCholNB * kCholNB; t += blockDim.x) {
            const int r = t >> 5, c = t & 31;
            work[(int64_t)(kb / kCholNB) * 1024 + t] = Di[r][c];
            if (r < nb && c <= r) S[(int64_t)(kb + r) * N + kb + c] = D[r][c];
        }
        // (3) panel: L_i = A_i D^-T, i.e. L[i][c] = sum_{q <= c} A[i][q] Di[c][q]
        const int r0 = kb + nb, R = N - r0;
        for (int i = tid; i < R; i += blockDim.x) {
            double *row = S + (int64_t)(r0 + i) * N + kb;
            double a[kCholNB];
#pragma unroll
            for (int q = 0; q < kCholNB; ++q) a[q] = (q < nb) ? row[q] : 0.0;
            double *pr = P + i * 33;
#pragma unroll 4
            for (int c = 0; c < kCholNB; ++c) {
                double sacc = 0.0;
#pragma unroll
                for (int q = 0; q < kCholNB; ++q) sacc += a[q] * Di[c][q];       // Di is lower triangular: zeros for q > c
                pr[c] = sacc;
                if (c < nb) row[c] = sacc;
            }
        }
        __syncthreads();
        // (4) trailing update of the lower triangle: 4x4 register tiles
        const int T = (R + 3) >> 2;
        for (int t = tid; t < T * T; t += blockDim.x) {
            const int ti = t / T, tj = t % T;
            if (tj > ti) continue;
            double acc[4][4];
#pragma unroll
            for (int a_ = 0; a_ < 4; ++a_)
#pragma unroll
                for (int b_ = 0; b_ < 4; ++b_) acc[a_][b_] = 0.0;
            const int i0 = 4 * ti, j0 = 4 * tj;
            for (int c = 0; c < kCholNB; ++c) {
                double pi[4], pj[4];
#pragma unroll
                for (int a_ = 0; a_ < 4; ++a_) { pi[a_] = (i0 + a_ < R) ? P[(i0 + a_) * 33 + c] : 0.0; pj[a_] = (j0 + a_ < R) ? P[(j0 + a_) * 33 + c] : 0.0; }
#pragma unroll
                for (int a_ = 0; a_ < 4; ++a_)
#pragma unroll
                    for (int b_ = 0; b_ < 4; ++b_) acc[a_][b_] += pi[a_] * pj[b_];
            }
#pragma unroll
            for (int a_ = 0; a_ < 4; ++a_)
#pragma unroll
                for (int b_ = 0; b_ < 4; ++b_) {
                    const int i = i0 + a_, j = j0 + b_;
                    if (i < R && j <= i) S[(int64_t)(r0 + i) * N + r0 + j] -= acc[a_][b_];
                }
        }
        __syncthreads();
    }
    if (fail) { if (tid == 0) status[0] = 0; return; }
    // forward substitution L y = E: per block y_b = Di_b rhs_b, then rhs_i -= L[i][b] y_b
    double *rhs = P;
    for (int i = tid; i < N; i += blockDim.x) rhs[i] = E[i];
    __syncthreads();
    for (int kb = 0; kb < N; kb += kCholNB) {
        const int nb = min(kCholNB, N - kb);
        for (int t = tid; t < kCholNB * kCholNB; t += blockDim.x) Di[t >> 5][t & 31] = work[(int64_t)(kb / kCholNB) * 1024 + t];
        __syncthreads();
        if (tid < kCholNB) {            // Di is zero above the diagonal and identity-padded: the full 32-term dot product is exact
            double s0 = 0.0, s1 = 0.0;
#pragma unroll
            for (int q = 0; q < kCholNB; q += 2) { s0 += Di[tid][q] * ((q < nb) ? rhs[kb + q] : 0.0); s1 += Di[tid][q + 1] * ((q + 1 < nb) ? rhs[kb + q + 1] : 0.0); }
            ys[tid] = (tid < nb) ? s0 + s1 : 0.0;
        }
        __syncthreads();
        if (tid < nb) rhs[kb + tid] = ys[tid];
        for (int i = kb + nb + tid; i < N; i += blockDim.x) {
            const double *row = S + (int64_t)i * N + kb;
            double rv[kCholNB];
#pragma unroll
            for (int c = 0; c < kCholNB; ++c) rv[c] = (c < nb) ? row[c] : 0.0;        // all loads in flight together
            double s0 = rhs[i], s1 = 0.0;
#pragma unroll
            for (int c = 0; c < kCholNB; c += 2) { s0 -= rv[c] * ys[c]; s1 -= rv[c + 1] * ys[c + 1]; }
            rhs[i] = s0 + s1;
        }
        __syncthreads();
    }
    // backward substitution L^T x = y: x_b = Di_b^T rhs_b, then rhs_i -= L[b][i]^T x_b for i above the block
    for (int kb = ((N - 1) / kCholNB) * kCholNB; kb >= 0; kb -= kCholNB) {
        const int nb = min(kCholNB, N - kb);
        for (int t = tid; t < kCholNB * kCholNB; t += blockDim.x) Di[t >> 5][t & 31] = work[(int64_t)(kb / kCholNB) * 1024 + t];
        __syncthreads();
        if (tid < kCholNB) {
            double s0 = 0.0, s1 = 0.0;
#pragma unroll
            for (int q = 0; q < kCholNB; q += 2) { s0 += Di[q][tid] * ((q < nb) ? rhs[kb + q] : 0.0); s1 += Di[q + 1][tid] * ((q + 1 < nb) ? rhs[kb + q + 1] : 0.0); }
            ys[tid] = (tid < nb) ? s0 + s1 : 0.0;
        }
        __syncthreads();
        if (tid < nb) rhs[kb + tid] = ys[tid];
        for (int i = tid; i < kb; i += blockDim.x) {
            double rv[kCholNB];
#pragma unroll
            for (int c = 0; c < kCholNB; ++c) rv[c] = (c < nb) ? S[(int64_t)(kb + c) * N + i] : 0.0;
            double s0 = rhs[i], s1 = 0.0;
#pragma unroll
            for (int c = 0; c < kCholNB; c += 2) { s0 -= rv[c] * ys[c]; s1 -= rv[c + 1] * ys[c + 1]; }
            rhs[i] = s0 + s1;
        }
        __syncthreads();
    }
    for (int i = tid; i < N; i += blockDim.x) x[i] = rhs[i];
    if (tid == 0) status[0] = 1;
}

// camera update: cam_new = cam + da (da = 0 for the mcon fixed cameras, sba_levmar.c:1211), and the camera part of
// ||dp||^2 and dL = sum dp (mu dp + J^T e) (sba_levmar.c:1261-1264, 1297-1298).  out = {da^2, dL_cam}.  One block.
__global__ void k_cam_update(int m, int mcon, ParamBufs pbuf, const LmState *__restrict__ st, const double *__restrict__ da_free,
                             const double *__restrict__ Ured, double *__restrict__ da_full,
                             double *__restrict__ out2)
{
    __shared__ double red[32];
    if (st->done) return;
    const double *cam = pb_cam(pbuf, st->cur_cam);
    double *cam_new = pb_cam(pbuf, st->cur_cam ^ 1);
    const double mu = st->mu;
    double s0 = 0.0, s1 = 0.0;
    for (int t = threadIdx.x; t < 16 * m; t += blockDim.x) {
        const int j = t >> 4, r = t & 15;
        const double d = (j >= mcon) ? da_free[t - 16 * mcon] : 0.0;
        da_full[t] = d;
        cam_new[t] = cam[t] + d;
        s0 += d * d;
        s1 += d * (mu * d + ((j >= mcon) ? Ured[j * kU + 136 + r] : 0.0));
    }
    s0 = block_sum(s0, red); s1 = block_sum(s1, red);
    if (threadIdx.x == 0) { out2[0] = s0; out2[1] = s1; }
}

}  // namespace argus
