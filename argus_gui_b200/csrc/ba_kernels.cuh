// BA kernels, generation 1 ("straight" kernels: one thread per point, block-pair Schur kernel).
// fp64 throughout.  Reference arithmetic replaced: sba-1.6p/sba_levmar.c:781-1300 (see each kernel).
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
__global__ void __launch_bounds__(256) k_residual(BaProblem pb, const double *__restrict__ cam, const double *__restrict__ pts,
                                                  double *__restrict__ hx, double *__restrict__ part)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    CamConst *sc = reinterpret_cast<CamConst *>(smem_raw);
    __shared__ double red[32];
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

// K2+K3: linearise at p (img_projsKDRTS_jac_x, sbaprojs.c:1193-1250) and accumulate
//   U_j = sum A^T A, ea_j = sum A^T e   (sba_levmar.c:837-869)   -> per-block partials Upart[block][m][152]
//   V_i = sum B^T B, eb_i = sum B^T e   (sba_levmar.c:879-907)   -> V[n][6], eb[n][3]
//   W_ij = A^T B                         (sba_levmar.c:912-938)   -> W[nobs][16][3]
// plus per-block partials of {||e||^2, max|eb|, max diag V, sum pts^2} (sba_levmar.c:941-961).
// The Jacobian itself is never written to memory.
__global__ void __launch_bounds__(128) k_linearize(BaProblem pb, const double *__restrict__ cam, const double *__restrict__ pts,
                                                   double *__restrict__ W, double *__restrict__ V, double *__restrict__ eb,
                                                   double *__restrict__ Upart, double *__restrict__ scpart)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    CamConst *sc = reinterpret_cast<CamConst *>(smem_raw);
    double *Us = reinterpret_cast<double *>(sc + pb.m);     // m x 152
    __shared__ double red[32];
    for (int t = threadIdx.x; t < pb.m * kU; t += blockDim.x) Us[t] = 0.0;
    load_cameras(sc, cam, pb.rot0, pb.m);
    double cost = 0.0, ebmax = 0.0, vdmax = -1.79769313486231570e308, ptsq = 0.0;
    for (int64_t base = (int64_t)blockIdx.x * blockDim.x; base < pb.n; base += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = base + threadIdx.x;
        if (i >= pb.n) continue;
        const double M0 = pts[3 * i], M1 = pts[3 * i + 1], M2 = pts[3 * i + 2];
        ptsq += M0 * M0 + M1 * M1 + M2 * M2;
        const bool pfree = i >= pb.ncon;
        double v00 = 0, v01 = 0, v02 = 0, v11 = 0, v12 = 0, v22 = 0, b0 = 0, b1 = 0, b2 = 0;
        const int k1 = pb.obs_ptr[i + 1];
        for (int k = pb.obs_ptr[i]; k < k1; ++k) {
            const int j = pb.obs_cam[k];
            double u, v, A[32], B[6];
            cam_project_jac(sc[j], M0, M1, M2, pb.nfix_k, pb.nfix_d, u, v, A, B);
            const double2 z = pb.obs_uv[k];
            const double e0 = z.x - u, e1 = z.y - v;
            cost += e0 * e0 + e1 * e1;
            if (pfree) {
                v00 += B[0] * B[0] + B[3] * B[3]; v01 += B[0] * B[1] + B[3] * B[4]; v02 += B[0] * B[2] + B[3] * B[5];
                v11 += B[1] * B[1] + B[4] * B[4]; v12 += B[1] * B[2] + B[4] * B[5]; v22 += B[2] * B[2] + B[5] * B[5];
                b0 += B[0] * e0 + B[3] * e1; b1 += B[1] * e0 + B[4] * e1; b2 += B[2] * e0 + B[5] * e1;
            }
            if (j >= pb.mcon) {
                double *Uj = Us + j * kU;
                int q = 0;
#pragma unroll
                for (int r = 0; r < 16; ++r) {
#pragma unroll
                    for (int c = r; c < 16; ++c, ++q) {
                        const double s = A[r] * A[c] + A[16 + r] * A[16 + c];
                        if (s != 0.0) atomicAdd(Uj + q, s);
                    }
                }
#pragma unroll
                for (int r = 0; r < 16; ++r) {
                    const double s = A[r] * e0 + A[16 + r] * e1;
                    if (s != 0.0) atomicAdd(Uj + 136 + r, s);
                }
                if (pfree) {
                    double *Wk = W + 48 * (int64_t)k;
#pragma unroll
                    for (int r = 0; r < 16; ++r) {
                        Wk[3 * r + 0] = A[r] * B[0] + A[16 + r] * B[3];
                        Wk[3 * r + 1] = A[r] * B[1] + A[16 + r] * B[4];
                        Wk[3 * r + 2] = A[r] * B[2] + A[16 + r] * B[5];
                    }
                }
            }
        }
        double *Vi = V + 6 * i;
        Vi[0] = v00; Vi[1] = v01; Vi[2] = v02; Vi[3] = v11; Vi[4] = v12; Vi[5] = v22;
        eb[3 * i] = b0; eb[3 * i + 1] = b1; eb[3 * i + 2] = b2;
        if (pfree) {
            ebmax = fmax(ebmax, fmax(fabs(b0), fmax(fabs(b1), fabs(b2))));
            vdmax = fmax(vdmax, fmax(v00, fmax(v11, v22)));
        }
    }
    __syncthreads();
    for (int t = threadIdx.x; t < pb.m * kU; t += blockDim.x) Upart[(int64_t)blockIdx.x * pb.m * kU + t] = Us[t];
    const double s0 = block_sum(cost, red), s1 = block_max(ebmax, red), s2 = block_max(vdmax, red), s3 = block_sum(ptsq, red);
    if (threadIdx.x == 0) {
        scpart[4 * blockIdx.x + 0] = s0; scpart[4 * blockIdx.x + 1] = s1;
        scpart[4 * blockIdx.x + 2] = s2; scpart[4 * blockIdx.x + 3] = s3;
    }
}

// out[i] = sum_c part[c * stride + i] in fixed order c = 0..nparts-1 (deterministic partial-sum reduction)
__global__ void k_reduce_sum(const double *__restrict__ part, int nparts, int64_t stride, int64_t len, double *__restrict__ out)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= len) return;
    double s = 0.0;
    for (int c = 0; c < nparts; ++c) s += part[c * stride + i];
    out[i] = s;
}

// reduce the linearize scalar partials: out = {cost (sum), ebmax (max), vdmax (max), ptsq (sum)}; one block of 32
__global__ void k_reduce_lin_scalars(const double *__restrict__ scpart, int nparts, double *__restrict__ out_sum2, double *__restrict__ out_max2)
{
    if (threadIdx.x == 0) {
        double c = 0.0, p = 0.0, a = 0.0, b = -1.79769313486231570e308;
        for (int i = 0; i < nparts; ++i) {
            c += scpart[4 * i]; a = fmax(a, scpart[4 * i + 1]); b = fmax(b, scpart[4 * i + 2]); p += scpart[4 * i + 3];
        }
        out_sum2[0] = c; out_sum2[1] = p; out_max2[0] = a; out_max2[1] = b;
    }
}

// After the (optional) all-reduce: host-visible scalars of one linearisation (sba_levmar.c:941-987).
//   Ured: m x 152 (reduced U / ea), sum2 = {cost, sum pts^2}, max2 = {max|eb|, max diag V}
//   scal[0]=cost, [1]=||J^T e||_inf, [2]=||p||^2, [3]=max diag(U,V) over free cameras/points
__global__ void k_lin_finalize(const double *__restrict__ Ured, const double *__restrict__ cam, int m, int mcon,
                               const double *__restrict__ sum2, const double *__restrict__ max2, double *__restrict__ scal)
{
    if (blockIdx.x != 0 || threadIdx.x >= 32) return;
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

// K4a: per point, damp and factor V*_i (sba_levmar.c:992-1016 with sba_symat_invert_BK replaced by a 3x3 LDL^T),
// Y_ij = W_ij V*_i^-1 (sba_levmar.c:1051-1072), c_i = V*_i^-1 eb_i, and the E_j partials sum_i Y_ij eb_i
// (sba_levmar.c:1164-1185) -> Epart[block][m][16].
//   use_state: this is a retry after a rejected step in sba-1.6 compatibility mode: the diagonal to damp is the
//   one the previous attempt left behind (vstate = diag of the previous inverse).
__global__ void __launch_bounds__(128) k_prepare(BaProblem pb, double mu, int use_state, int write_state,
                                                 const double *__restrict__ V, const double *__restrict__ eb,
                                                 double *__restrict__ vstate, const double *__restrict__ W, double *__restrict__ Y,
                                                 double *__restrict__ Vinv, double *__restrict__ cvec,
                                                 double *__restrict__ Epart, int *__restrict__ flag)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *Es = reinterpret_cast<double *>(smem_raw);      // m x 16
    for (int t = threadIdx.x; t < pb.m * 16; t += blockDim.x) Es[t] = 0.0;
    __syncthreads();
    for (int64_t base = (int64_t)blockIdx.x * blockDim.x; base < pb.n; base += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = base + threadIdx.x;
        if (i >= pb.n || i < pb.ncon) continue;
        double v[6];
#pragma unroll
        for (int t = 0; t < 6; ++t) v[t] = V[6 * i + t];
        if (use_state) { v[0] = vstate[3 * i]; v[3] = vstate[3 * i + 1]; v[5] = vstate[3 * i + 2]; }
        v[0] += mu; v[3] += mu; v[5] += mu;
        double li[3], di[3];
        if (!sym3_ldl_inv(v, li, di)) { atomicOr(flag, 1); continue; }
        // inverse = L^-T D^-1 L^-1 with L^-1 = [[1,0,0],[li0,1,0],[li1,li2,1]]
        const double n00 = di[0] + li[0] * li[0] * di[1] + li[1] * li[1] * di[2];
        const double n01 = li[0] * di[1] + li[1] * li[2] * di[2];
        const double n02 = li[1] * di[2];
        const double n11 = di[1] + li[2] * li[2] * di[2];
        const double n12 = li[2] * di[2];
        const double n22 = di[2];
        double *Ni = Vinv + 6 * i;
        Ni[0] = n00; Ni[1] = n01; Ni[2] = n02; Ni[3] = n11; Ni[4] = n12; Ni[5] = n22;
        if (write_state) { vstate[3 * i] = n00; vstate[3 * i + 1] = n11; vstate[3 * i + 2] = n22; }
        const double b0 = eb[3 * i], b1 = eb[3 * i + 1], b2 = eb[3 * i + 2];
        const double c0 = n00 * b0 + n01 * b1 + n02 * b2, c1 = n01 * b0 + n11 * b1 + n12 * b2, c2 = n02 * b0 + n12 * b1 + n22 * b2;
        cvec[3 * i] = c0; cvec[3 * i + 1] = c1; cvec[3 * i + 2] = c2;
        const int k1 = pb.obs_ptr[i + 1];
        for (int k = pb.obs_ptr[i]; k < k1; ++k) {
            const int j = pb.obs_cam[k];
            if (j < pb.mcon) continue;
            const double *Wk = W + 48 * (int64_t)k;
            double *Yk = Y + 48 * (int64_t)k;
#pragma unroll
            for (int r = 0; r < 16; ++r) {
                const double w0 = Wk[3 * r], w1 = Wk[3 * r + 1], w2 = Wk[3 * r + 2];
                Yk[3 * r + 0] = w0 * n00 + w1 * n01 + w2 * n02;
                Yk[3 * r + 1] = w0 * n01 + w1 * n11 + w2 * n12;
                Yk[3 * r + 2] = w0 * n02 + w1 * n12 + w2 * n22;
                const double s = w0 * c0 + w1 * c1 + w2 * c2;     // (Y_ij eb_i)_r = (W_ij c_i)_r
                if (s != 0.0) atomicAdd(Es + 16 * j + r, s);
            }
        }
    }
    __syncthreads();
    for (int t = threadIdx.x; t < pb.m * 16; t += blockDim.x) Epart[(int64_t)blockIdx.x * pb.m * 16 + t] = Es[t];
}

// K4b (generation 1): Schur-complement blocks  sum_i Y_ij W_ik^T  for one camera pair (j <= k) and one chunk of
// points (sba_levmar.c:1075-1114).  grid = (npairs, nchunks), 256 threads = one element of the 16x16 block each.
__global__ void __launch_bounds__(256) k_schur_pairs(BaProblem pb, const double *__restrict__ Y, const double *__restrict__ W,
                                                     int64_t chunk_pts, double *__restrict__ Spart)
{
    const int mf = pb.m - pb.mcon;
    int a = 0, rem = blockIdx.x;
    while (rem >= mf - a) { rem -= mf - a; ++a; }
    const int j = pb.mcon + a, k = j + rem;
    const int r = threadIdx.x >> 4, c = threadIdx.x & 15;
    const uint64_t bj = 1ull << j, bk = 1ull << k;
    int64_t i0 = (int64_t)blockIdx.y * chunk_pts, i1 = i0 + chunk_pts;
    if (i0 < pb.ncon) i0 = pb.ncon;
    if (i1 > pb.n) i1 = pb.n;
    double acc = 0.0;
    for (int64_t i = i0; i < i1; ++i) {
        const uint64_t mk = pb.pt_mask[i];
        if ((mk & bj) && (mk & bk)) {
            const int64_t o = pb.obs_ptr[i];
            const double *y = Y + 48 * (o + __popcll(mk & (bj - 1))) + 3 * r;
            const double *w = W + 48 * (o + __popcll(mk & (bk - 1))) + 3 * c;
            acc += y[0] * w[0] + y[1] * w[1] + y[2] * w[2];
        }
    }
    const int npairs = gridDim.x;
    Spart[((int64_t)blockIdx.y * npairs + blockIdx.x) * 256 + threadIdx.x] = acc;
}

// Assemble the reduced camera system (sba_levmar.c:1116-1185): S_jk = delta_jk U*_j - sum_i Y_ij W_ik^T, mirrored;
// E_j = ea_j - sum_i Y_ij eb_i.  S is N x N row-major (symmetric), N = 16 (m - mcon).
//   Ured: m x 152; Sred: npairs x 256 (pair-major, upper blocks); Ered: m x 16; mu_u: damping added to U's diagonal.
__global__ void k_assemble(int m, int mcon, const double *__restrict__ Ured, const double *__restrict__ Sred,
                           const double *__restrict__ Ered, double mu_u, double *__restrict__ S, double *__restrict__ E)
{
    const int mf = m - mcon, N = 16 * mf;
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < (int64_t)N * N) {
        const int row = (int)(t / N), col = (int)(t % N);
        int a = row >> 4, b = col >> 4, r = row & 15, c = col & 15;
        const bool lower = a > b;
        if (lower) { int q = a; a = b; b = q; q = r; r = c; c = q; }
        const int pair = a * mf - (a * (a - 1)) / 2 + (b - a);
        double s = -Sred[(int64_t)pair * 256 + r * 16 + c];
        if (a == b) {
            const double *Uj = Ured + (mcon + a) * kU;
            s += (r <= c) ? Uj[tri16(r, c)] : Uj[tri16(c, r)];
            if (r == c) s += mu_u;
        }
        S[t] = s;
    }
    if (t < N) {
        const int j = mcon + (int)(t >> 4), r = (int)(t & 15);
        E[t] = Ured[j * kU + 136 + r] - Ered[j * 16 + r];
    }
}

// K5: dense Cholesky solve S da = E (sba_Axb_Chol: dpotrf + dpotrs, sba_lapack.c:373-479) in ONE thread block.
// Blocked right-looking factorisation (panel 32) of the lower triangle, panel kept in shared memory, then
// blocked forward/backward substitution.  status[0] = 1 on success, 0 when a pivot is not positive (the
// reference then increases the damping, sba_levmar.c:1201-1210).  S is overwritten.
constexpr int kCholNB = 32;
constexpr int kCholThreads = 512;
// NB = panel width: 32 by default; 16 for systems whose 32-wide panel would not fit in shared memory (N > 864, i.e. 55..64 free cameras)
template <int NB>
__global__ void __launch_bounds__(kCholThreads) k_chol_solve_t(double *__restrict__ S, const double *__restrict__ E, double *__restrict__ x, int N,
                                                             int *__restrict__ status)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *P = reinterpret_cast<double *>(smem_raw);          // N x 33 panel (also the running rhs later)
    __shared__ double D[NB][NB + 1];
    __shared__ double ys[NB];
    __shared__ int fail;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) fail = 0;
    __syncthreads();
    for (int kb = 0; kb < N; kb += NB) {
        const int nb = min(NB, N - kb);
        // (1) diagonal block (lower triangle, row r >= col c)
        for (int t = tid; t < nb * nb; t += blockDim.x) { const int r = t / nb, c = t % nb; D[r][c] = S[(int64_t)(kb + r) * N + kb + c]; }
        __syncthreads();
        for (int j = 0; j < nb; ++j) {
            if (tid == 0) {
                const double d = D[j][j];
                if (!(d > 0.0)) fail = 1; else D[j][j] = sqrt(d);
            }
            __syncthreads();
            if (fail) break;
            if (tid > j && tid < nb) D[tid][j] /= D[j][j];
            __syncthreads();
            const int w = nb - 1 - j;        // trailing (w x w) lower update
            for (int t = tid; t < w * w; t += blockDim.x) {
                const int r = j + 1 + t / w, c = j + 1 + t % w;
                if (c <= r) D[r][c] -= D[r][j] * D[c][j];
            }
            __syncthreads();
        }
        if (fail) break;
        for (int t = tid; t < nb * nb; t += blockDim.x) { const int r = t / nb, c = t % nb; if (c <= r) S[(int64_t)(kb + r) * N + kb + c] = D[r][c]; }
        // (2) panel: rows below the diagonal block, L_i = A_i D^-T (row i kept in P[i][0..nb))
        const int r0 = kb + nb, R = N - r0;
        for (int i = tid; i < R; i += blockDim.x) {
            double *row = S + (int64_t)(r0 + i) * N + kb;
            double *pr = P + i * (NB + 1);
            for (int c = 0; c < nb; ++c) {
                double s = row[c];
                for (int q = 0; q < c; ++q) s -= pr[q] * D[c][q];
                s /= D[c][c];
                pr[c] = s; row[c] = s;
            }
            for (int c = nb; c < NB; ++c) pr[c] = 0.0;
        }
        __syncthreads();
        // (3) trailing update of the lower triangle: 4x4 register tiles
        const int T = (R + 3) >> 2;
        for (int t = tid; t < T * T; t += blockDim.x) {
            const int ti = t / T, tj = t % T;
            if (tj > ti) continue;
            double acc[4][4];
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) acc[a][b] = 0.0;
            const int i0 = 4 * ti, j0 = 4 * tj;
            for (int c = 0; c < NB; ++c) {
                double pi[4], pj[4];
#pragma unroll
                for (int a = 0; a < 4; ++a) { pi[a] = (i0 + a < R) ? P[(i0 + a) * (NB + 1) + c] : 0.0; pj[a] = (j0 + a < R) ? P[(j0 + a) * (NB + 1) + c] : 0.0; }
#pragma unroll
                for (int a = 0; a < 4; ++a)
#pragma unroll
                    for (int b = 0; b < 4; ++b) acc[a][b] += pi[a] * pj[b];
            }
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    const int i = i0 + a, j = j0 + b;
                    if (i < R && j <= i) S[(int64_t)(r0 + i) * N + r0 + j] -= acc[a][b];
                }
        }
        __syncthreads();
    }
    if (fail) { if (tid == 0) status[0] = 0; return; }
    // forward substitution L y = E (blocked by 32): rhs[0..N) is the running right-hand side
    double *rhs = P;
    for (int i = tid; i < N; i += blockDim.x) rhs[i] = E[i];
    __syncthreads();
    for (int kb = 0; kb < N; kb += NB) {
        const int nb = min(NB, N - kb);
        if (warp == 0) {      // warp 0 solves the diagonal block
            double b = (lane < nb) ? rhs[kb + lane] : 0.0;
            for (int j = 0; j < nb; ++j) {
                const double ljj = S[(int64_t)(kb + j) * N + kb + j];
                const double yj = __shfl_sync(0xffffffffu, b, j) / ljj;
                if (lane == j) b = yj;
                else if (lane > j && lane < nb) b -= S[(int64_t)(kb + lane) * N + kb + j] * yj;
            }
            if (lane < nb) { ys[lane] = b; rhs[kb + lane] = b; }
        }
        __syncthreads();
        for (int i = kb + nb + tid; i < N; i += blockDim.x) {
            const double *row = S + (int64_t)i * N + kb;
            double s = rhs[i];
            for (int c = 0; c < nb; ++c) s -= row[c] * ys[c];
            rhs[i] = s;
        }
        __syncthreads();
    }
    // backward substitution L^T x = y
    for (int kb = ((N - 1) / NB) * NB; kb >= 0; kb -= NB) {
        const int nb = min(NB, N - kb);
        if (warp == 0) {
            double b = (lane < nb) ? rhs[kb + lane] : 0.0;
            for (int j = nb - 1; j >= 0; --j) {
                const double ljj = S[(int64_t)(kb + j) * N + kb + j];
                const double xj = __shfl_sync(0xffffffffu, b, j) / ljj;
                if (lane == j) b = xj;
                else if (lane < j) b -= S[(int64_t)(kb + j) * N + kb + lane] * xj;
            }
            if (lane < nb) { ys[lane] = b; rhs[kb + lane] = b; }
        }
        __syncthreads();
        for (int i = tid; i < kb; i += blockDim.x) {
            double s = rhs[i];
            for (int c = 0; c < nb; ++c) s -= S[(int64_t)(kb + c) * N + i] * ys[c];
            rhs[i] = s;
        }
        __syncthreads();
    }
    for (int i = tid; i < N; i += blockDim.x) x[i] = rhs[i];
    if (tid == 0) status[0] = 1;
}

// K5, second generation: same contract as k_chol_solve (one CTA, S overwritten by its lower Cholesky factor, status[0] = 1 / 0),
// restructured so that no long dependent chain touches global memory:
//   * the 32x32 diagonal block is factored in shared memory by the whole CTA;
//   * its inverse (lower triangular) is formed by all warps (one column per warp pass, shuffle broadcast) and kept in
//     work[kb/32][32][32]: the panel solve L = A D^-T becomes a small matrix product with full instruction-level
//     parallelism, and both substitutions become per-block triangular matrix-vector products + parallel updates.
// work: (ceil(N/32)) x 1024 doubles of scratch.
__global__ void __launch_bounds__(kCholThreads) k_chol_solve2(double *__restrict__ S, const double *__restrict__ E, double *__restrict__ x, int N,
                                                              double *__restrict__ work, int *__restrict__ status)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *P = reinterpret_cast<double *>(smem_raw);          // rows below the panel x 33 (also the running rhs later)
    __shared__ double D[kCholNB][kCholNB + 1];
    __shared__ double Di[kCholNB][kCholNB + 1];
    __shared__ double ys[kCholNB];
    __shared__ int fail;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = blockDim.x >> 5;
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
__global__ void k_cam_update(int m, int mcon, const double *__restrict__ cam, const double *__restrict__ da_free,
                             const double *__restrict__ Ured, double mu, double *__restrict__ cam_new, double *__restrict__ da_full,
                             double *__restrict__ out2)
{
    __shared__ double red[32];
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

// K6 + K1: back-substitution db_i = V*_i^-1 (eb_i - sum_j W_ij^T da_j) (sba_levmar.c:1215-1256), trial point
// pdp = p + dp, and the cost at the trial point (func(pdp) + nrmL2xmy, sba_levmar.c:1281-1288).
// Per-block partials part[block] = {||e_new||^2, ||db||^2, sum db (mu db + eb)}.
__global__ void __launch_bounds__(128) k_backsub_eval(BaProblem pb, double mu, const double *__restrict__ cam_new,
                                                      const double *__restrict__ da, const double *__restrict__ pts,
                                                      const double *__restrict__ W, const double *__restrict__ Vinv,
                                                      const double *__restrict__ eb, double *__restrict__ pts_new,
                                                      double *__restrict__ dbout, double *__restrict__ part)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    CamConst *sc = reinterpret_cast<CamConst *>(smem_raw);
    double *das = reinterpret_cast<double *>(sc + pb.m);     // 16 m
    __shared__ double red[32];
    for (int t = threadIdx.x; t < 16 * pb.m; t += blockDim.x) das[t] = da[t];
    load_cameras(sc, cam_new, pb.rot0, pb.m);
    double cost = 0.0, dbsq = 0.0, dl = 0.0;
    for (int64_t base = (int64_t)blockIdx.x * blockDim.x; base < pb.n; base += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = base + threadIdx.x;
        if (i >= pb.n) continue;
        double d0 = 0.0, d1 = 0.0, d2 = 0.0;
        const int k0 = pb.obs_ptr[i], k1 = pb.obs_ptr[i + 1];
        if (i >= pb.ncon) {
            const double b0 = eb[3 * i], b1 = eb[3 * i + 1], b2 = eb[3 * i + 2];
            double t0 = b0, t1 = b1, t2 = b2;
            for (int k = k0; k < k1; ++k) {
                const int j = pb.obs_cam[k];
                if (j < pb.mcon) continue;
                const double *Wk = W + 48 * (int64_t)k;
                const double *dj = das + 16 * j;
#pragma unroll
                for (int r = 0; r < 16; ++r) { t0 -= Wk[3 * r] * dj[r]; t1 -= Wk[3 * r + 1] * dj[r]; t2 -= Wk[3 * r + 2] * dj[r]; }
            }
            const double *Ni = Vinv + 6 * i;
            d0 = Ni[0] * t0 + Ni[1] * t1 + Ni[2] * t2;
            d1 = Ni[1] * t0 + Ni[3] * t1 + Ni[4] * t2;
            d2 = Ni[2] * t0 + Ni[4] * t1 + Ni[5] * t2;
            dbsq += d0 * d0 + d1 * d1 + d2 * d2;
            dl += d0 * (mu * d0 + b0) + d1 * (mu * d1 + b1) + d2 * (mu * d2 + b2);
        }
        if (dbout) { dbout[3 * i] = d0; dbout[3 * i + 1] = d1; dbout[3 * i + 2] = d2; }
        const double M0 = pts[3 * i] + d0, M1 = pts[3 * i + 1] + d1, M2 = pts[3 * i + 2] + d2;
        pts_new[3 * i] = M0; pts_new[3 * i + 1] = M1; pts_new[3 * i + 2] = M2;
        for (int k = k0; k < k1; ++k) {
            double u, v;
            cam_project(sc[pb.obs_cam[k]], M0, M1, M2, u, v);
            const double2 z = pb.obs_uv[k];
            const double e0 = z.x - u, e1 = z.y - v;
            cost += e0 * e0 + e1 * e1;
        }
    }
    const double s0 = block_sum(cost, red), s1 = block_sum(dbsq, red), s2 = block_sum(dl, red);
    if (threadIdx.x == 0) { part[3 * blockIdx.x] = s0; part[3 * blockIdx.x + 1] = s1; part[3 * blockIdx.x + 2] = s2; }
}

// pack the scalars of one attempt for the host: att = {solved, ||dp||^2, dL, ||e_new||^2, singular V* flag}
__global__ void k_attempt_finalize(const int *__restrict__ flags, const double *__restrict__ cam2, const double *__restrict__ pt3,
                                   double *__restrict__ att)
{
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        att[0] = (double)flags[1];
        att[1] = cam2[0] + pt3[1];
        att[2] = cam2[1] + pt3[2];
        att[3] = pt3[0];
        att[4] = (double)flags[0];
    }
}

}  // namespace argus
