// K5, multi-SM version: the reduced camera system S da = E (sba_Axb_Chol: dpotrf + dpotrs, sba_lapack.c:373-479) factored by
// a COOPERATIVE grid (one CTA per SM, grid-wide barriers) instead of one CTA.  The single-CTA kernels are latency bound
// (0.36 ms at 256 unknowns, 1.6 ms at 512, several ms at 864); here the O(N^3) parts - panel solve and trailing update - are
// spread over all SMs and only the 32 x 32 diagonal factorisations and the substitutions stay serial:
//   per 32-column panel:  [owner CTA: factor the diagonal block, invert it]  grid barrier
//                         [all: L_panel = A_panel D^-T, one 32 x 32 block per CTA pass]  grid barrier
//                         [all: trailing 32 x 32 tiles -= L_i L_j^T]   (the CTA that updates the next diagonal tile goes on to
//                                                                        factor it before the barrier)
//   then CTA 0: forward and backward substitution with the stored block inverses.
// Same contract as k_chol_solve: S (N x N, symmetric, row-major) is overwritten by its lower factor, x = S^-1 E,
// status[0] = 1, or 0 when a pivot is not positive (LAPACK's dpotrf failure -> the LM loop raises the damping).
// work: ceil(N/32) x 1024 doubles (the inverses of the diagonal blocks) + N doubles (running right-hand side).
#pragma once
#include <cooperative_groups.h>
#include "ba_common.cuh"

namespace argus {

constexpr int kCoopThreads = 256;

// factor the 32 x 32 diagonal block at kb (lower triangle of S), write L back and its inverse to Dinv (global).
// The factorisation and the inversion are the serial part of every panel, so they run in ONE warp without block barriers:
// lane i keeps row i of the block in registers, the pivot row travels by shuffles (right-looking Cholesky, 1/sqrt(pivot) from
// rsqrt); then lane c forms column c of L^-1 by forward substitution with L read from shared memory (broadcast loads).
__device__ void coop_factor_diag(double *__restrict__ S, int N, int kb, double *__restrict__ Dinv, int *fail,
                                 double (*D)[33], double (*Di)[33])
{
    __shared__ double s_rinv[32];
    __shared__ int s_fail;
    const int tid = threadIdx.x, nb = min(32, N - kb);
    for (int t = tid; t < 32 * 32; t += blockDim.x) {
        const int r = t >> 5, c = t & 31;
        D[r][c] = (r < nb && c < nb && c <= r) ? S[(int64_t)(kb + r) * N + kb + c] : ((r == c) ? 1.0 : 0.0);
    }
    if (tid == 0) s_fail = 0;
    __syncthreads();
    if (tid < 32) {
        const int lane = tid;
        double a[32];
#pragma unroll
        for (int c = 0; c < 32; ++c) a[c] = D[lane][c];
        bool bad = false;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            const double piv = __shfl_sync(0xffffffffu, a[j], j);
            if (!(piv > 0.0)) bad = true;
            const double ri = rsqrt(piv);
            const double lij = a[j] * ri;                       // lane j: sqrt(pivot); lanes below: l_ij
            a[j] = lij;
            if (lane == j) s_rinv[j] = ri;
#pragma unroll
            for (int c = j + 1; c < 32; ++c) {
                const double lc = __shfl_sync(0xffffffffu, lij, c);
                a[c] -= lij * lc;                                // meaningful for c <= lane
            }
        }
        if (bad) { if (lane == 0) s_fail = 1; }
#pragma unroll
        for (int c = 0; c < 32; ++c) D[lane][c] = (c <= lane) ? a[c] : 0.0;
        __syncwarp();
        // column `lane` of L^-1: x_r = (delta_r,lane - sum_{q<r} L[r][q] x_q) / L[r][r]   (x_q = 0 for q < lane)
        double x[32];
#pragma unroll
        for (int r = 0; r < 32; ++r) {
            double s0 = (r == lane) ? 1.0 : 0.0, s1 = 0.0;
#pragma unroll
            for (int q = 0; q + 1 < r; q += 2) { s0 -= D[r][q] * x[q]; s1 -= D[r][q + 1] * x[q + 1]; }
            if (r & 1) s0 -= D[r][r - 1] * x[r - 1];
            x[r] = (r >= lane) ? (s0 + s1) * s_rinv[r] : 0.0;
        }
#pragma unroll
        for (int r = 0; r < 32; ++r) Di[r][lane] = x[r];
    }
    __syncthreads();
    if (s_fail) { if (tid == 0) { *(volatile int *)fail = 1; __threadfence(); } __syncthreads(); return; }
    for (int t = tid; t < 32 * 32; t += blockDim.x) {
        const int r = t >> 5, c = t & 31;
        Dinv[t] = Di[r][c];
        if (r < nb && c < nb && c <= r) S[(int64_t)(kb + r) * N + kb + c] = D[r][c];
    }
    __syncthreads();                                                // D / Di are the caller's tile buffers
}

__global__ void __launch_bounds__(kCoopThreads) k_chol_coop(double *__restrict__ S, const double *__restrict__ E, double *__restrict__ x,
                                                            int N, double *__restrict__ work, int *__restrict__ status,
                                                            int *fail /* zero on entry (the controller clears it); written by the CTA that factors a diagonal block */,
                                                            const int *__restrict__ done)
{
    if (*done) return;
    namespace cg = cooperative_groups;
    cg::grid_group grid = cg::this_grid();
    __shared__ double A[32][33], B[32][33];
    const int tid = threadIdx.x, nblk = (N + 31) / 32;
    double *rhs = work + (size_t)nblk * 1024;

    if (blockIdx.x == 0) coop_factor_diag(S, N, 0, work, fail, A, B);
    grid.sync();
    for (int kbi = 0; kbi < nblk; ++kbi) {
        if (*(volatile int *)fail) break;                           // read after a grid barrier: every CTA sees the same value
        const int kb = kbi * 32;
        const double *Dinv = work + (size_t)kbi * 1024;
        // ---- panel: L[i][kb-block] = A[i][kb-block] * Dinv^T for the row blocks i > kbi ----
        for (int i = kbi + 1 + blockIdx.x; i < nblk; i += gridDim.x) {
            const int r0 = i * 32, nr = min(32, N - r0);
            for (int t = tid; t < 1024; t += blockDim.x) {
                const int r = t >> 5, c = t & 31;
                A[r][c] = (r < nr) ? S[(int64_t)(r0 + r) * N + kb + c] : 0.0;
                B[r][c] = Dinv[t];
            }
            __syncthreads();
            for (int t = tid; t < 1024; t += blockDim.x) {
                const int r = t >> 5, c = t & 31;
                double s = 0.0;
#pragma unroll 8
                for (int q = 0; q <= c; ++q) s += A[r][q] * B[c][q];       // (Dinv^T)[q][c] = Dinv[c][q], lower triangular
                if (r < nr) S[(int64_t)(r0 + r) * N + kb + c] = s;
            }
            __syncthreads();
        }
        grid.sync();
        // ---- trailing update: tile (i, j), kbi < j <= i, -= L_i L_j^T; the owner of the next diagonal tile factors it ----
        const int R = nblk - kbi - 1;
        const int ntile = R * (R + 1) / 2;
        // CTA 0 takes only tile 0 (the next diagonal block) and factors it while the other CTAs share the remaining tiles
        const int nworker = (gridDim.x > 1) ? (int)gridDim.x - 1 : 1;
        const int tfirst = (gridDim.x > 1) ? ((blockIdx.x == 0) ? 0 : (int)blockIdx.x) : 0;
        const int tstep = (gridDim.x > 1) ? ((blockIdx.x == 0) ? ntile : nworker) : 1;
        for (int tix = tfirst; tix < ntile; tix += tstep) {
            int ti = 0, rem = tix;                                  // row-major lower triangle: tile tix -> (ti, tj)
            while (rem > ti) { rem -= ti + 1; ++ti; }
            const int tj = rem;
            const int i = kbi + 1 + ti, j = kbi + 1 + tj;
            const int r0 = i * 32, c0 = j * 32, nr = min(32, N - r0), nc = min(32, N - c0);
            for (int t = tid; t < 1024; t += blockDim.x) {
                const int r = t >> 5, c = t & 31;
                A[r][c] = (r < nr) ? S[(int64_t)(r0 + r) * N + kb + c] : 0.0;
                B[r][c] = (r < nc) ? S[(int64_t)(c0 + r) * N + kb + c] : 0.0;
            }
            __syncthreads();
            for (int t = tid; t < 1024; t += blockDim.x) {
                const int r = t >> 5, c = t & 31;
                if (r < nr && c < nc && (i > j || c <= r)) {
                    double s = 0.0;
#pragma unroll 8
                    for (int q = 0; q < 32; ++q) s += A[r][q] * B[c][q];
                    S[(int64_t)(r0 + r) * N + c0 + c] -= s;
                }
            }
            __syncthreads();
            if (ti == 0) {                                          // tile (kbi + 1, kbi + 1): the next diagonal block is final
                __threadfence_block();
                coop_factor_diag(S, N, (kbi + 1) * 32, work + (size_t)(kbi + 1) * 1024, fail, A, B);
            }
        }
        grid.sync();
    }
    if (*(volatile int *)fail) { if (blockIdx.x == 0 && tid == 0) status[0] = 0; return; }
    if (blockIdx.x != 0) return;
    // ---- substitutions (CTA 0): L y = E, L^T x = y, block by block with the stored inverses ----
    __shared__ double ys[32];
    for (int i = tid; i < N; i += blockDim.x) rhs[i] = E[i];
    __syncthreads();
    for (int b = 0; b < nblk; ++b) {
        const int kb = b * 32, nb = min(32, N - kb);
        const double *Dinv = work + (size_t)b * 1024;
        if (tid < 32) {
            double s = 0.0;
            if (tid < nb) for (int q = 0; q <= tid; ++q) s += Dinv[tid * 32 + q] * rhs[kb + q];
            ys[tid] = s;
        }
        __syncthreads();
        if (tid < nb) rhs[kb + tid] = ys[tid];
        for (int i = kb + nb + tid; i < N; i += blockDim.x) {
            const double *row = S + (int64_t)i * N + kb;
            double s = rhs[i];
            for (int q = 0; q < nb; ++q) s -= row[q] * ys[q];
            rhs[i] = s;
        }
        __syncthreads();
    }
    for (int b = nblk - 1; b >= 0; --b) {
        const int kb = b * 32, nb = min(32, N - kb);
        const double *Dinv = work + (size_t)b * 1024;
        if (tid < 32) {                                             // x_b = Dinv^T y_b
            double s = 0.0;
            if (tid < nb) for (int q = tid; q < nb; ++q) s += Dinv[q * 32 + tid] * rhs[kb + q];
            ys[tid] = s;
        }
        __syncthreads();
        if (tid < nb) rhs[kb + tid] = ys[tid];
        for (int i = tid; i < kb; i += blockDim.x) {                // rhs[i] -= sum_q L[kb + q][i] x_b[q]
            double s = rhs[i];
            for (int q = 0; q < nb; ++q) s -= S[(int64_t)(kb + q) * N + i] * ys[q];
            rhs[i] = s;
        }
        __syncthreads();
    }
    for (int i = tid; i < N; i += blockDim.x) x[i] = rhs[i];
    if (tid == 0) status[0] = 1;
}

}  // namespace argus
