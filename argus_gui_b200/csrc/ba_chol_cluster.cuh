// K5, thread-block-cluster version: the reduced camera system S da = E (sba_Axb_Chol: dpotrf + dpotrs, sba_lapack.c:373-479)
// for N <= 512 unknowns (up to 32 free cameras), factored by ONE cluster of 8 CTAs.
//
// Why.  The solve sits between the Schur complement and the back-substitution of every damping attempt and nothing can
// overlap it; it does not shrink when the points are sharded over more GPUs, so it is THE fixed cost of strong scaling
// (0.24 ms of a 1.3 ms attempt at 8 GPUs on BASELINE config 4) and a seventh of an attempt at config 3.  It is latency bound:
// 5.6 Mflop at N = 256.  The cooperative-grid kernel (ba_chol_coop.cuh) pays two grid-wide barriers (2-3 us each), the inversion
// of every diagonal block and serial substitutions; here
//   * the CTAs synchronise with hardware cluster barriers (barrier.cluster, a fraction of a microsecond),
//   * the right-hand side travels through the factorisation as one more row of the matrix (row N of [S; E^T]): after the
//     last panel it holds y = L^-1 E - no forward substitution,
//   * the panel solve is a per-row triangular solve in registers (one thread per row, L_kk broadcast from shared memory) - no
//     inverse of the diagonal block on the critical path,
//   * the back-substitution keeps its vector in shared memory, loads a block's rows of L before the 32-step triangular solve it
//     does not depend on, and the next diagonal block while the current one is used.
// Per 32-column panel k:  [all groups: rows below the panel  <-  rows * L_kk^-T]                      cluster barrier
//                         [group 0: next diagonal tile -= L L^T, factor it (one warp, registers + shuffles);
//                          groups of the other CTAs: the remaining trailing tiles -= L_i L_j^T]     cluster barrier
// A group = 128 threads (4 per CTA, 32 per cluster) with its own pair of 32 x 32 tiles in shared memory and its own named barrier.
// The matrix stays in global memory (L2): tiles written by one CTA are read by others after a cluster barrier (release /
// acquire at cluster scope) with ld.global.cg.
// Same contract as k_chol_coop: S (N x N, symmetric, row-major) is overwritten by its lower factor, x = S^-1 E,
// status[0] = 1, or 0 when a pivot is not positive (LAPACK's dpotrf failure -> the LM loop raises the damping).
// work: N + 32 + 1056 ceil(N / 32) doubles (the right-hand side row, the reciprocals of the factor's diagonal, the inverses of
// its diagonal blocks).
#pragma once
#include "ba_common.cuh"

namespace argus {

constexpr int kClCtas = 8, kClThreads = 512, kClGroups = 4, kClGT = 128;
constexpr int kClLdA = 34;        // row stride of the A tile (even: rows are read as 16-byte pairs)
constexpr int kClLdB = 33;        // row stride of the B tile (read down a column, one row per lane)
constexpr int kClMaxN = 512;
struct __align__(16) ClBuf { double A[32 * kClLdA]; double B[32 * kClLdB + 1]; };
__host__ __device__ constexpr size_t chol_cluster_smem() { return kClGroups * sizeof(ClBuf) + (kClMaxN + 32) * sizeof(double); }

#ifdef ARGUS_CHOL_TRACE
// debug builds only (tools/chol_bench.cu): clock64 stamps of group 0 of CTA 0 (the critical path), in program order
__device__ long long *g_cl_trace = nullptr;
__device__ int g_cl_trace_n = 0;
#define CL_TRACE() do { if (g_cl_trace && blockIdx.x == 0 && threadIdx.x == 0) g_cl_trace[g_cl_trace_n++] = clock64(); } while (0)
#else
#define CL_TRACE() do { } while (0)
#endif

__device__ __forceinline__ void cl_group_bar(int grp) { asm volatile("bar.sync %0, %1;" ::"r"(grp + 1), "r"(kClGT) : "memory"); }
__device__ __forceinline__ void cl_cluster_bar()
{
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void cl_cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cl_cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
// 1 / sqrt(x) on the pivot chain of the factorisation, branch free (the unrolled factorisation stays one basic block, which lets
// the updates of one step fill the latency of the next step's pivot): MUFU seed from the upper word (rsqrt.approx.ftz.f64,
// relative error below 2^-19 over the whole double range) and one third-order step, y (1 + e/2 + 3 e^2 / 8), e = 1 - x y^2
__device__ __forceinline__ double cl_rsqrt(double x)
{
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double e = fma(-x * y, y, 1.0);
    return fma(y * e, fma(0.375, e, 0.5), y);
}
// row r of row block ib; block nblk is the right-hand side (one row, kept in `rhs`)
__device__ __forceinline__ double *cl_rowp(double *S, double *rhs, int N, int nblk, int ib, int r)
{
    return ib < nblk ? S + (int64_t)(ib * 32 + r) * N : rhs;
}
__device__ __forceinline__ int cl_rows(int N, int nblk, int ib) { return ib < nblk ? min(32, N - ib * 32) : 1; }

// 32 x 32 tile of doubles, rows `nr` x columns `nc` valid (the rest reads as `pad` off the diagonal / `dpad` on it), row r at
// row0 + r * ld (ld = 0: the single right-hand-side row), into shared memory with row stride lds.  All eight loads of a thread
// are issued before the first one is used: the tiles come from L2 and a load -> store -> load sequence costs a round trip each.
__device__ __forceinline__ void cl_load_tile(double *dst, int lds, const double *row0, int64_t ld, int nr, int nc, bool lower_only, double dpad, int gt)
{
    double v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int t = gt + i * kClGT, r = t >> 5, c = t & 31;
        const bool ok = r < nr && c < nc && (!lower_only || c <= r);
        const double *p = row0 + (int64_t)(ok ? r : 0) * ld + (ok ? c : 0);
        v[i] = __ldcg(p);
        if (!ok) v[i] = (r == c) ? dpad : 0.0;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) { const int t = gt + i * kClGT; dst[(t >> 5) * lds + (t & 31)] = v[i]; }
}

// Tile (ib, jb) of the trailing matrix -= L_ib L_jb^T over the panel columns [kb, kb + 32)  (have_panel = false: no update, the
// tile as it is).  to_smem: the result (a diagonal tile, ib == jb) goes to T.B (identity-padded, upper part zero) to be factored.
__device__ __forceinline__ void cl_tile_update(double *S, double *rhs, int N, int nblk, int kb, int ib, int jb, ClBuf &T, int gt, int grp,
                                               bool have_panel, bool to_smem)
{
    const int nr = cl_rows(N, nblk, ib), nc = cl_rows(N, nblk, jb), nb = min(32, N - kb);
    const bool diag = ib == jb;
    const int w = gt >> 5, c = gt & 31;
    double old[8], acc[8];
    bool valid[8];
    const double *irow0 = cl_rowp(S, rhs, N, nblk, ib, 0);
    const int64_t ild = ib < nblk ? N : 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int r = 8 * w + k;
        valid[k] = r < nr && c < nc && (!diag || c <= r);
        old[k] = __ldcg(irow0 + (valid[k] ? (int64_t)r * ild + jb * 32 + c : 0));
        acc[k] = 0.0;
    }
    if (have_panel) {
        cl_load_tile(T.A, kClLdA, irow0 + kb, ild, nr, nb, false, 0.0, gt);
        cl_load_tile(T.B, kClLdB, S + (int64_t)(jb * 32) * N + kb, N, nc, nb, false, 0.0, gt);
    }
    cl_group_bar(grp);
    if (have_panel && 8 * w < nr) {
#pragma unroll 4
        for (int q = 0; q < 32; q += 2) {
            const double b0 = T.B[c * kClLdB + q], b1 = T.B[c * kClLdB + q + 1];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const double2 a = *reinterpret_cast<const double2 *>(&T.A[(8 * w + k) * kClLdA + q]);
                acc[k] = fma(a.x, b0, acc[k]);
                acc[k] = fma(a.y, b1, acc[k]);
            }
        }
    }
    if (to_smem) {
        cl_group_bar(grp);                                  // every dot product has read T.B
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int r = 8 * w + k;
            T.B[r * kClLdB + c] = valid[k] ? old[k] - acc[k] : ((r == c) ? 1.0 : 0.0);
        }
    } else {
#pragma unroll
        for (int k = 0; k < 8; ++k)
            if (valid[k]) const_cast<double *>(irow0)[(int64_t)(8 * w + k) * ild + jb * 32 + c] = old[k] - acc[k];
    }
    cl_group_bar(grp);
}

// Factor the diagonal tile in T.B (32 x 32, lower triangle, identity-padded) with one warp: lane i keeps row i in registers, the
// pivot row travels by shuffles (right-looking Cholesky, 1/sqrt(pivot) from rsqrt); the factor goes to S.
__device__ __forceinline__ void cl_factor_store(double *S, double *rdiag, int N, int kb, ClBuf &T, int gt, int grp, int *fail)
{
    const int nb = min(32, N - kb);
    if ((gt >> 5) == 0) {
        const int lane = gt & 31;
        double a[32], myri = 1.0;
#pragma unroll
        for (int c = 0; c < 32; ++c) a[c] = T.B[lane * kClLdB + c];
        bool bad = false;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            const double piv = __shfl_sync(0xffffffffu, a[j], j);
            if (!(piv > 0.0)) bad = true;
            const double ri = cl_rsqrt(piv);
            const double lij = a[j] * ri;                       // lane j: sqrt(pivot); lanes below: l_ij
            a[j] = lij;
            if (lane == j) myri = ri;
#pragma unroll
            for (int c = j + 1; c < 32; ++c) {
                const double lc = __shfl_sync(0xffffffffu, lij, c);
                a[c] = fma(-lij, lc, a[c]);                      // meaningful for c <= lane
            }
        }
        if (bad && lane == 0) { *(volatile int *)fail = 1; __threadfence(); }
        rdiag[kb + lane] = myri;                                // 1 / L_jj for the panel solves and the back-substitution
#pragma unroll
        for (int c = 0; c < 32; ++c) T.B[lane * kClLdB + c] = (c <= lane) ? a[c] : 0.0;
    }
    cl_group_bar(grp);
    for (int t = gt; t < 1024; t += kClGT) {
        const int r = t >> 5, c = t & 31;
        if (r < nb && c <= r) S[(int64_t)(kb + r) * N + kb + c] = T.B[r * kClLdB + c];
    }
    cl_group_bar(grp);
}

// rows of row block ib, panel columns [kb, kb + 32):  X <- X L_kk^-T.  Four lanes per row (lane part p keeps the columns p, p + 4, ...,
// eight values): x_c is formed by its owner and handed to the other three by a width-4 shuffle, so a step of the forward
// substitution costs a multiplication, a shuffle and a fused multiply-add on the chain and 8 of them off it; L_kk is read from
// shared memory (four addresses per warp instruction, in different banks).  The solved rows stay in T.A (keep = true: no
// trailing barrier, the caller goes on with them).
__device__ __forceinline__ void cl_panel_solve(double *S, double *rhs, const double *rdiag, int N, int nblk, int kb, int ib, ClBuf &T, int gt, int grp,
                                               bool keep)
{
    const int nr = cl_rows(N, nblk, ib), nb = min(32, N - kb);
    const int part = gt & 3;
    double rinv[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) rinv[k] = __ldcg(rdiag + kb + part + 4 * k);       // 1 / L_cc, left behind by the factorisation
    cl_load_tile(T.A, kClLdA, cl_rowp(S, rhs, N, nblk, ib, 0) + kb, ib < nblk ? N : 0, nr, nb, false, 0.0, gt);
    cl_load_tile(T.B, kClLdB, S + (int64_t)kb * N + kb, N, nb, nb, true, 1.0, gt);
    cl_group_bar(grp);
    {
        const int r = gt >> 2;
        double a[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) a[k] = T.A[r * kClLdA + part + 4 * k];
#pragma unroll
        for (int c = 0; c < 32; ++c) {
            const double xc = __shfl_sync(0xffffffffu, a[c >> 2] * rinv[c >> 2], c & 3, 4);
            if (part == (c & 3)) a[c >> 2] = xc;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                if (4 * k + 3 > c) {                                   // (compile time) some lane of this k still has a column behind c
                    const int q = part + 4 * k;
                    const double l = T.B[q * kClLdB + c];              // zero above the diagonal: columns q <= c are left alone
                    if (q > c) a[k] = fma(-xc, l, a[k]);
                }
            }
        }
        cl_group_bar(grp);                                             // (all loads of T.A are done)
#pragma unroll
        for (int k = 0; k < 8; ++k) T.A[r * kClLdA + part + 4 * k] = a[k];
    }
    cl_group_bar(grp);
    for (int t = gt; t < 1024; t += kClGT) {
        const int r = t >> 5, c = t & 31;
        if (r < nr && c < nb) cl_rowp(S, rhs, N, nblk, ib, r)[kb + c] = T.A[r * kClLdA + c];
    }
    if (!keep) cl_group_bar(grp);
}

// L_kk^-1 (lower triangular, 32 x 32) -> inv (global, row-major 32 x 32, zero above the diagonal): one warp, lane = column of
// the inverse, forward substitution with the accumulators of all rows in registers.  Off the critical path (a group that has no
// panel unit does it while the others solve and update); the back-substitution then multiplies instead of solving.
__device__ __forceinline__ void cl_invert_diag(const double *S, const double *rdiag, double *inv, int N, int kb, ClBuf &T, int gt, int grp)
{
    const int nb = min(32, N - kb);
    cl_load_tile(T.B, kClLdB, S + (int64_t)kb * N + kb, N, nb, nb, true, 1.0, gt);
    cl_group_bar(grp);
    if ((gt >> 5) == 0) {
        const int lane = gt & 31;
        const double ri = (lane < nb) ? __ldcg(rdiag + kb + lane) : 1.0;
        double a[32];
#pragma unroll
        for (int r = 0; r < 32; ++r) a[r] = (r == lane) ? 1.0 : 0.0;
#pragma unroll
        for (int q = 0; q < 32; ++q) {                         // x_q = a_q / L_qq;  a_r -= L_rq x_q for r > q  (x_q = 0 for q < lane)
            const double xq = a[q] * __shfl_sync(0xffffffffu, ri, q);
            a[q] = xq;
#pragma unroll
            for (int r = q + 1; r < 32; ++r) a[r] = fma(-T.B[r * kClLdB + q], xq, a[r]);
        }
#pragma unroll
        for (int r = 0; r < 32; ++r) T.A[r * kClLdA + lane] = a[r];
    }
    cl_group_bar(grp);
    for (int t = gt; t < 1024; t += kClGT) inv[t] = T.A[(t >> 5) * kClLdA + (t & 31)];
    cl_group_bar(grp);
}

// The next diagonal tile for the group that has just solved its row block (the solved rows X are still in T.A):
// D = old - X X^T (lower triangle, identity-padded) -> T.B, ready to be factored.  old[] was loaded before the panel solve.
__device__ __forceinline__ void cl_diag_from_panel(ClBuf &T, int gt, int grp, const double (&old)[8], const bool (&valid)[8])
{
    for (int t = gt; t < 1024; t += kClGT) { const int r = t >> 5, q = t & 31; T.B[r * kClLdB + q] = T.A[r * kClLdA + q]; }
    cl_group_bar(grp);
    const int w = gt >> 5, c = gt & 31;
    double acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = 0.0;
#pragma unroll 4
    for (int q = 0; q < 32; q += 2) {
        const double b0 = T.B[c * kClLdB + q], b1 = T.B[c * kClLdB + q + 1];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const double2 a = *reinterpret_cast<const double2 *>(&T.A[(8 * w + k) * kClLdA + q]);
            acc[k] = fma(a.x, b0, acc[k]);
            acc[k] = fma(a.y, b1, acc[k]);
        }
    }
    cl_group_bar(grp);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int r = 8 * w + k;
        T.B[r * kClLdB + c] = valid[k] ? old[k] - acc[k] : ((r == c) ? 1.0 : 0.0);
    }
    cl_group_bar(grp);
}

__global__ void __cluster_dims__(kClCtas, 1, 1) __launch_bounds__(kClThreads, 1)
k_chol_cluster(double *__restrict__ S, const double *__restrict__ E, double *__restrict__ x, int N, double *__restrict__ work,
               int *__restrict__ status, int *fail /* zero on entry (the controller clears it) */, const int *__restrict__ done)
{
    if (*done) return;
    extern __shared__ __align__(16) unsigned char cl_smem[];
    ClBuf *bufs = reinterpret_cast<ClBuf *>(cl_smem);
    double *ys = reinterpret_cast<double *>(bufs + kClGroups);
    // the warp index through a shuffle: the compiler then knows that everything derived from it (group, roles) is warp uniform,
    // and the shuffles inside the role branches stay plain SHFLs instead of divergence-safe collective sequences
    const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), grp = warp >> 2;
    const int gt = ((warp & 3) << 5) | (tid & 31);
    const int cta = blockIdx.x, gidx = cta * kClGroups + grp;
    constexpr int kWorkers = (kClCtas - 1) * kClGroups;          // CTA 0 keeps its load/store unit for the diagonal tile
    const int nblk = (N + 31) / 32;
    double *rhs = work, *rdiag = work + N + 32;                  // rdiag: nblk * 32 reciprocals of the factor's diagonal
    double *linv = rdiag + nblk * 32;                            // nblk inverses of the diagonal blocks of the factor
    ClBuf &T = bufs[grp];

    if (cta == 0) for (int i = tid; i < N; i += kClThreads) rhs[i] = E[i];
    CL_TRACE();
    // "panel" -1 has no columns: group 0 only fetches and factors the first diagonal tile (one call site for the unrolled
    // factorisation: the instruction cache is cold at every launch)
    for (int kbi = -1; kbi < nblk; ++kbi) {
        if (kbi >= 0 && *(volatile int *)fail) break;               // read after a cluster barrier: every CTA sees the same value
        const int kb = kbi * 32, R = nblk - kbi - 1;
        // ---- panel: the R row blocks below the diagonal block and the right-hand-side row.  Group 0 solves the first row block,
        //      goes straight on to the next diagonal tile with the rows it still holds (no barrier, no reload) and factors it
        //      while the other groups update the rest of the trailing matrix ----
        if (gidx == 0 && R > 0) {
            if (kbi >= 0) {
                const int w = gt >> 5, c = gt & 31, ib = kbi + 1, nd = cl_rows(N, nblk, ib);
                double old[8]; bool valid[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) {                       // the tile's current values: in flight during the panel solve
                    const int r = 8 * w + k;
                    valid[k] = r < nd && c <= r;
                    old[k] = __ldcg(S + (int64_t)(ib * 32) * N + ib * 32 + (valid[k] ? (int64_t)r * N + c : 0));
                }
                cl_panel_solve(S, rhs, rdiag, N, nblk, kb, ib, T, gt, grp, true);
                CL_TRACE();
                cl_cluster_arrive();
                cl_diag_from_panel(T, gt, grp, old, valid);
            } else {
                cl_tile_update(S, rhs, N, nblk, 0, 0, 0, T, gt, grp, false, true);
            }
            CL_TRACE();
            cl_factor_store(S, rdiag, N, kb + 32, T, gt, grp, fail);
            CL_TRACE();
            if (kbi < 0) cl_cluster_arrive();                       // (no trailing barrier in this round: the factor is published here)
        } else {
            if (kbi >= 0) {
                // units dealt CTA-major (unit u -> CTA u mod 8, group u div 8): up to eight units, one per SM
                for (int u = cta + kClCtas * grp; u <= R; u += kClCtas * kClGroups)
                    if (u > 0 || R == 0) cl_panel_solve(S, rhs, rdiag, N, nblk, kb, kbi + 1 + u, T, gt, grp, false);
                if (gidx == kClCtas * kClGroups - 1) cl_invert_diag(S, rdiag, linv + (size_t)kbi * 1024, N, kb, T, gt, grp);
            }
            cl_cluster_arrive();
        }
        cl_cluster_wait();
        CL_TRACE();
        if (kbi < 0) continue;
        // ---- trailing update: tiles (ti, tj), tj <= ti < R, and (right-hand side, tj); tile 0 (the next diagonal block) is done ----
        const int ntile = R * (R + 1) / 2 + R;
        if (cta > 0) {
            for (int tix = 1 + gidx - kClGroups; tix < ntile; tix += kWorkers) {
                int ti = 0, rem = tix;                              // row-major lower triangle; the last row (ti == R) is the right-hand side
                while (rem > ti) { rem -= ti + 1; ++ti; }
                cl_tile_update(S, rhs, N, nblk, kb, kbi + 1 + ti, kbi + 1 + rem, T, gt, grp, true, false);
            }
        }
        cl_cluster_bar();
        CL_TRACE();
    }
    if (*(volatile int *)fail) { if (cta == 0 && tid == 0) status[0] = 0; return; }
    if (cta != 0) return;

    // ---- back-substitution (CTA 0): L^T x = y, block by block from the bottom; y = the right-hand-side row.  x_b = L_bb^-T y_b is
    //      a product with the stored inverse (no 32-step chain); the rows of L that take x_b to the blocks above do not depend on
    //      it: their loads are issued first ----
    for (int i = tid; i < N + 32; i += kClThreads) ys[i] = (i < N) ? __ldcg(rhs + i) : 0.0;      // (the tail: a partial last block multiplies it by zero)
    double *D = bufs[0].B, *xs = bufs[1].A;                         // L_bb^-1 (stride kClLdB), x_b
    double v0 = __ldcg(linv + (size_t)(nblk - 1) * 1024 + tid), v1 = __ldcg(linv + (size_t)(nblk - 1) * 1024 + tid + kClThreads);
    for (int b = nblk - 1; b >= 0; --b) {
        const int kb = b * 32, nb = min(32, N - kb);
        D[(tid >> 5) * kClLdB + (tid & 31)] = v0;                   // (the previous block's product ended before the last barrier)
        D[((tid + kClThreads) >> 5) * kClLdB + (tid & 31)] = v1;
        double lv[32];
        if (tid < kb) {
            const double *p = S + (int64_t)kb * N + tid;
#pragma unroll
            for (int q = 0; q < 32; ++q) lv[q] = __ldcg(p + (int64_t)min(q, nb - 1) * N);      // rows past nb: x_b is zero there
        }
        if (b > 0) { v0 = __ldcg(linv + (size_t)(b - 1) * 1024 + tid); v1 = __ldcg(linv + (size_t)(b - 1) * 1024 + tid + kClThreads); }
        __syncthreads();
        if (warp == 0) {                                            // x_c = sum_{q >= c} Linv[q][c] y_q
            double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll
            for (int q = 0; q < 32; q += 4) {
                s0 = fma(D[q * kClLdB + tid], ys[kb + q], s0); s1 = fma(D[(q + 1) * kClLdB + tid], ys[kb + q + 1], s1);
                s2 = fma(D[(q + 2) * kClLdB + tid], ys[kb + q + 2], s2); s3 = fma(D[(q + 3) * kClLdB + tid], ys[kb + q + 3], s3);
            }
            xs[tid] = (tid < nb) ? (s0 + s1) + (s2 + s3) : 0.0;
        }
        __syncthreads();
        if (tid < nb) ys[kb + tid] = xs[tid];
        if (tid < kb) {
            double s0 = ys[tid], s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll
            for (int q = 0; q < 32; q += 4) {
                s0 = fma(-lv[q], xs[q], s0); s1 = fma(-lv[q + 1], xs[q + 1], s1);
                s2 = fma(-lv[q + 2], xs[q + 2], s2); s3 = fma(-lv[q + 3], xs[q + 3], s3);
            }
            ys[tid] = (s0 + s1) + (s2 + s3);
        }
    }
    __syncthreads();
    for (int i = tid; i < N; i += kClThreads) x[i] = ys[i];
    if (tid == 0) status[0] = 1;
    CL_TRACE();
}

}  // namespace argus
