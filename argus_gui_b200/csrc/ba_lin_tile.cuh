// Generation-2 linearisation: one CTA per tile of points (<= 32 points, <= TO observations), one thread per observation.
//
// Fuses, per LM attempt, everything the reference does between the Jacobian evaluation and the reduced system's right
// hand side (sba_levmar.c:795-1016 and the E_j part of :1164-1185), without ever materialising the Jacobian or W:
//   P1  per observation: project, Jacobian A (2x16), B (2x3), residual e              (img_projsKDRTS_jac_x)
//   P2  per point (8 threads, shuffle reduction): V_i, eb_i, damping, V*_i = L D L^T, V*^-1, c_i = V*^-1 eb_i
//   P3  per observation: e' = e - B c_i (so that E_j = sum A^T e'), W = A^T B, Z = W L^-T |D|^-1/2  -> HBM (384 B / obs)
//   P4  per camera (one warp each, DMMA m8n8k4): Gram matrix of G = [A^T ; e^T ; e'^T] over the camera's observations
//       in the tile: U_j (16x16), ea_j = col 16, E_j = col 17.  Deterministic: fixed order inside a tile, per-CTA
//       shared-memory accumulators with a single owner per element, fixed-order reduction over CTAs afterwards.
// mode 0 (first iteration, damping not known yet) does P1, P2 (V, eb only) and P4 (U, ea): the statistics for
// mu = tau * max diag (sba_levmar.c:980-987).
#pragma once
#include "ba_common.cuh"
#include "ba_schur_mma.cuh"

namespace argus {

constexpr int kGa = 38;          // doubles per observation in the Gram operand tile: 32 (A rows, (g, g+8) pairs) + 4 (e, e') + 2 pad:
                                 // 19 16-byte units per row (odd) -> consecutive threads hit distinct shared-memory banks
constexpr int kBs = 10;          // doubles per observation in the B/e tile: 6 + 2 + 2 pad (5 16-byte units, odd)
constexpr int kUfrag = 320;      // per camera: 5 DMMA tiles x 64 fragment elements (lane * 10 + q)
constexpr int kPs = 16;          // per point scratch: li[3], sq[3], c[3], sign, free

struct LinTilePlan {
    const uint8_t *corder;       // nobs: per tile, local observation indices sorted by camera
    const uint16_t *cstart;      // ntiles x (m+1): start of each camera's run in corder (local)
    const uint32_t *row_info;    // Z rows: what the row of a tile's Z block holds - local observation index | camera << 8 | local point << 16,
                                 // 0xFFFFFFFF = unused row
    const double2 *row_uv;       // Z rows: the measured pixel of that observation (a row-ordered copy of obs_uv: the tile kernel's
                                 // per-row inputs are then two independent, coalesced loads)
};

struct LinTileOut {
    double *Z;                   // Z rows x 48 (permuted, see zperm; chunk-major tile blocks, z_index)
    uint8_t *zs;                 // ntiles x 256: sign bits of D of the point behind every Z row (what the Schur kernel reads)
    uint8_t *zsign;              // n
    double *Vinv;                // n x 6
    double *V;                   // n x 6 (undamped, for the parity hook)
    double *eb;                  // n x 3
    double *vstate;              // n x 3 (sba-1.6 compatibility state)
    double *Upart;               // grid x m x 320
    double *scpart;              // grid x 4: cost, max|eb|, max diag V, sum pts^2
    int *flag;
};

// ---- static index construction (once per upload; visibility never changes during a solve) ----
// The visibility mask arrives as the reference's n x m byte matrix (sba_levmar.c:639-650 turns it into a CRS on the host);
// here the CSR (obs_ptr, obs_cam) and the per-point camera bit masks are built on the device from the raw bytes:
//   k_index_count : per point bit mask + per-block observation totals   (kIdxPts points per block)
//   k_index_scan  : exclusive scan of the block totals (one CTA)
//   k_index_fill  : obs_ptr of every point, obs_cam of every observation
constexpr int kIdxPer = 8;                    // points per thread
constexpr int kIdxPts = 256 * kIdxPer;        // points per block

__global__ void __launch_bounds__(256) k_index_count(int64_t n, int m, const uint8_t *__restrict__ vmask, uint64_t *__restrict__ pt_mask,
                                                     int32_t *__restrict__ btot)
{
    __shared__ int s_w[8];
    const int64_t base = (int64_t)blockIdx.x * kIdxPts + (int64_t)threadIdx.x * kIdxPer;
    int cnt = 0;
    for (int q = 0; q < kIdxPer; ++q) {
        const int64_t i = base + q;
        if (i >= n) break;
        const uint8_t *row = vmask + (size_t)i * m;
        uint64_t mk = 0;
        for (int j = 0; j < m; ++j) if (row[j]) mk |= 1ull << j;
        pt_mask[i] = mk;
        cnt += __popcll(mk);
    }
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_down_sync(0xffffffffu, cnt, o);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) { int t = 0; for (int w = 0; w < 8; ++w) t += s_w[w]; btot[blockIdx.x] = t; }
}

__global__ void __launch_bounds__(1024) k_index_scan(int nb, int32_t *__restrict__ btot /* in: totals, out: exclusive offsets; [nb] = grand total */)
{
    __shared__ int s_w[32];
    __shared__ int s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int b0 = 0; b0 < nb; b0 += 1024) {
        const int b = b0 + threadIdx.x;
        const int v = (b < nb) ? btot[b] : 0;
        int inc = v;
        for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, inc, o); if ((threadIdx.x & 31) >= o) inc += u; }
        if ((threadIdx.x & 31) == 31) s_w[threadIdx.x >> 5] = inc;
        __syncthreads();
        if (threadIdx.x < 32) {
            int w = s_w[threadIdx.x];
            for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, w, o); if (threadIdx.x >= o) w += u; }
            s_w[threadIdx.x] = w;
        }
        __syncthreads();
        const int wofs = (threadIdx.x >> 5) ? s_w[(threadIdx.x >> 5) - 1] : 0;
        const int carry = s_carry;
        if (b < nb) btot[b] = carry + wofs + inc - v;
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = carry + wofs + inc;
        __syncthreads();
    }
    if (threadIdx.x == 0) btot[nb] = s_carry;
}

__global__ void __launch_bounds__(256) k_index_fill(int64_t n, const uint64_t *__restrict__ pt_mask, const int32_t *__restrict__ btot,
                                                    int32_t *__restrict__ obs_ptr, uint8_t *__restrict__ obs_cam)
{
    __shared__ int s_w[8];
    const int64_t base = (int64_t)blockIdx.x * kIdxPts + (int64_t)threadIdx.x * kIdxPer;
    uint64_t mk[kIdxPer];
    int cnt = 0;
#pragma unroll
    for (int q = 0; q < kIdxPer; ++q) { mk[q] = (base + q < n) ? pt_mask[base + q] : 0ull; cnt += __popcll(mk[q]); }
    int inc = cnt;
    for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, inc, o); if ((threadIdx.x & 31) >= o) inc += u; }
    if ((threadIdx.x & 31) == 31) s_w[threadIdx.x >> 5] = inc;
    __syncthreads();
    int k = btot[blockIdx.x] + inc - cnt;
    for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) k += s_w[w];
#pragma unroll
    for (int q = 0; q < kIdxPer; ++q) {
        if (base + q >= n) break;
        obs_ptr[base + q] = k;
        uint64_t b = mk[q];
        while (b) { obs_cam[k++] = (uint8_t)(__ffsll((long long)b) - 1); b &= b - 1; }
    }
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 255) obs_ptr[n] = k;     // last thread of the last block ends at the grand total
}

// per tile: camera-sorted (stable) observation order, the camera run starts, and what every row of the tile's Z block holds: the
// observation of rank r of a point sits in row pt_slot0[point] + 4 r (ba_schur_mma.cuh: four classes of points, rows 4 i + class);
// row_info is pre-filled with 0xFF bytes (unused row)
__global__ void __launch_bounds__(256) k_build_tile_index(int m, const int32_t *__restrict__ tile_pt, const int32_t *__restrict__ tile_o,
                                                          const int32_t *__restrict__ tile_r, const int32_t *__restrict__ obs_ptr,
                                                          const uint8_t *__restrict__ obs_cam, const double2 *__restrict__ obs_uv,
                                                          const uint8_t *__restrict__ pt_slot0, uint8_t *__restrict__ corder,
                                                          uint16_t *__restrict__ cstart, uint32_t *__restrict__ row_info, double2 *__restrict__ row_uv)
{
    __shared__ int s_pofs[kTilePts + 1];
    __shared__ uint8_t s_cam[256];
    __shared__ int s_cnt[ARGUS_BA_MAX_CAMS + 1];
    const int tile = blockIdx.x, tid = threadIdx.x;
    const int p0 = tile_pt[tile], npts = tile_pt[tile + 1] - p0;
    const int o0 = tile_o[tile], nobs_t = tile_o[tile + 1] - o0;
    if (tid <= npts) s_pofs[tid] = obs_ptr[p0 + tid] - o0;
    if (tid <= m) s_cnt[tid] = 0;
    if (tid < nobs_t) s_cam[tid] = obs_cam[o0 + tid];
    __syncthreads();
    if (tid < nobs_t) {
        int lp = 0;
        while (lp + 1 < npts && tid >= s_pofs[lp + 1]) ++lp;
        const size_t row = (size_t)tile_r[tile] + pt_slot0[p0 + lp] + 4 * (tid - s_pofs[lp]);
        row_info[row] = (uint32_t)tid | ((uint32_t)s_cam[tid] << 8) | ((uint32_t)lp << 16);
        row_uv[row] = obs_uv[o0 + tid];
        atomicAdd(&s_cnt[s_cam[tid] + 1], 1);               // integer histogram: order-independent
    }
    __syncthreads();
    if (tid == 0) for (int j = 0; j < m; ++j) s_cnt[j + 1] += s_cnt[j];
    __syncthreads();
    if (tid <= m) cstart[(size_t)tile * (m + 1) + tid] = (uint16_t)s_cnt[tid];
    if (tid < nobs_t) {                                      // stable position inside the camera's run: observations before me with my camera
        const int cam = s_cam[tid];
        int before = 0;
        for (int q = 0; q < tid; ++q) before += (s_cam[q] == cam);
        corder[o0 + s_cnt[cam] + before] = (uint8_t)tid;
    }
}

__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// 32-byte global store (sm_100: st.global.v4.f64): one full sector per store instruction
__device__ __forceinline__ void st_global_v4(double *p, double a, double b, double c, double d)
{
    asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
}

#ifdef ARGUS_LIN_TRACE
// debug builds only (tools/lin_trace.py): clock64 stamps per (CTA < 2, tile of the CTA < 64, warp, phase boundary)
__device__ long long *g_lin_trace = nullptr;
#define LIN_TRACE(slot_) do { if (g_lin_trace && blockIdx.x < 2 && lane == 0 && ltile < 64) \
    g_lin_trace[(((size_t)blockIdx.x * 64 + ltile) * 8 + warp) * 10 + (slot_)] = clock64(); } while (0)
#else
#define LIN_TRACE(slot_) do { } while (0)
#endif
template <int MC>
__global__ void __launch_bounds__(256, (MC <= 4) ? 2 : 1) k_lin_tile(BaProblem pb, TilePlan tp, LinTilePlan lp, ParamBufs pbuf,
                                                  const LmState *__restrict__ st, int mode, LinTileOut out)
{
    if (st->done) return;
    const double *__restrict__ cam = st->cur_cam ? pbuf.cam[1] : pbuf.cam[0];      // (a run-time index would put the pointer table in local memory)
    const double *__restrict__ pts = st->cur_pts ? pbuf.pts[1] : pbuf.pts[0];
    const double mu = st->mu;
    const int use_state = st->use_state, write_state = st->write_state;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    CamConst *sc = reinterpret_cast<CamConst *>(smem_raw);
    double *Ga = reinterpret_cast<double *>(smem_raw + (((size_t)pb.m * sizeof(CamConst) + 15) & ~(size_t)15));   // TO x 36 (16-byte aligned)
    double *Bs = Ga + (size_t)kTileRows * kGa;                      // 256 x kBs
    double *Ps = Bs + (size_t)kTileRows * kBs;                      // 32 x 16
    __shared__ double red[32];
    __shared__ double s_M[kTilePts * 3];          // the tile's points
    __shared__ int s_pofs[kTilePts + 1];          // local observation offset of each point
    __shared__ uint16_t s_cst[ARGUS_BA_MAX_CAMS + 2];   // start of each camera's run in s_cord
    __shared__ uint8_t s_cord[256];               // local observation indices sorted by camera

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g8 = lane >> 2, kk = lane & 3;
    load_cameras(sc, cam, pb.rot0, pb.m);
    // Gram accumulators of the cameras this warp owns (mcon + warp + 8 q), kept in registers across all tiles of the CTA
    double uacc[MC][10];
#pragma unroll
    for (int q = 0; q < MC; ++q)
#pragma unroll
        for (int t = 0; t < 10; ++t) uacc[q][t] = 0.0;

    double cost = 0.0, ebmax = 0.0, vdmax = -1.79769313486231570e308, ptsq = 0.0;

    // tile bounds are loaded one tile ahead (they end up in uniform registers), and the next tile's index data / measurements /
    // points are pulled into L2 while this tile computes: the dependent loads at the top of a tile then see L2, not DRAM, latency
    int nx_p0 = 0, nx_p1 = 0, nx_o0 = 0, nx_o1 = 0, nx_r0 = 0, nx_r1 = 0;
    if ((int)blockIdx.x < tp.ntiles) {
        nx_p0 = tp.tile_pt[blockIdx.x]; nx_p1 = tp.tile_pt[blockIdx.x + 1]; nx_o0 = tp.tile_o[blockIdx.x]; nx_o1 = tp.tile_o[blockIdx.x + 1];
        nx_r0 = tp.tile_r[blockIdx.x]; nx_r1 = tp.tile_r[blockIdx.x + 1];
    }
#ifdef ARGUS_LIN_TRACE
    int ltile = -1;
#endif
    for (int tile = blockIdx.x; tile < tp.ntiles; tile += gridDim.x) {
#ifdef ARGUS_LIN_TRACE
        ++ltile;
#endif
        LIN_TRACE(0);
        const int p0 = nx_p0, npts = nx_p1 - nx_p0;
        const int o0 = nx_o0, nobs_t = nx_o1 - nx_o0;
        const int r0 = nx_r0, nrows_t = nx_r1 - nx_r0;
        const int ntile = tile + gridDim.x;
        if (ntile < tp.ntiles) {
            nx_p0 = tp.tile_pt[ntile]; nx_p1 = tp.tile_pt[ntile + 1]; nx_o0 = tp.tile_o[ntile]; nx_o1 = tp.tile_o[ntile + 1];
            nx_r0 = tp.tile_r[ntile]; nx_r1 = tp.tile_r[ntile + 1];
        }
        // the tile's index data and points -> shared memory (the previous tile ended with a barrier).  In the per-observation
        // phases P1 and P3 a thread is a ROW of the tile's Z block (so that the Z stores of a warp are consecutive 32-byte
        // chunks); li is the observation that row holds - the index everything in shared memory is keyed by
        double2 z = make_double2(0.0, 0.0); uint32_t info = 0xFFFFFFFFu;
        if (tid < nrows_t) { info = lp.row_info[(size_t)r0 + tid]; z = lp.row_uv[(size_t)r0 + tid]; }
        const bool act = info != 0xFFFFFFFFu;
        const int li = (int)(info & 0xFFu), my_cam = (int)((info >> 8) & 0xFFu), my_lp = (int)((info >> 16) & 0xFFu);
        if (tid < nobs_t) s_cord[tid] = lp.corder[o0 + tid];
        if (tid <= npts) s_pofs[tid] = pb.obs_ptr[p0 + tid] - o0;
        if (tid < npts) { const int64_t i = p0 + tid; s_M[3 * tid] = pts[3 * i]; s_M[3 * tid + 1] = pts[3 * i + 1]; s_M[3 * tid + 2] = pts[3 * i + 2]; }
        if (tid <= pb.m) s_cst[tid] = lp.cstart[(size_t)tile * (pb.m + 1) + tid];
        __syncthreads();
        LIN_TRACE(1);

        // ---- P1: one thread per observation ----
        if (act) {
            const int64_t i = p0 + my_lp;
            const double M0 = s_M[3 * my_lp], M1 = s_M[3 * my_lp + 1], M2 = s_M[3 * my_lp + 2];
            double u, v, A[32], B[6];
            const int nfk = pb.cam_fix ? pb.cam_fix[2 * my_cam] : pb.nfix_k, nfd = pb.cam_fix ? pb.cam_fix[2 * my_cam + 1] : pb.nfix_d;
            cam_project_jac(sc[my_cam], M0, M1, M2, nfk, nfd, u, v, A, B);
            const double e0 = z.x - u, e1 = z.y - v;
            cost += e0 * e0 + e1 * e1;
            double *g = Ga + (size_t)li * kGa;
            if (my_cam >= pb.mcon) {
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    // row layout [r (8)][component u / v][parameter r, r + 8]: the four 16-byte elements a quarter-warp of the
                    // Gram phase reads from one observation (2 rows x 2 components) are 64 contiguous bytes
                    *reinterpret_cast<double2 *>(g + r * 4) = make_double2(A[r], A[r + 8]);
                    *reinterpret_cast<double2 *>(g + r * 4 + 2) = make_double2(A[16 + r], A[24 + r]);
                }
            } else {
#pragma unroll
                for (int r = 0; r < 16; ++r) *reinterpret_cast<double2 *>(g + r * 2) = make_double2(0.0, 0.0);
            }
            *reinterpret_cast<double2 *>(g + 32) = make_double2(e0, 0.0);      // [comp][{e, e'}]
            *reinterpret_cast<double2 *>(g + 34) = make_double2(e1, 0.0);
            double *b = Bs + (size_t)li * kBs;
            if (i >= pb.ncon) {
                *reinterpret_cast<double2 *>(b) = make_double2(B[0], B[1]); *reinterpret_cast<double2 *>(b + 2) = make_double2(B[2], B[3]);
                *reinterpret_cast<double2 *>(b + 4) = make_double2(B[4], B[5]);
            } else {
                *reinterpret_cast<double2 *>(b) = make_double2(0.0, 0.0); *reinterpret_cast<double2 *>(b + 2) = make_double2(0.0, 0.0);
                *reinterpret_cast<double2 *>(b + 4) = make_double2(0.0, 0.0);
            }
            *reinterpret_cast<double2 *>(b + 6) = make_double2(e0, e1);
        }
        if (ntile < tp.ntiles) {      // L2 prefetch of the next tile's inputs (one request per 128-byte line)
            const int no = nx_o1 - nx_o0, np = nx_p1 - nx_p0, nr = nx_r1 - nx_r0;
            if ((tid & 7) == 0 && tid < nr) prefetch_l2(lp.row_uv + nx_r0 + tid);
            if ((tid & 31) == 0 && tid < nr) prefetch_l2(lp.row_info + nx_r0 + tid);
            if (tid < 3 && tid * 128 < no + 127) prefetch_l2(lp.corder + nx_o0 + tid * 128);
            if (tid >= 32 && tid < 40 && (tid - 32) * 16 < 3 * np + 15) prefetch_l2(pts + 3 * (int64_t)nx_p0 + (tid - 32) * 16);
            if (tid == 64) { prefetch_l2(pb.obs_ptr + nx_p0); prefetch_l2(pb.obs_ptr + nx_p0 + 31); prefetch_l2(lp.cstart + (size_t)ntile * (pb.m + 1)); }
        }
        LIN_TRACE(2);
        __syncthreads();
        LIN_TRACE(3);

        // ---- P2: 8 threads per point ----
        {
            const int lpt = tid >> 3, q = tid & 7;
            double v00 = 0, v01 = 0, v02 = 0, v11 = 0, v12 = 0, v22 = 0, b0 = 0, b1 = 0, b2 = 0;
            const bool have = lpt < npts;
            const int64_t i = p0 + (have ? lpt : 0);
            if (have) {
                const int ob = s_pofs[lpt], oe = s_pofs[lpt + 1];
                for (int o = ob + q; o < oe; o += 8) {
                    const double *b = Bs + (size_t)o * kBs;
                    const double2 t0 = *reinterpret_cast<const double2 *>(b), t1 = *reinterpret_cast<const double2 *>(b + 2),
                                  t2 = *reinterpret_cast<const double2 *>(b + 4), t3 = *reinterpret_cast<const double2 *>(b + 6);
                    const double B0 = t0.x, B1 = t0.y, B2 = t1.x, B3 = t1.y, B4 = t2.x, B5 = t2.y, e0 = t3.x, e1 = t3.y;
                    v00 += B0 * B0 + B3 * B3; v01 += B0 * B1 + B3 * B4; v02 += B0 * B2 + B3 * B5;
                    v11 += B1 * B1 + B4 * B4; v12 += B1 * B2 + B4 * B5; v22 += B2 * B2 + B5 * B5;
                    b0 += B0 * e0 + B3 * e1; b1 += B1 * e0 + B4 * e1; b2 += B2 * e0 + B5 * e1;
                }
            }
#pragma unroll
            for (int off = 4; off > 0; off >>= 1) {       // fixed-order butterfly: every lane of the group ends with the same sum
                v00 += __shfl_xor_sync(0xffffffffu, v00, off); v01 += __shfl_xor_sync(0xffffffffu, v01, off);
                v02 += __shfl_xor_sync(0xffffffffu, v02, off); v11 += __shfl_xor_sync(0xffffffffu, v11, off);
                v12 += __shfl_xor_sync(0xffffffffu, v12, off); v22 += __shfl_xor_sync(0xffffffffu, v22, off);
                b0 += __shfl_xor_sync(0xffffffffu, b0, off); b1 += __shfl_xor_sync(0xffffffffu, b1, off);
                b2 += __shfl_xor_sync(0xffffffffu, b2, off);
            }
            if (have && q == 0) {
                const bool pfree = i >= pb.ncon;
                const double M0 = s_M[3 * lpt], M1 = s_M[3 * lpt + 1], M2 = s_M[3 * lpt + 2];
                ptsq += M0 * M0 + M1 * M1 + M2 * M2;
                double *Vi = out.V + 6 * i;
                Vi[0] = v00; Vi[1] = v01; Vi[2] = v02; Vi[3] = v11; Vi[4] = v12; Vi[5] = v22;
                out.eb[3 * i] = b0; out.eb[3 * i + 1] = b1; out.eb[3 * i + 2] = b2;
                double *P = Ps + lpt * kPs;
#pragma unroll
                for (int t = 0; t < 10; ++t) P[t] = 0.0;
                if (pfree) {
                    if (!(pb.wand_of && pb.wand_of[i] >= 0)) {       // points of a wand pair: k_wand_prepare adds the pair's share first
                        ebmax = fmax(ebmax, fmax(fabs(b0), fmax(fabs(b1), fabs(b2))));
                        vdmax = fmax(vdmax, fmax(v00, fmax(v11, v22)));
                    }
                    if (mode != 0) {
                        double vv[6] = {v00, v01, v02, v11, v12, v22};
                        if (use_state) { vv[0] = out.vstate[3 * i]; vv[3] = out.vstate[3 * i + 1]; vv[5] = out.vstate[3 * i + 2]; }
                        vv[0] += mu; vv[3] += mu; vv[5] += mu;
                        double li[3], di[3], rs[3];
                        if (!sym3_ldl_inv_rs(vv, li, di, rs)) {
                            atomicOr(out.flag, 1);
                            out.zsign[i] = 0;
                        } else {
                            const double n00 = di[0] + li[0] * li[0] * di[1] + li[1] * li[1] * di[2];
                            const double n01 = li[0] * di[1] + li[1] * li[2] * di[2];
                            const double n02 = li[1] * di[2];
                            const double n11 = di[1] + li[2] * li[2] * di[2];
                            const double n12 = li[2] * di[2];
                            const double n22 = di[2];
                            double *Ni = out.Vinv + 6 * i;
                            Ni[0] = n00; Ni[1] = n01; Ni[2] = n02; Ni[3] = n11; Ni[4] = n12; Ni[5] = n22;
                            if (write_state) { out.vstate[3 * i] = n00; out.vstate[3 * i + 1] = n11; out.vstate[3 * i + 2] = n22; }
                            P[0] = li[0]; P[1] = li[1]; P[2] = li[2];
                            P[3] = rs[0]; P[4] = rs[1]; P[5] = rs[2];
                            P[6] = n00 * b0 + n01 * b1 + n02 * b2; P[7] = n01 * b0 + n11 * b1 + n12 * b2; P[8] = n02 * b0 + n12 * b1 + n22 * b2;
                            const int sgn = (di[0] < 0.0 ? 1 : 0) | (di[1] < 0.0 ? 2 : 0) | (di[2] < 0.0 ? 4 : 0);
                            out.zsign[i] = (uint8_t)sgn;
                            P[9] = (double)sgn;
                        }
                    }
                } else if (mode != 0) {
                    out.zsign[i] = 0;
                }
            }
        }
        LIN_TRACE(4);
        __syncthreads();
        LIN_TRACE(5);

        // ---- P3: per observation: e' and Z ----
        if (mode != 0) out.zs[(size_t)tile * 256 + tid] = act ? (uint8_t)(int)Ps[my_lp * kPs + 9] : (uint8_t)0;
        if (mode != 0 && act) {
            const double *P = Ps + my_lp * kPs;
            const double *b = Bs + (size_t)li * kBs;
            const double2 t0 = *reinterpret_cast<const double2 *>(b), t1 = *reinterpret_cast<const double2 *>(b + 2),
                          t2 = *reinterpret_cast<const double2 *>(b + 4), t3 = *reinterpret_cast<const double2 *>(b + 6);
            const double B0 = t0.x, B1 = t0.y, B2 = t1.x, B3 = t1.y, B4 = t2.x, B5 = t2.y;
            double *g = Ga + (size_t)li * kGa;
            const double c0 = P[6], c1 = P[7], c2 = P[8];
            g[33] = t3.x - (B0 * c0 + B1 * c1 + B2 * c2);
            g[35] = t3.y - (B3 * c0 + B4 * c1 + B5 * c2);
            const double l0 = P[0], l1 = P[1], l2 = P[2], s0 = P[3], s1 = P[4], s2 = P[5];
            const size_t zpl = (size_t)(nrows_t + 1) * 4;                // doubles per chunk plane (one padding chunk, z_index)
            double *Zt = out.Z + 48 * ((int64_t)r0 + tile) + 4 * (int64_t)tid;     // chunk-major tile block: chunk c of this row at Zt + c * zpl
            // two rows (r, r+1) x three components x (lo, hi) = 12 doubles = three 32-byte stores (full sectors)
#pragma unroll
            for (int r = 0; r < 8; r += 2) {
                double zz[12];
#pragma unroll
                for (int rr = 0; rr < 2; ++rr) {
                    const double2 a0 = *reinterpret_cast<const double2 *>(g + (r + rr) * 4);          // du/da_r, du/da_{r+8}
                    const double2 a1 = *reinterpret_cast<const double2 *>(g + (r + rr) * 4 + 2);      // dv/da_r, dv/da_{r+8}
                    const double wl0 = a0.x * B0 + a1.x * B3, wl1 = a0.x * B1 + a1.x * B4, wl2 = a0.x * B2 + a1.x * B5;   // row r
                    const double wh0 = a0.y * B0 + a1.y * B3, wh1 = a0.y * B1 + a1.y * B4, wh2 = a0.y * B2 + a1.y * B5;   // row r + 8
                    // layout ((r%8)*3 + c)*2 + r/8
                    zz[rr * 6 + 0] = wl0 * s0; zz[rr * 6 + 1] = wh0 * s0;
                    zz[rr * 6 + 2] = (l0 * wl0 + wl1) * s1; zz[rr * 6 + 3] = (l0 * wh0 + wh1) * s1;
                    zz[rr * 6 + 4] = (l1 * wl0 + l2 * wl1 + wl2) * s2; zz[rr * 6 + 5] = (l1 * wh0 + l2 * wh1 + wh2) * s2;
                }
                double *zp = Zt + (size_t)((r >> 1) * 3) * zpl;       // chunks 3 r/2 .. 3 r/2 + 2: consecutive lanes write consecutive chunks
                st_global_v4(zp, zz[0], zz[1], zz[2], zz[3]);
                st_global_v4(zp + zpl, zz[4], zz[5], zz[6], zz[7]);
                st_global_v4(zp + 2 * zpl, zz[8], zz[9], zz[10], zz[11]);
            }
        }
        LIN_TRACE(6);
        __syncthreads();
        LIN_TRACE(7);

        // ---- P4: Gram per camera on the tensor pipe (the cameras a warp owns advance together: independent DMMA chains) ----
        {
            int c0q[MC], njq[MC], smax = 0;
#pragma unroll
            for (int q = 0; q < MC; ++q) {
                const int j = pb.mcon + warp + 8 * q;
                c0q[q] = 0; njq[q] = 0;
                if (j < pb.m) { c0q[q] = s_cst[j]; njq[q] = s_cst[j + 1] - c0q[q]; }
                smax = max(smax, njq[q]);
            }
            const int comp = kk & 1;
            int onx[MC];                                   // observation index of the NEXT step (one shared-memory latency off the chain)
#pragma unroll
            for (int q = 0; q < MC; ++q) onx[q] = s_cord[c0q[q] + (((kk >> 1) < njq[q]) ? (kk >> 1) : 0)];
            for (int s = 0; s < smax; s += 2) {
#pragma unroll
                for (int q = 0; q < MC; ++q) {
                    if (s < njq[q]) {
                        const int idx = s + (kk >> 1);
                        const bool valid = idx < njq[q];
                        const int o = onx[q];
                        onx[q] = s_cord[c0q[q] + ((idx + 2 < njq[q]) ? idx + 2 : 0)];
                        const double *g = Ga + (size_t)o * kGa;
                        double2 a = *reinterpret_cast<const double2 *>(g + g8 * 4 + comp * 2);
                        double ax = (g8 < 2) ? g[32 + comp * 2 + g8] : 0.0;
                        if (!valid) { a.x = 0.0; a.y = 0.0; ax = 0.0; }
                        dmma884(uacc[q][0], uacc[q][1], a.x, a.x);
                        dmma884(uacc[q][2], uacc[q][3], a.x, a.y);
                        dmma884(uacc[q][4], uacc[q][5], a.y, a.y);
                        dmma884(uacc[q][6], uacc[q][7], a.x, ax);
                        dmma884(uacc[q][8], uacc[q][9], a.y, ax);
                    }
                }
            }
        }
        LIN_TRACE(8);
        __syncthreads();
        LIN_TRACE(9);
    }

    {   // per-CTA Gram fragments; cameras nobody owns (fixed ones) are written as zeros
        double *up = out.Upart + (size_t)blockIdx.x * pb.m * kUfrag;
        for (int t = tid; t < pb.mcon * kUfrag; t += blockDim.x) up[t] = 0.0;
#pragma unroll
        for (int q = 0; q < MC; ++q) {
            const int j = pb.mcon + warp + 8 * q;
            if (j < pb.m) {
#pragma unroll
                for (int t = 0; t < 10; ++t) up[(size_t)j * kUfrag + lane * 10 + t] = uacc[q][t];
            }
        }
    }
    const double s0 = block_sum(cost, red), s1 = block_max(ebmax, red), s2 = block_max(vdmax, red), s3 = block_sum(ptsq, red);
    if (tid == 0) {
        out.scpart[4 * blockIdx.x + 0] = s0; out.scpart[4 * blockIdx.x + 1] = s1;
        out.scpart[4 * blockIdx.x + 2] = s2; out.scpart[4 * blockIdx.x + 3] = s3;
    }
}

// Fixed-order reduction of the per-CTA Gram fragments into the packed per-camera layout the rest of the solver uses:
//   Ured[j][0..135] upper triangle of U_j, [136..151] ea_j;  Eout[j][0..15] = E_j (= sum A^T e').
// Also reduces the nparts_sc scalar partials: sum2 = {cost, sum pts^2}, max2 = {max|eb|, max diag V}.  grid = m blocks of 1024.
constexpr int kLinRedSlices = 6;      // 6 x 168 = 1008 threads: every output element is summed by 6 threads over interleaved partials
__global__ void __launch_bounds__(1024) k_lin_tile_reduce(int m, int nparts, int nparts_sc, const double *__restrict__ Upart, const double *__restrict__ scpart,
                                                          double *__restrict__ Ured, double *__restrict__ Eout, double *__restrict__ sum2,
                                                          double *__restrict__ max2, const int *__restrict__ done)
{
    __shared__ double s_part[kLinRedSlices][168];
    const int j = blockIdx.x, t = threadIdx.x % 168, slice = threadIdx.x / 168;
    if (*done) return;
    if (slice < kLinRedSlices) {
        // element -> (tile, row, col) of the 24-column Gram matrix
        int r, c;
        if (t < 136) { r = 0; int q = t; while (q >= 16 - r) { q -= 16 - r; ++r; } c = r + q; }
        else if (t < 152) { r = t - 136; c = 16; }
        else { r = t - 152; c = 17; }
        int tl, rr = r & 7, cc;
        if (c < 8) { tl = 0; cc = c; }                    // (0,0)
        else if (c < 16) { tl = (r < 8) ? 1 : 2; cc = c - 8; }   // (0,1) or (1,1)
        else { tl = (r < 8) ? 3 : 4; cc = c - 16; }       // (0,2) or (1,2)
        const int ln = (rr << 2) | (cc >> 1);
        const int idx = ln * 10 + tl * 2 + (cc & 1);
        double s = 0.0;
        const double *p = Upart + (size_t)j * kUfrag + idx;
        for (int q = slice; q < nparts; q += kLinRedSlices) s += p[(size_t)q * m * kUfrag];       // fixed order inside a slice
        s_part[slice][t] = s;
    }
    __syncthreads();
    if (threadIdx.x < 168) {
        double s = 0.0;
#pragma unroll
        for (int q = 0; q < kLinRedSlices; ++q) s += s_part[q][t];                                // and over the slices
        if (t < 152) Ured[j * kU + t] = s; else Eout[j * 16 + (t - 152)] = s;
    }
    if (j == 0 && threadIdx.x >= 992) {                  // the last warp: scalar partials, lanes interleaved over the parts, fixed-order tree
        const int lane = threadIdx.x - 992;
        double c0 = 0.0, c3 = 0.0, m1 = 0.0, m2 = -1.79769313486231570e308;
        for (int q = lane; q < nparts_sc; q += 32) {        // nparts_sc >= nparts: the wand kernel's partials follow the tile kernel's
            c0 += scpart[4 * q]; m1 = fmax(m1, scpart[4 * q + 1]); m2 = fmax(m2, scpart[4 * q + 2]); c3 += scpart[4 * q + 3];
        }
        for (int o = 16; o > 0; o >>= 1) {
            c0 += __shfl_xor_sync(0xffffffffu, c0, o); c3 += __shfl_xor_sync(0xffffffffu, c3, o);
            m1 = fmax(m1, __shfl_xor_sync(0xffffffffu, m1, o)); m2 = fmax(m2, __shfl_xor_sync(0xffffffffu, m2, o));
        }
        if (lane == 0) { sum2[0] = c0; sum2[1] = c3; max2[0] = m1; max2[1] = m2; }
    }
}

// Generation-2 back-substitution + trial evaluation, one thread per point, nothing per-observation is read back from
// HBM except the measurement: the Jacobian is re-evaluated (about 6x cheaper than streaming 384 B of W per observation).
//   db_i = V*_i^-1 (eb_i - sum_j B_ij^T (A_ij da_j))      (sba_levmar.c:1215-1256, W_ij^T da_j = B_ij^T A_ij da_j)
//   trial point, cost at (cam_new, pts_new)                 (sba_levmar.c:1261-1288)
//   cam_moves = 0: the structure-only driver (trial cameras = current cameras, da = 0)
__global__ void __launch_bounds__(128) k_backsub_recompute(BaProblem pb, ParamBufs pbuf, const LmState *__restrict__ st, int cam_moves,
                                                           const double *__restrict__ da, const double *__restrict__ Vinv,
                                                           const double *__restrict__ eb, double *__restrict__ dbout,
                                                           double *__restrict__ part)
{
    if (st->done) return;
    const double mu = st->mu;
    const int ccur = st->cur_cam, pcur = st->cur_pts;
    const double *__restrict__ cam = ccur ? pbuf.cam[1] : pbuf.cam[0];
    const double *__restrict__ cam_new = (ccur ^ (cam_moves ? 1 : 0)) ? pbuf.cam[1] : pbuf.cam[0];
    const double *__restrict__ pts = pcur ? pbuf.pts[1] : pbuf.pts[0];
    double *__restrict__ pts_new = pcur ? pbuf.pts[0] : pbuf.pts[1];
    extern __shared__ __align__(16) unsigned char smem_raw[];
    CamConst *sc = reinterpret_cast<CamConst *>(smem_raw);          // old cameras
    CamConst *sn = sc + pb.m;                                        // trial cameras
    double *das = reinterpret_cast<double *>(sn + pb.m);             // m rows of 17 (16 + 1 pad: lanes with different cameras hit different banks)
    __shared__ double red[32];
    for (int t = threadIdx.x; t < 16 * pb.m; t += blockDim.x) das[(t >> 4) * 17 + (t & 15)] = da[t];
    for (int j = threadIdx.x; j < pb.m; j += blockDim.x) { cam_precompute(cam + 16 * j, pb.rot0 + 4 * j, sc[j]); cam_precompute(cam_new + 16 * j, pb.rot0 + 4 * j, sn[j]); }
    __syncthreads();
    double cost = 0.0, dbsq = 0.0, dl = 0.0;
    for (int64_t base = (int64_t)blockIdx.x * blockDim.x; base < pb.n; base += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = base + threadIdx.x;
        if (i >= pb.n) continue;
        if (pb.wand_of && pb.wand_of[i] >= 0) continue;             // the two points of a wand pair are coupled: k_wand_backsub
        const double M0 = pts[3 * i], M1 = pts[3 * i + 1], M2 = pts[3 * i + 2];
        const int k0 = pb.obs_ptr[i], k1 = pb.obs_ptr[i + 1];
        double d0 = 0.0, d1 = 0.0, d2 = 0.0;
        if (i >= pb.ncon) {
            const double b0 = eb[3 * i], b1 = eb[3 * i + 1], b2 = eb[3 * i + 2];
            double t0 = b0, t1 = b1, t2 = b2;
            for (int k = k0; k < k1; ++k) {
                const int j = pb.obs_cam[k];
                if (j < pb.mcon) continue;
                double B[6], s0, s1;
                const int nfk = pb.cam_fix ? pb.cam_fix[2 * j] : pb.nfix_k, nfd = pb.cam_fix ? pb.cam_fix[2 * j + 1] : pb.nfix_d;
                cam_step_dir(sc[j], M0, M1, M2, nfk, nfd, das + 17 * j, s0, s1, B);        // (A_ij da_j, B_ij) without forming A_ij
                t0 -= B[0] * s0 + B[3] * s1; t1 -= B[1] * s0 + B[4] * s1; t2 -= B[2] * s0 + B[5] * s1;
            }
            const double *Ni = Vinv + 6 * i;
            d0 = Ni[0] * t0 + Ni[1] * t1 + Ni[2] * t2;
            d1 = Ni[1] * t0 + Ni[3] * t1 + Ni[4] * t2;
            d2 = Ni[2] * t0 + Ni[4] * t1 + Ni[5] * t2;
            dbsq += d0 * d0 + d1 * d1 + d2 * d2;
            dl += d0 * (mu * d0 + b0) + d1 * (mu * d1 + b1) + d2 * (mu * d2 + b2);
        }
        if (dbout) { dbout[3 * i] = d0; dbout[3 * i + 1] = d1; dbout[3 * i + 2] = d2; }
        const double N0 = M0 + d0, N1 = M1 + d1, N2 = M2 + d2;
        pts_new[3 * i] = N0; pts_new[3 * i + 1] = N1; pts_new[3 * i + 2] = N2;
        for (int k = k0; k < k1; ++k) {
            double u, v;
            cam_project(sn[pb.obs_cam[k]], N0, N1, N2, u, v);
            const double2 z = pb.obs_uv[k];
            const double e0 = z.x - u, e1 = z.y - v;
            cost += e0 * e0 + e1 * e1;
        }
    }
    const double s0 = block_sum(cost, red), s1 = block_sum(dbsq, red), s2 = block_sum(dl, red);
    if (threadIdx.x == 0) { part[3 * blockIdx.x] = s0; part[3 * blockIdx.x + 1] = s1; part[3 * blockIdx.x + 2] = s2; }
}

}  // namespace argus
