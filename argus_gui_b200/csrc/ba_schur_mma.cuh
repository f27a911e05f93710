// Generation-2 Schur-complement kernel: fp64 tensor-pipe (DMMA m8n8k4) rank-3 updates with S-blocks stationary in
// registers, operands streamed through shared memory by bulk-async (TMA) copies.
//
// Math.  With V*_i = L_i D_i L_i^T (3x3, per point) and  Z_ij = W_ij L_i^-T |D_i|^-1/2  (16x3, per observation),
//     sum_i Y_ij W_ik^T = sum_i Z_ij diag(sign D_i) Z_ik^T            (sba_levmar.c:1075-1114 computes the left side)
// so the whole Schur complement is a sum of signed rank-3 outer products of ONE operand array Z (384 B per
// observation) instead of two (Y and W).  For a damped, positive-definite V* all signs are +.
//
// Work decomposition.  The upper-triangular 16x16 camera blocks (j <= k) are dealt to G groups x 8 warps x NB slots.
// A CTA = (chunk of point tiles, group); each warp keeps its NB blocks as DMMA accumulators (8 doubles per lane per
// block) for the whole chunk.  Per tile (<= 32 points, <= TO observations) the CTA's Z block arrives by cp.async.bulk,
// one copy per chunk plane (z_index), into a double-buffered shared-memory stage; per block the warp ballots the points that
// see both cameras, compacts them into a hit list and feeds the K dimension (3 per hit, 4 per DMMA step) without padding
// (groups of 4 hits fill 3 steps; a last partial group runs through the same steps with its missing slots on a zero row).
// Fragment layout of mma.m8n8k4.f64: A: lane -> (row = lane>>2, k = lane&3); B: (k = lane&3, col = lane>>2);
// C: (row = lane>>2, cols 2*(lane&3), +1).  The 48 values of an observation are ordered ((r%8)*3 + c)*2 + r/8 so one
// 16-byte load gives a lane the (row, row+8) pair of its fragment - the same load serves the A and the B operand;
// consecutive 32-byte chunks of that order live in consecutive chunk planes of the tile (coalesced producer stores).
#pragma once
#include "ba_common.cuh"

namespace argus {

constexpr int kTilePts = 32;        // points per tile (one ballot)
constexpr int kSchurWarps = 8;

__host__ __device__ __forceinline__ int zperm(int r, int c) { return ((r & 7) * 3 + c) * 2 + (r >> 3); }

struct TilePlan {
    int ntiles;
    int tile_obs;                   // TO: max observations per tile
    const int32_t *tile_pt;         // ntiles + 1: first point of each tile
    const int32_t *tile_o;          // ntiles + 1: first observation of each tile
};

struct SchurPlan {
    int G;                          // block groups
    int nchunk;                     // point-tile chunks (grid.x)
    int npairs;                     // upper-triangular camera pairs among the free cameras
    const uint16_t *blk_jk;         // [G][8][NB]: j | k << 8 (absolute camera ids), 0xFFFF = empty slot
    const int32_t *blk_pair;        // [G][8][NB]: pair index (row-major upper triangle over free cameras) or -1
};

// ---- PTX helpers: mbarrier + bulk async copy (TMA, non-tensor form) ----
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t done = 0;
    const uint32_t a = smem_u32(bar);
    while (!done) {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
                     : "=r"(done) : "r"(a), "r"(parity) : "memory");
    }
}
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// Z in HBM: per tile of points one block of 48 * nobs_t doubles, chunk-major inside the block: the 12 four-double chunks of an
// observation's 16 x 3 row (chunk = zperm / 4) live in 12 planes of nobs_t chunks each, so that the one-thread-per-observation
// producer writes consecutive 32-byte chunks (coalesced) and the Schur kernel copies a tile plane by plane.
//   element e (= zperm(r, c)) of local observation lk of a tile starting at observation o0:
__host__ __device__ __forceinline__ int64_t z_index(int64_t o0, int nobs_t, int lk, int e)
{
    return 48 * o0 + ((int64_t)(e >> 2) * nobs_t + lk) * 4 + (e & 3);
}

// The Schur kernel.  grid = (nchunk, G); NW consumer warps, each owning NB blocks; the copy producer is lane 0 of warp 0
// (PROD = 0) or a dedicated extra warp (PROD = 1).
// Dynamic shared memory: 2 x [Z tile: 12 chunk planes of (TO + 1) x 32 bytes | rank tile: mp x 32 bytes], then 384 B of zeros.
// Spart[chunk][pair][256]: per-chunk partial sums in fragment order (lane * 8 + q), q = (ri*2 + ci)*2 + e for the
// 8x8 sub-tile (ri, ci) of the block: element (row = ri*8 + lane>>2, col = ci*8 + 2*(lane&3) + e).
//
// pt_rank[tile][mp][32] (uint8, static): position of camera j in point i's observation list, 0xFF when i does not see j (or i is
// a fixed point).  It travels with the Z tile, so finding the points that see both cameras of a block is two byte loads.
// K mapping: one hit (K = 3, padded to the DMMA's 4) per step; the k-lanes 0..2 of a step read the three components of the
// same observation at consecutive 16-byte addresses, which is free of shared-memory bank conflicts (a 4-hits-per-3-steps
// packing was measured slower: its four k-lanes read different observations at the same in-row offset = 4-way conflicts).
__device__ __forceinline__ double2 lds_f64x2(uint32_t addr)
{
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ double flip_sign(double v, uint32_t bit)      // bit = 0 or 1
{
    return __hiloint2double(__double2hiint(v) ^ (int)(bit << 31), __double2loint(v));
}

template <int NB, int NW, int PROD>
__global__ void __launch_bounds__((NW + PROD) * 32, 1) k_schur_mma(BaProblem pb, TilePlan tp, SchurPlan sp, const double *__restrict__ Z,
                                                                 const uint8_t *__restrict__ zsign, const uint8_t *__restrict__ pt_rank,
                                                                 int mp, double *__restrict__ Spart, const int *__restrict__ done)
{
    if (*done) return;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const uint32_t plane = (uint32_t)tp.tile_obs * 32u + 32u;                   // one chunk plane (+32 B: consecutive planes start 8 banks apart)
    const uint32_t zbytes = 12u * plane;
    const uint32_t stage = zbytes + (uint32_t)kTilePts * (uint32_t)mp;          // multiple of 16
    unsigned char *zero_blk = smem_raw + 2 * stage;
    __shared__ __align__(8) uint64_t mfull[2], mempty[2];
    __shared__ uint2 hitlist[NW][kTilePts + 4];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g8 = lane >> 2, kk = lane & 3;
    const int chunk = blockIdx.x, grp = blockIdx.y;
    const int ntl = (tp.ntiles > chunk) ? (tp.ntiles - chunk + sp.nchunk - 1) / sp.nchunk : 0;   // tiles of this chunk

    if (threadIdx.x < 96) reinterpret_cast<uint32_t *>(zero_blk)[threadIdx.x] = 0u;        // a whole Z row of zeros (384 B)
    // the pad slot (index TO) of every chunk plane of both stages is never written by the copies: it is the all-zero row
    for (int t = threadIdx.x; t < 2 * 12 * 8; t += blockDim.x)
        reinterpret_cast<uint32_t *>(smem_raw + (size_t)(t / 96) * stage + (size_t)((t % 96) / 8) * plane + (size_t)tp.tile_obs * 32u)[t % 8] = 0u;
    if (threadIdx.x == 0) { mbar_init(&mfull[0], 1); mbar_init(&mfull[1], 1); mbar_init(&mempty[0], NW); mbar_init(&mempty[1], NW); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();

    // producer duty (warp 0, lane 0): bulk-async copies of the Z rows and the rank rows of local tile `it` into stage it & 1,
    // after every warp has released that stage (mempty)
    auto produce = [&](int it) {
        if (warp == (PROD ? NW : 0) && lane == 0) {
            const int b = it & 1;
            if (it >= 2) mbar_wait(&mempty[b], (uint32_t)(((it >> 1) - 1) & 1));
            const int t = chunk + it * sp.nchunk;
            const int p0 = tp.tile_pt[t], p1 = tp.tile_pt[t + 1];
            const int o0 = pb.obs_ptr[p0], o1 = pb.obs_ptr[p1];
            const uint32_t zb = (uint32_t)(o1 - o0) * 384u, rb = (uint32_t)kTilePts * (uint32_t)mp; (void)p1;
            unsigned char *dst = smem_raw + (size_t)b * stage;
            const char *src = reinterpret_cast<const char *>(Z + 48 * (int64_t)o0);
            mbar_expect_tx(&mfull[b], zb + rb);
            const uint32_t pl = (uint32_t)(o1 - o0) * 32u;                     // bytes of one chunk plane of this tile
            if (pl) for (uint32_t q = 0; q < 12u; ++q) bulk_g2s(dst + q * plane, src + (size_t)q * pl, pl, &mfull[b]);
            bulk_g2s(dst + zbytes, pt_rank + (size_t)t * rb, rb, &mfull[b]);
        }
        __syncwarp();
    };
    if (PROD) {                      // dedicated producer warp: runs ahead of the consumers by up to two stages
        if (warp == NW) { for (int it = 0; it < ntl; ++it) produce(it); return; }
    } else if (ntl > 0) produce(0);

    // ---------------- consumer warps ----------------
    int bj[NB], bk[NB];
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        const uint16_t jk = sp.blk_jk[((size_t)grp * NW + warp) * NB + b];
        bj[b] = (jk == 0xFFFF) ? -1 : (jk & 0xFF);
        bk[b] = (jk == 0xFFFF) ? -1 : (jk >> 8);
    }
    double acc[NB][8];
#pragma unroll
    for (int b = 0; b < NB; ++b)
#pragma unroll
        for (int q = 0; q < 8; ++q) acc[b][q] = 0.0;

    const uint32_t smem_base = smem_u32(smem_raw);
    uint2 *hl = hitlist[warp];
    // byte offset of 16-byte element x = g8 * 3 + component of a row, relative to the row's chunk slot in plane 0
    auto elem_off = [&](int x) { return (uint32_t)(x >> 1) * plane + (uint32_t)(x & 1) * 16u; };
    const uint32_t lane_off = elem_off(g8 * 3 + (kk < 3 ? kk : 0));
    // (hit-in-group, component) of this k-lane in the three steps of a full 4-hit group
    const int gh0 = (kk == 3) ? 1 : 0, gh1 = (kk == 0 || kk == 1) ? 1 : 2, gh2 = (kk == 0) ? 2 : 3;
    const int gc0 = (kk == 3) ? 0 : kk, gc1 = (kk == 0) ? 1 : (kk == 1) ? 2 : (kk == 2) ? 0 : 1, gc2 = (kk == 0) ? 2 : kk - 1;
    const uint32_t goff0 = elem_off(g8 * 3 + gc0), goff1 = elem_off(g8 * 3 + gc1), goff2 = elem_off(g8 * 3 + gc2);

    // per-lane metadata of the point this lane represents in a tile: local observation base and the sign bits of D.
    // Two-level prefetch so no load waits on another: tile bounds two tiles ahead, per-point data one tile ahead.
    auto load_bounds = [&](int it, int &p0, int &p1, int &o0) {
        p0 = 0; p1 = 0; o0 = 0;
        if (it < ntl) { const int t = chunk + it * sp.nchunk; p0 = tp.tile_pt[t]; p1 = tp.tile_pt[t + 1]; o0 = tp.tile_o[t]; }
    };
    auto load_meta = [&](int p0, int p1, int o0, int &ob, uint32_t &sg, bool &have) {
        ob = 0; sg = 0; have = false;
        const int p = p0 + lane;
        if (p < p1) { ob = pb.obs_ptr[p] - o0; sg = zsign[p]; have = true; }
    };
    int bp0, bp1, bo0;                       // bounds of tile it + 1 (then it + 2)
    int ob_n; uint32_t sg_n; bool hv_n;
    load_bounds(0, bp0, bp1, bo0);
    load_meta(bp0, bp1, bo0, ob_n, sg_n, hv_n);
    load_bounds(1, bp0, bp1, bo0);

    for (int it = 0; it < ntl; ++it) {
        const int ob = ob_n; const uint32_t sg = sg_n; const bool have = hv_n;
        if (!PROD && it + 1 < ntl) produce(it + 1);
        load_meta(bp0, bp1, bo0, ob_n, sg_n, hv_n);          // tile it + 1 (bounds arrived during the previous tile)
        load_bounds(it + 2, bp0, bp1, bo0);
        const int sb = it & 1;
        mbar_wait(&mfull[sb], (uint32_t)((it >> 1) & 1));
        const uint32_t zoff = (uint32_t)sb * stage;                           // byte offset of this stage from smem_base
        const uint32_t zbase = smem_base + zoff;
        const uint32_t zrow_addr = zbase + (uint32_t)tp.tile_obs * 32u;       // plane-0 address of the all-zero row of this stage
        const unsigned char *rk = smem_raw + zoff + zbytes + lane;          // rank tile is camera-major: [cam][32 points]
        const bool anyneg = __any_sync(0xffffffffu, sg != 0);

#pragma unroll
        for (int b = 0; b < NB; ++b) {
            const int j = bj[b], k = bk[b];
            if (j < 0) continue;
            const uint32_t rj = have ? rk[j * kTilePts] : 0xFFu, rkk = have ? rk[k * kTilePts] : 0xFFu;
            const bool hit = (rj | rkk) < 0x80u;
            const uint32_t bal = __ballot_sync(0xffffffffu, hit);
            const int nh = __popc(bal);
            if (nh == 0) continue;
            if (hit) {
                const int pos = __popc(bal & ((1u << lane) - 1u));
                // shared-memory byte addresses of the two Z rows of this hit; bits 28..30 of .x carry the sign bits of D
                const uint32_t aj = zbase + (uint32_t)(ob + (int)rj) * 32u, ak = zbase + (uint32_t)(ob + (int)rkk) * 32u;
                hl[pos] = make_uint2(aj | (sg << 28), ak);
            }
            if (lane < 3) hl[nh + lane] = make_uint2(zrow_addr, zrow_addr);     // missing hits of the last group -> the zero slot
            __syncwarp();
            // K packing.  Full groups of 4 hits fill 3 DMMA steps completely (12 K-slots): k-lane kk handles the (hit, component)
            // pairs  kk0: (0,0)(1,1)(2,2)  kk1: (0,1)(1,2)(3,0)  kk2: (0,2)(2,0)(3,1)  kk3: (1,0)(2,1)(3,2)  of the group in steps
            // 0,1,2.  The remaining 1..3 hits take one step each (k-lanes 0..2 = the three components, k-lane 3 reads zeros).
            // Every group of up to 4 hits goes through the same three packed steps; the slots of a last, partial group that have no
            // hit read the all-zero pad slot of the chunk planes (hit-list entries nh .. nh + 2 point there), and steps that would
            // only multiply zeros are skipped (a group of r hits needs min(r, 3) steps).
            const int nfull = nh & ~3, rem = nh - nfull;
            if (j == k) {
                for (int q = 0; q < nfull; q += 4) {
                    const uint2 e0 = hl[q + gh0], e1 = hl[q + gh1], e2 = hl[q + gh2];
                    const double2 a0 = lds_f64x2((e0.x & 0x0FFFFFFFu) + goff0), a1 = lds_f64x2((e1.x & 0x0FFFFFFFu) + goff1),
                                  a2 = lds_f64x2((e2.x & 0x0FFFFFFFu) + goff2);
                    double2 s0 = a0, s1 = a1, s2 = a2;
                    if (anyneg) {
                        const uint32_t b0 = ((e0.x >> 28) >> gc0) & 1u, b1 = ((e1.x >> 28) >> gc1) & 1u, b2 = ((e2.x >> 28) >> gc2) & 1u;
                        s0.x = flip_sign(s0.x, b0); s0.y = flip_sign(s0.y, b0); s1.x = flip_sign(s1.x, b1); s1.y = flip_sign(s1.y, b1);
                        s2.x = flip_sign(s2.x, b2); s2.y = flip_sign(s2.y, b2);
                    }
                    dmma884(acc[b][0], acc[b][1], s0.x, a0.x); dmma884(acc[b][2], acc[b][3], s0.x, a0.y); dmma884(acc[b][6], acc[b][7], s0.y, a0.y);
                    dmma884(acc[b][0], acc[b][1], s1.x, a1.x); dmma884(acc[b][2], acc[b][3], s1.x, a1.y); dmma884(acc[b][6], acc[b][7], s1.y, a1.y);
                    dmma884(acc[b][0], acc[b][1], s2.x, a2.x); dmma884(acc[b][2], acc[b][3], s2.x, a2.y); dmma884(acc[b][6], acc[b][7], s2.y, a2.y);
                }
                if (rem) {          // last, partial group: same packed steps, all loads issued before the first DMMA
                    const uint2 e0 = hl[nfull + gh0];
                    uint2 e1 = e0, e2 = e0;
                    double2 a0 = lds_f64x2((e0.x & 0x0FFFFFFFu) + goff0), a1 = a0, a2 = a0;
                    if (rem >= 2) { e1 = hl[nfull + gh1]; a1 = lds_f64x2((e1.x & 0x0FFFFFFFu) + goff1); }
                    if (rem >= 3) { e2 = hl[nfull + gh2]; a2 = lds_f64x2((e2.x & 0x0FFFFFFFu) + goff2); }
                    double2 s0 = a0, s1 = a1, s2 = a2;
                    if (anyneg) {
                        const uint32_t b0 = ((e0.x >> 28) >> gc0) & 1u, b1 = ((e1.x >> 28) >> gc1) & 1u, b2 = ((e2.x >> 28) >> gc2) & 1u;
                        s0.x = flip_sign(s0.x, b0); s0.y = flip_sign(s0.y, b0); s1.x = flip_sign(s1.x, b1); s1.y = flip_sign(s1.y, b1);
                        s2.x = flip_sign(s2.x, b2); s2.y = flip_sign(s2.y, b2);
                    }
                    dmma884(acc[b][0], acc[b][1], s0.x, a0.x); dmma884(acc[b][2], acc[b][3], s0.x, a0.y); dmma884(acc[b][6], acc[b][7], s0.y, a0.y);
                    if (rem >= 2) { dmma884(acc[b][0], acc[b][1], s1.x, a1.x); dmma884(acc[b][2], acc[b][3], s1.x, a1.y); dmma884(acc[b][6], acc[b][7], s1.y, a1.y); }
                    if (rem >= 3) { dmma884(acc[b][0], acc[b][1], s2.x, a2.x); dmma884(acc[b][2], acc[b][3], s2.x, a2.y); dmma884(acc[b][6], acc[b][7], s2.y, a2.y); }
                }
            } else {
                for (int q = 0; q < nfull; q += 4) {
                    const uint2 e0 = hl[q + gh0], e1 = hl[q + gh1], e2 = hl[q + gh2];
                    double2 a0 = lds_f64x2((e0.x & 0x0FFFFFFFu) + goff0), a1 = lds_f64x2((e1.x & 0x0FFFFFFFu) + goff1),
                            a2 = lds_f64x2((e2.x & 0x0FFFFFFFu) + goff2);
                    const double2 b0 = lds_f64x2(e0.y + goff0), b1 = lds_f64x2(e1.y + goff1), b2 = lds_f64x2(e2.y + goff2);
                    if (anyneg) {
                        const uint32_t t0 = ((e0.x >> 28) >> gc0) & 1u, t1 = ((e1.x >> 28) >> gc1) & 1u, t2 = ((e2.x >> 28) >> gc2) & 1u;
                        a0.x = flip_sign(a0.x, t0); a0.y = flip_sign(a0.y, t0); a1.x = flip_sign(a1.x, t1); a1.y = flip_sign(a1.y, t1);
                        a2.x = flip_sign(a2.x, t2); a2.y = flip_sign(a2.y, t2);
                    }
                    dmma884(acc[b][0], acc[b][1], a0.x, b0.x); dmma884(acc[b][2], acc[b][3], a0.x, b0.y);
                    dmma884(acc[b][4], acc[b][5], a0.y, b0.x); dmma884(acc[b][6], acc[b][7], a0.y, b0.y);
                    dmma884(acc[b][0], acc[b][1], a1.x, b1.x); dmma884(acc[b][2], acc[b][3], a1.x, b1.y);
                    dmma884(acc[b][4], acc[b][5], a1.y, b1.x); dmma884(acc[b][6], acc[b][7], a1.y, b1.y);
                    dmma884(acc[b][0], acc[b][1], a2.x, b2.x); dmma884(acc[b][2], acc[b][3], a2.x, b2.y);
                    dmma884(acc[b][4], acc[b][5], a2.y, b2.x); dmma884(acc[b][6], acc[b][7], a2.y, b2.y);
                }
                if (rem) {          // last, partial group: same packed steps, all loads issued before the first DMMA
                    const uint2 e0 = hl[nfull + gh0];
                    uint2 e1 = e0, e2 = e0;
                    double2 a0 = lds_f64x2((e0.x & 0x0FFFFFFFu) + goff0), a1 = a0, a2 = a0;
                    double2 b0 = lds_f64x2(e0.y + goff0), b1 = b0, b2 = b0;
                    if (rem >= 2) { e1 = hl[nfull + gh1]; a1 = lds_f64x2((e1.x & 0x0FFFFFFFu) + goff1); b1 = lds_f64x2(e1.y + goff1); }
                    if (rem >= 3) { e2 = hl[nfull + gh2]; a2 = lds_f64x2((e2.x & 0x0FFFFFFFu) + goff2); b2 = lds_f64x2(e2.y + goff2); }
                    if (anyneg) {
                        const uint32_t t0 = ((e0.x >> 28) >> gc0) & 1u, t1 = ((e1.x >> 28) >> gc1) & 1u, t2 = ((e2.x >> 28) >> gc2) & 1u;
                        a0.x = flip_sign(a0.x, t0); a0.y = flip_sign(a0.y, t0); a1.x = flip_sign(a1.x, t1); a1.y = flip_sign(a1.y, t1);
                        a2.x = flip_sign(a2.x, t2); a2.y = flip_sign(a2.y, t2);
                    }
                    dmma884(acc[b][0], acc[b][1], a0.x, b0.x); dmma884(acc[b][2], acc[b][3], a0.x, b0.y);
                    dmma884(acc[b][4], acc[b][5], a0.y, b0.x); dmma884(acc[b][6], acc[b][7], a0.y, b0.y);
                    if (rem >= 2) {
                        dmma884(acc[b][0], acc[b][1], a1.x, b1.x); dmma884(acc[b][2], acc[b][3], a1.x, b1.y);
                        dmma884(acc[b][4], acc[b][5], a1.y, b1.x); dmma884(acc[b][6], acc[b][7], a1.y, b1.y);
                    }
                    if (rem >= 3) {
                        dmma884(acc[b][0], acc[b][1], a2.x, b2.x); dmma884(acc[b][2], acc[b][3], a2.x, b2.y);
                        dmma884(acc[b][4], acc[b][5], a2.y, b2.x); dmma884(acc[b][6], acc[b][7], a2.y, b2.y);
                    }
                }
            }
            __syncwarp();
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&mempty[sb]);       // this warp is done with the stage
    }

#pragma unroll
    for (int b = 0; b < NB; ++b) {
        const int pair = sp.blk_pair[((size_t)grp * NW + warp) * NB + b];
        if (pair < 0) continue;
        double *o = Spart + ((size_t)chunk * sp.npairs + pair) * 256 + lane * 8;
#pragma unroll
        for (int q = 0; q < 8; q += 2) *reinterpret_cast<double2 *>(o + q) = make_double2(acc[b][q], acc[b][q + 1]);
    }
}

// Assemble S, E from fragment-ordered Schur sums (see k_schur_mma) and the reduced U / ea:
//   S_jk = delta_jk (U_j + mu_u I) - P_jk,  E_j = Ered_j (the tile lineariser produces E = sum A^T e' directly); the lower block triangle (and the (hi,lo) sub-tile of the
//   diagonal blocks, which the MMA kernel skips) is mirrored.
__global__ void k_assemble_mma(int m, int mcon, const double *__restrict__ Ured, const double *__restrict__ Sred,
                               const double *__restrict__ Ered, const LmState *__restrict__ st, double *__restrict__ S, double *__restrict__ E)
{
    if (st->done) return;
    const double mu_u = st->mu_u;
    const int mf = m - mcon, N = 16 * mf;
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < (int64_t)N * N) {
        const int row = (int)(t / N), col = (int)(t % N);
        int a = row >> 4, b = col >> 4, r = row & 15, c = col & 15;
        if (a > b || (a == b && r > c)) { int q = a; a = b; b = q; q = r; r = c; c = q; }
        const int pair = a * mf - (a * (a - 1)) / 2 + (b - a);
        const int ri = r >> 3, ci = c >> 3;
        const int ln = ((r & 7) << 2) | ((c & 7) >> 1);
        const int q = (ri * 2 + ci) * 2 + (c & 1);
        double s = -Sred[(int64_t)pair * 256 + ln * 8 + q];
        if (a == b) {
            s += Ured[(mcon + a) * kU + tri16(r, c)];
            if (r == c) s += mu_u;
        }
        S[t] = s;
    }
    if (t < N) {
        const int j = mcon + (int)(t >> 4), r = (int)(t & 15);
        E[t] = Ered[j * 16 + r];
    }
}

}  // namespace argus
