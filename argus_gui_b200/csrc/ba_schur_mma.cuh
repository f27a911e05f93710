// Schur-complement kernel: fp64 tensor-pipe (DMMA m8n8k4) rank-3 updates with the S blocks stationary in registers, operands
// streamed through shared memory by bulk-async (TMA) copies, and STATIC hit lists.
//
// Math.  With V*_i = L_i D_i L_i^T (3x3, per point) and  Z_ij = W_ij L_i^-T |D_i|^-1/2  (16x3, per observation),
//     sum_i Y_ij W_ik^T = sum_i Z_ij diag(sign D_i) Z_ik^T            (sba_levmar.c:1075-1114 computes the left side)
// so the whole Schur complement is a sum of signed rank-3 outer products of ONE operand array Z (384 B per
// observation) instead of two (Y and W).  For a damped, positive-definite V* all signs are +.
//
// Work decomposition.  The upper-triangular 16x16 camera blocks (j <= k) are dealt to G groups x NW warps x NB slots.
// A CTA = (chunk of point tiles, group); each warp keeps its NB blocks as DMMA accumulators (8 doubles per lane per block)
// for the whole chunk.  Per tile (<= 32 points, <= 256 rows) the CTA's Z block arrives by cp.async.bulk, one copy per chunk
// plane, into a double-buffered shared-memory stage, together with the tile's slice of the static index.
//
// Static index (built once per upload, visibility never changes during a solve; k_hits_*): for every (tile, block) the list of
// "hits" - points of the tile that see both cameras of the block - as (row of Z_ij, row of Z_ik) byte pairs, plus a table
// (offset, count) per block.  The kernel does no visibility test at all: it walks lists.
//
// Rows, classes and bank conflicts.  A tile's Z block has one 32-byte slot per row and chunk plane, so the shared-memory bank
// group (16 bytes) of element x = 3 (row pair g) + component of row s is (x + 2 s) mod 8.  Rows are not the tile's
// observations in order: every point is given one of four CLASSES and its observations occupy rows 4 i + class.  A group of
// four hits of four different classes is processed in three DMMA steps, step t = component t of all four hits, k-lane = hit:
// the eight lanes of a quarter warp then read eight different bank groups - conflict free - and the K dimension of the
// DMMA is full (4 hits x 3 components = 3 steps x 4).  The index build orders every list so that its groups of four mix the
// classes as far as the list allows; the 1-3 hits left over take one step each (k-lanes 0-2 = the three components of one
// hit, six consecutive elements per quarter warp: conflict free by construction).
// Fragment layout of mma.m8n8k4.f64: A: lane -> (row = lane>>2, k = lane&3); B: (k = lane&3, col = lane>>2);
// C: (row = lane>>2, cols 2*(lane&3), +1).  The 48 values of an observation are ordered ((r%8)*3 + c)*2 + r/8 so one
// 16-byte load gives a lane the (row, row+8) pair of its fragment - the same load serves the A and the B operand;
// consecutive 32-byte chunks of that order live in consecutive chunk planes of the tile (coalesced producer stores).
#pragma once
#include "ba_common.cuh"

namespace argus {

constexpr int kTilePts = 32;        // points per tile
constexpr int kTileRows = 256;      // Z rows per tile: 4 classes x 64
constexpr int kClassCap = 64;       // observations per class and tile
constexpr int kSchurWarps = 8;

__host__ __device__ __forceinline__ int zperm(int r, int c) { return ((r & 7) * 3 + c) * 2 + (r >> 3); }

struct TilePlan {
    int ntiles;
    const int32_t *tile_pt;         // ntiles + 1: first point of each tile
    const int32_t *tile_o;          // ntiles + 1: first observation of each tile
    const int32_t *tile_r;          // ntiles + 1: first Z row of each tile (a tile has 4 * max class fill rows)
};

struct SchurPlan {
    int G;                          // block groups
    int nchunk;                     // point-tile chunks (grid.x)
    int npairs;                     // upper-triangular camera pairs among the free cameras
    int nslots;                     // NW * NB block slots per group
    const uint16_t *blk_jk;         // [G][NW][NB]: j | k << 8 (absolute camera ids), 0xFFFF = empty slot
    const int32_t *blk_pair;        // [G][NW][NB]: pair index (row-major upper triangle over free cameras) or -1
    const uint16_t *hit_blob;       // per (tile, group) ONE contiguous record, copied into shared memory by one bulk copy:
                                    //   table, nslots + 4 words: [b] = offset of block b's list (in entries) | hits << 16 | diagonal block << 31,
                                    //                            [nslots] = Z rows of the tile;
                                    //   then the entries: row of Z_ij | row of Z_ik << 8
    const int32_t *hit_ptr;         // [ntiles * G + 1]: start of every record in hit_blob, in 16-bit units (multiples of 8)
    int *pace;                      // G == 2 only: [nchunk][8] tiles fetched so far by the two CTAs of a chunk (zeroed before the launch)
                                    // or null.  They read the same Z tiles; kept within pace_slack tiles of each other, the
                                    // second one finds them in L2
    int pace_slack;
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
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
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
__device__ __forceinline__ double2 lds_f64x2(uint32_t addr)
{
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint32_t lds_u16(uint32_t addr)
{
    uint32_t v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ double flip_sign(double v, uint32_t bit)      // bit = 0 or 1
{
    return __hiloint2double(__double2hiint(v) ^ (int)(bit << 31), __double2loint(v));
}
__device__ __forceinline__ double2 flip_sign2(double2 v, uint32_t bit) { v.x = flip_sign(v.x, bit); v.y = flip_sign(v.y, bit); return v; }

// Z in HBM: per tile one block of 48 * (nrows + 1) doubles, chunk-major inside the block: the 12 four-double chunks of a row
// (chunk = zperm / 4) live in 12 planes of nrows + 1 chunks each (the last one is padding: with nrows a multiple of 4,
// consecutive planes start 32 bytes = 8 banks apart in shared memory), so that the one-thread-per-row producer writes
// consecutive 32-byte chunks (coalesced) and the Schur kernel brings the WHOLE tile into shared memory with one bulk copy.
//   element e (= zperm(r, c)) of row `row` of tile `tile`, whose first row is r0 (the padding adds one row per earlier tile):
__host__ __device__ __forceinline__ int64_t z_index(int64_t r0, int64_t tile, int nrows, int row, int e)
{
    return 48 * (r0 + tile) + ((int64_t)(e >> 2) * (nrows + 1) + row) * 4 + (e & 3);
}

// ---------------------------------------------------------------------------------------------------------------------------
// Static hit lists.  One CTA per tile; a warp takes block slots round robin, lane = point of the tile.
//   pass 0 (k_hits_count): per (tile, group) the size of the record (table + entries, in 16-bit units, rounded up to 8: the
//                          16-byte granularity of the bulk copies)
//   scan (k_index_scan of ba_lin_tile.cuh) -> hit_ptr
//   pass 1 (k_hits_fill): the table and the entries in processing order
// Order of a list of n hits: sorted by class, the classes by ascending population, then dealt round robin to the n / 4 full
// groups (hit at sorted position q < 4 (n / 4) -> group q mod (n / 4), k-lane q div (n / 4)): a group receives two hits of one
// class only when that class holds more than n / 4 of the hits; the n mod 4 hits left over come from the most populous class.
// pt_slot0[p] = 4 * (first row index of point p inside its class) + class;  its observation of rank r sits in row pt_slot0 + 4 r.
// ---------------------------------------------------------------------------------------------------------------------------
struct HitCtx {
    uint64_t mask;      // cameras seeing this lane's point (0 for lanes without a free point)
    uint32_t slot0;     // pt_slot0 of this lane's point
};
__device__ __forceinline__ HitCtx hit_ctx(int lane, int p0, int npts, int64_t ncon, const uint64_t *__restrict__ pt_mask,
                                          const uint8_t *__restrict__ pt_slot0)
{
    HitCtx h; h.mask = 0; h.slot0 = 0;
    if (lane < npts && (int64_t)p0 + lane >= ncon) { h.mask = pt_mask[p0 + lane]; h.slot0 = pt_slot0[p0 + lane]; }
    return h;
}

// hits per camera pair over all free points (j <= k, out[j * m + k]): the weights of the block -> warp assignment.  Integer
// atomics only (order independent).
__global__ void __launch_bounds__(256) k_pair_hist(int64_t n, int64_t ncon, int m, const uint64_t *__restrict__ pt_mask,
                                                   unsigned long long *__restrict__ out)
{
    extern __shared__ unsigned int s_h[];      // m * m
    for (int t = threadIdx.x; t < m * m; t += blockDim.x) s_h[t] = 0u;
    __syncthreads();
    for (int64_t i = ncon + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        uint64_t a = pt_mask[i];
        while (a) {
            const int j = __ffsll((long long)a) - 1;
            uint64_t b = a;
            while (b) { const int k = __ffsll((long long)b) - 1; b &= b - 1; atomicAdd(&s_h[j * m + k], 1u); }
            a &= a - 1;
        }
    }
    __syncthreads();
    for (int t = threadIdx.x; t < m * m; t += blockDim.x) if (s_h[t]) atomicAdd(&out[t], (unsigned long long)s_h[t]);
}

__global__ void __launch_bounds__(256) k_hits_count(int ntiles, int G, int nslots, const int32_t *__restrict__ tile_pt, int64_t ncon,
                                                    const uint64_t *__restrict__ pt_mask, const uint8_t *__restrict__ pt_slot0,
                                                    const uint16_t *__restrict__ blk_jk, int32_t *__restrict__ tg_count)
{
    extern __shared__ int s_tot[];       // G
    const int tile = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int g = threadIdx.x; g < G; g += blockDim.x) s_tot[g] = 0;
    __syncthreads();
    const int p0 = tile_pt[tile], npts = tile_pt[tile + 1] - p0;
    const HitCtx h = hit_ctx(lane, p0, npts, ncon, pt_mask, pt_slot0);
    for (int q = warp; q < G * nslots; q += 8) {
        const uint32_t jk = blk_jk[q];
        if (jk == 0xFFFFu) continue;
        const int j = jk & 0xFF, k = jk >> 8;
        const bool hit = ((h.mask >> j) & 1ull) && ((h.mask >> k) & 1ull);
        const int n = __popc(__ballot_sync(0xffffffffu, hit));
        if (lane == 0 && n) atomicAdd(&s_tot[q / nslots], n);          // integer: order independent
    }
    __syncthreads();
    for (int g = threadIdx.x; g < G; g += blockDim.x) tg_count[(size_t)tile * G + g] = (2 * (nslots + 4) + s_tot[g] + 7) & ~7;
}

__global__ void __launch_bounds__(256) k_hits_fill(int ntiles, int G, int nslots, const int32_t *__restrict__ tile_pt, int64_t ncon,
                                                   const uint64_t *__restrict__ pt_mask, const uint8_t *__restrict__ pt_slot0,
                                                   const uint16_t *__restrict__ blk_jk, const int32_t *__restrict__ tile_r,
                                                   const int32_t *__restrict__ hit_ptr, uint16_t *__restrict__ hit_blob)
{
    extern __shared__ int s_n[];         // G * nslots: hits per slot, then offsets
    const int tile = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int p0 = tile_pt[tile], npts = tile_pt[tile + 1] - p0;
    const HitCtx h = hit_ctx(lane, p0, npts, ncon, pt_mask, pt_slot0);
    const int nq = G * nslots;
    for (int q = warp; q < nq; q += 8) {
        const uint32_t jk = blk_jk[q];
        int n = 0;
        if (jk != 0xFFFFu) {
            const int j = jk & 0xFF, k = jk >> 8;
            n = __popc(__ballot_sync(0xffffffffu, ((h.mask >> j) & 1ull) && ((h.mask >> k) & 1ull)));
        }
        if (lane == 0) s_n[q] = n;
    }
    __syncthreads();
    for (int g = threadIdx.x; g < G; g += blockDim.x) {          // exclusive offsets inside each (tile, group) block
        int acc = 0;
        for (int q = g * nslots; q < (g + 1) * nslots; ++q) {
            const int n = s_n[q];
            const uint32_t jk = blk_jk[q];
            s_n[q] = (int)((uint32_t)acc | ((uint32_t)n << 16) | ((jk != 0xFFFFu && (jk & 0xFF) == (jk >> 8)) ? 0x80000000u : 0u));
            acc += n;
        }
    }
    __syncthreads();
    for (int q = warp; q < nq; q += 8) {
        const uint32_t tb = (uint32_t)s_n[q];
        uint16_t *rec = hit_blob + (size_t)hit_ptr[(size_t)tile * G + q / nslots];
        if (lane == 0) reinterpret_cast<uint32_t *>(rec)[q % nslots] = tb;
        if (lane == 1 && q % nslots == 0) reinterpret_cast<uint32_t *>(rec)[nslots] = (uint32_t)(tile_r[tile + 1] - tile_r[tile]);
        const int n = (int)((tb >> 16) & 0x7FFFu);
        if (!n) continue;
        const uint32_t jk = blk_jk[q];
        const int j = jk & 0xFF, k = jk >> 8;
        const bool hit = ((h.mask >> j) & 1ull) && ((h.mask >> k) & 1ull);
        const int cls = (int)(h.slot0 & 3u);
        // class populations, classes ranked by (population, class id) ascending
        uint32_t bal[4]; int cnt[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) { bal[c] = __ballot_sync(0xffffffffu, hit && cls == c); cnt[c] = __popc(bal[c]); }
        if (hit) {
            int before = 0, mine = 0; uint32_t mybal = 0;
#pragma unroll
            for (int c = 0; c < 4; ++c) if (c == cls) { mine = cnt[c]; mybal = bal[c]; }
#pragma unroll
            for (int c = 0; c < 4; ++c)
                if (cnt[c] < mine || (cnt[c] == mine && c < cls)) before += cnt[c];        // hits of classes ranked before mine
            const int pos = before + __popc(mybal & ((1u << lane) - 1u));
            const int F = n >> 2;
            const int idx = (pos < 4 * F) ? 4 * (pos % F) + pos / F : pos;
            const uint32_t ra = h.slot0 + 4u * (uint32_t)__popcll(h.mask & ((1ull << j) - 1ull));
            const uint32_t rb = h.slot0 + 4u * (uint32_t)__popcll(h.mask & ((1ull << k) - 1ull));
            rec[2 * (nslots + 4) + (tb & 0xFFFFu) + idx] = (uint16_t)(ra | (rb << 8));
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------------
// The Schur kernel.  grid = (nchunk, G); NW consumer warps, each owning NB blocks; the copy producer is lane 0 of warp 0.
// Dynamic shared memory: 2 stages of [Z tile: 12 chunk planes of (nrows + 1) x 32 bytes | the (tile, group) record: table, entries |
// sign bytes], filled by THREE bulk copies per tile (the copies are issued by one lane of a warp that also computes: with one
// copy per plane that warp fell a whole tile behind the others, which then waited for it at every tile).  The metadata of the
// next tile (first row, record bounds) reaches shared memory through cp.async, off that warp's critical path as well.
// Spart[chunk][pair][256]: per-chunk partial sums in fragment order (lane * 8 + q), q = (ri*2 + ci)*2 + e for the
// 8x8 sub-tile (ri, ci) of the block: element (row = ri*8 + lane>>2, col = ci*8 + 2*(lane&3) + e).
// ---------------------------------------------------------------------------------------------------------------------------
#ifdef ARGUS_SCHUR_TRACE
// debug builds only (tools/schur_trace.py): per (CTA < 4, tile < 256, warp) clock64 stamps {wait begin, wait end, tile done, loads issued}
__device__ long long *g_schur_trace = nullptr;
#define ARGUS_TRACE(slot_) do { if (g_schur_trace && blockIdx.x < 2 && it < 256) \
    g_schur_trace[((((size_t)(blockIdx.y * 2 + blockIdx.x) * 256 + it) * NW + warp) * 4) + (slot_)] = clock64(); } while (0)
#else
#define ARGUS_TRACE(slot_) do { } while (0)
#endif
constexpr int kPaceInts = 8;         // counters per chunk (G <= 8), one 32-byte row
constexpr int kPaceSpins = 1 << 14;  // polls without success after which a CTA stops pacing itself (the groups of a chunk are co-resident
                                     // whenever the GPU is ours alone - grid <= SMs, one CTA per SM - but nothing may depend on that)
constexpr int kSchurUnroll = 1;      // groups of four hits per trip of the inner loop
constexpr bool kSchurDiagPath = false;   // a separate 9-DMMA code path for the diagonal blocks (j == k): fewer DMMAs for 16 of 136 blocks, twice the code
constexpr uint32_t kZBytes = 12u * (kTileRows + 1) * 32u;            // largest tile: 12 planes of 257 chunks
__host__ __device__ constexpr uint32_t schur_ent_bytes(int nslots) { return (uint32_t)nslots * kTilePts * 2u; }
__host__ __device__ constexpr uint32_t schur_tab_bytes(int nslots) { return ((uint32_t)nslots + 4u) * 4u; }        // table + the tile's row count, 16-byte multiple
__host__ __device__ constexpr uint32_t schur_stage_bytes(int nslots) { return kZBytes + schur_tab_bytes(nslots) + schur_ent_bytes(nslots) + 16u + 256u; }

// one pacing step of the copy-issuing lane.  Not inlined, and its state (the snapshot of the other CTA's counter, "still
// pacing") lives in shared memory: the hot loop, which sits at the register limit, keeps nothing alive for it.
__device__ __noinline__ void schur_pace_step(int *row, int grp, int it, int slack, int *s_pace /* [0..3] snapshot of the row, [4] pacing on */)
{
    if (!s_pace[4]) return;
    asm volatile("st.relaxed.gpu.global.s32 [%0], %1;" ::"l"(row + grp), "r"(it) : "memory");
    const int need = it - slack;
    if (need > 1 && s_pace[grp ^ 1] < need) {            // rare: really ahead (the first snapshot is taken while tile 1 is fetched)
        int v = 0;
        for (int spins = 0; v < need; ++spins) {
            if (spins >= kPaceSpins) { s_pace[4] = 0; break; }
            asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(row + (grp ^ 1)) : "memory");
        }
    }
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(s_pace)), "l"(row) : "memory");   // next snapshot (L2 only)
}

// DIAG (compile time): a code path of its own for the diagonal blocks; skip_hl (run time, warp uniform): in the common path,
// leave out the (rows 8-15) x (columns 0-7) sub-tile, which a diagonal block gets from its transpose (k_assemble_mma)
template <bool DIAG, bool SIGNED>
__device__ __forceinline__ void schur_block(double (&acc)[8], uint32_t ent_addr, int n, uint32_t zbase, uint32_t off0, uint32_t off1,
                                            uint32_t off2, uint32_t offk, int kk, uint32_t zs_addr, bool skip_hl)
{
    constexpr bool anyneg = SIGNED;
    const int nfull = n >> 2, rem = n & 3;
    // ---- groups of four hits: step t = component t of the four hits, k-lane = hit ----
    uint32_t e = (nfull > 0) ? lds_u16(ent_addr + 2u * (uint32_t)kk) : 0u;
#pragma unroll kSchurUnroll
    for (int g = 0; g < nfull; ++g) {
        const uint32_t ra = zbase + (e & 0xFFu) * 32u, rb = zbase + (e >> 8) * 32u;
        double2 a0 = lds_f64x2(ra + off0), a1 = lds_f64x2(ra + off1), a2 = lds_f64x2(ra + off2);
        double2 b0, b1, b2;
        if (!DIAG) { b0 = lds_f64x2(rb + off0); b1 = lds_f64x2(rb + off1); b2 = lds_f64x2(rb + off2); }
        else { b0 = a0; b1 = a1; b2 = a2; }
        if (anyneg) {
            uint32_t sg;
            asm volatile("ld.shared.u8 %0, [%1];" : "=r"(sg) : "r"(zs_addr + (e & 0xFFu)));
            a0 = flip_sign2(a0, sg & 1u); a1 = flip_sign2(a1, (sg >> 1) & 1u); a2 = flip_sign2(a2, (sg >> 2) & 1u);
        }
        if (g + 1 < nfull) e = lds_u16(ent_addr + 8u * (uint32_t)(g + 1) + 2u * (uint32_t)kk);      // next group's entry, off the chain
        dmma884(acc[0], acc[1], a0.x, b0.x); dmma884(acc[2], acc[3], a0.x, b0.y);
        if (!DIAG && !skip_hl) dmma884(acc[4], acc[5], a0.y, b0.x);
        dmma884(acc[6], acc[7], a0.y, b0.y);
        dmma884(acc[0], acc[1], a1.x, b1.x); dmma884(acc[2], acc[3], a1.x, b1.y);
        if (!DIAG && !skip_hl) dmma884(acc[4], acc[5], a1.y, b1.x);
        dmma884(acc[6], acc[7], a1.y, b1.y);
        dmma884(acc[0], acc[1], a2.x, b2.x); dmma884(acc[2], acc[3], a2.x, b2.y);
        if (!DIAG && !skip_hl) dmma884(acc[4], acc[5], a2.y, b2.x);
        dmma884(acc[6], acc[7], a2.y, b2.y);
    }
    // ---- left-over hits: one step each, k-lanes 0..2 = the three components, k-lane 3 reads the zero row ----
    if (rem) {
        double2 av[3], bv[3];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            if (r < rem) {
                const uint32_t er = lds_u16(ent_addr + 8u * (uint32_t)nfull + 2u * (uint32_t)r);
                av[r] = make_double2(0.0, 0.0); bv[r] = av[r];
                if (kk < 3) {                                   // the fourth k-lane contributes zeros (no load: nothing to conflict with)
                    av[r] = lds_f64x2(zbase + (er & 0xFFu) * 32u + offk);
                    bv[r] = DIAG ? av[r] : lds_f64x2(zbase + (er >> 8) * 32u + offk);
                }
                if (anyneg) {
                    uint32_t sg;
                    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(sg) : "r"(zs_addr + (er & 0xFFu)));
                    av[r] = flip_sign2(av[r], (kk < 3) ? ((sg >> kk) & 1u) : 0u);
                }
            }
        }
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            if (r < rem) {
                dmma884(acc[0], acc[1], av[r].x, bv[r].x); dmma884(acc[2], acc[3], av[r].x, bv[r].y);
                if (!DIAG && !skip_hl) dmma884(acc[4], acc[5], av[r].y, bv[r].x);
                dmma884(acc[6], acc[7], av[r].y, bv[r].y);
            }
        }
    }
}

template <int NB, int NW>
__global__ void __launch_bounds__(NW * 32, 1) k_schur_mma(TilePlan tp, SchurPlan sp, const double *__restrict__ Z,
                                                          const uint8_t *__restrict__ zs, double *__restrict__ Spart,
                                                          const int *__restrict__ done)
{
    if (*done) return;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int kSlots = NW * NB;
    constexpr uint32_t kEnt = schur_ent_bytes(kSlots), kTab = schur_tab_bytes(kSlots), kStage = schur_stage_bytes(kSlots);
    __shared__ __align__(8) uint64_t mfull[2], mempty[2];
    __shared__ __align__(16) int s_pace[8];            // pacing (G == 2): [0..3] snapshot of the chunk's counters, [4] still pacing

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g8 = lane >> 2, kk = lane & 3;
    const int chunk = blockIdx.x, grp = blockIdx.y;
    const int ntl = (tp.ntiles > chunk) ? (tp.ntiles - chunk + sp.nchunk - 1) / sp.nchunk : 0;   // tiles of this chunk
    __shared__ __align__(16) int32_t s_meta[2][4];      // per tile: first Z row, first Z row of the next tile, record begin, record end

    if (threadIdx.x == 0) { s_pace[0] = 0; s_pace[1] = 0; s_pace[4] = 1; mbar_init(&mfull[0], 1); mbar_init(&mfull[1], 1); mbar_init(&mempty[0], NW); mbar_init(&mempty[1], NW); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();

    // producer duty (warp 0, lane 0): the tile's Z block, its (tile, group) record and its sign bytes into stage it & 1, after every
    // warp has released that stage (mempty)
    auto fetch_meta = [&](int it) {                      // asynchronous 4-byte copies: nothing to wait for now
        const int t = chunk + it * sp.nchunk;
        const uint32_t d = smem_u32(&s_meta[it & 1][0]);
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(tp.tile_r + t) : "memory");
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d + 4u), "l"(tp.tile_r + t + 1) : "memory");
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d + 8u), "l"(sp.hit_ptr + (size_t)t * sp.G + grp) : "memory");
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d + 12u), "l"(sp.hit_ptr + (size_t)t * sp.G + grp + 1) : "memory");
    };
    // pacing of the two CTAs of a chunk when G == 2 (they read the same Z tiles): each publishes the tile it is about to fetch and
    // looks at a snapshot of the other's counter that an asynchronous copy brought into shared memory during the previous tile
    // - nothing on the copy-issuing warp's critical path unless it really is more than pace_slack tiles ahead; then it waits (the
    // launch ends with the slower CTA anyway) and the tile comes from DRAM once instead of twice
    auto produce = [&](int it) {
        if (warp == 0 && lane == 0) {
            const int b = it & 1;
            asm volatile("cp.async.wait_all;" ::: "memory");
            int4 mt;
            asm volatile("ld.shared.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(mt.x), "=r"(mt.y), "=r"(mt.z), "=r"(mt.w) : "r"(smem_u32(&s_meta[b][0])) : "memory");
            if (it + 1 < ntl) fetch_meta(it + 1);
            if (sp.pace != nullptr) schur_pace_step(sp.pace + (size_t)chunk * kPaceInts, grp, it, sp.pace_slack, s_pace);
            if (it >= 2) mbar_wait(&mempty[b], (uint32_t)(((it >> 1) - 1) & 1));
            const int t = chunk + it * sp.nchunk;
            const uint32_t zb = (uint32_t)(mt.y - mt.x + 1) * 384u, rb = (uint32_t)(mt.w - mt.z) * 2u;
            unsigned char *dst = smem_raw + (size_t)b * kStage;
            mbar_expect_tx(&mfull[b], zb + rb + 256u);
            bulk_g2s(dst, Z + 48 * ((int64_t)mt.x + t), zb, &mfull[b]);
            bulk_g2s(dst + kZBytes, sp.hit_blob + mt.z, rb, &mfull[b]);
            bulk_g2s(dst + kZBytes + kTab + kEnt + 16u, zs + (size_t)t * 256, 256u, &mfull[b]);
#ifdef ARGUS_SCHUR_TRACE
            if (g_schur_trace && blockIdx.x < 2 && it < 256) g_schur_trace[((((size_t)(blockIdx.y * 2 + blockIdx.x) * 256 + it) * NW + 0) * 4) + 3] = clock64();
#endif
        }
        __syncwarp();
    };
    if (ntl > 0) {
        if (warp == 0 && lane == 0) fetch_meta(0);
        produce(0);
    }

    double acc[NB][8];
#pragma unroll
    for (int b = 0; b < NB; ++b)
#pragma unroll
        for (int q = 0; q < 8; ++q) acc[b][q] = 0.0;

    const uint32_t smem_base = smem_u32(smem_raw);

    for (int it = 0; it < ntl; ++it) {
        if (it + 1 < ntl) produce(it + 1);
        const int sb = it & 1;
        if (lane == 0) ARGUS_TRACE(0);
        mbar_wait(&mfull[sb], (uint32_t)((it >> 1) & 1));
        if (lane == 0) ARGUS_TRACE(1);
        const uint32_t zbase = smem_base + (uint32_t)sb * kStage;
        const uint32_t tab_base = zbase + kZBytes, ent_base = tab_base + kTab, zs_addr = ent_base + kEnt + 16u;
        // byte offset of the 16-byte element x = g8 * 3 + component of a row, relative to the row's slot in plane 0; a plane
        // has one chunk per row of THIS tile plus one of padding
        uint32_t pstride;
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(pstride) : "r"(tab_base + 4u * (uint32_t)kSlots));
        pstride = (pstride + 1u) * 32u;
        auto elem_off = [&](int x) { return (uint32_t)(x >> 1) * pstride + (uint32_t)(x & 1) * 16u; };
        const uint32_t off0 = elem_off(g8 * 3), off1 = elem_off(g8 * 3 + 1), off2 = elem_off(g8 * 3 + 2);
        const uint32_t offk = (kk == 1) ? off1 : (kk == 2) ? off2 : off0;
        uint32_t s0, s1;
        asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(s0), "=r"(s1) : "r"(zs_addr + 8u * (uint32_t)lane));
        const bool anyneg = __any_sync(0xffffffffu, (s0 | s1) != 0u);
        // table entry of a block: list offset | hits << 16 | diagonal-block flag << 31 (empty slots have no hits)
#define ARGUS_SCHUR_TILE(SIGNED_)                                                                                                   \
        _Pragma("unroll") for (int b = 0; b < NB; ++b) {                                                                            \
            uint32_t tb;                                                                                                            \
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tb) : "r"(tab_base + 4u * (uint32_t)(warp * NB + b)));                    \
            const int n = (int)((tb >> 16) & 0x7FFFu);                                                                              \
            if (n == 0) continue;                                                                                                   \
            const uint32_t ent_addr = ent_base + 2u * (tb & 0xFFFFu);                                                               \
            if (kSchurDiagPath && (tb >> 31)) schur_block<true, SIGNED_>(acc[b], ent_addr, n, zbase, off0, off1, off2, offk, kk, zs_addr, false); \
            else schur_block<false, SIGNED_>(acc[b], ent_addr, n, zbase, off0, off1, off2, offk, kk, zs_addr, (tb >> 31) != 0u);   \
        }
        if (anyneg) { ARGUS_SCHUR_TILE(true) } else { ARGUS_SCHUR_TILE(false) }
#undef ARGUS_SCHUR_TILE
        __syncwarp();
        if (lane == 0) ARGUS_TRACE(2);
        if (lane == 0) mbar_arrive(&mempty[sb]);       // this warp is done with the stage
    }

    if (sp.pace != nullptr && threadIdx.x == 0)      // nobody waits for a CTA that has fetched everything
        asm volatile("st.relaxed.gpu.global.s32 [%0], %1;" ::"l"(sp.pace + (size_t)chunk * kPaceInts + grp), "r"(0x7fffffff) : "memory");
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
