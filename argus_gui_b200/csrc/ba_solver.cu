// Sparse Levenberg-Marquardt bundle adjustment on B200: host side of the solver + C ABI.
//
// Replaces sba_motstr_levmar_x / sba_mot_levmar_x / sba_str_levmar_x (sba-1.6p/sba_levmar.c:449-1428, :1437-1980, :1987-2540)
// with their KDRTS / KDRT / KDS projection callbacks (libsbaprojs/sbaprojs.c:1135-1445).  Every array stays resident in HBM.
// The damping loop itself runs on the device (ba_lm_controller.cuh): one damping attempt is a fixed sequence of kernels (a
// CUDA graph) that reads its inputs from the LmState block and ends with the controller kernel; the host only queues attempts
// a few ahead of the device and watches asynchronous snapshots of that block.  With several GPUs (one process - or one host
// thread - per GPU, points sharded, cameras replicated) every rank takes identical decisions from all-reduced numbers; the two
// collectives of an attempt go through NCCL (bound at run time, nccl_dyn.h) or through a caller-supplied hook.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <float.h>
#include <vector>
#include <algorithm>
#include <chrono>
#include <thread>
#include <mutex>
#include <functional>
#include "../../include/argus_b200.h"
#include "ba_kernels.cuh"
#include "ba_schur_mma.cuh"
#include "ba_lin_tile.cuh"
#include "ba_blockdiag.cuh"
#include "ba_chol_coop.cuh"
#include "ba_chol_cluster.cuh"
#include "ba_lm_controller.cuh"
#include "ba_wand.cuh"
#include "nccl_dyn.h"

using namespace argus;

namespace {

enum ProfId { P_RESIDUAL = 0, P_LINEARIZE, P_LIN_REDUCE, P_SCHUR, P_SCHUR_REDUCE, P_ASSEMBLE, P_CHOL, P_BACKSUB, P_ALLREDUCE, P_CONTROL, P_LIN_STATS, P_COUNT };
const char *kProfNames[P_COUNT] = {"residual", "linearize", "lin_reduce", "schur", "schur_reduce", "assemble", "chol_solve", "backsub_eval",
                                   "allreduce", "controller", "linearize_stats"};

constexpr int kProfEvery = 4;            // ARGUS_BA_PROFILE with graphs: every fourth attempt carries the event-record nodes
constexpr int kDepth = 2;               // damping attempts the host queues ahead of the one it has seen the outcome of
constexpr int kSnap = kDepth + 2;       // ring of pinned LmState snapshots
constexpr int kLogCap = 4096;           // verbose > 1: attempts remembered for the per-attempt report

// One device allocation per uploaded problem: every buffer is carved out of it (two passes: measure, then assign).
struct Arena {
    char *base = nullptr; size_t size = 0, used = 0; bool measuring = false;
    void *take(size_t bytes) {
        const size_t off = (used + 255) & ~(size_t)255;
        used = off + bytes;
        return (measuring || !base) ? nullptr : base + off;
    }
};
thread_local Arena *g_arena = nullptr;     // set only inside argus_ba_upload (per calling thread)

template <typename T> struct DevBuf {
    T *p = nullptr; size_t cap = 0; bool owned = false;
    cudaError_t ensure(size_t n) {
        if (g_arena) {                       // arena mode: pointers are (re)assigned on every upload
            if (owned && p) cudaFree(p);
            owned = false;
            p = static_cast<T *>(g_arena->take((n ? n : 1) * sizeof(T)));
            cap = n;
            return cudaSuccess;
        }
        if (n <= cap && p) return cudaSuccess;
        if (owned && p) cudaFree(p);
        p = nullptr; cap = 0;
        cudaError_t e = cudaMalloc(&p, (n ? n : 1) * sizeof(T));
        if (e == cudaSuccess) { cap = n; owned = true; }
        return e;
    }
    void release() { if (owned && p) cudaFree(p); p = nullptr; cap = 0; owned = false; }
};

// host-side helper: split [0, n) over host threads (counting the observations per point at upload is a pure host streaming
// job); the thread count is divided by the number of ranks sharing the host
void parallel_for(int64_t n, int world, const std::function<void(int64_t, int64_t)> &fn, int64_t grain = 65536)
{
    unsigned hw = std::thread::hardware_concurrency();
    int cap = (int)std::min<unsigned>(hw ? hw : 4, 16) / std::max(1, world);
    if (cap < 1) cap = 1;
    int T = (int)std::min<int64_t>(cap, (n + grain - 1) / grain);
    if (T <= 1) { fn(0, n); return; }
    std::vector<std::thread> th;
    for (int q = 0; q < T; ++q) th.emplace_back([&, q] { fn(n * q / T, n * (q + 1) / T); });
    for (auto &x : th) x.join();
}

// process-wide bits that are expensive to set up per context (a one-shot call creates and destroys a context):
// kernel attributes already applied per device, and a free list of pinned blocks for the state snapshots
std::mutex g_once_mutex;
unsigned long long g_attr_done = 0;
std::vector<void *> g_pinned_free;
constexpr size_t kPinnedBytes = kSnap * sizeof(LmState) + 256;
void *pinned_take()
{
    {
        std::lock_guard<std::mutex> lk(g_once_mutex);
        if (!g_pinned_free.empty()) { void *p = g_pinned_free.back(); g_pinned_free.pop_back(); return p; }
    }
    void *p = nullptr;
    return cudaMallocHost(&p, kPinnedBytes) == cudaSuccess ? p : nullptr;
}
void pinned_give(void *p) { std::lock_guard<std::mutex> lk(g_once_mutex); g_pinned_free.push_back(p); }

// ... and small device arenas: cudaMalloc + cudaFree are a third of a one-shot call on a wand-sized problem (config 1).  A
// destroyed context parks its arena here if it is small (its stream has been synchronised); an upload that needs no more takes it.
constexpr size_t kArenaPoolMaxBytes = 32u << 20;
constexpr size_t kArenaPoolSlots = 4;
struct PooledArena { int device; char *base; size_t size; };
std::vector<PooledArena> g_arena_pool;
char *arena_take(int device, size_t need, size_t *got)
{
    if (need <= kArenaPoolMaxBytes) {
        std::lock_guard<std::mutex> lk(g_once_mutex);
        for (size_t i = 0; i < g_arena_pool.size(); ++i)
            if (g_arena_pool[i].device == device && g_arena_pool[i].size >= need) {
                char *b = g_arena_pool[i].base; *got = g_arena_pool[i].size;
                g_arena_pool.erase(g_arena_pool.begin() + i);
                return b;
            }
    }
    char *b = nullptr;
    if (cudaMalloc(&b, need) != cudaSuccess) return nullptr;
    *got = need;
    return b;
}
void arena_give(int device, char *base, size_t size)
{
    if (!base) return;
    if (size <= kArenaPoolMaxBytes) {
        std::lock_guard<std::mutex> lk(g_once_mutex);
        if (g_arena_pool.size() < kArenaPoolSlots) { g_arena_pool.push_back({device, base, size}); return; }
    }
    cudaFree(base);
}

int g_default_device = 0;      // device of the one-shot calls (argus_set_default_device)

}  // namespace

struct argus_ba_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    int sms = 148;
    bool coop_ok = true, cluster_ok = true;
    // comm
    int rank = 0, world = 1;
    argus_allreduce_fn allreduce = nullptr;
    void *comm_user = nullptr;
    argus_nccl::ncclComm_t nccl = nullptr;
    bool nccl_owned = false;
    // problem
    int64_t n = 0, ncon = 0, nobs = 0;
    int m = 0, mcon = 0, nfix_k = 0, nfix_d = 0;
    int mode = kModeMotStr;   // which half of the parameters moves (argus_ba_set_mode)
    DevBuf<int32_t> obs_ptr; DevBuf<uint8_t> obs_cam; DevBuf<double2> obs_uv; DevBuf<uint64_t> pt_mask;
    DevBuf<double> rot0, camb[2], ptsb[2], cam0, pts0;
    int cur_cam = 0, cur_pts = 0;      // host mirror of LmState::cur_* between solves
    // work
    DevBuf<double> V, eb, vstate, Vinv, dbuf;
    DevBuf<double> scpart, red12, max2, Spart, S, E, da_free, da_full, part3, att3, cam2, scal, respart, Upart2, cholwork;
    double *red1 = nullptr, *red2 = nullptr;     // red12 = [U, ea, cost, sum p^2 | Schur sums, E]: one all-reduce
    int64_t r1len = 0;
    DevBuf<int> flags;     // [0] singular V*, [1] Cholesky status, [2] cooperative Cholesky failure flag
    DevBuf<LmState> st;
    LmState *h_snap = nullptr;        // kSnap pinned snapshots
    cudaEvent_t snap_ev[kSnap] = {};
    double *lmlog = nullptr;          // device, verbose > 1 only
    // tile / MMA plan
    DevBuf<int32_t> tile_pt, tile_o, tile_r; DevBuf<uint16_t> blk_jk; DevBuf<int32_t> blk_pair; DevBuf<uint8_t> zsign, zs; DevBuf<double> Z;
    DevBuf<uint8_t> corder, pt_slot0, vmask_dev; DevBuf<int32_t> idx_btot; DevBuf<uint32_t> row_info; DevBuf<double2> row_uv;
    DevBuf<int> schur_pace; int pace_slack = 0;
    DevBuf<uint16_t> hit_blob; DevBuf<int32_t> hit_ptr;     // static hit lists of the Schur kernel: one record per (tile, group)
    DevBuf<uint8_t> cam_fix; bool cam_fix_on = false; DevBuf<uint16_t> cstart;
    // wand pairs (extension, ba_wand.cuh); their buffers live outside the arena: argus_ba_set_wand comes after the upload
    int64_t wand_n = 0; double wand_w = 0.0; int wand_grid = 0, wand_chunks = 0;
    DevBuf<int32_t> wand_ia, wand_ib, wand_of; DevBuf<double> wand_len, wand_rec, wand_Q, wand_part;
    int grid_lin = 1; size_t lin_smem = 0;
    int64_t nrows = 0;                 // Z rows (>= nobs: a tile has 4 * max class fill rows)
    int ntiles = 0, schur_G = 1, schur_NB = 1, schur_NW = 8, schur_nchunk = 1;
    Arena arena;
    int grid_pts128 = 1, grid_pts256 = 1, npairs = 1;
    // damping attempt as a CUDA graph (captured at the first solve after an upload)
    bool use_graph = true;
    cudaStream_t copy_stream = nullptr; cudaEvent_t copy_ev[2] = {nullptr, nullptr};     // upload: the large host arrays travel beside the index build
    cudaGraphExec_t graph_exec = nullptr;
    cudaGraphExec_t graph_exec_prof = nullptr;     // the same attempt with an event-record node on either side of every phase (ARGUS_BA_PROFILE)
    bool capture_prof = false;
    int64_t graph_launches = 0, graph_allreduces = 0;
    // profile
    bool profiling = false, capturing = false;
    struct Ev { int id; cudaEvent_t a, b; };
    std::vector<Ev> evs;
    std::vector<Ev> graph_evs;                     // events recorded by the nodes of graph_exec_prof: hold the times of its LAST replay
    std::vector<cudaEvent_t> ev_pool;
    double prof_ms[P_COUNT]; int64_t prof_n[P_COUNT];
    int64_t launches = 0, h2d_bytes = 0, d2h_bytes = 0, n_allreduce = 0;

    TilePlan tile_plan() const { TilePlan tp; tp.ntiles = ntiles; tp.tile_pt = tile_pt.p; tp.tile_o = tile_o.p; tp.tile_r = tile_r.p; return tp; }
    SchurPlan schur_plan() const {
        SchurPlan sp; sp.G = schur_G; sp.nchunk = schur_nchunk; sp.npairs = npairs; sp.nslots = schur_NW * schur_NB; sp.blk_jk = blk_jk.p;
        sp.blk_pair = blk_pair.p; sp.hit_blob = hit_blob.p; sp.hit_ptr = hit_ptr.p;
        sp.pace = (pace_slack > 0 && schur_G == 2 && schur_nchunk * schur_G <= sms) ? schur_pace.p : nullptr;
        sp.pace_slack = pace_slack;
        return sp;
    }
    LinTilePlan lin_plan() const { LinTilePlan lp; lp.corder = corder.p; lp.cstart = cstart.p; lp.row_info = row_info.p; lp.row_uv = row_uv.p; return lp; }
    LinTileOut lin_out() {
        LinTileOut o; o.Z = Z.p; o.zs = zs.p; o.zsign = zsign.p; o.Vinv = Vinv.p; o.V = V.p; o.eb = eb.p; o.vstate = vstate.p; o.Upart = Upart2.p;
        o.scpart = scpart.p; o.flag = flags.p; return o;
    }
    ParamBufs pbufs() const { ParamBufs b; b.cam[0] = camb[0].p; b.cam[1] = camb[1].p; b.pts[0] = ptsb[0].p; b.pts[1] = ptsb[1].p; return b; }
    const int *done_ptr() const { return &st.p->done; }
    BaProblem problem() const {
        BaProblem pb;
        pb.n = n; pb.ncon = (mode == kModeMot) ? n : ncon; pb.m = m; pb.mcon = (mode == kModeStr) ? m : mcon; pb.nobs = nobs; pb.nfix_k = nfix_k; pb.nfix_d = nfix_d; pb.cam_fix = cam_fix_on ? cam_fix.p : nullptr;
        pb.obs_ptr = obs_ptr.p; pb.obs_cam = obs_cam.p; pb.obs_uv = obs_uv.p; pb.pt_mask = pt_mask.p; pb.rot0 = rot0.p;
        pb.wand_of = wand_n > 0 ? wand_of.p : nullptr;
        return pb;
    }
    WandPlan wand_plan() const {
        WandPlan wp; wp.npairs = wand_n; wp.ia = wand_ia.p; wp.ib = wand_ib.p; wp.len = wand_len.p; wp.w = wand_w; wp.rec = wand_rec.p; wp.Q = wand_Q.p;
        return wp;
    }
};

namespace {

cudaEvent_t get_event(argus_ba_ctx *c)
{
    if (!c->ev_pool.empty()) { cudaEvent_t e = c->ev_pool.back(); c->ev_pool.pop_back(); return e; }
    cudaEvent_t e; cudaEventCreate(&e); return e;
}
struct ProfScope {
    argus_ba_ctx *c; int id; cudaEvent_t a = nullptr, b = nullptr; bool on;
    ProfScope(argus_ba_ctx *c_, int id_) : c(c_), id(id_), on(c_->capturing ? c_->capture_prof : c_->profiling) {
        if (on) { a = get_event(c); b = get_event(c); rec(a); }
    }
    ~ProfScope() { if (on) { rec(b); (c->capturing ? c->graph_evs : c->evs).push_back({id, a, b}); } }
    // while capturing, cudaEventRecordExternal makes the call an event-record NODE of the graph (a plain record only expresses a
    // dependency between captured streams)
    void rec(cudaEvent_t e) { if (c->capturing) cudaEventRecordWithFlags(e, c->stream, cudaEventRecordExternal); else cudaEventRecord(e, c->stream); }
};
void prof_collect(argus_ba_ctx *c)
{
    for (auto &e : c->evs) {
        float ms = 0.f;
        cudaEventSynchronize(e.b);
        cudaEventElapsedTime(&ms, e.a, e.b);
        c->prof_ms[e.id] += ms; c->prof_n[e.id] += 1;
        c->ev_pool.push_back(e.a); c->ev_pool.push_back(e.b);
    }
    c->evs.clear();
}

void drop_graph(argus_ba_ctx *c)
{
    if (c->graph_exec) { cudaGraphExecDestroy(c->graph_exec); c->graph_exec = nullptr; }
    if (c->graph_exec_prof) { cudaGraphExecDestroy(c->graph_exec_prof); c->graph_exec_prof = nullptr; }
    for (auto &e : c->graph_evs) { c->ev_pool.push_back(e.a); c->ev_pool.push_back(e.b); }
    c->graph_evs.clear();
}

// one collective step: every (buffer, count, op) of the list is reduced over the ranks, in place, on the solver's stream
struct Red { double *buf; int64_t count; int op; };
int allreduce_many(argus_ba_ctx *c, const Red *r, int nr)
{
    if (c->world <= 1) return ARGUS_OK;
    ProfScope ps(c, P_ALLREDUCE);
    c->n_allreduce += 1;
    if (c->nccl) {
        argus_nccl::Api &a = argus_nccl::api();
        int rc = 0;
        if (nr > 1) rc |= a.GroupStart();
        for (int q = 0; q < nr; ++q)
            rc |= a.AllReduce(r[q].buf, r[q].buf, (size_t)r[q].count, argus_nccl::ncclFloat64, r[q].op ? argus_nccl::ncclMax : argus_nccl::ncclSum,
                              c->nccl, c->stream);
        if (nr > 1) rc |= a.GroupEnd();
        if (rc) { fprintf(stderr, "argus_b200: NCCL all-reduce failed\n"); return ARGUS_E_COMM; }
        return ARGUS_OK;
    }
    if (!c->allreduce) return ARGUS_E_COMM;
    for (int q = 0; q < nr; ++q)
        if (c->allreduce(c->comm_user, r[q].buf, r[q].count, r[q].op, (void *)c->stream) != 0) return ARGUS_E_COMM;
    return ARGUS_OK;
}

size_t cam_smem(int m) { return (size_t)m * sizeof(CamConst); }

// ---- ||x - hx(p)||^2 at the current (or, motion-only driver, the trial) cameras; sum left in att3[0] ----
int enqueue_cost(argus_ba_ctx *c, int trial_cam, double *hx)
{
    {
        ProfScope ps(c, P_RESIDUAL);
        k_residual<<<c->grid_pts256, 256, cam_smem(c->m), c->stream>>>(c->problem(), c->pbufs(), c->st.p, trial_cam, hx, c->respart.p);
        if (c->wand_n > 0) {
            k_wand_cost<<<c->wand_grid, kWandThreads, 0, c->stream>>>(c->wand_plan(), c->pbufs(), c->st.p, c->respart.p + c->grid_pts256);
            c->launches += 1;
        }
        k_reduce_cost<<<1, 32, 0, c->stream>>>(c->respart.p, c->grid_pts256 + c->wand_grid, c->att3.p, c->done_ptr());
        c->launches += 2;
    }
    ARGUS_CUDA_CHECK(cudaGetLastError());
    Red r[1] = {{c->att3.p, 4, 0}};
    return allreduce_many(c, r, 1);
}

// ---- tile linearisation (mode 0: statistics only; mode 1: + damping, Z, E) and its fixed-order reduction into red1 ----
int enqueue_lin(argus_ba_ctx *c, int mode)
{
    const int m = c->m;
    {
        ProfScope ps(c, mode ? P_LINEARIZE : P_LIN_STATS);
#define LIN_LAUNCH(MC_) k_lin_tile<MC_><<<c->grid_lin, 256, c->lin_smem, c->stream>>>(c->problem(), c->tile_plan(), c->lin_plan(), c->pbufs(), c->st.p, mode, c->lin_out())
        const int mc = (m - c->mcon + 7) / 8;
        if (mc <= 2) LIN_LAUNCH(2); else if (mc <= 4) LIN_LAUNCH(4); else LIN_LAUNCH(8);
#undef LIN_LAUNCH
        if (c->wand_n > 0) {
            k_wand_prepare<<<c->wand_grid, kWandThreads, 0, c->stream>>>(c->problem(), c->wand_plan(), c->pbufs(), c->st.p, mode, c->V.p, c->eb.p,
                                                                         c->Vinv.p, c->scpart.p + 4 * (size_t)c->grid_lin);
            c->launches += 1;
            if (mode != 0) {
                const int64_t work = c->wand_n * (m - c->mcon);
                const int gq = (int)std::min<int64_t>((work + 127) / 128, (int64_t)c->sms * 16);
                k_wand_q<<<gq, 128, cam_smem(m), c->stream>>>(c->problem(), c->wand_plan(), c->pbufs(), c->st.p);
                c->launches += 1;
            }
        }
    }
    {
        ProfScope ps(c, P_LIN_REDUCE);
        k_lin_tile_reduce<<<m, 1024, 0, c->stream>>>(m, c->grid_lin, c->grid_lin + c->wand_grid, c->Upart2.p, c->scpart.p, c->red1, c->red2 + (int64_t)c->npairs * 256,
                                                   c->red1 + (int64_t)m * kU, c->max2.p, c->done_ptr());
    }
    c->launches += 2;
    ARGUS_CUDA_CHECK(cudaGetLastError());
    return ARGUS_OK;
}

// scalars of a linearisation from the (all-reduced) red1 / max2: scal = {cost, ||J^T e||_inf, ||p||^2, max diagonal}
int enqueue_lin_scalars(argus_ba_ctx *c)
{
    ProfScope ps(c, P_LIN_REDUCE);
    const int m = c->m;
    if (c->mode == kModeMotStr) k_lin_finalize<<<1, 32, 0, c->stream>>>(c->red1, c->pbufs(), c->st.p, m, c->mcon, c->red1 + (int64_t)m * kU, c->max2.p, c->scal.p);
    else k_bd_finalize<<<1, 32, 0, c->stream>>>(c->mode, c->red1, c->pbufs(), c->st.p, m, c->mcon, c->red1 + (int64_t)m * kU, c->max2.p, c->scal.p);
    c->launches += 1;
    ARGUS_CUDA_CHECK(cudaGetLastError());
    return ARGUS_OK;
}

// statistics-only linearisation at the current point with its collective (first iteration, parity hooks)
int enqueue_lin0(argus_ba_ctx *c)
{
    int rc = enqueue_lin(c, 0);
    if (rc) return rc;
    Red r[2] = {{c->red1, (int64_t)c->m * kU + 2, 0}, {c->max2.p, 2, 1}};
    if ((rc = allreduce_many(c, r, 2))) return rc;
    return enqueue_lin_scalars(c);
}

int enqueue_controller(argus_ba_ctx *c)
{
    ProfScope ps(c, P_CONTROL);
    k_lm_controller<<<1, 32, 0, c->stream>>>(c->st.p, c->flags.p, c->cam2.p, c->att3.p, c->scal.p);
    c->launches += 1;
    ARGUS_CUDA_CHECK(cudaGetLastError());
    return ARGUS_OK;
}

// ---- one damping attempt of the motion + structure driver: linearise with the damping applied (Z, E), Schur complement,
//      solve, back-substitute, evaluate the trial point; the controller decides ----
int enqueue_attempt_motstr(argus_ba_ctx *c, bool controller, double *dbout)
{
    const int m = c->m, mcon = c->mcon, mf = m - mcon, N = 16 * mf;
    cudaStream_t s = c->stream;
    const int *done = c->done_ptr();
    int rc = enqueue_lin(c, 1);
    if (rc) return rc;
    {
        ProfScope ps(c, P_SCHUR);
        dim3 grid(c->schur_nchunk, c->schur_G);
        if (c->schur_plan().pace) ARGUS_CUDA_CHECK(cudaMemsetAsync(c->schur_pace.p, 0, (size_t)c->schur_nchunk * kPaceInts * sizeof(int), s));
#define SCHUR_LAUNCH(NB_, NW_) k_schur_mma<NB_, NW_><<<grid, (NW_) * 32, 2 * schur_stage_bytes((NB_) * (NW_)), s>>>(c->tile_plan(), c->schur_plan(), c->Z.p, \
                                                                                                      c->zs.p, c->Spart.p, done)
        switch (c->schur_NB * 100 + c->schur_NW) {
            case 108: SCHUR_LAUNCH(1, 8); break;
            case 308: SCHUR_LAUNCH(3, 8); break;
            case 508: SCHUR_LAUNCH(5, 8); break;
            default: SCHUR_LAUNCH(6, 12); break;
        }
#undef SCHUR_LAUNCH
    }
    const int64_t slen = (int64_t)c->npairs * 256;
    {
        ProfScope ps(c, P_SCHUR_REDUCE);
        k_reduce_sum<<<(unsigned)((slen + 127) / 128), 128, 0, s>>>(c->Spart.p, c->schur_nchunk, slen, slen, c->red2, done);
        if (c->wand_n > 0) {     // rank-one corrections of the wand pairs: Schur sums -= sigma q q^T, E -= beta q
            const int64_t wlen = slen + 16 * m;
            k_wand_schur<<<dim3((c->npairs + kWandBlocks - 1) / kWandBlocks, c->wand_chunks), 256, wand_schur_smem(m), s>>>(c->problem(), c->wand_plan(), mf,
                                                                                                                          c->st.p, c->wand_part.p);
            k_wand_apply<<<(unsigned)((wlen + 255) / 256), 256, 0, s>>>(c->npairs, m, c->wand_chunks, c->wand_part.p, c->red2, done);
            c->launches += 2;
        }
    }
    c->launches += 2;
    ARGUS_CUDA_CHECK(cudaGetLastError());
    // one collective for [U, ea, cost, sum p^2 | Schur sums, E] (contiguous) and the two maxima
    Red r[2] = {{c->red1, c->r1len + slen + (int64_t)m * 16, 0}, {c->max2.p, 2, 1}};
    if ((rc = allreduce_many(c, r, 2))) return rc;
    if ((rc = enqueue_lin_scalars(c))) return rc;
    {
        ProfScope ps(c, P_ASSEMBLE);
        const int64_t tot = (int64_t)N * N;
        k_assemble_mma<<<(unsigned)((tot + 255) / 256), 256, 0, s>>>(m, mcon, c->red1, c->red2, c->red2 + slen, c->st.p, c->S.p, c->E.p);
    }
    {
        ProfScope ps(c, P_CHOL);
        if (c->cluster_ok && N <= kClMaxN) {
            // one thread-block cluster, hardware cluster barriers, the right-hand side carried through the factorisation
            k_chol_cluster<<<kClCtas, kClThreads, chol_cluster_smem(), s>>>(c->S.p, c->E.p, c->da_free.p, N, c->cholwork.p, c->flags.p + 1, c->flags.p + 2, done);
        } else if (c->coop_ok && N >= 192) {
            // multi-SM factorisation (cooperative grid): a single CTA is latency bound from about 200 unknowns
            double *Sp = c->S.p, *xp = c->da_free.p, *wp = c->cholwork.p; const double *Ep = c->E.p; int Nn = N;
            int *stt = c->flags.p + 1, *fl = c->flags.p + 2;
            void *args[] = {&Sp, &Ep, &xp, &Nn, &wp, &stt, &fl, &done};
            const int nblk = (N + 31) / 32;
            int grid = std::min(c->sms, std::max(2, (nblk - 1) * nblk / 2));
            ARGUS_CUDA_CHECK(cudaLaunchCooperativeKernel((void *)k_chol_coop, dim3(grid), dim3(kCoopThreads), args, 0, s));
        } else {
            if (N > 832) { fprintf(stderr, "argus_b200: %d unknowns in the reduced system need cooperative launches\n", N); return ARGUS_E_CUDA; }
            size_t rows = (size_t)(N > kCholNB ? N - kCholNB : 1);
            size_t sm = rows * (kCholNB + 1) * sizeof(double);
            if (sm < (size_t)N * sizeof(double)) sm = (size_t)N * sizeof(double);
            k_chol_solve2<<<1, kCholThreads, sm, s>>>(c->S.p, c->E.p, c->da_free.p, N, c->cholwork.p, c->flags.p + 1, done);
        }
    }
    {
        ProfScope ps(c, P_BACKSUB);
        k_cam_update<<<1, 256, 0, s>>>(m, mcon, c->pbufs(), c->st.p, c->da_free.p, c->red1, c->da_full.p, c->cam2.p);
        const size_t sm = 2 * cam_smem(m) + (size_t)m * 17 * sizeof(double);
        k_backsub_recompute<<<c->grid_pts128, 128, sm, s>>>(c->problem(), c->pbufs(), c->st.p, 1, c->da_full.p, c->Vinv.p, c->eb.p, dbout, c->part3.p);
        if (c->wand_n > 0) {
            k_wand_backsub<<<c->wand_grid, kWandThreads, sm, s>>>(c->problem(), c->wand_plan(), c->pbufs(), c->st.p, c->da_full.p, c->Vinv.p, c->eb.p,
                                                                   dbout, c->part3.p + 3 * (size_t)c->grid_pts128);
            c->launches += 1;
        }
        k_reduce_attempt<<<1, 32, 0, s>>>(c->part3.p, c->grid_pts128 + c->wand_grid, c->flags.p, c->att3.p, done);
    }
    c->launches += 5;
    ARGUS_CUDA_CHECK(cudaGetLastError());
    Red r3[1] = {{c->att3.p, 4, 0}};
    if ((rc = allreduce_many(c, r3, 1))) return rc;
    return controller ? enqueue_controller(c) : ARGUS_OK;
}

// ---- one attempt of the motion-only / structure-only drivers (sba_levmar.c:1838-1950, :2383-2495): block-diagonal solves ----
int enqueue_attempt_blockdiag(argus_ba_ctx *c)
{
    const int m = c->m;
    cudaStream_t s = c->stream;
    const int *done = c->done_ptr();
    int rc = enqueue_lin0(c);            // U / ea or V / eb do not depend on the damping: the attempt re-derives them at the current point
    if (rc) return rc;
    if (c->mode == kModeMot) {
        {
            ProfScope ps(c, P_CHOL);
            k_mot_solve<<<1, 64, 0, s>>>(m, c->mcon, c->red1, c->pbufs(), c->st.p, c->cam2.p, c->flags.p);
            c->launches += 1;
        }
        if ((rc = enqueue_cost(c, 1, nullptr))) return rc;
    } else {
        {
            ProfScope ps(c, P_BACKSUB);
            k_str_invert<<<(unsigned)((c->n + 255) / 256), 256, 0, s>>>(c->n, c->ncon, c->st.p, c->V.p, c->Vinv.p, c->flags.p);
            const size_t sm = 2 * cam_smem(m) + (size_t)m * 17 * sizeof(double);
            k_backsub_recompute<<<c->grid_pts128, 128, sm, s>>>(c->problem(), c->pbufs(), c->st.p, 0, c->da_full.p, c->Vinv.p, c->eb.p, nullptr, c->part3.p);
            k_reduce_attempt<<<1, 32, 0, s>>>(c->part3.p, c->grid_pts128, c->flags.p, c->att3.p, done);
            c->launches += 3;
        }
        ARGUS_CUDA_CHECK(cudaGetLastError());
        Red r3[1] = {{c->att3.p, 4, 0}};
        if ((rc = allreduce_many(c, r3, 1))) return rc;
    }
    return enqueue_controller(c);
}

int enqueue_attempt(argus_ba_ctx *c)
{
    return (c->mode == kModeMotStr) ? enqueue_attempt_motstr(c, true, nullptr) : enqueue_attempt_blockdiag(c);
}

// capture one attempt into a graph; on any failure the solver simply keeps launching kernels directly
void capture_attempt_graph(argus_ba_ctx *c, bool prof)
{
    cudaGraphExec_t &ge = prof ? c->graph_exec_prof : c->graph_exec;
    if (ge) { cudaGraphExecDestroy(ge); ge = nullptr; }
    if (prof) { for (auto &e : c->graph_evs) { c->ev_pool.push_back(e.a); c->ev_pool.push_back(e.b); } c->graph_evs.clear(); }
    const int64_t l0 = c->launches, a0 = c->n_allreduce;
    if (cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) { cudaGetLastError(); c->use_graph = false; return; }
    c->capturing = true; c->capture_prof = prof;
    const int rc = enqueue_attempt(c);
    c->capturing = false; c->capture_prof = false;
    cudaGraph_t g = nullptr;
    const cudaError_t e = cudaStreamEndCapture(c->stream, &g);
    c->graph_launches = c->launches - l0; c->graph_allreduces = c->n_allreduce - a0;
    c->launches = l0; c->n_allreduce = a0;
    if (rc != ARGUS_OK || e != cudaSuccess || !g) { if (g) cudaGraphDestroy(g); cudaGetLastError(); c->use_graph = false; return; }
    if (cudaGraphInstantiate(&ge, g, 0) != cudaSuccess) { cudaGetLastError(); ge = nullptr; c->use_graph = false; }
    cudaGraphDestroy(g);
}

// global problem size over the ranks: {projections, points, free points}
int global_counts(argus_ba_ctx *c, double out[3])
{
    out[0] = (double)c->nobs; out[1] = (double)c->n; out[2] = (double)(c->n - c->ncon);
    if (c->world <= 1) return ARGUS_OK;
    ARGUS_CUDA_CHECK(cudaMemcpyAsync(c->att3.p, out, 3 * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    ARGUS_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    Red r[1] = {{c->att3.p, 3, 0}};
    int rc = allreduce_many(c, r, 1);
    if (rc) return rc;
    ARGUS_CUDA_CHECK(cudaMemcpyAsync(out, c->att3.p, 3 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    ARGUS_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    c->d2h_bytes += 24;
    return ARGUS_OK;
}

}  // namespace

extern "C" {

#ifdef ARGUS_LIN_TRACE
int argus_debug_set_lin_trace(long long *dev_ptr) { return cudaMemcpyToSymbol(g_lin_trace, &dev_ptr, sizeof(dev_ptr)) == cudaSuccess ? 0 : -1; }
#endif
#ifdef ARGUS_SCHUR_TRACE
int argus_debug_set_schur_trace(long long *dev_ptr) { return cudaMemcpyToSymbol(g_schur_trace, &dev_ptr, sizeof(dev_ptr)) == cudaSuccess ? 0 : -1; }
#endif

const char *argus_b200_version(void) { return "argus_b200 0.2 (sm_100a, fp64)"; }

int argus_set_default_device(int device)
{
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) return ARGUS_E_ARG;
    g_default_device = device;
    return ARGUS_OK;
}

int argus_ba_create(argus_ba_ctx **out, int device, void *stream)
{
    if (!out) return ARGUS_E_ARG;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev) {
        fprintf(stderr, "argus_b200: no CUDA device available (there is no CPU fallback)\n");
        return ARGUS_E_CUDA;
    }
    ARGUS_CUDA_CHECK(cudaSetDevice(device));
    argus_ba_ctx *c = new argus_ba_ctx();
    c->device = device;
    memset(c->prof_ms, 0, sizeof(c->prof_ms)); memset(c->prof_n, 0, sizeof(c->prof_n));
    auto fail = [&](const char *what, cudaError_t e) {
        fprintf(stderr, "argus_b200: %s failed: %s\n", what, cudaGetErrorString(e));
        argus_ba_destroy(c);
        return ARGUS_E_CUDA;
    };
    cudaError_t e;
    if (stream) { c->stream = (cudaStream_t)stream; c->own_stream = false; }
    else {
        if ((e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)) != cudaSuccess) { c->stream = nullptr; return fail("cudaStreamCreate", e); }
        c->own_stream = true;
    }
    if ((e = cudaDeviceGetAttribute(&c->sms, cudaDevAttrMultiProcessorCount, device)) != cudaSuccess) return fail("cudaDeviceGetAttribute", e);
    { int coop = 0; cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, device); c->coop_ok = coop != 0; }
    c->h_snap = static_cast<LmState *>(pinned_take());
    if (!c->h_snap) return fail("cudaMallocHost", cudaErrorMemoryAllocation);
    for (int q = 0; q < kSnap; ++q)
        if ((e = cudaEventCreateWithFlags(&c->snap_ev[q], cudaEventDisableTiming)) != cudaSuccess) { c->snap_ev[q] = nullptr; return fail("cudaEventCreate", e); }
    {   // opt in to large dynamic shared memory: once per device and process, always to the worst case of each kernel
        std::lock_guard<std::mutex> lk(g_once_mutex);
        if (device < 64 && !(g_attr_done & (1ull << device))) {
            const int lin_max = (int)((((size_t)ARGUS_BA_MAX_CAMS * sizeof(CamConst) + 15) & ~(size_t)15) + (size_t)256 * (kGa + kBs) * sizeof(double) +
                                      (size_t)kTilePts * kPs * sizeof(double));
            e = cudaFuncSetAttribute(k_chol_solve2, cudaFuncAttributeMaxDynamicSharedMemorySize, 214000);
            if (e == cudaSuccess) e = cudaFuncSetAttribute(k_schur_mma<1, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * (int)schur_stage_bytes(8));
            if (e == cudaSuccess) e = cudaFuncSetAttribute(k_schur_mma<3, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * (int)schur_stage_bytes(24));
            if (e == cudaSuccess) e = cudaFuncSetAttribute(k_schur_mma<5, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * (int)schur_stage_bytes(40));
            if (e == cudaSuccess) e = cudaFuncSetAttribute(k_schur_mma<6, 12>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * (int)schur_stage_bytes(72));
            if (e == cudaSuccess) e = cudaFuncSetAttribute(k_wand_schur, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wand_schur_smem(64));
            if (e == cudaSuccess) e = cudaFuncSetAttribute(k_chol_cluster, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)chol_cluster_smem());
            if (e == cudaSuccess) e = cudaFuncSetAttribute(k_lin_tile<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, lin_max);
            if (e == cudaSuccess) e = cudaFuncSetAttribute(k_lin_tile<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, lin_max);
            if (e == cudaSuccess) e = cudaFuncSetAttribute(k_lin_tile<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, lin_max);
            if (e != cudaSuccess) return fail("cudaFuncSetAttribute", e);
            g_attr_done |= 1ull << device;
        }
    }
    {   // can one cluster of kClCtas CTAs of the reduced-system solver be resident here (a partitioned GPU may say no)?
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(kClCtas); cfg.blockDim = dim3(kClThreads); cfg.dynamicSmemBytes = chol_cluster_smem();
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = kClCtas; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        int ncl = 0;
        if (cudaOccupancyMaxActiveClusters(&ncl, (void *)k_chol_cluster, &cfg) != cudaSuccess) { cudaGetLastError(); ncl = 0; }
        c->cluster_ok = ncl > 0;
    }
    *out = c;
    return ARGUS_OK;
}

void argus_ba_destroy(argus_ba_ctx *c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    drop_graph(c);
    if (c->nccl && c->nccl_owned) argus_nccl::api().CommDestroy(c->nccl);
    c->obs_ptr.release(); c->obs_cam.release(); c->obs_uv.release(); c->pt_mask.release();
    DevBuf<double> *bufs[] = {&c->rot0, &c->camb[0], &c->camb[1], &c->ptsb[0], &c->ptsb[1], &c->cam0, &c->pts0, &c->V, &c->eb, &c->vstate, &c->Vinv,
                              &c->dbuf, &c->scpart, &c->red12, &c->max2, &c->Spart, &c->S, &c->E, &c->da_free, &c->da_full, &c->part3, &c->att3,
                              &c->cam2, &c->scal, &c->respart, &c->Upart2, &c->cholwork, &c->Z};
    for (auto *b : bufs) b->release();
    c->corder.release(); c->row_info.release(); c->row_uv.release(); c->pt_slot0.release(); c->tile_r.release(); c->zs.release();
    c->hit_blob.release(); c->hit_ptr.release(); c->schur_pace.release(); c->cam_fix.release(); c->cstart.release(); c->vmask_dev.release();
    c->idx_btot.release(); c->flags.release(); c->tile_pt.release(); c->tile_o.release(); c->blk_jk.release(); c->blk_pair.release();
    c->zsign.release(); c->st.release();
    c->wand_ia.release(); c->wand_ib.release(); c->wand_of.release(); c->wand_len.release(); c->wand_rec.release(); c->wand_Q.release(); c->wand_part.release();
    if (c->lmlog) cudaFree(c->lmlog);
    for (auto e : c->ev_pool) cudaEventDestroy(e);
    for (auto &e : c->evs) { cudaEventDestroy(e.a); cudaEventDestroy(e.b); }
    for (int q = 0; q < kSnap; ++q) if (c->snap_ev[q]) cudaEventDestroy(c->snap_ev[q]);
    arena_give(c->device, c->arena.base, c->arena.size);
    if (c->copy_stream) { cudaStreamDestroy(c->copy_stream); cudaEventDestroy(c->copy_ev[0]); cudaEventDestroy(c->copy_ev[1]); }
    if (c->h_snap) pinned_give(c->h_snap);
    if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

int argus_ba_set_comm(argus_ba_ctx *c, int rank, int world, argus_allreduce_fn fn, void *user)
{
    if (!c || world < 1 || rank < 0 || rank >= world || (world > 1 && !fn)) return ARGUS_E_ARG;
    c->rank = rank; c->world = world; c->allreduce = fn; c->comm_user = user;
    drop_graph(c);
    return ARGUS_OK;
}

int argus_nccl_unique_id(void *out128)
{
    if (!out128) return ARGUS_E_ARG;
    argus_nccl::Api &a = argus_nccl::api();
    if (!a.ok) return ARGUS_E_COMM;
    argus_nccl::ncclUniqueId id;
    if (a.GetUniqueId(&id) != argus_nccl::ncclSuccess) return ARGUS_E_COMM;
    memcpy(out128, &id, sizeof(id));
    return ARGUS_OK;
}

void *argus_ba_get_comm_nccl(const argus_ba_ctx *c) { return c ? (void *)c->nccl : nullptr; }

int argus_ba_set_comm_nccl(argus_ba_ctx *c, int rank, int world, void *nccl_comm)
{
    if (!c || world < 1 || rank < 0 || rank >= world || (world > 1 && !nccl_comm)) return ARGUS_E_ARG;
    if (world > 1 && !argus_nccl::api().ok) return ARGUS_E_COMM;
    if (c->nccl && c->nccl_owned) argus_nccl::api().CommDestroy(c->nccl);
    c->rank = rank; c->world = world; c->nccl = (argus_nccl::ncclComm_t)nccl_comm; c->nccl_owned = false; c->allreduce = nullptr;
    drop_graph(c);
    return ARGUS_OK;
}

int argus_ba_comm_init_nccl(argus_ba_ctx *c, int rank, int world, const void *unique_id128)
{
    if (!c || world < 1 || rank < 0 || rank >= world || !unique_id128) return ARGUS_E_ARG;
    argus_nccl::Api &a = argus_nccl::api();
    if (!a.ok) return ARGUS_E_COMM;
    ARGUS_CUDA_CHECK(cudaSetDevice(c->device));
    argus_nccl::ncclUniqueId id;
    memcpy(&id, unique_id128, sizeof(id));
    argus_nccl::ncclComm_t comm = nullptr;
    const int rc = a.CommInitRank(&comm, world, id, rank);
    if (rc != argus_nccl::ncclSuccess) {
        fprintf(stderr, "argus_b200: ncclCommInitRank failed: %s\n", a.GetErrorString ? a.GetErrorString(rc) : "?");
        return ARGUS_E_COMM;
    }
    if (c->nccl && c->nccl_owned) a.CommDestroy(c->nccl);
    c->rank = rank; c->world = world; c->nccl = comm; c->nccl_owned = true; c->allreduce = nullptr;
    drop_graph(c);
    return ARGUS_OK;
}

int64_t argus_ba_num_observations(const argus_ba_ctx *c) { return c ? c->nobs : 0; }

int argus_ba_upload(argus_ba_ctx *c, int64_t n, int64_t ncon, int m, int mcon, const char *vmask, const double *p,
                    const double *rot0, const double *x, int nccalib, int ncdist)
{
    if (!c || !vmask || !p || !rot0 || !x) return ARGUS_E_ARG;
    if (m < 1 || m > ARGUS_BA_MAX_CAMS || n < 0 || ncon < 0 || ncon > n || mcon < 0 || mcon >= m) return ARGUS_E_ARG;
    if (nccalib < 0 || nccalib > 5 || ncdist < 0 || ncdist > 5) return ARGUS_E_ARG;
    ARGUS_CUDA_CHECK(cudaSetDevice(c->device));
    drop_graph(c);
    // visibility: the host only counts (observations per point, for the sizes and the tile plan); the CSR and the per-point
    // camera bit masks are built on the device from the raw mask bytes (k_index_*; sba_levmar.c:639-650 builds its CRS on the host)
    std::vector<uint8_t> cnt((size_t)n);
    parallel_for(n, c->world, [&](int64_t lo, int64_t hi) {
        for (int64_t i = lo; i < hi; ++i) {
            const char *row = vmask + i * m;
            int k = 0;
            for (int j = 0; j < m; ++j) k += row[j] != 0;
            cnt[i] = (uint8_t)k;
        }
    });
    c->n = n; c->ncon = ncon; c->m = m; c->mcon = mcon; c->nfix_k = nccalib; c->nfix_d = ncdist; c->cam_fix_on = false;
    c->wand_n = 0; c->wand_grid = 0; c->wand_chunks = 0;       // pairs refer to the points of one upload
    const int mf = m - mcon, N = 16 * mf;
    c->npairs = mf * (mf + 1) / 2;
    c->grid_pts128 = (int)std::min<int64_t>((n + 127) / 128, (int64_t)c->sms * 4);   /* one wave: 4 CTAs of 128 threads x 120 registers per SM, one camera prologue each */ if (c->grid_pts128 < 1) c->grid_pts128 = 1;
    c->grid_pts256 = (int)std::min<int64_t>((n + 255) / 256, (int64_t)c->sms * 8); if (c->grid_pts256 < 1) c->grid_pts256 = 1;
    // ---- plans: point tiles (<= 32 points; every point gets one of four classes, a class holds <= 64 observations per tile;
    //      the observation of rank r of a point sits in Z row 4 (first index inside the class + r) + class, see ba_schur_mma.cuh)
    //      and the block -> (group, warp, slot) map ----
    std::vector<int32_t> tiles, tile_o, tile_r;
    std::vector<uint8_t> slot0((size_t)n);
    int64_t nobs = 0, nrows = 0, hits_bound = 0;
    {
        // the points are cut into segments of a FIXED length (so that the tiling, and with it every summation order, does not
        // depend on how many host threads there are); the segments are planned in parallel, a tile never spans two of them
        constexpr int64_t kSeg = 32768;
        const int64_t nseg = (n + kSeg - 1) / kSeg;
        struct SegPlan { std::vector<int32_t> first, obs, rows; int64_t hits = 0; };
        std::vector<SegPlan> seg((size_t)nseg);
        parallel_for(nseg, c->world, [&](int64_t lo, int64_t hi) {
            for (int64_t sgi = lo; sgi < hi; ++sgi) {
                SegPlan &sp = seg[(size_t)sgi];
                const int64_t end = std::min(n, (sgi + 1) * kSeg);
                int64_t i = sgi * kSeg;
                while (i < end) {
                    sp.first.push_back((int32_t)i);
                    int fill[4] = {0, 0, 0, 0}, npts = 0, acc = 0;
                    while (i < end && npts < kTilePts) {
                        const int v = cnt[i];
                        int cl = 0;
                        for (int q = 1; q < 4; ++q) if (fill[q] < fill[cl]) cl = q;        // emptiest class, lowest id on ties
                        if (fill[cl] + v > kClassCap) break;
                        slot0[i] = (uint8_t)(4 * fill[cl] + cl);
                        fill[cl] += v; acc += v; ++npts; ++i;
                        sp.hits += (int64_t)v * (v + 1) / 2;
                    }
                    sp.obs.push_back(acc);
                    sp.rows.push_back(4 * std::max(std::max(fill[0], fill[1]), std::max(fill[2], fill[3])));
                }
            }
        }, 1);
        size_t nt = 0;
        for (auto &sp : seg) nt += sp.first.size();
        tiles.reserve(nt + 1); tile_o.reserve(nt + 1); tile_r.reserve(nt + 1);
        for (auto &sp : seg) {
            for (size_t t = 0; t < sp.first.size(); ++t) {
                tiles.push_back(sp.first[t]); tile_o.push_back((int32_t)nobs); tile_r.push_back((int32_t)nrows);
                nobs += sp.obs[t]; nrows += sp.rows[t];
                if (nrows > 0x7fffffffLL || nobs > 0x7fffffffLL) return ARGUS_E_ARG;
            }
            hits_bound += sp.hits;
        }
        tiles.push_back((int32_t)n); tile_o.push_back((int32_t)nobs); tile_r.push_back((int32_t)nrows);
        c->ntiles = (int)tiles.size() - 1;
    }
    c->nobs = nobs; c->nrows = nrows;
    {
        const int cand[3] = {1, 3, 5};
        int NB = 6, NW = 12;
        for (int q = 0; q < 3; ++q) if (kSchurWarps * cand[q] >= c->npairs) { NB = cand[q]; NW = 8; break; }
        c->schur_NB = NB; c->schur_NW = NW;
        c->schur_G = (c->npairs + NW * NB - 1) / (NW * NB);
        int nch = c->sms / c->schur_G; if (nch < 1) nch = 1;
        if (nch > c->ntiles) nch = c->ntiles > 0 ? c->ntiles : 1;
        c->schur_nchunk = nch;
        // tiles a group CTA may run ahead of the slowest group of its chunk: as many as a QUARTER of the L2 holds over all chunks
        // (measured: with half of it - 9 tiles at config 4 - the second CTA misses again; the L2 is two partitions and lines
        // read from the far one are held twice)
        int l2 = 0; cudaDeviceGetAttribute(&l2, cudaDevAttrL2CacheSize, c->device);
        const double tile_bytes = 384.0 * kTileRows;
        int slack = (int)(0.25 * (double)l2 / ((double)nch * tile_bytes));
        c->pace_slack = slack < 2 ? 0 : (slack > 64 ? 64 : slack);
    }
    // corder, cstart, row_info / row_uv and the hit lists are derived on the device (k_index_*, k_build_tile_index, k_hits_*)
    c->lin_smem = (((size_t)m * sizeof(CamConst) + 15) & ~(size_t)15) + (size_t)kTileRows * (kGa + kBs) * sizeof(double) +
                  (size_t)kTilePts * kPs * sizeof(double);
    {
        int occ = (int)((220 * 1024) / (c->lin_smem + 1024)); if (occ < 1) occ = 1; if (occ > 2) occ = 2;
        c->grid_lin = std::min(c->ntiles > 0 ? c->ntiles : 1, c->sms * occ);
    }
    // block -> (group, warp, slot): filled further down, once the device has counted the hits of every camera pair
    std::vector<uint16_t> hjk((size_t)c->schur_G * c->schur_NW * c->schur_NB, 0xFFFF);
    std::vector<int32_t> hpair((size_t)c->schur_G * c->schur_NW * c->schur_NB, -1);
    const int64_t slen = (int64_t)c->npairs * 256;
    c->r1len = ((int64_t)m * kU + 2 + 1) & ~(int64_t)1;
    // every device buffer of the problem comes out of ONE allocation: measure, allocate, assign
    auto plan = [&]() -> cudaError_t {
#define PLAN_CHECK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return e_; } while (0)
#define PLAN_ENS(buf, cnt_) PLAN_CHECK((buf).ensure((size_t)(cnt_)))
        PLAN_ENS(c->tile_pt, tiles.size()); PLAN_ENS(c->tile_o, tiles.size()); PLAN_ENS(c->tile_r, tiles.size()); PLAN_ENS(c->blk_jk, hjk.size());
        PLAN_ENS(c->blk_pair, hpair.size()); PLAN_ENS(c->pt_slot0, n); PLAN_ENS(c->row_info, nrows); PLAN_ENS(c->row_uv, nrows); PLAN_ENS(c->zs, (size_t)c->ntiles * 256);
        PLAN_ENS(c->zsign, n);
        if (c->mode == kModeMotStr) {
            PLAN_ENS(c->Z, 48 * (nrows + c->ntiles));          // one padding chunk per plane and tile (z_index)
            PLAN_ENS(c->hit_ptr, (size_t)c->ntiles * c->schur_G + 1);
            PLAN_ENS(c->hit_blob, hits_bound + (2 * (int64_t)(c->schur_NW * c->schur_NB + 4) + 8) * c->ntiles * c->schur_G + 8);
        }
        PLAN_ENS(c->corder, nobs); PLAN_ENS(c->cstart, (size_t)c->ntiles * (m + 1));
        PLAN_ENS(c->Upart2, (int64_t)c->grid_lin * m * kUfrag); PLAN_ENS(c->cholwork, (int64_t)((N + 31) / 32) * 1056 + N + 64);
        PLAN_ENS(c->vmask_dev, n * m); PLAN_ENS(c->idx_btot, (n + kIdxPts - 1) / kIdxPts + 1);
        PLAN_ENS(c->obs_ptr, n + 1); PLAN_ENS(c->obs_cam, nobs); PLAN_ENS(c->obs_uv, nobs); PLAN_ENS(c->pt_mask, n);
        PLAN_ENS(c->rot0, 4 * m); PLAN_ENS(c->camb[0], 16 * m); PLAN_ENS(c->camb[1], 16 * m); PLAN_ENS(c->cam0, 16 * m);
        PLAN_ENS(c->ptsb[0], 3 * n); PLAN_ENS(c->ptsb[1], 3 * n); PLAN_ENS(c->pts0, 3 * n);
        PLAN_ENS(c->V, 6 * n); PLAN_ENS(c->eb, 3 * n); PLAN_ENS(c->vstate, 3 * n); PLAN_ENS(c->Vinv, 6 * n); PLAN_ENS(c->dbuf, 3 * n);
        PLAN_ENS(c->scpart, 4 * (c->grid_lin + kWandGridMax)); PLAN_ENS(c->red12, c->r1len + slen + 16 * m); PLAN_ENS(c->max2, 2);
        PLAN_ENS(c->Spart, (int64_t)c->schur_nchunk * slen); PLAN_ENS(c->schur_pace, (size_t)c->schur_nchunk * kPaceInts);
        PLAN_ENS(c->S, (int64_t)N * N); PLAN_ENS(c->E, N); PLAN_ENS(c->da_free, N); PLAN_ENS(c->da_full, 16 * m);
        PLAN_ENS(c->part3, 3 * (c->grid_pts128 + kWandGridMax)); PLAN_ENS(c->att3, 4); PLAN_ENS(c->cam2, 2); PLAN_ENS(c->scal, 8); PLAN_ENS(c->respart, c->grid_pts256 + kWandGridMax);
        PLAN_ENS(c->flags, 4); PLAN_ENS(c->st, 1);
#undef PLAN_ENS
#undef PLAN_CHECK
        return cudaSuccess;
    };
    {
        Arena &ar = c->arena;
        ar.measuring = true; ar.used = 0;
        g_arena = &ar;
        cudaError_t pe = plan();
        const size_t need = ar.used + 256;
        if (pe == cudaSuccess && need > ar.size) {
            arena_give(c->device, ar.base, ar.size);
            ar.base = nullptr; ar.size = 0;
            size_t got = 0;
            ar.base = arena_take(c->device, need, &got);
            if (ar.base) ar.size = got; else pe = cudaErrorMemoryAllocation;
        }
        if (pe == cudaSuccess) { ar.measuring = false; ar.used = 0; pe = plan(); }
        g_arena = nullptr;
        ARGUS_CUDA_CHECK(pe);
    }
    c->red1 = c->red12.p; c->red2 = c->red12.p + c->r1len;
    cudaStream_t s = c->stream;
    ARGUS_CUDA_CHECK(cudaMemsetAsync(c->red12.p, 0, (size_t)(c->r1len + slen + 16 * m) * sizeof(double), s));
    ARGUS_CUDA_CHECK(cudaMemsetAsync(c->da_full.p, 0, 16 * m * sizeof(double), s));
    ARGUS_CUDA_CHECK(cudaMemsetAsync(c->cam2.p, 0, 2 * sizeof(double), s));
    ARGUS_CUDA_CHECK(cudaMemsetAsync(c->st.p, 0, sizeof(LmState), s));
    ARGUS_CUDA_CHECK(cudaMemcpyAsync(c->tile_pt.p, tiles.data(), tiles.size() * sizeof(int32_t), cudaMemcpyHostToDevice, s));
    ARGUS_CUDA_CHECK(cudaMemcpyAsync(c->tile_o.p, tile_o.data(), tile_o.size() * sizeof(int32_t), cudaMemcpyHostToDevice, s));
    ARGUS_CUDA_CHECK(cudaMemcpyAsync(c->tile_r.p, tile_r.data(), tile_r.size() * sizeof(int32_t), cudaMemcpyHostToDevice, s));
    if (n > 0) ARGUS_CUDA_CHECK(cudaMemcpyAsync(c->pt_slot0.p, slot0.data(), (size_t)n, cudaMemcpyHostToDevice, s));
    ARGUS_CUDA_CHECK(cudaMemcpyAsync(c->vmask_dev.p, vmask, (size_t)n * m, cudaMemcpyHostToDevice, s));
    if (n > 0) {
        const int nb = (int)((n + kIdxPts - 1) / kIdxPts);
        k_index_count<<<nb, 256, 0, s>>>(n, m, c->vmask_dev.p, c->pt_mask.p, c->idx_btot.p);
        k_index_scan<<<1, 1024, 0, s>>>(nb, c->idx_btot.p);
        k_index_fill<<<nb, 256, 0, s>>>(n, c->pt_mask.p, c->idx_btot.p, c->obs_ptr.p, c->obs_cam.p);
    } else {
        ARGUS_CUDA_CHECK(cudaMemsetAsync(c->obs_ptr.p, 0, sizeof(int32_t), s));
    }
    // the two large host arrays (measurements, points: 368 of the 400 MB at config 4) travel on a second stream while the
    // visibility index, the block assignment and the hit lists - which need neither - are built; k_build_tile_index waits for them
    if (!c->copy_stream) {
        ARGUS_CUDA_CHECK(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
        ARGUS_CUDA_CHECK(cudaEventCreateWithFlags(&c->copy_ev[0], cudaEventDisableTiming));
        ARGUS_CUDA_CHECK(cudaEventCreateWithFlags(&c->copy_ev[1], cudaEventDisableTiming));
    }
    // (the small copies first: one copy engine serves both streams in issue order, and the solver's stream is synchronised for
    // the hit histogram long before the large ones are through)
    ARGUS_CUDA_CHECK(cudaMemcpyAsync(c->rot0.p, rot0, 4 * m * sizeof(double), cudaMemcpyHostToDevice, s));
    ARGUS_CUDA_CHECK(cudaMemcpyAsync(c->cam0.p, p, 16 * m * sizeof(double), cudaMemcpyHostToDevice, s));
    ARGUS_CUDA_CHECK(cudaEventRecord(c->copy_ev[0], s));                       // (earlier work on the solver's stream may still use the buffers)
    ARGUS_CUDA_CHECK(cudaStreamWaitEvent(c->copy_stream, c->copy_ev[0], 0));
    ARGUS_CUDA_CHECK(cudaMemcpyAsync(c->obs_uv.p, x, nobs * sizeof(double2), cudaMemcpyHostToDevice, c->copy_stream));
    ARGUS_CUDA_CHECK(cudaMemcpyAsync(c->pts0.p, p + 16 * m, 3 * n * sizeof(double), cudaMemcpyHostToDevice, c->copy_stream));
    ARGUS_CUDA_CHECK(cudaEventRecord(c->copy_ev[1], c->copy_stream));
    {
        // Every warp of the Schur kernel keeps its blocks for the whole launch and the warps of a CTA advance tile by tile
        // together (two shared-memory stages), so the slowest warp sets the pace: deal the blocks by weight.  Weight = hits of
        // the camera pair over all points (counted on the device) - a diagonal block needs 3 of the 4 sub-tiles - plus a fixed
        // cost per tile; longest-processing-time-first into the G x NW warps, each with NB slots.
        std::vector<unsigned long long> hist((size_t)m * m, 0ull);
        if (n > ncon && c->mode == kModeMotStr) {
            unsigned long long *dh = reinterpret_cast<unsigned long long *>(c->S.p);      // scratch: S (N x N doubles >= m x m) is not in use yet
            if ((size_t)N * N >= (size_t)m * m) {
                ARGUS_CUDA_CHECK(cudaMemsetAsync(dh, 0, (size_t)m * m * sizeof(unsigned long long), s));
                const int gh = (int)std::min<int64_t>((n - ncon + 255) / 256, (int64_t)c->sms * 4);
                k_pair_hist<<<gh, 256, (size_t)m * m * sizeof(unsigned int), s>>>(n, ncon, m, c->pt_mask.p, dh);
                ARGUS_CUDA_CHECK(cudaMemcpyAsync(hist.data(), dh, hist.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
                ARGUS_CUDA_CHECK(cudaStreamSynchronize(s));
            }
        }
        struct Blk { int a, b, pidx; double w; };
        std::vector<Blk> blks;
        int pidx = 0;
        for (int a = 0; a < mf; ++a)
            for (int b = a; b < mf; ++b, ++pidx) {
                const double h = (double)hist[(size_t)(mcon + a) * m + (mcon + b)];
                blks.push_back({a, b, pidx, h * (a == b ? 0.75 : 1.0) + 1.5 * (double)c->ntiles});
            }
        std::stable_sort(blks.begin(), blks.end(), [](const Blk &x, const Blk &y) { return x.w > y.w; });
        const int nbins = c->schur_G * c->schur_NW, NBs = c->schur_NB;
        std::vector<double> load((size_t)nbins, 0.0);
        std::vector<std::vector<int>> bins((size_t)nbins);
        // warp 0 of every CTA also issues the tile copies (about 1200 clocks per tile, measured with tools/schur_trace.py: the
        // time of ~6 hits); without this head start the other warps wait for it at every tile
        for (int q = 0; q < nbins; q += c->schur_NW) load[q] = 6.0 * (double)c->ntiles;
        for (int i = 0; i < (int)blks.size(); ++i) {
            int best = -1;
            for (int q = 0; q < nbins; ++q)
                if ((int)bins[q].size() < NBs && (best < 0 || load[q] < load[best])) best = q;
            bins[best].push_back(i); load[best] += blks[i].w;
        }
        // local search on the heaviest warp: the slot limit makes plain LPT leave it up to 10 % above what a move or a swap reaches
        for (int iter = 0; iter < 4096; ++iter) {
            int bm = 0;
            for (int q = 1; q < nbins; ++q) if (load[q] > load[bm]) bm = q;
            double bestv = load[bm]; int bi = -1, bc = -1, bj = -1;
            for (int ii = 0; ii < (int)bins[bm].size(); ++ii) {
                const double wi = blks[bins[bm][ii]].w;
                for (int q = 0; q < nbins; ++q) {
                    if (q == bm) continue;
                    if ((int)bins[q].size() < NBs) {
                        const double v = std::max(load[bm] - wi, load[q] + wi);
                        if (v < bestv * (1.0 - 1e-12)) { bestv = v; bi = ii; bc = q; bj = -1; }
                    }
                    for (int jj = 0; jj < (int)bins[q].size(); ++jj) {
                        const double wj = blks[bins[q][jj]].w;
                        const double v = std::max(load[bm] - wi + wj, load[q] + wi - wj);
                        if (v < bestv * (1.0 - 1e-12)) { bestv = v; bi = ii; bc = q; bj = jj; }
                    }
                }
            }
            if (bi < 0) break;
            const int i = bins[bm][bi];
            if (bj < 0) {
                bins[bm].erase(bins[bm].begin() + bi); bins[bc].push_back(i);
                load[bm] -= blks[i].w; load[bc] += blks[i].w;
            } else {
                const int j = bins[bc][bj];
                bins[bm][bi] = j; bins[bc][bj] = i;
                load[bm] += blks[j].w - blks[i].w; load[bc] += blks[i].w - blks[j].w;
            }
        }
        for (int q = 0; q < nbins; ++q)
            for (int t = 0; t < (int)bins[q].size(); ++t) {
                const Blk &bk = blks[bins[q][t]];
                const size_t at = (size_t)q * NBs + t;
                hjk[at] = (uint16_t)((mcon + bk.a) | ((mcon + bk.b) << 8));
                hpair[at] = bk.pidx;
            }
        ARGUS_CUDA_CHECK(cudaMemcpyAsync(c->blk_jk.p, hjk.data(), hjk.size() * sizeof(uint16_t), cudaMemcpyHostToDevice, s));
        ARGUS_CUDA_CHECK(cudaMemcpyAsync(c->blk_pair.p, hpair.data(), hpair.size() * sizeof(int32_t), cudaMemcpyHostToDevice, s));
    }
    // derived static indices, built on the device
    if (n > 0 && c->ntiles > 0) {
        ARGUS_CUDA_CHECK(cudaMemsetAsync(c->row_info.p, 0xFF, (size_t)nrows * sizeof(uint32_t), s));
        ARGUS_CUDA_CHECK(cudaMemsetAsync(c->row_uv.p, 0, (size_t)nrows * sizeof(double2), s));
        ARGUS_CUDA_CHECK(cudaMemsetAsync(c->zs.p, 0, (size_t)c->ntiles * 256, s));
        if (c->mode == kModeMotStr) {
            const int G = c->schur_G, nslots = c->schur_NW * c->schur_NB, ntg = c->ntiles * G;
            k_hits_count<<<c->ntiles, 256, (size_t)G * sizeof(int), s>>>(c->ntiles, G, nslots, c->tile_pt.p, ncon, c->pt_mask.p, c->pt_slot0.p,
                                                                         c->blk_jk.p, c->hit_ptr.p);
            k_index_scan<<<1, 1024, 0, s>>>(ntg, c->hit_ptr.p);
            k_hits_fill<<<c->ntiles, 256, (size_t)G * nslots * sizeof(int), s>>>(c->ntiles, G, nslots, c->tile_pt.p, ncon, c->pt_mask.p,
                                                                               c->pt_slot0.p, c->blk_jk.p, c->tile_r.p, c->hit_ptr.p, c->hit_blob.p);
            // Z itself needs no initialisation: the lineariser writes every row that holds an observation before the Schur
            // kernel runs, the hit lists name only such rows, and the unused rows and the padding chunk of every plane travel to
            // shared memory without ever being read from there
        }
    }
    ARGUS_CUDA_CHECK(cudaStreamWaitEvent(s, c->copy_ev[1], 0));               // measurements and points have arrived
    if (n > 0 && c->ntiles > 0) {
        k_build_tile_index<<<c->ntiles, 256, 0, s>>>(m, c->tile_pt.p, c->tile_o.p, c->tile_r.p, c->obs_ptr.p, c->obs_cam.p, c->obs_uv.p, c->pt_slot0.p,
                                                     c->corder.p, c->cstart.p, c->row_info.p, c->row_uv.p);
        ARGUS_CUDA_CHECK(cudaGetLastError());
    }
    ARGUS_CUDA_CHECK(cudaStreamSynchronize(s));      // host staging vectors go out of scope
    c->h2d_bytes += (size_t)n * m + nobs * sizeof(double2) + (4 + 16) * m * sizeof(double) + 3 * n * sizeof(double) +
                    3 * tiles.size() * sizeof(int32_t) + (size_t)n;
    return argus_ba_reset(c);
}

int argus_ba_set_mode(argus_ba_ctx *c, int mode)
{
    if (!c || mode < kModeMotStr || mode > kModeStr) return ARGUS_E_ARG;
    if (c->m != 0 && mode != c->mode && (mode == kModeMotStr) != (c->mode == kModeMotStr)) return ARGUS_E_ARG;   // set it before upload
    if (mode != c->mode) drop_graph(c);
    c->mode = mode;
    return ARGUS_OK;
}

int argus_ba_set_camera_fix(argus_ba_ctx *c, const int *nccalib, const int *ncdist)
{
    if (!c || c->m == 0) return ARGUS_E_ARG;
    ARGUS_CUDA_CHECK(cudaSetDevice(c->device));
    drop_graph(c);
    if (!nccalib || !ncdist) { c->cam_fix_on = false; return ARGUS_OK; }
    std::vector<uint8_t> h(2 * (size_t)c->m);
    for (int j = 0; j < c->m; ++j) {
        if (nccalib[j] < 0 || nccalib[j] > 5 || ncdist[j] < 0 || ncdist[j] > 5) return ARGUS_E_ARG;
        h[2 * j] = (uint8_t)nccalib[j]; h[2 * j + 1] = (uint8_t)ncdist[j];
    }
    ARGUS_CUDA_CHECK(c->cam_fix.ensure(h.size()));
    ARGUS_CUDA_CHECK(cudaMemcpy(c->cam_fix.p, h.data(), h.size(), cudaMemcpyHostToDevice));
    c->cam_fix_on = true;
    return ARGUS_OK;
}

int argus_ba_set_wand(argus_ba_ctx *c, int64_t npairs, const int64_t *ia, const int64_t *ib, const double *length, double weight)
{
    if (!c || c->m == 0 || npairs < 0 || (npairs > 0 && (!ia || !ib || !length)) || !(weight >= 0.0)) return ARGUS_E_ARG;
    ARGUS_CUDA_CHECK(cudaSetDevice(c->device));
    drop_graph(c);
    c->wand_n = 0; c->wand_grid = 0; c->wand_chunks = 0; c->wand_w = weight;
    if (npairs == 0 || weight == 0.0) return ARGUS_OK;          // off: the solver is the unconstrained one, bit for bit
    if (c->mode != kModeMotStr || npairs > 0x7fffffffLL) return ARGUS_E_ARG;
    std::vector<int32_t> of((size_t)c->n, -1), ha((size_t)npairs), hb((size_t)npairs);
    for (int64_t p = 0; p < npairs; ++p) {
        const int64_t a = ia[p], b = ib[p];
        if (a < 0 || b < 0 || a >= c->n || b >= c->n || a == b || of[a] >= 0 || of[b] >= 0 || !(length[p] > 0.0)) return ARGUS_E_ARG;
        of[a] = (int32_t)p; of[b] = (int32_t)p; ha[p] = (int32_t)a; hb[p] = (int32_t)b;
    }
    const int grid = (int)std::min<int64_t>(kWandGridMax, (npairs + kWandThreads - 1) / kWandThreads);
    const int nsub = (c->npairs + kWandBlocks - 1) / kWandBlocks;
    const int chunks = (int)std::max<int64_t>(1, std::min<int64_t>(std::max(1, 3 * c->sms / nsub) /* one wave at 3 CTAs per SM */, (npairs + 4 * wand_batch(c->m) - 1) / (4 * wand_batch(c->m))));
    ARGUS_CUDA_CHECK(c->wand_ia.ensure(npairs)); ARGUS_CUDA_CHECK(c->wand_ib.ensure(npairs)); ARGUS_CUDA_CHECK(c->wand_of.ensure(c->n));
    ARGUS_CUDA_CHECK(c->wand_len.ensure(npairs)); ARGUS_CUDA_CHECK(c->wand_rec.ensure((size_t)npairs * kWandRec));
    ARGUS_CUDA_CHECK(c->wand_Q.ensure((size_t)npairs * c->m * 16));
    ARGUS_CUDA_CHECK(c->wand_part.ensure((size_t)chunks * ((size_t)c->npairs * 256 + 16 * (size_t)c->m)));
    ARGUS_CUDA_CHECK(cudaMemcpy(c->wand_ia.p, ha.data(), ha.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
    ARGUS_CUDA_CHECK(cudaMemcpy(c->wand_ib.p, hb.data(), hb.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
    ARGUS_CUDA_CHECK(cudaMemcpy(c->wand_of.p, of.data(), of.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
    ARGUS_CUDA_CHECK(cudaMemcpy(c->wand_len.p, length, (size_t)npairs * sizeof(double), cudaMemcpyHostToDevice));
    c->h2d_bytes += (size_t)npairs * (2 * sizeof(int32_t) + sizeof(double)) + (size_t)c->n * sizeof(int32_t);
    c->wand_n = npairs; c->wand_grid = grid; c->wand_chunks = chunks;
    return ARGUS_OK;
}

int argus_ba_reset(argus_ba_ctx *c)
{
    if (!c || c->m == 0) return ARGUS_E_ARG;
    ARGUS_CUDA_CHECK(cudaSetDevice(c->device));
    c->cur_cam = 0; c->cur_pts = 0;
    ARGUS_CUDA_CHECK(cudaMemcpyAsync(c->camb[0].p, c->cam0.p, 16 * c->m * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
    ARGUS_CUDA_CHECK(cudaMemcpyAsync(c->ptsb[0].p, c->pts0.p, 3 * c->n * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
    return ARGUS_OK;
}

int argus_ba_download(argus_ba_ctx *c, double *p)
{
    if (!c || !p) return ARGUS_E_ARG;
    ARGUS_CUDA_CHECK(cudaSetDevice(c->device));
    ARGUS_CUDA_CHECK(cudaMemcpyAsync(p, c->camb[c->cur_cam].p, 16 * c->m * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    ARGUS_CUDA_CHECK(cudaMemcpyAsync(p + 16 * c->m, c->ptsb[c->cur_pts].p, 3 * c->n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    ARGUS_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    c->d2h_bytes += (16 * c->m + 3 * c->n) * sizeof(double);
    return ARGUS_OK;
}

int argus_ba_get_counters(argus_ba_ctx *c, int64_t out[4])
{
    if (!c || !out) return ARGUS_E_ARG;
    out[0] = c->launches; out[1] = c->n_allreduce; out[2] = c->h2d_bytes; out[3] = c->d2h_bytes;
    return ARGUS_OK;
}

int argus_ba_get_profile(argus_ba_ctx *c, const char **names, double *ms, int64_t *launches, int cap)
{
    if (!c) return 0;
    int k = 0;
    for (int i = 0; i < P_COUNT && k < cap; ++i, ++k) { names[k] = kProfNames[i]; ms[k] = c->prof_ms[i]; launches[k] = c->prof_n[i]; }
    return k;
}

int argus_ba_set_graph(argus_ba_ctx *c, int on)
{
    if (!c) return ARGUS_E_ARG;
    c->use_graph = on != 0;
    if (!on) drop_graph(c);
    return ARGUS_OK;
}

int argus_ba_project(argus_ba_ctx *c, double *hx)
{
    if (!c || !hx || c->m == 0) return ARGUS_E_ARG;
    ARGUS_CUDA_CHECK(cudaSetDevice(c->device));
    DevBuf<double> d;
    ARGUS_CUDA_CHECK(d.ensure(2 * c->nobs));
    k_lm_set_attempt<<<1, 32, 0, c->stream>>>(c->st.p, 0.0, c->cur_cam, c->cur_pts, c->mode, c->flags.p);
    int rc = enqueue_cost(c, 0, d.p);
    if (rc == ARGUS_OK) {
        cudaError_t e = cudaMemcpyAsync(hx, d.p, 2 * c->nobs * sizeof(double), cudaMemcpyDeviceToHost, c->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
        if (e != cudaSuccess) rc = ARGUS_E_CUDA;
    }
    d.release();
    return rc;
}

int argus_ba_normal_equations(argus_ba_ctx *c, double mu, double *U, double *ea, double *V, double *eb, double *S, double *E,
                              double *dp, double *scal)
{
    if (!c || c->m == 0 || c->mode != kModeMotStr) return ARGUS_E_ARG;
    ARGUS_CUDA_CHECK(cudaSetDevice(c->device));
    const int m = c->m, N = 16 * (m - c->mcon);
    cudaStream_t s = c->stream;
    k_lm_set_attempt<<<1, 32, 0, s>>>(c->st.p, mu, c->cur_cam, c->cur_pts, c->mode, c->flags.p);
    int rc = enqueue_lin0(c);
    if (rc) return rc;
    double sc[8] = {0};
    ARGUS_CUDA_CHECK(cudaMemcpyAsync(sc, c->scal.p, 4 * sizeof(double), cudaMemcpyDeviceToHost, s));
    rc = enqueue_attempt_motstr(c, false, c->dbuf.p);
    if (rc) return rc;
    double at[4], c2[2]; int fl[2];
    ARGUS_CUDA_CHECK(cudaMemcpyAsync(at, c->att3.p, sizeof(at), cudaMemcpyDeviceToHost, s));
    ARGUS_CUDA_CHECK(cudaMemcpyAsync(c2, c->cam2.p, sizeof(c2), cudaMemcpyDeviceToHost, s));
    ARGUS_CUDA_CHECK(cudaMemcpyAsync(fl, c->flags.p, sizeof(fl), cudaMemcpyDeviceToHost, s));
    ARGUS_CUDA_CHECK(cudaStreamSynchronize(s));
    sc[4] = c2[0] + at[1]; sc[5] = c2[1] + at[2]; sc[6] = at[0]; sc[7] = (double)fl[1];
    if (scal) memcpy(scal, sc, sizeof(sc));
    if (U || ea) {
        std::vector<double> h((size_t)m * kU);
        ARGUS_CUDA_CHECK(cudaMemcpyAsync(h.data(), c->red1, h.size() * sizeof(double), cudaMemcpyDeviceToHost, s));
        ARGUS_CUDA_CHECK(cudaStreamSynchronize(s));
        for (int j = 0; j < m; ++j) {
            if (U) for (int r = 0; r < 16; ++r) for (int q = 0; q < 16; ++q)
                U[(j * 16 + r) * 16 + q] = (r <= q) ? h[j * kU + tri16(r, q)] : h[j * kU + tri16(q, r)];
            if (ea) for (int r = 0; r < 16; ++r) ea[j * 16 + r] = h[j * kU + 136 + r];
        }
    }
    if (V) ARGUS_CUDA_CHECK(cudaMemcpyAsync(V, c->V.p, 6 * c->n * sizeof(double), cudaMemcpyDeviceToHost, s));
    if (eb) ARGUS_CUDA_CHECK(cudaMemcpyAsync(eb, c->eb.p, 3 * c->n * sizeof(double), cudaMemcpyDeviceToHost, s));
    if (dp) {
        ARGUS_CUDA_CHECK(cudaMemcpyAsync(dp, c->da_full.p, 16 * m * sizeof(double), cudaMemcpyDeviceToHost, s));
        ARGUS_CUDA_CHECK(cudaMemcpyAsync(dp + 16 * m, c->dbuf.p, 3 * c->n * sizeof(double), cudaMemcpyDeviceToHost, s));
    }
    if (S || E) {
        // re-assemble (the solve factored S in place)
        const int64_t slen = (int64_t)c->npairs * 256, tot = (int64_t)N * N;
        k_assemble_mma<<<(unsigned)((tot + 255) / 256), 256, 0, s>>>(m, c->mcon, c->red1, c->red2, c->red2 + slen, c->st.p, c->S.p, c->E.p);
        ARGUS_CUDA_CHECK(cudaGetLastError());
        if (S) ARGUS_CUDA_CHECK(cudaMemcpyAsync(S, c->S.p, tot * sizeof(double), cudaMemcpyDeviceToHost, s));
        if (E) ARGUS_CUDA_CHECK(cudaMemcpyAsync(E, c->E.p, N * sizeof(double), cudaMemcpyDeviceToHost, s));
    }
    ARGUS_CUDA_CHECK(cudaStreamSynchronize(s));
    return ARGUS_OK;
}

int argus_ba_solve(argus_ba_ctx *c, int itmax, int verbose, const double opts[5], double info[10], int flags)
{
    if (!c || !opts || c->m == 0) return ARGUS_E_ARG;
    ARGUS_CUDA_CHECK(cudaSetDevice(c->device));
    const bool compat = (flags & ARGUS_BA_COMPAT_SBA16) != 0;
    c->profiling = (flags & ARGUS_BA_PROFILE) != 0;
    memset(c->prof_ms, 0, sizeof(c->prof_ms)); memset(c->prof_n, 0, sizeof(c->prof_n));
    const int m = c->m, mode = c->mode;
    cudaStream_t s = c->stream;
    int rc;
    if (c->wand_n > 0 && mode != kModeMotStr) return ARGUS_E_ARG;      // the wand term couples points: motion + structure driver only
    // the reference refuses problems with fewer measurements than unknowns (sba_levmar.c:632-637); counts are global over the ranks
    double gc[3];
    if ((rc = global_counts(c, gc))) return rc;
    {
        const double nobs2 = 2.0 * gc[0];
        const double nvars = (mode == kModeMot) ? 16.0 * m : (mode == kModeStr) ? 3.0 * gc[1] : 16.0 * m + 3.0 * gc[1];
        if (nobs2 < nvars) {
            fprintf(stderr, "argus_b200: cannot solve a problem with fewer measurements [%.0f] than unknowns [%.0f]\n", nobs2, nvars);
            return ARGUS_E_SBA;
        }
    }
    if (itmax == 0) return 0;   // the reference runs its Jacobian checker here (sba_levmar.c:749-753); see tests instead
    if (verbose > 1 && !c->lmlog) ARGUS_CUDA_CHECK(cudaMalloc(&c->lmlog, (size_t)kLogCap * kLmLogDoubles * sizeof(double)));

    LmInit in;
    for (int q = 0; q < 5; ++q) in.opts[q] = opts[q];
    in.nblocks = (mode == kModeMot) ? (double)(m - c->mcon) : (mode == kModeStr) ? gc[2] : 1.0;
    in.itmax = itmax; in.compat = compat ? 1 : 0; in.mode = mode; in.cur_cam = c->cur_cam; in.cur_pts = c->cur_pts;
    in.logcap = (verbose > 1) ? kLogCap : 0; in.lmlog = (verbose > 1) ? c->lmlog : nullptr;
    k_lm_init<<<1, 32, 0, s>>>(c->st.p, in, c->flags.p);
    c->launches += 1;
    // initial cost, first linearisation, mu = tau * max diag
    if ((rc = enqueue_cost(c, 0, nullptr))) return rc;
    if ((rc = enqueue_lin0(c))) return rc;
    k_lm_begin<<<1, 32, 0, s>>>(c->st.p, c->att3.p, c->scal.p);
    c->launches += 1;
    ARGUS_CUDA_CHECK(cudaGetLastError());

    // damping attempts: queued kDepth ahead of the last one whose outcome has been seen.  Attempt q is queued only after
    // attempt q - kDepth - 1 is known not to have ended the loop - a rule every rank follows identically, so that all ranks
    // queue the same number of (collective-carrying) attempts; attempts queued behind the end of the loop return at once.
    // ARGUS_BA_PROFILE: every kProfEvery-th attempt is replayed from a second graph that has an event-record node on either side
    // of every phase.  Before such a replay the host waits for the outcome of everything queued so far, so that it is a real
    // attempt (not one queued behind the end of the loop) and its events - the ones that can be read after the solve - hold the
    // times of real work; the other attempts run as always.
    const bool graph = c->use_graph && (c->world == 1 || c->nccl != nullptr);
    if (graph && !c->graph_exec) capture_attempt_graph(c, false);
    const bool prof_graph = c->profiling && graph && c->graph_exec;
    if (prof_graph && !c->graph_exec_prof) capture_attempt_graph(c, true);
    int queued = 0, confirmed = 0, last_prof = -1;
    bool done = false;
    std::vector<size_t> ev_mark;          // profiling without graphs: first event of each attempt (attempts queued behind the end are dropped)
    while (!done) {
        const bool prof_this = prof_graph && c->graph_exec_prof && (queued % kProfEvery == kProfEvery - 1);
        const int depth = prof_this ? 0 : kDepth;
        while (queued - confirmed > depth) {
            ARGUS_CUDA_CHECK(cudaEventSynchronize(c->snap_ev[confirmed % kSnap]));
            if (c->h_snap[confirmed % kSnap].done) { done = true; break; }
            ++confirmed;
        }
        if (done) break;
        if (graph && c->graph_exec) {
            ARGUS_CUDA_CHECK(cudaGraphLaunch(prof_this ? c->graph_exec_prof : c->graph_exec, s));
            if (prof_this) last_prof = queued;
            c->launches += c->graph_launches; c->n_allreduce += c->graph_allreduces;
        } else {
            if (c->profiling) ev_mark.push_back(c->evs.size());
            if ((rc = enqueue_attempt(c))) return rc;
        }
        ARGUS_CUDA_CHECK(cudaMemcpyAsync(&c->h_snap[queued % kSnap], c->st.p, sizeof(LmState), cudaMemcpyDeviceToHost, s));
        ARGUS_CUDA_CHECK(cudaEventRecord(c->snap_ev[queued % kSnap], s));
        c->d2h_bytes += sizeof(LmState);
        ++queued;
    }
    ARGUS_CUDA_CHECK(cudaStreamSynchronize(s));
    const LmState fin = c->h_snap[(queued - 1) % kSnap];
    c->cur_cam = fin.cur_cam; c->cur_pts = fin.cur_pts;
    if (c->profiling) {
        // `confirmed` is the attempt that ended the loop; the ones queued behind it did nothing
        if ((size_t)confirmed + 1 < ev_mark.size()) {
            for (size_t q = ev_mark[confirmed + 1]; q < c->evs.size(); ++q) { c->ev_pool.push_back(c->evs[q].a); c->ev_pool.push_back(c->evs[q].b); }
            c->evs.resize(ev_mark[confirmed + 1]);
        }
        prof_collect(c);
        if (prof_graph && last_prof >= 0)                      // per-phase times of the last profiled attempt (one sample per phase and solve)
            for (auto &e : c->graph_evs) {
                float ms = 0.f;
                if (cudaEventElapsedTime(&ms, e.a, e.b) == cudaSuccess) { c->prof_ms[e.id] += ms; c->prof_n[e.id] += 1; } else cudaGetLastError();
            }
    }

    const double nvis = gc[0];
    const char *tag = (mode == kModeMot) ? "mot" : (mode == kModeStr) ? "str" : "motstr";
    if (verbose) printf("initial %s-SBA error %g [%g]\n", tag, fin.init_p_eL2, fin.init_p_eL2 / nvis);
    for (int q = 0; q < fin.nsingular; ++q) fprintf(stderr, "argus_b200: singular matrix V*_i, increasing damping\n");
    if (verbose > 1 && fin.nlog > 0) {
        std::vector<double> lg((size_t)fin.nlog * kLmLogDoubles);
        ARGUS_CUDA_CHECK(cudaMemcpy(lg.data(), c->lmlog, lg.size() * sizeof(double), cudaMemcpyDeviceToHost));
        for (int q = 0; q < fin.nlog; ++q) {
            const double *e = &lg[(size_t)q * kLmLogDoubles];
            if (e[5] != 0.0) { printf("\ndamping term %8g: %s\n", e[0], e[5] == 1.0 ? "singular V*_i" : "reduced system not positive definite"); continue; }
            printf("\ndamping term %8g, gain ratio %8g, errors %8g / %8g = %g\n", e[0], e[1] != 0.0 ? e[2] / e[1] : e[2] / DBL_EPSILON,
                   e[3] / nvis, e[4] / nvis, e[3] / e[4]);
        }
    }
    if (fin.err) {
        fprintf(stderr, "argus_b200: the matrix of the augmented normal equations is almost singular,\n"
                        "     minimization should be restarted from the current solution with an increased damping term\n");
        return ARGUS_E_SBA;
    }
    if (fin.stop == 6) fprintf(stderr, "argus_b200: too many failed attempts to increase the damping factor! Singular Hessian matrix?\n");
    if (info) {
        info[0] = fin.init_p_eL2; info[1] = fin.p_eL2; info[2] = fin.eab_inf; info[3] = fin.dp_L2; info[4] = fin.mu / fin.dmax;
        info[5] = fin.itno; info[6] = fin.stop; info[7] = fin.nfev; info[8] = fin.njev; info[9] = fin.nlss;
    }
    return (fin.stop != 7) ? fin.itno : ARGUS_E_SBA;
}

namespace {
int oneshot(int mode, int device, int64_t n, int64_t ncon, int m, int mcon, const char *vmask, double *p, const double *rot0, const double *x,
            int nccalib, int ncdist, int itmax, int verbose, const double opts[5], double info[10], int flags)
{
    argus_ba_ctx *c = nullptr;
    int rc = argus_ba_create(&c, device, nullptr);
    if (rc) return rc;
    c->use_graph = false;        // a graph pays off from the second solve of a resident problem, not for one call
    rc = argus_ba_set_mode(c, mode);
    if (rc == ARGUS_OK) rc = argus_ba_upload(c, n, ncon, m, mcon, vmask, p, rot0, x, nccalib, ncdist);
    int ret = rc;
    if (rc == ARGUS_OK) {
        ret = argus_ba_solve(c, itmax, verbose, opts, info, flags);
        if (ret >= 0 || ret == ARGUS_E_SBA) { rc = argus_ba_download(c, p); if (rc) ret = rc; }
    }
    argus_ba_destroy(c);
    return ret;
}

int blockdiag_oneshot(int mode, int64_t n, int64_t ncon, int m, int mcon, const char *vmask, double *moving, const double *frozen,
                      const double *rot0, const double *x, int nccalib, int ncdist, int itmax, int verbose, const double opts[5],
                      double info[10], int flags)
{
    if (!moving || !frozen || n < 0 || m < 1) return ARGUS_E_ARG;
    std::vector<double> p((size_t)16 * m + (size_t)3 * n);
    const bool mot = mode == kModeMot;
    memcpy(p.data(), mot ? moving : frozen, (size_t)16 * m * sizeof(double));
    memcpy(p.data() + 16 * m, mot ? frozen : moving, (size_t)3 * n * sizeof(double));
    const int ret = oneshot(mode, g_default_device, n, ncon, m, mcon, vmask, p.data(), rot0, x, nccalib, ncdist, itmax, verbose, opts, info, flags);
    if (ret >= 0 || ret == ARGUS_E_SBA)
        memcpy(moving, mot ? p.data() : p.data() + 16 * m, (mot ? (size_t)16 * m : (size_t)3 * n) * sizeof(double));
    return ret;
}
}  // namespace

int argus_ba_motstr_kdrts(int64_t n, int64_t ncon, int m, int mcon, const char *vmask, double *p, const double *rot0,
                          const double *x, int nccalib, int ncdist, int itmax, int verbose, const double opts[5], double info[10],
                          int flags)
{
    return oneshot(kModeMotStr, g_default_device, n, ncon, m, mcon, vmask, p, rot0, x, nccalib, ncdist, itmax, verbose, opts, info, flags);
}

int argus_ba_mot_kdrts(int64_t n, int m, int mcon, const char *vmask, double *cams, const double *pts, const double *rot0,
                       const double *x, int nccalib, int ncdist, int itmax, int verbose, const double opts[5], double info[10], int flags)
{
    return blockdiag_oneshot(kModeMot, n, 0, m, mcon, vmask, cams, pts, rot0, x, nccalib, ncdist, itmax, verbose, opts, info, flags);
}

int argus_ba_str_kdrts(int64_t n, int64_t ncon, int m, const char *vmask, double *pts, const double *cams, const double *rot0,
                       const double *x, int itmax, int verbose, const double opts[5], double info[10], int flags)
{
    return blockdiag_oneshot(kModeStr, n, ncon, m, 0, vmask, pts, cams, rot0, x, 0, 0, itmax, verbose, opts, info, flags);
}

}  // extern "C"
