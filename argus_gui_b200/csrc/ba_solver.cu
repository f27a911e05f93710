// Sparse Levenberg-Marquardt bundle adjustment on one B200 (one process per GPU; optional all-reduce hook).
// Host-side controller + C ABI.  Follows the control flow of sba_motstr_levmar_x
// (sba-1.6p/sba_levmar.c:449-1428) exactly - including, in ARGUS_BA_COMPAT_SBA16 mode, its behaviour after a
// rejected step - while every array stays resident in HBM and all arithmetic runs in the kernels of
// ba_kernels.cuh / ba_schur_mma.cuh.
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

using namespace argus;

namespace {

enum ProfId { P_RESIDUAL = 0, P_LINEARIZE, P_LIN_REDUCE, P_PREPARE, P_SCHUR, P_SCHUR_REDUCE, P_ASSEMBLE, P_CHOL, P_BACKSUB,
              P_ALLREDUCE, P_COUNT };
const char *kProfNames[P_COUNT] = {"residual", "linearize", "lin_reduce", "prepare", "schur", "schur_reduce", "assemble",
                                   "chol_solve", "backsub_eval", "allreduce"};

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

// host-side helper: split [0, n) over the host cores (the index build at upload is a pure host streaming job)
void parallel_for(int64_t n, const std::function<void(int64_t, int64_t)> &fn)
{
    unsigned hw = std::thread::hardware_concurrency();
    int T = (int)std::min<int64_t>(std::min<unsigned>(hw ? hw : 4, 16), (n + 65535) / 65536);
    if (T <= 1) { fn(0, n); return; }
    std::vector<std::thread> th;
    for (int q = 0; q < T; ++q) th.emplace_back([&, q] { fn(n * q / T, n * (q + 1) / T); });
    for (auto &x : th) x.join();
}

// process-wide bits that are expensive to set up per context (a one-shot call creates and destroys a context):
// kernel attributes already applied per device, and a free list of 128-byte pinned scratch blocks for the scalar read-backs
std::mutex g_once_mutex;
unsigned long long g_attr_done = 0;
std::vector<double *> g_pinned_free;
double *pinned_take()
{
    {
        std::lock_guard<std::mutex> lk(g_once_mutex);
        if (!g_pinned_free.empty()) { double *p = g_pinned_free.back(); g_pinned_free.pop_back(); return p; }
    }
    double *p = nullptr;
    return cudaMallocHost(&p, 16 * sizeof(double)) == cudaSuccess ? p : nullptr;
}
void pinned_give(double *p) { std::lock_guard<std::mutex> lk(g_once_mutex); g_pinned_free.push_back(p); }

}  // namespace

struct argus_ba_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    int sms = 148;
    // comm
    int rank = 0, world = 1;
    argus_allreduce_fn allreduce = nullptr;
    void *comm_user = nullptr;
    // problem
    int64_t n = 0, ncon = 0, nobs = 0;
    int m = 0, mcon = 0, nfix_k = 0, nfix_d = 0;
    DevBuf<int32_t> obs_ptr; DevBuf<uint8_t> obs_cam; DevBuf<double2> obs_uv; DevBuf<uint64_t> pt_mask;
    DevBuf<double> rot0, cam, pts, cam_new, pts_new, cam0, pts0;
    // work
    DevBuf<double> W, Y, V, eb, vstate, Vinv, cvec, dbuf;
    DevBuf<double> Upart, scpart, red1, max2, Epart, Spart, red2, S, E, da_free, da_full, part3, red3, cam2, scal, att, respart;
    DevBuf<int> flags;     // [0] singular V*, [1] chol status
    // generation-2 (tile / MMA) plan
    int gen = 3;
    int mode = kModeMotStr;   // which half of the parameters moves (argus_ba_set_mode)
    int schur_nw_pref = 12;
    bool chol_coop = true;     // multi-SM Cholesky (ARGUS_CHOL_COOP=0 falls back to the single-CTA kernels)
    int chol_coop_min = 192;   // smallest reduced system that goes to it (ARGUS_CHOL_COOP_MIN)
    DevBuf<int32_t> tile_pt, tile_o; DevBuf<uint16_t> blk_jk; DevBuf<int32_t> blk_pair; DevBuf<uint8_t> zsign; DevBuf<double> Z;
    DevBuf<uint8_t> obs_lpt, corder, pt_rank, vmask_dev; DevBuf<int32_t> idx_btot; int rank_mp = 16;
    DevBuf<uint8_t> cam_fix; bool cam_fix_on = false; DevBuf<uint16_t> cstart; DevBuf<double> Upart2, Edirect, cholwork;
    int grid_lin = 1; size_t lin_smem = 0;
    unsigned long long *lin_dbg = nullptr;
    int ntiles = 0, tile_obs = 256, schur_G = 1, schur_NB = 1, schur_NW = 8, schur_nchunk = 1;
    double *h_pinned = nullptr;   // 16 doubles
    Arena arena;
    int grid_pts128 = 1, grid_pts256 = 1, nchunks = 1, npairs = 1;
    int64_t chunk_pts = 1;
    // profile
    bool profiling = false;
    struct Ev { int id; cudaEvent_t a, b; };
    std::vector<Ev> evs;
    std::vector<cudaEvent_t> ev_pool;
    double prof_ms[P_COUNT]; int64_t prof_n[P_COUNT];
    int64_t launches = 0, h2d_bytes = 0, d2h_bytes = 0, n_allreduce = 0;

    TilePlan tile_plan() const { TilePlan tp; tp.ntiles = ntiles; tp.tile_obs = tile_obs; tp.tile_pt = tile_pt.p; tp.tile_o = tile_o.p; return tp; }
    SchurPlan schur_plan() const {
        SchurPlan sp; sp.G = schur_G; sp.nchunk = schur_nchunk; sp.npairs = npairs; sp.blk_jk = blk_jk.p; sp.blk_pair = blk_pair.p;
        return sp;
    }
    LinTilePlan lin_plan() const { LinTilePlan lp; lp.obs_lpt = obs_lpt.p; lp.corder = corder.p; lp.cstart = cstart.p; return lp; }
    LinTileOut lin_out() {
        LinTileOut o; o.Z = Z.p; o.zsign = zsign.p; o.Vinv = Vinv.p; o.V = V.p; o.eb = eb.p; o.vstate = vstate.p; o.Upart = Upart2.p;
        o.scpart = scpart.p; o.flag = flags.p; o.dbg = lin_dbg; return o;
    }
    BaProblem problem() const {
        BaProblem pb;
        pb.n = n; pb.ncon = (mode == kModeMot) ? n : ncon; pb.m = m; pb.mcon = (mode == kModeStr) ? m : mcon; pb.nobs = nobs; pb.nfix_k = nfix_k; pb.nfix_d = nfix_d; pb.cam_fix = cam_fix_on ? cam_fix.p : nullptr;
        pb.obs_ptr = obs_ptr.p; pb.obs_cam = obs_cam.p; pb.obs_uv = obs_uv.p; pb.pt_mask = pt_mask.p; pb.rot0 = rot0.p;
        return pb;
    }
};

namespace {

cudaEvent_t get_event(argus_ba_ctx *c)
{
    if (!c->ev_pool.empty()) { cudaEvent_t e = c->ev_pool.back(); c->ev_pool.pop_back(); return e; }
    cudaEvent_t e; cudaEventCreate(&e); return e;
}
struct ProfScope {
    argus_ba_ctx *c; int id; cudaEvent_t a = nullptr, b = nullptr;
    ProfScope(argus_ba_ctx *c_, int id_) : c(c_), id(id_) {
        if (c->profiling) { a = get_event(c); b = get_event(c); cudaEventRecord(a, c->stream); }
    }
    ~ProfScope() { if (c->profiling) { cudaEventRecord(b, c->stream); c->evs.push_back({id, a, b}); } }
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

int do_allreduce(argus_ba_ctx *c, double *buf, int64_t count, int op)
{
    if (c->world <= 1) return ARGUS_OK;
    if (!c->allreduce) return ARGUS_E_COMM;
    ProfScope ps(c, P_ALLREDUCE);
    c->n_allreduce += 1;
    return c->allreduce(c->comm_user, buf, count, op, (void *)c->stream) == 0 ? ARGUS_OK : ARGUS_E_COMM;
}

size_t cam_smem(int m) { return (size_t)m * sizeof(CamConst); }

// ---- one evaluation of ||x - hx(p)||^2 on (cam, pts); result left in red3[0] (after optional all-reduce) ----
int run_residual(argus_ba_ctx *c, const double *cam, const double *pts, double *hx)
{
    {
        ProfScope ps(c, P_RESIDUAL);
        k_residual<<<c->grid_pts256, 256, cam_smem(c->m), c->stream>>>(c->problem(), cam, pts, hx, c->respart.p);
        k_reduce_sum<<<1, 32, 0, c->stream>>>(c->respart.p, c->grid_pts256, 1, 1, c->red3.p);
        c->launches += 2;
    }
    ARGUS_CUDA_CHECK(cudaGetLastError());
    return do_allreduce(c, c->red3.p, 1, 0);
}

int run_lin_tile(argus_ba_ctx *c, int mode, double mu, int use_state, int write_state);

// ---- linearise at (cam, pts): W, V, eb, reduced U/ea in red1, scalars in scal[0..3] ----
int run_linearize(argus_ba_ctx *c)
{
    const int m = c->m;
    if (c->gen >= 3) return run_lin_tile(c, 0, 0.0, 0, 0);
    {
        ProfScope ps(c, P_LINEARIZE);
        const size_t sm = cam_smem(m) + (size_t)m * kU * sizeof(double);
        k_linearize<<<c->grid_pts128, 128, sm, c->stream>>>(c->problem(), c->cam.p, c->pts.p, c->W.p, c->V.p, c->eb.p,
                                                           c->Upart.p, c->scpart.p);
    }
    {
        ProfScope ps(c, P_LIN_REDUCE);
        const int64_t len = (int64_t)m * kU;
        k_reduce_sum<<<(unsigned)((len + 127) / 128), 128, 0, c->stream>>>(c->Upart.p, c->grid_pts128, len, len, c->red1.p);
        k_reduce_lin_scalars<<<1, 32, 0, c->stream>>>(c->scpart.p, c->grid_pts128, c->red1.p + len, c->max2.p);
    }
    ARGUS_CUDA_CHECK(cudaGetLastError());
    int rc = do_allreduce(c, c->red1.p, (int64_t)m * kU + 2, 0);
    if (rc) return rc;
    rc = do_allreduce(c, c->max2.p, 2, 1);
    if (rc) return rc;
    k_lin_finalize<<<1, 32, 0, c->stream>>>(c->red1.p, c->cam.p, m, c->mcon, c->red1.p + (int64_t)m * kU, c->max2.p, c->scal.p);
    c->launches += 4;
    ARGUS_CUDA_CHECK(cudaGetLastError());
    return ARGUS_OK;
}

// ---- generation 2: tile linearisation (mode 0: statistics only; mode 1: + damping, Z, E) ----
int run_lin_tile(argus_ba_ctx *c, int mode, double mu, int use_state, int write_state)
{
    const int m = c->m;
    if (mode != 0) ARGUS_CUDA_CHECK(cudaMemsetAsync(c->flags.p, 0, sizeof(int), c->stream));
    {
        ProfScope ps(c, P_LINEARIZE);
#define LIN_LAUNCH(MC_) k_lin_tile<MC_><<<c->grid_lin, 256, c->lin_smem, c->stream>>>(c->problem(), c->tile_plan(), c->lin_plan(), c->cam.p, \
                                                                                    c->pts.p, mode, mu, use_state, write_state, c->lin_out())
        const int mc = (m - c->mcon + 7) / 8;
        if (mc <= 2) LIN_LAUNCH(2); else if (mc <= 4) LIN_LAUNCH(4); else LIN_LAUNCH(8);
#undef LIN_LAUNCH
    }
    const int64_t slen = (int64_t)c->npairs * 256;
    {
        ProfScope ps(c, P_LIN_REDUCE);
        k_lin_tile_reduce<<<m, 192, 0, c->stream>>>(m, c->grid_lin, c->Upart2.p, c->scpart.p, c->red1.p, c->red2.p + slen,
                                                   c->red1.p + (int64_t)m * kU, c->max2.p);
    }
    c->launches += 2;
    ARGUS_CUDA_CHECK(cudaGetLastError());
    int rc = do_allreduce(c, c->red1.p, (int64_t)m * kU + 2, 0);
    if (rc) return rc;
    rc = do_allreduce(c, c->max2.p, 2, 1);
    if (rc) return rc;
    k_lin_finalize<<<1, 32, 0, c->stream>>>(c->red1.p, c->cam.p, m, c->mcon, c->red1.p + (int64_t)m * kU, c->max2.p, c->scal.p);
    c->launches += 1;
    ARGUS_CUDA_CHECK(cudaGetLastError());
    return ARGUS_OK;
}

// red3 = {cost, ||db||^2, dL} summed over ranks; the "a block was singular" flag travels in red3[3] so that every rank takes the
// same accept / reject decision
__global__ void k_flag_to_double(const int *__restrict__ flags, double *__restrict__ dst) { if (threadIdx.x == 0) *dst = (double)flags[0]; }
__global__ void k_double_to_flag(const double *__restrict__ src, int *__restrict__ flags) { if (threadIdx.x == 0) flags[0] = (*src != 0.0) ? 1 : 0; }
int reduce_attempt_scalars(argus_ba_ctx *c)
{
    if (c->world <= 1) return ARGUS_OK;
    k_flag_to_double<<<1, 32, 0, c->stream>>>(c->flags.p, c->red3.p + 3);
    int rc = do_allreduce(c, c->red3.p, 4, 0);
    if (rc) return rc;
    k_double_to_flag<<<1, 32, 0, c->stream>>>(c->red3.p + 3, c->flags.p);
    c->launches += 2;
    return ARGUS_OK;
}

// ---- one damping attempt: prepare, Schur complement, solve, back-substitute, evaluate; scalars in att[0..3] ----
int run_attempt(argus_ba_ctx *c, double mu, double mu_u, int use_state, int write_state, double *dbout, bool need_lin = true)
{
    const int m = c->m, mcon = c->mcon, mf = m - mcon, N = 16 * mf;
    if (c->gen < 3) ARGUS_CUDA_CHECK(cudaMemsetAsync(c->flags.p, 0, 2 * sizeof(int), c->stream));
    else ARGUS_CUDA_CHECK(cudaMemsetAsync(c->flags.p + 1, 0, sizeof(int), c->stream));
    if (c->gen >= 2) {
        if (c->gen >= 3) {
            if (need_lin) { int rc0 = run_lin_tile(c, 1, mu, use_state, write_state); if (rc0) return rc0; }
        } else {
            ProfScope ps(c, P_PREPARE);
            k_make_z<<<c->grid_pts128, 128, (size_t)m * 16 * sizeof(double), c->stream>>>(
                c->problem(), c->tile_plan(), mu, use_state, write_state, c->V.p, c->eb.p, c->vstate.p, c->W.p, c->Z.p, c->zsign.p, c->Vinv.p,
                c->Epart.p, c->flags.p);
        }
        {
            ProfScope ps(c, P_SCHUR);
            dim3 grid(c->schur_nchunk, c->schur_G);
            const size_t sm = (size_t)2 * ((size_t)c->tile_obs * 384 + 384 + (size_t)kTilePts * c->rank_mp) + 384;
#define SCHUR_LAUNCH(NB_, NW_, PR_) k_schur_mma<NB_, NW_, PR_><<<grid, (NW_ + PR_) * 32, sm, c->stream>>>(c->problem(), c->tile_plan(), c->schur_plan(), c->Z.p, \
                                                                                       c->zsign.p, c->pt_rank.p, c->rank_mp, c->Spart.p)
            switch (c->schur_NB * 100 + c->schur_NW) {
                case 108: SCHUR_LAUNCH(1, 8, 0); break;
                case 308: SCHUR_LAUNCH(3, 8, 0); break;
                case 508: SCHUR_LAUNCH(5, 8, 0); break;
                case 612: SCHUR_LAUNCH(6, 12, 0); break;
                case 711: SCHUR_LAUNCH(7, 11, 1); break;
                default: SCHUR_LAUNCH(9, 8, 0); break;
            }
#undef SCHUR_LAUNCH
        }
    } else {
        {
            ProfScope ps(c, P_PREPARE);
            k_prepare<<<c->grid_pts128, 128, (size_t)m * 16 * sizeof(double), c->stream>>>(
                c->problem(), mu, use_state, write_state, c->V.p, c->eb.p, c->vstate.p, c->W.p, c->Y.p, c->Vinv.p, c->cvec.p,
                c->Epart.p, c->flags.p);
        }
        {
            ProfScope ps(c, P_SCHUR);
            dim3 grid(c->npairs, c->nchunks);
            k_schur_pairs<<<grid, 256, 0, c->stream>>>(c->problem(), c->Y.p, c->W.p, c->chunk_pts, c->Spart.p);
        }
    }
    const int64_t slen = (int64_t)c->npairs * 256;
    {
        ProfScope ps(c, P_SCHUR_REDUCE);
        k_reduce_sum<<<(unsigned)((slen + 127) / 128), 128, 0, c->stream>>>(c->Spart.p, c->gen >= 2 ? c->schur_nchunk : c->nchunks, slen, slen, c->red2.p);
        if (c->gen < 3)
            k_reduce_sum<<<(unsigned)((m * 16 + 127) / 128), 128, 0, c->stream>>>(c->Epart.p, c->grid_pts128, (int64_t)m * 16, (int64_t)m * 16,
                                                                                 c->red2.p + slen);
    }
    ARGUS_CUDA_CHECK(cudaGetLastError());
    int rc = do_allreduce(c, c->red2.p, slen + (int64_t)m * 16, 0);
    if (rc) return rc;
    {
        ProfScope ps(c, P_ASSEMBLE);
        const int64_t tot = (int64_t)N * N;
        if (c->gen >= 2) k_assemble_mma<<<(unsigned)((tot + 255) / 256), 256, 0, c->stream>>>(m, mcon, c->red1.p, c->red2.p, c->red2.p + slen, mu_u, c->S.p, c->E.p, c->gen >= 3);
        else k_assemble<<<(unsigned)((tot + 255) / 256), 256, 0, c->stream>>>(m, mcon, c->red1.p, c->red2.p, c->red2.p + slen, mu_u, c->S.p, c->E.p);
    }
    {
        ProfScope ps(c, P_CHOL);
        const int nbw = (N > 864) ? 16 : 32;                   // panel width: the 32-wide panel of N > 864 exceeds the shared memory of one SM
        size_t rows = (size_t)(N > nbw ? N - nbw : 1);
        size_t sm = rows * (nbw + 1) * sizeof(double);
        if (sm < (size_t)N * sizeof(double)) sm = (size_t)N * sizeof(double);
        if (c->gen >= 3 && c->chol_coop && N >= c->chol_coop_min) {
            // multi-SM factorisation (cooperative grid): the single-CTA kernels below are latency bound
            ARGUS_CUDA_CHECK(cudaMemsetAsync(c->flags.p + 2, 0, sizeof(int), c->stream));
            double *Sp = c->S.p, *xp = c->da_free.p, *wp = c->cholwork.p; const double *Ep = c->E.p; int Nn = N;
            int *st = c->flags.p + 1, *fl = c->flags.p + 2;
            void *args[] = {&Sp, &Ep, &xp, &Nn, &wp, &st, &fl};
            const int nblk = (N + 31) / 32;
            int grid = std::min(c->sms, std::max(2, (nblk - 1) * nblk / 2));
            ARGUS_CUDA_CHECK(cudaLaunchCooperativeKernel((void *)k_chol_coop, dim3(grid), dim3(kCoopThreads), args, 0, c->stream));
        } else if (c->gen >= 3 && N <= 832) k_chol_solve2<<<1, kCholThreads, sm, c->stream>>>(c->S.p, c->E.p, c->da_free.p, N, c->cholwork.p, c->flags.p + 1);
        else if (nbw == 32) k_chol_solve_t<32><<<1, kCholThreads, sm, c->stream>>>(c->S.p, c->E.p, c->da_free.p, N, c->flags.p + 1);
        else k_chol_solve_t<16><<<1, kCholThreads, sm, c->stream>>>(c->S.p, c->E.p, c->da_free.p, N, c->flags.p + 1);
    }
    {
        ProfScope ps(c, P_BACKSUB);
        k_cam_update<<<1, 256, 0, c->stream>>>(m, mcon, c->cam.p, c->da_free.p, c->red1.p, mu, c->cam_new.p, c->da_full.p, c->cam2.p);
        if (c->gen >= 3) {
            const size_t sm = 2 * cam_smem(m) + (size_t)m * 17 * sizeof(double);
            k_backsub_recompute<<<c->grid_pts128, 128, sm, c->stream>>>(c->problem(), mu, c->cam.p, c->cam_new.p, c->da_full.p, c->pts.p,
                                                                       c->Vinv.p, c->eb.p, c->pts_new.p, dbout, c->part3.p);
        } else {
            const size_t sm = cam_smem(m) + (size_t)m * 16 * sizeof(double);
            k_backsub_eval<<<c->grid_pts128, 128, sm, c->stream>>>(c->problem(), mu, c->cam_new.p, c->da_full.p, c->pts.p, c->W.p, c->Vinv.p,
                                                                  c->eb.p, c->pts_new.p, dbout, c->part3.p);
        }
        k_reduce_sum<<<1, 32, 0, c->stream>>>(c->part3.p, c->grid_pts128, 3, 3, c->red3.p);
    }
    ARGUS_CUDA_CHECK(cudaGetLastError());
    rc = reduce_attempt_scalars(c);
    if (rc) return rc;
    k_attempt_finalize<<<1, 32, 0, c->stream>>>(c->flags.p, c->cam2.p, c->red3.p, c->att.p);
    c->launches += 10;
    ARGUS_CUDA_CHECK(cudaGetLastError());
    return ARGUS_OK;
}

int read_back(argus_ba_ctx *c, const double *dsrc, int count, double *hdst)
{
    ARGUS_CUDA_CHECK(cudaMemcpyAsync(c->h_pinned, dsrc, count * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    ARGUS_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    memcpy(hdst, c->h_pinned, count * sizeof(double));
    c->d2h_bytes += count * sizeof(double);
    return ARGUS_OK;
}


// ---- motion-only / structure-only LM (sba_mot_levmar_x sba_levmar.c:1437-1980, sba_str_levmar_x :1987-2540) ----
int run_bd_linearize(argus_ba_ctx *c)
{
    const int m = c->m;
    {
        ProfScope ps(c, P_LINEARIZE);
#define LIN_LAUNCH(MC_) k_lin_tile<MC_><<<c->grid_lin, 256, c->lin_smem, c->stream>>>(c->problem(), c->tile_plan(), c->lin_plan(), c->cam.p, \
                                                                                    c->pts.p, 0, 0.0, 0, 0, c->lin_out())
        const int mc = (m - c->mcon + 7) / 8;
        if (mc <= 2) LIN_LAUNCH(2); else if (mc <= 4) LIN_LAUNCH(4); else LIN_LAUNCH(8);
#undef LIN_LAUNCH
    }
    {
        ProfScope ps(c, P_LIN_REDUCE);
        k_lin_tile_reduce<<<m, 192, 0, c->stream>>>(m, c->grid_lin, c->Upart2.p, c->scpart.p, c->red1.p, c->red2.p + (int64_t)c->npairs * 256,
                                                   c->red1.p + (int64_t)m * kU, c->max2.p);
    }
    ARGUS_CUDA_CHECK(cudaGetLastError());
    int rc = do_allreduce(c, c->red1.p, (int64_t)m * kU + 2, 0);
    if (rc) return rc;
    rc = do_allreduce(c, c->max2.p, 2, 1);
    if (rc) return rc;
    k_bd_finalize<<<1, 32, 0, c->stream>>>(c->mode, c->red1.p, c->cam.p, m, c->mcon, c->red1.p + (int64_t)m * kU, c->max2.p, c->scal.p);
    c->launches += 3;
    ARGUS_CUDA_CHECK(cudaGetLastError());
    return ARGUS_OK;
}

// one attempt: att = {all blocks solved, ||dp||^2, dL, cost at the trial point, 0}
int run_bd_attempt(argus_ba_ctx *c, double mu, double mu_diag)
{
    const int m = c->m;
    ARGUS_CUDA_CHECK(cudaMemsetAsync(c->flags.p, 0, 2 * sizeof(int), c->stream));
    if (c->mode == kModeMot) {
        {
            ProfScope ps(c, P_CHOL);
            k_mot_solve<<<1, 64, 0, c->stream>>>(m, c->mcon, c->red1.p, mu_diag, mu, c->cam.p, c->cam_new.p, c->cam2.p, c->flags.p);
        }
        ARGUS_CUDA_CHECK(cudaMemsetAsync(c->red3.p, 0, 4 * sizeof(double), c->stream));
        int rc = run_residual(c, c->cam_new.p, c->pts.p, nullptr);           // cost -> red3[0] (all-reduced)
        if (rc) return rc;
        c->launches += 1;
    } else {
        ProfScope ps(c, P_BACKSUB);
        ARGUS_CUDA_CHECK(cudaMemsetAsync(c->da_full.p, 0, 16 * m * sizeof(double), c->stream));
        ARGUS_CUDA_CHECK(cudaMemsetAsync(c->cam2.p, 0, 2 * sizeof(double), c->stream));
        k_str_invert<<<(unsigned)((c->n + 255) / 256), 256, 0, c->stream>>>(c->n, c->ncon, mu_diag, c->V.p, c->Vinv.p, c->flags.p);
        const size_t sm = 2 * cam_smem(m) + (size_t)m * 17 * sizeof(double);
        k_backsub_recompute<<<c->grid_pts128, 128, sm, c->stream>>>(c->problem(), mu, c->cam.p, c->cam.p, c->da_full.p, c->pts.p, c->Vinv.p,
                                                                   c->eb.p, c->pts_new.p, nullptr, c->part3.p);
        k_reduce_sum<<<1, 32, 0, c->stream>>>(c->part3.p, c->grid_pts128, 3, 3, c->red3.p);
        ARGUS_CUDA_CHECK(cudaGetLastError());
        int rc = reduce_attempt_scalars(c);
        if (rc) return rc;
        c->launches += 3;
    }
    k_attempt_finalize<<<1, 32, 0, c->stream>>>(c->flags.p, c->cam2.p, c->red3.p, c->att.p);
    c->launches += 1;
    ARGUS_CUDA_CHECK(cudaGetLastError());
    return ARGUS_OK;
}

int solve_blockdiag(argus_ba_ctx *c, int itmax, int verbose, const double opts[5], double info[10], int flags)
{
    const bool compat = (flags & ARGUS_BA_COMPAT_SBA16) != 0;
    const bool mot = c->mode == kModeMot;
    const int m = c->m;
    const int64_t nvis = c->nobs;
    double h[8];
    int rc;
    // number of block systems per attempt (info[9] counts every one of them, :1851 / :2396), over all ranks
    double nblocks = mot ? (double)(m - c->mcon) : (double)(c->n - c->ncon);
    if (c->world == 1) {
        const int64_t nobs2 = nvis * 2, nvars = mot ? (int64_t)m * 16 : c->n * 3;
        if (nobs2 < nvars) {
            fprintf(stderr, "argus_b200: cannot solve a problem with fewer measurements [%lld] than unknowns [%lld]\n",
                    (long long)nobs2, (long long)nvars);
            return ARGUS_E_SBA;
        }
    } else if (!mot) {
        ARGUS_CUDA_CHECK(cudaMemcpyAsync(c->red3.p, &nblocks, sizeof(double), cudaMemcpyHostToDevice, c->stream));
        ARGUS_CUDA_CHECK(cudaStreamSynchronize(c->stream));
        if ((rc = do_allreduce(c, c->red3.p, 1, 0))) return rc;
        if ((rc = read_back(c, c->red3.p, 1, h))) return rc;
        nblocks = h[0];
    }
    if (itmax == 0) return 0;
    const double tau = fabs(opts[0]), eps1 = fabs(opts[1]), eps2 = fabs(opts[2]), eps2_sq = opts[2] * opts[2],
                 eps3_sq = opts[3] * opts[3], eps4_sq = opts[4] * opts[4];
    double mu = 0.0, g_inf = 0.0, p_eL2, pdp_eL2, p_L2 = 0.0, dp_L2 = DBL_MAX, dF, dL, dmax = DBL_MIN, init_p_eL2, nlss = 0.0;
    int nu = 2, nu2, stop = 0, nfev = 0, njev = 0, itno;

    if ((rc = run_residual(c, c->cam.p, c->pts.p, nullptr))) return rc;
    if ((rc = read_back(c, c->red3.p, 1, h))) return rc;
    p_eL2 = h[0]; nfev = 1;
    if (verbose) printf("initial %s-SBA error %g [%g]\n", mot ? "mot" : "str", p_eL2, p_eL2 / (double)nvis);
    init_p_eL2 = p_eL2;
    if (!isfinite(p_eL2)) stop = 7;

    for (itno = 0; itno < itmax && !stop; ++itno) {
        if ((rc = run_bd_linearize(c))) return rc;
        ++njev;
        if ((rc = read_back(c, c->scal.p, 4, h))) return rc;
        g_inf = h[1]; p_L2 = h[2]; dmax = h[3];
        if (g_inf <= eps1) { dp_L2 = 0.0; stop = 1; break; }
        if (itno == 0) mu = tau * dmax;
        double mu_diag = 0.0;       // what sits on the diagonals: never taken off after a rejected step in sba 1.6 (:1932, :2477 are #if 0)
        while (1) {
            mu_diag = compat ? mu_diag + mu : mu;
            if ((rc = run_bd_attempt(c, mu, mu_diag))) return rc;
            if ((rc = read_back(c, c->att.p, 5, h))) return rc;
            nlss += nblocks;
            const bool solved = mot ? (h[0] != 0.0) : (h[4] == 0.0);
            if (solved) {
                dp_L2 = h[1];
                if (dp_L2 <= eps2_sq * p_L2) { stop = 2; break; }
                if (dp_L2 >= (p_L2 + eps2) / 1e-24) {
                    fprintf(stderr, "argus_b200: the matrix of the augmented normal equations is almost singular,\n"
                                    "     minimization should be restarted from the current solution with an increased damping term\n");
                    if (c->profiling) prof_collect(c);
                    return ARGUS_E_SBA;
                }
                ++nfev;
                pdp_eL2 = h[3];
                if (!isfinite(pdp_eL2)) { stop = 7; break; }
                dL = h[2];
                dF = p_eL2 - pdp_eL2;
                if (verbose > 1)
                    printf("\ndamping term %8g, gain ratio %8g, errors %8g / %8g = %g\n", mu, dL != 0.0 ? dF / dL : dF / DBL_EPSILON,
                           p_eL2 / nvis, pdp_eL2 / nvis, p_eL2 / pdp_eL2);
                if (dL > 0.0 && dF > 0.0) {
                    double tmp = (2.0 * dF / dL - 1.0);
                    tmp = 1.0 - tmp * tmp * tmp;
                    mu = mu * ((tmp >= 0.3333333334) ? tmp : 0.3333333334);
                    nu = 2;
                    if (pdp_eL2 - 2.0 * sqrt(p_eL2 * pdp_eL2) < (eps4_sq - 1.0) * p_eL2) stop = 4;
                    if (mot) { std::swap(c->cam.p, c->cam_new.p); std::swap(c->cam.cap, c->cam_new.cap); }
                    else { std::swap(c->pts.p, c->pts_new.p); std::swap(c->pts.cap, c->pts_new.cap); }
                    p_eL2 = pdp_eL2;
                    break;
                }
            }
            mu *= nu;
            nu2 = (int)((unsigned)nu << 1);
            if (nu2 <= nu) {
                fprintf(stderr, "argus_b200: too many failed attempts to increase the damping factor! Singular Hessian matrix?\n");
                stop = 6;
                break;
            }
            nu = nu2;
        }
        if (p_eL2 <= eps3_sq) stop = 5;
    }
    if (itno >= itmax) stop = 3;
    if (c->profiling) prof_collect(c);
    if (info) {
        info[0] = init_p_eL2; info[1] = p_eL2; info[2] = g_inf; info[3] = dp_L2; info[4] = mu / dmax;
        info[5] = itno; info[6] = stop; info[7] = nfev; info[8] = njev; info[9] = nlss;
    }
    return (stop != 7) ? itno : ARGUS_E_SBA;
}

}  // namespace

extern "C" {

const char *argus_b200_version(void) { return "argus_b200 0.1 (sm_100a, fp64)"; }

int argus_ba_create(argus_ba_ctx **out, int device, void *stream)
{
    if (!out) return ARGUS_E_ARG;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || device >= ndev) {
        fprintf(stderr, "argus_b200: no CUDA device available (there is no CPU fallback)\n");
        return ARGUS_E_CUDA;
    }
    ARGUS_CUDA_CHECK(cudaSetDevice(device));
    argus_ba_ctx *c = new argus_ba_ctx();
    c->device = device;
    if (stream) { c->stream = (cudaStream_t)stream; c->own_stream = false; }
    else { ARGUS_CUDA_CHECK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)); c->own_stream = true; }
    ARGUS_CUDA_CHECK(cudaDeviceGetAttribute(&c->sms, cudaDevAttrMultiProcessorCount, device));
    c->h_pinned = pinned_take();
    if (!c->h_pinned) { delete c; return ARGUS_E_CUDA; }
    memset(c->prof_ms, 0, sizeof(c->prof_ms)); memset(c->prof_n, 0, sizeof(c->prof_n));
    if (const char *e = getenv("ARGUS_BA_GEN")) c->gen = atoi(e);
    {   // opt in to large dynamic shared memory where needed: once per device and process
        std::lock_guard<std::mutex> lk(g_once_mutex);
        if (device < 64 && !(g_attr_done & (1ull << device))) {
            const int smx = 2 * (256 * 384 + 384 + 32 * 64) + 384;
            ARGUS_CUDA_CHECK(cudaFuncSetAttribute(k_linearize, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
            ARGUS_CUDA_CHECK(cudaFuncSetAttribute(k_chol_solve_t<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 223000));
            ARGUS_CUDA_CHECK(cudaFuncSetAttribute(k_chol_solve_t<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 223000));
            ARGUS_CUDA_CHECK(cudaFuncSetAttribute(k_chol_solve2, cudaFuncAttributeMaxDynamicSharedMemorySize, 214000));
            ARGUS_CUDA_CHECK(cudaFuncSetAttribute(k_schur_mma<1, 8, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smx));
            ARGUS_CUDA_CHECK(cudaFuncSetAttribute(k_schur_mma<3, 8, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smx));
            ARGUS_CUDA_CHECK(cudaFuncSetAttribute(k_schur_mma<5, 8, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smx));
            ARGUS_CUDA_CHECK(cudaFuncSetAttribute(k_schur_mma<9, 8, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smx));
            ARGUS_CUDA_CHECK(cudaFuncSetAttribute(k_schur_mma<6, 12, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smx));
            ARGUS_CUDA_CHECK(cudaFuncSetAttribute(k_schur_mma<7, 11, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smx));
            g_attr_done |= 1ull << device;
        }
    }
    if (const char *e = getenv("ARGUS_SCHUR_NW")) c->schur_nw_pref = atoi(e);
    if (const char *e = getenv("ARGUS_CHOL_COOP")) c->chol_coop = atoi(e) != 0;
    if (const char *e = getenv("ARGUS_CHOL_COOP_MIN")) c->chol_coop_min = atoi(e);
    { int coop = 0; cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, device); if (!coop) c->chol_coop = false; }
    if (getenv("ARGUS_LIN_DEBUG")) { ARGUS_CUDA_CHECK(cudaMalloc(&c->lin_dbg, 8 * sizeof(unsigned long long))); ARGUS_CUDA_CHECK(cudaMemset(c->lin_dbg, 0, 8 * sizeof(unsigned long long))); }
    *out = c;
    return ARGUS_OK;
}

void argus_ba_destroy(argus_ba_ctx *c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    if (c->lin_dbg) {
        unsigned long long h[8]; cudaMemcpy(h, c->lin_dbg, sizeof(h), cudaMemcpyDeviceToHost);
        fprintf(stderr, "argus_b200 lin-tile phase clocks (sum over CTAs): top/fill %llu  P1 %llu  P2 %llu  P3 %llu  P4 %llu\n", h[0], h[1], h[2], h[3], h[4]);
        cudaFree(c->lin_dbg);
    }
    c->obs_ptr.release(); c->obs_cam.release(); c->obs_uv.release(); c->pt_mask.release();
    DevBuf<double> *bufs[] = {&c->rot0, &c->cam, &c->pts, &c->cam_new, &c->pts_new, &c->cam0, &c->pts0, &c->W, &c->Y, &c->V, &c->eb,
                              &c->vstate, &c->Vinv, &c->cvec, &c->dbuf, &c->Upart, &c->scpart, &c->red1, &c->max2, &c->Epart, &c->Spart,
                              &c->red2, &c->S, &c->E, &c->da_free, &c->da_full, &c->part3, &c->red3, &c->cam2, &c->scal, &c->att, &c->respart};
    for (auto *b : bufs) b->release();
    c->obs_lpt.release(); c->corder.release(); c->pt_rank.release(); c->cam_fix.release(); c->cstart.release(); c->Upart2.release(); c->Edirect.release(); c->cholwork.release();
    c->flags.release(); c->tile_pt.release(); c->tile_o.release(); c->blk_jk.release(); c->blk_pair.release(); c->zsign.release(); c->Z.release();
    for (auto e : c->ev_pool) cudaEventDestroy(e);
    if (c->arena.base) cudaFree(c->arena.base);
    if (c->h_pinned) pinned_give(c->h_pinned);
    if (c->own_stream) cudaStreamDestroy(c->stream);
    delete c;
}

int argus_ba_set_comm(argus_ba_ctx *c, int rank, int world, argus_allreduce_fn fn, void *user)
{
    if (!c || world < 1 || rank < 0 || rank >= world || (world > 1 && !fn)) return ARGUS_E_ARG;
    c->rank = rank; c->world = world; c->allreduce = fn; c->comm_user = user;
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
    const bool timing = getenv("ARGUS_TIMING") != nullptr;
    auto tnow = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t_start = tnow();
    // visibility: the host only counts (observations per point, for the sizes and the tile plan); the CSR and the per-point
    // camera bit masks are built on the device from the raw mask bytes (k_index_*; sba_levmar.c:639-650 builds its CRS on the host)
    std::vector<uint8_t> cnt((size_t)n);
    parallel_for(n, [&](int64_t lo, int64_t hi) {
        for (int64_t i = lo; i < hi; ++i) {
            const char *row = vmask + i * m;
            int k = 0;
            for (int j = 0; j < m; ++j) k += row[j] != 0;
            cnt[i] = (uint8_t)k;
        }
    });
    c->n = n; c->ncon = ncon; c->m = m; c->mcon = mcon; c->nfix_k = nccalib; c->nfix_d = ncdist; c->cam_fix_on = false;
    const int mf = m - mcon, N = 16 * mf;
    c->npairs = mf * (mf + 1) / 2;
    c->grid_pts128 = (int)std::min<int64_t>((n + 127) / 128, (int64_t)c->sms * 8); if (c->grid_pts128 < 1) c->grid_pts128 = 1;
    c->grid_pts256 = (int)std::min<int64_t>((n + 255) / 256, (int64_t)c->sms * 8); if (c->grid_pts256 < 1) c->grid_pts256 = 1;
    {
        int want = (c->sms * 8 + c->npairs - 1) / c->npairs;
        int64_t maxc = (n + 63) / 64; if (maxc < 1) maxc = 1;
        if (want > maxc) want = (int)maxc;
        if (want < 1) want = 1;
        c->nchunks = want;
        c->chunk_pts = (n + want - 1) / want; if (c->chunk_pts < 1) c->chunk_pts = 1;
    }
    // ---- generation-2 plans: point tiles (<= 32 points, <= tile_obs observations) and the block -> (group, warp, slot) map ----
    std::vector<int32_t> tiles, tile_o;
    c->tile_obs = 256;
    int64_t nobs = 0;
    {
        int64_t i = 0;
        while (i < n) {
            tiles.push_back((int32_t)i); tile_o.push_back((int32_t)nobs);
            int64_t e = i; int acc = 0;
            while (e < n && e - i < kTilePts && acc + cnt[e] <= c->tile_obs) acc += cnt[e++];
            if (e == i) acc = cnt[e++];
            i = e; nobs += acc;
            if (nobs > 0x7fffffffLL) return ARGUS_E_ARG;
        }
        tiles.push_back((int32_t)n); tile_o.push_back((int32_t)nobs);
        c->ntiles = (int)tiles.size() - 1;
    }
    c->nobs = nobs;
    {
        const int cand[4] = {1, 3, 5, 9};
        int NB = 9, NW = 8;
        for (int q = 0; q < 4; ++q) if (kSchurWarps * cand[q] >= c->npairs) { NB = cand[q]; break; }
        if (NB == 9 && c->schur_nw_pref == 12) { NB = 6; NW = 12; }
        if (NB == 9 && c->schur_nw_pref == 11) { NB = 7; NW = 11; }
        c->schur_NB = NB; c->schur_NW = NW;
        c->schur_G = (c->npairs + NW * NB - 1) / (NW * NB);
        int nch = c->sms / c->schur_G; if (nch < 1) nch = 1;
        if (nch > c->ntiles) nch = c->ntiles > 0 ? c->ntiles : 1;
        c->schur_nchunk = nch;
    }
    c->rank_mp = (m + 15) & ~15;
    // obs_lpt, corder, cstart and pt_rank are derived on the device from obs_ptr / obs_cam / tile_pt (k_index_*, k_build_tile_index)
    c->lin_smem = (((size_t)m * sizeof(CamConst) + 15) & ~(size_t)15) + (size_t)c->tile_obs * (kGa + kBs) * sizeof(double) +
                  (size_t)kTilePts * kPs * sizeof(double);
    {
        int occ = (int)((220 * 1024) / (c->lin_smem + 1024)); if (occ < 1) occ = 1; if (occ > 2) occ = 2;
        c->grid_lin = std::min(c->ntiles > 0 ? c->ntiles : 1, c->sms * occ);
    }
    std::vector<uint16_t> hjk((size_t)c->schur_G * c->schur_NW * c->schur_NB, 0xFFFF);
    std::vector<int32_t> hpair((size_t)c->schur_G * c->schur_NW * c->schur_NB, -1);
    {
        int pidx = 0;
        for (int a = 0; a < mf; ++a)
            for (int b = a; b < mf; ++b, ++pidx) {
                const int grp = pidx % c->schur_G, q = pidx / c->schur_G, w = q % c->schur_NW, slot = q / c->schur_NW;
                const size_t at = ((size_t)grp * c->schur_NW + w) * c->schur_NB + slot;
                hjk[at] = (uint16_t)((mcon + a) | ((mcon + b) << 8));
                hpair[at] = pidx;
            }
    }
    const double t_index = tnow();
    ARGUS_CUDA_CHECK(cudaFuncSetAttribute(k_lin_tile<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->lin_smem));
    ARGUS_CUDA_CHECK(cudaFuncSetAttribute(k_lin_tile<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->lin_smem));
    ARGUS_CUDA_CHECK(cudaFuncSetAttribute(k_lin_tile<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->lin_smem));
    // every device buffer of the problem comes out of ONE allocation: measure, allocate, assign
    auto plan = [&]() -> cudaError_t {
#define PLAN_CHECK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return e_; } while (0)
#define PLAN_ENS(buf, cnt) PLAN_CHECK((buf).ensure((size_t)(cnt)))
    PLAN_CHECK(c->tile_pt.ensure(tiles.size())); PLAN_CHECK(c->tile_o.ensure(tiles.size())); PLAN_CHECK(c->blk_jk.ensure(hjk.size())); PLAN_CHECK(c->blk_pair.ensure(hpair.size()));
    PLAN_CHECK(c->zsign.ensure((size_t)n)); if (c->gen >= 2 && c->mode == kModeMotStr) { PLAN_ENS(c->Z, 48 * nobs); }
    PLAN_CHECK(c->obs_lpt.ensure((size_t)nobs)); PLAN_CHECK(c->corder.ensure((size_t)nobs)); PLAN_CHECK(c->cstart.ensure((size_t)c->ntiles * (m + 1)));
    PLAN_ENS(c->Upart2, (int64_t)c->grid_lin * m * kUfrag); PLAN_ENS(c->cholwork, (int64_t)((N + 31) / 32) * 1024 + N + 32);
    PLAN_CHECK(c->pt_rank.ensure(c->mode == kModeMotStr ? (size_t)c->ntiles * kTilePts * c->rank_mp : 1));
    PLAN_ENS(c->vmask_dev, n * m); PLAN_ENS(c->idx_btot, (n + kIdxPts - 1) / kIdxPts + 1);
    PLAN_ENS(c->obs_ptr, n + 1); PLAN_ENS(c->obs_cam, nobs); PLAN_ENS(c->obs_uv, nobs); PLAN_ENS(c->pt_mask, n);
    PLAN_ENS(c->rot0, 4 * m); PLAN_ENS(c->cam, 16 * m); PLAN_ENS(c->cam_new, 16 * m); PLAN_ENS(c->cam0, 16 * m);
    PLAN_ENS(c->pts, 3 * n); PLAN_ENS(c->pts_new, 3 * n); PLAN_ENS(c->pts0, 3 * n);
    if (c->gen < 3 && c->mode == kModeMotStr) { PLAN_ENS(c->W, 48 * nobs); }
    if (c->gen < 2 && c->mode == kModeMotStr) { PLAN_ENS(c->Y, 48 * nobs); } PLAN_ENS(c->V, 6 * n); PLAN_ENS(c->eb, 3 * n); PLAN_ENS(c->vstate, 3 * n); PLAN_ENS(c->Vinv, 6 * n);
    PLAN_ENS(c->cvec, 3 * n); PLAN_ENS(c->dbuf, 3 * n);
    PLAN_ENS(c->Upart, (int64_t)c->grid_pts128 * m * kU); PLAN_ENS(c->scpart, 4 * std::max(c->grid_pts128, c->grid_lin)); PLAN_ENS(c->red1, m * kU + 2); PLAN_ENS(c->max2, 2);
    PLAN_ENS(c->Epart, (int64_t)c->grid_pts128 * m * 16); PLAN_ENS(c->Spart, (int64_t)std::max(c->nchunks, c->schur_nchunk) * c->npairs * 256);
    PLAN_ENS(c->red2, (int64_t)c->npairs * 256 + m * 16); PLAN_ENS(c->S, (int64_t)N * N); PLAN_ENS(c->E, N); PLAN_ENS(c->da_free, N); PLAN_ENS(c->da_full, 16 * m);
    PLAN_ENS(c->part3, 3 * c->grid_pts128); PLAN_ENS(c->red3, 4); PLAN_ENS(c->cam2, 2); PLAN_ENS(c->scal, 16); c->att.p = c->scal.p + 8; c->att.cap = 8; c->att.owned = false; PLAN_ENS(c->respart, c->grid_pts256);
    PLAN_CHECK(c->flags.ensure(4));
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
            if (ar.base) cudaFree(ar.base);
            ar.base = nullptr; ar.size = 0;
            pe = cudaMalloc(&ar.base, need);
            if (pe == cudaSuccess) ar.size = need;
        }
        if (pe == cudaSuccess) { ar.measuring = false; ar.used = 0; pe = plan(); }
        g_arena = nullptr;
        ARGUS_CUDA_CHECK(pe);
    }
    const double t_alloc = tnow();
    cudaStream_t s = c->stream;
    ARGUS_CUDA_CHECK(cudaMemcpyAsync(c->tile_pt.p, tiles.data(), tiles.size() * sizeof(int32_t), cudaMemcpyHostToDevice, s));
    ARGUS_CUDA_CHECK(cudaMemcpyAsync(c->tile_o.p, tile_o.data(), tile_o.size() * sizeof(int32_t), cudaMemcpyHostToDevice, s));
    ARGUS_CUDA_CHECK(cudaMemcpyAsync(c->blk_jk.p, hjk.data(), hjk.size() * sizeof(uint16_t), cudaMemcpyHostToDevice, s));
    ARGUS_CUDA_CHECK(cudaMemcpyAsync(c->blk_pair.p, hpair.data(), hpair.size() * sizeof(int32_t), cudaMemcpyHostToDevice, s));
    ARGUS_CUDA_CHECK(cudaMemcpyAsync(c->vmask_dev.p, vmask, (size_t)n * m, cudaMemcpyHostToDevice, s));
    if (n > 0) {
        const int nb = (int)((n + kIdxPts - 1) / kIdxPts);
        k_index_count<<<nb, 256, 0, s>>>(n, m, c->vmask_dev.p, c->pt_mask.p, c->idx_btot.p);
        k_index_scan<<<1, 1024, 0, s>>>(nb, c->idx_btot.p);
        k_index_fill<<<nb, 256, 0, s>>>(n, c->pt_mask.p, c->idx_btot.p, c->obs_ptr.p, c->obs_cam.p);
    } else {
        ARGUS_CUDA_CHECK(cudaMemsetAsync(c->obs_ptr.p, 0, sizeof(int32_t), s));
    }
    ARGUS_CUDA_CHECK(cudaMemcpyAsync(c->obs_uv.p, x, nobs * sizeof(double2), cudaMemcpyHostToDevice, s));
    ARGUS_CUDA_CHECK(cudaMemcpyAsync(c->rot0.p, rot0, 4 * m * sizeof(double), cudaMemcpyHostToDevice, s));
    ARGUS_CUDA_CHECK(cudaMemcpyAsync(c->cam0.p, p, 16 * m * sizeof(double), cudaMemcpyHostToDevice, s));
    ARGUS_CUDA_CHECK(cudaMemcpyAsync(c->pts0.p, p + 16 * m, 3 * n * sizeof(double), cudaMemcpyHostToDevice, s));
    // derived static indices, built on the device
    if (n > 0) {
        if (c->mode == kModeMotStr && c->ntiles > 0)
            ARGUS_CUDA_CHECK(cudaMemsetAsync(c->pt_rank.p, 0xFF, (size_t)c->ntiles * kTilePts * c->rank_mp, s));
        if (c->ntiles > 0)
            k_build_tile_index<<<c->ntiles, 256, 0, s>>>(m, c->tile_pt.p, c->tile_o.p, c->obs_ptr.p, c->obs_cam.p, c->obs_lpt.p, c->corder.p,
                                                         c->cstart.p, ncon, c->rank_mp, c->mode == kModeMotStr ? c->pt_rank.p : nullptr);
        ARGUS_CUDA_CHECK(cudaGetLastError());
    }
    ARGUS_CUDA_CHECK(cudaStreamSynchronize(s));      // host staging vectors go out of scope
    c->h2d_bytes += (size_t)n * m + nobs * sizeof(double2) + (4 + 16) * m * sizeof(double) + 3 * n * sizeof(double) +
                    2 * tiles.size() * sizeof(int32_t);
    if (timing) fprintf(stderr, "argus_b200 upload: host counting + tile plan %.1f ms, device alloc %.1f ms, H2D + device index build %.1f ms\n", 1e3 * (t_index - t_start), 1e3 * (t_alloc - t_index), 1e3 * (tnow() - t_alloc));
    return argus_ba_reset(c);
}

int argus_ba_set_mode(argus_ba_ctx *c, int mode)
{
    if (!c || mode < kModeMotStr || mode > kModeStr) return ARGUS_E_ARG;
    if (c->m != 0 && mode != c->mode && (mode == kModeMotStr) != (c->mode == kModeMotStr)) return ARGUS_E_ARG;   // set it before upload
    c->mode = mode;
    return ARGUS_OK;
}

int argus_ba_set_camera_fix(argus_ba_ctx *c, const int *nccalib, const int *ncdist)
{
    if (!c || c->m == 0) return ARGUS_E_ARG;
    ARGUS_CUDA_CHECK(cudaSetDevice(c->device));
    if (!nccalib || !ncdist) { c->cam_fix_on = false; return ARGUS_OK; }
    if (c->gen < 3) return ARGUS_E_ARG;            // only the tile kernels implement the per-camera extension
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

int argus_ba_reset(argus_ba_ctx *c)
{
    if (!c || c->m == 0) return ARGUS_E_ARG;
    ARGUS_CUDA_CHECK(cudaSetDevice(c->device));
    ARGUS_CUDA_CHECK(cudaMemcpyAsync(c->cam.p, c->cam0.p, 16 * c->m * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
    ARGUS_CUDA_CHECK(cudaMemcpyAsync(c->pts.p, c->pts0.p, 3 * c->n * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
    return ARGUS_OK;
}

int argus_ba_download(argus_ba_ctx *c, double *p)
{
    if (!c || !p) return ARGUS_E_ARG;
    ARGUS_CUDA_CHECK(cudaSetDevice(c->device));
    ARGUS_CUDA_CHECK(cudaMemcpyAsync(p, c->cam.p, 16 * c->m * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    ARGUS_CUDA_CHECK(cudaMemcpyAsync(p + 16 * c->m, c->pts.p, 3 * c->n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
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

int argus_ba_project(argus_ba_ctx *c, double *hx)
{
    if (!c || !hx) return ARGUS_E_ARG;
    ARGUS_CUDA_CHECK(cudaSetDevice(c->device));
    DevBuf<double> d;
    ARGUS_CUDA_CHECK(d.ensure(2 * c->nobs));
    int rc = run_residual(c, c->cam.p, c->pts.p, d.p);
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
    if (!c) return ARGUS_E_ARG;
    ARGUS_CUDA_CHECK(cudaSetDevice(c->device));
    const int m = c->m, N = 16 * (m - c->mcon);
    int rc = run_linearize(c);
    if (rc) return rc;
    double sc[8] = {0};
    rc = read_back(c, c->scal.p, 4, sc);
    if (rc) return rc;
    // assemble a copy of S/E before the in-place Cholesky destroys it: run the attempt, but snapshot S, E first
    // (k_chol_solve overwrites S), so split: everything up to assemble, copy, then the rest via a second attempt call.
    rc = run_attempt(c, mu, mu, 0, 0, c->dbuf.p);
    if (rc) return rc;
    double at[4];
    rc = read_back(c, c->att.p, 4, at);
    if (rc) return rc;
    sc[4] = at[1]; sc[5] = at[2]; sc[6] = at[3]; sc[7] = at[0];
    if (scal) memcpy(scal, sc, sizeof(sc));
    cudaStream_t s = c->stream;
    if (U || ea) {
        std::vector<double> h((size_t)m * kU);
        ARGUS_CUDA_CHECK(cudaMemcpyAsync(h.data(), c->red1.p, h.size() * sizeof(double), cudaMemcpyDeviceToHost, s));
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
        if (c->gen >= 2) k_assemble_mma<<<(unsigned)((tot + 255) / 256), 256, 0, s>>>(m, c->mcon, c->red1.p, c->red2.p, c->red2.p + slen, mu, c->S.p, c->E.p, c->gen >= 3);
        else k_assemble<<<(unsigned)((tot + 255) / 256), 256, 0, s>>>(m, c->mcon, c->red1.p, c->red2.p, c->red2.p + slen, mu, c->S.p, c->E.p);
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
    if (c->mode != kModeMotStr) return solve_blockdiag(c, itmax, verbose, opts, info, flags);
    const int m = c->m;
    const int64_t nvis = c->nobs;
    if (c->world == 1) {      // with several ranks the host wrapper checks the global counts
        const int64_t nobs2 = nvis * 2, nvars = (int64_t)m * 16 + c->n * 3;
        if (nobs2 < nvars) {
            fprintf(stderr, "argus_b200: cannot solve a problem with fewer measurements [%lld] than unknowns [%lld]\n",
                    (long long)nobs2, (long long)nvars);
            return ARGUS_E_SBA;
        }
    }
    if (itmax == 0) return 0;   // the reference runs its Jacobian checker here (sba_levmar.c:749-753); see tests instead
    const double tau = fabs(opts[0]), eps1 = fabs(opts[1]), eps2 = fabs(opts[2]), eps2_sq = opts[2] * opts[2],
                 eps3_sq = opts[3] * opts[3], eps4_sq = opts[4] * opts[4];
    double mu = 0.0, eab_inf = 0.0, p_eL2, pdp_eL2, p_L2 = 0.0, dp_L2 = DBL_MAX, dF, dL, dmax = DBL_MIN, init_p_eL2;
    int nu = 2, nu2, stop = 0, nfev = 0, njev = 0, nlss = 0, itno, rc;
    double h[8];

    if ((rc = run_residual(c, c->cam.p, c->pts.p, nullptr))) return rc;
    if ((rc = read_back(c, c->red3.p, 1, h))) return rc;
    p_eL2 = h[0]; nfev = 1;
    if (verbose) printf("initial motstr-SBA error %g [%g]\n", p_eL2, p_eL2 / (double)nvis);
    init_p_eL2 = p_eL2;
    if (!isfinite(p_eL2)) stop = 7;

    for (itno = 0; itno < itmax && !stop; ++itno) {
        // generation 3: once mu is known the tile kernel linearises AND prepares the first attempt (Z, E) in one pass
        bool z_valid = false, deferred = false;
        if (c->gen >= 3 && itno > 0) { rc = run_lin_tile(c, 1, mu, 0, compat ? 1 : 0); z_valid = true; deferred = true; }
        else rc = run_linearize(c);
        if (rc) return rc;
        ++njev;
        if (!deferred) {
            if ((rc = read_back(c, c->scal.p, 4, h))) return rc;
            eab_inf = h[1]; p_L2 = h[2]; dmax = h[3];
            if (eab_inf <= eps1) { dp_L2 = 0.0; stop = 1; break; }
            if (itno == 0) mu = tau * dmax;
        }
        double mu_u = 0.0;          // damping accumulated on U's diagonal within this iteration (compat: not restored)
        bool first = true;
        while (1) {
            mu_u = compat ? mu_u + mu : mu;
            // generation 3: Z / E depend on mu, so the tile linearisation runs once per attempt with the damping applied
            if ((rc = run_attempt(c, mu, mu_u, compat && !first, compat, nullptr, !z_valid))) return rc;
            z_valid = false;
            if (deferred) {
                // mu was known, so the first attempt was launched behind the linearisation without a host round trip; the
                // linearisation's scalars arrive with the attempt's (scal and att are one buffer).  When the gradient test
                // stops the iteration here, the attempt's results are simply never looked at.
                double hh[16];
                if ((rc = read_back(c, c->scal.p, 16, hh))) return rc;
                deferred = false;
                eab_inf = hh[1]; p_L2 = hh[2]; dmax = hh[3];
                if (eab_inf <= eps1) { dp_L2 = 0.0; stop = 1; break; }
                for (int q = 0; q < 8; ++q) h[q] = hh[8 + q];
            } else if ((rc = read_back(c, c->att.p, 5, h))) return rc;
            first = false;
            const bool singular = h[4] != 0.0;
            if (singular) fprintf(stderr, "argus_b200: singular matrix V*_i, increasing damping\n");
            if (!singular) {
                ++nlss;
                const bool issolved = h[0] != 0.0;
                if (issolved) {
                    dp_L2 = h[1];
                    if (dp_L2 <= eps2_sq * p_L2) { stop = 2; break; }
                    if (dp_L2 >= (p_L2 + eps2) / 1e-24) {
                        fprintf(stderr, "argus_b200: the matrix of the augmented normal equations is almost singular,\n"
                                        "     minimization should be restarted from the current solution with an increased damping term\n");
                        c->profiling ? prof_collect(c) : (void)0;
                        return ARGUS_E_SBA;
                    }
                    ++nfev;
                    pdp_eL2 = h[3];
                    if (!isfinite(pdp_eL2)) { stop = 7; break; }
                    dL = h[2];
                    dF = p_eL2 - pdp_eL2;
                    if (verbose > 1)
                        printf("\ndamping term %8g, gain ratio %8g, errors %8g / %8g = %g\n", mu, dL != 0.0 ? dF / dL : dF / DBL_EPSILON,
                               p_eL2 / nvis, pdp_eL2 / nvis, p_eL2 / pdp_eL2);
                    if (dL > 0.0 && dF > 0.0) {
                        double tmp = (2.0 * dF / dL - 1.0);
                        tmp = 1.0 - tmp * tmp * tmp;
                        mu = mu * ((tmp >= 0.3333333334) ? tmp : 0.3333333334);
                        nu = 2;
                        if (pdp_eL2 - 2.0 * sqrt(p_eL2 * pdp_eL2) < (eps4_sq - 1.0) * p_eL2) stop = 4;
                        std::swap(c->cam.p, c->cam_new.p); std::swap(c->cam.cap, c->cam_new.cap);
                        std::swap(c->pts.p, c->pts_new.p); std::swap(c->pts.cap, c->pts_new.cap);
                        p_eL2 = pdp_eL2;
                        break;
                    }
                }
            }
            // rejected: the linear system could not be solved or the error did not decrease
            mu *= nu;
            nu2 = (int)((unsigned)nu << 1);
            if (nu2 <= nu) {
                fprintf(stderr, "argus_b200: too many failed attempts to increase the damping factor! Singular Hessian matrix?\n");
                stop = 6;
                break;
            }
            nu = nu2;
        }
        if (stop == 1) break;       // gradient test (deferred read-back): leaves the outer loop like the reference's break (:977-981)
        if (p_eL2 <= eps3_sq) stop = 5;
    }
    if (itno >= itmax) stop = 3;
    if (c->profiling) prof_collect(c);
    if (info) {
        info[0] = init_p_eL2; info[1] = p_eL2; info[2] = eab_inf; info[3] = dp_L2; info[4] = mu / dmax;
        info[5] = itno; info[6] = stop; info[7] = nfev; info[8] = njev; info[9] = nlss;
    }
    return (stop != 7) ? itno : ARGUS_E_SBA;
}

int argus_ba_motstr_kdrts(int64_t n, int64_t ncon, int m, int mcon, const char *vmask, double *p, const double *rot0,
                          const double *x, int nccalib, int ncdist, int itmax, int verbose, const double opts[5], double info[10],
                          int flags)
{
    const bool timing = getenv("ARGUS_TIMING") != nullptr;
    auto tnow = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t0 = tnow();
    argus_ba_ctx *c = nullptr;
    int rc = argus_ba_create(&c, 0, nullptr);
    if (rc) return rc;
    const double t1 = tnow();
    rc = argus_ba_upload(c, n, ncon, m, mcon, vmask, p, rot0, x, nccalib, ncdist);
    const double t2 = tnow();
    double t3 = t2, t4 = t2;
    int ret = rc;
    if (rc == ARGUS_OK) {
        ret = argus_ba_solve(c, itmax, verbose, opts, info, flags);
        t3 = tnow();
        if (ret >= 0 || ret == ARGUS_E_SBA) { rc = argus_ba_download(c, p); if (rc) ret = rc; }
        t4 = tnow();
    }
    argus_ba_destroy(c);
    if (timing) fprintf(stderr, "argus_b200 one-shot: create %.1f ms, upload %.1f ms, solve %.1f ms, download %.1f ms, destroy %.1f ms\n",
                        1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * (t3 - t2), 1e3 * (t4 - t3), 1e3 * (tnow() - t4));
    return ret;
}

namespace {
int blockdiag_oneshot(int mode, int64_t n, int64_t ncon, int m, int mcon, const char *vmask, double *moving, const double *frozen,
                      const double *rot0, const double *x, int nccalib, int ncdist, int itmax, int verbose, const double opts[5],
                      double info[10], int flags)
{
    if (!moving || !frozen || n < 0 || m < 1) return ARGUS_E_ARG;
    std::vector<double> p((size_t)16 * m + (size_t)3 * n);
    const bool mot = mode == kModeMot;
    memcpy(p.data(), mot ? moving : frozen, (size_t)16 * m * sizeof(double));
    memcpy(p.data() + 16 * m, mot ? frozen : moving, (size_t)3 * n * sizeof(double));
    argus_ba_ctx *c = nullptr;
    int rc = argus_ba_create(&c, 0, nullptr);
    if (rc) return rc;
    rc = argus_ba_set_mode(c, mode);
    if (rc == ARGUS_OK) rc = argus_ba_upload(c, n, ncon, m, mcon, vmask, p.data(), rot0, x, nccalib, ncdist);
    int ret = rc;
    if (rc == ARGUS_OK) {
        ret = argus_ba_solve(c, itmax, verbose, opts, info, flags);
        if (ret >= 0 || ret == ARGUS_E_SBA) {
            rc = argus_ba_download(c, p.data());
            if (rc) ret = rc;
            else memcpy(moving, mot ? p.data() : p.data() + 16 * m, (mot ? (size_t)16 * m : (size_t)3 * n) * sizeof(double));
        }
    }
    argus_ba_destroy(c);
    return ret;
}
}  // namespace

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
