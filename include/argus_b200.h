/* argus_b200.h - C ABI of the B200-native bundle-adjustment + DLT hot path.
 *
 * Every entry point is `extern "C"`, takes plain pointers and sizes (no torch / C++ types) and is what a
 * binding on the reference side (python-sba's ctypes layer, or argus_gui/tools.py) would bind instead of
 * the reference's own native layer.  INTEGRATION.md shows those bindings.
 *
 * Reference interfaces replaced (R/ = /root/reference, SBA/ = inside
 * R/argus_gui/resources/sba_lib_sources.tar.gz, directory sba-1.6p/):
 *
 *   argus_ba_motstr_kdrts      <- sba_motstr_levmar_x           SBA/sba.h:111-116, SBA/sba_levmar.c:449-1428
 *                                 with func/fjac = img_projsKDRTS_x / img_projsKDRTS_jac_x
 *                                 (SBA/libsbaprojs/sbaprojs.h:113-114, sbaprojs.c:1135-1250) and
 *                                 adata = struct globs_ {rot0params, nccalib, ncdist, cnp=16, pnp=3, mnp=2}
 *                                 (sbaprojs.h:123-156); call site R/argus_gui/sbaDriver.py:292 through
 *                                 python-sba's sba.SparseBundleAdjust.
 *   argus_ba_mot_kdrts         <- sba_mot_levmar_x              SBA/sba.h:119-123, SBA/sba_levmar.c:1437-1980
 *                                 with img_projsKDRT_x / _jac_x (sbaprojs.c:1261-1370), globs_.ptparams
 *   argus_ba_str_kdrts         <- sba_str_levmar_x              SBA/sba.h:126-130, SBA/sba_levmar.c:1987-2540
 *                                 with img_projsKDS_x / _jac_x (sbaprojs.c:1375-1445), globs_.camparams
 *   argus_dlt_triangulate      <- uv_to_xyz                     R/argus_gui/tools.py:296-337
 *   argus_dlt_reproj_error     <- get_repo_errors               R/argus_gui/tools.py:360-423
 *   argus_dlt_reconstruct      <- both of the above in one pass (call sites argus-click:3283-3305, utils/DLCbatch.py:646-666)
 *   argus_undistort_points     <- undistort_pts (pinhole)       R/argus_gui/tools.py:129-148
 *   argus_dlt_bootstrap        <- bootstrapXYZs, the 250-iteration loop      R/argus_gui/tools.py:448-544 (:513-527)
 *   argus_dlt_fit              <- solve_dlt / OutlierWindow.getCoefficients   tools.py:220-280, sbaDriver.py:1027-1079
 *   argus_pair_triangulate     <- Triangulator.triangulate      R/argus_gui/triangulate.py:27-212
 *   argus_ba_create ... argus_ba_normal_equations: the resident-data form of the three LM drivers (no counterpart in the
 *                                 reference, which re-allocates everything per call) and parity hooks for the tests
 *   argus_ba_comm_init_nccl / argus_ba_set_comm_nccl / argus_nccl_unique_id: the multi-GPU exchange (one all-reduce of the
 *                                 reduced camera system + one of four scalars per damping attempt) through NCCL, which the
 *                                 library binds at run time; the reference is single-process and has no counterpart
 *
 * Ownership: the caller owns every array; outputs are written into caller-allocated memory.  No entry point
 * calls exit(); errors are negative return codes (ARGUS_E_*).  There is no CPU fallback: without a CUDA
 * device every compute entry point returns ARGUS_E_CUDA.
 */
#ifndef ARGUS_B200_H_
#define ARGUS_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ARGUS_OK            0
#define ARGUS_E_SBA        -1   /* == SBA_ERROR (SBA/sba.h:62): the reference's failure return            */
#define ARGUS_E_CUDA       -2   /* CUDA runtime error / no device                                         */
#define ARGUS_E_ARG        -3   /* bad argument (m > 64, null pointer, ...)                               */
#define ARGUS_E_COMM       -4   /* the all-reduce callback reported an error                              */

#define ARGUS_BA_CNP 16   /* camera parameters: fu,u0,v0,ar,s | k1,k2,p1,p2,k3 | local rot (3) | t (3) */
#define ARGUS_BA_PNP 3
#define ARGUS_BA_MNP 2
#define ARGUS_BA_MAX_CAMS 64

/* flags for argus_ba_solve / argus_ba_motstr_kdrts */
#define ARGUS_BA_COMPAT_SBA16  1  /* emulate sba 1.6's damping behaviour after a rejected step: U/V diagonals are
                                     not restored (SBA/sba_levmar.c:1339-1353 is #if 0) and V's diagonal holds the
                                     inverse's diagonal (SBA/sba_lapack.c:1118-1121).  Without it the diagonals are
                                     restored (textbook LM). */
#define ARGUS_BA_PROFILE       2  /* record per-phase CUDA-event times (argus_ba_get_profile): every fourth damping
                                     attempt is replayed from a graph with an event-record node on either side of every
                                     phase, after the host has seen the outcome of the attempts before it; the profile
                                     holds ONE sample per phase and solve (the last such attempt), plus the once-per-solve
                                     phases (initial cost, statistics-only linearisation)                              */

typedef struct argus_ba_ctx argus_ba_ctx;

/* All-reduce hook for the multi-GPU path (one process per GPU).  `buf` is DEVICE memory holding `count`
 * doubles; op 0 = sum, 1 = max; `stream` is the cudaStream_t the solver enqueues on - the hook must order
 * the collective after work already enqueued on it and make later work on it wait for the result
 * (torch.distributed.all_reduce issued under torch.cuda.stream(ExternalStream(stream)) does exactly that).
 * Return 0 on success. */
typedef int (*argus_allreduce_fn)(void *user, double *buf, int64_t count, int op, void *stream);

/* ---- one-shot, host pointers: the drop-in for sba_motstr_levmar_x + img_projsKDRTS_x/_jac_x ----
 * n, ncon, m, mcon, vmask (n x m, non-zero = visible), p ([16 per camera | 3 per point], updated in place),
 * x (2 per visible projection, point-major, cameras ascending), itmax, verbose, opts[5], info[10]: exactly
 * the reference's arguments (SBA/sba_levmar.c:449-518).  rot0 (4 per camera, scalar first), nccalib, ncdist:
 * the fields of struct globs_.  cnp/pnp/mnp are fixed at 16/3/2 and covx is not supported (Argus passes NULL).
 * Returns the number of iterations (>= 0) or ARGUS_E_*; info[] as the reference fills it. */
int argus_ba_motstr_kdrts(int64_t n, int64_t ncon, int m, int mcon, const char *vmask, double *p,
                          const double *rot0, const double *x, int nccalib, int ncdist,
                          int itmax, int verbose, const double opts[5], double info[10], int flags);

/* ---- motion-only and structure-only bundle adjustment (SBA/sba.h:118-130) ----
 * argus_ba_mot_kdrts: the drop-in for sba_mot_levmar_x (SBA/sba_levmar.c:1437-1980) + img_projsKDRT_x/_jac_x
 * (libsbaprojs/sbaprojs.c:1261-1370): cams (16 per camera) is updated in place, pts (3 per point) is what the reference
 * passes in globs_.ptparams.
 * argus_ba_str_kdrts: the drop-in for sba_str_levmar_x (:1987-2540) + img_projsKDS_x/_jac_x (sbaprojs.c:1375-1445): pts is
 * updated in place, cams is globs_.camparams.  The reference's KDS callbacks read entries 10..12 of a camera row as the
 * vector part of the camera's FULL rotation quaternion; that is rot0 = (1,0,0,0) per camera here, while a general rot0
 * keeps the (rot0, local rotation) form the motstr driver returns.  info[9] counts every block system solved. */
int argus_ba_mot_kdrts(int64_t n, int m, int mcon, const char *vmask, double *cams, const double *pts,
                       const double *rot0, const double *x, int nccalib, int ncdist,
                       int itmax, int verbose, const double opts[5], double info[10], int flags);
int argus_ba_str_kdrts(int64_t n, int64_t ncon, int m, const char *vmask, double *pts, const double *cams,
                       const double *rot0, const double *x,
                       int itmax, int verbose, const double opts[5], double info[10], int flags);

/* CUDA device the one-shot calls above run on (default 0) */
int argus_set_default_device(int device);

/* ---- resident-data interface (what the one-shot call is built from; used by bench.py and the multi-GPU host) ---- */
int  argus_ba_create(argus_ba_ctx **ctx, int device, void *stream /* cudaStream_t or NULL */);
void argus_ba_destroy(argus_ba_ctx *ctx);
/* Multi-GPU (one context per GPU, points sharded, cameras replicated; ncon/mcon passed to upload are LOCAL counts).
 * Preferred: NCCL inside the library.  argus_nccl_unique_id writes the 128-byte ncclUniqueId the ranks must share (rank 0
 * creates it, the host distributes it by whatever means it has); argus_ba_comm_init_nccl creates this rank's communicator
 * from it (collective over the ranks: ncclCommInitRank) on the context's device; argus_ba_set_comm_nccl adopts a
 * communicator the caller owns (an ncclComm_t).  With NCCL a damping attempt is one CUDA graph including its two
 * collectives.  Alternative: argus_ba_set_comm with a caller-supplied all-reduce hook (used by the CPU / single-GPU tests
 * of the plumbing).  world == 1 (default) needs neither. */
int  argus_nccl_unique_id(void *out128);
int  argus_ba_comm_init_nccl(argus_ba_ctx *ctx, int rank, int world, const void *unique_id128);
int  argus_ba_set_comm_nccl(argus_ba_ctx *ctx, int rank, int world, void *nccl_comm);
/* the context's communicator (created by argus_ba_comm_init_nccl or adopted), so that further contexts of the same process
 * can share it through argus_ba_set_comm_nccl instead of paying ncclCommInitRank again; NULL if none */
void *argus_ba_get_comm_nccl(const argus_ba_ctx *ctx);
int  argus_ba_set_comm(argus_ba_ctx *ctx, int rank, int world, argus_allreduce_fn fn, void *user);
/* a damping attempt is replayed as a CUDA graph from the first solve after an upload (default on); 0 launches kernel by kernel */
int  argus_ba_set_graph(argus_ba_ctx *ctx, int on);
/* copy a problem (or this rank's shard of the points, with all cameras) to the device and build the indices */
int  argus_ba_upload(argus_ba_ctx *ctx, int64_t n, int64_t ncon, int m, int mcon, const char *vmask,
                     const double *p, const double *rot0, const double *x, int nccalib, int ncdist);
/* which parameters move: ARGUS_BA_MODE_MOTSTR (default), _MOT (points frozen) or _STR (cameras frozen); call before
 * argus_ba_upload (p keeps the [16 per camera | 3 per point] layout in every mode) */
#define ARGUS_BA_MODE_MOTSTR 0
#define ARGUS_BA_MODE_MOT    1
#define ARGUS_BA_MODE_STR    2
int  argus_ba_set_mode(argus_ba_ctx *ctx, int mode);
/* EXTENSION (no counterpart in the reference, whose nccalib / ncdist are global, sbaprojs.c:1229-1247): per-camera numbers
 * of fixed trailing intrinsics / distortion coefficients, m entries each, after argus_ba_upload; NULL, NULL restores the global
 * values.  BASELINE config 5 ("mixed fixed/free intrinsics") needs it. */
int  argus_ba_set_camera_fix(argus_ba_ctx *ctx, const int *nccalib, const int *ncdist);
/* EXTENSION, parity unpinned (the reference has no counterpart: it applies the wand length after the adjustment, as one global
 * scale, argus_gui/sbaDriver.py:703-713, and reports the spread as the wand score, :779-781): wand-length constraint inside
 * the solver.  Every pair (ia[p], ib[p]) of points of the uploaded problem (indices local to this rank, a point in at most one
 * pair) adds the residual  weight * (length[p] - ||X_ia - X_ib||)  to the cost (weight: pixels per length unit).  Call
 * after argus_ba_upload, motion + structure mode only; npairs = 0 or weight = 0 switches it off, and the solver is then bit
 * for bit the unconstrained one.  info[0], info[1] then include the wand residuals. */
int  argus_ba_set_wand(argus_ba_ctx *ctx, int64_t npairs, const int64_t *ia, const int64_t *ib, const double *length, double weight);
/* restore the parameter vector uploaded last (device-side copy), e.g. between timed repetitions */
int  argus_ba_reset(argus_ba_ctx *ctx);
int  argus_ba_solve(argus_ba_ctx *ctx, int itmax, int verbose, const double opts[5], double info[10], int flags);
int  argus_ba_download(argus_ba_ctx *ctx, double *p);
int64_t argus_ba_num_observations(const argus_ba_ctx *ctx);

/* profile of the last solve (ARGUS_BA_PROFILE): names[i] (static strings), summed ms and number of samples per phase;
 * returns the number of entries written (<= cap). */
int  argus_ba_get_profile(argus_ba_ctx *ctx, const char **names, double *ms, int64_t *launches, int cap);

/* counters since creation: out = {kernel launches, all-reduce calls, host->device bytes, device->host bytes} */
int  argus_ba_get_counters(argus_ba_ctx *ctx, int64_t out[4]);

/* ---- parity hooks (the pieces of one LM attempt, for tests against the oracle) ---- */
/* hx (2 per projection) at the resident p  (img_projsKDRTS_x) */
int  argus_ba_project(argus_ba_ctx *ctx, double *hx);
/* One linearisation + one damping attempt at the resident p with damping mu.  Any output may be NULL.
 * U (m x 16 x 16), ea (m x 16), V (n x 6, upper triangle 00 01 02 11 12 22), eb (n x 3),
 * S (Sdim x Sdim, Sdim = 16 (m - mcon), symmetric), E (Sdim), dp (16 m + 3 n), scal[8] =
 * {||e||^2, ||J^T e||_inf, ||p||^2, max diag, ||dp||^2, dL, ||e(p+dp)||^2, solved flag}. */
int  argus_ba_normal_equations(argus_ba_ctx *ctx, double mu, double *U, double *ea, double *V, double *eb,
                               double *S, double *E, double *dp, double *scal);

/* ---- DLT path ---- */
/* pts: N x (2 m k): for every frame, k track blocks of 2m values (u,v per camera, NaN = not seen) - the layout of
 * the Clicker/DLC point matrix (argus-click:3287 slices one track block per uv_to_xyz call; k = 1 reproduces that
 * call exactly).  prof: m x 8 [f,cx,cy,k1,k2,p1,p2,k3]; dlt: m x 11.
 * xyz: N x 3k (NaN where fewer than two cameras see the point or a component is exactly 0, as the reference).
 * flags bit 0: textbook 2n-row system instead of the reference's (n+1)-row system (tools.py:317-328). */
int  argus_dlt_triangulate(const double *pts, int64_t N, int m, int k, const double *prof, const double *dlt,
                           double *xyz, int flags);
/* xyz: N x 3k, pts: N x 2mk, err: k x N (tools.py:360-423) */
int  argus_dlt_reproj_error(const double *xyz, const double *pts, int64_t N, int m, int k,
                            const double *prof, const double *dlt, double *err);
/* both in one host round trip (pts copied to the device once): xyz N x 3k, err k x N */
int  argus_dlt_reconstruct(const double *pts, int64_t N, int m, int k, const double *prof, const double *dlt,
                           double *xyz, double *err, int flags);
/* Bootstrap standard deviations of the triangulated points (the 250-iteration loop of bootstrapXYZs, tools.py:513-530):
 * rmses N x k, sd N x 3k (sample standard deviation, ddof = 1, over bs_iter perturbed re-triangulations).  noise: optional
 * explicit standard-normal numbers [k][bs_iter][N][2m] (the reference's np.random.randn draw order) for exact parity tests;
 * NULL = Philox generator on the device seeded with `seed`. */
int  argus_dlt_bootstrap(const double *pts, int64_t N, int m, int k, const double *prof, const double *dlt, const double *rmses,
                         int bs_iter, const double *noise, uint64_t seed, double *sd);
/* pts: N x 2 pixels -> out: N x 2 undistorted pixels, float32-faithful to cv2.undistortPoints(P=K) (tools.py:129-148) */
int  argus_undistort_points(const double *pts, int64_t N, const double *prof8, float *out);

/* device-pointer variants (inputs already resident in HBM; used for the kernel-only timings) */
int  argus_dlt_triangulate_dev(const double *d_pts, int64_t N, int m, int k, const double *d_prof, const double *d_dlt,
                               double *d_xyz, int flags, void *stream);
int  argus_dlt_reconstruct_dev(const double *d_pts, int64_t N, int m, int k, const double *d_prof, const double *d_dlt,
                               double *d_xyz, double *d_err, int flags, void *stream);
int  argus_dlt_reproj_error_dev(const double *d_xyz, const double *d_pts, int64_t N, int m, int k,
                                const double *d_prof, const double *d_dlt, double *d_err, void *stream);

const char *argus_b200_version(void);

/* Measured fp64 throughput of `device` in TFLOP/s (2 flop per FMA): mma.sync m8n8k4 f64 (the tensor-pipe form the Schur
 * kernel uses) and plain DFMA, each run back to back for about `seconds`.  bench.py's roofline denominators. */
int argus_fp64_peak(int device, double seconds, double *dmma_tflops, double *dfma_tflops);

/* ---- two-view initialisation (what feeds the bundle adjustment) ----
 * One camera paired with the reference camera: Triangulator(p1, p2, f1, f2, c1, c2, dist1, dist2).triangulate() of
 * argus_gui/triangulate.py:27-212 as one call.  p1 / p2: N x 2 pixel coordinates in camera k / in the reference camera (N >= 8,
 * no NaN); prof1 / prof2: [f, cx, cy, k1, k2, p1, p2, k3].  Outputs: xyz (N x 3, in the reference camera's frame), R (3 x 3
 * row-major, Triangulator.R before its Gram-Schmidt), t (3, Triangulator.t), F (3 x 3, optional, may be NULL: the 8-point
 * fundamental matrix of cv2.findFundamentalMat(p2, p1, FM_8POINT)). */
int argus_pair_triangulate(const double *p1, const double *p2, int64_t N, const double *prof1, const double *prof2,
                           double *xyz, double *R, double *t, double *F);

/* ---- 11-parameter DLT fit of one camera (the step after the bundle adjustment) ----
 * tools.solve_dlt (argus_gui/tools.py:220-280) / OutlierWindow.getCoefficients (argus_gui/sbaDriver.py:1027-1075):
 * xyz (N x 3), uv (N x 2 undistorted pixels; NaN in u - or, with flags & 1, in u or v - skips the point),
 * L (11) = least-squares DLT coefficients, errors (N, optional, may be NULL): reprojection distance of every used point
 * (NaN for skipped ones), stats = {points used, sum of distances, mean, population standard deviation}.
 * Returns ARGUS_E_SBA when the system is rank deficient. */
int argus_dlt_fit(const double *xyz, const double *uv, int64_t N, int flags, double *L, double *errors, double *stats);

#ifdef __cplusplus
}
#endif
#endif /* ARGUS_B200_H_ */
