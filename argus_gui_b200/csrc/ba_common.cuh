// Shared declarations of the BA device path (problem layout in HBM, launch helpers).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "camera_model.cuh"

namespace argus {

// ---- problem layout in HBM (one rank's shard; cameras replicated) ----------------------------------
//   observations are point-major, cameras ascending inside a point: exactly the order of the
//   reference's measurement vector x (sba_levmar.c:466-470) so hx / e need no permutation.
struct BaProblem {
    int64_t n;        // points on this rank
    int64_t ncon;     // leading fixed points on this rank
    int     m;        // cameras (all ranks)
    int     mcon;     // leading fixed cameras
    int64_t nobs;     // visible projections on this rank
    int     nfix_k, nfix_d;   // nccalib / ncdist (all cameras)
    const uint8_t *cam_fix;   // optional per-camera override: [2j] = nccalib_j, [2j+1] = ncdist_j (extension; nullptr = global)

    const int32_t  *obs_ptr;  // n+1: first observation of each point
    const uint8_t  *obs_cam;  // nobs: camera of each observation
    const double2  *obs_uv;   // nobs: measured pixel
    const uint64_t *pt_mask;  // n: bit j set <=> point seen by camera j
    const double   *rot0;     // 4m
    const int32_t  *wand_of;  // optional (extension, ba_wand.cuh): n entries, index of the wand pair a point belongs to or -1; nullptr = no pairs
};

// ---- Levenberg-Marquardt state, resident on the device ------------------------------------------------
// The damping loop of sba_motstr_levmar_x / sba_mot_levmar_x / sba_str_levmar_x (sba_levmar.c:791-1357, :1694-1950,
// :2239-2495) runs as a single-thread controller kernel at the end of every damping attempt (ba_lm_controller.cuh); every
// other kernel of an attempt reads what it needs (damping, which parameter buffer is current, "the loop has ended") from
// this block.  The host therefore never sits between two attempts: it queues attempts (a CUDA graph each) ahead of the
// device and only looks at asynchronous snapshots of this block to know when to stop queueing.
struct LmState {
    // inputs of the next attempt
    double mu;            // damping term
    double mu_u;          // what the attempt adds to U's (and, block-diagonal drivers, V's) diagonal: mu, or in sba-1.6 mode
                          // the sum of the dampings of the attempts rejected in this iteration (diagonals never restored)
    int use_state;        // sba-1.6 mode retry: V's diagonal is what the previous inversion left behind
    int write_state;
    int cur_cam, cur_pts; // which buffer of each parameter pair holds the current estimate
    int done;             // the loop has ended: every queued kernel returns immediately
    int fresh;            // the next attempt linearises at a new point (counts as a Jacobian evaluation, gradient stop test)
    // controller
    int mode, compat, itmax, itno, stop, nu, nfev, njev, err, nsingular, nlog, logcap;
    double nlss, nblocks;
    double tau, eps1, eps2, eps2_sq, eps3_sq, eps4_sq;
    double p_eL2, init_p_eL2, eab_inf, p_L2, dmax, dp_L2;
    double *lmlog;        // optional per-attempt log (verbose > 1), kLmLogDoubles per entry
};

// the two copies of the parameter vector (current estimate / trial point); LmState::cur_* says which is which
struct ParamBufs {
    double *cam[2];
    double *pts[2];
};
// selection by a run-time index through a conditional: indexing the kernel parameter itself would copy the table to local memory
__device__ __forceinline__ double *pb_cam(const ParamBufs &b, int which) { return which ? b.cam[1] : b.cam[0]; }
__device__ __forceinline__ double *pb_pts(const ParamBufs &b, int which) { return which ? b.pts[1] : b.pts[0]; }

constexpr int kU = 152;       // per camera: 136 upper-triangle entries of U_j + 16 of ea_j

// index of (r,c), r<=c, in the packed upper triangle of a 16x16 symmetric block
__host__ __device__ __forceinline__ int tri16(int r, int c) { return r * 16 - (r * (r - 1)) / 2 + (c - r); }

#define ARGUS_CUDA_CHECK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
    fprintf(stderr, "argus_b200: CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
    return ARGUS_E_CUDA; } } while (0)

// deterministic block sum (all threads get the result); blockDim.x <= 1024, multiple of 32
__device__ __forceinline__ double block_sum(double v, double *sh /* >= 32 */)
{
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = blockDim.x >> 5;
    __syncthreads();
    if (l == 0) sh[w] = v;
    __syncthreads();
    double r = 0.0;
    if (w == 0) {
        r = (l < nw) ? sh[l] : 0.0;
        for (int o = 16; o > 0; o >>= 1) r += __shfl_down_sync(0xffffffffu, r, o);
        if (l == 0) sh[0] = r;
    }
    __syncthreads();
    r = sh[0];
    return r;
}

__device__ __forceinline__ double block_max(double v, double *sh)
{
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_down_sync(0xffffffffu, v, o));
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = blockDim.x >> 5;
    __syncthreads();
    if (l == 0) sh[w] = v;
    __syncthreads();
    double r = -1.79769313486231570e308;
    if (w == 0) {
        r = (l < nw) ? sh[l] : -1.79769313486231570e308;
        for (int o = 16; o > 0; o >>= 1) r = fmax(r, __shfl_down_sync(0xffffffffu, r, o));
        if (l == 0) sh[0] = r;
    }
    __syncthreads();
    r = sh[0];
    return r;
}

}  // namespace argus
