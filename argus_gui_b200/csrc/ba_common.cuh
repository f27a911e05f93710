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
};

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
