// Host build of the product camera model (argus_gui_b200/csrc/camera_model.cuh) so that the from-scratch
// projection / Jacobian can be checked on CPU against the compiled reference (tests/test_camera_model.py).
#include "../../argus_gui_b200/csrc/camera_model.cuh"
#include <stdint.h>

extern "C" {

// p = [16 per camera | 3 per point]; obs k is (pt[k], cam[k]); hx is nobs x 2
void host_project(const double *p, const double *rot0, int m, int64_t nobs, const int32_t *pt, const int32_t *cam, double *hx)
{
    const double *pb = p + 16 * m;
    for (int64_t k = 0; k < nobs; ++k) {
        argus::CamConst c;
        argus::cam_precompute(p + 16 * cam[k], rot0 + 4 * cam[k], c);
        const double *M = pb + 3 * (int64_t)pt[k];
        argus::cam_project(c, M[0], M[1], M[2], hx[2 * k], hx[2 * k + 1]);
    }
}

void host_jacobian(const double *p, const double *rot0, int m, int64_t nobs, const int32_t *pt, const int32_t *cam,
                   int nccalib, int ncdist, double *hx, double *A, double *B)
{
    const double *pb = p + 16 * m;
    for (int64_t k = 0; k < nobs; ++k) {
        argus::CamConst c;
        argus::cam_precompute(p + 16 * cam[k], rot0 + 4 * cam[k], c);
        const double *M = pb + 3 * (int64_t)pt[k];
        argus::cam_project_jac(c, M[0], M[1], M[2], nccalib, ncdist, hx[2 * k], hx[2 * k + 1], A + 32 * k, B + 6 * k);
    }
}

}
