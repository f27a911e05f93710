// The Levenberg-Marquardt damping loop as device code: one single-thread kernel at the end of every damping attempt takes
// the decisions the reference takes on the host between its linear solves (sba_motstr_levmar_x sba_levmar.c:941-987 and
// :1267-1357; the motion-only / structure-only drivers :1774-1950 / :2319-2495 have the same loop around block-diagonal
// solves) and leaves the inputs of the next attempt in LmState.  One controller serves the three drivers.
//
// What it reproduces, in the reference's order:
//   after a linearisation at a new point:  ||J^T e||_inf <= eps1 -> stop 1 (the attempt queued behind that linearisation
//                                          is then simply not looked at); first iteration: mu = tau * max diag
//   after the solve:   singular V*_i -> more damping (not counted as a linear system); Cholesky failed -> more damping;
//                      ||dp||^2 <= eps2^2 ||p||^2 -> stop 2;  ||dp||^2 >= (||p||^2 + eps2) / 1e-24 -> error return;
//                      non-finite trial cost -> stop 7;  dL > 0 and dF > 0 -> accept: mu *= max(1/3, 1 - (2 dF/dL - 1)^3),
//                      nu = 2, stop 4 test, trial point becomes current;  otherwise mu *= nu, nu *= 2 (overflow -> stop 6)
//   after an iteration: ||e||^2 <= eps3^2 -> stop 5;  itno == itmax -> stop 3 (which replaces any other reason reached in
//                      the last iteration, as in the reference where it is tested after the loop)
// sba-1.6 mode (ARGUS_BA_COMPAT_SBA16): after a rejected step the damping of the new attempt is added on top of what the
// rejected attempts left on the diagonals (mu_u accumulates, V's diagonal state is reused) - SURVEY.md appendix B Q1/Q2.
#pragma once
#include "ba_common.cuh"
#include "ba_blockdiag.cuh"

namespace argus {

constexpr int kLmLogDoubles = 6;     // per attempt: mu, dL, dF, ||e||^2 before, ||e||^2 at the trial point, kind (0 solved, 1 singular V*, 2 Cholesky failed)

struct LmInit {
    double opts[5];
    double nblocks;          // block systems per attempt (block-diagonal drivers: info[9] counts every one)
    int itmax, compat, mode, cur_cam, cur_pts, logcap;
    double *lmlog;
};

// start of a solve: zero the counters, keep which buffers are current
__global__ void k_lm_init(LmState *st, LmInit in, int *flags)
{
    if (threadIdx.x || blockIdx.x) return;
    st->mu = 0.0; st->mu_u = 0.0; st->use_state = 0; st->write_state = in.compat;
    st->cur_cam = in.cur_cam; st->cur_pts = in.cur_pts; st->done = 0; st->fresh = 1;
    st->mode = in.mode; st->compat = in.compat; st->itmax = in.itmax; st->itno = 0; st->stop = 0; st->nu = 2;
    st->nfev = 0; st->njev = 0; st->err = 0; st->nsingular = 0; st->nlog = 0; st->logcap = in.logcap; st->lmlog = in.lmlog;
    st->nlss = 0.0; st->nblocks = in.nblocks;
    st->tau = fabs(in.opts[0]); st->eps1 = fabs(in.opts[1]); st->eps2 = fabs(in.opts[2]); st->eps2_sq = in.opts[2] * in.opts[2];
    st->eps3_sq = in.opts[3] * in.opts[3]; st->eps4_sq = in.opts[4] * in.opts[4];
    st->p_eL2 = 0.0; st->init_p_eL2 = 0.0; st->eab_inf = 0.0; st->p_L2 = 0.0; st->dmax = 2.2250738585072014e-308; st->dp_L2 = 1.79769313486231570e308;
    flags[0] = 0; flags[1] = 0; flags[2] = 0; flags[3] = 0;
}

// parity hooks: one attempt with a given damping at the resident point, no loop
__global__ void k_lm_set_attempt(LmState *st, double mu, int cur_cam, int cur_pts, int mode, int *flags)
{
    if (threadIdx.x || blockIdx.x) return;
    st->mu = mu; st->mu_u = mu; st->use_state = 0; st->write_state = 0; st->cur_cam = cur_cam; st->cur_pts = cur_pts; st->done = 0;
    st->fresh = 0; st->mode = mode; st->compat = 0; st->err = 0; st->logcap = 0; st->lmlog = nullptr;
    flags[0] = 0; flags[1] = 0; flags[2] = 0; flags[3] = 0;
}

__device__ __forceinline__ void lm_finish(LmState *st)
{
    if (st->itno >= st->itmax) st->stop = 3;     // tested after the loop in the reference: it wins over a reason set in the last iteration
    st->done = 1;
}
// the inner loop was left (stop 2 / 4 / 6 / 7 or an accepted step): end of the iteration body
__device__ __forceinline__ void lm_end_iteration(LmState *st)
{
    if (st->p_eL2 <= st->eps3_sq) st->stop = 5;
    st->itno += 1;
    if (st->stop || st->itno >= st->itmax) { lm_finish(st); return; }
    st->fresh = 1; st->mu_u = st->mu; st->use_state = 0;
}

// after the initial cost (red3[0]) and the first linearisation (scal): the part of iteration 0 that precedes the inner loop
__global__ void k_lm_begin(LmState *st, const double *__restrict__ red3, const double *__restrict__ scal)
{
    if (threadIdx.x || blockIdx.x || st->done) return;
    st->p_eL2 = red3[0]; st->init_p_eL2 = red3[0]; st->nfev = 1;
    if (!isfinite(st->p_eL2)) { st->stop = 7; st->done = 1; return; }          // the loop is never entered (itno stays 0)
    st->njev = 1;
    st->eab_inf = scal[1]; st->p_L2 = scal[2]; st->dmax = scal[3];
    if (st->eab_inf <= st->eps1) { st->dp_L2 = 0.0; st->stop = 1; lm_finish(st); return; }
    st->mu = st->tau * st->dmax;
    st->mu_u = st->mu; st->use_state = 0; st->fresh = 0;
}

// att3: {||e(p + dp)||^2, ||db||^2, dL over the points, singular-V* flag} (summed over ranks); cam2: {||da||^2, dL over the cameras}
__global__ void k_lm_controller(LmState *st, int *flags, const double *__restrict__ cam2, const double *__restrict__ att3,
                                const double *__restrict__ scal)
{
    if (threadIdx.x || blockIdx.x || st->done) return;
    const int mode = st->mode;
    if (st->fresh) {                      // this attempt linearised at a new point
        st->njev += 1;
        st->eab_inf = scal[1]; st->p_L2 = scal[2]; st->dmax = scal[3];
        st->fresh = 0;
        if (st->eab_inf <= st->eps1) { st->dp_L2 = 0.0; st->stop = 1; lm_finish(st); return; }     // `break` out of the outer loop: itno is not advanced
    }
    const bool singular = att3[3] != 0.0;
    const int chol_ok = flags[1];
    flags[0] = 0; flags[2] = 0;           // ready for the next attempt
    bool solved;
    int kind = 0;
    if (mode == kModeMotStr) {
        if (singular) { st->nsingular += 1; solved = false; kind = 1; }
        else { st->nlss += 1.0; solved = chol_ok != 0; if (!solved) kind = 2; }
    } else {
        st->nlss += st->nblocks;
        solved = (mode == kModeMot) ? (chol_ok != 0) : !singular;
        if (!solved) kind = 2;
    }
    const double mu = st->mu;
    double dL = 0.0, dF = 0.0, pdp = 0.0;
    bool accepted = false;
    if (solved) {
        const double dp_L2 = (mode == kModeMot) ? cam2[0] : (mode == kModeStr) ? att3[1] : cam2[0] + att3[1];
        st->dp_L2 = dp_L2;
        if (dp_L2 <= st->eps2_sq * st->p_L2) { st->stop = 2; lm_end_iteration(st); return; }
        if (dp_L2 >= (st->p_L2 + st->eps2) / 1e-24) { st->err = 1; st->done = 1; return; }
        st->nfev += 1;
        pdp = att3[0];
        if (!isfinite(pdp)) { st->stop = 7; lm_end_iteration(st); return; }
        dL = (mode == kModeMot) ? cam2[1] : (mode == kModeStr) ? att3[2] : cam2[1] + att3[2];
        dF = st->p_eL2 - pdp;
        accepted = dL > 0.0 && dF > 0.0;
    }
    if (st->lmlog && st->nlog < st->logcap) {
        double *e = st->lmlog + (size_t)st->nlog * kLmLogDoubles;
        e[0] = mu; e[1] = dL; e[2] = dF; e[3] = st->p_eL2; e[4] = pdp; e[5] = (double)kind;
        st->nlog += 1;
    }
    if (accepted) {
        double t = 2.0 * dF / dL - 1.0;
        t = 1.0 - t * t * t;
        st->mu = mu * ((t >= 0.3333333334) ? t : 0.3333333334);
        st->nu = 2;
        if (pdp - 2.0 * sqrt(st->p_eL2 * pdp) < (st->eps4_sq - 1.0) * st->p_eL2) st->stop = 4;
        if (mode != kModeStr) st->cur_cam ^= 1;
        if (mode != kModeMot) st->cur_pts ^= 1;
        st->p_eL2 = pdp;
        lm_end_iteration(st);
        return;
    }
    // rejected: the linear system could not be solved or the error did not decrease
    st->mu = mu * (double)st->nu;
    const int nu2 = (int)((unsigned)st->nu << 1);
    if (nu2 <= st->nu) { st->stop = 6; lm_end_iteration(st); return; }
    st->nu = nu2;
    st->mu_u = st->compat ? st->mu_u + st->mu : st->mu;
    st->use_state = st->compat;
}

// fixed-order sum of the per-block partials {trial cost, ||db||^2, dL} of the back-substitution + the singular-V* flag as a
// double, so that one all-reduce carries the four numbers on which every rank must take the same decision
__global__ void k_reduce_attempt(const double *__restrict__ part3, int nparts, const int *__restrict__ flags, double *__restrict__ att3,
                                 const int *__restrict__ done)
{
    if (*done) return;
    const int t = threadIdx.x;
    if (t < 3) {
        double s = 0.0;
        for (int c = 0; c < nparts; ++c) s += part3[3 * c + t];
        att3[t] = s;
    } else if (t == 3) att3[3] = (double)flags[0];
}

// motion-only driver: {trial cost} from the residual partials, the rest of att3 is zero
__global__ void k_reduce_cost(const double *__restrict__ part, int nparts, double *__restrict__ att3, const int *__restrict__ done)
{
    if (*done) return;
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int c = 0; c < nparts; ++c) s += part[c];
        att3[0] = s;
    } else if (threadIdx.x < 4) att3[threadIdx.x] = 0.0;
}

}  // namespace argus
