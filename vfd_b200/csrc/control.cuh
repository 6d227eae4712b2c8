// control.cuh — what the solver does with a finished grid-wide reduction: the loop-carried decisions of the
// reference's host loops (DFSPHImplementation.cu: ComputeDivergence :506-575, ComputePressure :443-504,
// ComputeTimeStepSize :395-428, SolveViscosity :617-808), taken on the device.
//
// On one GPU the last block of the reducing kernel calls control_site() directly.  With several ranks the
// last block only publishes its rank's partial result in DevState::red; an NCCL all-reduce combines the
// ranks' values and k_control (solver.cu) then takes the decision — identically on every rank.
#pragma once
#include "common.cuh"

namespace vfd {

enum ReductionSite { SITE_DIV = 0, SITE_PRESS, SITE_CFL, SITE_VISC_BB, SITE_VISC_INIT, SITE_VISC_PQ, SITE_VISC_UPDATE, SITE_COUNT };

#ifdef __CUDACC__
__device__ __forceinline__ void control_site(int site, const Params& P, DevState* S, const double* tot) {
    switch (site) {
    case SITE_DIV: {
        // the reference folds with thrust::minus from 0 (DFSPHImplementation.cu:554-560): as a left fold that is
        // -(sum), the mean of rho0*|residuum|
        const float err = (float)(-tot[0]) / (float)P.nGlobal;
        const uint32_t it = S->divIt + 1;
        const float eta = S->dtInv * P.divErrScale;
        S->divIt = it; S->divErr = err;
        S->divActive = ((err > eta || it < P.minDivIt) && it < P.maxDivIt) ? 1u : 0u;
        break; }
    case SITE_PRESS: {
        const float err = (float)(-tot[0]) / (float)P.nGlobal;      // :483-489
        const uint32_t it = S->pressIt + 1;
        S->pressIt = it; S->pressErr = err;
        S->pressActive = ((err > P.etaPressure || it < P.minPressIt) && it < P.maxPressIt) ? 1u : 0u;
        break; }
    case SITE_CFL: {
        float vmax2 = fmaxf((float)tot[0], 0.1f);              // initial value 0.1 (DFSPHImplementation.cu:397)
        if (vmax2 < 1.0e-9f) vmax2 = 1.0e-9f;
        float ndt = 0.4f * (P.d / sqrtf(vmax2));
        ndt = fminf(ndt, P.maxDt);
        ndt = fmaxf(ndt, P.minDt);
        S->vmax2 = vmax2;
        S->dt = ndt; S->dt2 = ndt * ndt; S->dtInv = 1.0f / ndt; S->dt2Inv = 1.0f / (ndt * ndt);
        S->sampleCount = P.csdFix > 0 ? (uint32_t)P.csdFix : (uint32_t)(int)((float)P.csd * ndt);
        S->mcFactor = P.mcFactor;
        S->frameTime += ndt;
        S->stepCount += 1;
        break; }
    case SITE_VISC_BB:
        S->rhsNorm2 = (float)tot[0];
        S->viscIt = 0;
        break;
    case SITE_VISC_INIT: {
        const float rhs = S->rhsNorm2;
        const float rr = (float)tot[0];
        S->resNorm2 = rr;
        S->delta = fabsf((float)tot[1]);
        if (rhs == 0.0f) {
            S->viscActive = 2u;          // g := 0, error 0 (DFSPHImplementation.cu:656-661); applied by k_visc_apply
            S->viscErr = 0.0f;
        } else {
            const float thr = fmaxf(P.viscErr2 * rhs, FLT_MIN);
            S->threshold = thr;
            S->viscErr = sqrtf(rr / rhs);
            // loop condition "it >= Min && it < Max" at it = 0 (:693)
            S->viscActive = (!(rr < thr) && 0u >= P.minViscIt && 0u < P.maxViscIt) ? 1u : 0u;
        }
        break; }
    case SITE_VISC_PQ:
        S->alpha = S->delta / (float)tot[0];
        break;
    case SITE_VISC_UPDATE: {
        const float rr = (float)tot[0];
        S->resNorm2 = rr;
        S->viscErr = sqrtf(rr / S->rhsNorm2);
        if (rr < S->threshold) {
            S->viscActive = 0u;                       // break: this iteration is not counted (:770-772)
        } else {
            const float dNew = fabsf((float)tot[1]);
            S->beta = dNew / S->delta;
            S->delta = dNew;
            const uint32_t it = S->viscIt + 1;
            S->viscIt = it;
            S->viscActive = (it >= P.minViscIt && it < P.maxViscIt) ? 1u : 0u;
        }
        break; }
    default: break;
    }
}

// called by thread 0 of the last block with the folded result
template<int NV>
__device__ __forceinline__ void finish_reduction(int site, const Params& P, DevState* S, const double (&tot)[NV]) {
    if (P.nRanks > 1) {
        #pragma unroll
        for (int q = 0; q < NV; q++) S->red[site * 2 + q] = tot[q];
        if (NV < 2) S->red[site * 2 + 1] = 0.0;
    } else {
        control_site(site, P, S, tot);
    }
}
#endif

} // namespace vfd
