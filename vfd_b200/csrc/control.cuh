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

// Per-rank control block in peer-mapped memory: what the ranks of a decomposition write into EACH OTHER over NVLink.
//   red / redFlag   all-reduce of a solver decision's two doubles: rank r stores its partial result into slot [seq & 3][r] of
//                   every rank's block, then the sequence number into redFlag; every rank sums the eight entries in rank order
//                   (the same value everywhere) as soon as all flags show the number — no collective launch, no host;
//   haloReady/Data  hand-shake of a halo exchange with the left [0] / right [1] neighbour (distributed.cu: k_halo_exchange).
struct PeerCtl {
    double red[4][8][2];
    uint32_t redFlag[4][8];
    uint32_t haloReady[2], haloData[2];
    uint32_t redSeq, haloSeq;             // the owner's own sequence numbers (never reset while the slab lives: a restarted bake goes on counting)
    uint32_t pad[26];
};

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

// All-reduce (sum or max) of two doubles over the ranks through peer memory; called by ONE thread per rank, in the same
// sequence on every rank (the solver's control flow is identical everywhere).  A rank can be at most one reduction ahead of
// another (it needs the other's contribution to finish), so four slots used in turn never collide.
__device__ __forceinline__ void peer_allreduce(const Params& P, DevState* S, double (&v)[2], bool isMax) {
    PeerCtl* me = reinterpret_cast<PeerCtl*>(P.peerCtl[P.rank]);
    const uint32_t seq = me->redSeq + 1u, slot = seq & 3u;
    me->redSeq = seq;
    for (uint32_t r = 0; r < P.nRanks; r++) {
        volatile double* d = reinterpret_cast<PeerCtl*>(P.peerCtl[r])->red[slot][P.rank];
        d[0] = v[0]; d[1] = v[1];
    }
    __threadfence_system();
    for (uint32_t r = 0; r < P.nRanks; r++) *(volatile uint32_t*)&reinterpret_cast<PeerCtl*>(P.peerCtl[r])->redFlag[slot][P.rank] = seq;
    for (uint32_t r = 0; r < P.nRanks; r++) while (*(volatile uint32_t*)&me->redFlag[slot][r] != seq) { }
    __threadfence_system();
    double a = isMax ? -DBL_MAX : 0.0, b = a;
    for (uint32_t r = 0; r < P.nRanks; r++) {
        const volatile double* d = me->red[slot][r];
        a = isMax ? fmax(a, d[0]) : a + d[0];
        b = isMax ? fmax(b, d[1]) : b + d[1];
    }
    v[0] = a; v[1] = b;
}

// called by thread 0 of the last block with the folded result
template<int NV>
__device__ __forceinline__ void finish_reduction(int site, const Params& P, DevState* S, const double (&tot)[NV], bool isMax = false) {
    if (P.nRanks > 1 && P.peerCtl[0]) {
        double v[2] = { tot[0], NV > 1 ? tot[NV - 1] : 0.0 };
        peer_allreduce(P, S, v, isMax);
        control_site(site, P, S, v);
    } else if (P.nRanks > 1) {
        #pragma unroll
        for (int q = 0; q < NV; q++) S->red[site * 2 + q] = tot[q];
        if (NV < 2) S->red[site * 2 + 1] = 0.0;
    } else {
        control_site(site, P, S, tot);
    }
}
#endif

} // namespace vfd
