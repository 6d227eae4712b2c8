// surface_tension.cu — Monte-Carlo surface tension (Zorilla et al. 2020) as the reference implements it.
//
// Replaces ComputeSurfaceTensionClassificationKernel (reference: DFSPHKernels.cu:881-958),
// ComputeSurfaceTensionNormalsAndCurvatureKernel :960-1006, ComputeSurfaceTensionBlendingKernel
// :1008-1045 and DFSPHImplementation::SolveSurfaceTension (DFSPHImplementation.cu:810-839).
// The 16 384-point Halton sphere table (reference data file HaltonVec323.cuh) is regenerated on the
// host (tables.cpp) and read through the read-only path.  The normal and the curvature of a particle
// are written as one float4 so the smoothing pass gathers both with a single 16-B load.
#include "solver.h"
#include <algorithm>

namespace vfd {

#define FOR_EACH_TILE(p) for (uint32_t p = blockIdx.x * blockDim.x + threadIdx.x; p < P.n; p += gridDim.x * blockDim.x)

__device__ __forceinline__ float3 normalize_if_nonzero(float3 v) {   // DFSPHKernels.cu:874-879
    if (dot3(v, v) > 0.0f) v = normalize3(v);
    return v;
}

// T1: classify surface particles by the centre of mass of their neighbourhood; for those, estimate
// the normal and curvature from the Halton samples on the support sphere not covered by a neighbour.
__global__ void __launch_bounds__(VFD_TPB) k_st_classify(Params P, Arrays A, const DevState* __restrict__ S, const float* __restrict__ halton) {
    const float4* __restrict__ pos = A.posRho;
    const uint32_t sampleCount = S->sampleCount;
    const float mcFactor = S->mcFactor;
    const float radiusRatio = P.nbrRadius / P.r;
    const float cover2 = radiusRatio * radiusRatio * P.h2;
    FOR_EACH_TILE(p) {
        const float3 xi = f3(pos[p]);
        const uint32_t m = A.cnt[p];
        const uint32_t* col = nbr_column(A.list, p);
        float3 n = f3(0.0f, 0.0f, 0.0f);
        float curv = A.curv[p];             // left untouched for interior particles
        if (m == 0u) {
            curv = 1.0f / P.h;
        } else {
            float3 com = f3(0.0f, 0.0f, 0.0f);
            for (uint32_t k = 0; k < m; k++) {
                const uint32_t j = col[(size_t)k * 32];
                com += f3(pos[j]) - xi;
            }
            com = com / P.h;
            const float clsIn = sqrtf(dot3(com, com)) / (float)m;
            const float onLine = P.clsSlope * clsIn + P.clsConst + 0.0f;
            if ((float)m <= onLine) {
                uint32_t kept = 0u;
                // the sample window depends on the particle's *original* index (DFSPHKernels.cu:911)
                const uint32_t s = A.id[p] * sampleCount / 3u * 3u;
                for (uint32_t q = 0; q < sampleCount; q++) {
                    const uint32_t i3 = s + 3u * q;
                    const float3 pt = P.h * f3(__ldg(halton + i3 % VFD_HALTON_N), __ldg(halton + (i3 + 1u) % VFD_HALTON_N), __ldg(halton + (i3 + 2u) % VFD_HALTON_N));
                    bool covered = false;
                    for (uint32_t k = 0; k < m; k++) {
                        const uint32_t j = col[(size_t)k * 32];
                        const float3 dir = f3(pos[j]) - xi;
                        const float3 v = pt - dir;
                        if (dot3(v, v) <= cover2) { covered = true; break; }
                    }
                    if (!covered) { n += pt; kept++; }
                }
                if (kept > 0u) {
                    n = normalize_if_nonzero(n);
                    curv = 1.0f / P.h * -2.0f * sqrtf(1.0f - P.nbrRadius * P.nbrRadius / (P.r * P.r)) *
                           cosf(acosf(1.0f - 2.0f * ((float)kept / (float)sampleCount)) + mcFactor);
                } else {
                    n = f3(0.0f, 0.0f, 0.0f);
                    curv = 0.0f;
                }
            }
        }
        A.curv[p] = curv;
        A.nrm[p] = make_float4(n.x, n.y, n.z, curv);
    }
}

// T2: neighbour-weighted smoothing of normal and curvature among surface particles
__global__ void __launch_bounds__(VFD_TPB) k_st_smooth(Params P, Arrays A) {
    const float4* __restrict__ pos = A.posRho;
    const float4* __restrict__ nrm = A.nrm;
    const float tau = P.smoothing;
    FOR_EACH_TILE(p) {
        const float4 ni4 = nrm[p];
        if (ni4.x != 0.0f || ni4.y != 0.0f || ni4.z != 0.0f) {
            const float3 xi = f3(pos[p]);
            const uint32_t m = A.cnt[p];
            const uint32_t* col = nbr_column(A.list, p);
            float3 nc = f3(0.0f, 0.0f, 0.0f);
            float cc = 0.0f, wsum = 0.0f;
            for (uint32_t k = 0; k < m; k++) {
                const uint32_t j = col[(size_t)k * 32];
                const float4 nj = nrm[j];
                if (nj.x != 0.0f || nj.y != 0.0f || nj.z != 0.0f) {
                    const float3 d = f3(pos[j]) - xi;
                    const float dist = sqrtf(dot3(d, d));
                    const float w = 1.0f - dist / P.h;
                    nc += f3(nj) * w;
                    cc += nj.w * w;
                    wsum += w;
                }
            }
            nc = normalize_if_nonzero(nc);
            float3 ns = (1.0f - tau) * f3(ni4) + tau * nc;
            ns = normalize_if_nonzero(ns);
            A.nbar[p] = make_float4(ns.x, ns.y, ns.z, 0.0f);
            A.curvS[p] = ((1.0f - tau) * ni4.w + tau * cc) / (1.0f - tau + tau * wsum);
        }
    }
}

// T3: apply the force (once per smoothing pass: SURVEY.md Q18)
__global__ void __launch_bounds__(VFD_TPB) k_st_apply(Params P, Arrays A) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P.n) return;
    const float4 n = A.nrm[p];
    if (n.x != 0.0f || n.y != 0.0f || n.z != 0.0f) {
        const float4 fn = A.nbar[p];
        float c = A.curvS[p];
        if (P.temporalSmoothing) c = 0.05f * c + 0.95f * A.curvD[p];
        const float3 force = f3(fn) * P.sigma * c;      // (n * sigma) * c
        float4 a = A.acc[p];
        a.x -= P.massInv * force.x; a.y -= P.massInv * force.y; a.z -= P.massInv * force.z;
        A.acc[p] = a;
        A.curvD[p] = c;
    } else {
        float c = 0.0f;
        if (P.temporalSmoothing) c = 0.95f * A.curvD[p];
        A.curvD[p] = c;
    }
}

void launch_surface_tension(const LaunchCfg& L, const Params& P, const Arrays& A, const DevState* S, const float* halton, uint32_t passes) {
    const uint32_t tiles = (P.n + VFD_TPB - 1) / VFD_TPB;
    const uint32_t g = std::max(1u, std::min<uint32_t>(tiles, (uint32_t)L.numSMs * 6u));
    { LaunchScope ls(L, KID_ST_CLASSIFY); k_st_classify<<<g, VFD_TPB, 0, L.stream>>>(P, A, S, halton); }
    { LaunchScope ls(L, KID_ST_SMOOTH); k_st_smooth<<<g, VFD_TPB, 0, L.stream>>>(P, A); }
    for (uint32_t i = 0; i < passes; i++) { LaunchScope ls(L, KID_ST_APPLY); k_st_apply<<<tiles, VFD_TPB, 0, L.stream>>>(P, A); }
}

} // namespace vfd
