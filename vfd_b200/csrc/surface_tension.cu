// surface_tension.cu — Monte-Carlo surface tension (Zorilla et al. 2020) as the reference implements it.
//
// Replaces ComputeSurfaceTensionClassificationKernel (reference: DFSPHKernels.cu:881-958),
// ComputeSurfaceTensionNormalsAndCurvatureKernel :960-1006, ComputeSurfaceTensionBlendingKernel
// :1008-1045 and DFSPHImplementation::SolveSurfaceTension (DFSPHImplementation.cu:810-839).
// The 16 384-point Halton sphere table (reference data file HaltonVec323.cuh) is regenerated on the
// host (tables.cpp) and read through the read-only path.  The normal and the curvature of a particle
// are written as one float4 so the smoothing pass gathers both with a single 16-B load.
#include "solver.h"
#include "tile.cuh"
#include <algorithm>

namespace vfd {

extern __shared__ __align__(128) unsigned char smemRaw[];
#define FOR_EACH_TILE(p) for (uint32_t p = blockIdx.x * blockDim.x + threadIdx.x; p < P.n; p += gridDim.x * blockDim.x)

__device__ __forceinline__ float3 normalize_if_nonzero(float3 v) {   // DFSPHKernels.cu:874-879
    if (dot3(v, v) > 0.0f) v = normalize3(v);
    return v;
}

// T1: classify surface particles by the centre of mass of their neighbourhood; for those, estimate
// the normal and curvature from the Halton samples on the support sphere not covered by a neighbour.
struct StClassifyOp {
    typedef float4 Payload;
    static constexpr bool READ_COUNT = true;
    const Params& P; const Arrays& A;
    const float* __restrict__ halton;
    uint32_t sampleCount;
    float mcFactor, cover2;
    __device__ __forceinline__ float4 load(uint32_t g) const { return A.posRho[g]; }
    template<class Acc>
    __device__ __forceinline__ void particle(uint32_t p, uint32_t m, size_t ell, const Acc& acc) {
        const float3 xi = f3(A.posRho[p]);
        const uint16_t* col = A.list16 + ell;
        float3 n = f3(0.0f, 0.0f, 0.0f);
        float curv = A.curv[p];             // left untouched for interior particles
        if (m == 0u) {
            curv = 1.0f / P.h;
        } else {
            float3 com = f3(0.0f, 0.0f, 0.0f);
            for (uint32_t k = 0; k < m; k++) com += f3(acc(col[(size_t)k * 32])) - xi;
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
                        const float3 dir = f3(acc(col[(size_t)k * 32])) - xi;
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
};

__global__ void __launch_bounds__(TILE_THREADS) k_st_classify(const __grid_constant__ Params P, const __grid_constant__ Arrays A, const DevState* __restrict__ S, const float* __restrict__ halton) {
    const float radiusRatio = P.nbrRadius / P.r;
    StClassifyOp op{ P, A, halton, S->sampleCount, S->mcFactor, radiusRatio * radiusRatio * P.h2 };
    tile_pass(S, A.cellBegin, A.cnt, smem_header(smemRaw), reinterpret_cast<float4*>(smemRaw + smem_header_bytes()), STAGE_CAP16, op);
}

// T2: neighbour-weighted smoothing of normal and curvature among surface particles
struct StSmoothOp {
    typedef Pay32 Payload;           // position, (normal, curvature)
    static constexpr bool READ_COUNT = true;
    const Params& P; const Arrays& A;
    __device__ __forceinline__ Pay32 load(uint32_t g) const { return Pay32{ A.posRho[g], A.nrm[g] }; }
    template<class Acc>
    __device__ __forceinline__ void particle(uint32_t p, uint32_t m, size_t ell, const Acc& acc) {
        const float tau = P.smoothing;
        const float4 ni4 = A.nrm[p];
        if (ni4.x != 0.0f || ni4.y != 0.0f || ni4.z != 0.0f) {
            const float3 xi = f3(A.posRho[p]);
            const uint16_t* col = A.list16 + ell;
            float3 nc = f3(0.0f, 0.0f, 0.0f);
            float cc = 0.0f, wsum = 0.0f;
            for (uint32_t k = 0; k < m; k++) {
                const Pay32 nb = acc(col[(size_t)k * 32]);
                const float4 nj = nb.b;
                if (nj.x != 0.0f || nj.y != 0.0f || nj.z != 0.0f) {
                    const float3 d = f3(nb.a) - xi;
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
};

__global__ void __launch_bounds__(TILE_THREADS) k_st_smooth(const __grid_constant__ Params P, const __grid_constant__ Arrays A, const DevState* __restrict__ S) {
    StSmoothOp op{ P, A };
    tile_pass(S, A.cellBegin, A.cnt, smem_header(smemRaw), reinterpret_cast<Pay32*>(smemRaw + smem_header_bytes()), STAGE_CAP32, op);
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
    const size_t s1 = smem_header_bytes() + (size_t)STAGE_CAP16 * sizeof(float4), s2 = smem_header_bytes() + (size_t)STAGE_CAP32 * sizeof(Pay32);
    static thread_local int per1 = 0, per2 = 0;
    if (!per1) {
        cudaFuncSetAttribute(k_st_classify, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s1);
        cudaFuncSetAttribute(k_st_smooth, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s2);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per1, k_st_classify, TILE_THREADS, s1);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per2, k_st_smooth, TILE_THREADS, s2);
        per1 = std::max(per1, 1); per2 = std::max(per2, 1);
    }
    { LaunchScope ls(L, KID_ST_CLASSIFY); k_st_classify<<<per1 * L.numSMs, TILE_THREADS, s1, L.stream>>>(P, A, S, halton); }
    { LaunchScope ls(L, KID_ST_SMOOTH); k_st_smooth<<<per2 * L.numSMs, TILE_THREADS, s2, L.stream>>>(P, A, S); }
    for (uint32_t i = 0; i < passes; i++) { LaunchScope ls(L, KID_ST_APPLY); k_st_apply<<<tiles, VFD_TPB, 0, L.stream>>>(P, A); }
}

} // namespace vfd
