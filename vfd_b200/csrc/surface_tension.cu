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
#define OWNED_INDEX(p) uint32_t ownB_, ownE_; owned_range(P, A.cellBegin, ownB_, ownE_); \
    const uint32_t p = ownB_ + blockIdx.x * blockDim.x + threadIdx.x; if (p >= ownE_) return

__device__ __forceinline__ float3 normalize_if_nonzero(float3 v) {   // DFSPHKernels.cu:874-879
    if (dot3(v, v) > 0.0f) v = normalize3(v);
    return v;
}

// T1: classify surface particles by the centre of mass of their neighbourhood; for those, estimate
// the normal and curvature from the Halton samples on the support sphere not covered by a neighbour.
//
// Two phases per tile.  A: every lane classifies its own particle (one pass over its neighbours); surface particles go into
// the tile's work queue (tile.cuh: TileQueue).  B: the consumer warps of the CTA take surface particles from the queue one at
// a time and process each WITH THE WHOLE WARP — lane l tests samples l, l+32, ... against the particle's neighbours
// (warp-uniform shared-memory reads: broadcasts) — instead of one lane running SampleCount x m distance tests while 31 lanes
// idle, and instead of the few warps whose batches lie in the surface shell doing all of it.  The uncovered samples are
// summed in sample order, as the reference's sequential loop does, so the result is bit-identical to it.
struct StClassifyOp {
    static constexpr bool CUSTOM = true, TILE_QUEUE = true;
    using Cfg = PipeCfg<24, 2, 2>;
    static constexpr int NPAY = 1, BBYTES = 0, COEF = 0, NRED = 0, NLUT = 0;
    const Params& P; const Arrays& A;
    const float* __restrict__ halton;
    uint32_t sampleCount;
    float mcFactor, cover2;
    __device__ __forceinline__ const float4* srcA() const { return A.posRho; }
    __device__ __forceinline__ const void* srcB() const { return nullptr; }
    __device__ __forceinline__ void prefetch_own(uint32_t, uint32_t, uint32_t) const {}
    __device__ __forceinline__ float4 loadA(uint32_t g) const { return A.posRho[g]; }
    __device__ __forceinline__ float4 loadB(uint32_t) const { return make_float4(0.0f, 0.0f, 0.0f, 0.0f); }
    // phase A; returns true when the particle is a surface particle (to be queued)
    template<class Acc>
    __device__ __forceinline__ bool particle(uint32_t p, bool valid, const Acc& acc, const StageHeader&) {
        if (!valid) return false;
        const uint32_t m = A.cnt[p] & VFD_COUNT_MASK;
        const float3 xi = f3(A.posRho[p]);
        float curv = A.curv[p];           // left untouched for interior particles
        bool surface = false;
        if (m == 0u) {
            curv = 1.0f / P.h;
        } else {
            const uint2* col = ell_list(A.list16, p);                  // group g (slots 4g..4g+3): col[g * 32]
            float3 com = f3(0.0f, 0.0f, 0.0f);
            const uint32_t nG = (m + 3u) >> 2;
            uint2 w = __ldg(col);
            for (uint32_t g = 0; g < nG; g++) {
                const uint2 wn = __ldg(col + (size_t)min(g + 1u, nG - 1u) * 32);      // next group in flight
                uint32_t L[4];
                ell_unpack(w, L);
                float4 x[4];
                #pragma unroll
                for (int u = 0; u < 4; u++) x[u] = acc(L[u]);           // unused slots of the last group hold index 0: a valid read
                #pragma unroll
                for (int u = 0; u < 4; u++) if (g * 4u + (uint32_t)u < m) com += f3(x[u]) - xi;      // in list order
                w = wn;
            }
            com = com / P.h;
            const float clsIn = sqrtf(dot3(com, com)) / (float)m;
            const float onLine = P.clsSlope * clsIn + P.clsConst + 0.0f;
            surface = (float)m <= onLine;
        }
        if (!surface) {
            A.curv[p] = curv;
            A.nrm[p] = make_float4(0.0f, 0.0f, 0.0f, curv);
        }
        return surface;
    }
    // phase B: one surface particle, by all 32 lanes
    template<class Acc>
    __device__ __forceinline__ void item(uint32_t pS, const Acc& acc, const StageHeader&) {
        const int lane = threadIdx.x & 31;
        const uint32_t mS = __ldg(A.cnt + pS) & VFD_COUNT_MASK;
        const float3 xS = f3(A.posRho[pS]);
        // the sample window depends on the particle's *original* index (DFSPHKernels.cu:911)
        const uint32_t sS = A.id[pS] * sampleCount / 3u * 3u;
        // the particle's whole list in registers: lane g holds group g (18 groups), handed round by shuffle
        const uint32_t nG = (mS + 3u) >> 2;
        uint2 myW = make_uint2(0u, 0u);
        if ((uint32_t)lane < nG) myW = __ldg(ell_list(A.list16, pS) + (size_t)lane * 32);
        float3 nS = f3(0.0f, 0.0f, 0.0f);
        uint32_t kept = 0u;
        for (uint32_t q0 = 0; q0 < sampleCount; q0 += 32u) {
            const uint32_t q = q0 + (uint32_t)lane;
            float3 pt = f3(0.0f, 0.0f, 0.0f);
            bool open = false;
            if (q < sampleCount) {
                const uint32_t i3 = sS + 3u * q;
                pt = P.h * f3(__ldg(halton + i3 % VFD_HALTON_N), __ldg(halton + (i3 + 1u) % VFD_HALTON_N), __ldg(halton + (i3 + 2u) % VFD_HALTON_N));
                open = true;
            }
            // every lane walks the same neighbour list (uniform addresses); a lane's sample drops out once covered
            for (uint32_t g = 0; g < nG; g++) {
                uint32_t L[4];
                ell_unpack(make_uint2(__shfl_sync(0xffffffffu, myW.x, g), __shfl_sync(0xffffffffu, myW.y, g)), L);
                float4 x[4];
                #pragma unroll
                for (int u = 0; u < 4; u++) x[u] = acc(L[u]);
                #pragma unroll
                for (int u = 0; u < 4; u++) {
                    if (g * 4u + (uint32_t)u < mS) {
                        const float3 dir = f3(x[u]) - xS;
                        const float3 v = pt - dir;
                        if (dot3(v, v) <= cover2) open = false;
                    }
                }
                if (!__any_sync(0xffffffffu, open)) break;
            }
            uint32_t keep = __ballot_sync(0xffffffffu, open);
            kept += __popc(keep);
            // n += pt in sample order (the reference's sequential sum)
            while (keep) {
                const int l = __ffs(keep) - 1;
                keep &= keep - 1u;
                nS.x += __shfl_sync(0xffffffffu, pt.x, l);
                nS.y += __shfl_sync(0xffffffffu, pt.y, l);
                nS.z += __shfl_sync(0xffffffffu, pt.z, l);
            }
        }
        if (lane == 0) {
            float3 n = f3(0.0f, 0.0f, 0.0f);
            float curv = 0.0f;
            if (kept > 0u) {
                n = normalize_if_nonzero(nS);
                curv = 1.0f / P.h * -2.0f * sqrtf(1.0f - P.nbrRadius * P.nbrRadius / (P.r * P.r)) *
                       cosf(acosf(1.0f - 2.0f * ((float)kept / (float)sampleCount)) + mcFactor);
            }
            A.curv[pS] = curv;
            A.nrm[pS] = make_float4(n.x, n.y, n.z, curv);
        }
    }
};

// On the asynchronous tile pipeline: the few warps of a tile that hold surface particles (and run the sample tests) no
// longer keep the rest of the CTA waiting at a tile barrier (57 % of the warp-stall samples of the barrier version,
// profiles/r01_ncu_full_step_before_pipeline.txt) — the other warps move on to the next tiles.
__global__ void __launch_bounds__(StClassifyOp::Cfg::THREADS, 1) k_st_classify(const __grid_constant__ Params P, const __grid_constant__ Arrays A, DevState* S, const float* __restrict__ halton) {
    const float radiusRatio = P.nbrRadius / P.r;
    StClassifyOp op{ P, A, halton, S->sampleCount, S->mcFactor, radiusRatio * radiusRatio * P.h2 };
    pipe_pass(S, A, pipe_header(smemRaw), pipe_pay<StClassifyOp>(smemRaw), op, P.tile0, P.tile1);
}

// T2: neighbour-weighted smoothing of normal and curvature among surface particles
struct StSmoothOp {
    static constexpr bool CUSTOM = false;
    using Cfg = PipeCfgMany;
    static constexpr int NPAY = 2, BBYTES = 16, NOWN = 4, NSUM = 5, COEF = 0, NRED = 0, NLUT = 0;       // payload: position, (normal, curvature)
    const Params& P; const Arrays& A;
    __device__ __forceinline__ const float4* srcA() const { return A.posRho; }
    __device__ __forceinline__ const void* srcB() const { return A.nrm; }
    __device__ __forceinline__ const float* coef_in() const { return nullptr; }
    __device__ __forceinline__ float* coef_out() const { return nullptr; }
    __device__ __forceinline__ void prefetch_own(uint32_t, uint32_t, uint32_t) const {}
    __device__ __forceinline__ float4 loadA(uint32_t g) const { return A.posRho[g]; }
    __device__ __forceinline__ float4 loadB(uint32_t g) const { return A.nrm[g]; }
    static constexpr bool PAD_SAFE = false;
    __device__ __forceinline__ void own_from(uint32_t, float4 x, float4 n, float (&own)[NOWN]) const {
        own[0] = x.x; own[1] = x.y; own[2] = x.z;
        own[3] = (n.x != 0.0f || n.y != 0.0f || n.z != 0.0f) ? 1.0f : 0.0f;
    }
    __device__ __forceinline__ void pair(const float (&o)[NOWN], float4 a, float4 nj, float&, float (&acc)[NSUM]) const {
        if (o[3] != 0.0f && (nj.x != 0.0f || nj.y != 0.0f || nj.z != 0.0f)) {
            const float3 d = f3(a) - f3(o[0], o[1], o[2]);
            const float dist = sqrtf(dot3(d, d));
            const float w = 1.0f - dist / P.h;
            acc[0] += nj.x * w; acc[1] += nj.y * w; acc[2] += nj.z * w;
            acc[3] += nj.w * w;
            acc[4] += w;
        }
    }
    __device__ __forceinline__ void finish(uint32_t p, uint32_t, const float (&o)[NOWN], const float (&sum)[NSUM]) const {
        if (o[3] == 0.0f) return;
        const float tau = P.smoothing;
        const float4 ni4 = A.nrm[p];
        const float3 nc = normalize_if_nonzero(f3(sum[0], sum[1], sum[2]));
        float3 ns = (1.0f - tau) * f3(ni4) + tau * nc;
        ns = normalize_if_nonzero(ns);
        A.nbar[p] = make_float4(ns.x, ns.y, ns.z, 0.0f);
        A.curvS[p] = ((1.0f - tau) * ni4.w + tau * sum[3]) / (1.0f - tau + tau * sum[4]);
    }
};

__global__ void __launch_bounds__(StSmoothOp::Cfg::THREADS, 1) k_st_smooth(const __grid_constant__ Params P, const __grid_constant__ Arrays A, DevState* S) {
    StSmoothOp op{ P, A };
    pipe_pass(S, A, pipe_header(smemRaw), pipe_pay<StSmoothOp>(smemRaw), op, P.tile0, P.tile1);
}

// T3: apply the force (once per smoothing pass: SURVEY.md Q18)
__global__ void __launch_bounds__(VFD_TPB) k_st_apply(Params P, Arrays A) {
    OWNED_INDEX(p);
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

void launch_st_classify(const LaunchCfg& L, const Params& P, const Arrays& A, DevState* S, const float* halton) {
    const size_t sp = pipe_smem_bytes<StClassifyOp>();
    static thread_local int dev = -1;
    if (launch_device_changed(dev)) cudaFuncSetAttribute(k_st_classify, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sp);
    LaunchScope ls(L, KID_ST_CLASSIFY);
    k_st_classify<<<L.numSMs, StClassifyOp::Cfg::THREADS, sp, L.stream>>>(P, A, S, halton);
}
void launch_st_smooth(const LaunchCfg& L, const Params& P, const Arrays& A, DevState* S) {
    const size_t sp = pipe_smem_bytes<StSmoothOp>();
    static thread_local int dev = -1;
    if (launch_device_changed(dev)) cudaFuncSetAttribute(k_st_smooth, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sp);
    LaunchScope ls(L, KID_ST_SMOOTH);
    k_st_smooth<<<L.numSMs, StSmoothOp::Cfg::THREADS, sp, L.stream>>>(P, A, S);
}
void launch_st_apply(const LaunchCfg& L, const Params& P, const Arrays& A) {
    LaunchScope ls(L, KID_ST_APPLY);
    k_st_apply<<<std::max(1u, (P.n + VFD_TPB - 1) / VFD_TPB), VFD_TPB, 0, L.stream>>>(P, A);
}

} // namespace vfd
