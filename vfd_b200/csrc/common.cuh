// common.cuh — shared device/host definitions of the B200-native DFSPH solver step.
//
// Data layout (all arrays live in HBM, particle-major, in the *sorted* order produced by
// search.cu; `id[p]` is the persistent original index of the particle stored at slot p):
//   float4 arrays : gathered per neighbour with one 16-B load (pos, vel, pressure acceleration,
//                   PCG direction ...);  .w is a per-array payload or unused.
//   float  arrays : per-particle scalars, streamed coalesced.
//   neighbour list: "warp-blocked ELL" of 16-bit TILE-LOCAL indices in groups of four (tile.cuh): neighbours
//                   4g..4g+3 of the particle in lane l of 32-group w are the 8-byte word
//                   list16[(w*18 + g)*32 + l]  — a warp reads one 256-B block per four neighbours,
//                   only groups below the largest count in the warp are ever touched.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include <float.h>

#define VFD_MAX_NEIGHBORS 70          // reference: ParticleSearchKernels.cuh:7 (CUDA_MAX_NEIGHBORS)
#define VFD_LUT_RES 10000             // reference: Kernel/DFSPHKernels.h:142 (DFSPHKernel<10000u>)
#define VFD_LUT_N (VFD_LUT_RES - 1)   // combined (midpoint) entries: pos in [0, RES-2]
#define VFD_HALTON_N 49152u           // reference: DFSPHKernels.cu:920 (haltonVec323 size)
// reference: Core/Math/Math.h:6 — EPS is the *double* literal 1.0e-5, so `floatExpr > EPS` compares in double.
// float(1e-5) = 9.99999974737875e-06 lies just below the double value and no float lies between the two,
// hence for any float x:  (double)x > 1.0e-5  <=>  x > 1.0e-5f.  The fp32 comparison is exact, not approximate.
#define VFD_EPS_F 1.0e-5f
#define VFD_TPB 256                   // reference launch shape: DFSPHKernels.cuh:11
#define VFD_MAX_BODIES 8
// cnt[p] = neighbour count | VFD_NEAR_BODY when at least one rigid body has a boundary sample for the particle (set by
// the boundary kernel): the epilogues of the neighbour passes read the per-body arrays only then
#define VFD_NEAR_BODY 0x80000000u
#define VFD_COUNT_MASK 0x7fffffffu

namespace vfd {

// Everything constant between two SetDescription() calls; passed to kernels by value.
struct Params {
    uint32_t n;                 // particles held by this rank (owned + ghost copies of the neighbour slabs' edge tiles)
    uint32_t nGlobal;           // particles of the whole simulation
    uint32_t nBodies;
    uint32_t nRanks, rank;      // spatial decomposition in slabs of tile columns along x (distributed.cu); 1, 0 on one GPU
    uint32_t tile0, tile1;      // owned tiles [tile0, tile1) of this rank's local grid; tile1 = 0xffffffff: all tiles
    float h, h2, r, d;          // SupportRadius, SupportRadius2, ParticleRadius, ParticleDiameter
    float volume, rho0, mass, massInv;
    float mu, muB, tangentialDistance;
    float sigma, clsSlope, clsConst, smoothing, nbrRadius, mcFactor;
    int   temporalSmoothing;
    float gx, gy, gz;
    float lutInvStep, lutRadius, lutRadius2, wZero;
    float minDt, maxDt;
    int   csdFix, csd;
    float frameLength;
    float etaPressure;          // MaxPressureSolverError * 0.0001 * rho0   (DFSPHImplementation.cu:452)
    float divErrScale;          // MaxDivergenceSolverError * 0.0001 * rho0 (eta = dtInv * this, :523)
    float viscErr2;             // MaxViscositySolverError^2 * 0.0001       (:663)
    uint32_t minPressIt, maxPressIt, minDivIt, maxDivIt, minViscIt, maxViscIt;
    int   searchFma;
    int   tune[8];              // experiment knobs (env VFD_TUNE0..7; 0 = default behaviour)
    // several ranks: every rank's control block (control.cuh: PeerCtl) mapped into this process over NVLink (CUDA IPC), own one
    // included; null when the ranks talk through NCCL only
    void* peerCtl[8];
    // ... and where the PCG direction of this rank's first / last owned tile column lives in the left / right neighbour's ghost
    // range ([neighbour][(x, y, z, p.x) | (p.y, p.z)]), with the particle ranges of those columns: ownB, edgeLEnd, edgeRBegin, ownE.
    // The kernel that rewrites the direction stores these particles' words there as well (viscosity.cu: k_visc_step).
    unsigned char* haloP[2][2];
    uint32_t haloRange[4];
};

// Device-resident mutable scalars: the time step and all solver control state.  Kernels read dt
// from here (as the reference's kernels read d_Info), so a step needs no host round trip.
struct DevState {
    float dt, dt2, dtInv, dt2Inv;
    uint32_t sampleCount;       // SurfaceTensionSampleCount (0 until the first CFL update — SURVEY Q9)
    float mcFactor;             // MonteCarloFactor          (0 until the first CFL update)
    float vmax2;                // MaxVelocityMagnitude (squared, as the reference stores it)
    float frameTime;
    uint32_t frameIndex;
    uint32_t captureFlag;
    uint32_t stepCount;
    // search grid (derived on device each step)
    int32_t  gridMinCell[3];
    uint32_t gridDim[3];        // cells per axis, a multiple of 4
    uint32_t tileDim[3];        // tiles (4x4x4 cells) per axis
    uint32_t nTiles;
    uint32_t nCells;            // nTiles * 64
    float    gridOrigin[3];
    int32_t  boundsMin[3], boundsMax[3];   // cell = floor(x/h) extrema (ParticleSearchKernels.cu:40-62)
    uint32_t errorFlags;        // bit0: grid larger than capacity
    // Jacobi solvers
    uint32_t divIt, pressIt; uint32_t divActive, pressActive;
    float divErr, pressErr;
    // PCG
    uint32_t viscIt, viscActive;
    float viscErr;
    float rhsNorm2, threshold, resNorm2, delta, alpha, beta;
    // last-block tickets (one per reduction site)
    uint32_t ticket[8];
    // multi-GPU: this rank's partial reduction results per site (control.cuh), all-reduced in place
    double red[16];
    // multi-GPU: global cell coordinates of the local grid's cell (0,0,0)
    int32_t cellOffset[3];
    uint32_t fallbackTiles;     // tile passes whose halo box exceeded the shared-memory stage (slow, exact path); cumulative
    // dynamic tile queue of the pipelined passes (tile.cuh): next entry of the tile list, CTAs that have finished the
    // running pass (the last one resets both)
    uint32_t tileCursor, doneCtas;
    uint32_t stepGen;           // generation counter of the grid barrier of the fused PCG vector kernel (viscosity.cu: k_visc_step)
};

struct float3x3 { float m[9]; };   // column-major like glm::mat3x3: m[3*c + r]

#ifdef __CUDACC__
// ---- small vector helpers -------------------------------------------------------------------
__device__ __forceinline__ float3 f3(float x, float y, float z) { return make_float3(x, y, z); }
__device__ __forceinline__ float3 f3(const float4& v) { return make_float3(v.x, v.y, v.z); }
__device__ __forceinline__ float3 operator+(float3 a, float3 b) { return f3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ float3 operator-(float3 a, float3 b) { return f3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ float3 operator-(float3 a) { return f3(-a.x, -a.y, -a.z); }
__device__ __forceinline__ float3 operator*(float s, float3 a) { return f3(s * a.x, s * a.y, s * a.z); }
__device__ __forceinline__ float3 operator*(float3 a, float s) { return f3(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ float3 operator/(float3 a, float s) { return f3(a.x / s, a.y / s, a.z / s); }
__device__ __forceinline__ void operator+=(float3& a, float3 b) { a.x += b.x; a.y += b.y; a.z += b.z; }
__device__ __forceinline__ void operator-=(float3& a, float3 b) { a.x -= b.x; a.y -= b.y; a.z -= b.z; }
// glm::dot for vec3 is (x*x' + y*y') + z*z' (glm/detail/func_geometric.inl compute_dot<vec<3>>)
__device__ __forceinline__ float dot3(float3 a, float3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ float3 cross3(float3 x, float3 y) {
    return f3(x.y * y.z - y.y * x.z, x.z * y.x - y.z * x.x, x.x * y.y - y.x * x.y);
}
// glm::normalize = v * (1 / sqrt(dot(v,v)))
__device__ __forceinline__ float3 normalize3(float3 v) { return v * (1.0f / sqrtf(dot3(v, v))); }

// ---- cubic-spline lookup tables in shared memory ---------------------------------------------
// The reference tabulates W and (dW/dr)/r at 10 000 points and returns the mean of two adjacent
// entries (Kernel/DFSPHKernels.h:54-80).  0.5f*(T[p]+T[p+1]) is a pure function of p, so the
// host precombines it (bit-identical fp32 operations) into one 9 999-entry table per function and
// a lookup is a single shared-memory load.
struct Lut {
    const float* W;   // combined W table   (shared memory)
    const float* G;   // combined grad table (shared memory)
    float invStep, radius, radius2;
    __device__ __forceinline__ uint32_t index(float rl) const {
        return min(static_cast<uint32_t>(rl * invStep), static_cast<uint32_t>(VFD_LUT_RES - 2));
    }
    // GetW(vec3): r2 <= Radius2 test, then sqrt (DFSPHKernels.h:54-65)
    __device__ __forceinline__ float w(float3 r) const {
        const float r2 = dot3(r, r);
        float res = 0.0f;
        if (r2 <= radius2) res = W[index(sqrtf(r2))];
        return res;
    }
    // GetGradientW(vec3): rl <= Radius test on the length (DFSPHKernels.h:67-80)
    __device__ __forceinline__ float3 gradW(float3 r) const {
        const float rl = sqrtf(dot3(r, r));
        if (rl <= radius) return G[index(rl)] * r;
        return f3(0.0f, 0.0f, 0.0f);
    }
    __device__ __forceinline__ float gradWScalar(float3 r) const {   // g such that gradW = g * r
        const float rl = sqrtf(dot3(r, r));
        return rl <= radius ? G[index(rl)] : 0.0f;
    }
};

__device__ __forceinline__ void load_lut(float* dst, const float* __restrict__ src) {
    // 9 999 floats; vectorised cooperative copy global -> shared
    const float4* s4 = reinterpret_cast<const float4*>(src);
    float4* d4 = reinterpret_cast<float4*>(dst);
    for (int i = threadIdx.x; i < (VFD_LUT_RES / 4); i += blockDim.x) d4[i] = __ldg(s4 + i);
}

// ---- opting kernels into > 48 KB of dynamic shared memory -----------------------------------------
// cudaFuncSetAttribute acts on the CURRENT device, so "already done" is a fact about (kernel, device).  The launchers cache it per
// host thread and forget it when that thread's current device is another one than at their last call (a process that drives
// handles on two GPUs); with one device per thread — the only pattern this repo's tests and bench use — nothing changes.
static inline bool launch_device_changed(int& lastDevice) {
    int dev = -1;
    if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); dev = 0; }      // (then: behave as with one device)
    if (dev == lastDevice) return false;
    lastDevice = dev;
    return true;
}

// ---- deterministic grid-wide reductions ------------------------------------------------------
// Each block reduces to one double per quantity, stores it in `partials`, and the last block to
// finish (ticket counter) folds the partials in a fixed order: bit-reproducible run to run,
// unlike atomics or the reference's thrust::transform_reduce.
template<int NV>
__device__ __forceinline__ bool block_reduce_publish(double (&v)[NV], double* __restrict__ partials, uint32_t* ticket,
                                                     double* sh /* >= NV*32 doubles */, bool isMax = false) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    #pragma unroll
    for (int q = 0; q < NV; q++) {
        double x = v[q];
        #pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double y = __shfl_down_sync(0xffffffffu, x, o);
            x = isMax ? fmax(x, y) : x + y;
        }
        if (lane == 0) sh[q * 32 + warp] = x;
    }
    __syncthreads();
    __shared__ bool amLast;
    if (warp == 0) {
        #pragma unroll
        for (int q = 0; q < NV; q++) {
            double x = lane < nwarps ? sh[q * 32 + lane] : (isMax ? -DBL_MAX : 0.0);
            #pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const double y = __shfl_down_sync(0xffffffffu, x, o);
                x = isMax ? fmax(x, y) : x + y;
            }
            if (lane == 0) partials[(size_t)q * gridDim.x + blockIdx.x] = x;
        }
        if (lane == 0) {
            __threadfence();
            const uint32_t t = atomicAdd(ticket, 1u);
            amLast = (t == gridDim.x - 1);
        }
    }
    __syncthreads();
    return amLast;
}

// Called by every thread of the last block: folds partials[q*gridDim.x + b] over b in a fixed order.
template<int NV>
__device__ __forceinline__ void last_block_fold(double (&out)[NV], const double* __restrict__ partials, double* sh, bool isMax = false) {
    __threadfence();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    #pragma unroll
    for (int q = 0; q < NV; q++) {
        double x = isMax ? -DBL_MAX : 0.0;
        for (uint32_t b = threadIdx.x; b < gridDim.x; b += blockDim.x) {
            const double y = __ldcg(partials + (size_t)q * gridDim.x + b);
            x = isMax ? fmax(x, y) : x + y;
        }
        #pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double y = __shfl_down_sync(0xffffffffu, x, o);
            x = isMax ? fmax(x, y) : x + y;
        }
        __syncthreads();
        if (lane == 0) sh[q * 32 + warp] = x;
    }
    __syncthreads();
    #pragma unroll
    for (int q = 0; q < NV; q++) {
        double x = isMax ? -DBL_MAX : 0.0;
        for (int w = 0; w < nwarps; w++) { const double y = sh[q * 32 + w]; x = isMax ? fmax(x, y) : x + y; }
        out[q] = x;
    }
}

// Reductions of the dynamically scheduled tile passes: which CTA processes a tile is decided at run time, so partial
// sums are kept per TILE, in the slot of the tile's entry in the tile list (tile.cuh: RedRecord); the last CTA to
// finish folds the slots in list order — bit-reproducible run to run whatever the schedule was.  Result valid in every thread.
template<int NV>
__device__ __forceinline__ void fold_slots(double (&out)[NV], const double* __restrict__ slots, uint32_t nSlots, size_t stride, double* sh) {
    __threadfence();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    #pragma unroll
    for (int q = 0; q < NV; q++) {
        double x0 = 0.0, x1 = 0.0, x2 = 0.0, x3 = 0.0;
        uint32_t b = threadIdx.x;
        for (; b + 3u * blockDim.x < nSlots; b += 4u * blockDim.x) {
            const double* p = slots + (size_t)q * stride + b;
            const double y0 = __ldcg(p), y1 = __ldcg(p + blockDim.x), y2 = __ldcg(p + 2u * blockDim.x), y3 = __ldcg(p + 3u * blockDim.x);
            x0 += y0; x1 += y1; x2 += y2; x3 += y3;
        }
        for (; b < nSlots; b += blockDim.x) x0 += __ldcg(slots + (size_t)q * stride + b);
        double x = (x0 + x1) + (x2 + x3);
        #pragma unroll
        for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
        __syncthreads();
        if (lane == 0) sh[q * 32 + warp] = x;
    }
    __syncthreads();
    #pragma unroll
    for (int q = 0; q < NV; q++) {
        double x = 0.0;
        for (int w = 0; w < nwarps; w++) x += sh[q * 32 + w];
        out[q] = x;
    }
}

#endif // __CUDACC__

} // namespace vfd
