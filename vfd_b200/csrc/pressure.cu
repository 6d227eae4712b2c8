// pressure.cu — the DFSPH core: density, factor, the divergence-free and constant-density Jacobi
// solvers, time integration and the CFL update.
//
// Replaces (reference: DFSPHKernels.cu) ComputeDensityKernel :152, ComputeDFSPHFactorKernel :189,
// ClearAccelerationKernel :24, ComputeDensityChangeKernel :483, ComputePressureAccelerationAnd-
// DivergenceKernel :388, DivergenceSolveIterationKernel :531, ComputePressureAccelerationAndFactor-
// Kernel :555, ComputeDensityAdvectionKernel :233, ComputePressureAccelerationKernel :342,
// PressureSolveIterationKernel :317, ComputePressureAccelerationAndVelocityKernel :434,
// ComputeVelocityKernel :39, ComputePositionKernel :54, and the host logic around them
// (DFSPHImplementation.cu: ComputeDivergence :506-575, ComputePressure :443-504,
// ComputeTimeStepSize :395-428, ComputeMaxVelocityMagnitude :430-441).
//
// The neighbour-sum kernels are tile passes (tile.cuh): persistent CTAs, the 40-kB kernel lookup table
// staged into shared memory once per CTA, the neighbour payload of a tile's 6x6x6-cell halo box staged
// once per tile, groups of 8 lanes per particle streaming 16-bit local neighbour indices and reducing
// the kernel-weighted sums with warp shuffles.  Solver control
// (iteration counters, residual means, continue flags) lives in DevState: the loop-carried decision
// of the reference's host loops is taken by the last block of the iteration kernel.
#include "solver.h"
#include "tile.cuh"
#include "control.cuh"
#include <algorithm>

namespace vfd {

extern __shared__ __align__(128) unsigned char smemRaw[];

#define FOR_EACH_OWNED(p) uint32_t ownB_, ownE_; owned_range(P, A.cellBegin, ownB_, ownE_); \
    for (uint32_t p = ownB_ + blockIdx.x * blockDim.x + threadIdx.x; p < ownE_; p += gridDim.x * blockDim.x)
#define OWNED_INDEX(p) uint32_t ownB_, ownE_; owned_range(P, A.cellBegin, ownB_, ownE_); \
    const uint32_t p = ownB_ + blockIdx.x * blockDim.x + threadIdx.x; if (p >= ownE_) return
#define NO_B __device__ __forceinline__ float4 loadB(uint32_t) const { return make_float4(0.0f, 0.0f, 0.0f, 0.0f); }

// ---- K2 + K3 + K8 fused: density, DFSPH factor, a = g --------------------------------------
struct DensityFactorOp {
    static constexpr bool CUSTOM = false;
    static constexpr int NPAY = 1, NOWN = 3, NSUM = 5, COEF = 0;
    const Params& P; const Arrays& A; Lut K;
    __device__ __forceinline__ float4 loadA(uint32_t g) const { return A.pos[g]; }
    NO_B
    __device__ __forceinline__ void load_own(uint32_t p, float (&own)[NOWN]) const {
        const float4 x = A.pos[p]; own[0] = x.x; own[1] = x.y; own[2] = x.z;
    }
    __device__ __forceinline__ void pair(const float (&o)[NOWN], float4 a, float4, float&, float (&acc)[NSUM]) const {
        const float3 xij = f3(o[0], o[1], o[2]) - f3(a);
        acc[0] += P.volume * K.w(xij);
        const float3 gj = -P.volume * K.gradW(xij);
        acc[1] += dot3(gj, gj);
        acc[2] -= gj.x; acc[3] -= gj.y; acc[4] -= gj.z;
    }
    __device__ __forceinline__ void finish(uint32_t p, uint32_t, const float (&o)[NOWN], const float (&sum)[NSUM]) const {
        const float3 xi = f3(o[0], o[1], o[2]);
        float rho = P.volume * P.wZero + sum[0];
        float sumK = sum[1];
        float3 gradI = f3(sum[2], sum[3], sum[4]);
        for (uint32_t b = 0; b < P.nBodies; b++) {
            const float4 bx = A.bx[b][p];
            if (bx.w > 0.0f) {
                const float3 xib = xi - f3(bx);
                rho += bx.w * K.w(xib);
                const float3 gj = -bx.w * K.gradW(xib);
                gradI -= gj;
            }
        }
        rho *= P.rho0;
        sumK += dot3(gradI, gradI);
        A.rho[p] = rho;
        A.posRho[p] = make_float4(xi.x, xi.y, xi.z, rho);
        A.alpha[p] = (sumK > VFD_EPS_F) ? 1.0f / sumK : 0.0f;
        A.acc[p] = make_float4(P.gx, P.gy, P.gz, 0.0f);
    }
};

__global__ void __launch_bounds__(TT_LUT2) k_density_factor(const __grid_constant__ Params P, const __grid_constant__ Arrays A, DevState* S,
                                                                 const float* __restrict__ lutW, const float* __restrict__ lutG) {
    float* sW = smem_lut<2>(smemRaw);
    float* sG = sW + VFD_LUT_RES;
    load_lut_tile(sW, lutW);
    load_lut_tile(sG, lutG);
    DensityFactorOp op{ P, A, Lut{ sW, sG, P.lutInvStep, P.lutRadius, P.lutRadius2 } };
    tile_pass(S, A, smem_header(smemRaw), smem_pay_a<2>(smemRaw), nullptr, STAGE_CAP, op, P.tile0, P.tile1);
}

// ---- K4 / K10: solver source terms ----------------------------------------------------------
// rate = V * sum_j (v_i - v_j) . gradW_ij + sum_b V_b v_i . gradW_ib
template<bool DIV>
struct SourceOp {
    static constexpr bool CUSTOM = false;
    static constexpr int NPAY = 2, NOWN = 6, NSUM = 1, COEF = 0;
    const Params& P; const Arrays& A; Lut K;
    float dt, dtInv, dt2Inv;
    __device__ __forceinline__ float4 loadA(uint32_t g) const { return A.posRho[g]; }
    __device__ __forceinline__ float4 loadB(uint32_t g) const { return A.vel[g]; }
    __device__ __forceinline__ void load_own(uint32_t p, float (&own)[NOWN]) const {
        const float4 x = A.posRho[p], v = A.vel[p];
        own[0] = x.x; own[1] = x.y; own[2] = x.z; own[3] = v.x; own[4] = v.y; own[5] = v.z;
    }
    __device__ __forceinline__ void pair(const float (&o)[NOWN], float4 a, float4 b, float&, float (&acc)[NSUM]) const {
        acc[0] += dot3(f3(o[3], o[4], o[5]) - f3(b), K.gradW(f3(o[0], o[1], o[2]) - f3(a)));
    }
    __device__ __forceinline__ void finish(uint32_t p, uint32_t m, const float (&o)[NOWN], const float (&sum)[NSUM]) const {
        const float3 xi = f3(o[0], o[1], o[2]), vi = f3(o[3], o[4], o[5]);
        float s = sum[0] * P.volume;
        for (uint32_t b = 0; b < P.nBodies; b++) {
            const float4 bx = A.bx[b][p];
            if (bx.w > 0.0f) s += bx.w * dot3(vi, K.gradW(xi - f3(bx)));
        }
        if (DIV) {
            const float adv = m < 20u ? 0.0f : fmaxf(s, 0.0f);
            const float factor = A.alpha[p] * dtInv;
            A.rhoAdv[p] = adv;
            A.alpha[p] = factor;
            A.kappaV[p] = adv * factor;
        } else {
            const float adv = A.rho[p] / P.rho0 + dt * s;
            const float factor = A.alpha[p] * dt2Inv;
            const float residuum = fminf(1.0f - adv, 0.0f);
            A.rhoAdv[p] = adv;
            A.alpha[p] = factor;
            A.kappa[p] = -residuum * factor;
        }
    }
};

template<bool DIV>
__global__ void __launch_bounds__(TT_LUT, 2) k_source(const __grid_constant__ Params P, const __grid_constant__ Arrays A, DevState* S, const float* __restrict__ lutG) {
    float* sG = smem_lut<1>(smemRaw);
    load_lut_tile(sG, lutG);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        // loop entry of ComputeDivergence / ComputePressure (DFSPHImplementation.cu:526-532, 455-461): error 0, so the
        // loop is entered only through the minimum iteration count
        if (DIV) { S->divIt = 0; S->divErr = 0.0f; S->divActive = (0u < P.minDivIt && 0u < P.maxDivIt) ? 1u : 0u; }
        else     { S->pressIt = 0; S->pressErr = 0.0f; S->pressActive = (0u < P.minPressIt && 0u < P.maxPressIt) ? 1u : 0u; }
    }
    SourceOp<DIV> op{ P, A, Lut{ nullptr, sG, P.lutInvStep, P.lutRadius, P.lutRadius2 }, S->dt, S->dtInv, S->dt2Inv };
    tile_pass(S, A, smem_header(smemRaw), smem_pay_a<1>(smemRaw), smem_pay_b<1>(smemRaw, STAGE_CAP_LUT), STAGE_CAP_LUT, op, P.tile0, P.tile1);
}

// ---- K5 / K7 / K11 / K13: pressure acceleration from kappa ---------------------------------
enum { ACC_DIV_ITER = 0, ACC_DIV_FINISH = 1, ACC_PRESS_ITER = 2, ACC_PRESS_FINISH = 3 };

template<int MODE>
struct AccelOp {
    static constexpr bool CUSTOM = false;
    static constexpr int NPAY = 1, NOWN = 4, NSUM = 3, COEF = 0;       // payload (x, y, z, kappa)
    const Params& P; const Arrays& A; Lut K;
    const float* __restrict__ kap;
    float dt;
    __device__ __forceinline__ float4 loadA(uint32_t g) const { float4 x = A.posRho[g]; x.w = kap[g]; return x; }
    NO_B
    __device__ __forceinline__ void load_own(uint32_t p, float (&own)[NOWN]) const {
        const float4 x = A.posRho[p]; own[0] = x.x; own[1] = x.y; own[2] = x.z; own[3] = kap[p];
    }
    __device__ __forceinline__ void pair(const float (&o)[NOWN], float4 a, float4, float&, float (&acc)[NSUM]) const {
        const float ks = o[3] + a.w;
        if (fabsf(ks) > VFD_EPS_F) {
            const float3 gj = -P.volume * K.gradW(f3(o[0], o[1], o[2]) - f3(a));
            acc[0] += ks * gj.x; acc[1] += ks * gj.y; acc[2] += ks * gj.z;
        }
    }
    __device__ __forceinline__ void finish(uint32_t p, uint32_t, const float (&o)[NOWN], const float (&sum)[NSUM]) const {
        const float3 xi = f3(o[0], o[1], o[2]);
        const float ki = o[3];
        float3 a = f3(sum[0], sum[1], sum[2]);
        if (fabsf(ki) > VFD_EPS_F) {
            for (uint32_t b = 0; b < P.nBodies; b++) {
                const float4 bx = A.bx[b][p];
                if (bx.w > 0.0f) {
                    const float3 gj = -bx.w * K.gradW(xi - f3(bx));
                    a += ki * gj;
                }
            }
        }
        A.pacc[p] = make_float4(a.x, a.y, a.z, 0.0f);
        if (MODE == ACC_DIV_FINISH || MODE == ACC_PRESS_FINISH) {
            float4 v = A.vel[p];
            v.x += dt * a.x; v.y += dt * a.y; v.z += dt * a.z;
            A.vel[p] = v;
        }
        if (MODE == ACC_DIV_FINISH) A.alpha[p] *= dt;
    }
};

template<int MODE>
__global__ void __launch_bounds__(TT_LUT, 2) k_pressure_accel(const __grid_constant__ Params P, const __grid_constant__ Arrays A, DevState* S, const float* __restrict__ lutG) {
    if (MODE == ACC_DIV_ITER && !S->divActive) return;
    if (MODE == ACC_PRESS_ITER && !S->pressActive) return;
    float* sG = smem_lut<1>(smemRaw);
    load_lut_tile(sG, lutG);
    AccelOp<MODE> op{ P, A, Lut{ nullptr, sG, P.lutInvStep, P.lutRadius, P.lutRadius2 },
                      (MODE == ACC_DIV_ITER || MODE == ACC_DIV_FINISH) ? A.kappaV : A.kappa, S->dt };
    tile_pass(S, A, smem_header(smemRaw), smem_pay_a<1>(smemRaw), nullptr, STAGE_CAP, op, P.tile0, P.tile1);
}

// ---- K6 / K12 (+ R1 / R3): one Jacobi update and the fused residual reduction ----------------
template<bool DIV>
struct SolveOp {
    static constexpr bool CUSTOM = false;
    static constexpr int NPAY = 2, NOWN = 6, NSUM = 1, COEF = 0;       // payload: position, pressure acceleration
    const Params& P; const Arrays& A; Lut K;
    float scale;
    float errSum;
    __device__ __forceinline__ float4 loadA(uint32_t g) const { return A.posRho[g]; }
    __device__ __forceinline__ float4 loadB(uint32_t g) const { return A.pacc[g]; }
    __device__ __forceinline__ void load_own(uint32_t p, float (&own)[NOWN]) const {
        const float4 x = A.posRho[p], a = A.pacc[p];
        own[0] = x.x; own[1] = x.y; own[2] = x.z; own[3] = a.x; own[4] = a.y; own[5] = a.z;
    }
    __device__ __forceinline__ void pair(const float (&o)[NOWN], float4 a, float4 b, float&, float (&acc)[NSUM]) const {
        acc[0] += dot3(f3(o[3], o[4], o[5]) - f3(b), K.gradW(f3(o[0], o[1], o[2]) - f3(a)));
    }
    __device__ __forceinline__ void finish(uint32_t p, uint32_t m, const float (&o)[NOWN], const float (&sum)[NSUM]) {
        const float3 xi = f3(o[0], o[1], o[2]), ai = f3(o[3], o[4], o[5]);
        float s = sum[0] * P.volume;
        for (uint32_t b = 0; b < P.nBodies; b++) {
            const float4 bx = A.bx[b][p];
            if (bx.w > 0.0f) s += bx.w * dot3(ai, K.gradW(xi - f3(bx)));
        }
        s *= scale;
        float residuum;
        if (DIV) {
            residuum = m < 20u ? 0.0f : fminf(-A.rhoAdv[p] - s, 0.0f);
            A.kappaV[p] -= residuum * A.alpha[p];
        } else {
            residuum = fminf(1.0f - A.rhoAdv[p] - s, 0.0f);
            A.kappa[p] -= residuum * A.alpha[p];
        }
        A.res[p] = residuum;
        errSum += P.rho0 * residuum;
    }
};

template<bool DIV>
__global__ void __launch_bounds__(TT_LUT, 2) k_solve_iteration(const __grid_constant__ Params P, const __grid_constant__ Arrays A, DevState* S, const float* __restrict__ lutG) {
    if (DIV ? !S->divActive : !S->pressActive) return;
    TileShared& sh = smem_header(smemRaw);
    float* sG = smem_lut<1>(smemRaw);
    load_lut_tile(sG, lutG);
    SolveOp<DIV> op{ P, A, Lut{ nullptr, sG, P.lutInvStep, P.lutRadius, P.lutRadius2 }, DIV ? S->dt : S->dt2, 0.0f };
    tile_pass(S, A, sh, smem_pay_a<1>(smemRaw), smem_pay_b<1>(smemRaw, STAGE_CAP_LUT), STAGE_CAP_LUT, op, P.tile0, P.tile1);
    __syncthreads();
    double v[1] = { (double)op.errSum };
    uint32_t* ticket = &S->ticket[DIV ? 1 : 2];
    if (block_reduce_publish<1>(v, A.partials, ticket, sh.red)) {
        double tot[1];
        last_block_fold<1>(tot, A.partials, sh.red);
        if (threadIdx.x == 0) {
            finish_reduction<1>(DIV ? SITE_DIV : SITE_PRESS, P, S, tot);
            *ticket = 0;
        }
    }
}

// ---- R2 + ComputeTimeStepSize: CFL on the device ---------------------------------------------
__global__ void __launch_bounds__(VFD_TPB) k_cfl(Params P, Arrays A, DevState* S) {
    __shared__ double shRed[32];
    const float dt = S->dt;
    float mx = 0.0f;
    FOR_EACH_OWNED(p) {
        const float4 v = A.vel[p], a = A.acc[p];
        const float3 w = f3(v.x + a.x * dt, v.y + a.y * dt, v.z + a.z * dt);
        mx = fmaxf(mx, dot3(w, w));
    }
    double v[1] = { (double)mx };
    if (block_reduce_publish<1>(v, A.partials, &S->ticket[3], shRed, true)) {
        double tot[1];
        last_block_fold<1>(tot, A.partials, shRed, true);
        if (threadIdx.x == 0) {
            finish_reduction<1>(SITE_CFL, P, S, tot);
            S->ticket[3] = 0;
        }
    }
}

// K9: v += dt * a
__global__ void __launch_bounds__(VFD_TPB) k_velocity(Params P, Arrays A, DevState* S) {
    OWNED_INDEX(p);
    const float dt = S->dt;
    float4 v = A.vel[p];
    const float4 a = A.acc[p];
    v.x += dt * a.x; v.y += dt * a.y; v.z += dt * a.z;
    A.vel[p] = v;
}

// K14: x += dt * v
__global__ void __launch_bounds__(VFD_TPB) k_position(Params P, Arrays A, DevState* S) {
    OWNED_INDEX(p);
    const float dt = S->dt;
    float4 x = A.posRho[p];
    const float4 v = A.vel[p];
    x.x += dt * v.x; x.y += dt * v.y; x.z += dt * v.z;
    A.pos[p] = x;
}

__global__ void k_clear_acc(Params P, Arrays A) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < P.n) A.acc[p] = make_float4(P.gx, P.gy, P.gz, 0.0f);
}

// ---- launchers -------------------------------------------------------------------------------
// persistent grid: resident CTAs per SM (from the occupancy calculator) x SMs
template<typename Kern>
static uint32_t tile_grid(Kern kern, size_t smem, const LaunchCfg& L, int threads) {
    static thread_local const void* cachedK[32]; static thread_local int cachedV[32]; static thread_local int nc = 0;
    int perSM = 0;
    for (int i = 0; i < nc; i++) if (cachedK[i] == (const void*)kern) perSM = cachedV[i];
    if (!perSM) {
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, kern, threads, smem);
        if (perSM < 1) perSM = 1;
        if (nc < 32) { cachedK[nc] = (const void*)kern; cachedV[nc] = perSM; nc++; }
    }
    return (uint32_t)(perSM * L.numSMs);
}

void launch_density_factor(const LaunchCfg& L, const Params& P, const Arrays& A, DevState* S, const float* lutW, const float* lutG) {
    const size_t smem = tile_smem_bytes<2, 1>(STAGE_CAP);
    const uint32_t g = tile_grid(k_density_factor, smem, L, TT_LUT2);
    LaunchScope ls(L, KID_DENSITY_FACTOR);
    k_density_factor<<<g, TT_LUT2, smem, L.stream>>>(P, A, S, lutW, lutG);
}
void launch_divergence_source(const LaunchCfg& L, const Params& P, const Arrays& A, DevState* S, const float* lutG) {
    const size_t smem = tile_smem_bytes<1, 2>(STAGE_CAP_LUT);
    const uint32_t g = tile_grid(k_source<true>, smem, L, TT_LUT);
    LaunchScope ls(L, KID_DIV_SOURCE);
    k_source<true><<<g, TT_LUT, smem, L.stream>>>(P, A, S, lutG);
}
void launch_divergence_accel(const LaunchCfg& L, const Params& P, const Arrays& A, DevState* S, const float* lutG) {
    const size_t s1 = tile_smem_bytes<1, 1>(STAGE_CAP);
    const uint32_t g1 = tile_grid(k_pressure_accel<ACC_DIV_ITER>, s1, L, TT_LUT);
    LaunchScope ls(L, KID_DIV_ACCEL);
    k_pressure_accel<ACC_DIV_ITER><<<g1, TT_LUT, s1, L.stream>>>(P, A, S, lutG);
}
void launch_divergence_solve(const LaunchCfg& L, const Params& P, const Arrays& A, DevState* S, const float* lutG) {
    const size_t s2 = tile_smem_bytes<1, 2>(STAGE_CAP_LUT);
    const uint32_t g2 = tile_grid(k_solve_iteration<true>, s2, L, TT_LUT);
    LaunchScope ls(L, KID_DIV_SOLVE);
    k_solve_iteration<true><<<g2, TT_LUT, s2, L.stream>>>(P, A, S, lutG);
}
void launch_divergence_finish(const LaunchCfg& L, const Params& P, const Arrays& A, DevState* S, const float* lutG) {
    const size_t smem = tile_smem_bytes<1, 1>(STAGE_CAP);
    const uint32_t g = tile_grid(k_pressure_accel<ACC_DIV_FINISH>, smem, L, TT_LUT);
    LaunchScope ls(L, KID_DIV_FINISH);
    k_pressure_accel<ACC_DIV_FINISH><<<g, TT_LUT, smem, L.stream>>>(P, A, S, lutG);
}
void launch_pressure_source(const LaunchCfg& L, const Params& P, const Arrays& A, DevState* S, const float* lutG) {
    const size_t smem = tile_smem_bytes<1, 2>(STAGE_CAP_LUT);
    const uint32_t g = tile_grid(k_source<false>, smem, L, TT_LUT);
    LaunchScope ls(L, KID_PRESS_SOURCE);
    k_source<false><<<g, TT_LUT, smem, L.stream>>>(P, A, S, lutG);
}
void launch_pressure_accel(const LaunchCfg& L, const Params& P, const Arrays& A, DevState* S, const float* lutG) {
    const size_t s1 = tile_smem_bytes<1, 1>(STAGE_CAP);
    const uint32_t g1 = tile_grid(k_pressure_accel<ACC_PRESS_ITER>, s1, L, TT_LUT);
    LaunchScope ls(L, KID_PRESS_ACCEL);
    k_pressure_accel<ACC_PRESS_ITER><<<g1, TT_LUT, s1, L.stream>>>(P, A, S, lutG);
}
void launch_pressure_solve(const LaunchCfg& L, const Params& P, const Arrays& A, DevState* S, const float* lutG) {
    const size_t s2 = tile_smem_bytes<1, 2>(STAGE_CAP_LUT);
    const uint32_t g2 = tile_grid(k_solve_iteration<false>, s2, L, TT_LUT);
    LaunchScope ls(L, KID_PRESS_SOLVE);
    k_solve_iteration<false><<<g2, TT_LUT, s2, L.stream>>>(P, A, S, lutG);
}
void launch_pressure_finish(const LaunchCfg& L, const Params& P, const Arrays& A, DevState* S, const float* lutG) {
    const size_t smem = tile_smem_bytes<1, 1>(STAGE_CAP);
    const uint32_t g = tile_grid(k_pressure_accel<ACC_PRESS_FINISH>, smem, L, TT_LUT);
    LaunchScope ls(L, KID_PRESS_FINISH);
    k_pressure_accel<ACC_PRESS_FINISH><<<g, TT_LUT, smem, L.stream>>>(P, A, S, lutG);
}
void launch_clear_acceleration(const LaunchCfg& L, const Params& P, const Arrays& A) {
    LaunchScope ls(L, KID_CLEAR_ACC);
    k_clear_acc<<<std::max(1u, (P.n + VFD_TPB - 1) / VFD_TPB), VFD_TPB, 0, L.stream>>>(P, A);
}
void launch_cfl(const LaunchCfg& L, const Params& P, const Arrays& A, DevState* S) {
    const uint32_t tiles = std::max(1u, (P.n + VFD_TPB - 1) / VFD_TPB);
    LaunchScope ls(L, KID_CFL);
    k_cfl<<<std::max(1u, std::min<uint32_t>(tiles, (uint32_t)L.numSMs * 8u)), VFD_TPB, 0, L.stream>>>(P, A, S);
}
void launch_velocity(const LaunchCfg& L, const Params& P, const Arrays& A, DevState* S) {
    LaunchScope ls(L, KID_VELOCITY);
    k_velocity<<<std::max(1u, (P.n + VFD_TPB - 1) / VFD_TPB), VFD_TPB, 0, L.stream>>>(P, A, S);
}
void launch_positions(const LaunchCfg& L, const Params& P, const Arrays& A, DevState* S) {
    LaunchScope ls(L, KID_POSITION);
    k_position<<<std::max(1u, (P.n + VFD_TPB - 1) / VFD_TPB), VFD_TPB, 0, L.stream>>>(P, A, S);
}

} // namespace vfd
