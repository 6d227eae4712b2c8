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
// The neighbour-sum kernels are pipelined tile passes (tile.cuh: pipe_pass): one persistent CTA per SM, producer
// warps copying the neighbour payload of a tile's 6x6x6-cell halo box into a shared-memory ring with cp.async,
// consumer warps (one particle per lane) streaming 16-bit local neighbour indices and the per-pair kernel-gradient
// factor and gathering the payload from shared memory.  Solver control
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

// Positions and neighbour lists are frozen between two searches, so the kernel-gradient factor of a pair,
//   g_ij  with  gradW(x_ij) = g_ij x_ij      (the table lookup of PrecomputedDFSPHCubicKernel::GetGradientW,
//   Kernel/DFSPHKernels.h:67-80: one square root, one index conversion, two dependent table loads),
// is evaluated ONCE per step — by the density pass, which needs it anyway — and streamed next to the neighbour list
// (A.gcoef, same ELL layout, 4 B per pair) by the ~14 later neighbour passes of the step, bit-identical to evaluating
// it again.  The same holds for the boundary term: gradW(x_i - x_b) per particle and body is stored as A.bgrad
// (w = V_b; zero when the body is out of range).  Only the density pass keeps the lookup tables in shared memory.
#define NO_PREFETCH __device__ __forceinline__ void prefetch_own(uint32_t, uint32_t, uint32_t) const {}
#define COEF_IN_G __device__ __forceinline__ const float* coef_in() const { return A.gcoef; } \
                  __device__ __forceinline__ float* coef_out() const { return nullptr; }

// ---- K2 + K3 + K8 fused: density, DFSPH factor, a = g; writes the pair factors g_ij and the boundary gradients ----
struct DensityFactorOp {
    static constexpr bool CUSTOM = false;
    using Cfg = PipeCfgMany;
    static constexpr int NPAY = 1, BBYTES = 0, NOWN = 3, NSUM = 5, COEF = 2, NRED = 0, NLUT = 2;
    const Params& P; const Arrays& A; Lut K;
    __device__ __forceinline__ const float4* srcA() const { return A.pos; }
    __device__ __forceinline__ const void* srcB() const { return nullptr; }
    __device__ __forceinline__ const float* coef_in() const { return nullptr; }
    __device__ __forceinline__ float* coef_out() const { return A.gcoef; }
    NO_PREFETCH
    __device__ __forceinline__ float4 loadA(uint32_t g) const { return A.pos[g]; }
    NO_B
    static constexpr bool PAD_SAFE = false;
    __device__ __forceinline__ void own_from(uint32_t, float4 x, float4, float (&own)[NOWN]) const { own[0] = x.x; own[1] = x.y; own[2] = x.z; }
    __device__ __forceinline__ void pair(const float (&o)[NOWN], float4 a, float4, float& c, float (&acc)[NSUM]) const {
        const float3 xij = f3(o[0], o[1], o[2]) - f3(a);
        acc[0] += P.volume * K.w(xij);
        c = K.gradWScalar(xij);
        const float3 gj = -P.volume * (c * xij);
        acc[1] += dot3(gj, gj);
        acc[2] -= gj.x; acc[3] -= gj.y; acc[4] -= gj.z;
    }
    __device__ __forceinline__ void finish(uint32_t p, uint32_t mf, const float (&o)[NOWN], const float (&sum)[NSUM]) const {
        const float3 xi = f3(o[0], o[1], o[2]);
        float rho = P.volume * P.wZero + sum[0];
        float sumK = sum[1];
        float3 gradI = f3(sum[2], sum[3], sum[4]);
        if (mf & VFD_NEAR_BODY) for (uint32_t b = 0; b < P.nBodies; b++) {
            const float4 bx = A.bx[b][p];
            float4 bg = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            if (bx.w > 0.0f) {
                const float3 xib = xi - f3(bx);
                rho += bx.w * K.w(xib);
                const float3 gw = K.gradW(xib);
                const float3 gj = -bx.w * gw;
                gradI -= gj;
                bg = make_float4(gw.x, gw.y, gw.z, bx.w);
            }
            A.bgrad[b][p] = bg;
        }
        rho *= P.rho0;
        sumK += dot3(gradI, gradI);
        A.rho[p] = rho;
        A.posRho[p] = make_float4(xi.x, xi.y, xi.z, rho);
        A.alpha[p] = (sumK > VFD_EPS_F) ? 1.0f / sumK : 0.0f;
        A.acc[p] = make_float4(P.gx, P.gy, P.gz, 0.0f);
    }
};

__global__ void __launch_bounds__(DensityFactorOp::Cfg::THREADS, 1) k_density_factor(const __grid_constant__ Params P, const __grid_constant__ Arrays A, DevState* S,
                                                                    const float* __restrict__ lutW, const float* __restrict__ lutG) {
    float* sW = pipe_lut<DensityFactorOp>(smemRaw);
    float* sG = sW + VFD_LUT_RES;
    load_lut_tile(sW, lutW);
    load_lut_tile(sG, lutG);
    DensityFactorOp op{ P, A, Lut{ sW, sG, P.lutInvStep, P.lutRadius, P.lutRadius2 } };
    pipe_pass(S, A, pipe_header(smemRaw), pipe_pay<DensityFactorOp>(smemRaw), op, P.tile0, P.tile1);
}

// ---- K4 / K10: solver source terms ----------------------------------------------------------
// rate = V * sum_j (v_i - v_j) . gradW_ij + sum_b V_b v_i . gradW_ib
template<bool DIV>
struct SourceOp {
    static constexpr bool CUSTOM = false;
    using Cfg = PipeCfgWide;
    static constexpr int NPAY = 2, BBYTES = 16, NOWN = 6, NSUM = 1, COEF = 1, NRED = 0, NLUT = 0;
    const Params& P; const Arrays& A;
    float dt, dtInv, dt2Inv;
    __device__ __forceinline__ const float4* srcA() const { return A.posRho; }
    __device__ __forceinline__ const void* srcB() const { return A.vel; }
    COEF_IN_G
    NO_PREFETCH
    __device__ __forceinline__ float4 loadA(uint32_t g) const { return A.posRho[g]; }
    __device__ __forceinline__ float4 loadB(uint32_t g) const { return A.vel[g]; }
    static constexpr bool PAD_SAFE = true;                 // the pair term is linear in the kernel-gradient factor
    __device__ __forceinline__ void own_from(uint32_t, float4 x, float4 v, float (&own)[NOWN]) const {
        own[0] = x.x; own[1] = x.y; own[2] = x.z; own[3] = v.x; own[4] = v.y; own[5] = v.z;
    }
    __device__ __forceinline__ void pair(const float (&o)[NOWN], float4 a, float4 b, float& c, float (&acc)[NSUM]) const {
        acc[0] += dot3(f3(o[3], o[4], o[5]) - f3(b), c * (f3(o[0], o[1], o[2]) - f3(a)));
    }
    __device__ __forceinline__ void finish(uint32_t p, uint32_t mf, const float (&o)[NOWN], const float (&sum)[NSUM]) const {
        const uint32_t m = mf & VFD_COUNT_MASK;
        const float3 vi = f3(o[3], o[4], o[5]);
        float s = sum[0] * P.volume;
        if (mf & VFD_NEAR_BODY) for (uint32_t b = 0; b < P.nBodies; b++) {
            const float4 bg = A.bgrad[b][p];
            if (bg.w > 0.0f) s += bg.w * dot3(vi, f3(bg));
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
__global__ void __launch_bounds__(SourceOp<DIV>::Cfg::THREADS, 1) k_source(const __grid_constant__ Params P, const __grid_constant__ Arrays A, DevState* S) {
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        // loop entry of ComputeDivergence / ComputePressure (DFSPHImplementation.cu:526-532, 455-461): error 0, so the
        // loop is entered only through the minimum iteration count
        if (DIV) { S->divIt = 0; S->divErr = 0.0f; S->divActive = (0u < P.minDivIt && 0u < P.maxDivIt) ? 1u : 0u; }
        else     { S->pressIt = 0; S->pressErr = 0.0f; S->pressActive = (0u < P.minPressIt && 0u < P.maxPressIt) ? 1u : 0u; }
    }
    SourceOp<DIV> op{ P, A, S->dt, S->dtInv, S->dt2Inv };
    pipe_pass(S, A, pipe_header(smemRaw), pipe_pay<SourceOp<DIV>>(smemRaw), op, P.tile0, P.tile1);
}

// ---- K5 / K7 / K11 / K13: pressure acceleration from kappa ---------------------------------
enum { ACC_DIV_ITER = 0, ACC_DIV_FINISH = 1, ACC_PRESS_ITER = 2, ACC_PRESS_FINISH = 3 };

template<int MODE>
struct AccelOp {
    static constexpr bool CUSTOM = false;
    using Cfg = PipeCfgMany;
    static constexpr int NPAY = 2, BBYTES = 4, NOWN = 4, NSUM = 3, COEF = 1, NRED = 0, NLUT = 0;     // payload: (x, y, z, rho) and kappa
    const Params& P; const Arrays& A;
    const float* __restrict__ kap;
    float dt;
    __device__ __forceinline__ const float4* srcA() const { return A.posRho; }
    __device__ __forceinline__ const void* srcB() const { return kap; }
    COEF_IN_G
    NO_PREFETCH
    __device__ __forceinline__ float4 loadA(uint32_t g) const { return A.posRho[g]; }
    __device__ __forceinline__ float4 loadB(uint32_t g) const { return make_float4(kap[g], 0.0f, 0.0f, 0.0f); }
    static constexpr bool PAD_SAFE = true;
    __device__ __forceinline__ void own_from(uint32_t, float4 x, float4 k, float (&own)[NOWN]) const { own[0] = x.x; own[1] = x.y; own[2] = x.z; own[3] = k.x; }
    __device__ __forceinline__ void pair(const float (&o)[NOWN], float4 a, float4 b, float& c, float (&acc)[NSUM]) const {
        const float ks = o[3] + b.x;
        if (fabsf(ks) > VFD_EPS_F) {
            const float3 gj = -P.volume * (c * (f3(o[0], o[1], o[2]) - f3(a)));
            acc[0] += ks * gj.x; acc[1] += ks * gj.y; acc[2] += ks * gj.z;
        }
    }
    __device__ __forceinline__ void finish(uint32_t p, uint32_t mf, const float (&o)[NOWN], const float (&sum)[NSUM]) const {
        const float ki = o[3];
        float3 a = f3(sum[0], sum[1], sum[2]);
        if ((mf & VFD_NEAR_BODY) && fabsf(ki) > VFD_EPS_F) {
            for (uint32_t b = 0; b < P.nBodies; b++) {
                const float4 bg = A.bgrad[b][p];
                if (bg.w > 0.0f) {
                    const float3 gj = -bg.w * f3(bg);
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
__global__ void __launch_bounds__(AccelOp<MODE>::Cfg::THREADS, 1) k_pressure_accel(const __grid_constant__ Params P, const __grid_constant__ Arrays A, DevState* S) {
    if (MODE == ACC_DIV_ITER && !S->divActive) return;
    if (MODE == ACC_PRESS_ITER && !S->pressActive) return;
    AccelOp<MODE> op{ P, A, (MODE == ACC_DIV_ITER || MODE == ACC_DIV_FINISH) ? A.kappaV : A.kappa, S->dt };
    pipe_pass(S, A, pipe_header(smemRaw), pipe_pay<AccelOp<MODE>>(smemRaw), op, P.tile0, P.tile1);
}

// ---- K6 / K12 (+ R1 / R3): one Jacobi update and the fused residual reduction ----------------
template<bool DIV>
struct SolveOp {
    static constexpr bool CUSTOM = false;
    using Cfg = PipeCfgWide;
    static constexpr int NPAY = 2, BBYTES = 16, NOWN = 6, NSUM = 1, COEF = 1, NRED = 1, NLUT = 0;   // payload: position, pressure acceleration
    const Params& P; const Arrays& A;
    float scale;
    float red[1];                                          // rho0 * residuum of the lane's particle (summed grid-wide: tile.cuh RedRecord)
    __device__ __forceinline__ const float4* srcA() const { return A.posRho; }
    __device__ __forceinline__ const void* srcB() const { return A.pacc; }
    COEF_IN_G
    NO_PREFETCH
    __device__ __forceinline__ float4 loadA(uint32_t g) const { return A.posRho[g]; }
    __device__ __forceinline__ float4 loadB(uint32_t g) const { return A.pacc[g]; }
    static constexpr bool PAD_SAFE = true;
    __device__ __forceinline__ void own_from(uint32_t, float4 x, float4 a, float (&own)[NOWN]) const {
        own[0] = x.x; own[1] = x.y; own[2] = x.z; own[3] = a.x; own[4] = a.y; own[5] = a.z;
    }
    __device__ __forceinline__ void pair(const float (&o)[NOWN], float4 a, float4 b, float& c, float (&acc)[NSUM]) const {
        acc[0] += dot3(f3(o[3], o[4], o[5]) - f3(b), c * (f3(o[0], o[1], o[2]) - f3(a)));
    }
    __device__ __forceinline__ void finish(uint32_t p, uint32_t mf, const float (&o)[NOWN], const float (&sum)[NSUM]) {
        const uint32_t m = mf & VFD_COUNT_MASK;
        const float3 ai = f3(o[3], o[4], o[5]);
        float s = sum[0] * P.volume;
        if (mf & VFD_NEAR_BODY) for (uint32_t b = 0; b < P.nBodies; b++) {
            const float4 bg = A.bgrad[b][p];
            if (bg.w > 0.0f) s += bg.w * dot3(ai, f3(bg));
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
        red[0] = P.rho0 * residuum;
    }
};

template<bool DIV>
__global__ void __launch_bounds__(SolveOp<DIV>::Cfg::THREADS, 1) k_solve_iteration(const __grid_constant__ Params P, const __grid_constant__ Arrays A, DevState* S) {
    if (DIV ? !S->divActive : !S->pressActive) return;
    PipeShared& ps = pipe_header(smemRaw);
    SolveOp<DIV> op{ P, A, DIV ? S->dt : S->dt2, { 0.0f } };
    if (pipe_pass(S, A, ps, pipe_pay<SolveOp<DIV>>(smemRaw), op, P.tile0, P.tile1)) {
        double tot[1];
        fold_slots<1>(tot, A.slotSums, __ldg(A.tileList), A.slotStride, ps.red);
        if (threadIdx.x == 0) finish_reduction<1>(DIV ? SITE_DIV : SITE_PRESS, P, S, tot);
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
            finish_reduction<1>(SITE_CFL, P, S, tot, true);
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
// pipelined tile passes: one persistent CTA per SM
template<typename Kern>
static void pipe_attr(Kern kern, size_t smem) {
    static thread_local const void* done[32]; static thread_local int nd = 0;
    static thread_local int dev = -1;
    if (launch_device_changed(dev)) nd = 0;
    for (int i = 0; i < nd; i++) if (done[i] == (const void*)kern) return;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (nd < 32) done[nd++] = (const void*)kern;
}
#define PIPE_LAUNCH(kid, kern, OpT, smem, ...) do { using OPT_ = OpT; const size_t sm_ = (smem); pipe_attr(kern, sm_); LaunchScope ls(L, kid); \
    kern<<<L.numSMs, OpT::Cfg::THREADS, sm_, L.stream>>>(__VA_ARGS__); } while (0)

void launch_density_factor(const LaunchCfg& L, const Params& P, const Arrays& A, DevState* S, const float* lutW, const float* lutG) {
    PIPE_LAUNCH(KID_DENSITY_FACTOR, k_density_factor, DensityFactorOp, (pipe_smem_bytes<OPT_>()), P, A, S, lutW, lutG);
}
void launch_divergence_source(const LaunchCfg& L, const Params& P, const Arrays& A, DevState* S, const float*) {
    PIPE_LAUNCH(KID_DIV_SOURCE, k_source<true>, SourceOp<true>, (pipe_smem_bytes<OPT_>()), P, A, S);
}
void launch_divergence_accel(const LaunchCfg& L, const Params& P, const Arrays& A, DevState* S, const float*) {
    PIPE_LAUNCH(KID_DIV_ACCEL, k_pressure_accel<ACC_DIV_ITER>, AccelOp<0>, (pipe_smem_bytes<OPT_>()), P, A, S);
}
void launch_divergence_solve(const LaunchCfg& L, const Params& P, const Arrays& A, DevState* S, const float*) {
    PIPE_LAUNCH(KID_DIV_SOLVE, k_solve_iteration<true>, SolveOp<true>, (pipe_smem_bytes<OPT_>()), P, A, S);
}
void launch_divergence_finish(const LaunchCfg& L, const Params& P, const Arrays& A, DevState* S, const float*) {
    PIPE_LAUNCH(KID_DIV_FINISH, k_pressure_accel<ACC_DIV_FINISH>, AccelOp<0>, (pipe_smem_bytes<OPT_>()), P, A, S);
}
void launch_pressure_source(const LaunchCfg& L, const Params& P, const Arrays& A, DevState* S, const float*) {
    PIPE_LAUNCH(KID_PRESS_SOURCE, k_source<false>, SourceOp<true>, (pipe_smem_bytes<OPT_>()), P, A, S);
}
void launch_pressure_accel(const LaunchCfg& L, const Params& P, const Arrays& A, DevState* S, const float*) {
    PIPE_LAUNCH(KID_PRESS_ACCEL, k_pressure_accel<ACC_PRESS_ITER>, AccelOp<0>, (pipe_smem_bytes<OPT_>()), P, A, S);
}
void launch_pressure_solve(const LaunchCfg& L, const Params& P, const Arrays& A, DevState* S, const float*) {
    PIPE_LAUNCH(KID_PRESS_SOLVE, k_solve_iteration<false>, SolveOp<true>, (pipe_smem_bytes<OPT_>()), P, A, S);
}
void launch_pressure_finish(const LaunchCfg& L, const Params& P, const Arrays& A, DevState* S, const float*) {
    PIPE_LAUNCH(KID_PRESS_FINISH, k_pressure_accel<ACC_PRESS_FINISH>, AccelOp<0>, (pipe_smem_bytes<OPT_>()), P, A, S);
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
