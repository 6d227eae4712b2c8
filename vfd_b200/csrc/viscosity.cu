// viscosity.cu — implicit viscosity (Weiler et al. 2018): matrix-free preconditioned CG with a 3x3
// block-Jacobi preconditioner, warm-started from the previous velocity change.
//
// Replaces ComputeViscosityPreconditionerKernel (reference: DFSPHKernels.cu:604-684),
// ComputeViscosityGradientKernel :686-756, ComputeMatrixVecProdFunctionKernel :758-842,
// ApplyViscosityForceKernel :844-862 and the host PCG DFSPHImplementation::SolveViscosity
// (DFSPHImplementation.cu:617-808: one mat-vec kernel + 11 thrust launches + 3 host-synchronous
// scalar reductions per iteration).  Here one PCG iteration is three kernels and no host round trip:
//   k_visc_matvec  q = A p            (+ p.q reduction, alpha = delta / p.q by the last block)
//   k_visc_update  g += alpha p; r -= alpha q; z = M^-1 r   (+ |r|^2, r.z; break test, beta)
//   k_visc_direction  p = z + beta p
// The recurrence, the break test and the iteration accounting are the reference's.
//
// Positions, densities and neighbour lists are frozen during the solve, so the per-pair scalar
//   c_ij = 10 mu (m / rho_j) / (|x_ij|^2 + 0.01 h^2) * (dW/dr)/r
// is computed once per step by the setup pass (which needs it for the preconditioner anyway) and
// stored next to the neighbour list (same ELL layout, 4 B/pair); the ~45 mat-vecs of a step then
// need no square root, no division and no table lookup per pair:  (A p)_i = p_i - dt/rho_i sum_j
// c_ij ((p_i - p_j) . x_ij) x_ij, with x_j and p_j gathered from the tile's shared-memory stage.
#include "solver.h"
#include "tile.cuh"
#include "control.cuh"
#include <algorithm>
#include <stdlib.h>
#include <stdio.h>

namespace vfd {

extern __shared__ __align__(128) unsigned char smemRaw[];
#define FOR_EACH_OWNED(p) uint32_t ownB_, ownE_; owned_range(P, A.cellBegin, ownB_, ownE_); \
    for (uint32_t p = ownB_ + blockIdx.x * blockDim.x + threadIdx.x; p < ownE_; p += gridDim.x * blockDim.x)
#define OWNED_INDEX(p) uint32_t ownB_, ownE_; owned_range(P, A.cellBegin, ownB_, ownE_); \
    const uint32_t p = ownB_ + blockIdx.x * blockDim.x + threadIdx.x; if (p >= ownE_) return

// Core/Math/Math.h:19-33 (GetOrthogonalVectors)
__device__ __forceinline__ void orthogonal_vectors(float3 n, float3& t1, float3& t2) {
    float3 v = f3(1.0f, 0.0f, 0.0f);
    if (fabsf(dot3(v, n)) >= 0.999f) v = f3(0.0f, 1.0f, 0.0f);   // "> 0.999" against a double literal: see EPS note in common.cuh
    t1 = cross3(n, v);
    t2 = cross3(n, t1);
    t1 = normalize3(t1);
    t2 = normalize3(t2);
}

// The four tangentially displaced boundary samples of Weiler's friction model
// (DFSPHKernels.cu:650-668 / :809-827): x_i - (x_b -/+ delta t1), x_i - (x_b -/+ delta t2)
__device__ __forceinline__ bool boundary_samples(const Params& P, float3 xi, float3 xb, float3 (&pd)[4]) {
    const float3 dir = xi - xb;
    float3 n = -dir;
    const float nl = sqrtf(dot3(n, n));
    if (!(nl > 0.0001f)) return false;
    n = n / nl;
    float3 t1, t2;
    orthogonal_vectors(n, t1, t2);
    pd[0] = xi - (xb - t1 * P.tangentialDistance);
    pd[1] = xi - (xb + t1 * P.tangentialDistance);
    pd[2] = xi - (xb - t2 * P.tangentialDistance);
    pd[3] = xi - (xb + t2 * P.tangentialDistance);
    return true;
}

// ---- V1 + V2 fused: preconditioner blocks, per-pair coefficients, warm start, |b|^2 ----------
struct ViscSetupOp {
    static constexpr bool CUSTOM = false;
    using Cfg = PipeCfgMany;
    static constexpr int NPAY = 1, BBYTES = 0, NOWN = 3, NSUM = 9, COEF = 3, NRED = 1, NLUT = 1; // payload (x, y, z, rho); reads the pair's kernel-gradient
    const Params& P; const Arrays& A; Lut K;                                   // factor g_ij (pressure.cu), writes the pair coefficients
    float dt, eps2;
    float red[1];                                          // |b_i|^2 of the lane's particle
    __device__ __forceinline__ const float4* srcA() const { return A.posRho; }
    __device__ __forceinline__ const void* srcB() const { return nullptr; }
    __device__ __forceinline__ const float* coef_in() const { return A.gcoef; }
    __device__ __forceinline__ float* coef_out() const { return A.coef; }
    __device__ __forceinline__ void prefetch_own(uint32_t, uint32_t, uint32_t) const {}
    __device__ __forceinline__ float4 loadA(uint32_t g) const { return A.posRho[g]; }
    __device__ __forceinline__ float4 loadB(uint32_t) const { return make_float4(0.0f, 0.0f, 0.0f, 0.0f); }
    static constexpr bool PAD_SAFE = false;
    __device__ __forceinline__ void own_from(uint32_t, float4 x, float4, float (&own)[NOWN]) const { own[0] = x.x; own[1] = x.y; own[2] = x.z; }
    __device__ __forceinline__ void pair(const float (&o)[NOWN], float4 xj, float4, float& coef, float (&M)[NSUM]) const {
        const float3 d = f3(o[0], o[1], o[2]) - f3(xj);
        const float g = coef;                             // gradW(d) = g d
        const float3 gw = g * d;
        const float s = 10.0f * P.mu * (P.mass / xj.w) / (dot3(d, d) + eps2);
        coef = s * g;
        // glm::outerProduct(c, r): column i = c * r[i]   (zero-initialised accumulator: SURVEY.md F6/Q3)
        M[0] += s * (d.x * gw.x); M[1] += s * (d.y * gw.x); M[2] += s * (d.z * gw.x);
        M[3] += s * (d.x * gw.y); M[4] += s * (d.y * gw.y); M[5] += s * (d.z * gw.y);
        M[6] += s * (d.x * gw.z); M[7] += s * (d.y * gw.z); M[8] += s * (d.z * gw.z);
    }
    __device__ __forceinline__ void finish(uint32_t p, uint32_t mf, const float (&o)[NOWN], const float (&sum)[NSUM]) {
        const float3 xi = f3(o[0], o[1], o[2]);
        const float rhoI = A.posRho[p].w;
        float M[9];
        #pragma unroll
        for (int q = 0; q < 9; q++) M[q] = sum[q];
        if (P.muB != 0.0f && (mf & VFD_NEAR_BODY)) {
            for (uint32_t b = 0; b < P.nBodies; b++) {
                const float4 bx = A.bx[b][p];
                float4 cb = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                if (bx.w > 0.0f) {
                    float3 pd[4];
                    if (boundary_samples(P, xi, f3(bx), pd)) {
                        const float vol = 0.25f * bx.w;
                        float cq[4];
                        #pragma unroll
                        for (int q = 0; q < 4; q++) {
                            const float g = K.gradWScalar(pd[q]);
                            const float3 gw = g * pd[q];
                            const float s = 10.0f * P.muB * vol / (dot3(pd[q], pd[q]) + eps2);
                            cq[q] = s * g;
                            const float3 c = pd[0];    // the reference uses positionDirection1 in all four products (DFSPHKernels.cu:674-677)
                            M[0] += s * (c.x * gw.x); M[1] += s * (c.y * gw.x); M[2] += s * (c.z * gw.x);
                            M[3] += s * (c.x * gw.y); M[4] += s * (c.y * gw.y); M[5] += s * (c.z * gw.y);
                            M[6] += s * (c.x * gw.z); M[7] += s * (c.y * gw.z); M[8] += s * (c.z * gw.z);
                        }
                        cb = make_float4(cq[0], cq[1], cq[2], cq[3]);
                    }
                }
                A.bcoef[b][p] = cb;
            }
        }
        // inverse(I - dt/rho_i * M), cofactor formula of glm::inverse(mat3) (glm/detail/func_matrix.inl:322-344)
        const float f = dt / rhoI;
        float a[9];
        #pragma unroll
        for (int q = 0; q < 9; q++) a[q] = ((q == 0 || q == 4 || q == 8) ? 1.0f : 0.0f) - f * M[q];
        // a[3*c + r] = m[c][r]
        const float c00 = a[4] * a[8] - a[7] * a[5];
        const float c01 = a[1] * a[8] - a[7] * a[2];
        const float c02 = a[1] * a[5] - a[4] * a[2];
        const float invDet = 1.0f / (a[0] * c00 - a[3] * c01 + a[6] * c02);
        float inv[9];
        inv[0] = +(a[4] * a[8] - a[7] * a[5]) * invDet;   // [0][0]
        inv[3] = -(a[3] * a[8] - a[6] * a[5]) * invDet;   // [1][0]
        inv[6] = +(a[3] * a[7] - a[6] * a[4]) * invDet;   // [2][0]
        inv[1] = -(a[1] * a[8] - a[7] * a[2]) * invDet;   // [0][1]
        inv[4] = +(a[0] * a[8] - a[6] * a[2]) * invDet;   // [1][1]
        inv[7] = -(a[0] * a[7] - a[6] * a[1]) * invDet;   // [2][1]
        inv[2] = +(a[1] * a[5] - a[4] * a[2]) * invDet;   // [0][2]
        inv[5] = -(a[0] * a[5] - a[3] * a[2]) * invDet;   // [1][2]
        inv[8] = +(a[0] * a[4] - a[3] * a[1]) * invDet;   // [2][2]
        #pragma unroll
        for (int q = 0; q < 9; q++) A.minv[(size_t)q * P.n + p] = inv[q];
        // V2: b = v (the boundary term multiplies a zero vector: DFSPHKernels.cu:744-754), g = v + dv_prev
        const float4 v = A.vel[p], dv = A.dv[p];
        const float3 g0 = f3(v.x + dv.x, v.y + dv.y, v.z + dv.z);
        A.cgG[p] = make_float4(g0.x, g0.y, g0.z, 0.0f);
        A.cgXG[p] = make_float4(xi.x, xi.y, xi.z, g0.x);     // what the first mat-vec gathers per neighbour (see ViscMatvecOp)
        A.cgGyz[p] = make_float2(g0.y, g0.z);
        red[0] = (v.x * v.x + v.y * v.y) + v.z * v.z;
    }
};

__global__ void __launch_bounds__(ViscSetupOp::Cfg::THREADS, 1) k_visc_setup(const __grid_constant__ Params P, const __grid_constant__ Arrays A, DevState* S, const float* __restrict__ lutG) {
    PipeShared& ps = pipe_header(smemRaw);
    float* sG = pipe_lut<ViscSetupOp>(smemRaw);                     // the boundary-friction samples still need the table
    load_lut_tile(sG, lutG);
    ViscSetupOp op{ P, A, Lut{ nullptr, sG, P.lutInvStep, P.lutRadius, P.lutRadius2 }, S->dt, 0.01f * P.h2, { 0.0f } };
    if (pipe_pass(S, A, ps, pipe_pay<ViscSetupOp>(smemRaw), op, P.tile0, P.tile1)) {
        double tot[1];
        fold_slots<1>(tot, A.slotSums, __ldg(A.tileList), A.slotStride, ps.red);
        if (threadIdx.x == 0) finish_reduction<1>(SITE_VISC_BB, P, S, tot);
    }
}

// ---- V3: matrix-free product with the viscosity operator -------------------------------------
__device__ __forceinline__ float3 mat_vec(const float* __restrict__ minv, uint32_t n, uint32_t p, float3 r) {
    // glm mat3 * vec3: m[0]*v.x + m[1]*v.y + m[2]*v.z
    float M[9];
    #pragma unroll
    for (int q = 0; q < 9; q++) M[q] = minv[(size_t)q * n + p];
    return f3(M[0] * r.x + M[3] * r.y + M[6] * r.z, M[1] * r.x + M[4] * r.y + M[7] * r.z, M[2] * r.x + M[5] * r.y + M[8] * r.z);
}

// What a pair needs of the neighbour is its position and the vector's value there: six floats.  They are kept as ONE
// 16-byte word (x, y, z, v.x) and one 8-byte word (v.y, v.z) per particle — A.cgXG/cgGyz for the start-up product with g,
// A.cgXP/cgPyz for the search direction p — so a neighbour costs an LDS.128 and an LDS.64 from the shared-memory stage
// (six wavefronts per warp and slot at best) instead of two LDS.128 on (x, y, z, rho) and (v, 0): the pass is bound by
// the shared-memory/L1 data path (profiles/r01_ncu_full_step.txt: 17.2 M wavefronts = 59 us of a 117-us launch).
// The position part never changes during a solve; whoever rewrites p rewrites the whole 16-byte word.
template<bool INIT>
struct ViscMatvecOp {
    static constexpr bool CUSTOM = false, PAD_SAFE = true;      // the pair term is linear in the pair coefficient
    using Cfg = PipeCfgMatvec;
    static constexpr int NPAY = 2, NOWN = 7, NSUM = 3, COEF = 1, NRED = INIT ? 2 : 1, NLUT = 0;
    const Params& P; const Arrays& A;
    float dt;
    float red[2];                    // INIT: |r_i|^2, r_i.z_i; else p_i.q_i
    static constexpr int BBYTES = 8;
    __device__ __forceinline__ const float4* srcA() const { return INIT ? A.cgXG : A.cgXP; }
    __device__ __forceinline__ const void* srcB() const { return INIT ? A.cgGyz : A.cgPyz; }
    __device__ __forceinline__ const float* coef_in() const { return A.coef; }
    __device__ __forceinline__ float* coef_out() const { return nullptr; }
    __device__ __forceinline__ void prefetch_own(uint32_t pt, uint32_t b0, uint32_t e0) const {
        if (P.muB != 0.0f && pt < P.nBodies) {
            l2_prefetch(A.bx[pt], (size_t)b0 * 16, (size_t)e0 * 16);
            l2_prefetch(A.bcoef[pt], (size_t)b0 * 16, (size_t)e0 * 16);
        }
        if (INIT && pt == 8) l2_prefetch(A.vel, (size_t)b0 * 16, (size_t)e0 * 16);
    }
    __device__ __forceinline__ float4 loadA(uint32_t g) const { return srcA()[g]; }
    __device__ __forceinline__ float4 loadB(uint32_t g) const { const float2 t = reinterpret_cast<const float2*>(srcB())[g]; return make_float4(t.x, t.y, 0.0f, 0.0f); }
    __device__ __forceinline__ void own_from(uint32_t p, float4 a, float4 b, float (&own)[NOWN]) const {
        own[0] = a.x; own[1] = a.y; own[2] = a.z; own[3] = a.w; own[4] = b.x; own[5] = b.y;
        own[6] = __ldg(A.rho + p);                          // for the epilogue; in flight during the gather
    }
    __device__ __forceinline__ void pair(const float (&o)[NOWN], float4 a, float4 b, float& c, float (&acc)[NSUM]) const {
        // the six differences as three packed operations (sm_100 FFMA2: a * -1 + o is o - a exactly): 10 floating-point
        // instructions per pair instead of 13, the same roundings in the same order as the scalar form
        const float2 m1 = make_float2(-1.0f, -1.0f);
        const float2 dxy = __ffma2_rn(make_float2(a.x, a.y), m1, make_float2(o[0], o[1]));       // (dx, dy)
        const float2 dzp = __ffma2_rn(make_float2(a.z, a.w), m1, make_float2(o[2], o[3]));       // (dz, p_i.x - p_j.x)
        const float2 dpq = __ffma2_rn(make_float2(b.x, b.y), m1, make_float2(o[4], o[5]));       // (p_i.y - p_j.y, p_i.z - p_j.z)
        const float w = c * __fmaf_rn(dpq.y, dzp.x, __fmaf_rn(dpq.x, dxy.y, dzp.y * dxy.x));
        acc[0] = __fmaf_rn(w, dxy.x, acc[0]); acc[1] = __fmaf_rn(w, dxy.y, acc[1]); acc[2] = __fmaf_rn(w, dzp.x, acc[2]);
    }
    __device__ __forceinline__ void finish(uint32_t p, uint32_t mf, const float (&o)[NOWN], const float (&acc)[NSUM]) {
        const float3 xi = f3(o[0], o[1], o[2]), vi = f3(o[3], o[4], o[5]);
        const float rhoI = o[6];
        float3 sum = f3(acc[0], acc[1], acc[2]);
        if (P.muB != 0.0f && (mf & VFD_NEAR_BODY)) {
            for (uint32_t b = 0; b < P.nBodies; b++) {
                const float4 bx = A.bx[b][p];
                if (bx.w > 0.0f) {
                    float3 pd[4];
                    if (boundary_samples(P, xi, f3(bx), pd)) {
                        const float4 cb = A.bcoef[b][p];
                        const float3 a0 = (cb.x * dot3(vi, pd[0])) * pd[0], a1 = (cb.y * dot3(vi, pd[1])) * pd[1];
                        const float3 a2 = (cb.z * dot3(vi, pd[2])) * pd[2], a3 = (cb.w * dot3(vi, pd[3])) * pd[3];
                        sum += ((a0 + a1) + a2) + a3;
                    }
                }
            }
        }
        const float3 q = vi - (dt / rhoI) * sum;
        if (INIT) {
            // r = b - A g ; p = M^-1 r ; |r|^2 ; r.p      (DFSPHImplementation.cu:638-691)
            const float3 r = f3(A.vel[p]) - q;
            const float3 z = mat_vec(A.minv, P.n, p, r);
            A.cgR[p] = make_float4(r.x, r.y, r.z, 0.0f);
            A.cgXP[p] = make_float4(xi.x, xi.y, xi.z, z.x);
            A.cgPyz[p] = make_float2(z.y, z.z);
            red[0] = (r.x * r.x + r.y * r.y) + r.z * r.z;
            red[1] = (r.x * z.x + r.y * z.y) + r.z * z.z;
        } else {
            A.cgQ[p] = make_float4(q.x, q.y, q.z, 0.0f);
            red[0] = (vi.x * q.x + vi.y * q.y) + vi.z * q.z;
        }
    }
};

// waitHalo: the neighbours' edge columns of the direction are written into this rank's ghost ranges by THEIR k_visc_step (no
// exchange kernel in between): wait for both "landed" flags first; the block that finishes last advances the exchange counter.
template<bool INIT>
__global__ void __launch_bounds__(ViscMatvecOp<INIT>::Cfg::THREADS, 1) k_visc_matvec_pipe(const __grid_constant__ Params P, const __grid_constant__ Arrays A, DevState* S, uint32_t waitHalo) {
    if (!INIT && S->viscActive != 1u) return;
    PipeShared& ps = pipe_header(smemRaw);
    uint32_t haloSeq = 0;
    if (!INIT && waitHalo) {
        if (threadIdx.x == 0) {
            PeerCtl* me = reinterpret_cast<PeerCtl*>(P.peerCtl[P.rank]);
            haloSeq = *(volatile uint32_t*)&me->haloSeq + 1u;
            if (P.rank > 0u) while (*(volatile uint32_t*)&me->haloData[0] != haloSeq) { }
            if (P.rank + 1u < P.nRanks) while (*(volatile uint32_t*)&me->haloData[1] != haloSeq) { }
            __threadfence_system();
        }
        __syncthreads();
    }
    ViscMatvecOp<INIT> op{ P, A, S->dt, { 0.0f, 0.0f } };
    if (pipe_pass(S, A, ps, pipe_pay<ViscMatvecOp<INIT>>(smemRaw), op, P.tile0, P.tile1)) {
        if (!INIT && waitHalo && threadIdx.x == 0) *(volatile uint32_t*)&reinterpret_cast<PeerCtl*>(P.peerCtl[P.rank])->haloSeq = haloSeq;
        double tot[2] = { 0.0, 0.0 };
        if (INIT) {
            fold_slots<2>(tot, A.slotSums, __ldg(A.tileList), A.slotStride, ps.red);
            if (threadIdx.x == 0) finish_reduction<2>(SITE_VISC_INIT, P, S, tot);
        } else {
            double t1[1];
            fold_slots<1>(t1, A.slotSums, __ldg(A.tileList), A.slotStride, ps.red);
            if (threadIdx.x == 0) finish_reduction<1>(SITE_VISC_PQ, P, S, t1);
        }
    }
}

// g += alpha p ; r -= alpha q ; |r|^2 ; z = M^-1 r ; r.z       (DFSPHImplementation.cu:714-779)
__global__ void __launch_bounds__(VFD_TPB) k_visc_update(Params P, Arrays A, DevState* S) {
    if (S->viscActive != 1u) return;
    __shared__ double shRed[64];
    const float alpha = S->alpha;
    double s0 = 0.0, s1 = 0.0;                  // sums of fp32 terms in fp64: the same value whatever the order (and the kernel)
    FOR_EACH_OWNED(p) {
        const float4 xp = A.cgXP[p], qq = A.cgQ[p];
        const float2 pyz = A.cgPyz[p];
        const float3 pp = f3(xp.w, pyz.x, pyz.y);
        float4 g = A.cgG[p], r = A.cgR[p];
        g.x = g.x + pp.x * alpha; g.y = g.y + pp.y * alpha; g.z = g.z + pp.z * alpha;
        r.x = r.x - qq.x * alpha; r.y = r.y - qq.y * alpha; r.z = r.z - qq.z * alpha;
        A.cgG[p] = g; A.cgR[p] = r;
        const float3 z = mat_vec(A.minv, P.n, p, f3(r));
        A.cgZ[p] = make_float4(z.x, z.y, z.z, 0.0f);
        s0 += (double)((r.x * r.x + r.y * r.y) + r.z * r.z);
        s1 += (double)((r.x * z.x + r.y * z.y) + r.z * z.z);
    }
    double v[2] = { s0, s1 };
    if (block_reduce_publish<2>(v, A.partials, &S->ticket[6], shRed)) {
        double tot[2];
        last_block_fold<2>(tot, A.partials, shRed);
        if (threadIdx.x == 0) {
            finish_reduction<2>(SITE_VISC_UPDATE, P, S, tot);
            S->ticket[6] = 0;
        }
    }
}

// p = beta p + z   (:788-802)
__global__ void __launch_bounds__(VFD_TPB) k_visc_direction(Params P, Arrays A, DevState* S) {
    if (S->viscActive != 1u) return;
    OWNED_INDEX(p);
    const float beta = S->beta;
    const float4 z = A.cgZ[p];
    float4 xp = A.cgXP[p];
    float2 pyz = A.cgPyz[p];
    xp.w = xp.w * beta + z.x; pyz.x = pyz.x * beta + z.y; pyz.y = pyz.y * beta + z.z;
    A.cgXP[p] = xp; A.cgPyz[p] = pyz;
}

// ---- the vector work of one PCG iteration in ONE launch -------------------------------------------------------------
// k_visc_update and k_visc_direction are separated by a grid-wide decision (|r|^2 against the threshold, beta) and nothing
// else: the direction update needs z and p of the SAME particle only.  On one GPU, with a scene that fits (<= STEP_MAXK x 1024
// particles per SM), both run as one cooperative launch of one CTA per SM: each CTA keeps the (x, y, z, p) words of its
// particles in shared memory and z in registers across a grid barrier (last block folds the partial sums, takes the solver's
// decision — control.cuh — and releases the others), then writes p = z + beta p.  Per iteration this saves a launch, the write
// and re-read of z and the re-read of p: 164 B per particle instead of 220.  Same expressions, same values as the two kernels.
#define STEP_THREADS 1024
#define STEP_MAXK 8
__global__ void __launch_bounds__(STEP_THREADS, 1) k_visc_step(const __grid_constant__ Params P, const __grid_constant__ Arrays A, DevState* S) {
    if (*(volatile uint32_t*)&S->viscActive != 1u) return;       // every block reads this before the last one can change it
    __shared__ double shRed[64];
    __shared__ uint32_t shGo;
    float4* sXP = reinterpret_cast<float4*>(smemRaw);
    float2* sPyz = reinterpret_cast<float2*>(smemRaw + (size_t)STEP_MAXK * STEP_THREADS * sizeof(float4));
    const uint32_t gen0 = *(volatile uint32_t*)&S->stepGen;
    uint32_t ownB, ownE;
    owned_range(P, A.cellBegin, ownB, ownE);
    const uint32_t chunk = ((ownE - ownB + gridDim.x - 1u) / gridDim.x + 31u) & ~31u;
    const uint32_t lo = min(ownB + blockIdx.x * chunk, ownE), hi = min(lo + chunk, ownE);
    const float alpha = S->alpha;
    float z[STEP_MAXK][3];
    double s0 = 0.0, s1 = 0.0;
    #pragma unroll
    for (int k = 0; k < STEP_MAXK; k++) {
        const uint32_t i = (uint32_t)k * STEP_THREADS + threadIdx.x, p = lo + i;
        z[k][0] = z[k][1] = z[k][2] = 0.0f;
        if (p < hi) {
            const float4 xp = A.cgXP[p], qq = A.cgQ[p];
            const float2 pyz = A.cgPyz[p];
            const float3 pp = f3(xp.w, pyz.x, pyz.y);
            float4 g = A.cgG[p], r = A.cgR[p];
            g.x = g.x + pp.x * alpha; g.y = g.y + pp.y * alpha; g.z = g.z + pp.z * alpha;
            r.x = r.x - qq.x * alpha; r.y = r.y - qq.y * alpha; r.z = r.z - qq.z * alpha;
            A.cgG[p] = g; A.cgR[p] = r;
            const float3 zz = mat_vec(A.minv, P.n, p, f3(r));
            z[k][0] = zz.x; z[k][1] = zz.y; z[k][2] = zz.z;
            sXP[i] = xp; sPyz[i] = pyz;
            s0 += (double)((r.x * r.x + r.y * r.y) + r.z * r.z);
            s1 += (double)((r.x * zz.x + r.y * zz.y) + r.z * zz.z);
        }
    }
    double v[2] = { s0, s1 };
    if (block_reduce_publish<2>(v, A.partials, &S->ticket[6], shRed)) {
        double tot[2];
        last_block_fold<2>(tot, A.partials, shRed);
        if (threadIdx.x == 0) {
            finish_reduction<2>(SITE_VISC_UPDATE, P, S, tot);
            S->ticket[6] = 0;
            __threadfence();
            atomicAdd(&S->stepGen, 1u);                         // release: the decision and beta are visible
        }
    }
    if (threadIdx.x == 0) {
        while (*(volatile uint32_t*)&S->stepGen == gen0) { }
        __threadfence();
        shGo = *(volatile uint32_t*)&S->viscActive;
    }
    __syncthreads();
    if (shGo != 1u) return;                                      // converged (or the iteration cap): no next direction
    const float beta = *(volatile float*)&S->beta;
    #pragma unroll
    for (int k = 0; k < STEP_MAXK; k++) {
        const uint32_t i = (uint32_t)k * STEP_THREADS + threadIdx.x, p = lo + i;
        if (p < hi) {
            float4 xp = sXP[i];
            float2 pyz = sPyz[i];
            xp.w = xp.w * beta + z[k][0]; pyz.x = pyz.x * beta + z[k][1]; pyz.y = pyz.y * beta + z[k][2];
            A.cgXP[p] = xp; A.cgPyz[p] = pyz;
            // several ranks over peer memory: the edge columns' direction goes straight into the neighbours' ghost ranges.  They
            // are done reading the previous direction: this rank's mat-vec ended with an all-reduce every rank's mat-vec fed
            // after its last tile.
            if (P.haloP[0][0] && p >= P.haloRange[0] && p < P.haloRange[1]) {
                const uint32_t j = p - P.haloRange[0];
                reinterpret_cast<float4*>(P.haloP[0][0])[j] = xp; reinterpret_cast<float2*>(P.haloP[0][1])[j] = pyz;
            }
            if (P.haloP[1][0] && p >= P.haloRange[2] && p < P.haloRange[3]) {
                const uint32_t j = p - P.haloRange[2];
                reinterpret_cast<float4*>(P.haloP[1][0])[j] = xp; reinterpret_cast<float2*>(P.haloP[1][1])[j] = pyz;
            }
        }
    }
    if (P.nRanks > 1u && P.peerCtl[0] && (P.haloP[0][0] || P.haloP[1][0])) {
        // "the direction has landed": told to both neighbours by the block that finishes last; their next mat-vec waits for it
        __threadfence_system();
        __syncthreads();
        if (threadIdx.x == 0 && atomicAdd(&S->ticket[7], 1u) == gridDim.x - 1u) {
            S->ticket[7] = 0u;
            __threadfence_system();
            const uint32_t seq = *(volatile uint32_t*)&reinterpret_cast<PeerCtl*>(P.peerCtl[P.rank])->haloSeq + 1u;
            if (P.rank > 0u) *(volatile uint32_t*)&reinterpret_cast<PeerCtl*>(P.peerCtl[P.rank - 1u])->haloData[1] = seq;
            if (P.rank + 1u < P.nRanks) *(volatile uint32_t*)&reinterpret_cast<PeerCtl*>(P.peerCtl[P.rank + 1u])->haloData[0] = seq;
        }
    }
}

// V5: a += (g - v)/dt ; dv = g - v
__global__ void __launch_bounds__(VFD_TPB) k_visc_apply(Params P, Arrays A, DevState* S) {
    OWNED_INDEX(p);
    const float dt = S->dt;
    float4 g = A.cgG[p];
    if (S->viscActive == 2u) g = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    const float4 v = A.vel[p];
    const float3 d = f3(g.x - v.x, g.y - v.y, g.z - v.z);
    const float s = 1.0f / dt;
    float4 a = A.acc[p];
    a.x += s * d.x; a.y += s * d.y; a.z += s * d.z;
    A.acc[p] = a;
    A.dv[p] = make_float4(d.x, d.y, d.z, 0.0f);
}

template<typename Kern>
static void pipe_attr(Kern kern, size_t smem) {
    static thread_local const void* done[16]; static thread_local int nd = 0;
    static thread_local int dev = -1;
    if (launch_device_changed(dev)) nd = 0;
    for (int i = 0; i < nd; i++) if (done[i] == (const void*)kern) return;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (nd < 16) done[nd++] = (const void*)kern;
}

void launch_viscosity_setup(const LaunchCfg& L, const Params& P, const Arrays& A, DevState* S, const float* lutG) {
    const size_t sp = pipe_smem_bytes<ViscSetupOp>();
    pipe_attr(k_visc_setup, sp);
    LaunchScope ls(L, KID_VISC_SETUP);
    k_visc_setup<<<L.numSMs, ViscSetupOp::Cfg::THREADS, sp, L.stream>>>(P, A, S, lutG);
}

void launch_viscosity_matvec(const LaunchCfg& L, const Params& P, const Arrays& A, DevState* S, bool init, bool waitHalo) {
    const size_t sp = pipe_smem_bytes<ViscMatvecOp<false>>();
    LaunchScope ls(L, init ? KID_VISC_MATVEC0 : KID_VISC_MATVEC);
    if (init) { pipe_attr(k_visc_matvec_pipe<true>, sp); k_visc_matvec_pipe<true><<<L.numSMs, ViscMatvecOp<true>::Cfg::THREADS, sp, L.stream>>>(P, A, S, 0u); }
    else      { pipe_attr(k_visc_matvec_pipe<false>, sp); k_visc_matvec_pipe<false><<<L.numSMs, ViscMatvecOp<false>::Cfg::THREADS, sp, L.stream>>>(P, A, S, waitHalo ? 1u : 0u); }
}
#ifdef PIPE_TRACE
// diagnostic builds: one traced launch of the initial mat-vec; the event log of CTA 0 goes to `path` (u64 pairs)
void trace_viscosity_matvec(const LaunchCfg& L, const Params& P, const Arrays& A, DevState* S, const char* path) {
    unsigned int zero = 0, one = 1;
    cudaMemcpyToSymbol(g_pipeTraceN, &zero, 4); cudaMemcpyToSymbol(g_pipeTraceOn, &one, 4);
    // the PCG's own product (q = A p) unless VFD_TRACE_INIT is set; it only runs while the solver is active
    const bool init = getenv("VFD_TRACE_INIT") != nullptr;
    uint32_t active = 1u, saved = 0u;
    cudaMemcpy(&saved, &S->viscActive, 4, cudaMemcpyDeviceToHost);
    if (!init) cudaMemcpy(&S->viscActive, &active, 4, cudaMemcpyHostToDevice);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 3; i++) launch_viscosity_matvec(L, P, A, S, init);      // warm: the traced launch is a steady-state one
    cudaStreamSynchronize(L.stream);
    cudaMemcpyToSymbol(g_pipeTraceN, &zero, 4);
    cudaEventRecord(e0, L.stream);
    launch_viscosity_matvec(L, P, A, S, init);
    cudaEventRecord(e1, L.stream);
    cudaStreamSynchronize(L.stream);
    float ms = 0.0f; cudaEventElapsedTime(&ms, e0, e1);
    printf("traced launch (%s): %.4f ms\n", init ? "r = b - A g" : "q = A p", ms);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaMemcpy(&S->viscActive, &saved, 4, cudaMemcpyHostToDevice);
    cudaMemcpyToSymbol(g_pipeTraceOn, &zero, 4);
    unsigned int n = 0;
    cudaMemcpyFromSymbol(&n, g_pipeTraceN, 4);
    if (n > PIPE_TRACE_CAP) n = PIPE_TRACE_CAP;
    unsigned long long* h = (unsigned long long*)malloc((size_t)n * 16);
    cudaMemcpyFromSymbol(h, g_pipeTrace, (size_t)n * 16);
    if (FILE* f = fopen(path, "wb")) { fwrite(h, 16, n, f); fclose(f); }
    free(h);
}
#endif
void launch_viscosity_update(const LaunchCfg& L, const Params& P, const Arrays& A, DevState* S) {
    const uint32_t tiles = std::max(1u, (P.n + VFD_TPB - 1) / VFD_TPB);
    const uint32_t g2 = std::max(1u, std::min<uint32_t>(tiles, (uint32_t)L.numSMs * 8u));
    LaunchScope ls(L, KID_VISC_UPDATE);
    k_visc_update<<<g2, VFD_TPB, 0, L.stream>>>(P, A, S);
}
// true when the fused kernel can carry this scene: the decision is taken inside the kernel (one rank, or ranks that all-reduce
// through peer memory) and the rank holds <= STEP_MAXK x 1024 particles per SM
bool viscosity_step_fits(const LaunchCfg& L, const Params& P) {
    // the kernel walks the OWNED particles only (ghost copies of the neighbour slabs are not updated here)
    const uint64_t owned = P.nRanks > 1u ? (uint64_t)(P.haloRange[3] - P.haloRange[0]) : (uint64_t)P.n;
    return (P.nRanks == 1 || P.peerCtl[0] != nullptr) && P.tune[5] != 1 && owned <= (uint64_t)L.numSMs * STEP_MAXK * STEP_THREADS - 32ull * (uint64_t)L.numSMs;
}
int launch_viscosity_step(const LaunchCfg& L, const Params& P, const Arrays& A, DevState* S) {
    const size_t sm = (size_t)STEP_MAXK * STEP_THREADS * (sizeof(float4) + sizeof(float2));
    pipe_attr(k_visc_step, sm);
    LaunchScope ls(L, KID_VISC_STEP);
    void* args[] = { (void*)&P, (void*)&A, (void*)&S };
    // cooperative: the grid barrier needs every CTA resident (one per SM; the launch fails rather than deadlock)
    return (int)cudaLaunchCooperativeKernel((const void*)k_visc_step, dim3((unsigned)L.numSMs), dim3(STEP_THREADS), args, sm, L.stream);
}
void launch_viscosity_direction(const LaunchCfg& L, const Params& P, const Arrays& A, DevState* S) {
    LaunchScope ls(L, KID_VISC_DIRECTION);
    k_visc_direction<<<std::max(1u, (P.n + VFD_TPB - 1) / VFD_TPB), VFD_TPB, 0, L.stream>>>(P, A, S);
}
void launch_viscosity_apply(const LaunchCfg& L, const Params& P, const Arrays& A, DevState* S) {
    LaunchScope ls(L, KID_VISC_APPLY);
    k_visc_apply<<<std::max(1u, (P.n + VFD_TPB - 1) / VFD_TPB), VFD_TPB, 0, L.stream>>>(P, A, S);
}

} // namespace vfd
