// solver.cu — host orchestration of one DFSPH step on one B200 and the dump/restore kernels.
//
// Solver::step() is DFSPHImplementation::OnUpdate (reference: DFSPHImplementation.cu:63-170) re-expressed
// as a stream of kernels with no host round trip unless a solver is convergence-driven (then the
// continue flag is polled with one batch of look-ahead).  Solver::simulate() is Simulate() (:33-61).
#include "solver_impl.h"
#include "tile.cuh"
#include "control.cuh"
#include <algorithm>
#include <cstring>
#include <limits.h>

namespace vfd {
#ifdef PIPE_TRACE
void trace_viscosity_matvec(const LaunchCfg& L, const Params& P, const Arrays& A, DevState* S, const char* path);
#endif

#define CK(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) return fail_cuda(e__, #call, __LINE__); } while (0)

// ---------------------------------------------------------------------------------------------
// dump / restore kernels
// ---------------------------------------------------------------------------------------------
__global__ void k_export_aos(Params P, Arrays A, VfdParticle* __restrict__ out) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P.n) return;
    VfdParticle q;
    const float4 x = A.pos[p], v = A.vel[p], a = A.acc[p], pa = A.pacc[p], dv = A.dv[p], n = A.nrm[p], nb = A.nbar[p];
    q.Position[0] = x.x; q.Position[1] = x.y; q.Position[2] = x.z;
    q.Velocity[0] = v.x; q.Velocity[1] = v.y; q.Velocity[2] = v.z;
    q.Acceleration[0] = a.x; q.Acceleration[1] = a.y; q.Acceleration[2] = a.z;
    q.PressureAcceleration[0] = pa.x; q.PressureAcceleration[1] = pa.y; q.PressureAcceleration[2] = pa.z;
    q.PressureResiduum = A.res[p]; q.Density = A.rho[p]; q.DensityAdvection = A.rhoAdv[p];
    q.PressureRho2 = A.kappa[p]; q.PressureRho2V = A.kappaV[p]; q.Factor = A.alpha[p];
    q.VelocityDifference[0] = dv.x; q.VelocityDifference[1] = dv.y; q.VelocityDifference[2] = dv.z;
    q.MonteCarloSurfaceNormal[0] = n.x; q.MonteCarloSurfaceNormal[1] = n.y; q.MonteCarloSurfaceNormal[2] = n.z;
    q.MonteCarloSurfaceNormalSmooth[0] = nb.x; q.MonteCarloSurfaceNormalSmooth[1] = nb.y; q.MonteCarloSurfaceNormalSmooth[2] = nb.z;
    q.MonteCarloSurfaceCurvature = A.curv[p]; q.MonteCarloSurfaceCurvatureSmooth = A.curvS[p]; q.DeltaFinalCurvature = A.curvD[p];
    out[A.id[p]] = q;
}

__global__ void k_import_aos(Params P, Arrays A, const VfdParticle* __restrict__ in) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P.n) return;
    const VfdParticle q = in[p];
    A.pos[p] = make_float4(q.Position[0], q.Position[1], q.Position[2], 0.0f);
    A.posRho[p] = make_float4(q.Position[0], q.Position[1], q.Position[2], q.Density);
    A.vel[p] = make_float4(q.Velocity[0], q.Velocity[1], q.Velocity[2], 0.0f);
    A.acc[p] = make_float4(q.Acceleration[0], q.Acceleration[1], q.Acceleration[2], 0.0f);
    A.pacc[p] = make_float4(q.PressureAcceleration[0], q.PressureAcceleration[1], q.PressureAcceleration[2], 0.0f);
    A.res[p] = q.PressureResiduum; A.rho[p] = q.Density; A.rhoAdv[p] = q.DensityAdvection;
    A.kappa[p] = q.PressureRho2; A.kappaV[p] = q.PressureRho2V; A.alpha[p] = q.Factor;
    A.dv[p] = make_float4(q.VelocityDifference[0], q.VelocityDifference[1], q.VelocityDifference[2], 0.0f);
    A.nrm[p] = make_float4(q.MonteCarloSurfaceNormal[0], q.MonteCarloSurfaceNormal[1], q.MonteCarloSurfaceNormal[2], q.MonteCarloSurfaceCurvature);
    A.nbar[p] = make_float4(q.MonteCarloSurfaceNormalSmooth[0], q.MonteCarloSurfaceNormalSmooth[1], q.MonteCarloSurfaceNormalSmooth[2], 0.0f);
    A.curv[p] = q.MonteCarloSurfaceCurvature; A.curvS[p] = q.MonteCarloSurfaceCurvatureSmooth; A.curvD[p] = q.DeltaFinalCurvature;
    A.id[p] = p;
    A.cnt[p] = 0;
}

// initial state: SetFluidObjects zero-fills everything but position and velocity (DFSPHImplementation.cu:198-220)
__global__ void k_reset_state(Params P, Arrays A, const float4* __restrict__ pos0, const float4* __restrict__ vel0, const uint32_t* __restrict__ ids0) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P.n) return;
    const float4 z = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    A.pos[p] = pos0[p]; A.posRho[p] = pos0[p]; A.vel[p] = vel0[p];
    A.acc[p] = z; A.pacc[p] = z; A.dv[p] = z; A.nrm[p] = z; A.nbar[p] = z;
    A.res[p] = 0.0f; A.rho[p] = 0.0f; A.rhoAdv[p] = 0.0f; A.kappa[p] = 0.0f; A.kappaV[p] = 0.0f; A.alpha[p] = 0.0f;
    A.curv[p] = 0.0f; A.curvS[p] = 0.0f; A.curvD[p] = 0.0f;
    A.id[p] = ids0 ? ids0[p] : p; A.cnt[p] = 0;
    for (uint32_t b = 0; b < P.nBodies; b++) A.bx[b][p] = z;
}

__global__ void k_pack_posvel(uint32_t n, const float* __restrict__ pos, const float* __restrict__ vel, float4* __restrict__ pos4, float4* __restrict__ vel4) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    pos4[i] = make_float4(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2], 0.0f);
    vel4[i] = vel ? make_float4(vel[3 * i], vel[3 * i + 1], vel[3 * i + 2], 0.0f) : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
}

// The frame decision of DFSPHImplementation::OnUpdate (:148: FrameTime >= FrameLength, FrameTime accumulated with every new
// time step in the CFL update, :427) taken on the device, so that the host need not read the time step back every step.
__global__ void k_frame_decide(DevState* S, float frameLength, uint32_t frameCount) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    if (S->frameIndex < frameCount && S->frameTime >= frameLength) { S->captureFlag = 1u; S->frameTime = 0.0f; S->frameIndex += 1u; }
    else S->captureFlag = 0u;
}

// K15: ConvertParticlesToBuffer (DFSPHKernels.cu:6-22), written in original particle order.  conditional: only if
// k_frame_decide said so (meta[2] tells the frame pipe's worker)
__global__ void k_export_frame(Params P, Arrays A, VfdParticleSimple* __restrict__ out, const DevState* __restrict__ S, float* __restrict__ meta, int conditional) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    const bool capture = !conditional || S->captureFlag != 0u;
    if (p == 0 && meta) { meta[0] = S->vmax2; meta[1] = S->dt; meta[2] = capture ? 1.0f : 0.0f; }     // DFSPHParticleFrame::MaxVelocityMagnitude, CurrentTimeStep
    if (!capture) return;
    if (p >= P.n) return;
    const float4 x = A.pos[p], v = A.vel[p], a = A.acc[p];
    VfdParticleSimple q;
    q.Position[0] = x.x; q.Position[1] = x.y; q.Position[2] = x.z;
    q.Velocity[0] = v.x; q.Velocity[1] = v.y; q.Velocity[2] = v.z;
    q.Acceleration[0] = a.x; q.Acceleration[1] = a.y; q.Acceleration[2] = a.z;
    out[A.id[p]] = q;
}

// neighbour lists hold tile-local indices: translate them back per tile (thread per particle; test/inspection only)
__global__ void __launch_bounds__(512) k_export_neighbors(const __grid_constant__ Arrays A, const DevState* S, uint32_t* __restrict__ counts, uint32_t* __restrict__ idsPadded) {
    __shared__ TileShared sh;
    const uint32_t nTiles = S->nTiles;
    for (uint32_t tile = blockIdx.x; tile < nTiles; tile += gridDim.x) {
        const uint32_t b0 = A.cellBegin[tile * TILE_CELLS], e0 = A.cellBegin[tile * TILE_CELLS + TILE_CELLS];
        if (b0 == e0) continue;
        __syncthreads();
        const TileInfo t = tile_setup(S, A.cellBegin, tile, sh, 0u);
        (void)t.staged;
        for (uint32_t p = t.begin + threadIdx.x; p < t.end; p += blockDim.x) {
            const uint32_t o = A.id[p], m = A.cnt[p] & VFD_COUNT_MASK;
            counts[o] = m;
            const uint2* col = ell_list(A.list16, p);
            for (uint32_t k = 0; k < m; k++) {
                uint32_t L[4];
                ell_unpack(col[(size_t)(k >> 2) * 32], L);
                idsPadded[(size_t)o * VFD_MAX_NEIGHBORS + k] = A.id[tile_local_to_global(sh, L[k & 3u])];
            }
        }
    }
}

__global__ void k_export_boundary(Params P, Arrays A, uint32_t body, float* __restrict__ xj, float* __restrict__ vol) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P.n) return;
    const uint32_t o = A.id[p];
    const float4 b = A.bx[body][p];
    xj[3 * (size_t)o] = b.x; xj[3 * (size_t)o + 1] = b.y; xj[3 * (size_t)o + 2] = b.z;
    vol[o] = b.w;
}

static inline uint32_t nblk(uint32_t n) { return n ? (n + VFD_TPB - 1) / VFD_TPB : 1u; }     // never a zero-block launch (a rank that owns no particle yet); the kernels bound-check

// ---------------------------------------------------------------------------------------------
// per-kernel device timers
// ---------------------------------------------------------------------------------------------
const char* const kKernelNames[KID_COUNT] = {
    "search_bounds", "search_hist", "search_scan", "search_scatter", "search_reorder", "search_build_list",
    "boundary", "density_factor",
    "div_source", "div_accel", "div_solve", "div_finish",
    "st_classify", "st_smooth", "st_apply",
    "visc_setup", "visc_matvec0", "visc_matvec", "visc_update", "visc_direction", "visc_step", "visc_apply",
    "cfl", "velocity",
    "press_source", "press_accel", "press_solve", "press_finish",
    "position", "clear_acc", "io" };

cudaEvent_t KernelProf::take() {
    if (used == pool.size()) { cudaEvent_t e; cudaEventCreate(&e); pool.push_back(e); }
    return pool[used++];
}
void KernelProf::begin(int kid, cudaStream_t s) {
    Rec r{ kid, take(), take() };
    cudaEventRecord(r.a, s);
    pending.push_back(r);
}
void KernelProf::end(cudaStream_t s) { cudaEventRecord(pending.back().b, s); }
void KernelProf::drain() {
    std::vector<float> dur(pending.size());
    float mx[KID_COUNT] = {};
    for (size_t i = 0; i < pending.size(); i++) {
        float t = 0.0f;
        cudaEventElapsedTime(&t, pending[i].a, pending[i].b);
        dur[i] = t;
        mx[pending[i].kid] = std::max(mx[pending[i].kid], t);
    }
    for (size_t i = 0; i < pending.size(); i++) {
        const int k = pending[i].kid;
        ms[k] += dur[i]; launches[k]++;
        if (dur[i] >= 0.3f * mx[k]) { msActive[k] += dur[i]; launchesActive[k]++; }
    }
    pending.clear();
    used = 0;
}
void KernelProf::reset() {
    for (int k = 0; k < KID_COUNT; k++) { ms[k] = msActive[k] = 0.0; launches[k] = launchesActive[k] = 0; }
}
KernelProf::~KernelProf() { for (cudaEvent_t e : pool) cudaEventDestroy(e); }

// ---------------------------------------------------------------------------------------------
// Solver
// ---------------------------------------------------------------------------------------------
int Solver::fail(int code, const std::string& msg) {
    std::lock_guard<std::mutex> g(errMutex);
    lastError = msg;
    return code;
}
int Solver::fail_cuda(cudaError_t e, const char* what, int line) {
    char buf[512];
    snprintf(buf, sizeof buf, "CUDA error %d (%s) at solver.cu:%d: %s", (int)e, cudaGetErrorString(e), line, what);
    cudaGetLastError();
    return fail(VFD_E_CUDA, buf);
}

int Solver::init(const VfdDfsphDescription& d, int dev) {
    device = dev;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) { cudaGetLastError(); return fail(VFD_E_CUDA, "no CUDA device available: libvfd_dfsph has no CPU fallback"); }
    if (dev < 0 || dev >= count) return fail(VFD_E_INVALID, "device index out of range");
    CK(cudaSetDevice(device));
    CK(cudaDeviceGetAttribute(&numSMs, cudaDevAttrMultiProcessorCount, device));
    CK(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    CK(cudaMalloc(&dState, sizeof(DevState)));
    CK(cudaMemset(dState, 0, sizeof(DevState)));
    CK(cudaMallocHost(&hState, sizeof(DevState) * 4));
    CK(cudaMallocHost(&hFlags, sizeof(uint32_t) * 64));
    for (int i = 0; i < 4; i++) CK(cudaEventCreateWithFlags(&pollEvent[i], cudaEventDisableTiming));
    for (int i = 0; i < 7; i++) CK(cudaEventCreate(&phaseEvent[i]));
    CK(cudaMalloc(&dLutW, VFD_LUT_RES * sizeof(float)));
    CK(cudaMalloc(&dLutG, VFD_LUT_RES * sizeof(float)));
    CK(cudaMalloc(&dHalton, VFD_HALTON_N * sizeof(float)));
    build_halton_table(halton);
    CK(cudaMemcpy(dHalton, halton.data(), VFD_HALTON_N * sizeof(float), cudaMemcpyHostToDevice));
    memset(&desc, 0, sizeof desc);
    memset(&info, 0, sizeof info);          // zero start state
    return set_description(d);
}

Solver::~Solver() {
    if (device >= 0) cudaSetDevice(device);
    pipe.drain();
    free_particles();
    free_bodies();
    if (dStagePos) cudaFree(dStagePos);
    if (dStageVel) cudaFree(dStageVel);
    cudaFree(dState); cudaFreeHost(hState); cudaFreeHost(hFlags);
    cudaFree(dLutW); cudaFree(dLutG); cudaFree(dHalton);
    for (int i = 0; i < 4; i++) if (pollEvent[i]) cudaEventDestroy(pollEvent[i]);
    for (int i = 0; i < 7; i++) if (phaseEvent[i]) cudaEventDestroy(phaseEvent[i]);
    for (int i = 0; i < 16; i++) if (userEvent[i]) cudaEventDestroy(userEvent[i]);
    if (stream) cudaStreamDestroy(stream);
}

// SetDescription (DFSPHImplementation.cu:318-368): all fp32 host arithmetic, same operations
int Solver::set_description(const VfdDfsphDescription& d) {
    CK(cudaSetDevice(device));
    desc = d;
    if (desc.ParticleRadius != info.ParticleRadius) {
        tables.build(4.0f * desc.ParticleRadius);
        CK(cudaMemcpy(dLutW, tables.Wc.data(), VFD_LUT_RES * sizeof(float), cudaMemcpyHostToDevice));
        CK(cudaMemcpy(dLutG, tables.Gc.data(), VFD_LUT_RES * sizeof(float), cudaMemcpyHostToDevice));
    }
    info.ParticleRadius = desc.ParticleRadius;
    info.ParticleDiameter = 2.0f * info.ParticleRadius;
    info.SupportRadius = 4.0f * info.ParticleRadius;
    info.SupportRadius2 = info.SupportRadius * info.SupportRadius;
    info.Volume = 0.8f * info.ParticleDiameter * info.ParticleDiameter * info.ParticleDiameter;
    info.Density0 = 1000.0f;
    info.ParticleMass = info.Volume * info.Density0;
    info.ParticleMassInverse = 1.0f / info.ParticleMass;
    info.Viscosity = desc.Viscosity;
    info.BoundaryViscosity = desc.BoundaryViscosity;
    info.DynamicViscosity = info.Viscosity * info.Density0;
    info.DynamicBoundaryViscosity = info.BoundaryViscosity * info.Density0;
    info.TangentialDistanceFactor = desc.TangentialDistanceFactor;
    info.TangentialDistance = info.TangentialDistanceFactor * info.SupportRadius;
    info.TimeStepSize = desc.TimeStepSize;
    info.TimeStepSize2 = desc.TimeStepSize * desc.TimeStepSize;
    info.TimeStepSizeInverse = 1.0f / info.TimeStepSize;
    info.TimeStepSize2Inverse = 1.0f / info.TimeStepSize2;
    info.SurfaceTension = desc.SurfaceTension;
    info.ClassifierSlope = 74.688796680497925f;
    info.ClassifierConstant = 12.0f;
    info.TemporalSmoothing = desc.TemporalSmoothing ? 1 : 0;
    info.SmoothingFactor = 0.5f;
    info.Factor = 0.8f;
    info.NeighborParticleRadius = info.ParticleRadius * info.Factor;
    info.Gravity[0] = desc.Gravity[0]; info.Gravity[1] = desc.Gravity[1]; info.Gravity[2] = desc.Gravity[2];
    refresh_params();
    return VFD_OK;
}

void Solver::refresh_params() {
    Params& P = params;
    P.n = info.ParticleCount; P.nBodies = info.RigidBodyCount;
    P.nGlobal = info.ParticleCount; P.nRanks = 1; P.rank = 0; P.tile0 = 0; P.tile1 = 0xffffffffu;
    P.h = info.SupportRadius; P.h2 = info.SupportRadius2; P.r = info.ParticleRadius; P.d = info.ParticleDiameter;
    P.volume = info.Volume; P.rho0 = info.Density0; P.mass = info.ParticleMass; P.massInv = info.ParticleMassInverse;
    P.mu = info.DynamicViscosity; P.muB = info.DynamicBoundaryViscosity; P.tangentialDistance = info.TangentialDistance;
    P.sigma = info.SurfaceTension; P.clsSlope = info.ClassifierSlope; P.clsConst = info.ClassifierConstant;
    P.smoothing = info.SmoothingFactor; P.nbrRadius = info.NeighborParticleRadius;
    // MonteCarloFactor = asin(NeighborParticleRadius / ParticleRadius) (DFSPHImplementation.cu:422), host libm like the reference
    P.mcFactor = asinf(info.NeighborParticleRadius / info.ParticleRadius);
    P.temporalSmoothing = info.TemporalSmoothing;
    P.gx = info.Gravity[0]; P.gy = info.Gravity[1]; P.gz = info.Gravity[2];
    P.lutInvStep = tables.invStep; P.lutRadius = tables.radius; P.lutRadius2 = tables.radius2; P.wZero = tables.wZero;
    P.minDt = desc.MinTimeStepSize; P.maxDt = desc.MaxTimeStepSize;
    P.csdFix = desc.CSDFix; P.csd = desc.CSD;
    P.frameLength = desc.FrameLength;
    P.etaPressure = desc.MaxPressureSolverError * 0.0001f * info.Density0;
    P.divErrScale = desc.MaxDivergenceSolverError * 0.0001f * info.Density0;
    P.viscErr2 = (desc.MaxViscositySolverError * desc.MaxViscositySolverError) * 0.0001f;
    P.minPressIt = desc.MinPressureSolverIterations; P.maxPressIt = desc.MaxPressureSolverIterations;
    P.minDivIt = desc.MinDivergenceSolverIterations; P.maxDivIt = desc.MaxDivergenceSolverIterations;
    P.minViscIt = desc.MinViscositySolverIterations; P.maxViscIt = desc.MaxViscositySolverIterations;
    P.searchFma = optSearchFma;
    for (int k = 0; k < 8; k++) { char nm[16]; snprintf(nm, sizeof nm, "VFD_TUNE%d", k); const char* e = getenv(nm); P.tune[k] = e ? atoi(e) : 0; }
    if (dist) dist_params(P);
}

// + 32 elements: the bulk copies of an 8-byte payload array read whole 16-byte granules, i.e. up to one element past a range's end (tile.cuh)
template<typename T> static cudaError_t dalloc(T*& p, size_t count) {
    const size_t bytes = (std::max<size_t>(count, 1) + 32) * sizeof(T);
    cudaError_t e = cudaMalloc((void**)&p, bytes);
    return e == cudaSuccess ? cudaMemset(p, 0, bytes) : e;        // the slack is read (never used): keep it defined
}

void Solver::free_particles() {
    Arrays& A = arrays;
    dist_free_slab();                        // the arrays that live in the peer-memory slab (several ranks) are not freed one by one
    void* ptrs[] = { A.pos, A.vel, A.dv, A.nbar, A.curv, A.curvS, A.curvD, A.id, A.pos2, A.vel2, A.dv2, A.nbar2, A.curv2, A.curvS2, A.curvD2, A.id2,
                     A.posRho, A.acc, A.pacc, A.nrm, A.res, A.rho, A.rhoAdv, A.kappa, A.kappaV, A.alpha, A.cgG, A.cgR, A.cgQ, A.cgZ, A.cgXG, A.cgXP, A.cgGyz, A.cgPyz, A.minv,
                     A.cnt, A.list16, A.coef, A.gcoef, A.key, A.rank, A.tmpIdx, A.cellCount, A.cellBegin, A.tileSums, A.tileList, A.partials, A.slotSums, dPos0, dVel0 };
    for (void* p : ptrs) if (p && !(pcgArena && (char*)p >= (char*)pcgArena && (char*)p < (char*)pcgArena + pcgArenaBytes)) cudaFree(p);
    if (pcgArena) { cudaFree(pcgArena); pcgArena = nullptr; pcgArenaBytes = 0; }
    if (dIds0) { cudaFree(dIds0); dIds0 = nullptr; }
    if (dFrame) { cudaFree(dFrame); dFrame = nullptr; dFrameCapacity = 0; }
    for (int b = 0; b < VFD_MAX_BODIES; b++) { if (A.bx[b]) cudaFree(A.bx[b]); if (A.bcoef[b]) cudaFree(A.bcoef[b]); if (A.bgrad[b]) cudaFree(A.bgrad[b]); }
    memset(&A, 0, sizeof A);
    bodySampleSlots = 0;
    dPos0 = dVel0 = nullptr;
    allocBytes = 0;
    allocParticles = 0;
}

// L2 residency of an address range for everything launched on the solver's stream from now on (persisting carve-out of the
// L2 cache; VFD_L2_PERSIST_MB caps it, 0 switches it off)
void Solver::l2_window(void* base, size_t bytes) {
    const char* e = getenv("VFD_L2_PERSIST_MB");
    int maxPersist = 0, maxWindow = 0;
    cudaDeviceGetAttribute(&maxPersist, cudaDevAttrMaxPersistingL2CacheSize, device);
    cudaDeviceGetAttribute(&maxWindow, cudaDevAttrMaxAccessPolicyWindowSize, device);
    size_t carve = e ? (size_t)atoll(e) << 20 : (size_t)maxPersist;
    carve = std::min(carve, (size_t)maxPersist);
    cudaStreamAttrValue attr;
    memset(&attr, 0, sizeof attr);
    if (carve == 0 || bytes == 0 || maxWindow <= 0) {
        attr.accessPolicyWindow.num_bytes = 0;
        cudaStreamSetAttribute(stream, cudaStreamAttributeAccessPolicyWindow, &attr);
        cudaGetLastError();
        return;
    }
    cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, carve);
    const size_t win = std::min(bytes, (size_t)maxWindow);
    attr.accessPolicyWindow.base_ptr = base;
    attr.accessPolicyWindow.num_bytes = win;
    attr.accessPolicyWindow.hitRatio = (float)std::min(1.0, (double)carve / (double)win);
    attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    cudaStreamSetAttribute(stream, cudaStreamAttributeAccessPolicyWindow, &attr);
    cudaGetLastError();
    l2Carve = carve; l2Window = win;
}

int Solver::alloc_particles(uint32_t n, const float* bboxMin, const float* bboxMax) {
    Arrays& A = arrays;
    const size_t np = ((size_t)(dist ? std::max(dist->capacity, n) : n) + 31) / 32 * 32;
    // Re-baking the same scene (the editor's "Simulate" button: Simulate() restarts from the stored initial state,
    // DFSPHImplementation.cu:36-51) keeps every buffer: same particle capacity, same bodies, grid capacity still enough.
    {
        double cells0 = 1.0;
        for (int k = 0; k < 3; k++) cells0 *= std::ceil((double)(bboxMax[k] - bboxMin[k]) / info.SupportRadius) + 9.0;
        bool bodiesOk = true;
        for (uint32_t b = 0; b < VFD_MAX_BODIES; b++) bodiesOk = bodiesOk && ((A.bx[b] != nullptr) == (b < info.RigidBodyCount));
        if (!dist && A.pos && allocParticles == np && bodiesOk && 8.0 * cells0 <= (double)cellCapacity) {
            cellEstimate = (uint32_t)std::min<double>(cells0, (double)optMaxCells);
            return VFD_OK;
        }
    }
    free_particles();
    allocParticles = np;
    // several ranks: the arrays that halo exchanges touch come from one slab that the other ranks map (distributed.cu)
    if (dist) { int rc = dist_alloc_slab(np); if (rc) return rc; allocBytes += dist->slabBytes; }
    // One GPU: the vectors every PCG iteration reads and rewrites — direction (x, y, z, p.x) + (p.y, p.z), q = A p, residual,
    // solution: 72 B per particle — come from one arena, so that one access-policy window can keep them resident in L2 across
    // the ~45 iterations of a solve while the neighbour list and the pair coefficients (430 MB per product) stream past them.
    if (!dist) {
        pcgArenaBytes = (np + 32) * 72;
        CK(cudaMalloc(&pcgArena, pcgArenaBytes));
        CK(cudaMemset(pcgArena, 0, pcgArenaBytes));
        char* a = (char*)pcgArena;
        A.cgXP = (float4*)a; a += (np + 8) * 16;
        A.cgQ = (float4*)a;  a += (np + 8) * 16;
        A.cgR = (float4*)a;  a += (np + 8) * 16;
        A.cgG = (float4*)a;  a += (np + 8) * 16;
        A.cgPyz = (float2*)a;
        l2_window(pcgArena, pcgArenaBytes);
    }
    float4** f4s[] = { &A.pos, &A.vel, &A.dv, &A.nbar, &A.pos2, &A.vel2, &A.dv2, &A.nbar2, &A.posRho, &A.acc, &A.pacc, &A.nrm,
                       &A.cgG, &A.cgR, &A.cgQ, &A.cgZ, &A.cgXG, &A.cgXP, &dPos0, &dVel0 };
    CK(dalloc(dIds0, np));
    for (float4** p : f4s) if (!*p) { CK(dalloc(*p, np)); allocBytes += np * 16; }
    float** f1s[] = { &A.curv, &A.curvS, &A.curvD, &A.curv2, &A.curvS2, &A.curvD2, &A.res, &A.rho, &A.rhoAdv, &A.kappa, &A.kappaV, &A.alpha };
    for (float** p : f1s) if (!*p) { CK(dalloc(*p, np)); allocBytes += np * 4; }
    CK(dalloc(A.minv, np * 9)); allocBytes += np * 36;
    if (!A.cgGyz) { CK(dalloc(A.cgGyz, np)); allocBytes += np * 8; }
    if (!A.cgPyz) { CK(dalloc(A.cgPyz, np)); allocBytes += np * 8; }
    uint32_t** u1s[] = { &A.id, &A.id2, &A.cnt, &A.key, &A.rank, &A.tmpIdx };
    for (uint32_t** p : u1s) { CK(dalloc(*p, np)); allocBytes += np * 4; }
    CK(dalloc(A.list16, np * ELL_SLOTS)); searchBytes = np * ELL_SLOTS * 2 + np * 4 * 4;
    CK(dalloc(A.coef, np * ELL_SLOTS));
    CK(dalloc(A.gcoef, np * ELL_SLOTS));
    allocBytes += np * ELL_SLOTS * 10;
    // search grid capacity: 64x the cells of the initial bounding box (4x per axis of head room)
    double cells0 = 1.0;
    for (int k = 0; k < 3; k++) cells0 *= std::ceil((double)(bboxMax[k] - bboxMin[k]) / info.SupportRadius) + 9.0;
    cellEstimate = (uint32_t)std::min<double>(cells0, (double)optMaxCells);
    cellCapacity = (uint32_t)std::min<double>(std::max<double>(64.0 * cells0, (double)(1u << 20)), (double)optMaxCells);
    if (dist) {
        // fixed local grid: owned tile columns + ghost columns
        DevState g; memset(&g, 0, sizeof g);
        dist_apply_grid(g);
        cellEstimate = g.nCells;
        // the slab's bounds move (re-balancing): room for the whole domain's tile columns plus the two ghost columns
        cellCapacity = (dist->gtiles[0] + 2u) * dist->gtiles[1] * dist->gtiles[2] * 64u + 64u;
    }
    CK(dalloc(A.cellCount, (size_t)cellCapacity + 4)); CK(dalloc(A.cellBegin, (size_t)cellCapacity + 4));
    CK(cudaMemset(A.cellCount, 0, ((size_t)cellCapacity + 4) * 4));
    CK(cudaMemset(A.cellBegin, 0, ((size_t)cellCapacity + 4) * 4));
    CK(dalloc(A.tileSums, (size_t)cellCapacity / 4096 + 8));
    CK(dalloc(A.tileList, (size_t)cellCapacity / TILE_CELLS + 8));
    searchBytes += ((size_t)cellCapacity * 2 + 8) * 4;
    allocBytes += ((size_t)cellCapacity * 2 + 8) * 4;
    CK(dalloc(A.partials, (size_t)4 * 65536));
    A.slotStride = (uint32_t)((size_t)cellCapacity / TILE_CELLS + 8);      // one reduction slot per listed tile
    CK(dalloc(A.slotSums, (size_t)2 * A.slotStride));
    CK(cudaMemset(A.slotSums, 0, (size_t)2 * A.slotStride * sizeof(double)));
    for (uint32_t b = 0; b < info.RigidBodyCount; b++) { CK(dalloc(A.bx[b], np)); CK(dalloc(A.bcoef[b], np)); CK(dalloc(A.bgrad[b], np)); allocBytes += np * 48; }
    return VFD_OK;
}

int Solver::set_particles(const float* pos, const float* vel, uint32_t n, bool onDevice) {
    CK(cudaSetDevice(device));
    if (!pos && n) return fail(VFD_E_INVALID, "set_particles: null positions");
    state = VFD_STATE_NONE;
    info.ParticleCount = n;
    refresh_params();
    pipe.clear();
    began = false;
    if (n == 0) { free_particles(); return VFD_OK; }
    float *dPos = nullptr, *dVel = nullptr;
    float bmin[3] = { 0, 0, 0 }, bmax[3] = { 1, 1, 1 };
    std::vector<float> hostPos;
    const float* hp = pos;
    if (onDevice) {
        hostPos.resize((size_t)3 * n);
        CK(cudaMemcpy(hostPos.data(), pos, (size_t)12 * n, cudaMemcpyDeviceToHost));
        hp = hostPos.data();
    }
    for (int k = 0; k < 3; k++) { bmin[k] = FLT_MAX; bmax[k] = -FLT_MAX; }
    for (size_t i = 0; i < n; i++) for (int k = 0; k < 3; k++) { const float v = hp[3 * i + k]; bmin[k] = std::min(bmin[k], v); bmax[k] = std::max(bmax[k], v); }
    int rc = alloc_particles(n, bmin, bmax);
    if (rc) return rc;
    CK(pipe.configure(device, n));
    if (onDevice) { dPos = const_cast<float*>(pos); dVel = const_cast<float*>(vel); }
    else {
        // the upload's staging buffers stay with the handle (see set_rigid_bodies: no cudaMalloc / cudaFree per bake)
        if (stageSlots < n) {
            if (dStagePos) cudaFree(dStagePos);
            if (dStageVel) cudaFree(dStageVel);
            dStagePos = dStageVel = nullptr; stageSlots = 0;
            CK(cudaMalloc(&dStagePos, (size_t)12 * n)); CK(cudaMalloc(&dStageVel, (size_t)12 * n));
            stageSlots = n;
        }
        dPos = dStagePos;
        CK(cudaMemcpyAsync(dPos, pos, (size_t)12 * n, cudaMemcpyHostToDevice, stream));
        if (vel) { dVel = dStageVel; CK(cudaMemcpyAsync(dVel, vel, (size_t)12 * n, cudaMemcpyHostToDevice, stream)); }
    }
    k_pack_posvel<<<nblk(n), VFD_TPB, 0, stream>>>(n, dPos, dVel, dPos0, dVel0);
    launches += 1;
    CK(cudaStreamSynchronize(stream));
    return begin();
}

// Distributed variant: this rank's particles (those inside its slab), their global ids, the global particle count and
// the number of particle slots to allocate (owned + ghosts + head room for migration).
int Solver::dist_set_particles(const float* pos, const float* vel, const uint32_t* ids, uint32_t n, uint32_t nGlobal, uint32_t capacity) {
    CK(cudaSetDevice(device));
    if (!dist) return fail(VFD_E_INVALID, "set_particles_distributed needs init_distributed first");
    if ((!pos || !ids) && n) return fail(VFD_E_INVALID, "set_particles_distributed: null positions / ids");
    Dist& D = *dist;
    D.nGlobal = nGlobal;
    D.capacity = std::max<uint32_t>(capacity, n + 1024u);
    // one tile column of ghosts per side + the migrants of a step
    const uint64_t column = (uint64_t)D.gtiles[1] * D.gtiles[2] * 64ull * 24ull;      // 24 particles per cell is twice the rest density
    D.haloCap = (uint32_t)std::min<uint64_t>(std::max<uint64_t>(column, 1u << 16), (uint64_t)D.capacity);
    cudaFree(D.sendL); cudaFree(D.sendR); cudaFree(D.recvL); cudaFree(D.recvR);
    D.sendL = D.sendR = D.recvL = D.recvR = nullptr;
    CK(cudaMalloc(&D.sendL, (size_t)D.haloCap * 80)); CK(cudaMalloc(&D.sendR, (size_t)D.haloCap * 80));
    CK(cudaMalloc(&D.recvL, (size_t)D.haloCap * 80)); CK(cudaMalloc(&D.recvR, (size_t)D.haloCap * 80));
    state = VFD_STATE_NONE;
    info.ParticleCount = n;
    pipe.clear();
    began = false;
    float bmin[3] = { 0, 0, 0 }, bmax[3] = { 1, 1, 1 };
    int rc = alloc_particles(n, bmin, bmax);
    if (rc) return rc;
    CK(pipe.configure(device, D.rank == 0 ? nGlobal : 0u));          // frames are whole-scene frames, gathered on rank 0
    // rigid-body sample arrays are sized with the particle arrays
    if (n) {
        float *dPos = nullptr, *dVel = nullptr;
        CK(cudaMalloc(&dPos, (size_t)12 * n));
        CK(cudaMemcpyAsync(dPos, pos, (size_t)12 * n, cudaMemcpyHostToDevice, stream));
        if (vel) { CK(cudaMalloc(&dVel, (size_t)12 * n)); CK(cudaMemcpyAsync(dVel, vel, (size_t)12 * n, cudaMemcpyHostToDevice, stream)); }
        CK(cudaMemcpyAsync(dIds0, ids, (size_t)4 * n, cudaMemcpyHostToDevice, stream));
        k_pack_posvel<<<nblk(n), VFD_TPB, 0, stream>>>(n, dPos, dVel, dPos0, dVel0);
        launches += 1;
        CK(cudaStreamSynchronize(stream));
        cudaFree(dPos); if (dVel) cudaFree(dVel);
    }
    return begin();
}

void Solver::free_bodies() {
    for (auto& p : bodyAllocs) cudaFree(p);
    bodyAllocs.clear();
    bodyAllocBytes.clear();
    memset(&bodies, 0, sizeof bodies);
}

int Solver::set_rigid_bodies(uint32_t count, const VfdVolumeMap* maps) {
    CK(cudaSetDevice(device));
    if (count > VFD_MAX_BODIES) return fail(VFD_E_INVALID, "too many rigid bodies (VFD_MAX_BODIES = 8)");
    if (count && !maps) return fail(VFD_E_INVALID, "set_rigid_bodies: null maps");
    CK(cudaStreamSynchronize(stream));
    state = VFD_STATE_NONE;
    // Re-baking calls this with the same bodies again and again (the editor's "Bake": SetFluidObjects + SetRigidBodies + Simulate):
    // device allocations of unchanged size are kept — cudaFree / cudaMalloc of tens of megabytes stall for up to seconds now and
    // then once gigabytes of pinned frame storage exist, which showed as a 2x spread of the end-to-end figure.
    std::vector<size_t> want;
    for (uint32_t b = 0; b < count; b++) {
        const VfdVolumeMap& m = maps[b];
        if (m.fieldCount < 2 || !m.nodes || !m.cells || !m.cellMap) return fail(VFD_E_INVALID, "volume map needs two fields (SDF, volume) and its three arrays");
        want.push_back((size_t)m.fieldCount * m.nodeCount * 4); want.push_back((size_t)m.fieldCount * m.cellCount * 32 * 4); want.push_back((size_t)m.fieldCount * m.cellMapCount * 4);
    }
    if (want != bodyAllocBytes) {
        free_bodies();
        for (size_t bytes : want) { void* p = nullptr; CK(cudaMalloc(&p, std::max<size_t>(bytes, 4))); bodyAllocs.push_back(p); }
        bodyAllocBytes = want;
    }
    memset(&bodies, 0, sizeof bodies);
    for (uint32_t b = 0; b < count; b++) {
        const VfdVolumeMap& m = maps[b];
        DevVolumeMap& d = bodies.map[b];
        for (int k = 0; k < 3; k++) { d.dmin[k] = m.domainMin[k]; d.dmax[k] = m.domainMax[k]; d.res[k] = m.resolution[k]; d.cell[k] = m.cellSize[k]; d.cellInv[k] = m.cellSizeInverse[k]; }
        d.fieldCount = m.fieldCount; d.nodeCount = m.nodeCount; d.cellCount = m.cellCount; d.cellMapCount = m.cellMapCount;
        float* dn = (float*)bodyAllocs[3 * b]; uint32_t* dc = (uint32_t*)bodyAllocs[3 * b + 1]; uint32_t* dm = (uint32_t*)bodyAllocs[3 * b + 2];
        CK(cudaMemcpy(dn, m.nodes, want[3 * b], cudaMemcpyHostToDevice));
        CK(cudaMemcpy(dc, m.cells, want[3 * b + 1], cudaMemcpyHostToDevice));
        CK(cudaMemcpy(dm, m.cellMap, want[3 * b + 2], cudaMemcpyHostToDevice));
        d.nodes = dn; d.cells = dc; d.cellMap = dm;
    }
    // per-body boundary sample arrays are sized by the particle count (RigidBody.cu:13-16): particles first
    const size_t np = ((size_t)(dist ? std::max(dist->capacity, info.ParticleCount) : info.ParticleCount) + 31) / 32 * 32;
    const bool need = info.ParticleCount || dist;
    for (int b = 0; b < VFD_MAX_BODIES; b++) {
        const bool keep = need && (uint32_t)b < count && arrays.bx[b] && bodySampleSlots == np;
        if (keep) continue;
        if (arrays.bx[b]) { cudaFree(arrays.bx[b]); arrays.bx[b] = nullptr; }
        if (arrays.bcoef[b]) { cudaFree(arrays.bcoef[b]); arrays.bcoef[b] = nullptr; }
        if (arrays.bgrad[b]) { cudaFree(arrays.bgrad[b]); arrays.bgrad[b] = nullptr; }
    }
    for (uint32_t b = 0; b < count && need; b++) {
        if (!arrays.bx[b]) { CK(dalloc(arrays.bx[b], np)); CK(dalloc(arrays.bcoef[b], np)); CK(dalloc(arrays.bgrad[b], np)); }
        else { CK(cudaMemset(arrays.bx[b], 0, np * 16)); CK(cudaMemset(arrays.bcoef[b], 0, np * 16)); CK(cudaMemset(arrays.bgrad[b], 0, np * 16)); }
    }
    bodySampleSlots = np;
    info.RigidBodyCount = count;
    refresh_params();
    return VFD_OK;
}

// What Simulate() does before its loop (DFSPHImplementation.cu:36-51)
int Solver::begin() {
    CK(cudaSetDevice(device));
    if (info.ParticleCount == 0 && !dist) { began = true; return VFD_OK; }
    if (dist) { dist->nLocal = info.ParticleCount; dist->ownB = 0; dist->ownE = info.ParticleCount; dist->edgeLEnd = 0; dist->edgeRBegin = info.ParticleCount; }
    refresh_params();
    if (params.n) {                                      // a rank of a decomposition may own nothing at start-up
        k_reset_state<<<nblk(params.n), VFD_TPB, 0, stream>>>(params, arrays, dPos0, dVel0, dist ? dIds0 : nullptr);
        launches += 1;
    }
    DevState s;
    memset(&s, 0, sizeof s);
    s.dt = desc.TimeStepSize; s.dt2 = desc.TimeStepSize * desc.TimeStepSize;
    s.dtInv = 1.0f / s.dt; s.dt2Inv = 1.0f / s.dt2;
    s.sampleCount = info.SurfaceTensionSampleCount; s.mcFactor = info.MonteCarloFactor;
    for (int k = 0; k < 3; k++) { s.boundsMin[k] = INT_MAX; s.boundsMax[k] = INT_MIN; s.gridDim[k] = 4; s.tileDim[k] = 1; }
    s.nCells = 64; s.nTiles = 1;
    if (dist) dist_apply_grid(s);
    hState[0] = s;
    CK(cudaMemcpyAsync(dState, &hState[0], sizeof(DevState), cudaMemcpyHostToDevice, stream));
    CK(cudaStreamSynchronize(stream));
    info.TimeStepSize = s.dt; info.TimeStepSize2 = s.dt2; info.TimeStepSizeInverse = s.dtInv; info.TimeStepSize2Inverse = s.dt2Inv;
    {
        std::lock_guard<std::mutex> g(dbgMutex);
        memset(&debug, 0, sizeof debug);
    }
    pipe.clear();
    frameTimeHost = 0.0f; frameIndexHost = 0; stepsIssued = 0;
    began = true;
    searched = false;
    return VFD_OK;
}

int Solver::read_state(DevState& out) {
    CK(cudaMemcpyAsync(&hState[1], dState, sizeof(DevState), cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    out = hState[1];
    cellEstimate = std::max<uint32_t>(out.nCells, 27u);
    if (out.errorFlags & 1u) return fail(VFD_E_CAPACITY, "search grid exceeds the cell capacity (raise VFD_OPT_MAX_CELLS; a particle escaped far from the fluid?)");
    if (out.errorFlags & 2u) return fail(VFD_E_CAPACITY, "more than 65535 particles in one 6x6x6-cell neighbourhood: beyond the 16-bit tile-local neighbour index");
    return VFD_OK;
}

// polls a device flag with one batch of look-ahead; returns true if the loop may stop
int Solver::run_polled_loop(uint32_t maxIt, uint32_t already, int batch, uint32_t* dFlag, const std::function<int()>& enqueueIteration, uint32_t continueValue) {
    uint32_t issued = already;
    int slot = 0, pending = -1;
    while (issued < maxIt) {
        for (int b = 0; b < batch && issued < maxIt; b++, issued++) { int rc = enqueueIteration(); if (rc) return rc; }
        CK(cudaMemcpyAsync(&hFlags[slot], dFlag, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
        CK(cudaEventRecord(pollEvent[slot], stream));
        if (pending >= 0) {
            CK(cudaEventSynchronize(pollEvent[pending]));
            if (hFlags[pending] != continueValue) return VFD_OK;
        }
        pending = slot;
        slot ^= 1;
    }
    return VFD_OK;
}

int Solver::search_only() {
    CK(cudaSetDevice(device));
    if (!began) { int rc = begin(); if (rc) return rc; }
    if (info.ParticleCount == 0) return VFD_OK;
    refresh_params();
    LaunchCfg L{ stream, numSMs, &launches, &prof };
    launch_search(L, params, arrays, dState, cellCapacity, cellEstimate);
    searched = true;
    CK(cudaGetLastError());
    return VFD_OK;
}

#define RC(call) do { int rc__ = (call); if (rc__) return rc__; } while (0)

int Solver::step() {
    CK(cudaSetDevice(device));
    if (!began) { int rc = begin(); if (rc) return rc; }
    if (info.ParticleCount == 0 && !dist) return VFD_OK;        // DFSPHImplementation.cu:65-67
    // several ranks: particles that left the slab migrate, the ghost copies of the neighbours' edge columns are renewed
    if (dist) RC(dist_exchange_state());
    refresh_params();
    const Params& P = params;
    Arrays& A = arrays;
    LaunchCfg L{ stream, numSMs, &launches, &prof };
    const bool T = optTimers;
    if (T) cudaEventRecord(phaseEvent[0], stream);

    // 1. neighbourhood search (:71)
    launch_search(L, P, A, dState, cellCapacity, cellEstimate);
    searched = true;
    if (dist) RC(dist_read_ranges());
    if (T) cudaEventRecord(phaseEvent[1], stream);

    // 2. boundary samples, density, factor (:79-104); a = g (:112) is fused into the density pass
    launch_boundary(L, P, A, bodies);
    launch_density_factor(L, P, A, dState, dLutW, dLutG);
    RC(halo4(A.posRho));
    if (T) cudaEventRecord(phaseEvent[2], stream);

    // 3. divergence-free solve (:108, :506-575)
    if (desc.EnableDivergenceSolverError) {
        launch_divergence_source(L, P, A, dState, dLutG);
        RC(halo1(A.kappaV));
        auto iteration = [&]() -> int {
            launch_divergence_accel(L, P, A, dState, dLutG);
            RC(halo4(A.pacc));
            launch_divergence_solve(L, P, A, dState, dLutG);
            RC(reduce(SITE_DIV));
            RC(halo1(A.kappaV));
            return VFD_OK;
        };
        const uint32_t fixed = std::min(P.minDivIt, P.maxDivIt);
        for (uint32_t i = 0; i < fixed; i++) RC(iteration());
        if (fixed > 0 && fixed < P.maxDivIt) RC(run_polled_loop(P.maxDivIt, fixed, 2, &dState->divActive, iteration, 1u));
        launch_divergence_finish(L, P, A, dState, dLutG);
    }
    if (T) cudaEventRecord(phaseEvent[3], stream);

    // 5. surface tension (:119, :810-839)
    if (desc.EnableSurfaceTensionSolver) {
        launch_st_classify(L, P, A, dState, dHalton);
        RC(halo4(A.nrm));
        launch_st_smooth(L, P, A, dState);
        for (uint32_t i = 0; i < desc.SurfaceTensionSmoothPassCount; i++) launch_st_apply(L, P, A);
    }
    if (T) cudaEventRecord(phaseEvent[4], stream);

    // 6. implicit viscosity (:123, :577-808)
    if (desc.EnableViscositySolver) {
        launch_viscosity_setup(L, P, A, dState, dLutG);
        RC(reduce(SITE_VISC_BB));
        RC(halo42(A.cgXG, A.cgGyz));
        launch_viscosity_matvec(L, P, A, dState, true);
        RC(reduce(SITE_VISC_INIT));
        RC(halo42(A.cgXP, A.cgPyz));
        const bool fusedStep = viscosity_step_fits(L, P);
        // several ranks over peer memory: the fused vector kernel writes the new direction of the edge columns straight into
        // the neighbours' ghost ranges, and the next mat-vec waits for their "landed" flags: no exchange kernel per iteration
        // (VFD_DIST_FUSED_HALO=1; every rank takes the same decision: from the largest owned count over all ranks)
        const bool fusedHalo = fusedStep && dist && dist->fusedNow && P.nRanks > 1u && P.peerCtl[0] != nullptr;
        bool haloPending = false;
        auto iteration = [&]() -> int {
            launch_viscosity_matvec(L, P, A, dState, false, haloPending);
            RC(reduce(SITE_VISC_PQ));
            if (fusedStep) {
                CK((cudaError_t)launch_viscosity_step(L, P, A, dState));
            } else {
                launch_viscosity_update(L, P, A, dState);
                RC(reduce(SITE_VISC_UPDATE));
                launch_viscosity_direction(L, P, A, dState);
            }
            if (fusedHalo) { haloPending = true; dist_count_fused_halo(24); }
            else RC(halo42(A.cgXP, A.cgPyz));
            return VFD_OK;
        };
        if (P.minViscIt == 0 && P.maxViscIt > 0) RC(run_polled_loop(P.maxViscIt, 0, 4, &dState->viscActive, iteration, 1u));
        launch_viscosity_apply(L, P, A, dState);
    }
    if (T) cudaEventRecord(phaseEvent[5], stream);

    // 7.-8. CFL time step, v += dt a (:127-130)
    launch_cfl(L, P, A, dState);
    RC(reduce(SITE_CFL, true));
    launch_velocity(L, P, A, dState);
    RC(halo4(A.vel));

    // 9. constant-density solve (:137, :443-504)
    launch_pressure_source(L, P, A, dState, dLutG);
    RC(halo1(A.kappa));
    {
        auto iteration = [&]() -> int {
            launch_pressure_accel(L, P, A, dState, dLutG);
            RC(halo4(A.pacc));
            launch_pressure_solve(L, P, A, dState, dLutG);
            RC(reduce(SITE_PRESS));
            RC(halo1(A.kappa));
            return VFD_OK;
        };
        const uint32_t fixed = std::min(P.minPressIt, P.maxPressIt);
        for (uint32_t i = 0; i < fixed; i++) RC(iteration());
        if (fixed > 0 && fixed < P.maxPressIt) RC(run_polled_loop(P.maxPressIt, fixed, 2, &dState->pressActive, iteration, 1u));
        launch_pressure_finish(L, P, A, dState, dLutG);
    }
    if (T) cudaEventRecord(phaseEvent[6], stream);

    // 10. x += dt v (:141)
    launch_positions(L, P, A, dState);
    CK(cudaGetLastError());
    stepsIssued++;
    if (prof.enabled) { CK(cudaStreamSynchronize(stream)); prof.drain(); }

    // 11. frame capture (:148-167).  FrameLength <= 0: every step is a frame, so there is nothing to ask the device —
    // the frame's scalars travel with it and the host keeps queueing the next step (debug info is refreshed when the
    // bake ends).  Otherwise the accumulated frame time is needed on the host.
    // A rank of a decomposition holds a slab in local order: its frames go through dist_capture_frame (owned particles with
    // their persistent ids, gathered on rank 0), never through the original-order export of the whole scene.
    if (dist) return dist_frame_step();
    if (frameIndexHost < desc.FrameCount && desc.FrameLength <= 0.0f && !T && state == VFD_STATE_SIMULATING) {
        VfdParticleSimple* d = pipe.acquire();
        if (!d) { cudaGetLastError(); return fail(VFD_E_CUDA, "frame capture: cannot allocate the frame ring buffers"); }
        k_export_frame<<<nblk(params.n), VFD_TPB, 0, stream>>>(params, arrays, d, dState, pipe.meta_slot(), 0);
        launches += 1;
        CK(pipe.submit(stream, 0.0f, 0.0f, true));
        frameTimeHost = 0.0f;
        frameIndexHost++;
    } else if (asyncFrames) {
        // Simulate() with a frame length (the reference's default, 0.0016 s): whether this step's state is a frame depends on the
        // time steps the device chose.  The device decides (k_frame_decide), the export and the copy to the host are enqueued
        // for every step, and the frame pipe's worker keeps the frame or drops it: no host round trip per step.
        VfdParticleSimple* d = pipe.acquire();
        if (!d) { cudaGetLastError(); return fail(VFD_E_CUDA, "frame capture: cannot allocate the frame ring buffers"); }
        k_frame_decide<<<1, 32, 0, stream>>>(dState, desc.FrameLength, desc.FrameCount);
        k_export_frame<<<nblk(params.n), VFD_TPB, 0, stream>>>(params, arrays, d, dState, pipe.meta_slot(), 1);
        launches += 2;
        CK(pipe.submit(stream, 0.0f, 0.0f, true, true));
    } else if (frameIndexHost < desc.FrameCount || T) {
        DevState s;
        int rc = read_state(s);
        if (rc) return rc;
        frameTimeHost += s.dt;                  // m_DebugInfo.FrameTime += TimeStepSize (:427)
        update_debug(s, T);
        if (frameIndexHost < desc.FrameCount && frameTimeHost >= desc.FrameLength) {
            rc = capture_frame(s);
            if (rc) return rc;
        }
    }
    return VFD_OK;
}

void Solver::update_debug(const DevState& s, bool timers) {
    std::lock_guard<std::mutex> g(dbgMutex);
    debug.IterationCount = s.stepCount;
    debug.DivergenceSolverIterationCount = desc.EnableDivergenceSolverError ? s.divIt : 0;
    debug.PressureSolverIterationCount = s.pressIt;
    debug.ViscositySolverIterationCount = desc.EnableViscositySolver ? s.viscIt : 0;
    debug.DivergenceSolverError = desc.EnableDivergenceSolverError ? s.divErr : 0.0f;
    debug.PressureSolverError = s.pressErr;
    debug.ViscositySolverError = desc.EnableViscositySolver ? s.viscErr : 0.0f;
    debug.FrameTime = frameTimeHost;
    debug.FrameIndex = frameIndexHost;
    info.TimeStepSize = s.dt; info.TimeStepSize2 = s.dt2; info.TimeStepSizeInverse = s.dtInv; info.TimeStepSize2Inverse = s.dt2Inv;
    info.SurfaceTensionSampleCount = s.sampleCount; info.MonteCarloFactor = s.mcFactor;
    maxVel2 = s.vmax2;
    if (timers) {
        float ms[6];
        for (int i = 0; i < 6; i++) { ms[i] = 0.0f; cudaEventElapsedTime(&ms[i], phaseEvent[i], phaseEvent[i + 1]); }
        debug.NeighborhoodSearchUs = ms[0] * 1000.0f; debug.BaseSolverUs = ms[1] * 1000.0f; debug.DivergenceSolverUs = ms[2] * 1000.0f;
        debug.SurfaceTensionSolverUs = ms[3] * 1000.0f; debug.ViscositySolverUs = ms[4] * 1000.0f; debug.PressureSolverUs = ms[5] * 1000.0f;
    }
}

int Solver::capture_frame(const DevState& s) {
    // K15 + the frame cache: export in original particle order on the solver's stream, then hand the buffer to the
    // asynchronous pipe (copy stream + host worker): the next step does not wait for PCIe or for the host copy.
    VfdParticleSimple* d = pipe.acquire();
    if (!d) { cudaGetLastError(); return fail(VFD_E_CUDA, "frame capture: cannot allocate the frame ring buffers"); }
    k_export_frame<<<nblk(params.n), VFD_TPB, 0, stream>>>(params, arrays, d, dState, nullptr, 0);
    launches += 1;
    CK(pipe.submit(stream, s.vmax2, s.dt));
    frameTimeHost = 0.0f;              // FrameTime = 0 (:164)
    frameIndexHost++;
    std::lock_guard<std::mutex> g(dbgMutex);
    debug.FrameTime = 0.0f; debug.FrameIndex = frameIndexHost;
    return VFD_OK;
}

int Solver::simulate() {
    int rc = begin();
    if (rc) return rc;
    state = VFD_STATE_SIMULATING;
    // frame length > 0 on one GPU: the device decides which steps are frames (see step()); the host only has to stop in time —
    // it issues a step only if the bake would be incomplete even if every step still in flight turned out to be a frame
    asyncFrames = desc.FrameLength > 0.0f && !optTimers && !dist && info.ParticleCount != 0 && getenv("VFD_SYNC_FRAMES") == nullptr;
    if (asyncFrames) {
        for (;;) {
            size_t pub = 0, pend = 0;
            pipe.progress(pub, pend);
            if (pub >= desc.FrameCount) break;
            if (pub + pend >= desc.FrameCount) { pipe.wait_pending_below(pend); continue; }
            rc = step();
            if (rc) { state = VFD_STATE_NONE; asyncFrames = false; return rc; }
        }
        asyncFrames = false;
        frameIndexHost = desc.FrameCount;
    }
    while (frameIndexHost < desc.FrameCount) {
        if (info.ParticleCount == 0) break;    // the reference would spin forever here (OnUpdate returns at once)
        rc = step();
        if (rc) { state = VFD_STATE_NONE; return rc; }
    }
    CK(pipe.drain());                  // every baked frame is in the host cache when Simulate() returns
    rc = sync_debug();                 // also surfaces device-side error flags of the steps that ran unobserved
    if (rc) { state = VFD_STATE_NONE; return rc; }
    {
        std::lock_guard<std::mutex> g(dbgMutex);
        debug.FrameIndex = frameIndexHost;
    }
    state = VFD_STATE_READY;
    return VFD_OK;
}

int Solver::synchronize() {
    CK(cudaSetDevice(device));
    CK(cudaStreamSynchronize(stream));
    CK(pipe.drain());
    return VFD_OK;
}

int Solver::sync_debug() {
    CK(cudaSetDevice(device));
    if (!began || info.ParticleCount == 0) return VFD_OK;
    DevState s;
    int rc = read_state(s);
    if (rc) return rc;
    update_debug(s, false);
    return VFD_OK;
}

int Solver::get_particles(VfdParticle* out) {
    CK(cudaSetDevice(device));
    if (dist) return fail(VFD_E_INVALID, "original-order dumps are single-GPU calls: use get_owned on each rank");
    if (!out) return fail(VFD_E_INVALID, "null output");
    if (info.ParticleCount == 0) return VFD_OK;
    VfdParticle* d = nullptr;
    CK(cudaMalloc(&d, (size_t)info.ParticleCount * sizeof(VfdParticle)));
    refresh_params();
    k_export_aos<<<nblk(params.n), VFD_TPB, 0, stream>>>(params, arrays, d);
    launches += 1;
    cudaError_t e = cudaMemcpyAsync(out, d, (size_t)info.ParticleCount * sizeof(VfdParticle), cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    cudaFree(d);
    if (e != cudaSuccess) return fail_cuda(e, "get_particles", __LINE__);
    return VFD_OK;
}

int Solver::set_particles_full(const VfdParticle* in) {
    CK(cudaSetDevice(device));
    if (dist) return fail(VFD_E_INVALID, "original-order dumps are single-GPU calls: use get_owned on each rank");
    if (!in) return fail(VFD_E_INVALID, "null input");
    if (!began) { int rc = begin(); if (rc) return rc; }
    if (info.ParticleCount == 0) return VFD_OK;
    VfdParticle* d = nullptr;
    CK(cudaMalloc(&d, (size_t)info.ParticleCount * sizeof(VfdParticle)));
    cudaError_t e = cudaMemcpyAsync(d, in, (size_t)info.ParticleCount * sizeof(VfdParticle), cudaMemcpyHostToDevice, stream);
    refresh_params();
    if (e == cudaSuccess) { k_import_aos<<<nblk(params.n), VFD_TPB, 0, stream>>>(params, arrays, d); launches += 1; e = cudaStreamSynchronize(stream); }
    cudaFree(d);
    if (e != cudaSuccess) return fail_cuda(e, "set_particles_full", __LINE__);
    searched = false;
    return VFD_OK;
}

int Solver::set_time_step(float dt) {
    CK(cudaSetDevice(device));
    if (!began) { int rc = begin(); if (rc) return rc; }
    float v[4] = { dt, dt * dt, 1.0f / dt, 1.0f / (dt * dt) };
    CK(cudaMemcpyAsync(&dState->dt, v, sizeof v, cudaMemcpyHostToDevice, stream));
    CK(cudaStreamSynchronize(stream));
    info.TimeStepSize = v[0]; info.TimeStepSize2 = v[1]; info.TimeStepSizeInverse = v[2]; info.TimeStepSize2Inverse = v[3];
    return VFD_OK;
}

int Solver::set_st_state(uint32_t sampleCount, float mcFactor) {
    CK(cudaSetDevice(device));
    if (!began) { int rc = begin(); if (rc) return rc; }
    CK(cudaMemcpyAsync(&dState->sampleCount, &sampleCount, 4, cudaMemcpyHostToDevice, stream));
    CK(cudaMemcpyAsync(&dState->mcFactor, &mcFactor, 4, cudaMemcpyHostToDevice, stream));
    CK(cudaStreamSynchronize(stream));
    info.SurfaceTensionSampleCount = sampleCount; info.MonteCarloFactor = mcFactor;
    return VFD_OK;
}

int Solver::get_current_frame(VfdParticleSimple* out) {
    CK(cudaSetDevice(device));
    if (dist) return fail(VFD_E_INVALID, "original-order dumps are single-GPU calls: use get_owned on each rank");
    if (!out) return fail(VFD_E_INVALID, "null output");
    if (info.ParticleCount == 0) return VFD_OK;
    if (dFrameCapacity < info.ParticleCount) {           // the handle may have been given more particles since the last call
        if (dFrame) { cudaFree(dFrame); dFrame = nullptr; dFrameCapacity = 0; }
        CK(cudaMalloc(&dFrame, (size_t)info.ParticleCount * sizeof(VfdParticleSimple)));
        dFrameCapacity = info.ParticleCount;
    }
    refresh_params();
    k_export_frame<<<nblk(params.n), VFD_TPB, 0, stream>>>(params, arrays, dFrame, dState, nullptr, 0);
    launches += 1;
    CK(cudaMemcpyAsync(out, dFrame, (size_t)info.ParticleCount * sizeof(VfdParticleSimple), cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    return VFD_OK;
}

int Solver::get_neighbors(uint32_t* counts, uint32_t* offsets, uint32_t* ids, uint64_t capacity, uint64_t* total) {
    CK(cudaSetDevice(device));
    if (dist) return fail(VFD_E_INVALID, "original-order dumps are single-GPU calls: use get_owned on each rank");
    const uint32_t n = info.ParticleCount;
    if (!searched) return fail(VFD_E_INVALID, "get_neighbors: no neighbour search has run on the current state");
    uint32_t *dC = nullptr, *dI = nullptr;
    CK(cudaMalloc(&dC, (size_t)n * 4));
    CK(cudaMalloc(&dI, (size_t)n * VFD_MAX_NEIGHBORS * 4));
    refresh_params();
    k_export_neighbors<<<numSMs * 2, 512, 0, stream>>>(arrays, dState, dC, dI);
    launches += 1;
    std::vector<uint32_t> c(n), padded((size_t)n * VFD_MAX_NEIGHBORS);
    cudaError_t e = cudaMemcpyAsync(c.data(), dC, (size_t)n * 4, cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(padded.data(), dI, padded.size() * 4, cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    cudaFree(dC); cudaFree(dI);
    if (e != cudaSuccess) return fail_cuda(e, "get_neighbors", __LINE__);
    uint64_t tot = 0;
    for (uint32_t i = 0; i < n; i++) { if (counts) counts[i] = c[i]; if (offsets) offsets[i] = (uint32_t)tot; tot += c[i]; }
    if (total) *total = tot;
    if (ids) {
        if (capacity < tot) return fail(VFD_E_INVALID, "get_neighbors: ids capacity too small");
        uint64_t o = 0;
        for (uint32_t i = 0; i < n; i++) {
            uint32_t* row = padded.data() + (size_t)i * VFD_MAX_NEIGHBORS;
            std::sort(row, row + c[i]);
            memcpy(ids + o, row, (size_t)c[i] * 4);
            o += c[i];
        }
    }
    return VFD_OK;
}

int Solver::get_boundary(uint32_t body, float* xj, float* vol) {
    CK(cudaSetDevice(device));
    if (dist) return fail(VFD_E_INVALID, "original-order dumps are single-GPU calls: use get_owned on each rank");
    if (body >= info.RigidBodyCount) return fail(VFD_E_INVALID, "body index out of range");
    const uint32_t n = info.ParticleCount;
    float *dX = nullptr, *dV = nullptr;
    CK(cudaMalloc(&dX, (size_t)n * 12)); CK(cudaMalloc(&dV, (size_t)n * 4));
    refresh_params();
    k_export_boundary<<<nblk(n), VFD_TPB, 0, stream>>>(params, arrays, body, dX, dV);
    launches += 1;
    cudaError_t e = cudaMemcpyAsync(xj, dX, (size_t)n * 12, cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(vol, dV, (size_t)n * 4, cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    cudaFree(dX); cudaFree(dV);
    if (e != cudaSuccess) return fail_cuda(e, "get_boundary", __LINE__);
    return VFD_OK;
}

int Solver::record_event(uint32_t slot) {
    CK(cudaSetDevice(device));
    if (slot >= 16) return fail(VFD_E_INVALID, "event slot out of range");
    if (!userEvent[slot]) CK(cudaEventCreate(&userEvent[slot]));
    CK(cudaEventRecord(userEvent[slot], stream));
    return VFD_OK;
}

int Solver::elapsed_ms(uint32_t from, uint32_t to, float* ms) {
    CK(cudaSetDevice(device));
    if (from >= 16 || to >= 16 || !userEvent[from] || !userEvent[to]) return fail(VFD_E_INVALID, "event slot not recorded");
    CK(cudaEventSynchronize(userEvent[to]));
    CK(cudaEventElapsedTime(ms, userEvent[from], userEvent[to]));
    return VFD_OK;
}

int Solver::tile_stats(uint64_t* st) {
    st[0] = st[1] = st[2] = 0; st[3] = pipe.bytesCopied;
    if (!began || info.ParticleCount == 0) return VFD_OK;
    DevState s;
    CK(cudaSetDevice(device));
    int rc = read_state(s);
    if (rc) return rc;
    st[0] = s.nTiles; st[1] = s.nCells; st[2] = s.fallbackTiles;
    return VFD_OK;
}

// Tuning aid: `reps` back-to-back launches of the initial PCG mat-vec (r = b - A g: the pipelined pair pass with the
// heaviest payload) on the state the last step left behind, timed with events on the solver's stream.  The PCG work
// arrays it overwrites are rebuilt by the next step.
int Solver::time_matvec(uint32_t reps, float* ms) {
    CK(cudaSetDevice(device));
    if (!began || !searched || info.ParticleCount == 0) return fail(VFD_E_INVALID, "time_matvec: step the simulation first");
    LaunchCfg L{ stream, numSMs, &launches, &prof };
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    // the PCG's own product q = A p (VFD_TIME_MATVEC_INIT: the start-up product r = b - A g); it only runs while the solver is active
    const bool init = getenv("VFD_TIME_MATVEC_INIT") != nullptr;
    uint32_t active = 1u, saved = 0u;
    CK(cudaMemcpyAsync(&saved, &dState->viscActive, 4, cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    if (!init) CK(cudaMemcpyAsync(&dState->viscActive, &active, 4, cudaMemcpyHostToDevice, stream));
    launch_viscosity_matvec(L, params, arrays, dState, init);
    CK(cudaEventRecord(a, stream));
    for (uint32_t i = 0; i < reps; i++) launch_viscosity_matvec(L, params, arrays, dState, init);
    CK(cudaEventRecord(b, stream));
    CK(cudaMemcpyAsync(&dState->viscActive, &saved, 4, cudaMemcpyHostToDevice, stream));
    CK(cudaEventSynchronize(b));
    CK(cudaEventElapsedTime(ms, a, b));
    cudaEventDestroy(a); cudaEventDestroy(b);
#ifdef PIPE_TRACE
    { const char* path = getenv("VFD_TRACE_FILE"); if (path) trace_viscosity_matvec(L, params, arrays, dState, path); }
#endif
    CK(cudaStreamSynchronize(stream));
    CK(cudaGetLastError());
    return VFD_OK;
}

int Solver::get_bounds(float* bmin, float* bmax) {
    DevState s;
    CK(cudaSetDevice(device));
    int rc = read_state(s);
    if (rc) return rc;
    // ParticleSearch::ComputeMinMax (ParticleSearch.cu:50-54): min cell * h, (max cell + 1) * h
    for (int k = 0; k < 3; k++) {
        const int lo = s.gridMinCell[k], hi = lo + (int)s.gridDim[k] - 5;
        bmin[k] = (float)lo * info.SupportRadius;
        bmax[k] = (float)(hi + 1) * info.SupportRadius;
    }
    return VFD_OK;
}

} // namespace vfd
